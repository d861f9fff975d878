// FP32 CUDA-core tile GEMM used by the reference-exact (NRMS_MODE_FP32) path and by the
// backward contractions.  128x64x16 CTA tile, 256 threads, 8x4 register tile per thread,
// register-prefetched double buffering.  All leading dimensions / K / N must be multiples
// of 4 floats and all base pointers 16-byte aligned (true for every NRMS shape: 300/900/200).
#pragma once
#include "common.cuh"

namespace nrms {

constexpr int SG_BM = 128, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

enum { EPI_STORE = 0, EPI_ACCUM = 1, EPI_ATOMIC = 2 };

// AT == 0: A is [M,K] row-major (K contiguous)      AT == 1: A is [K,M] row-major (M contiguous)
// BT == 0: B is [N,K] row-major (K contiguous)      BT == 1: B is [K,N] row-major (N contiguous)
// C[M,N] (row-major) = / += A*B (+ bias[N]).  blockIdx.z splits K in chunks of k_chunk (EPI_ATOMIC).
template <int AT, int BT, int EPI>
__global__ void __launch_bounds__(SG_THREADS)
sgemm_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
             const float* __restrict__ bias, float* __restrict__ C, int64_t ldc,
             int64_t M, int N, int64_t K, int64_t k_chunk) {
  __shared__ __align__(16) float As[2][SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Bs[2][SG_BK][SG_BN + 4];

  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * SG_BM;
  const int n0 = blockIdx.y * SG_BN;
  const int64_t k_begin = (int64_t)blockIdx.z * k_chunk;
  const int64_t k_end = (k_begin + k_chunk < K) ? (k_begin + k_chunk) : K;
  if (k_begin >= k_end) return;

  float4 ra[2], rb;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  auto load_tiles = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int f = tid + i * SG_THREADS;
      if (AT == 0) {
        int row = f >> 2, kq = f & 3;
        int64_t m = m0 + row, k = k0 + kq * 4;
        ra[i] = (m < M && k < k_end) ? *reinterpret_cast<const float4*>(A + m * lda + k) : zero4;
      } else {
        int krow = f >> 5, mq = f & 31;
        int64_t k = k0 + krow, m = m0 + mq * 4;
        ra[i] = (k < k_end && m < M) ? *reinterpret_cast<const float4*>(A + k * lda + m) : zero4;
      }
    }
    if (BT == 0) {
      int row = tid >> 2, kq = tid & 3;
      int n = n0 + row;
      int64_t k = k0 + kq * 4;
      rb = (n < N && k < k_end) ? *reinterpret_cast<const float4*>(B + (int64_t)n * ldb + k) : zero4;
    } else {
      int krow = tid >> 4, nq = tid & 15;
      int64_t k = k0 + krow;
      int n = n0 + nq * 4;
      rb = (k < k_end && n < N) ? *reinterpret_cast<const float4*>(B + k * ldb + n) : zero4;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int f = tid + i * SG_THREADS;
      if (AT == 0) {
        int row = f >> 2, kq = f & 3;
        As[buf][kq * 4 + 0][row] = ra[i].x;
        As[buf][kq * 4 + 1][row] = ra[i].y;
        As[buf][kq * 4 + 2][row] = ra[i].z;
        As[buf][kq * 4 + 3][row] = ra[i].w;
      } else {
        int krow = f >> 5, mq = f & 31;
        *reinterpret_cast<float4*>(&As[buf][krow][mq * 4]) = ra[i];
      }
    }
    if (BT == 0) {
      int row = tid >> 2, kq = tid & 3;
      Bs[buf][kq * 4 + 0][row] = rb.x;
      Bs[buf][kq * 4 + 1][row] = rb.y;
      Bs[buf][kq * 4 + 2][row] = rb.z;
      Bs[buf][kq * 4 + 3][row] = rb.w;
    } else {
      int krow = tid >> 4, nq = tid & 15;
      *reinterpret_cast<float4*>(&Bs[buf][krow][nq * 4]) = rb;
    }
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  load_tiles(k_begin);
  store_tiles(0);
  __syncthreads();
  int buf = 0;
  for (int64_t k0 = k_begin; k0 < k_end; k0 += SG_BK) {
    const bool has_next = (k0 + SG_BK) < k_end;
    if (has_next) load_tiles(k0 + SG_BK);
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (has_next) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  const int n = n0 + tx * 4;
  if (n >= N) return;
  float4 bv = zero4;
  if (EPI == EPI_STORE && bias != nullptr) bv = *reinterpret_cast<const float4*>(bias + n);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t m = m0 + ty * 8 + i;
    if (m >= M) break;
    float* cp = C + m * ldc + n;
    if (EPI == EPI_STORE) {
      *reinterpret_cast<float4*>(cp) =
          make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
    } else if (EPI == EPI_ACCUM) {
      float4 c = *reinterpret_cast<float4*>(cp);
      *reinterpret_cast<float4*>(cp) =
          make_float4(c.x + acc[i][0], c.y + acc[i][1], c.z + acc[i][2], c.w + acc[i][3]);
    } else {
      atomicAdd(reinterpret_cast<float4*>(cp), make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    }
  }
}

template <int AT, int BT, int EPI>
inline cudaError_t sgemm_launch(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias,
                                float* C, int64_t ldc, int64_t M, int N, int64_t K, int splits,
                                cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return cudaSuccess;
  int64_t k_chunk = K;
  if (splits > 1) {
    k_chunk = ((K + splits - 1) / splits + SG_BK - 1) / SG_BK * SG_BK;
    splits = (int)((K + k_chunk - 1) / k_chunk);
  } else {
    splits = 1;
  }
  dim3 grid((unsigned)((M + SG_BM - 1) / SG_BM), (N + SG_BN - 1) / SG_BN, splits);
  sgemm_kernel<AT, BT, EPI><<<grid, SG_THREADS, 0, st>>>(A, lda, B, ldb, bias, C, ldc, M, N, K, k_chunk);
  count_launch();
  return cudaGetLastError();
}

}  // namespace nrms
