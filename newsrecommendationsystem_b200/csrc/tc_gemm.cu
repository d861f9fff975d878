// Generic TF32 tensor-core GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T (+ bias).
//
//   CTA tile 128 x 256, K streamed in 32-float (128-byte) chunks through a 2-stage
//   shared-memory ring in the canonical UMMA SWIZZLE_128B K-major layout.
//   warps 0-3 : producers (ld.global.v4 -> cvt.rna.tf32 -> swizzled st.shared -> proxy fence ->
//               mbarrier arrive), then the epilogue (tcgen05.ld -> +bias -> st.global)
//   warp  4   : TMEM allocation + single-thread tcgen05.mma issue, tcgen05.commit to free stages
//   Accumulator: 128 lanes x 256 fp32 columns of TMEM.  Two CTAs are resident per SM (2 x 96 KB
//   smem, 2 x 256 TMEM columns) so one CTA's epilogue overlaps the other's main loop.
#include "tc_common.cuh"
#include "tc_api.cuh"

namespace nrms {
using namespace tc;

constexpr int TG_BM = 128, TG_BN = 256, TG_BK = 32, TG_STAGES = 2;
constexpr int TG_A_BYTES = TG_BM * TG_BK * 4;  // 16 KB
constexpr int TG_B_BYTES = TG_BN * TG_BK * 4;  // 32 KB
constexpr int TG_STAGE_BYTES = TG_A_BYTES + TG_B_BYTES;
constexpr int TG_THREADS = 160;
constexpr int TG_SMEM = TG_STAGES * TG_STAGE_BYTES + 1024 /*align slack*/ + 64 /*barriers*/;

__global__ void __launch_bounds__(TG_THREADS)
tc_gemm_nt_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
                  const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int64_t M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;          // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + TG_STAGES * TG_STAGE_BYTES;  // full[2], empty[2], tmem_full, tmem_ptr
  const uint32_t full_bar = bars, empty_bar = bars + 8 * TG_STAGES, tmem_full_bar = bars + 16 * TG_STAGES;
  volatile uint32_t* tmem_ptr_smem =
      reinterpret_cast<volatile uint32_t*>(base_ptr + TG_STAGES * TG_STAGE_BYTES + 16 * TG_STAGES + 8);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * TG_BM;
  const int n0 = blockIdx.y * TG_BN;
  int n_valid = N - n0;
  if (n_valid > TG_BN) n_valid = TG_BN;
  const int n_mma = (n_valid + 15) & ~15;  // UMMA N (multiple of 16 for M=128); padded B rows are zero
  const int num_chunks = (K + TG_BK - 1) / TG_BK;

  if (tid == 0) {
    for (int s = 0; s < TG_STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 128);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(smem_u32((const void*)tmem_ptr_smem), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < 4) {
    // ------------------------------ producers ------------------------------------------
    for (int kc = 0; kc < num_chunks; ++kc) {
      const int s = kc % TG_STAGES;
      const uint32_t ph = (kc / TG_STAGES) & 1;
      mbar_wait(empty_bar + 8 * s, ph ^ 1);
      uint8_t* sa = base_ptr + s * TG_STAGE_BYTES;
      uint8_t* sb = sa + TG_A_BYTES;
      const int k0 = kc * TG_BK;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = tid + 128 * i;
        const int r = idx >> 3, c = idx & 7;
        const int64_t m = m0 + r;
        const int k = k0 + 4 * c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < M && k < K) v = to_tf32(__ldg(reinterpret_cast<const float4*>(A + m * lda + k)));
        *reinterpret_cast<float4*>(sa + r * 128 + ((c ^ (r & 7)) << 4)) = v;
      }
      for (int idx = tid; idx < n_mma * 8; idx += 128) {
        const int r = idx >> 3, c = idx & 7;
        const int k = k0 + 4 * c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n_valid && k < K) v = to_tf32(__ldg(reinterpret_cast<const float4*>(B + (int64_t)(n0 + r) * ldb + k)));
        *reinterpret_cast<float4*>(sb + r * 128 + ((c ^ (r & 7)) << 4)) = v;
      }
      fence_proxy_async_smem();
      mbar_arrive(full_bar + 8 * s);
    }
    // ------------------------------ epilogue -------------------------------------------
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int64_t m = m0 + tid;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int col = 0; col < n_mma; col += 16) {
      float v[16];
      tmem_ld16(trow + col, v);
      if (m < M) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = n0 + col + 4 * q;
          if (n < N) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias) b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
            *reinterpret_cast<float4*>(C + m * ldc + n) =
                make_float4(v[4 * q] + b4.x, v[4 * q + 1] + b4.y, v[4 * q + 2] + b4.z, v[4 * q + 3] + b4.w);
          }
        }
      }
    }
  } else {
    // ------------------------------ MMA issuer (warp 4) --------------------------------
    const uint32_t idesc = umma_idesc_tf32(TG_BM, n_mma);
    for (int kc = 0; kc < num_chunks; ++kc) {
      const int s = kc % TG_STAGES;
      const uint32_t ph = (kc / TG_STAGES) & 1;
      mbar_wait(full_bar + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = base + s * TG_STAGE_BYTES;
        const uint32_t sb = sa + TG_A_BYTES;
        int ksteps = (K - kc * TG_BK + 7) / 8;
        if (ksteps > 4) ksteps = 4;
        for (int ks = 0; ks < ksteps; ++ks) {
          umma_tf32_ss(tmem_base, umma_desc_k_sw128(sa + ks * 32), umma_desc_k_sw128(sb + ks * 32), idesc,
                       (kc | ks) ? 1u : 0u);
        }
        umma_commit(empty_bar + 8 * s);                       // stage reusable once these MMAs retire
        if (kc == num_chunks - 1) umma_commit(tmem_full_bar);  // accumulator complete
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 256);
}

int tc_gemm_nt(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, float* C, int64_t ldc,
               int64_t M, int N, int K, cudaStream_t st) {
  if (M <= 0) return NRMS_OK;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_gemm_nt_kernel)");
    configured = true;
  }
  dim3 grid((unsigned)((M + TG_BM - 1) / TG_BM), (N + TG_BN - 1) / TG_BN);
  tc_gemm_nt_kernel<<<grid, TG_THREADS, TG_SMEM, st>>>(A, lda, B, ldb, bias, C, ldc, M, N, K);
  NRMS_LAUNCH_CHECK("tc_gemm_nt");
  return NRMS_OK;
}

}  // namespace nrms
