// K2 v2 -- additive_pool_f16_kernel: additive-attention pooling over the fp16 context rows written by K1 v2.
//
//   T = C W_a^T on tcgen05 (kind::f16, fp32 accumulate, 128 x 208 x 16) with W_a (fp16, 128 KB) RESIDENT in
//   shared memory for the whole kernel -- only the context rows stream (4-stage TMA ring) -- then
//   s_i = tanh(T_i + b_a) . q_a, stable softmax over the sequence, out = sum_i w_i C_i.
//   warp 0 TMA, warp 1 MMA, warps 2-5 "score" (TMEM -> tanh.q -> softmax weights), warps 6-13 "pool"
//   (weighted row sum of the previous tile, read back from L2), double-buffered through mbarriers.
#include <cuda.h>
#include <cuda_fp16.h>
#include "tc_common.cuh"
#include "tc_api.cuh"

namespace nrms {
using namespace tc;

int make_tmap_k_major_f16(CUtensorMap* out, const void* base, int64_t rows, int cols, int64_t ld, int box_rows);

namespace k2v2 {

constexpr int CP = 320;                    // pitch (halfs) of the fp16 context rows and of the W_a copy
constexpr int KCH = 5;                     // 64-half K chunks
constexpr int BN = 208;                    // UMMA N (200 real)
constexpr int B_CHUNK = BN * 128;          // 26,624
constexpr int NSTA = 4;
constexpr int A_SLOT = 16384;
constexpr int OFF_B = 0;
constexpr int OFF_A = KCH * B_CHUNK;       // 133,120
constexpr int OFF_MISC = OFF_A + NSTA * A_SLOT;   // 198,656: sc[2][128] wv[2][128] ba[208] qa[208] | barriers
constexpr int SMEM = OFF_MISC + 4096 + 1024;
constexpr int POOL_THREADS = 256;          // 8 pool warps: one output float4 per thread and tile for S = 50
constexpr int THREADS = 192 + POOL_THREADS;

__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes)
               : "memory");
}
// One MUFU operation per element (tanh.approx.f32, max relative error 2^-11 -- the precision of the fp16 operands
// of the GEMM that feeds it) instead of ex2 + rcp: the 200 tanh per context row make the score warps MUFU-bound
// (16 operations per clock per SM, profiles/tmem_ld_probe.cu).
__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int S, int SPT>
__global__ void __launch_bounds__(THREADS, 1)
additive_pool_f16_kernel(const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_wa,
                         const __half* __restrict__ C, const float* __restrict__ ba, const float* __restrict__ qa,
                         float* __restrict__ out, int64_t n_seq, uint32_t a_tx_bytes) {
  constexpr int ROWS = S * SPT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  float* sc = reinterpret_cast<float*>(sm + OFF_MISC);     // [2][128]
  float* wv = sc + 256;                                    // [2][128]
  float* ba_s = wv + 256;
  float* qa_s = ba_s + 208;
  const uint32_t bars = base + OFF_MISC + 3840;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * NSTA, tfull_bar = bars + 16 * NSTA, tempty_bar = tfull_bar + 16;
  const uint32_t b_full = tfull_bar + 32, wv_ready = tfull_bar + 40, wv_free = tfull_bar + 56;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(sm + OFF_MISC + 3840 + 16 * NSTA + 80);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler (converged MMA issue)
  const int64_t n_tiles = (n_seq + SPT - 1) / SPT;

  if (tid == 0) {
    for (int s = 0; s < NSTA; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);
      mbar_init(tempty_bar + 8 * a, 4);
      mbar_init(wv_ready + 8 * a, 128);
      mbar_init(wv_free + 8 * a, POOL_THREADS);
    }
    mbar_init(b_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_ptr_smem), 512);
  for (int i = tid; i < 208; i += THREADS) {
    ba_s[i] = i < QD ? ba[i] : 0.f;
    qa_s[i] = i < QD ? qa[i] : 0.f;
  }
  // rows 200..207 of every resident W_a chunk are never written by the TMA (box = 200 rows): keep them zero
  for (int i = tid; i < KCH * 8 * 8; i += THREADS) {
    const int kc = i / 64, r = 200 + (i % 64) / 8, c = i % 8;
    *reinterpret_cast<uint4*>(sm + OFF_B + kc * B_CHUNK + r * 128 + c * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      expect_tx(b_full, KCH * QD * 128);
      for (int kc = 0; kc < KCH; ++kc) tma_load_2d(base + OFF_B + kc * B_CHUNK, &tmap_wa, kc * 64, 0, b_full);
      uint32_t it = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int r0 = (int)(t * ROWS);
        for (int kc = 0; kc < KCH; ++kc, ++it) {
          const int s = it % NSTA;
          mbar_wait(empty_bar + 8 * s, ((it / NSTA) & 1) ^ 1);
          expect_tx(full_bar + 8 * s, a_tx_bytes);
          tma_load_2d(base + OFF_A + s * A_SLOT, &tmap_c, kc * 64, r0, full_bar + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_f16(128, BN);
    const uint64_t desc0 = umma_desc_k_sw128(0);
    const uint32_t el = elect_one_u32();     // whole warp converged: see tc_common.cuh umma_*_p
    mbar_wait(b_full, 0);
    uint32_t it = 0, tile_it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_it) {
      const uint32_t as = tile_it & 1;
      mbar_wait(tempty_bar + 8 * as, ((tile_it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * 256;
      for (int kc = 0; kc < KCH; ++kc, ++it) {
        const int s = it % NSTA;
        mbar_wait(full_bar + 8 * s, (it / NSTA) & 1);
        tc_fence_after();
        const uint32_t sa = (base + OFF_A + s * A_SLOT) >> 4;
        const uint32_t sb = (base + OFF_B + kc * B_CHUNK) >> 4;
        const int ksteps = (kc == KCH - 1) ? 3 : 4;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          if (ks < ksteps)
            umma_f16_ss_p(d_tmem, desc0 | (uint64_t)((sa + 2 * ks) & 0x3FFF), desc0 | (uint64_t)((sb + 2 * ks) & 0x3FFF),
                          idesc, (kc | ks) ? 1u : 0u, el);
        umma_commit_p(empty_bar + 8 * s, el);
        if (kc == KCH - 1) umma_commit_p(tfull_bar + 8 * as, el);
      }
    }
  } else if (warp <= 5) {
    // ------------------------------ score warps -------------------------------------------------
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const bool row_ok = row < ROWS;
    float qa_l1 = 0.f;                       // bound of |s|: sum of |attention_query_vector| (qa_s is padded with zeros)
    for (int j = 0; j < 208; ++j) qa_l1 += fabsf(qa_s[j]);
    uint32_t tile_it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_it) {
      const uint32_t as = tile_it & 1;
      mbar_wait(tfull_bar + 8 * as, (tile_it >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + as * 256 + ((uint32_t)(q4 * 32) << 16);
      float s = 0.f;
#pragma unroll 1
      for (int col = 0; col < 192; col += 32) {
        uint32_t v[32];
        tmem_ld16_nw(trow + col, v);
        tmem_ld16_nw(trow + col + 16, v + 16);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) s = fmaf(fast_tanh(__uint_as_float(v[j]) + ba_s[col + j]), qa_s[col + j], s);
      }
      {   // columns 192..199 (200..207 are padding)
        uint32_t v[8];
        tmem_ld8_nw(trow + 192, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) s = fmaf(fast_tanh(__uint_as_float(v[j]) + ba_s[192 + j]), qa_s[192 + j], s);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * as);
      mbar_wait(wv_free + 8 * as, ((tile_it >> 1) & 1) ^ 1);   // the pool warps are done with this buffer
      // softmax over the sequence.  |s_i| <= sum_j |q_j| =: L1 because |tanh| <= 1, so exp(s_i - L1) can neither
      // overflow nor (for L1 <= 40) underflow: every row takes ONE exponential and the 50-element scan is a plain sum,
      // instead of a max scan plus 50 exponentials per row on the critical path of the score warps.  Softmax is
      // shift-invariant, so this is the reference's stable softmax (additive.py:37-39) up to fp32 rounding; a query
      // vector with L1 > 40 keeps the max-subtracting form.
      if (qa_l1 <= 40.f) {
        sc[as * 128 + row] = __expf(s - qa_l1);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (row_ok) {
          const int sq = row / S;
          const float* scs = sc + as * 128 + sq * S;
          float s0 = 0.f, s1 = 0.f;
#pragma unroll 5
          for (int j = 0; j < S; j += 2) { s0 += scs[j]; s1 += scs[j + 1]; }
          wv[as * 128 + row] = __fdividef(sc[as * 128 + row], s0 + s1);
        }
      } else {
        sc[as * 128 + row] = s;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (row_ok) {
          const int sq = row / S;
          const float* scs = sc + as * 128 + sq * S;
          float m = -INFINITY;
          for (int j = 0; j < S; ++j) m = fmaxf(m, scs[j]);
          float sum = 0.f;
          for (int j = 0; j < S; ++j) sum += __expf(scs[j] - m);
          wv[as * 128 + row] = __fdividef(__expf(s - m), sum);
        }
      }
      mbar_arrive(wv_ready + 8 * as);
    }
  } else {
    // ------------------------------ pool warps --------------------------------------------------
    const int pt = (warp - 6) * 32 + lane;     // 0..POOL_THREADS-1
    uint32_t tile_it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_it) {
      const int64_t seq0 = t * SPT;
      const uint32_t as = tile_it & 1;
      mbar_wait(wv_ready + 8 * as, (tile_it >> 1) & 1);
      const float* w = wv + as * 128;
      // pooled[seq, 4l..4l+3] = sum_i w_i C[seq*S+i, 4l..4l+3]; rows are L2-hot (K1 just wrote them)
      for (int o = pt; o < SPT * DV4; o += POOL_THREADS) {
        const int sq = o / DV4, l = o - sq * DV4;
        if (seq0 + sq < n_seq) {
          const uint2* cp = reinterpret_cast<const uint2*>(C + (seq0 + sq) * S * CP) + l;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          // the rows come back from L2 (~800 cycles under load): keep 25 (S = 50) / 20 (S = 20) loads in flight
          constexpr int PB = (S % 25 == 0) ? 25 : 20;
          static_assert(S % PB == 0, "pool batch");
#pragma unroll 1
          for (int i0 = 0; i0 < S; i0 += PB) {
            uint2 c4[PB];
#pragma unroll
            for (int i = 0; i < PB; ++i) c4[i] = __ldg(cp + (i0 + i) * (CP / 4));
#pragma unroll
            for (int i = 0; i < PB; ++i) {
              const float wi = w[sq * S + i0 + i];
              const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&c4[i].x));
              const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&c4[i].y));
              acc.x = fmaf(wi, lo.x, acc.x); acc.y = fmaf(wi, lo.y, acc.y);
              acc.z = fmaf(wi, hi.x, acc.z); acc.w = fmaf(wi, hi.y, acc.w);
            }
          }
          reinterpret_cast<float4*>(out + (seq0 + sq) * D)[l] = acc;
        }
      }
      mbar_arrive(wv_free + 8 * as);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// fp16 copy of W_a: [200][320] halfs, K tail zero
__global__ void __launch_bounds__(256) pack_wa16_kernel(const float* __restrict__ wa, __half* __restrict__ out) {
  const int n = QD * CP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / CP, k = i - r * CP;
    out[i] = __float2half_rn(k < D ? wa[r * D + k] : 0.f);
  }
}

}  // namespace k2v2

int k2v2_prepare(const float* wa, void* wa16, CUtensorMap* twa, cudaStream_t st) {
  k2v2::pack_wa16_kernel<<<64, 256, 0, st>>>(wa, reinterpret_cast<__half*>(wa16));
  NRMS_LAUNCH_CHECK("pack_wa16_kernel");
  return make_tmap_k_major_f16(twa, wa16, QD, k2v2::CP, k2v2::CP, QD);
}

template <int S, int SPT>
static int launch_k2v2(const CUtensorMap& twa, const void* Cbuf, int64_t n, const float* ba, const float* qa,
                       float* out, cudaStream_t st) {
  static bool configured[64] = {false};
  cudaError_t e = set_max_dynamic_smem(k2v2::additive_pool_f16_kernel<S, SPT>, k2v2::SMEM, configured);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(additive_pool_f16_kernel)");
  constexpr int ROWS = S * SPT;
  alignas(64) CUtensorMap tc_;
  const int box_c = (int)((n * S < ROWS) ? n * S : ROWS);
  if (int rc = make_tmap_k_major_f16(&tc_, Cbuf, n * S, k2v2::CP, k2v2::CP, box_c)) return rc;
  const int64_t tiles = (n + SPT - 1) / SPT;
  int grid = num_sms();
  if (tiles < grid) grid = (int)tiles;
  k2v2::additive_pool_f16_kernel<S, SPT><<<grid, k2v2::THREADS, k2v2::SMEM, st>>>(
      tc_, twa, reinterpret_cast<const __half*>(Cbuf), ba, qa, out, n, (uint32_t)box_c * 128u);
  NRMS_LAUNCH_CHECK("additive_pool_f16_kernel");
  return NRMS_OK;
}

int k2v2_run(int S, const CUtensorMap& twa, const void* Cbuf, int64_t n, const float* ba, const float* qa, float* out,
             cudaStream_t st) {
  if (S == 20) return launch_k2v2<20, 5>(twa, Cbuf, n, ba, qa, out, st);
  if (S == 50) return launch_k2v2<50, 2>(twa, Cbuf, n, ba, qa, out, st);
  set_error("additive_pool_f16_kernel compiled for S = 20 or 50, got %d", S);
  return NRMS_E_UNSUPPORTED;
}

}  // namespace nrms
