// Generic TF32 tensor-core GEMM for sm_100a:  C[M,N] (=, +=, atomic +=) A[M,K] * B[N,K]^T (+ bias), optionally
// split along K (one work item = output tile x K range; the training backward's weight gradients reduce over
// 10^5 rows into a handful of tiles).
//
// Persistent, warp-specialised, TMA-fed:
//   warp 0     : TMA producer.  One thread issues cp.async.bulk.tensor.2d loads of a 128x32 (A) and a
//                256x32 (B) fp32 box per 128-byte K chunk into a 4-stage ring; the tensor maps use
//                SWIZZLE_128B, i.e. exactly the canonical UMMA K-major layout; out-of-bounds rows / K
//                tail are zero-filled by the TMA unit.
//   warp 1     : TMEM allocation (512 columns = two 256-column accumulator stages) and single-thread
//                tcgen05.mma.kind::tf32 issue; tcgen05.commit releases smem stages / publishes accumulators.
//   warps 2..5 : epilogue: tcgen05.ld (32 lanes x 16 columns) -> +bias -> st.global, overlapped with the
//                next tile's main loop through the second accumulator stage.
// Tiles are walked n-fastest so the A tile of an M block stays L2-resident across its N tiles.
// The tensor maps are typed TFLOAT32: the TMA unit rounds the fp32 operands to TF32 while copying, so
// callers pass plain fp32 tensors (no pre-rounding pass) and the tensor core never truncates.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>
#include "tc_common.cuh"
#include "tc_api.cuh"

namespace nrms {
using namespace tc;

constexpr int TG_BM = 128, TG_BN = 256, TG_BK = 32, TG_STAGES = 4;
constexpr int TG_A_BYTES = TG_BM * TG_BK * 4;  // 16 KB
constexpr int TG_B_BYTES = TG_BN * TG_BK * 4;  // 32 KB
constexpr int TG_STAGE_BYTES = TG_A_BYTES + TG_B_BYTES;
constexpr int TG_THREADS = 192;
constexpr int TG_SMEM = TG_STAGES * TG_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
// TMA-store epilogue (fp16 output): 3 pipeline stages + one 16 KB staging buffer per epilogue warp (4 boxes of 32 rows x 128 B)
constexpr int TG_STAGES_TMA = 3;
constexpr int TG_STG_WARP = 4 * 32 * 128;
constexpr int TG_SMEM_TMA = TG_STAGES_TMA * TG_STAGE_BYTES + 4 * TG_STG_WARP + 1024 + 256;
static_assert(TG_SMEM_TMA <= 232448, "shared memory");

__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes)
               : "memory");
}

// F16 = false: fp32 operands rounded to TF32 by the TMA unit, kind::tf32 (K = 8 per MMA, 32 floats per 128-byte chunk);
// F16 = true : fp16 operands, kind::f16 (K = 16 per MMA, 64 halfs per chunk): half the operand bytes per flop and half the
//              MMA count -- the contraction the table projection of K1g uses (bias rides in a 1.0 column of A).
// TMA_EPI = true (F16 only): the epilogue rounds the accumulator to fp16 into a 128B-swizzled staging tile and leaves
//              with bulk tensor stores (one per 32 rows x 64 columns) instead of per-thread stores: with thread == row
//              every store instruction touched 32 different rows (32 requests of 16 bytes), which bound the table
//              projection of K1g at 147 us for 35 GFLOP.
// (MN-major operand descriptors for kind::tf32 -- which would let the weight gradients dW = dY^T X read the activations
// in place -- were tried in round 2: with either major bit of the instruction descriptor set the MMA returns zeros on
// sm_100a, as on Hopper where only 16-bit operands may be MN-major.  The transposed copies of the backward stay.)
template <bool F16, bool TMA_EPI = false>
__global__ void __launch_bounds__(TG_THREADS, 1)
tc_gemm_nt_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_c,
                  const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int64_t M, int N, int K,
                  uint32_t stage_tx_bytes, int k_splits, int chunks_per_split, int epi, float f16_scale, int f16_scale_cols) {
  constexpr int TG_STAGES = TMA_EPI ? TG_STAGES_TMA : nrms::TG_STAGES;
  constexpr int STG_BYTES = TMA_EPI ? 4 * TG_STG_WARP : 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + TG_STAGES * TG_STAGE_BYTES + STG_BYTES;
  const uint32_t full_bar = bars;                          // [STAGES]
  const uint32_t empty_bar = bars + 8 * TG_STAGES;         // [STAGES]
  const uint32_t tfull_bar = bars + 16 * TG_STAGES;        // [2]
  const uint32_t tempty_bar = tfull_bar + 16;              // [2]
  volatile uint32_t* tmem_ptr_smem =
      reinterpret_cast<volatile uint32_t*>(base_ptr + TG_STAGES * TG_STAGE_BYTES + STG_BYTES + 16 * TG_STAGES + 32);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler (converged MMA issue)
  const int n_tiles = (N + TG_BN - 1) / TG_BN;
  const int64_t m_tiles = (M + TG_BM - 1) / TG_BM;
  // Work items: a CTA takes ROW BLOCKS (m tile, K split) round-robin and walks ALL n tiles of each one back to back.
  // (Round 2, from the ncu captures of the training GEMMs: with items dealt out singly, w = blockIdx + k * grid and
  // n = w % n_tiles, a grid of 148 CTAs and N = 300 (n tiles of 256 and 44 columns) gave every even CTA only wide tiles and
  // every odd CTA only the 44-column slivers -- half the SMs idle -- and the two n tiles of a row block drifted apart in
  // time, so the A row block came from DRAM twice: 903 MB read for a 507 MB operand.)
  // With under half a wave of row blocks (the user-encoder GEMMs: 50 m tiles) the host launches one CTA per (row block,
  // n tile) instead and the n tiles are dealt out singly again (gridDim > row blocks tells the kernel).
  const int64_t total_rb = m_tiles * k_splits;
  const int ng = total_rb >= (int64_t)gridDim.x ? n_tiles : 1;          // n tiles walked back to back by one CTA
  const int64_t total_groups = total_rb * (n_tiles / ng);
  const int64_t my_groups = (int64_t)blockIdx.x < total_groups ? (total_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t my_items = my_groups * ng;
  auto item_w = [&](int64_t i) -> int64_t {
    const int64_t grp = blockIdx.x + (i / ng) * gridDim.x;
    const int64_t rb = ng == 1 ? grp / n_tiles : grp;
    const int64_t n = ng == 1 ? grp % n_tiles : i % ng;
    return ((rb / k_splits) * n_tiles + n) * k_splits + (rb % k_splits);      // K split fastest, as chunk_range expects
  };
  constexpr int BK = F16 ? 2 * TG_BK : TG_BK;          // elements per 128-byte K chunk
  constexpr int KSTEP = F16 ? 16 : 8;                  // elements per MMA
  const int all_chunks = (K + BK - 1) / BK;
  // chunk range of work item w (the host guarantees every split is non-empty)
  auto chunk_range = [&](int64_t w, int& c0, int& c1) {
    c0 = (int)(w % k_splits) * chunks_per_split;
    c1 = c0 + chunks_per_split;
    if (c1 > all_chunks) c1 = all_chunks;
  };

  if (tid == 0) {
    for (int s = 0; s < TG_STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);
      mbar_init(tempty_bar + 8 * a, 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_ptr_smem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------ TMA producer ---------------------------------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t i = 0; i < my_items; ++i) {
        const int64_t w = item_w(i);
        const int64_t t = w / k_splits;
        const int m0 = (int)(t / n_tiles) * TG_BM;
        const int n0 = (int)(t % n_tiles) * TG_BN;
        int c0, c1;
        chunk_range(w, c0, c1);
        for (int kc = c0; kc < c1; ++kc, ++it) {
          const int s = it % TG_STAGES;
          const uint32_t ph = (it / TG_STAGES) & 1;
          mbar_wait(empty_bar + 8 * s, ph ^ 1);
          mbar_expect_tx(full_bar + 8 * s, stage_tx_bytes);
          const uint32_t sa = base + s * TG_STAGE_BYTES;
          tma_load_2d(sa, &tmap_a, kc * BK, m0, full_bar + 8 * s);
          tma_load_2d(sa + TG_A_BYTES, &tmap_b, kc * BK, n0, full_bar + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (whole warp converged, see tc_common.cuh umma_*_p) ---
    const uint32_t el = elect_one_u32();
    const uint64_t desc0 = umma_desc_k_sw128(0);
    uint32_t it = 0, tile_it = 0;
    for (int64_t i = 0; i < my_items; ++i, ++tile_it) {
      const int64_t w = item_w(i);
      const int64_t t = w / k_splits;
      const int n0 = (int)(t % n_tiles) * TG_BN;
      int n_valid = N - n0;
      if (n_valid > TG_BN) n_valid = TG_BN;
      const int n_mma = (n_valid + 15) & ~15;
      const uint32_t idesc = F16 ? umma_idesc_f16(TG_BM, n_mma) : umma_idesc_tf32(TG_BM, n_mma);
      const uint32_t as = tile_it & 1;
      int c0, c1;
      chunk_range(w, c0, c1);
      mbar_wait(tempty_bar + 8 * as, ((tile_it >> 1) & 1) ^ 1);   // epilogue drained this accumulator stage
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * TG_BN;
      for (int kc = c0; kc < c1; ++kc, ++it) {
        const int s = it % TG_STAGES;
        const uint32_t ph = (it / TG_STAGES) & 1;
        mbar_wait(full_bar + 8 * s, ph);
        tc_fence_after();
        const uint32_t sa = (base + s * TG_STAGE_BYTES) >> 4;
        const uint32_t sb = sa + (TG_A_BYTES >> 4);
        int ksteps = (K - kc * BK + KSTEP - 1) / KSTEP;
        if (ksteps > 4) ksteps = 4;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          if (ks < ksteps) {
            if (F16)
              umma_f16_ss_p(d_tmem, desc0 | (uint64_t)((sa + 2 * ks) & 0x3FFF), desc0 | (uint64_t)((sb + 2 * ks) & 0x3FFF),
                            idesc, (kc > c0 || ks) ? 1u : 0u, el);
            else
              umma_tf32_ss_p(d_tmem, desc0 | (uint64_t)((sa + 2 * ks) & 0x3FFF), desc0 | (uint64_t)((sb + 2 * ks) & 0x3FFF),
                             idesc, (kc > c0 || ks) ? 1u : 0u, el);
          }
        umma_commit_p(empty_bar + 8 * s, el);
        if (kc == c1 - 1) umma_commit_p(tfull_bar + 8 * as, el);
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..5) ------------------------------
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    uint32_t tile_it = 0;
    for (int64_t i = 0; i < my_items; ++i, ++tile_it) {
      const int64_t w = item_w(i);
      const int64_t t = w / k_splits;
      const int64_t m0 = (t / n_tiles) * TG_BM;
      const int n0 = (int)(t % n_tiles) * TG_BN;
      int n_valid = N - n0;
      if (n_valid > TG_BN) n_valid = TG_BN;
      const bool add_bias = bias != nullptr && (w % k_splits) == 0;
      const uint32_t as = tile_it & 1;
      mbar_wait(tfull_bar + 8 * as, (tile_it >> 1) & 1);
      tc_fence_after();
      const int64_t m = m0 + q * 32 + lane;
      const uint32_t trow = tmem_base + as * TG_BN + ((uint32_t)(q * 32) << 16);
      if constexpr (TMA_EPI && !F16) {
        // fp32 result: the warp's 32 rows leave through a 128B-swizzled staging tile (4 boxes of 32 rows x 32 floats =
        // 128 columns at a time) and bulk tensor stores -- or bulk tensor REDUCTIONS (add.f32, performed in L2) for the
        // accumulate / split-K epilogues.  With thread == row a plain store instruction touched 32 different rows (32
        // requests of 16 bytes); the forward QKV projection of a training step ran at 1.3 TB/s of its 676 MB because of it.
        uint8_t* const stg = base_ptr + TG_STAGES * TG_STAGE_BYTES + (warp - 2) * TG_STG_WARP;
        const uint32_t stg_s = base + TG_STAGES * TG_STAGE_BYTES + (warp - 2) * TG_STG_WARP;
        for (int h0 = 0; h0 < n_valid; h0 += 128) {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // earlier stores have read the tile
          __syncwarp();
          const int h1 = (h0 + 128 < n_valid) ? h0 + 128 : n_valid;
          for (int col = h0; col < h1; col += 16) {
            float v[16];
            tmem_ld16(trow + col, v);
            if (add_bias) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int n = n0 + col + 4 * j;
                if (n < N) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
                  v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
                }
              }
            }
            uint8_t* const brow = stg + ((col - h0) >> 5) * 4096 + lane * 128;
            const int c4 = (col & 31) >> 2;                 // first 16-byte chunk of these 16 floats inside the 32-float row
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<float4*>(brow + (((c4 + j) ^ (lane & 7)) << 4)) =
                  make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (h1 == n_valid) tc_fence_before();             // the accumulator stage is fully read
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (h1 == n_valid) mbar_arrive(tempty_bar + 8 * as);
            for (int b = 0; h0 + 32 * b < h1; ++b) {
              if (epi == TC_EPI_STORE)
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                             ::"l"(reinterpret_cast<uint64_t>(&tmap_c)), "r"(n0 + h0 + 32 * b), "r"((int)(m0 + q * 32)),
                               "r"(stg_s + b * 4096) : "memory");
              else
                asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
                             ::"l"(reinterpret_cast<uint64_t>(&tmap_c)), "r"(n0 + h0 + 32 * b), "r"((int)(m0 + q * 32)),
                               "r"(stg_s + b * 4096) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        continue;
      }
      if constexpr (TMA_EPI && F16) {
        // fp16 tile rows of this warp -> staging (4 boxes of 32 rows x 64 halfs, 16-byte chunk c of row r at c ^ (r & 7))
        // -> bulk tensor stores; rows / columns past the tensor are clipped by the TMA unit
        uint8_t* const stg = base_ptr + TG_STAGES * TG_STAGE_BYTES + (warp - 2) * TG_STG_WARP;
        const uint32_t stg_s = base + TG_STAGES * TG_STAGE_BYTES + (warp - 2) * TG_STG_WARP;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // previous tile's stores have read it
        __syncwarp();
        for (int col = 0; col < n_valid; col += 16) {
          float v[16];
          tmem_ld16(trow + col, v);
          uint32_t h[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const __half2 p = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
            h[j] = *reinterpret_cast<const uint32_t*>(&p);
          }
          uint8_t* const brow = stg + (col >> 6) * 4096 + lane * 128;
          const int c16 = (col & 63) >> 3;
          *reinterpret_cast<uint4*>(brow + (((c16) ^ (lane & 7)) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(brow + (((c16 + 1) ^ (lane & 7)) << 4)) = make_uint4(h[4], h[5], h[6], h[7]);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(tempty_bar + 8 * as);
          for (int b = 0; b * 64 < n_valid; ++b)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                         ::"l"(reinterpret_cast<uint64_t>(&tmap_c)), "r"(n0 + 64 * b), "r"((int)(m0 + q * 32)), "r"(stg_s + b * 4096)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        continue;
      }
      for (int col = 0; col < n_valid; col += 16) {
        float v[16];
        tmem_ld16(trow + col, v);
        if (epi >= TC_EPI_STORE_F16) {
          // C is a __half matrix (ldc in halfs, rows 8-byte aligned): (acc + bias) * (n < f16_scale_cols ? f16_scale : 1)
          if (m < M) {
            __half* crow = reinterpret_cast<__half*>(C) + m * ldc;
            uint2 pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = n0 + col + 4 * j;
              float4 r = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              if (add_bias && n < N) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
                r = make_float4(r.x + b4.x, r.y + b4.y, r.z + b4.z, r.w + b4.w);
              }
              const float sc = n < f16_scale_cols ? f16_scale : 1.f;
              const __half2 lo = __floats2half2_rn(r.x * sc, r.y * sc), hi = __floats2half2_rn(r.z * sc, r.w * sc);
              pk[j].x = *reinterpret_cast<const uint32_t*>(&lo);
              pk[j].y = *reinterpret_cast<const uint32_t*>(&hi);
            }
            if (epi == TC_EPI_STORE_F16_QKV) {
              // K1g table row: [head group][q | k | v][5 heads][24 halfs]; column n = which*300 + head*20 + d.
              // The 4-half pad of every head slice goes with its last group: 0 (q, k) or (1,0,0,0) (v).  Groups that
              // are neighbours inside a slice (d = 0|4, 8|12, 16|pad) leave as ONE 16-byte store: the epilogue is
              // bound by the number of store requests (32 rows per warp instruction), not by bytes.
              bool merged = false;           // group j already left with group j-1
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int n = n0 + col + 4 * j;
                const bool skip = merged || n >= N;
                merged = false;
                if (skip) continue;
                const int which = n / 300, hd = n - which * 300, head = hd / 20, d = hd - head * 20;
                const int hgi = head / 5, hl = head - hgi * 5;
                __half* dst = crow + hgi * 360 + which * 120 + hl * 24 + d;
                const uint2 nxt = pk[j < 3 ? j + 1 : 3];
                if (d == 16) {
                  *reinterpret_cast<uint4*>(dst) = make_uint4(pk[j].x, pk[j].y, which == 2 ? 0x00003C00u : 0u, 0u);
                } else if ((d == 0 || d == 8) && j < 3 && n + 4 < N) {
                  *reinterpret_cast<uint4*>(dst) = make_uint4(pk[j].x, pk[j].y, nxt.x, nxt.y);
                  merged = true;
                } else {
                  *reinterpret_cast<uint2*>(dst) = pk[j];
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int n = n0 + col + 4 * j;
                if (n < N) *reinterpret_cast<uint2*>(crow + n) = pk[j];
              }
            }
          }
          continue;
        }
        if (m < M) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = n0 + col + 4 * j;
            if (n < N) {
              float4 r = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              if (add_bias) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
                r = make_float4(r.x + b4.x, r.y + b4.y, r.z + b4.z, r.w + b4.w);
              }
              float4* dst = reinterpret_cast<float4*>(C + m * ldc + n);
              if (epi == TC_EPI_STORE) {
                *dst = r;
              } else if (epi == TC_EPI_ACCUM) {
                const float4 o = *dst;
                *dst = make_float4(o.x + r.x, o.y + r.y, o.z + r.z, o.w + r.w);
              } else {
                atomicAdd(dst, r);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * as);
    }
  }
  if (TMA_EPI && warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- host: tensor maps ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 row-major [rows, cols] (row stride ld floats) -> SWIZZLE_128B boxes of box_rows x 32 floats
int make_tmap_k_major(CUtensorMap* out, const float* base, int64_t rows, int cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  NRMS_CHECK_ARG(fn != nullptr, NRMS_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  static int dtype_override = -1;
  if (dtype_override < 0) {
    // CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 makes the TMA unit ROUND fp32 to TF32 on the way into shared
    // memory (measured, profiles/tma_round_probe.py: max error 3.1e-4 vs 8.1e-4 with plain FLOAT32, where
    // the tensor core truncates the low 13 mantissa bits).  NRMS_TMA_TF32_DTYPE=0 selects FLOAT32 (debug).
    const char* e = getenv("NRMS_TMA_TF32_DTYPE");
    dtype_override = (e && e[0] == '0') ? 0 : 1;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(out, dtype_override ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NRMS_CHECK_ARG(r == CUDA_SUCCESS, NRMS_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return NRMS_OK;
}

// fp16 row-major [rows, cols] (row stride ld halfs, ld*2 % 16 == 0) -> SWIZZLE_128B boxes of box_rows x 64 halfs
int make_tmap_k_major_f16(CUtensorMap* out, const void* base, int64_t rows, int cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  NRMS_CHECK_ARG(fn != nullptr, NRMS_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NRMS_CHECK_ARG(r == CUDA_SUCCESS, NRMS_E_CUDA, "cuTensorMapEncodeTiled(f16) failed with CUresult %d", (int)r);
  return NRMS_OK;
}

// fp16 row-major [rows, cols], un-swizzled box of box_rows x box_cols halfs (box_cols*2 % 16 == 0): TMA stores
int make_tmap_store_f16(CUtensorMap* out, const void* base, int64_t rows, int cols, int64_t ld, int box_cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  NRMS_CHECK_ARG(fn != nullptr, NRMS_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NRMS_CHECK_ARG(r == CUDA_SUCCESS, NRMS_E_CUDA, "cuTensorMapEncodeTiled(store f16) failed with CUresult %d", (int)r);
  return NRMS_OK;
}

// fp16 row-major [rows, cols]: 128B-swizzled boxes of 32 rows x 64 halfs for the bulk-store epilogue
static int make_tmap_store_f16_sw128(CUtensorMap* out, const void* base, int64_t rows, int cols, int64_t ld) {
  EncodeTiledFn fn = get_encode_fn();
  NRMS_CHECK_ARG(fn != nullptr, NRMS_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, rows < 32 ? (cuuint32_t)rows : 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NRMS_CHECK_ARG(r == CUDA_SUCCESS, NRMS_E_CUDA, "cuTensorMapEncodeTiled(store f16 sw128) failed with CUresult %d", (int)r);
  return NRMS_OK;
}

static bool g_f32_tma_epilogue = true;      // "gemm_tma_epilogue" option (A/B): 0 = per-thread stores / atomics
void set_gemm_tma_epilogue(bool on) { g_f32_tma_epilogue = on; }

// fp32 row-major [rows, cols]: 128B-swizzled boxes of 32 rows x 32 floats for the bulk-store / bulk-reduce epilogue
static int make_tmap_store_f32_sw128(CUtensorMap* out, const void* base, int64_t rows, int cols, int64_t ld) {
  EncodeTiledFn fn = get_encode_fn();
  NRMS_CHECK_ARG(fn != nullptr, NRMS_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, rows < 32 ? (cuuint32_t)rows : 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NRMS_CHECK_ARG(r == CUDA_SUCCESS, NRMS_E_CUDA, "cuTensorMapEncodeTiled(store f32 sw128) failed with CUresult %d", (int)r);
  return NRMS_OK;
}

static int tc_gemm_launch(const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias, float* C, int64_t ldc,
                         int64_t M, int N, int K, int k_splits, int epi, float f16_scale, int f16_scale_cols, bool f16_in,
                         cudaStream_t st);

// fp16 operands, fp16 result C16[m][n] = half(sum_k A16[m][k] B16[n][k]) (row-major, ldc halfs, 16-byte aligned rows), leaving
// through bulk tensor stores.  No bias / scale: the callers fold them into B (a bias column met by a 1.0 column of A).
int tc_gemm_nt_f16_tma(const void* A16, int64_t lda, const void* B16, int64_t ldb, void* C16, int64_t ldc, int64_t M, int N,
                       int K, cudaStream_t st) {
  NRMS_CHECK_ARG((lda % 8) == 0 && (ldb % 8) == 0 && (ldc % 8) == 0 && aligned16(A16) && aligned16(B16) && aligned16(C16) &&
                     (K % 8) == 0 && (N % 8) == 0,
                 NRMS_E_INVALID, "fp16 GEMM operands must be 16-byte aligned rows");
  return tc_gemm_launch(A16, lda, B16, ldb, nullptr, reinterpret_cast<float*>(C16), ldc, M, N, K, 1, TC_EPI_STORE_F16_TMA,
                        1.f, 0, true, st);
}

int tc_gemm_nt_ex(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, float* C, int64_t ldc,
                  int64_t M, int N, int K, int k_splits, int epi, cudaStream_t st) {
  NRMS_CHECK_ARG(epi >= TC_EPI_STORE && epi <= TC_EPI_ATOMIC && (k_splits <= 1 || epi == TC_EPI_ATOMIC), NRMS_E_INVALID,
                 "split-K needs the atomic epilogue");
  return tc_gemm_launch(A, lda, B, ldb, bias, C, ldc, M, N, K, k_splits, epi, 1.f, 0, false, st);
}

int tc_gemm_nt_f16out(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, void* C16, int64_t ldc,
                      int64_t M, int N, int K, float scale, int scale_cols, int qkv_layout, cudaStream_t st) {
  NRMS_CHECK_ARG((ldc % 4) == 0 && (reinterpret_cast<uintptr_t>(C16) & 7) == 0 && (scale_cols % 4) == 0, NRMS_E_INVALID,
                 "fp16 output rows must be 8-byte aligned");
  NRMS_CHECK_ARG(!qkv_layout || (N == 900 && ldc >= 1080), NRMS_E_INVALID, "the q|k|v head-group layout is [*, 1080] from N = 900");
  return tc_gemm_launch(A, lda, B, ldb, bias, reinterpret_cast<float*>(C16), ldc, M, N, K, 1,
                        qkv_layout ? TC_EPI_STORE_F16_QKV : TC_EPI_STORE_F16, scale, scale_cols, false, st);
}

// fp16 operands (A16 [M, lda halfs], B16 [N, ldb halfs], rows 16-byte aligned), fp16 result; same epilogue as above
int tc_gemm_nt_f16(const void* A16, int64_t lda, const void* B16, int64_t ldb, void* C16, int64_t ldc, int64_t M, int N,
                   int K, float scale, int scale_cols, int qkv_layout, cudaStream_t st) {
  NRMS_CHECK_ARG((lda % 8) == 0 && (ldb % 8) == 0 && (ldc % 4) == 0 && aligned16(A16) && aligned16(B16) &&
                     (reinterpret_cast<uintptr_t>(C16) & 7) == 0 && (scale_cols % 4) == 0 && (K % 8) == 0,
                 NRMS_E_INVALID, "fp16 GEMM operands must be 16-byte aligned rows");
  NRMS_CHECK_ARG(!qkv_layout || (N == 900 && ldc >= 1080), NRMS_E_INVALID, "the q|k|v head-group layout is [*, 1080] from N = 900");
  return tc_gemm_launch(A16, lda, B16, ldb, nullptr, reinterpret_cast<float*>(C16), ldc, M, N, K, 1,
                        qkv_layout ? TC_EPI_STORE_F16_QKV : TC_EPI_STORE_F16, scale, scale_cols, true, st);
}

static int tc_gemm_launch(const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias, float* C, int64_t ldc,
                         int64_t M, int N, int K, int k_splits, int epi, float f16_scale, int f16_scale_cols, bool f16_in,
                         cudaStream_t st) {
  if (M <= 0) return NRMS_OK;
  NRMS_CHECK_ARG(M < (1ll << 31), NRMS_E_UNSUPPORTED, "M too large for one tensor map");
  static bool cfg_a[64] = {false}, cfg_b[64] = {false}, cfg_c[64] = {false}, cfg_d[64] = {false};
  cudaError_t e = set_max_dynamic_smem(tc_gemm_nt_kernel<false>, TG_SMEM, cfg_a);
  if (e == cudaSuccess) e = set_max_dynamic_smem(tc_gemm_nt_kernel<false, true>, TG_SMEM_TMA, cfg_d);
  if (e == cudaSuccess) e = set_max_dynamic_smem(tc_gemm_nt_kernel<true>, TG_SMEM, cfg_b);
  if (e == cudaSuccess) e = set_max_dynamic_smem(tc_gemm_nt_kernel<true, true>, TG_SMEM_TMA, cfg_c);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_gemm_nt_kernel)");
  alignas(64) CUtensorMap ta, tb, tcm;
  memset(&tcm, 0, sizeof(tcm));
  if (epi == TC_EPI_STORE_F16_TMA) {
    NRMS_CHECK_ARG(f16_in, NRMS_E_INVALID, "the bulk-store epilogue needs fp16 operands");
    if (int rc = make_tmap_store_f16_sw128(&tcm, C, M, N, ldc)) return rc;
  }
  // fp32 results (store / accumulate / split-K) leave through bulk tensor stores / reductions
  const bool f32_tma = !f16_in && epi <= TC_EPI_ATOMIC && g_f32_tma_epilogue;
  if (f32_tma) {
    if (int rc = make_tmap_store_f32_sw128(&tcm, C, M, N, ldc)) return rc;
  }
  // boxes never exceed the tensor extent (rows past it would only feed outputs that are not stored)
  const int box_a = M < TG_BM ? (int)M : TG_BM;
  const int box_b = N < TG_BN ? N : TG_BN;
  if (f16_in) {
    if (int rc = make_tmap_k_major_f16(&ta, A, M, K, lda, box_a)) return rc;
    if (int rc = make_tmap_k_major_f16(&tb, B, N, K, ldb, box_b)) return rc;
  } else {
    if (int rc = make_tmap_k_major(&ta, reinterpret_cast<const float*>(A), M, K, lda, box_a)) return rc;
    if (int rc = make_tmap_k_major(&tb, reinterpret_cast<const float*>(B), N, K, ldb, box_b)) return rc;
  }
  const uint32_t stage_tx = (uint32_t)(box_a + box_b) * TG_BK * 4;       // 128 bytes per row and chunk, either type
  const int bk = f16_in ? 2 * TG_BK : TG_BK;
  const int chunks = (K + bk - 1) / bk;
  if (k_splits < 1) k_splits = 1;
  if (k_splits > chunks) k_splits = chunks;
  const int cps = (chunks + k_splits - 1) / k_splits;
  k_splits = (chunks + cps - 1) / cps;            // every split owns at least one chunk
  int64_t items = ((M + TG_BM - 1) / TG_BM) * k_splits;     // row blocks (m tile, K split): see the kernel
  int grid = num_sms();
  if (2 * items < grid) items *= (N + TG_BN - 1) / TG_BN;    // under half a wave of row blocks: n tiles dealt out singly
  if (items < grid) grid = (int)items;
  if (epi == TC_EPI_STORE_F16_TMA)
    tc_gemm_nt_kernel<true, true><<<grid, TG_THREADS, TG_SMEM_TMA, st>>>(ta, tb, tcm, bias, C, ldc, M, N, K, stage_tx, k_splits,
                                                                          cps, epi, f16_scale, f16_scale_cols);
  else if (f32_tma)
    tc_gemm_nt_kernel<false, true><<<grid, TG_THREADS, TG_SMEM_TMA, st>>>(ta, tb, tcm, bias, C, ldc, M, N, K, stage_tx, k_splits,
                                                                           cps, epi, f16_scale, f16_scale_cols);
  else if (f16_in)
    tc_gemm_nt_kernel<true><<<grid, TG_THREADS, TG_SMEM, st>>>(ta, tb, tcm, bias, C, ldc, M, N, K, stage_tx, k_splits, cps, epi,
                                                               f16_scale, f16_scale_cols);
  else
    tc_gemm_nt_kernel<false><<<grid, TG_THREADS, TG_SMEM, st>>>(ta, tb, tcm, bias, C, ldc, M, N, K, stage_tx, k_splits, cps, epi,
                                                                f16_scale, f16_scale_cols);
  NRMS_LAUNCH_CHECK("tc_gemm_nt");
  return NRMS_OK;
}

int tc_gemm_nt(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, float* C, int64_t ldc,
               int64_t M, int N, int K, cudaStream_t st) {
  return tc_gemm_nt_ex(A, lda, B, ldb, bias, C, ldc, M, N, K, 1, TC_EPI_STORE, st);
}

// number of K splits that fills the machine for an [M,N] output (weight gradients: few tiles, very long K)
int tc_gemm_auto_splits(int64_t M, int N, int K) {
  // row blocks (m tile, K split) are the unit a CTA takes: one wave of them, rounded down so that no CTA gets a second
  const int64_t m_tiles = (M + TG_BM - 1) / TG_BM;
  (void)N;
  int64_t s = num_sms() / m_tiles;
  const int chunks = (K + TG_BK - 1) / TG_BK;
  if (s > chunks / 8) s = chunks / 8;             // keep the main loop long enough to amortise the epilogue
  return s < 1 ? 1 : (int)s;
}

// dst[c][r] = src[r][c]  (src [R,C] row stride ld_src, dst [C,R] row stride ld_dst); 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) transpose_f32_kernel(const float* __restrict__ src, int64_t ld_src,
                                                            float* __restrict__ dst, int64_t ld_dst, int64_t R, int C) {
  __shared__ float tile[32][33];
  const int c_tiles = (C + 31) / 32;
  const int64_t r_tiles = (R + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
  for (int64_t t = blockIdx.x; t < r_tiles * c_tiles; t += gridDim.x) {
    const int64_t r0 = (t / c_tiles) * 32;
    const int c0 = (int)(t % c_tiles) * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int64_t r = r0 + ty + j;
      const int c = c0 + tx;
      tile[ty + j][tx] = (r < R && c < C) ? __ldg(src + r * ld_src + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int c = c0 + ty + j;
      const int64_t r = r0 + tx;
      if (c < C && r < R) dst[(int64_t)c * ld_dst + r] = tile[tx][ty + j];
    }
    __syncthreads();
  }
}

int transpose_f32(const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t R, int C, cudaStream_t st) {
  if (R <= 0 || C <= 0) return NRMS_OK;
  const int64_t tiles = ((R + 31) / 32) * ((C + 31) / 32);
  int64_t grid = (int64_t)num_sms() * 16;
  if (tiles < grid) grid = tiles;
  transpose_f32_kernel<<<(unsigned)grid, 256, 0, st>>>(src, ld_src, dst, ld_dst, R, C);
  NRMS_LAUNCH_CHECK("transpose_f32_kernel");
  return NRMS_OK;
}

}  // namespace nrms
