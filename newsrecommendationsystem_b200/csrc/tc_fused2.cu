// K1 v2 -- encoder_attn_tc_kernel: fused gather -> QKV projection -> multi-head attention, with BOTH the
// projections and the attention contractions (Q K^T and P V) on tcgen05 (kind::f16, fp32 accumulate).
//
// Tile = SPT sequences laid out in SLOT-row slots (news: 5 titles x 24-row slots, users: 2 histories x 64-row
// slots; rows >= S inside a slot are zero padding and masked out of the softmax), so a sequence never
// shares an 8-key swizzle piece with another one and (users) never straddles a warp.
//
//   warp 0  TMA producer: fp16 weight copy re-ordered per head ([Q_h;K_h;V_h] = 60 contiguous rows), one
//           192-row x 64-K box per chunk (3 heads per pass, 5 passes, 5 chunks each) through an smem ring.
//   warp 6  projection MMA issuer (one thread): 5 passes x 19 tcgen05.mma (128x192x16) per tile.
//   warp 1  attention MMA issuer (one thread): per head S = Q_h K_h^T (128x128x32), then O = P V_h (128x32x128).
//   warps 11-12  gather: fill the A tile of the next tile as soon as a_free fires (overlaps the attention).
//   warps 2-9  workers, thread == tile row == TMEM lane, two warps (roles) per lane quarter:
//           W1  tcgen05.ld q/k/v (+bias, q pre-scaled by log2(e)/sqrt(20)) -> fp16 -> Q and K operand tiles
//               (SWIZZLE_128B K-major) and V^T (keys along K) in shared memory;
//           W2  tcgen05.ld the row's own score block, e = 2^s for the valid keys (the reference's
//               exp(s)/(sum+1e-8) without max-subtraction), row sum in fp32, P (fp16, unnormalised) into
//               the block-diagonal A tile (off-block columns stay zero from kernel start);
//           W3  tcgen05.ld O (20 columns), scale by 1/(Z+1e-8), store the context row to C as fp16
//               (pitch 320 halfs; K2 v2 consumes it as the fp16 A operand and for the pooling).
// TMEM: two 192-column projection accumulators + one 128-column S/O region (O aliases S) = 512 columns.
#include <cuda.h>
#include <cuda_fp16.h>
#include <type_traits>
#include "tc_common.cuh"
#include "tc_api.cuh"

namespace nrms {
using namespace tc;

int make_tmap_k_major_f16(CUtensorMap* out, const void* base, int64_t rows, int cols, int64_t ld, int box_rows);

namespace k1v2 {

// debug trace (clock64 stamps of worker thread 0 / block 0): read back with nrms_debug_read_trace
__device__ long long g_trace[2048];
__device__ int g_trace_n;
#ifdef NRMS_K1_TRACE
// cheap trace: stamps go to shared memory (per traced thread), dumped to g_trace at kernel end
#define TRACE(tag)                                                                  \
  do {                                                                              \
    if (blockIdx.x == 0 && (wt == 0 || wt == 128) && trace_n < 250) {               \
      trace_buf[(wt ? 250 : 0) + trace_n++] =                                       \
          ((long long)((tag) + (wt ? 100 : 0)) << 48) | (clock64() & 0xFFFFFFFFFFFFLL); \
    }                                                                               \
  } while (0)
#else
#define TRACE(tag) do { } while (0)
#endif

constexpr int HP = 3, NPASS = 5, KCH = 5;       // heads per pass, passes, 64-half K chunks
constexpr int PN = 192;                         // projection UMMA N (180 real columns)
constexpr int NST = 2;                          // weight ring stages
constexpr int B_STAGE = PN * 128;               // 24,576
constexpr int W16_ROWS = 960, W16_LD = 320;
constexpr int CP = 320;                        // pitch (halfs) of the fp16 context rows handed to K2
constexpr int THREADS = 480;             // warps: 0 TMA, 1 attention MMA, 2-9 workers, 10 projection MMA, 11-14 gather
constexpr int OFF_A = 0;                        // 5 x [128 rows x 128 B]
constexpr int OFF_B = 5 * 16384;                // 81,920
constexpr int OFF_Q = OFF_B + NST * B_STAGE;    // 131,072
constexpr int OFF_K = OFF_Q + 16384;
constexpr int OFF_VT = OFF_K + 16384;           // 2 x [32 rows x 128 B]
constexpr int OFF_P = OFF_VT + 8192;            // 2 x [128 rows x 128 B]
constexpr int OFF_BIAS = OFF_P + 32768;         // 204,800
constexpr int OFF_IDX = OFF_BIAS + 3712;
constexpr int OFF_Z = OFF_IDX + 1024;         // partial row sums [2 roles][128 rows]
constexpr int OFF_BAR = OFF_Z + 1024;
constexpr int OFF_TRACE = OFF_BAR + 256;     // 500 x 8 B (debug builds only)
#ifdef NRMS_K1_TRACE
constexpr int SMEM = OFF_TRACE + 4096 + 1024;
#else
constexpr int SMEM = OFF_BAR + 256 + 1024;
#endif
static_assert(SMEM <= 232448, "shared memory budget");
constexpr int TM_S = 2 * PN;                    // TMEM column of the S / O region (384)
constexpr float QSCALE = 1.4426950408889634f / 4.47213595499957939f;   // log2(e) / sqrt(20)

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
__device__ __forceinline__ void workers_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// S = sequence length, SLOT = padded slot, SPT = sequences per tile
template <int S, int SLOT, int SPT>
__global__ void __launch_bounds__(THREADS, 1)
encoder_attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const float* __restrict__ src,
                       const void* __restrict__ idx, int idx_kind, int64_t n_seq, const float* __restrict__ bqkv,
                       __half* __restrict__ C) {
  static_assert(SLOT % 8 == 0 && SLOT >= S && SPT * SLOT <= 128, "slot layout");
  constexpr int NPAIR = SPT * S;                 // real rows per tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  float* bias_s = reinterpret_cast<float*>(sm + OFF_BIAS);
  int64_t* rowid = reinterpret_cast<int64_t*>(sm + OFF_IDX);
  const uint32_t bars = base + OFF_BAR;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * NST;
  const uint32_t a_full = bars + 16 * NST, a_free = a_full + 8, acc_full = a_full + 16, acc_empty = a_full + 32;
  const uint32_t qk_ready = a_full + 48, s_ready = a_full + 56, p_ready = a_full + 64, o_ready = a_full + 72;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(sm + OFF_BAR + 16 * NST + 96);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (n_seq + SPT - 1) / SPT;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full + 8 * s, 1);
      mbar_init(acc_empty + 8 * s, 8);
    }
    mbar_init(a_full, 128);
    mbar_init(a_free, 1);
    mbar_init(qk_ready, 256);
    mbar_init(s_ready, 1);
    mbar_init(p_ready, 256);
    mbar_init(o_ready, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_ptr_smem), 512);
  for (int i = tid; i < D3; i += THREADS) bias_s[i] = bqkv[i];
  // zero the A tile (padding rows / K tail stay zero) and all attention operand tiles (K padding of Q/K,
  // padded keys of V^T, off-block columns of P)
  for (int i = tid; i < OFF_B / 16; i += THREADS) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < (OFF_BIAS - OFF_Q) / 16; i += THREADS)
    reinterpret_cast<uint4*>(sm + OFF_Q)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------ TMA producer --------------------------------------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int p = 0; p < NPASS; ++p) {
          for (int kc = 0; kc < KCH; ++kc, ++it) {
            const int s = it % NST;
            mbar_wait(empty_bar + 8 * s, ((it / NST) & 1) ^ 1);
            expect_tx(full_bar + 8 * s, B_STAGE);
            tma_load_2d(base + OFF_B + s * B_STAGE, &tmap_w, kc * 64, 60 * HP * p, full_bar + 8 * s);
          }
        }
      }
    }
  } else if (warp == 10) {
    // ------------------------------ projection MMA issuer (one thread) ----------------------------
    // Independent of the attention MMAs (different TMEM columns / smem), so it gets its own issuing
    // thread: the latency-critical S / PV MMAs never wait behind descriptor building or a TMA wait.
    const uint32_t idesc_proj = umma_idesc_f16(128, PN);
    const uint64_t desc0 = umma_desc_k_sw128(0);
    uint32_t ring_it = 0, pass_it = 0, tile_it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_it) {
      mbar_wait(a_full, tile_it & 1);
      for (int p = 0; p < NPASS; ++p, ++pass_it) {
        const uint32_t as = pass_it & 1;
        mbar_wait(acc_empty + 8 * as, ((pass_it >> 1) & 1) ^ 1);
        for (int kc = 0; kc < KCH; ++kc, ++ring_it) {
          const int s = ring_it % NST;
          mbar_wait(full_bar + 8 * s, (ring_it / NST) & 1);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = (base + OFF_A + kc * 16384) >> 4;
            const uint32_t sb = (base + OFF_B + s * B_STAGE) >> 4;
            const int ksteps = (kc == KCH - 1) ? 3 : 4;
            for (int ks = 0; ks < ksteps; ++ks)
              umma_f16_ss(tmem_base + as * PN, desc0 | (uint64_t)((sa + 2 * ks) & 0x3FFF),
                          desc0 | (uint64_t)((sb + 2 * ks) & 0x3FFF), idesc_proj, (kc | ks) ? 1u : 0u);
            umma_commit(empty_bar + 8 * s);
            if (kc == KCH - 1) {
              umma_commit(acc_full + 8 * as);
              if (p == NPASS - 1) umma_commit(a_free);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ attention MMA issuer (one thread) -----------------------------
    const uint32_t idesc_s = umma_idesc_f16(128, 128);
    const uint32_t idesc_o = umma_idesc_f16(128, 32);
    const uint64_t desc0 = umma_desc_k_sw128(0);
    const uint32_t q_a = (base + OFF_Q) >> 4, k_a = (base + OFF_K) >> 4;
    const uint32_t p_a = (base + OFF_P) >> 4, v_a = (base + OFF_VT) >> 4;
    uint32_t head_it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int hd = 0; hd < H; ++hd, ++head_it) {
        const uint32_t hp = head_it & 1;
        mbar_wait(qk_ready, hp);                 // S = Q K^T
        tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            umma_f16_ss(tmem_base + TM_S, desc0 | (uint64_t)((q_a + 2 * ks) & 0x3FFF),
                        desc0 | (uint64_t)((k_a + 2 * ks) & 0x3FFF), idesc_s, ks ? 1u : 0u);
          umma_commit(s_ready);
        }
        __syncwarp();
        mbar_wait(p_ready, hp);                  // O = P V
        tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_f16_ss(tmem_base + TM_S, desc0 | (uint64_t)((p_a + (ks >> 2) * 1024 + (ks & 3) * 2) & 0x3FFF),
                        desc0 | (uint64_t)((v_a + (ks >> 2) * 256 + (ks & 3) * 2) & 0x3FFF), idesc_o, ks ? 1u : 0u);
          umma_commit(o_ready);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 11) {
    // ------------------------------ gather warps (11..14) --------------------------------------
    // Fill the A tile of the NEXT tile as soon as the projections of the current one have retired
    // (a_free), i.e. while the workers are still busy with its attention: NPAIR real rows x 75 float4,
    // 16 independent 16-byte loads in flight per thread, fp32 -> fp16, swizzled slot layout.
    const int gt = (warp - 11) * 32 + lane;      // 0..127
    uint32_t tile_it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_it) {
      const int64_t seq0 = t * SPT;
      for (int pr = gt; pr < NPAIR; pr += 128) {
        const int64_t seq = seq0 + pr / S;
        int64_t id = 0;
        if (seq < n_seq) {
          const int64_t e = seq * S + (pr % S);
          id = idx_kind == 0 ? e : (idx_kind == 1 ? reinterpret_cast<const int64_t*>(idx)[e]
                                                  : (int64_t) reinterpret_cast<const int32_t*>(idx)[e]);
        }
        rowid[pr] = id;
      }
      mbar_wait(a_free, (tile_it & 1) ^ 1);      // previous tile's projection MMAs no longer read A
      asm volatile("bar.sync 2, 128;" ::: "memory");
      constexpr int TOTAL4 = NPAIR * DV4;
#pragma unroll 1
      for (int f0 = 0; f0 < TOTAL4; f0 += 128 * 16) {
        float4 v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int f = f0 + u * 128 + gt;
          if (f < TOTAL4) {
            const int pr = f / DV4, c4 = f - pr * DV4;
            v[u] = __ldg(reinterpret_cast<const float4*>(src + rowid[pr] * D) + c4);
          }
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int f = f0 + u * 128 + gt;
          if (f < TOTAL4) {
            const int pr = f / DV4, c4 = f - pr * DV4;
            const int r = (pr / S) * SLOT + (pr % S);
            uint2 pk;
            pk.x = pack_h2(v[u].x, v[u].y);
            pk.y = pack_h2(v[u].z, v[u].w);
            *reinterpret_cast<uint2*>(sm + OFF_A + (c4 >> 4) * 16384 + r * 128 + ((((c4 & 15) >> 1) ^ (r & 7)) << 4) +
                                      ((c4 & 1) << 3)) = pk;
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(a_full);
      asm volatile("bar.sync 2, 128;" ::: "memory");  // rowid is rewritten for the next tile
    }
  } else if (warp >= 2 && warp <= 9) {
    // ------------------------------ workers (warps 2..9) ----------------------------------------
    // Two warps share every TMEM lane quarter (thread == tile row == key): role 0 (warps 2-5) handles q, k,
    // the first half of the row's score block and context columns 0..11; role 1 (warps 6-9) handles v, the
    // second half of the score block and context columns 12..19.
    const int role = (warp - 2) >> 2;
    const int q4 = warp & 3;
    const int wt = ((warp - 2) & 3) * 32 + lane + role * 128;   // trace id (0 = warp 2 lane 0)
    const int row = q4 * 32 + lane;
    const int sq = row / SLOT, pos = row - sq * SLOT;
    const bool row_valid = (sq < SPT) && (pos < S);
    const int sq_lo = (q4 * 32) / SLOT;          // warp-uniform first slot of this warp
    const int own = sq - sq_lo;                  // 0 or 1 (SLOT = 24), always 0 (SLOT = 64)
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    // score-block split between the roles (whole 8-key pieces): role 0 gets [0, C0), role 1 [C0, SLOT)
    constexpr int C0 = (SLOT >= 32) ? SLOT / 2 : 16;
    float* zpart = reinterpret_cast<float*>(sm + OFF_Z);
    // row-constant shared-memory addresses (swizzle XOR hoisted out of the head loop)
    const int sw = row & 7;
    uint8_t* const qrow = sm + OFF_Q + row * 128;
    uint8_t* const krow = sm + OFF_K + row * 128;
    const int o0 = (0 ^ sw) << 4, o1 = (1 ^ sw) << 4, o2 = (2 ^ sw) << 4;
    uint8_t* const vt_base = sm + OFF_VT + (row >> 6) * 4096 + ((row & 7) << 1);
    int vt_off[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) vt_off[i] = ((((row & 63) >> 3) ^ i) << 4);
    uint8_t* const prow = sm + OFF_P + row * 128;
#ifdef NRMS_K1_TRACE
    long long* trace_buf = reinterpret_cast<long long*>(sm + OFF_TRACE);
    int trace_n = 0;
#endif
    uint32_t pass_it = 0, head_it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int64_t seq0 = t * SPT;
      const bool st_ok = row_valid && (seq0 + sq < n_seq);
      __half* const crow0 = C + ((seq0 + sq) * S + pos) * CP + (role ? 12 : 0);
      TRACE(6);
      for (int p = 0; p < NPASS; ++p, ++pass_it) {
        const uint32_t as = pass_it & 1;
        mbar_wait(acc_full + 8 * as, (pass_it >> 1) & 1);
        tc_fence_after();
        TRACE(8);
        const uint32_t tacc = tmem_base + as * PN + lane_addr;
#pragma unroll 1
        for (int hh = 0; hh < HP; ++hh, ++head_it) {
          const int h = p * HP + hh;
          const uint32_t hp = head_it & 1;
          // ================= W1: q,k (role 0) / v (role 1) of this row -> fp16 operand tiles =========
          TRACE(1);
          if (role == 0) {
            uint32_t qk[40];
            tmem_ld16_nw(tacc + 60 * hh, qk);
            tmem_ld16_nw(tacc + 60 * hh + 16, qk + 16);
            tmem_ld8_nw(tacc + 60 * hh + 32, qk + 32);
            tmem_ld_wait();
            if (hh == HP - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(acc_empty + 8 * as);
            }
            uint32_t qp[10], kp[10];
#pragma unroll
            for (int d = 0; d < DH; d += 2) {
              qp[d >> 1] = pack_h2((__uint_as_float(qk[d]) + bias_s[h * DH + d]) * QSCALE,
                                   (__uint_as_float(qk[d + 1]) + bias_s[h * DH + d + 1]) * QSCALE);
              kp[d >> 1] = pack_h2(__uint_as_float(qk[DH + d]) + bias_s[D + h * DH + d],
                                   __uint_as_float(qk[DH + d + 1]) + bias_s[D + h * DH + d + 1]);
            }
            *reinterpret_cast<uint4*>(qrow + o0) = make_uint4(qp[0], qp[1], qp[2], qp[3]);
            *reinterpret_cast<uint4*>(qrow + o1) = make_uint4(qp[4], qp[5], qp[6], qp[7]);
            *reinterpret_cast<uint2*>(qrow + o2) = make_uint2(qp[8], qp[9]);
            *reinterpret_cast<uint4*>(krow + o0) = make_uint4(kp[0], kp[1], kp[2], kp[3]);
            *reinterpret_cast<uint4*>(krow + o1) = make_uint4(kp[4], kp[5], kp[6], kp[7]);
            *reinterpret_cast<uint2*>(krow + o2) = make_uint2(kp[8], kp[9]);
          } else {
            uint32_t vi[DH];
            tmem_ld16_nw(tacc + 60 * hh + 40, vi);
            tmem_ld4_nw(tacc + 60 * hh + 56, vi + 16);
            tmem_ld_wait();
            TRACE(11);
            if (hh == HP - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(acc_empty + 8 * as);
            }
            // V^T: element (d, key=row); padded / invalid keys contribute exact zeros
            const float vmask = row_valid ? 1.f : 0.f;
            // all bias loads first: the compiler must keep shared-memory loads behind earlier (possibly
            // aliasing) shared-memory stores, which would serialise load->convert->store per element
            __half hv[DH];
#pragma unroll
            for (int d = 0; d < DH; ++d)
              hv[d] = __float2half_rn((__uint_as_float(vi[d]) + bias_s[2 * D + h * DH + d]) * vmask);
#pragma unroll
            for (int d = 0; d < DH; ++d) *reinterpret_cast<__half*>(vt_base + d * 128 + vt_off[d & 7]) = hv[d];
            TRACE(12);
          }
          fence_proxy_async_smem();
          mbar_arrive(qk_ready);
          TRACE(2);
          // ================= W2: this role's half of the score block -> unnormalised probabilities =====
          mbar_wait(s_ready, hp);
          tc_fence_after();
          TRACE(3);
          {
            const int g0 = sq * (SLOT / 8);
            float Z = 0.f;
            auto half_block = [&](auto lo_c, auto n_c) {
              constexpr int LO = decltype(lo_c)::value, NC = decltype(n_c)::value;   // columns [LO, LO+NC) of the block
              uint32_t sv[(SLOT >= 32) ? NC : 2 * NC];
              if constexpr (SLOT >= 32) {
#pragma unroll
                for (int c = 0; c < NC; c += 16) tmem_ld16_nw(tmem_base + TM_S + lane_addr + sq_lo * SLOT + LO + c, sv + c);
              } else {
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                  if (sq_lo + b < SPT) {
                    if constexpr (NC == 16) tmem_ld16_nw(tmem_base + TM_S + lane_addr + (sq_lo + b) * SLOT + LO, sv + b * NC);
                    else tmem_ld8_nw(tmem_base + TM_S + lane_addr + (sq_lo + b) * SLOT + LO, sv + b * NC);
                  } else {
#pragma unroll
                    for (int j = 0; j < NC; ++j) sv[b * NC + j] = 0u;
                  }
                }
              }
              tmem_ld_wait();
              tc_fence_before();
              float e[NC];
#pragma unroll
              for (int j = 0; j < NC; ++j) {
                uint32_t x = sv[j];
                if (SLOT < 32) x = own ? sv[NC + j] : sv[j];
                e[j] = (LO + j < S) ? ex2(__uint_as_float(x)) : 0.f;
                Z += e[j];
              }
#pragma unroll
              for (int m = 0; m < NC / 8; ++m) {
                const int g = g0 + LO / 8 + m;
                const uint4 pk = make_uint4(pack_h2(e[8 * m], e[8 * m + 1]), pack_h2(e[8 * m + 2], e[8 * m + 3]),
                                            pack_h2(e[8 * m + 4], e[8 * m + 5]), pack_h2(e[8 * m + 6], e[8 * m + 7]));
                if (sq < SPT) *reinterpret_cast<uint4*>(prow + (g >> 3) * 16384 + (((g & 7) ^ (row & 7)) << 4)) = pk;
              }
            };
            if (role == 0) half_block(std::integral_constant<int, 0>{}, std::integral_constant<int, C0>{});
            else half_block(std::integral_constant<int, C0>{}, std::integral_constant<int, SLOT - C0>{});
            zpart[role * 128 + row] = Z;
          }
          fence_proxy_async_smem();
          mbar_arrive(p_ready);
          TRACE(4);
          // ================= W3: context row (columns 0..11 role 0, 12..19 role 1) =================
          mbar_wait(o_ready, hp);
          tc_fence_after();
          TRACE(5);
          {
            const float inv = 1.f / (zpart[row] + zpart[128 + row] + 1e-8f);
            TRACE(13);
            const bool st = st_ok;
            __half* crow = crow0 + h * DH;
            if (role == 0) {
              uint32_t o[12];
              tmem_ld8_nw(tmem_base + TM_S + lane_addr, o);
              tmem_ld4_nw(tmem_base + TM_S + lane_addr + 8, o + 8);
              tmem_ld_wait();
              TRACE(14);
              tc_fence_before();
              if (st) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                  reinterpret_cast<uint2*>(crow)[c] =
                      make_uint2(pack_h2(__uint_as_float(o[4 * c]) * inv, __uint_as_float(o[4 * c + 1]) * inv),
                                 pack_h2(__uint_as_float(o[4 * c + 2]) * inv, __uint_as_float(o[4 * c + 3]) * inv));
              }
              TRACE(15);
            } else {
              uint32_t o[8];
              tmem_ld8_nw(tmem_base + TM_S + lane_addr + 12, o);
              tmem_ld_wait();
              tc_fence_before();
              if (st) {
#pragma unroll
                for (int c = 0; c < 2; ++c)
                  reinterpret_cast<uint2*>(crow)[c] =
                      make_uint2(pack_h2(__uint_as_float(o[4 * c]) * inv, __uint_as_float(o[4 * c + 1]) * inv),
                                 pack_h2(__uint_as_float(o[4 * c + 2]) * inv, __uint_as_float(o[4 * c + 3]) * inv));
                if (h == H - 1) {          // zero the K padding (columns 300..319) of this context row once
#pragma unroll
                  for (int c = 0; c < 5; ++c) reinterpret_cast<uint2*>(crow0 - 12 + D)[c] = make_uint2(0u, 0u);
                }
              }
            }
          }
        }
      }
    }
#ifdef NRMS_K1_TRACE
    if (blockIdx.x == 0 && (wt == 0 || wt == 128)) {
      for (int i = 0; i < trace_n; ++i) g_trace[(wt ? 250 : 0) + i] = trace_buf[(wt ? 250 : 0) + i];
      if (wt == 0) g_trace_n = trace_n; 
    }
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// fp16 weight copy, re-ordered per head: row 60*h + {0..19 | 20..39 | 40..59} = {W_Q, W_K, W_V}[20*h + ..]
__global__ void __launch_bounds__(256) pack_w16_kernel(const float* __restrict__ w, __half* __restrict__ out) {
  const int n = W16_ROWS * W16_LD;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / W16_LD, k = i - r * W16_LD;
    float x = 0.f;
    if (r < 3 * D && k < D) {
      const int h = r / 60, j = r - 60 * h;
      const int src_row = (j / DH) * D + h * DH + (j % DH);
      x = w[src_row * D + k];
    }
    out[i] = __float2half_rn(x);
  }
}
constexpr size_t W16_BYTES = (size_t)W16_ROWS * W16_LD * 2;   // 614,400

}  // namespace k1v2

size_t k1v2_w16_bytes() { return k1v2::W16_BYTES; }

// debug builds (-DNRMS_K1_TRACE): host[0..250) = role-0 stamps, host[250..500) = role-1 stamps; returns count per role
extern "C" int nrms_debug_read_trace(long long* host, int max_n) {
  int n = 0;
  cudaMemcpyFromSymbol(&n, k1v2::g_trace_n, sizeof(int));
  if (max_n < 500) return -1;
  cudaMemcpyFromSymbol(host, k1v2::g_trace, 500 * sizeof(long long));
  return n;
}

template <int S, int SLOT, int SPT>
static int launch_k1v2(const CUtensorMap& tw, const float* src, const void* idx, int idx_kind, int64_t n,
                       const float* bqkv, void* Cbuf, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k1v2::encoder_attn_tc_kernel<S, SLOT, SPT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, k1v2::SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(encoder_attn_tc_kernel)");
    configured = true;
  }
  const int64_t tiles = (n + SPT - 1) / SPT;
  int grid = num_sms();
  if (tiles < grid) grid = (int)tiles;
  k1v2::encoder_attn_tc_kernel<S, SLOT, SPT><<<grid, k1v2::THREADS, k1v2::SMEM, st>>>(
      tw, src, idx, idx_kind, n, bqkv, reinterpret_cast<__half*>(Cbuf));
  NRMS_LAUNCH_CHECK("encoder_attn_tc_kernel");
  return NRMS_OK;
}

// prepares the fp16 weight copy (once per encoder call) and returns its tensor map
int k1v2_prepare(const float* wqkv, void* w16, CUtensorMap* tw, cudaStream_t st) {
  k1v2::pack_w16_kernel<<<148, 256, 0, st>>>(wqkv, reinterpret_cast<__half*>(w16));
  NRMS_LAUNCH_CHECK("pack_w16_kernel");
  return make_tmap_k_major_f16(tw, w16, k1v2::W16_ROWS, k1v2::W16_LD, k1v2::W16_LD, k1v2::PN);
}

int k1v2_run(int S, const CUtensorMap& tw, const float* src, const void* idx, int idx_kind, int64_t n,
             const float* bqkv, void* Cbuf, cudaStream_t st) {
  if (S == 20) return launch_k1v2<20, 24, 5>(tw, src, idx, idx_kind, n, bqkv, Cbuf, st);
  if (S == 50) return launch_k1v2<50, 64, 2>(tw, src, idx, idx_kind, n, bqkv, Cbuf, st);
  set_error("encoder_attn_tc_kernel compiled for S = 20 or 50, got %d", S);
  return NRMS_E_UNSUPPORTED;
}

}  // namespace nrms
