// K1 v4 -- encoder_attn_tc4_kernel: gather -> QKV projection -> multi-head attention, all contractions on tcgen05,
// with the scalar (CUDA-core) work of the earlier generations cut to what cannot be avoided.
//
//   * Gather by the TMA unit: the input rows come from an fp16 copy of the source table ([rows+1, 320] halfs: 300
//     values, a constant 1.0 in column 300, zeros after it; the last row is an all-zero "null" row) and land in the
//     resident A tile (UMMA SWIZZLE_128B K-major, 5 chunks of 64 halfs) through cp.async.bulk.tensor ... tile::gather4:
//     one instruction moves four 128-byte row pieces, each lane of the gather warp owns one 4-row group.  The A tile
//     is a 5-slot ring (one full/free mbarrier pair per K chunk), so the next tile's rows stream in while the last
//     projection pass of the current tile is still running.  No thread touches the gathered data.
//   * Bias and the 1/sqrt(20) scale are folded into the GEMM: column 300 of the fp16 weight copy holds the bias and
//     meets the 1.0 column of A; the W_Q rows (and b_Q) are pre-multiplied by log2(e)/sqrt(20).  Padding rows of the
//     tile are all-zero rows of A, so their q, k and v are exact zeros without any masking.
//   * S = Q K^T is a TS-form kind::tf32 MMA whose A operand is the fp32 projection accumulator itself (tensor
//     memory): q is never read by a thread.  K goes to shared memory as tf32 (cvt.rna, 5 x 16-byte stores per row),
//     V^T as fp16.  The unnormalised probabilities P = 2^s live in tensor memory as packed fp16 and feed O = P V as
//     the A operand of a TS-form kind::f16 MMA.
//   * Two heads per projection pass (8 passes, UMMA N = 128) and two independent head sets in flight, so the
//     tensor-core round trips of one head hide behind the worker phases of the other.
//
// Warps: 0 weight TMA producer, 1 attention MMA issuer (+ TMEM allocator), 2-17 workers (thread == tile row == TMEM
// lane; per lane quarter one warp for each (head set, role)), 18 projection MMA issuer, 19 A-tile gather.
// TMEM (512 columns): projection accumulator [0,128) | set s: S/O at 128+192s (128 cols), P at +128 (64 cols).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <type_traits>
#include "tc_common.cuh"
#include "tc_api.cuh"

namespace nrms {
using namespace tc;

int make_tmap_k_major_f16(CUtensorMap* out, const void* base, int64_t rows, int cols, int64_t ld, int box_rows);
int make_tmap_store_f16(CUtensorMap* out, const void* base, int64_t rows, int cols, int64_t ld, int box_cols, int box_rows);

namespace k1v4 {

// debug trace (-DNRMS_K1_TRACE): clock64 stamps of lane 0 of four warps of block 0 (0 worker role 0, 1 worker
// role 1, 2 projection issuer, 3 attention issuer), written straight to global memory; nrms_debug_read_trace4
__device__ long long g_trace4[4][1024];
__device__ int g_trace4_n[4];
#ifdef NRMS_K1_TRACE
#define TRACE(who, tag)                                                                         \
  do {                                                                                          \
    if (blockIdx.x == 0 && lane == 0 && trace_n < 1024)                                         \
      g_trace4[who][trace_n++] = ((long long)(tag) << 48) | (clock64() & 0xFFFFFFFFFFFFLL);     \
  } while (0)
#define TRACE_DECL int trace_n = 0
#define TRACE_END(who) do { if (blockIdx.x == 0 && lane == 0) g_trace4_n[who] = trace_n; } while (0)
#else
#define TRACE(who, tag) do { } while (0)
#define TRACE_DECL do { } while (0)
#define TRACE_END(who) do { } while (0)
#endif

constexpr int NPASS = 8, KCH = 5;
constexpr int PN = 128;                         // projection UMMA N (120 real columns per pass)
constexpr int NST = 5;                          // weight ring stages (= one whole projection pass)
static_assert(NST == KCH, "the projection issuer relies on stage == K chunk");
constexpr int B_STAGE = PN * 128;               // 16,384
constexpr int W16_ROWS = 1024, W16_LD = 320;
constexpr int CP = 320;                         // pitch (halfs) of the fp16 context rows handed to K2
constexpr int SRC_LD = 320;                     // pitch (halfs) of the fp16 gather source
constexpr int THREADS = 640;                    // 20 warps: 4 service warps + 16 workers
constexpr int OFF_A = 0;                        // 5 x [128 rows x 128 B]
constexpr int OFF_B = KCH * 16384;              // 81,920
constexpr int OFF_SET = OFF_B + NST * B_STAGE;  // 163,840 ; per set: K (tf32) 16 KB | V^T (fp16) 8 KB
constexpr int SET_BYTES = 16384 + 8192;
constexpr int OFF_Z = OFF_SET + 2 * SET_BYTES;  // 212,992 ; partial row sums [2 sets][2 roles][128]
constexpr int OFF_STG = OFF_Z + 2048;           // context staging for the TMA store: [128 rows][2 heads x 20 halfs]
constexpr int STG_BYTES = 128 * 80;             // 10,240
constexpr int OFF_BAR = OFF_STG + STG_BYTES;
constexpr int SMEM = OFF_BAR + 512 + 1024;
static_assert(SMEM <= 232448, "shared memory budget");
constexpr int TM_SET = 128, TM_SET_STRIDE = 192, TM_P = 128;
// log2(e)/sqrt(20), times (1 + 2^-11): the S MMA reads q as tf32 by TRUNCATING the fp32 accumulator (mean relative
// error -2^-11); the pre-scale centres that error like a round-to-nearest would.
constexpr float QSCALE = 1.4426950408889634f / 4.47213595499957939f * (1.f + 1.f / 2048.f);

__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// four rows r0..r3 of the 2-D tensor (box = 64 halfs x 1 row) -> four consecutive 128-byte rows at dst_smem
__device__ __forceinline__ void tma_gather4(uint32_t dst_smem, const CUtensorMap* tmap, int col, int r0, int r1, int r2,
                                            int r3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}
// shared-memory box [rows][40 halfs] -> global tensor at {col, row}; completion tracked by bulk groups
__device__ __forceinline__ void tma_store_2d(uint32_t src_smem, const CUtensorMap* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(src_smem)
               : "memory");
}
__device__ __forceinline__ void expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// fp32 bits -> tf32 bits, round to nearest with ties away from zero (what cvt.rna.tf32.f32 computes for finite
// values), done on the integer pipe: cvt.rna runs on the quarter-rate conversion unit (measured 16 thread-ops per
// clock per SM, shared with ex2), which the exp2 of the softmax already saturates.
__device__ __forceinline__ uint32_t rna_tf32(uint32_t x) { return (x + 0x1000u) & 0xFFFFE000u; }
// D[tmem] (+)= A[tmem, fp32 read as tf32] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// S = sequence length, SLOT = padded slot (rows of the tile per sequence), SPT = sequences per tile
template <int S, int SLOT, int SPT>
__global__ void __launch_bounds__(THREADS, 1)
encoder_attn_tc4_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_src,
                        const __grid_constant__ CUtensorMap tmap_c, const void* __restrict__ idx, int idx_kind,
                        int64_t n_seq, int null_row) {
  static_assert(SLOT % 8 == 0 && SLOT >= S && SPT * SLOT <= 128, "slot layout");
  constexpr int GPS = (S + 3) / 4;               // 4-row gather groups per sequence
  static_assert(GPS * SPT <= 32 && GPS * 4 <= SLOT, "one gather group per lane");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t bars = base + OFF_BAR;
  const uint32_t w_full = bars, w_empty = bars + 8 * NST;                 // [NST] each
  const uint32_t a_full = bars + 16 * NST, a_free = a_full + 8 * KCH;      // [KCH] each
  const uint32_t acc_full = a_free + 8 * KCH, acc_empty = acc_full + 8;
  const uint32_t kv_ready = acc_empty + 8, s_ready = kv_ready + 16, p_ready = s_ready + 16, o_ready = p_ready + 16;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(sm + OFF_BAR + 16 * NST + 16 * KCH + 96);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler: role branches stay uniform
  const int64_t n_tiles = (n_seq + SPT - 1) / SPT;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(w_full + 8 * s, 1);
      mbar_init(w_empty + 8 * s, 1);
    }
    for (int k = 0; k < KCH; ++k) {
      mbar_init(a_full + 8 * k, 1);
      mbar_init(a_free + 8 * k, 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 17);         // 16 worker warps (k, v drained) + the attention issuer (q consumed by both S MMAs)
    for (int s = 0; s < 2; ++s) {
      mbar_init(kv_ready + 8 * s, 8);
      mbar_init(s_ready + 8 * s, 1);
      mbar_init(p_ready + 8 * s, 8);
      mbar_init(o_ready + 8 * s, 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_ptr_smem), 512);
  // zero the A tile (padding rows are never written by the gather) and both operand sets (K columns 20..31,
  // padded keys of V^T)
  for (int i = tid; i < OFF_B / 16; i += THREADS) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < (2 * SET_BYTES) / 16; i += THREADS)
    reinterpret_cast<uint4*>(sm + OFF_SET)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------ weight TMA producer: one 128-row box per K chunk ----------
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
#pragma unroll 1
        for (int p = 0; p < NPASS; ++p) {
#pragma unroll 1
          for (int kc = 0; kc < KCH; ++kc, ++it) {
            const int s = it % NST;
            mbar_wait(w_empty + 8 * s, ((it / NST) & 1) ^ 1);
#ifdef NRMS_DBG_NOWLOAD
            if (it >= NST) { mbar_arrive(w_full + 8 * s); continue; }     // timing experiment: stale weights
#endif
            expect_tx(w_full + 8 * s, B_STAGE);
            tma_load_2d(base + OFF_B + s * B_STAGE, &tmap_w, kc * 64, PN * p, w_full + 8 * s);
          }
        }
      }
    }
  } else if (warp == 19) {
    // ------------------------------ A-tile gather (TMA gather4, one 4-row group per lane) ------
    const int sq = lane / GPS, g = lane - sq * GPS;
    const uint32_t dst0 = base + OFF_A + (uint32_t)(sq * SLOT + 4 * g) * 128u;
    uint32_t tile_it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_it) {
      const int64_t seq0 = t * SPT;
      const int n_here = (n_seq - seq0 < SPT) ? (int)(n_seq - seq0) : SPT;
      const bool active = sq < n_here;
      int r[4] = {null_row, null_row, null_row, null_row};
      if (active) {
        const int64_t e0 = (seq0 + sq) * S + 4 * g;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (4 * g + j < S) {
            const int64_t e = e0 + j;
            r[j] = idx_kind == 0 ? (int)e : (idx_kind == 1 ? (int)reinterpret_cast<const int64_t*>(idx)[e]
                                                          : reinterpret_cast<const int32_t*>(idx)[e]);
          }
        }
      }
      const uint32_t tx = (uint32_t)(n_here * GPS) * 512u;
#pragma unroll 1
      for (int kc = 0; kc < KCH; ++kc) {
        mbar_wait(a_free + 8 * kc, (tile_it & 1) ^ 1);     // the previous tile's last pass is done with this chunk
        if (lane == 0) expect_tx(a_full + 8 * kc, tx);
        __syncwarp();
        if (active) tma_gather4(dst0 + kc * 16384, &tmap_src, kc * 64, r[0], r[1], r[2], r[3], a_full + 8 * kc);
      }
    }
  } else if (warp == 18) {
    // ------------------------------ projection MMA issuer (whole warp converged, see umma_*_p) ---
    const uint32_t idesc_proj = umma_idesc_f16(128, PN);
    const uint64_t desc0 = umma_desc_k_sw128(0);
    const uint32_t el = elect_one_u32();
    uint32_t pass_it = 0, tile_it = 0;
    TRACE_DECL;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_it) {
#pragma unroll 1
      for (int p = 0; p < NPASS; ++p, ++pass_it) {
        TRACE(2, 40);
        // NST == KCH: chunk kc of every pass lives in ring stage kc, and the whole pass is prefetched while the
        // previous one is being consumed -- so all operand waits come first, off the critical path
#pragma unroll
        for (int kc = 0; kc < KCH; ++kc) {
          if (p == 0) mbar_wait(a_full + 8 * kc, tile_it & 1);
          mbar_wait(w_full + 8 * kc, pass_it & 1);
        }
        TRACE(2, 47);
        mbar_wait(acc_empty, (pass_it & 1) ^ 1);       // previous pass: k, v drained and q consumed
        tc_fence_after();
        TRACE(2, 41);
#pragma unroll
        for (int kc = 0; kc < KCH; ++kc) {
          const uint32_t sa = (base + OFF_A + kc * 16384) >> 4;
          const uint32_t sb = (base + OFF_B + kc * B_STAGE) >> 4;
          const int ksteps = (kc == KCH - 1) ? 3 : 4;      // columns 256..303 (the bias column is 300)
#pragma unroll
          for (int ks = 0; ks < ksteps; ++ks)
            umma_f16_ss_p(tmem_base, desc0 | (uint64_t)((sa + 2 * ks) & 0x3FFF), desc0 | (uint64_t)((sb + 2 * ks) & 0x3FFF),
                          idesc_proj, (kc | ks) ? 1u : 0u, el);
          umma_commit_p(w_empty + 8 * kc, el);
          if (p == NPASS - 1) umma_commit_p(a_free + 8 * kc, el);
        }
        umma_commit_p(acc_full, el);
        TRACE(2, 48);
      }
    }
    TRACE_END(2);
  } else if (warp == 1) {
    // ------------------------------ attention MMA issuer (whole warp converged) ------------------
    // The workers publish in the fixed order kv(0), kv(1), p(0), p(1) every pass, so a static wait order works.
    const uint32_t idesc_s = umma_idesc_tf32(128, 128);
    const uint32_t idesc_o = umma_idesc_f16(128, 32);
    const uint64_t desc0 = umma_desc_k_sw128(0);
    const uint32_t el = elect_one_u32();
    uint32_t pass_it = 0;
    TRACE_DECL;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
#pragma unroll 1
      for (int p = 0; p < NPASS; ++p, ++pass_it) {
        const uint32_t ph = pass_it & 1;
        TRACE(3, 50);
#pragma unroll
        for (int set = 0; set < 2; ++set) {           // S = Q K^T   (A = q columns of the projection accumulator)
          mbar_wait(kv_ready + 8 * set, ph);
          tc_fence_after();
          TRACE(3, 51 + set);
          const uint32_t k_a = (base + OFF_SET + set * SET_BYTES) >> 4;
#pragma unroll
          for (int ks = 0; ks < 3; ++ks)
            umma_tf32_ts_p(tmem_base + TM_SET + TM_SET_STRIDE * set, tmem_base + 60 * set + 8 * ks,
                           desc0 | (uint64_t)((k_a + 2 * ks) & 0x3FFF), idesc_s, ks ? 1u : 0u, el);
          umma_commit_p(s_ready + 8 * set, el);
          if (set == 1) umma_commit_p(acc_empty, el);
          TRACE(3, 55);
        }
#pragma unroll
        for (int set = 0; set < 2; ++set) {           // O = P V   (A = P from tensor memory)
          mbar_wait(p_ready + 8 * set, ph);
          tc_fence_after();
          TRACE(3, 53 + set);
          const uint32_t v_a = (base + OFF_SET + set * SET_BYTES + 16384) >> 4;
          const uint32_t tset = tmem_base + TM_SET + TM_SET_STRIDE * set;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_f16_ts_p(tset, tset + TM_P + 8 * ks, desc0 | (uint64_t)((v_a + (ks >> 2) * 256 + (ks & 3) * 2) & 0x3FFF),
                          idesc_o, ks ? 1u : 0u, el);
          umma_commit_p(o_ready + 8 * set, el);
          TRACE(3, 56);
        }
      }
    }
    TRACE_END(3);
  } else if (warp >= 2 && warp <= 17) {
    // ------------------------------ workers (warps 2..17) ---------------------------------------
    // Four groups of four warps (one warp per TMEM lane quarter): group = (set, role).  A worker serves ONE head of
    // the pass (its set), so the two heads proceed side by side:
    //   role 0: k -> shared memory, first half of the score block, context columns 0..11
    //   role 1: v -> shared memory, second half of the score block, context columns 12..19
    const int grp = (warp - 2) >> 2;
    const int role = grp & 1, set = grp >> 1;
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const int sq = row / SLOT;
    const int sq_lo = (q4 * 32) / SLOT;
    const int own = sq - sq_lo;
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    constexpr int C0 = (SLOT >= 32) ? SLOT / 2 : 16;     // role 0 handles block columns [0,C0), role 1 [C0,SLOT)
    float* zpart = reinterpret_cast<float*>(sm + OFF_Z);
    uint8_t* const setp = sm + OFF_SET + set * SET_BYTES;
    const uint32_t tacc = tmem_base + lane_addr + 60 * set;
    const uint32_t tS = tmem_base + TM_SET + TM_SET_STRIDE * set + lane_addr;
    const int sw = row & 7;
    const int vt_row_off = (row >> 6) * 4096 + ((row & 7) << 1);
    int vt_off[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) vt_off[i] = ((((row & 63) >> 3) ^ i) << 4);
    // zero the P region of the set once (off-block columns must stay zero): each role clears one half
    {
      uint32_t z[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
      for (int c = 0; c < 32; c += 16) tmem_st16(tS + TM_P + 32 * role + c, z);
      tmem_st_wait();
      tc_fence_before();
      asm volatile("bar.sync 1, 512;" ::: "memory");
      tc_fence_after();
    }
    uint32_t pass_it = 0;
#ifdef NRMS_K1_TRACE
    int trace_n = 0;
    const bool tracer = (warp == 2 || warp == 6);
#define WTRACE(tag) do { if (tracer) TRACE(role, tag); } while (0)
#else
#define WTRACE(tag) do { } while (0)
#endif
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int64_t seq0 = t * SPT;
      const int n_here = (n_seq - seq0 < SPT) ? (int)(n_seq - seq0) : SPT;
#pragma unroll 1
      for (int p = 0; p < NPASS; ++p, ++pass_it) {
        const uint32_t ph = pass_it & 1;
        WTRACE(20);
        mbar_wait(acc_full, ph);
        tc_fence_after();
        WTRACE(21);
        // ================= W1: k (role 0) / v (role 1) of the head -> operand tiles =========
        {
          uint32_t x[DH];
          tmem_ld16_nw(tacc + (role ? 40 : 20), x);
          tmem_ld4_nw(tacc + (role ? 56 : 36), x + 16);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
          if (role == 0) {
            uint8_t* const krow = setp + row * 128;
#pragma unroll
            for (int c = 0; c < 5; ++c)
              *reinterpret_cast<uint4*>(krow + ((c ^ sw) << 4)) =
                  make_uint4(rna_tf32(x[4 * c]), rna_tf32(x[4 * c + 1]), rna_tf32(x[4 * c + 2]), rna_tf32(x[4 * c + 3]));
          } else {
            // packed conversions (cvt.rn.f16x2.f32 is a full-rate ALU instruction, the scalar cvt.rn.f16.f32 is not)
            uint8_t* const vt_base = setp + 16384 + vt_row_off;
#pragma unroll
            for (int d = 0; d < DH; d += 2) {
              const uint32_t h2 = pack_h2(__uint_as_float(x[d]), __uint_as_float(x[d + 1]));
              *reinterpret_cast<uint16_t*>(vt_base + d * 128 + vt_off[d & 7]) = (uint16_t)(h2 & 0xFFFFu);
              *reinterpret_cast<uint16_t*>(vt_base + (d + 1) * 128 + vt_off[(d + 1) & 7]) = (uint16_t)(h2 >> 16);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(kv_ready + 8 * set);
          WTRACE(22);
        }
        // ================= W2: score block -> P (tensor memory) =================
        {
          mbar_wait(s_ready + 8 * set, ph);
          tc_fence_after();
          WTRACE(24);
          float Z = 0.f;
          auto half_block = [&](auto lo_c, auto n_c) {
            constexpr int LO = decltype(lo_c)::value, NC = decltype(n_c)::value;
            if constexpr (SLOT >= 32) {
              uint32_t sv[NC];
#pragma unroll
              for (int c = 0; c < NC; c += 16) tmem_ld16_nw(tS + sq_lo * SLOT + LO + c, sv + c);
              tmem_ld_wait();
              uint32_t pk[NC / 2];
#pragma unroll
              for (int j = 0; j < NC; j += 2) {
                const float e0 = (LO + j < S) ? ex2(__uint_as_float(sv[j])) : 0.f;
                const float e1 = (LO + j + 1 < S) ? ex2(__uint_as_float(sv[j + 1])) : 0.f;
                Z += e0 + e1;
                pk[j >> 1] = pack_h2(e0, e1);
              }
              // keys sq*SLOT + LO + j  ->  P columns (sq*SLOT + LO)/2 + j/2   (warp-uniform: one slot per warp)
#pragma unroll
              for (int c = 0; c < NC / 2; c += 16) tmem_st16(tS + TM_P + (sq_lo * SLOT + LO) / 2 + c, pk + c);
            } else {
              // SLOT == 24: the warp's two candidate slots; lanes that do not own a slot store zeros there
              uint32_t sv[2 * NC];
#pragma unroll
              for (int b = 0; b < 2; ++b) {
                if (sq_lo + b < SPT) {
                  if constexpr (NC == 16) tmem_ld16_nw(tS + (sq_lo + b) * SLOT + LO, sv + b * NC);
                  else tmem_ld8_nw(tS + (sq_lo + b) * SLOT + LO, sv + b * NC);
                } else {
#pragma unroll
                  for (int j = 0; j < NC; ++j) sv[b * NC + j] = 0u;
                }
              }
              tmem_ld_wait();
#pragma unroll
              for (int b = 0; b < 2; ++b) {
                uint32_t pk[NC / 2];
                const bool mine = (own == b);
#pragma unroll
                for (int j = 0; j < NC; j += 2) {
                  const float e0 = (mine && LO + j < S) ? ex2(__uint_as_float(sv[b * NC + j])) : 0.f;
                  const float e1 = (mine && LO + j + 1 < S) ? ex2(__uint_as_float(sv[b * NC + j + 1])) : 0.f;
                  Z += e0 + e1;
                  pk[j >> 1] = pack_h2(e0, e1);
                }
                if (sq_lo + b < SPT) {
                  if constexpr (NC == 16) tmem_st8(tS + TM_P + ((sq_lo + b) * SLOT + LO) / 2, pk);
                  else tmem_st4(tS + TM_P + ((sq_lo + b) * SLOT + LO) / 2, pk);
                }
              }
            }
          };
          if (role == 0) half_block(std::integral_constant<int, 0>{}, std::integral_constant<int, C0>{});
          else half_block(std::integral_constant<int, C0>{}, std::integral_constant<int, SLOT - C0>{});
          zpart[(set * 2 + role) * 128 + row] = Z;
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_ready + 8 * set);
          WTRACE(26);
        }
        // ================= W3: context rows -> staging tile -> TMA store =================
        // The 40 context columns of the pass (heads 2p, 2p+1; the dummy 16th head gives the zero K padding
        // 300..319) are staged as [128 rows][80 B] and leave the SM as one 2-D bulk tensor store per sequence.
        // The bulk store of the previous pass must have read the staging tile before it is overwritten; its issuer
        // checks that here (long done by now) rather than stalling right after the issue.
        if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync 1, 512;" ::: "memory");
        {
          mbar_wait(o_ready + 8 * set, ph);
          tc_fence_after();
          WTRACE(28);
          const float inv = 1.f / (zpart[(set * 2) * 128 + row] + zpart[(set * 2 + 1) * 128 + row] + 1e-8f);
          uint8_t* const srow = sm + OFF_STG + row * 80 + set * 40 + (role ? 24 : 0);
          if (role == 0) {
            uint32_t o[12];
            tmem_ld8_nw(tS, o);
            tmem_ld4_nw(tS + 8, o + 8);
            tmem_ld_wait();
            tc_fence_before();
#pragma unroll
            for (int c = 0; c < 3; ++c)
              reinterpret_cast<uint2*>(srow)[c] =
                  make_uint2(pack_h2(__uint_as_float(o[4 * c]) * inv, __uint_as_float(o[4 * c + 1]) * inv),
                             pack_h2(__uint_as_float(o[4 * c + 2]) * inv, __uint_as_float(o[4 * c + 3]) * inv));
          } else {
            uint32_t o[8];
            tmem_ld8_nw(tS + 12, o);
            tmem_ld_wait();
            tc_fence_before();
#pragma unroll
            for (int c = 0; c < 2; ++c)
              reinterpret_cast<uint2*>(srow)[c] =
                  make_uint2(pack_h2(__uint_as_float(o[4 * c]) * inv, __uint_as_float(o[4 * c + 1]) * inv),
                             pack_h2(__uint_as_float(o[4 * c + 2]) * inv, __uint_as_float(o[4 * c + 3]) * inv));
          }
        }
        WTRACE(30);
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 512;" ::: "memory");     // staging tile complete
        WTRACE(31);
        if (warp == 2 && lane == 0) {
          for (int u = 0; u < n_here; ++u)
            tma_store_2d(base + OFF_STG + u * SLOT * 80, &tmap_c, 40 * p, (int)((seq0 + u) * S));
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#ifdef NRMS_K1_TRACE
    if (tracer) TRACE_END(role);
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// fp16 weight copy, two heads per 128-row pass block: row 128*p + 60*s + {0..19 | 20..39 | 40..59} =
// {W_Q, W_K, W_V}[20*(2p+s) + ..]; column 300 = the bias; W_Q rows and b_Q carry QSCALE.  Rows of the dummy 16th
// head and the last 8 rows of every block are zero.
__global__ void __launch_bounds__(256) pack_w16_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                                        __half* __restrict__ out) {
  const int n = W16_ROWS * W16_LD;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / W16_LD, k = i - r * W16_LD;
    const int p = r >> 7, j = r & 127;
    float x = 0.f;
    if (j < 120 && k <= D) {
      const int h = 2 * p + j / 60, jj = j % 60;
      if (h < H) {
        const int src_row = (jj / DH) * D + h * DH + (jj % DH);
        x = (k < D) ? w[src_row * D + k] : b[src_row];
        if (jj < DH) x *= QSCALE;
      }
    }
    out[i] = __float2half_rn(x);
  }
}

// fp16 gather source: dst [n_rows + 1][320] = {fp16(src[r][0..299]), 1.0, 0 x 19}; row n_rows (the null row) = 0.
// One thread per 8 output halfs (16-byte store).
__global__ void __launch_bounds__(256) pack_src16_kernel(const float* __restrict__ src, int64_t n_rows,
                                                          __half* __restrict__ dst) {
  const int64_t total = (n_rows + 1) * 40;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / 40;
    const int c = (int)(i - r * 40);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (r < n_rows) {
      const float4* s = reinterpret_cast<const float4*>(src + r * D) + 2 * c;
      if (c < 37) {
        const float4 a = __ldg(s), bq = __ldg(s + 1);
        o = make_uint4(pack_h2(a.x, a.y), pack_h2(a.z, a.w), pack_h2(bq.x, bq.y), pack_h2(bq.z, bq.w));
      } else if (c == 37) {
        const float4 a = __ldg(s);
        o = make_uint4(pack_h2(a.x, a.y), pack_h2(a.z, a.w), pack_h2(1.f, 0.f), 0u);
      }
    }
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
}

}  // namespace k1v4

// debug builds (-DNRMS_K1_TRACE): host[who*1024 ..] = stamps of tracer `who`; counts[4]
extern "C" int nrms_debug_read_trace4(long long* host, int* counts) {
  cudaMemcpyFromSymbol(counts, k1v4::g_trace4_n, 4 * sizeof(int));
  cudaMemcpyFromSymbol(host, k1v4::g_trace4, 4 * 1024 * sizeof(long long));
  return 0;
}

size_t k1v4_src16_bytes(int64_t n_rows) { return (size_t)(n_rows + 1) * k1v4::SRC_LD * 2; }

// fp16 weight copy (once per encoder call) + its tensor map
int k1v4_prepare(const float* wqkv, const float* bqkv, void* w16, CUtensorMap* tw, cudaStream_t st) {
  k1v4::pack_w16_kernel<<<148, 256, 0, st>>>(wqkv, bqkv, reinterpret_cast<__half*>(w16));
  NRMS_LAUNCH_CHECK("pack_w16_kernel(v4)");
  return make_tmap_k_major_f16(tw, w16, k1v4::W16_ROWS, k1v4::W16_LD, k1v4::W16_LD, k1v4::PN);
}

// fp16 copy [n_rows + 1][320] of n_rows fp32 rows (1.0 in column 300, zero tail, all-zero last row), no tensor map
int k1v4_pack_rows(const float* src, int64_t n_rows, void* src16, cudaStream_t st) {
  const int64_t total = (n_rows + 1) * 40;
  int64_t gb = (total + 255) / 256;
  if (gb > (int64_t)num_sms() * 16) gb = (int64_t)num_sms() * 16;
  k1v4::pack_src16_kernel<<<(unsigned)gb, 256, 0, st>>>(src, n_rows, reinterpret_cast<__half*>(src16));
  NRMS_LAUNCH_CHECK("pack_src16_kernel");
  return NRMS_OK;
}

// fp16 gather source of n_rows fp32 rows (+ the null row) and its gather4 tensor map (box = 64 halfs x 1 row)
int k1v4_pack_src(const float* src, int64_t n_rows, void* src16, CUtensorMap* ts, cudaStream_t st) {
  NRMS_CHECK_ARG((n_rows + 1) * k1v4::SRC_LD * 2 < (1ll << 32), NRMS_E_UNSUPPORTED,
                 "gather source too large (32-bit byte offsets in the K1 gather)");
  const int64_t total = (n_rows + 1) * 40;
  int64_t gb = (total + 255) / 256;
  if (gb > (int64_t)num_sms() * 16) gb = (int64_t)num_sms() * 16;
  k1v4::pack_src16_kernel<<<(unsigned)gb, 256, 0, st>>>(src, n_rows, reinterpret_cast<__half*>(src16));
  NRMS_LAUNCH_CHECK("pack_src16_kernel");
  return make_tmap_k_major_f16(ts, src16, n_rows + 1, k1v4::SRC_LD, k1v4::SRC_LD, 1);
}

template <int S, int SLOT, int SPT>
static int launch_k1v4(const CUtensorMap& tw, const CUtensorMap& ts, const void* idx, int idx_kind, int64_t n,
                       int null_row, void* Cbuf, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k1v4::encoder_attn_tc4_kernel<S, SLOT, SPT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, k1v4::SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(encoder_attn_tc4_kernel)");
    configured = true;
  }
  const int64_t tiles = (n + SPT - 1) / SPT;
  int grid = num_sms();
  if (const char* e = getenv("NRMS_K1_GRID")) { const int g = atoi(e); if (g > 0 && g < grid) grid = g; }   // experiments
  if (tiles < grid) grid = (int)tiles;
  alignas(64) CUtensorMap tc_;     // context rows [n*S][320] halfs, store box = S rows x 40 halfs (one sequence, one pass)
  if (int rc = make_tmap_store_f16(&tc_, Cbuf, n * S, k1v4::CP, k1v4::CP, 40, S)) return rc;
  k1v4::encoder_attn_tc4_kernel<S, SLOT, SPT><<<grid, k1v4::THREADS, k1v4::SMEM, st>>>(
      tw, ts, tc_, idx, idx_kind, n, null_row);
  NRMS_LAUNCH_CHECK("encoder_attn_tc4_kernel");
  return NRMS_OK;
}

// idx_kind 0: sequence s, position i reads source row s*S+i; 1: int64 ids; 2: int32 ids.  null_row = index of the
// all-zero row of the fp16 source (n_rows).
int k1v4_run(int S, const CUtensorMap& tw, const CUtensorMap& ts, const void* idx, int idx_kind, int64_t n,
             int null_row, void* Cbuf, cudaStream_t st) {
  if (S == 20) return launch_k1v4<20, 24, 5>(tw, ts, idx, idx_kind, n, null_row, Cbuf, st);
  if (S == 50) return launch_k1v4<50, 64, 2>(tw, ts, idx, idx_kind, n, null_row, Cbuf, st);
  set_error("encoder_attn_tc4_kernel compiled for S = 20 or 50, got %d", S);
  return NRMS_E_UNSUPPORTED;
}

}  // namespace nrms
