// K1g -- encoder self-attention over PRE-PROJECTED table rows (tensor-mode inference, indexed input).
//
// The Q/K/V projections of both encoders are row-wise linear maps of GATHERED rows (reference:
// src/model/general/attention/multihead_self.py:53-58 applied to the news vectors stacked at src/evaluate.py:220-224,
// and to the embedding rows of src/model/NRMS/news_encoder.py:38), and every row comes from one table.  Projecting the
// TABLE once per call (one [n_rows, 900] kind::f16 GEMM, k1g_project_table) and gathering q, k, v per row gives the same
// numbers with 1/56 (users: 65 k table rows against 3.66 M history rows) or 1/18 (news: 71 k vocabulary rows against
// 1.3 M token rows) of the projection work at MIND-small shapes.  What remains per sequence is 15 heads of S x S x 20
// attention on gathered rows.
//
//   table16 : [n_rows][3 head groups][q | k | v][5 heads][24 halfs] = 2,160 B per row; q pre-scaled by
//             log2(e)/sqrt(20), bias included, every 20-half head slice padded to 48 B (pad = 0 for q and k,
//             (1, 0, 0, 0) for v: the context MMA then also returns Z = sum_j P_ij in column 20)
//   stage   : one sequence = S rows x 2,160 B, one cp.async.bulk per row (completion on the stage's mbarrier);
//             2 stages for S = 50, 5 for S = 20.  The 2,160-byte pitch keeps every ldmatrix (8 rows x 16 B)
//             bank-conflict free.  (Six 36 KB head-group stages fed by 720-byte copies were tried first: bulk copies
//             cost ~75 cycles each whatever their size.)
//   warp 15 : producer (row indices -> bulk copies)
//   warp w  : head w (w = 0..14), FA2-style on mma.sync m16n8k16/k8 (fp16 in, fp32 accumulate): K and V fragments
//             of the head stay in registers for its 16-row query tiles; S = Q K^T -> P = 2^S (ex2.approx) -> fp16 A
//             fragments in registers -> O = P V -> O / (Z + 1e-8) (exp without max subtraction and the 1e-8 of
//             multihead_self.py:17-20) -> staged in place over the head's q slice -> fp16 context rows [S][300]
//             (+ 20 zero columns), the layout K2 reads.
//
// The tensor work here is ~3 MFLOP per user on tiles of 16 x 8: warp-level mma.sync is the right granularity (a 128-row
// tcgen05 tile would be 61 % padding, see K1 v6).  Measured: the kernel is bound by instruction issue first and the
// HMMA pipe second; legacy HMMA runs at 8 cycles per instruction and sub-partition on B200.
//
// Kernels in this file: seq_attn_kernel<SEQ, IdxT, SAFE> (both encoders) and the table projection's operand packing.
// (The first S = 50 forms -- table_attn_kernel with its component-removal switches and the 24-warp unit-parallel
// table_attn_units_kernel, measured slower -- were removed in round 2; their measurements are in DESIGN.md.)
#include <cuda.h>
#include <cuda_fp16.h>
#include "tc_common.cuh"
#include "tc_api.cuh"

namespace nrms {
namespace k1g {

constexpr int S = 50;              // history length
constexpr int H = 15;              // heads
constexpr int HG = 5;              // heads per group (one stage)
constexpr int SLICE = 48;          // bytes per (row, head) slice: 20 halfs + 4 pad
constexpr int GROUP = 3 * HG * SLICE;   // 720 B: q | k | v slices of one head group
constexpr int OFF_K = HG * SLICE;       // 240
constexpr int OFF_V = 2 * HG * SLICE;   // 480
constexpr int PITCH = 3 * GROUP;        // 2,160 B: one table16 row
constexpr int STAGE = S * PITCH;        // 108,000 B: one user
constexpr int NSTAGE = 2;
constexpr int ROW_BYTES = PITCH;
constexpr int HOT_REPLICAS_LOG2 = 6, HOT_REPLICAS = 1 << HOT_REPLICAS_LOG2;      // copies of a hot table row (see the producer)
constexpr int THREADS = 512;
constexpr int CP = 320;            // pitch (halfs) of the context rows K2 reads

__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
}
// D += A(16x16, row) * B(16x8, col), fp16 operands, fp32 accumulate
__device__ __forceinline__ void mma_k16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                        uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// D += A(16x8, row) * B(8x8, col)
__device__ __forceinline__ void mma_k8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ---------------------------------------------------------------------------------------------------------------
// The same kernel for any sequence length that fits the tiling (SEQ = 50 users' histories, SEQ = 20 title tokens):
// head per warp, MT = ceil(SEQ/16) query tiles, NT = ceil(SEQ/8) key tiles (NT odd for both lengths: the last key tile
// is the k8 step of O = P V), one stage = one sequence = SEQ rows of 2,160 B, as many stages as fit (2 / 5).
// IdxT = int32 (history rows into the news-vector table) or int64 (token ids into the embedding table).
// ---------------------------------------------------------------------------------------------------------------
template <int SEQ>
struct SeqCfg {
  static constexpr int MT = (SEQ + 15) / 16;
  static constexpr int NT = (SEQ + 7) / 8;
  static constexpr int KS16 = NT / 2;                 // full k16 steps of O = P V; key tile NT-1 is the k8 step
  static constexpr int REM = SEQ - 8 * (NT - 1);      // valid keys of the last key tile (2 / 4)
  static constexpr int STAGE_BYTES = SEQ * PITCH;
  static constexpr int NST = (SEQ == 50) ? 2 : 5;
  static constexpr int BAR = NST * STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR + 16 * NST + 64;
  static_assert((NT & 1) == 1 && REM % 2 == 0 && SEQ <= 64, "tiling of seq_attn_kernel");
  static_assert(SMEM_BYTES <= 232448, "shared memory");
};

// SAFE = the row-shifted softmax form (below).  Both instantiations are launched when the choice is left to the score
// bound (qk_bound != nullptr): each reads the 30 bounds first and the one that does not apply returns at once (~3 us),
// so the decision needs no host synchronisation and neither form pays for the other's code (a run-time branch inside one
// kernel cost the plain form 9 %: 468 instead of 428 us per 18,288 users).
template <int SEQ, typename IdxT, bool SAFE>
__global__ void __launch_bounds__(THREADS, 1)
seq_attn_kernel(const __half* __restrict__ table16, int64_t n_table_rows, const IdxT* __restrict__ seq_rows,
                int64_t n_seq, __half* __restrict__ ctx, const float* __restrict__ qk_bound, int64_t hot_row) {
  using Cfg = SeqCfg<SEQ>;
  if (qk_bound != nullptr) {
    const int l = threadIdx.x & 31;
    const bool big = l < H && !(qk_bound[l] * qk_bound[H + l] <= 225.f);      // (15 log2 units)^2: 2^15 < 65504
    if ((__ballot_sync(0xffffffffu, big) != 0u) != SAFE) return;
  }
  constexpr int MT = Cfg::MT, NT = Cfg::NT, KS16 = Cfg::KS16, NST = Cfg::NST, STG = Cfg::STAGE_BYTES;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = tc::smem_u32(smem);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t full_bar = sbase + Cfg::BAR, empty_bar = full_bar + 8 * NST;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      tc::mbar_init(full_bar + 8 * s, 1);
      tc::mbar_init(empty_bar + 8 * s, H);
    }
    tc::mbar_fence_init();
  }
  __syncthreads();

  if (warp == H) {
    // ------------------------------ producer: one 2,160-byte bulk copy per row ------------------------------
    // (16-byte cp.async pieces from this one warp were tried: 250 instructions per user, 8.4 ms instead of 2.5 for the user stage)
    uint32_t it = 0;
    for (int64_t u = blockIdx.x; u < n_seq; u += gridDim.x, ++it) {
      int64_t r0 = lane < SEQ ? (int64_t)seq_rows[u * SEQ + lane] : 0;
      int64_t r1 = lane + 32 < SEQ ? (int64_t)seq_rows[u * SEQ + lane + 32] : 0;
      r0 = r0 < 0 ? 0 : (r0 >= n_table_rows ? n_table_rows - 1 : r0);
      r1 = r1 < 0 ? 0 : (r1 >= n_table_rows ? n_table_rows - 1 : r1);
      // The hot row (the padding token of the titles: 45 % of all gathered rows) has HOT_REPLICAS copies behind the table;
      // its references are spread over them by position.  One row read by every SM all the time is an L2 hot spot:
      // measured on the user encoder's PADDED_NEWS row, 488 / 564 -> 417 us per launch (profiles/block_probe.py).
      if (r0 == hot_row) r0 = n_table_rows + (((uint32_t)(u * SEQ + lane) * 2654435761u) >> (32 - HOT_REPLICAS_LOG2));
      if (r1 == hot_row) r1 = n_table_rows + (((uint32_t)(u * SEQ + lane + 32) * 2654435761u) >> (32 - HOT_REPLICAS_LOG2));
      const uint32_t st = it % NST;
      tc::mbar_wait(empty_bar + 8 * st, ((it / NST) & 1) ^ 1);
      if (lane == 0) mbar_arrive_expect_tx(full_bar + 8 * st, STG);
      __syncwarp();
      const uint32_t dst = sbase + st * STG + lane * PITCH;
      if (lane < SEQ)
        bulk_copy_g2s(dst, reinterpret_cast<const char*>(table16) + r0 * ROW_BYTES, PITCH, full_bar + 8 * st);
      if (lane + 32 < SEQ)
        bulk_copy_g2s(dst + 32 * PITCH, reinterpret_cast<const char*>(table16) + r1 * ROW_BYTES, PITCH, full_bar + 8 * st);
    }
  } else {
    // ------------------------------ compute: warp = head ------------------------------
    const int hg = warp / HG, hl = warp - hg * HG;
    const int qoff = hg * GROUP + hl * SLICE, koff = OFF_K + qoff, voff = OFF_V + qoff;
    const int g = lane >> 2, t = lane & 3;
    const int mi = lane >> 3, rr = lane & 7;     // ldmatrix: this lane supplies row rr of matrix mi
    // lane-constant byte offsets of the ldmatrix rows inside a stage (padded rows read the last real row: finite values,
    // padded keys are masked out of P, padded query rows are never stored)
    auto rofs = [](int row) { return (uint32_t)((row < SEQ ? row : SEQ - 1) * PITCH); };
    constexpr int KP = (NT + 1) / 2, K8P = (NT + 3) / 4, VS = KS16 + 1;
    uint32_t k16o[KP], k8o[K8P], v4o[VS], v2o[VS], q4o[MT], q2o[MT];
#pragma unroll
    for (int p = 0; p < KP; ++p) k16o[p] = rofs(16 * p + 8 * (mi >> 1) + rr) + koff + (mi & 1) * 16;
#pragma unroll
    for (int p = 0; p < K8P; ++p) k8o[p] = rofs(32 * p + 8 * mi + rr) + koff + 32;
#pragma unroll
    for (int p = 0; p < VS; ++p) {
      v4o[p] = rofs(16 * p + 8 * (mi & 1) + rr) + voff + (mi >> 1) * 16;
      v2o[p] = rofs(16 * p + 8 * (mi & 1) + rr) + voff + 32;
    }
#pragma unroll
    for (int p = 0; p < MT; ++p) {
      q4o[p] = rofs(16 * p + 8 * (mi & 1) + rr) + qoff + (mi >> 1) * 16;
      q2o[p] = rofs(16 * p + 8 * (mi & 1) + rr) + qoff + 32;
    }
    // P = 2^s is packed to fp16 for the context MMA and overflows at s > 16 (a logit of 11.09) where the reference's fp32
    // exp (multihead_self.py:17) is finite to 88.  When the per-call bound max|q_h| max|k_h| over the projected table
    // (k1f_qk_bound: Cauchy-Schwarz, log2 units) allows scores above 15 the rows are shifted by their maximum (SAFE):
    // P = 2^(s - m), O / (Z + 1e-8 * 2^-m) -- algebraically exp(s) / (sum exp(s) + 1e-8) for any m.
    constexpr bool safe = SAFE;
    uint32_t it = 0;
    for (int64_t u = blockIdx.x; u < n_seq; u += gridDim.x, ++it) {
      const uint32_t st = it % NST;
      const uint32_t B = sbase + st * STG;
      tc::mbar_wait(full_bar + 8 * st, (it / NST) & 1);
      // K fragments: kb16[nt][0..1] (dims 0-7, 8-15), kb8[nt] (dims 16-23); V fragments (transposed loads):
      // vb[ks][dt][0..1] = keys 16ks..+7 / +8..15, dims 8dt..8dt+7.  They stay in registers for the MT query tiles.
      uint32_t kb16[2 * KP][2], kb8[4 * K8P], vb[VS][3][2];
#pragma unroll
      for (int p = 0; p < KP; ++p)
        ldsm_x4(B + k16o[p], kb16[2 * p][0], kb16[2 * p][1], kb16[2 * p + 1][0], kb16[2 * p + 1][1]);
#pragma unroll
      for (int p = 0; p < K8P; ++p) ldsm_x4(B + k8o[p], kb8[4 * p], kb8[4 * p + 1], kb8[4 * p + 2], kb8[4 * p + 3]);
#pragma unroll
      for (int ks = 0; ks < VS; ++ks) {
        ldsm_x4_t(B + v4o[ks], vb[ks][0][0], vb[ks][0][1], vb[ks][1][0], vb[ks][1][1]);
        ldsm_x2_t(B + v2o[ks], vb[ks][2][0], vb[ks][2][1]);
      }
      auto scores = [&](int mt, float (&sacc)[NT][4]) {      // S = Q K^T of one 16-row query tile
        uint32_t qa[4], qb[2];
        ldsm_x4(B + q4o[mt], qa[0], qa[1], qa[2], qa[3]);
        ldsm_x2(B + q2o[mt], qb[0], qb[1]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f;
          mma_k16(sacc[nt], qa[0], qa[1], qa[2], qa[3], kb16[nt][0], kb16[nt][1]);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_k8(sacc[nt], qb[0], qb[1], kb8[nt]);
      };
      float sacc[2][NT][4];
      scores(0, sacc[0]);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if (mt + 1 < MT) scores(mt + 1, sacc[(mt + 1) & 1]);
        float (&sc)[NT][4] = sacc[mt & 1];
        // P = 2^S (q carries log2(e)/sqrt(20)); the keys past SEQ in the last key tile are padding
        uint32_t pa[NT][2];
        const bool lower = mt + 1 < MT;              // rows 16mt+8..+15 exist only before the last tile
        float e0 = 1e-8f, e1 = 1e-8f;                // the epsilon of multihead_self.py:20, scaled with the row shift
        if (safe) {
          float m0 = sc[0][0], m1 = sc[0][2];
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            m0 = fmaxf(m0, fmaxf(sc[nt][0], sc[nt][1]));
            m1 = fmaxf(m1, fmaxf(sc[nt][2], sc[nt][3]));
          }
          m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
          m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
          m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
          m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
          m0 = fminf(fmaxf(m0, -120.f), 120.f);
          m1 = fminf(fmaxf(m1, -120.f), 120.f);
          e0 = 1e-8f * ex2f(-m0);
          e1 = 1e-8f * ex2f(-m1);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            float p0 = ex2f(sc[nt][0] - m0), p1 = ex2f(sc[nt][1] - m0);
            float p2 = 0.f, p3 = 0.f;
            if (lower) { p2 = ex2f(sc[nt][2] - m1); p3 = ex2f(sc[nt][3] - m1); }
            if (nt == NT - 1 && 2 * t >= Cfg::REM) { p0 = p1 = p2 = p3 = 0.f; }
            pa[nt][0] = pack_h2(p0, p1);
            pa[nt][1] = pack_h2(p2, p3);
          }
        } else {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            float p0 = ex2f(sc[nt][0]), p1 = ex2f(sc[nt][1]);
            float p2 = 0.f, p3 = 0.f;
            if (lower) { p2 = ex2f(sc[nt][2]); p3 = ex2f(sc[nt][3]); }
            if (nt == NT - 1 && 2 * t >= Cfg::REM) { p0 = p1 = p2 = p3 = 0.f; }
            pa[nt][0] = pack_h2(p0, p1);
            pa[nt][1] = pack_h2(p2, p3);
          }
        }
        float oacc[3][4];
#pragma unroll
        for (int dt = 0; dt < 3; ++dt) oacc[dt][0] = oacc[dt][1] = oacc[dt][2] = oacc[dt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS16; ++ks)
#pragma unroll
          for (int dt = 0; dt < 3; ++dt)
            mma_k16(oacc[dt], pa[2 * ks][0], pa[2 * ks][1], pa[2 * ks + 1][0], pa[2 * ks + 1][1], vb[ks][dt][0],
                    vb[ks][dt][1]);
#pragma unroll
        for (int dt = 0; dt < 3; ++dt) mma_k8(oacc[dt], pa[NT - 1][0], pa[NT - 1][1], vb[KS16][dt][0]);
        // Z of rows g / g+8 sits in column 20 = element 0 / 2 of dim tile 2 on the quad's lane t == 2
        const float z0 = __shfl_sync(0xffffffffu, oacc[2][0], (lane & ~3) | 2);
        const float z1 = __shfl_sync(0xffffffffu, oacc[2][2], (lane & ~3) | 2);
        const float i0 = __fdividef(1.f, z0 + e0), i1 = __fdividef(1.f, z1 + e1);
        // O / (Z + 1e-8) back in place over the head's q slice of these rows, then out as 8-byte pieces
        const int r0 = 16 * mt + g, r1 = r0 + 8;
        uint8_t* q0 = smem + st * STG + r0 * PITCH + qoff + 4 * t;
#pragma unroll
        for (int dt = 0; dt < 3; ++dt) {
          if (dt < 2 || t < 2) {
            if (r0 < SEQ) *reinterpret_cast<uint32_t*>(q0 + dt * 16) = pack_h2(oacc[dt][0] * i0, oacc[dt][1] * i0);
            if (r1 < SEQ) *reinterpret_cast<uint32_t*>(q0 + 8 * PITCH + dt * 16) = pack_h2(oacc[dt][2] * i1, oacc[dt][3] * i1);
          }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int p = lane + 32 * k;                 // 80 pieces: 16 rows x 5 x 8 bytes
          const int row = 16 * mt + p / 5, c = p % 5;
          if (p < 80 && row < SEQ) {
            const uint2 v = *reinterpret_cast<const uint2*>(smem + st * STG + row * PITCH + qoff + c * 8);
            uint8_t* orow = reinterpret_cast<uint8_t*>(ctx) + (u * SEQ + row) * (CP * 2);
            *reinterpret_cast<uint2*>(orow + warp * 40 + c * 8) = v;
            // the last head's warp also clears columns 300..319 (K2 multiplies them by zero weights: they must be finite)
            if (warp == H - 1) *reinterpret_cast<uint2*>(orow + 600 + c * 8) = make_uint2(0u, 0u);
          }
        }
        if (mt == MT - 1) {          // last shared-memory access of this stage: hand it back to the producer
          tc::fence_proxy_async_smem();      // the generic-proxy stores above precede the next bulk copy into these rows
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(empty_bar + 8 * st);
        }
      }
    }
  }
}

}  // namespace k1g

// fp16 copy of the packed projection weights IN THE ORDER OF A TABLE ROW: [1080][320] halfs.  Row j of the copy produces
// half j of a table16 row: j = head group * 360 + (q | k | v) * 120 + head * 24 + d.  d < 20: row (which * 300 + head * 20 + d)
// of [W_Q; W_K; W_V], its bias in column 300 (met by the 1.0 column of the operand rows), W_Q and b_Q scaled by
// log2(e)/sqrt(20); d >= 20: the slice pad -- a zero row, except d = 20 of v whose bias column is 1.0, so the GEMM itself
// writes the (1, 0, 0, 0) that returns Z from the context MMA.  The table then IS the row-major GEMM result.
constexpr int TABLE_COLS = k1g::PITCH / 2;      // 1,080
__global__ void __launch_bounds__(256) pack_wqkv16_table_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                                                 __half* __restrict__ out, float qscale) {
  const int n = TABLE_COLS * 320;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int j = i / 320, k = i - j * 320;
    const int hg = j / 360, r1 = j - hg * 360, which = r1 / 120, r2 = r1 - which * 120, hl = r2 / 24, d = r2 - hl * 24;
    float v = 0.f;
    if (d < DH) {
      const int src = which * D + (hg * k1g::HG + hl) * DH + d;
      const float sc = which == 0 ? qscale : 1.f;
      if (k < D) v = w[src * D + k] * sc;
      else if (k == D) v = bias[src] * sc;
    } else if (which == 2 && d == DH && k == D) {
      v = 1.f;
    }
    out[i] = __float2half_rn(v);
  }
}

int pack_rows16(const float* src, int64_t n_rows, void* src16, cudaStream_t st);      // pack.cu

// scratch of the table path: [table16: n_rows x 2,160 B][fp16 copy of the source rows: (n_rows + 1) x 640 B, unless the
// caller already has one][fp16 weights in table order]
static size_t k1g_table16_only(int64_t n_rows) { return align_up((size_t)(n_rows + k1g::HOT_REPLICAS) * k1g::ROW_BYTES, 1024); }

// rows [n_rows, n_rows + HOT_REPLICAS) of the projected table <- row hot_row
static __global__ void __launch_bounds__(256)
replicate_row_kernel(__half* __restrict__ table16, int64_t n_rows, int64_t hot_row) {
  const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(table16) + hot_row * k1g::ROW_BYTES);
  uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<char*>(table16) + (n_rows + blockIdx.x) * k1g::ROW_BYTES);
  for (int i = threadIdx.x; i < k1g::ROW_BYTES / 16; i += blockDim.x) dst[i] = src[i];
}
static size_t k1g_a16_bytes(int64_t n_rows) { return align_up((size_t)(n_rows + 1) * 640, 1024); }
size_t k1g_table16_bytes(int64_t n_rows, bool rows16_given) {
  return k1g_table16_only(n_rows) + k1g_a16_bytes(n_rows) * (rows16_given ? 0 : 1) + (size_t)TABLE_COLS * 640;
}
const void* k1g_table16_ptr(void* scratch) { return scratch; }

// table16 = half([table | 1] * B^T) with B = the weight copy above (rows in table order, bias column, q scale and pads
// included): fp16 copies of the table (a 1.0 in column 300 meets the bias column) and of the weights, then one kind::f16
// tensor-core GEMM with N = 1,080 whose row-major result IS the table; the epilogue leaves through bulk tensor stores.
// (Same operand precision as K1 v6's in-kernel projection.)  rows16: the fp16 copy of the source rows when the caller
// already holds one (pack_rows16 layout), else nullptr and `table` (fp32) is packed here.
int k1g_project_table(const float* table, const void* rows16, int64_t n_rows, const float* wqkv, const float* bqkv,
                      void* scratch, cudaStream_t st, int64_t hot_row) {
  char* base = reinterpret_cast<char*>(scratch);
  void* table16 = base;
  const void* a16 = rows16;
  char* next = base + k1g_table16_only(n_rows);
  if (a16 == nullptr) {
    if (int rc = pack_rows16(table, n_rows, next, st)) return rc;
    a16 = next;
    next += k1g_a16_bytes(n_rows);
  }
  void* b16 = next;
  const float qscale = 1.4426950408889634f / sqrtf((float)DH);
  pack_wqkv16_table_kernel<<<148, 256, 0, st>>>(wqkv, bqkv, reinterpret_cast<__half*>(b16), qscale);
  NRMS_LAUNCH_CHECK("pack_wqkv16_table_kernel");
  if (int rc = tc_gemm_nt_f16_tma(a16, 320, b16, 320, table16, TABLE_COLS, n_rows, TABLE_COLS, 304, st)) return rc;
  if (hot_row >= 0 && hot_row < n_rows) {
    replicate_row_kernel<<<k1g::HOT_REPLICAS, 256, 0, st>>>(reinterpret_cast<__half*>(table16), n_rows, hot_row);
    NRMS_LAUNCH_CHECK("replicate_row_kernel");
  }
  return NRMS_OK;
}

// S = 50 (int32 history rows / int64 ids) or S = 20; Cbuf: fp16 context rows [n_seq * S][320] (columns 300..319 cleared here)
template <int SEQ, typename IdxT>
static int launch_seq_attn(const void* table16, int64_t n_table_rows, const void* rows, int64_t n_seq, void* Cbuf,
                           const float* qk_bound, cudaStream_t st, int64_t hot_row) {
  int dev = 0;
  cudaGetDevice(&dev);
  static bool configured[64] = {false};
  if (dev < 0 || dev >= 64 || !configured[dev]) {      // the attribute is per device
    cudaError_t e = cudaFuncSetAttribute(k1g::seq_attn_kernel<SEQ, IdxT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         k1g::SeqCfg<SEQ>::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(k1g::seq_attn_kernel<SEQ, IdxT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               k1g::SeqCfg<SEQ>::SMEM_BYTES);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(seq_attn_kernel)");
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  if (n_seq <= 0) return NRMS_OK;
  NRMS_CHECK_ARG(n_table_rows > 0, NRMS_E_INVALID, "table row count out of range");
  int grid = num_sms();
  if (n_seq < grid) grid = (int)n_seq;
  int force = get_attn_safe_softmax();                // -1: both instantiations read the bound, one of them runs
  if (force < 0 && qk_bound == nullptr) force = 1;    // no bound given: the row-shifted form (always valid)
  const float* bound = force < 0 ? qk_bound : nullptr;
  if (force <= 0) {
    k1g::seq_attn_kernel<SEQ, IdxT, false><<<grid, k1g::THREADS, k1g::SeqCfg<SEQ>::SMEM_BYTES, st>>>(
        reinterpret_cast<const __half*>(table16), n_table_rows, reinterpret_cast<const IdxT*>(rows), n_seq,
        reinterpret_cast<__half*>(Cbuf), bound, hot_row);
    NRMS_LAUNCH_CHECK("seq_attn_kernel");
  }
  if (force != 0 && (force > 0 || bound != nullptr)) {
    k1g::seq_attn_kernel<SEQ, IdxT, true><<<grid, k1g::THREADS, k1g::SeqCfg<SEQ>::SMEM_BYTES, st>>>(
        reinterpret_cast<const __half*>(table16), n_table_rows, reinterpret_cast<const IdxT*>(rows), n_seq,
        reinterpret_cast<__half*>(Cbuf), bound, hot_row);
    NRMS_LAUNCH_CHECK("seq_attn_kernel(row-shifted)");
  }
  return NRMS_OK;
}

int k1g_run_seq(int S, int idx_kind, const void* table16, int64_t n_table_rows, const void* rows, int64_t n_seq,
                void* Cbuf, const float* qk_bound, cudaStream_t st, int64_t hot_row) {
  if (S == 50 && idx_kind == 2) return launch_seq_attn<50, int32_t>(table16, n_table_rows, rows, n_seq, Cbuf, qk_bound, st, hot_row);
  if (S == 50 && idx_kind == 1) return launch_seq_attn<50, int64_t>(table16, n_table_rows, rows, n_seq, Cbuf, qk_bound, st, hot_row);
  if (S == 20 && idx_kind == 1) return launch_seq_attn<20, int64_t>(table16, n_table_rows, rows, n_seq, Cbuf, qk_bound, st, hot_row);
  if (S == 20 && idx_kind == 2) return launch_seq_attn<20, int32_t>(table16, n_table_rows, rows, n_seq, Cbuf, qk_bound, st, hot_row);
  set_error("seq_attn_kernel: unsupported (S, index kind) = (%d, %d)", S, idx_kind);
  return NRMS_E_UNSUPPORTED;
}

}  // namespace nrms
