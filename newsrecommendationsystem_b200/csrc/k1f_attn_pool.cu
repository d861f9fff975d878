// K1f -- encoder self-attention over PRE-PROJECTED table rows with the additive-attention pooling IN THE SAME KERNEL
// (tensor-mode inference, indexed input).  The fp16 context rows never leave the SM: the round trip K1g -> HBM -> K2
// of the first table-path kernels (30 KB written and 30 KB read back per user) is gone.
//
// Reference math: src/model/general/attention/multihead_self.py:15-23,46-76 (exp-softmax with the 1e-8 in the
// denominator, no output projection) followed by src/model/general/attention/additive.py:27-53, i.e. the body of
// src/model/NRMS/user_encoder.py:15-26 and of src/model/NRMS/news_encoder.py:41-48 in eval mode.
//
//   table16 : [n_rows][3 head groups][q | k | v][5 heads][24 halfs] = 2,160 B per row (k1g_project_table): q pre-scaled
//             by log2(e)/sqrt(20), bias included, v pad = (1,0,0,0) so that O = P V also returns Z = sum_j P_ij.
//   stage   : one sequence = SEQ rows x 2,160 B, one cp.async.bulk per row.  Every head warp pulls its K, V AND Q
//             fragments into registers right after the stage lands and hands the stage back at once, so ONE stage
//             (S = 50; three for S = 20) keeps a whole sequence in flight behind the attention of the previous one.
//   warp h  : head h (0..14), FA2-style on mma.sync m16n8k16/k8: S = Q K^T -> P = 2^S -> O = P V -> O / (Z + 1e-8),
//             written as fp16 straight into the CONTEXT TILE: [64 rows][304 halfs] in the UMMA SWIZZLE_128B K-major
//             layout (5 chunks of 64 halfs), double buffered.  A tile holds one user (50 rows) or three titles (60).
//   warp 15 : producer (bulk copies, row indices fetched one sequence ahead), tcgen05 issuer, softmax over the sequence.
//   additive: T^T = W_a C^T on tcgen05 (kind::f16): A = W_a (fp16) RESIDENT IN TENSOR MEMORY for the whole kernel
//             (2 M-tiles of 128 hidden units x 152 packed columns -- shared memory has no room for its 122 KB next to
//             the gather stage), B = the context tile (N = 64 positions), D = 2 x 64 fp32 columns.  The head warps read
//             D back between two of their query tiles (thread = hidden unit): tanh(. + b_a) * q_a, a butterfly
//             reduce-scatter over the 32 hidden units of the warp, one partial logit row per (M-tile, lane quarter).
//             Warp 15 adds the seven partial rows in a fixed order (deterministic), takes the stable softmax
//             (additive.py:37-39) and leaves the weights in shared memory as mma.sync A fragments, split into two
//             fp16 terms (rows g / g+8), so the pooled sum carries fp32-accurate weights; one step later every head
//             warp pools its share of the 38 eight-column units of out = sum_i w_i c_i (context tile read back with
//             transposing ldmatrix) and writes them to global memory.  (A first version pooled on warp 15 alone:
//             3,300 cycles per user on the critical path, 3.05 ms per 73,152 users against 2.23 without it.)
//
// Large scores.  P = 2^s is packed to fp16 for the context MMA, which overflows at s > 16 (e^11.09) where the
// reference's fp32 exp is finite to 88.  A per-call bound max|q_h| max|k_h| (k1f_qk_bound over the projected table)
// selects, inside the kernel, between the plain form (bound <= 15: cannot overflow) and the row-shifted form
// P = 2^(s - m_i), O / (Z + 1e-8 * 2^-m_i) -- algebraically the reference's exp(s)/(sum exp(s) + 1e-8) for any m_i.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "tc_common.cuh"
#include "tc_api.cuh"

namespace nrms {
namespace k1f {

constexpr int NH = 15;             // heads
constexpr int HG = 5;              // heads per group
constexpr int SLICE = 48;          // bytes per (row, head) slice: 20 halfs + 4 pad
constexpr int GROUP = 3 * HG * SLICE;   // 720 B: q | k | v slices of one head group
constexpr int OFF_K = HG * SLICE;       // 240
constexpr int OFF_V = 2 * HG * SLICE;   // 480
constexpr int PITCH = 3 * GROUP;        // 2,160 B: one table16 row
constexpr int THREADS = 640;            // 15 head warps, warp 15 (producer / issuer / softmax), 4 additive-epilogue warps
// setmaxnreg targets of the two kinds of warpgroups.  The pool is the CTA's OWN launch allocation (640 threads x 96
// registers), not the SM's free registers: the increments must be covered by the decrement or setmaxnreg.inc waits forever.
constexpr int REG_LAUNCH = 96, REG_HEAD = 104, REG_EPI = 64;
static_assert(16 * REG_HEAD + 4 * REG_EPI <= 20 * REG_LAUNCH, "setmaxnreg budget");
constexpr int CTX_ROWS = 64;                       // positions per context tile = UMMA N
constexpr int CTX_CHUNK = CTX_ROWS * 128;          // 8,192 B: 64 rows x 64 halfs
constexpr int CTX_TILE = 5 * CTX_CHUNK;            // 40,960 B
constexpr int KSTEPS = 19;                         // 304 / 16
constexpr int TM_A = 0;                            // W_a: M-tile m at columns 152 m
constexpr int TM_D = 304;                          // D: M-tile m at columns 304 + 64 m
constexpr int N_EPI_WARPS = 4;                     // warps 16..19: one per TMEM lane quarter
constexpr float SAFE_BOUND_SQ = 225.f;             // (15 log2 units)^2: 2^15 < 65504

template <int SEQ>
struct Cfg {
  static constexpr int MT = (SEQ + 15) / 16;
  static constexpr int NT = (SEQ + 7) / 8;
  static constexpr int KS16 = NT / 2;                 // full k16 steps of O = P V; key tile NT-1 is the k8 step
  static constexpr int REM = SEQ - 8 * (NT - 1);      // valid keys of the last key tile (2 / 4)
  static constexpr int SPT = CTX_ROWS / SEQ;          // sequences per context tile (1 / 3)
  static constexpr int NST = (SEQ == 50) ? 1 : 3;
  static constexpr int STAGE_BYTES = SEQ * PITCH;
  static constexpr int OFF_CTX = 0;
  static constexpr int OFF_STAGE = 2 * CTX_TILE;
  static constexpr int OFF_PART = OFF_STAGE + NST * STAGE_BYTES;     // [2][8][64] fp32 partial logits
  static constexpr int OFF_W = OFF_PART + 2 * 8 * 64 * 4;            // [64] fp32 softmax weights
  static constexpr int OFF_AF = OFF_W + 256;                         // [2][4][32] uint4 A fragments of the pooling MMA
  static constexpr int OFF_BAR = OFF_AF + 2 * 4 * 32 * 16;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;            // barriers + tmem pointer | alignment slack
  static_assert((NT & 1) == 1 && REM % 2 == 0 && SEQ <= 64 && MT >= 2, "tiling");
  static_assert(OFF_PART % 16 == 0, "alignment");
  static_assert(SMEM_BYTES <= 232448, "shared memory");
};

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
__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void tmem_ld32_nw(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// qk_bound_kernel: bound[h] = max_r |q_r^h|^2, bound[15 + h] = max_r |k_r^h|^2 over the projected table (q carries
// log2(e)/sqrt(20), so sqrt(bound[h] bound[15+h]) bounds every score of head h in log2 units by Cauchy-Schwarz).
// Non-negative floats order like their bit patterns: atomicMax on the unsigned view.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qk_bound_kernel(const __half* __restrict__ table16, int64_t n_rows,
                                                       unsigned int* __restrict__ bound) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int which = lane / NH, head = lane - which * NH;       // lanes 0..29: (q | k, head)
  const int off = (head / HG) * GROUP + which * OFF_K + (head % HG) * SLICE;
  float best = 0.f;
  if (lane < 2 * NH) {
    for (int64_t r = warp0; r < n_rows; r += n_warps) {
      const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(table16) + r * PITCH + off);
      const uint4 a = __ldg(p), b = __ldg(p + 1);
      const uint2 c = __ldg(reinterpret_cast<const uint2*>(p + 2));
      const uint32_t w[10] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y};
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        s = fmaf(f.x, f.x, fmaf(f.y, f.y, s));
      }
      if (!(s <= 3.0e38f)) s = 3.0e38f;            // inf / NaN rows force the row-shifted form
      best = fmaxf(best, s);
    }
    atomicMax(bound + lane, __float_as_uint(best));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The fused kernel.  IdxT = int32 (history rows into the news-vector table) or int64 (token ids into the embedding table).
// force_safe: -1 = decide from qk_bound, 0 / 1 = plain / row-shifted form (tests).
// ---------------------------------------------------------------------------------------------------------------
template <int SEQ, typename IdxT>
__global__ void __launch_bounds__(THREADS, 1)
attn_pool_kernel(const __half* __restrict__ table16, int64_t n_table_rows, const IdxT* __restrict__ seq_rows,
                 int64_t n_seq, const __half* __restrict__ wa16, const float* __restrict__ ba,
                 const float* __restrict__ qa, const float* __restrict__ qk_bound, int force_safe, int dbg,
                 float* __restrict__ out) {
  // dbg (option "k1f_debug", timing experiments only -- results are garbage): 1 no gather copies, 2 no tcgen05 MMAs,
  // 4 no pooling, 8 no tanh / reduce in the additive epilogue, 16 no context stores, 32 no attention
  using C = Cfg<SEQ>;
  constexpr int MT = C::MT, NT = C::NT, KS16 = C::KS16, NST = C::NST, STG = C::STAGE_BYTES, SPT = C::SPT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (sbase - raw);
  float* part = reinterpret_cast<float*>(sm + C::OFF_PART);     // [2][8][64]
  float* wsm = reinterpret_cast<float*>(sm + C::OFF_W);         // [64]
  uint4* afrag = reinterpret_cast<uint4*>(sm + C::OFF_AF);      // [2][4][32]
  const uint32_t bars = sbase + C::OFF_BAR;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * NST;
  const uint32_t ctx_full = bars + 16 * NST, ctx_free = ctx_full + 16, part_full = ctx_free + 16;
  const uint32_t mma_done = part_full + 16, d_free = mma_done + 16, w_ready = d_free + 8;     // mma_done: one per M-tile
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(sm + C::OFF_BAR + 16 * NST + 104);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler
  // sequences of this CTA: u = blockIdx.x + it * gridDim.x, it < n_local; context tile T holds it = T*SPT .. T*SPT+SPT-1
  const int64_t n_local = (n_seq - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      tc::mbar_init(full_bar + 8 * s, 1);
      tc::mbar_init(empty_bar + 8 * s, NH);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(ctx_full + 8 * b, NH);
      tc::mbar_init(ctx_free + 8 * b, N_EPI_WARPS);
      tc::mbar_init(mma_done + 8 * b, 1);
      tc::mbar_init(part_full + 8 * b, N_EPI_WARPS);
      tc::mbar_init(w_ready + 8 * b, 1);
    }
    tc::mbar_init(d_free, N_EPI_WARPS);
    tc::mbar_fence_init();
  }
  if (warp == NH) tc::tmem_alloc(tc::smem_u32((const void*)tmem_ptr_smem), 512);
  // both context tiles start as zeros: rows past the last sequence of a tile are never written, and every column the
  // additive MMA multiplies must be finite
  for (int i = tid; i < (2 * CTX_TILE) / 16; i += THREADS) reinterpret_cast<uint4*>(sm + C::OFF_CTX)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < 2 * 8 * 64 + 64; i += THREADS) part[i] = 0.f;      // partial logits + weights
  __syncthreads();
  // column 300 of all 128 context rows = 1.0 (chunk 4, byte 88 of the 128-byte row: 16-byte unit 5, offset 8)
  if (tid < 2 * CTX_ROWS) {
    const int tl = tid >> 6, row = tid & 63;
    *reinterpret_cast<__half*>(sm + C::OFF_CTX + tl * CTX_TILE + 4 * CTX_CHUNK + row * 128 + ((5 ^ (row & 7)) << 4) + 8) = __float2half_rn(1.f);
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // ---- W_a -> tensor memory (once): lane = hidden unit, 152 packed-fp16x2 columns per M-tile; 16 warps = 4 lane
  //      quarters x (M-tile, column half) ----
  if (warp < 16) {
    const int q4 = warp & 3, j = warp >> 2, m = j >> 1, ch = j & 1;
    const int n = 128 * m + 32 * q4 + lane;
    const uint4* src = reinterpret_cast<const uint4*>(wa16 + (size_t)(n < QD ? n : 0) * 320 + 152 * ch);   // 304 B = 19 x 16 B
    const uint32_t tcol = tmem_base + TM_A + 152 * m + 76 * ch + ((uint32_t)(32 * q4) << 16);
    uint32_t r[16];
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {                      // 4 x 16 columns
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (n < QD) v = __ldg(src + 4 * c + i);
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
      }
      tc::tmem_st16(tcol + 16 * c, r);
    }
    {                                                  // + 8 + 4 columns = 76
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (n < QD) v = __ldg(src + 16 + i);
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
      }
      // k = 300 (packed column 150 = local column 74 of the upper half): b_a, met by the constant 1.0 in column 300 of
      // every context row -- the additive GEMM returns T + b_a
      if (ch == 1 && n < QD) r[10] = (r[10] & 0xFFFF0000u) | (uint32_t)__half_as_ushort(__float2half_rn(ba[n]));
      tc::tmem_st8(tcol + 64, r);
      tc::tmem_st4(tcol + 72, r + 8);
    }
    tc::tmem_st_wait();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();

  if (warp == NH) {
    // =========================== warp 15: producer, tcgen05 issuer, softmax over the sequence ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_HEAD));     // warpgroup 3 = heads 12..14 + this warp
    const uint32_t el = tc::elect_one_u32();
    const uint32_t idesc = tc::umma_idesc_f16(128, CTX_ROWS);
    const uint64_t desc0 = tc::umma_desc_k_sw128(0);
    const int g = lane >> 2, t = lane & 3;
    // row indices of the NEXT sequence to produce: loaded one step ahead, so that their global-memory latency is not
    // part of the chain "stage free -> copies issued"
    int64_t nr0 = 0, nr1 = 0;
    auto load_rows = [&](int64_t it) {
      if (it >= n_local) return;
      const int64_t u = blockIdx.x + it * gridDim.x;
      nr0 = lane < SEQ ? (int64_t)__ldg(seq_rows + u * SEQ + lane) : 0;
      nr1 = lane + 32 < SEQ ? (int64_t)__ldg(seq_rows + u * SEQ + lane + 32) : 0;
      nr0 = nr0 < 0 ? 0 : (nr0 >= n_table_rows ? n_table_rows - 1 : nr0);
      nr1 = nr1 < 0 ? 0 : (nr1 >= n_table_rows ? n_table_rows - 1 : nr1);
    };
    auto produce = [&](int64_t it) {
      const int64_t r0 = nr0, r1 = nr1;
      load_rows(it + 1);
      const uint32_t st = (uint32_t)(it % NST);
      tc::mbar_wait(empty_bar + 8 * st, (uint32_t)((it / NST) & 1) ^ 1u);
      if (dbg & 1) {
        if (lane == 0) tc::mbar_arrive(full_bar + 8 * st);
        return;
      }
      if (lane == 0) mbar_arrive_expect_tx(full_bar + 8 * st, STG);
      __syncwarp();
      const uint32_t dst = sbase + C::OFF_STAGE + st * STG + lane * PITCH;
      if (lane < SEQ)
        bulk_copy_g2s(dst, reinterpret_cast<const char*>(table16) + r0 * PITCH, PITCH, full_bar + 8 * st);
      if (lane + 32 < SEQ)
        bulk_copy_g2s(dst + 32 * PITCH, reinterpret_cast<const char*>(table16) + r1 * PITCH, PITCH, full_bar + 8 * st);
    };
    // additive GEMM of tile T: D[m] = W_a[m] (TMEM) x C^T (context tile T & 1), 2 x 19 MMAs 128 x 64 x 16
    auto issue_mma = [&](int64_t T) {
      const uint32_t b = (uint32_t)(T & 1);
      tc::mbar_wait(ctx_full + 8 * b, (uint32_t)((T >> 1) & 1));
      if (T > 0) tc::mbar_wait(d_free, (uint32_t)((T - 1) & 1));      // the head warps have read D of tile T-1
      tc::tc_fence_after();
      const uint32_t cb = (sbase + C::OFF_CTX + b * CTX_TILE) >> 4;
#pragma unroll
      for (int m = 0; m < 2; ++m) {
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks)
          if (!(dbg & 2)) tc::umma_f16_ts_p(tmem_base + TM_D + 64 * m, tmem_base + TM_A + 152 * m + 8 * ks,
                            desc0 | (uint64_t)((cb + (ks >> 2) * (CTX_CHUNK >> 4) + (ks & 3) * 2) & 0x3FFF), idesc,
                            ks ? 1u : 0u, el);
        tc::umma_commit_p(mma_done + 8 * m, el);       // the epilogue of M-tile 0 runs under the MMAs of M-tile 1
      }
    };
    // logits of tile T -> softmax over each of its sequences -> the A fragments of the pooling MMA (row g < SPT = fp16
    // head of the weights of sequence g, row g + 8 = the fp16 remainder: the sum of the two rows carries ~22 bits)
    auto softmax = [&](int64_t T) {
      const uint32_t b = (uint32_t)(T & 1);
      tc::mbar_wait(part_full + 8 * b, (uint32_t)((T >> 1) & 1));
      const float* pp = part + b * 512;
      float l0 = 0.f, l1 = 0.f;                       // logits of positions lane, lane + 32: seven partial rows, fixed order
#pragma unroll
      for (int un = 0; un < 7; ++un) { l0 += pp[un * 64 + lane]; l1 += pp[un * 64 + 32 + lane]; }
      float w0 = 0.f, w1 = 0.f;
#pragma unroll
      for (int s = 0; s < SPT; ++s) {
        const bool in0 = lane >= s * SEQ && lane < (s + 1) * SEQ;
        const bool in1 = lane + 32 >= s * SEQ && lane + 32 < (s + 1) * SEQ;
        const float mx = warp_max(fmaxf(in0 ? l0 : -INFINITY, in1 ? l1 : -INFINITY));
        const float e0 = in0 ? __expf(l0 - mx) : 0.f, e1 = in1 ? __expf(l1 - mx) : 0.f;
        const float inv = 1.f / warp_sum(e0 + e1);
        if (in0) w0 = e0 * inv;
        if (in1) w1 = e1 * inv;
      }
      wsm[lane] = w0;
      wsm[lane + 32] = w1;
      __syncwarp();
      uint4* af = afrag + b * 128;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[4];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int p = 16 * ks + 8 * hf + 2 * t;
          float x0 = 0.f, x1 = 0.f;
          if (g < SPT) {
            if (p >= g * SEQ && p < (g + 1) * SEQ) x0 = wsm[p];
            if (p + 1 >= g * SEQ && p + 1 < (g + 1) * SEQ) x1 = wsm[p + 1];
          }
          const __half2 hi = __floats2half2_rn(x0, x1);
          const float2 hf2 = __half22float2(hi);
          a[2 * hf] = *reinterpret_cast<const uint32_t*>(&hi);
          a[2 * hf + 1] = pack_h2(x0 - hf2.x, x1 - hf2.y);
        }
        af[ks * 32 + lane] = make_uint4(a[0], a[1], a[2], a[3]);
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(w_ready + 8 * b);
    };

    const int64_t n_tiles = (n_local + SPT - 1) / SPT;
    load_rows(0);
    for (int64_t it = 0; it < NST && it < n_local; ++it) produce(it);
    for (int64_t it = 0; it < n_local; ++it) {
      const int64_t T = it / SPT;
      const int s = (int)(it - T * SPT);
      if (it + NST < n_local) produce(it + NST);
      if (s == 0 && T >= 1) softmax(T - 1);
      if (s == SPT - 1 || it == n_local - 1) issue_mma(T);
    }
    softmax(n_tiles - 1);
  } else if (warp > NH) {
    // =========================== warps 16..19: additive epilogue, one TMEM lane quarter each ===========================
    // D^T tile of a context tile: [2 M-tiles][128 hidden units (lanes)][64 positions (columns)].  A warp reads its 32
    // lanes of both M-tiles in 32-column pieces: tanh(. + b_a) * q_a, a butterfly reduce-scatter over the 32 hidden
    // units (lane l ends with position l), one partial logit row per (M-tile, lane quarter).
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_EPI));
    const int q4 = warp & 3;
    const int n_m = q4 < 3 ? 2 : 1;                            // hidden units 224..255 do not exist
    float qa_n[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const int en = 128 * m + 32 * q4 + lane;
      qa_n[m] = en < QD ? qa[en] : 0.f;
    }
    const int g = lane >> 2, t = lane & 3, mi = lane >> 3, rr = lane & 7;
    const int64_t n_tiles = (n_local + SPT - 1) / SPT;
    for (int64_t Tp = 0; Tp < n_tiles; ++Tp) {
      const uint32_t pb = (uint32_t)(Tp & 1);
      float* prow = part + pb * 512 + q4 * 64 + lane;
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        if (m < n_m) {
          tc::mbar_wait(mma_done + 8 * m, (uint32_t)(Tp & 1));      // M-tile m of the additive GEMM has landed
          tc::tc_fence_after();
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t xr[32];
            tmem_ld32_nw(tmem_base + TM_D + 64 * m + 32 * ch + ((uint32_t)(32 * q4) << 16), xr);
            tc::tmem_ld_wait();
            if (m == n_m - 1 && ch == 1) {                     // last read of D: the next additive MMA may overwrite it
              tc::tc_fence_before();
              __syncwarp();
              if (lane == 0) tc::mbar_arrive(d_free);
            }
            float v[32];
            if (dbg & 8) {
              v[0] = __uint_as_float(xr[0] & 0x3f800000u);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fast_tanh(__uint_as_float(xr[i])) * qa_n[m];   // b_a rides in the GEMM
#pragma unroll
              for (int sft = 0; sft < 5; ++sft) {
                const int o = 16 >> sft;
                const bool upper = (lane & o) != 0;
#pragma unroll
                for (int i = 0; i < o; ++i) {
                  const float send = upper ? v[i] : v[i + o];
                  const float keep = upper ? v[i + o] : v[i];
                  v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
              }
            }
            prow[m * 256 + 32 * ch] = v[0];
          }
        }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(part_full + 8 * pb);
      // ---- pooling of the same tile: out = sum_i w_i c_i on mma.sync.  A = the weight fragments warp 15 leaves in
      //      shared memory, B = the context tile read back with transposing ldmatrix; this warp takes the 8-column
      //      units j = q4, q4 + 4, ...
      tc::mbar_wait(w_ready + 8 * pb, (uint32_t)((Tp >> 1) & 1));
      if (!(dbg & 4)) {
        const uint4* af = afrag + pb * 128;
        const uint4 a0 = af[lane], a1 = af[32 + lane], a2 = af[64 + lane], a3 = af[96 + lane];
        const uint32_t rowa = sbase + C::OFF_CTX + pb * CTX_TILE + (uint32_t)((8 * mi + rr) * 128);
        float* orow = nullptr;
        if (g < SPT && Tp * SPT + g < n_local) orow = out + (blockIdx.x + (Tp * SPT + g) * gridDim.x) * (int64_t)D + 2 * t;
#pragma unroll 2
        for (int j = q4; j < 38; j += 4) {              // 16-byte unit = 8 context columns
          const uint32_t ua = rowa + (uint32_t)((j >> 3) * CTX_CHUNK) + (uint32_t)((((j & 7) ^ rr)) << 4);
          uint32_t b0, b1, b2, b3, b4, b5, b6, b7;
          ldsm_x4_t(ua, b0, b1, b2, b3);                  // positions 0..31
          ldsm_x4_t(ua + 32 * 128, b4, b5, b6, b7);       // positions 32..63
          float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
          mma_k16(d0, a0.x, a0.y, a0.z, a0.w, b0, b1);
          mma_k16(d1, a1.x, a1.y, a1.z, a1.w, b2, b3);
          mma_k16(d0, a2.x, a2.y, a2.z, a2.w, b4, b5);
          mma_k16(d1, a3.x, a3.y, a3.z, a3.w, b6, b7);
          if (orow != nullptr && 8 * j + 2 * t < D)
            *reinterpret_cast<float2*>(orow + 8 * j) = make_float2((d0[0] + d1[0]) + (d0[2] + d1[2]), (d0[1] + d1[1]) + (d0[3] + d1[3]));
        }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(ctx_free + 8 * pb);
    }
  } else {
    // =========================== warps 0..14: one head each ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_HEAD));
    const int hg = warp / HG, hl = warp - hg * HG;
    const int qoff = hg * GROUP + hl * SLICE, koff = OFF_K + qoff, voff = OFF_V + qoff;
    const int g = lane >> 2, t = lane & 3;
    const int mi = lane >> 3, rr = lane & 7;     // ldmatrix: this lane supplies row rr of matrix mi
    // row-shifted form?  (uniform over the grid: every thread reads the same 30 bounds)
    bool safe = force_safe > 0;
    if (force_safe < 0 && qk_bound != nullptr) {
#pragma unroll 1
      for (int h = 0; h < NH; ++h) safe = safe || !(qk_bound[h] * qk_bound[NH + h] <= SAFE_BOUND_SQ);
    }
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
    // context-tile byte offsets of this lane's three column pairs (columns 20 h + 8 dt + 2 t, +1), without the row
    // term and the swizzle: chunk base | 16-byte unit | offset inside the unit
    uint32_t cchunk[3], cunit[3], cin[3];
#pragma unroll
    for (int dt = 0; dt < 3; ++dt) {
      const int bc = 2 * (DH * warp + 8 * dt + 2 * t);
      cchunk[dt] = (uint32_t)((bc >> 7) * CTX_CHUNK);
      cunit[dt] = (uint32_t)((bc & 127) >> 4);
      cin[dt] = (uint32_t)(bc & 15);
    }
    for (int64_t it = 0; it < n_local; ++it) {
      const int64_t T = it / SPT;
      const int s = (int)(it - T * SPT);
      const uint32_t b = (uint32_t)(T & 1);
      const uint32_t st = (uint32_t)(it % NST);
      const uint32_t B = sbase + C::OFF_STAGE + st * STG;
      tc::mbar_wait(full_bar + 8 * st, (uint32_t)((it / NST) & 1));
      // K fragments: kb16[nt][0..1] (dims 0-7, 8-15), kb8[nt] (dims 16-23); V fragments (transposed loads):
      // vb[ks][dt][0..1] = keys 16ks..+7 / +8..15, dims 8dt..8dt+7; Q fragments of every query tile.
      uint32_t kb16[2 * KP][2], kb8[4 * K8P], vb[VS][3][2], qf[MT][6];
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
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        ldsm_x4(B + q4o[mt], qf[mt][0], qf[mt][1], qf[mt][2], qf[mt][3]);
        ldsm_x2(B + q2o[mt], qf[mt][4], qf[mt][5]);
      }
      // everything this warp needs from the stage is in registers: hand it back (the next sequence streams in behind
      // the attention of this one)
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(empty_bar + 8 * st);
      // the context tile of this sequence: pooled and free again?
      if (s == 0 && T >= 2) tc::mbar_wait(ctx_free + 8 * b, (uint32_t)(((T >> 1) - 1) & 1));
      const int row0 = s * SEQ;
      const uint32_t xorv = (uint32_t)((row0 + g) & 7);
      uint8_t* const crow = sm + C::OFF_CTX + b * CTX_TILE + (row0 + g) * 128;
      uint32_t coff[3];
#pragma unroll
      for (int dt = 0; dt < 3; ++dt) coff[dt] = cchunk[dt] + ((cunit[dt] ^ xorv) << 4) + cin[dt];

#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if (dbg & 32) continue;
        float sc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
          mma_k16(sc[nt], qf[mt][0], qf[mt][1], qf[mt][2], qf[mt][3], kb16[nt][0], kb16[nt][1]);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_k8(sc[nt], qf[mt][4], qf[mt][5], kb8[nt]);
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
          m0 = fminf(fmaxf(m0, -120.f), 120.f);      // keeps 2^-m finite; scores beyond +-120 are outside fp32 exp anyway
          m1 = fminf(fmaxf(m1, -120.f), 120.f);
          e0 = 1e-8f * ex2f(-m0);
          e1 = 1e-8f * ex2f(-m1);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            float p0 = ex2f(sc[nt][0] - m0), p1 = ex2f(sc[nt][1] - m0);
            float p2 = 0.f, p3 = 0.f;
            if (lower) { p2 = ex2f(sc[nt][2] - m1); p3 = ex2f(sc[nt][3] - m1); }
            if (nt == NT - 1 && 2 * t >= C::REM) { p0 = p1 = p2 = p3 = 0.f; }
            pa[nt][0] = pack_h2(p0, p1);
            pa[nt][1] = pack_h2(p2, p3);
          }
        } else {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            float p0 = ex2f(sc[nt][0]), p1 = ex2f(sc[nt][1]);
            float p2 = 0.f, p3 = 0.f;
            if (lower) { p2 = ex2f(sc[nt][2]); p3 = ex2f(sc[nt][3]); }
            if (nt == NT - 1 && 2 * t >= C::REM) { p0 = p1 = p2 = p3 = 0.f; }
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
        // O / (Z + eps) -> fp16, straight into the swizzled context tile (rows row0 + 16 mt + g and + 8)
        uint8_t* const c0 = crow + mt * (16 * 128);
        const bool ok0 = 16 * mt + g < SEQ, ok1 = 16 * mt + g + 8 < SEQ;
#pragma unroll
        for (int dt = 0; dt < 3; ++dt) {
          if (dt < 2 || t < 2) {
            if (ok0 && !(dbg & 16)) *reinterpret_cast<uint32_t*>(c0 + coff[dt]) = pack_h2(oacc[dt][0] * i0, oacc[dt][1] * i0);
            if (ok1 && !(dbg & 16)) *reinterpret_cast<uint32_t*>(c0 + 8 * 128 + coff[dt]) = pack_h2(oacc[dt][2] * i1, oacc[dt][3] * i1);
          }
        }
      }
      if (s == SPT - 1 || it == n_local - 1) {     // tile complete: publish it to the tensor core (async proxy)
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(ctx_full + 8 * b);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == NH) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace k1f

// bound: 30 floats (zeroed here, then max |q_h|^2 / |k_h|^2 over the table)
int k1f_qk_bound(const void* table16, int64_t n_rows, float* bound, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(bound, 0, 32 * sizeof(float), st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(qk bound)");
  if (n_rows <= 0) return NRMS_OK;
  int64_t blocks = (n_rows + 7) / 8;
  if (blocks > (int64_t)num_sms() * 8) blocks = (int64_t)num_sms() * 8;
  k1f::qk_bound_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const __half*>(table16), n_rows,
                                                          reinterpret_cast<unsigned int*>(bound));
  NRMS_LAUNCH_CHECK("qk_bound_kernel");
  return NRMS_OK;
}

static int g_k1f_debug = 0;
void set_k1f_debug(int v) { g_k1f_debug = v; }
static int g_force_safe = -1;     // "attn_safe_softmax" option: -1 auto (qk bound), 0 plain, 1 row-shifted
void set_attn_safe_softmax(int v) { g_force_safe = v < 0 ? -1 : (v ? 1 : 0); }
int get_attn_safe_softmax() { return g_force_safe; }

template <int SEQ, typename IdxT>
static int launch_attn_pool(const void* table16, int64_t n_table_rows, const void* rows, int64_t n_seq, const void* wa16,
                            const float* ba, const float* qa, const float* bound, float* out, cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  static bool configured[64] = {false};
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k1f::attn_pool_kernel<SEQ, IdxT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         k1f::Cfg<SEQ>::SMEM_BYTES);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn_pool_kernel)");
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  if (n_seq <= 0) return NRMS_OK;
  NRMS_CHECK_ARG(n_table_rows > 0, NRMS_E_INVALID, "table row count out of range");
  int grid = num_sms();
  if (n_seq < grid) grid = (int)n_seq;
  k1f::attn_pool_kernel<SEQ, IdxT><<<grid, k1f::THREADS, k1f::Cfg<SEQ>::SMEM_BYTES, st>>>(
      reinterpret_cast<const __half*>(table16), n_table_rows, reinterpret_cast<const IdxT*>(rows), n_seq,
      reinterpret_cast<const __half*>(wa16), ba, qa, bound, (g_force_safe < 0 && bound == nullptr) ? 1 : g_force_safe,
      g_k1f_debug, out);
  NRMS_LAUNCH_CHECK("attn_pool_kernel");
  return NRMS_OK;
}

// idx_kind 1 = int64 ids, 2 = int32 rows (as in tc_encoder_fused)
int k1f_run(int S, int idx_kind, const void* table16, int64_t n_table_rows, const void* rows, int64_t n_seq,
            const void* wa16, const float* ba, const float* qa, const float* bound, float* out, cudaStream_t st) {
  if (S == 50 && idx_kind == 2) return launch_attn_pool<50, int32_t>(table16, n_table_rows, rows, n_seq, wa16, ba, qa, bound, out, st);
  if (S == 50 && idx_kind == 1) return launch_attn_pool<50, int64_t>(table16, n_table_rows, rows, n_seq, wa16, ba, qa, bound, out, st);
  if (S == 20 && idx_kind == 1) return launch_attn_pool<20, int64_t>(table16, n_table_rows, rows, n_seq, wa16, ba, qa, bound, out, st);
  if (S == 20 && idx_kind == 2) return launch_attn_pool<20, int32_t>(table16, n_table_rows, rows, n_seq, wa16, ba, qa, bound, out, st);
  set_error("attn_pool_kernel: unsupported (S, index kind) = (%d, %d)", S, idx_kind);
  return NRMS_E_UNSUPPORTED;
}

}  // namespace nrms
