// Tensor-mode TRAINING attention over 20-token titles on the tensor cores (mma.sync.m16n8k8 TF32, fp32 accumulate):
// forward and backward of ScaledDotProductAttention inside MultiHeadSelfAttention
// (reference src/model/general/attention/multihead_self.py:15-23 and its autograd), one warp per (title, head).
//
// Why mma.sync and not tcgen05 here: per (title, head) the five contractions of the backward are 20x20x20 each --
// 1/13 of one 128x128x16 tcgen05 tile -- and their operands chain through per-row softmax arithmetic in registers.  The
// CUDA-core kernels this replaces (encoder_kernels.cuh, still the FP32-mode path) spent one shared-memory load per four
// FMAs and ran at 0.8 ms per 7,040 titles in the backward (ncu launch list, profiles/), 20 % of the training step.
//
// Layout.  CTA = (title, chunk of 5 heads), one warp per head.  Per head four shared-memory tiles Q, K, V, G(=dO) of
// [24 rows][28 floats]: rows 20..23 and columns 20..27 stay zero (they pad the 20-long dimensions to MMA shapes: rows to
// 2 x m16 / 3 x n8, the head dimension to 3 x k8), values are rounded to TF32 (cvt.rna) when staged.  The row pitch of 28
// words makes every fragment load conflict-free (pitch mod 32 = 28: the 8 row groups of a fragment land on distinct
// 4-bank groups).
//
//   S = Q K^T            A = Q rows, B = K rows                      ("NT": both operands read along the head dimension)
//   P = exp(S / sqrt 20), attn = P / (sum_j P + 1e-8)                 in the accumulator fragments, quad shuffles
//   O = attn V           A = the attn accumulator fragments, used in place: an accumulator fragment holds columns
//                        (2t, 2t+1) of an 8-column block where an A fragment wants (t, t+4); the contraction index is a
//                        dummy, so V's rows are simply read in the matching order (8b+2t, 8b+2t+1) -- no shuffles
//   backward: dP = G V^T (NT), dS = attn (dP - sum_j attn dP) / sqrt 20, dQ = dS K (fragments in place, like O),
//             dK = dS^T Q and dV = attn^T G read dS / attn back TRANSPOSED from shared memory: they are parked in the V
//             and K tiles, which are dead by then (their padding stays zero: only the 20x20 block is written).
// Global loads: warp-private 16-byte async copies, no block-wide barrier in the title loop (see warp_fetch).
#include "common.cuh"
#include "tc_api.cuh"

namespace nrms {
namespace amma {

constexpr int S = 20;            // tokens per title
constexpr int RT = 24;           // tile rows (3 x 8)
constexpr int ST = 28;           // tile row pitch in floats
constexpr int TILE = RT * ST;    // floats per tile
constexpr int HC = 5;            // heads per CTA (one warp each)
constexpr int THREADS = HC * 32;
constexpr float SQRT_DH = 4.47213595499957939f;
constexpr float ATTN_EPS = 1e-8f;
constexpr uint32_t DROPOUT_STREAM_CTX = 2;     // encoder_kernels.cuh
constexpr int NBUF = 1;          // one set of tiles per warp (a second set, filled under the compute, halved the resident warps: slower)

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// operand load.  The tiles arrive raw (cp.async); their warp rounds them to TF32 in place once (warp_fetch: cvt.rna; a
// bare fp32 pattern would be truncated by the tensor core) -- one conversion per element instead of one per fragment load
__device__ __forceinline__ uint32_t ld(const float* p) { return __float_as_uint(*p); }
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

#ifdef NRMS_AMMA_TRACE
#define AT(k) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) tr[k] = clock64(); } while (0)
#else
#define AT(k)
#endif

typedef float Frag[2][3][4];     // [m tile][n tile][accumulator register]

__device__ __forceinline__ void zero(Frag& c) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) c[mt][nt][e] = 0.f;
}

// c[i][j] += sum_d A[i][d] B[j][d]        (A, B: tiles)
__device__ __forceinline__ void gemm_nt(Frag& c, const float* A, const float* B, int g, int t) {
#pragma unroll
  for (int ks = 0; ks < 3; ++ks) {
    uint32_t a[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      a[mt][0] = ld(A + r0 * ST + ks * 8 + t);
      a[mt][2] = ld(A + r0 * ST + ks * 8 + t + 4);
      a[mt][1] = r1 < RT ? ld(A + r1 * ST + ks * 8 + t) : 0u;
      a[mt][3] = r1 < RT ? ld(A + r1 * ST + ks * 8 + t + 4) : 0u;
    }
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      const uint32_t b0 = ld(B + (nt * 8 + g) * ST + ks * 8 + t), b1 = ld(B + (nt * 8 + g) * ST + ks * 8 + t + 4);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt][nt], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b0, b1);
    }
  }
}

// c[i][d] += sum_j P[i][j] B[j][d]        (P: accumulator fragments used in place as the A operand, B: tile)
__device__ __forceinline__ void gemm_frag_n(Frag& c, const Frag& p, const float* B, int g, int t) {
#pragma unroll
  for (int ks = 0; ks < 3; ++ks) {
    uint32_t a[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      a[mt][0] = __float_as_uint(tf32_round(p[mt][ks][0]));   // (row g,   contraction slot t)     <- column 8ks + 2t
      a[mt][2] = __float_as_uint(tf32_round(p[mt][ks][1]));   // (row g,   contraction slot t + 4) <- column 8ks + 2t + 1
      a[mt][1] = __float_as_uint(tf32_round(p[mt][ks][2]));   // (row g+8, slot t)
      a[mt][3] = __float_as_uint(tf32_round(p[mt][ks][3]));   // (row g+8, slot t + 4)
    }
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      const uint32_t b0 = ld(B + (ks * 8 + 2 * t) * ST + nt * 8 + g), b1 = ld(B + (ks * 8 + 2 * t + 1) * ST + nt * 8 + g);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt][nt], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b0, b1);
    }
  }
}

// c[j][d] += sum_i T[i][j] B[i][d]        (T, B: tiles; T is read transposed)
__device__ __forceinline__ void gemm_tn(Frag& c, const float* T, const float* B, int g, int t) {
#pragma unroll
  for (int ks = 0; ks < 3; ++ks) {
    const int i0 = ks * 8 + t, i1 = i0 + 4;
    uint32_t a[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int j0 = mt * 16 + g, j1 = j0 + 8;
      a[mt][0] = ld(T + i0 * ST + j0);
      a[mt][2] = ld(T + i1 * ST + j0);
      a[mt][1] = j1 < RT ? ld(T + i0 * ST + j1) : 0u;
      a[mt][3] = j1 < RT ? ld(T + i1 * ST + j1) : 0u;
    }
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      const uint32_t b0 = ld(B + i0 * ST + nt * 8 + g), b1 = ld(B + i1 * ST + nt * 8 + g);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt][nt], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b0, b1);
    }
  }
}

// raw scores -> attention weights in place: exp(s / sqrt 20) for key columns < 20, 0 for the padding columns,
// divided by (row sum + 1e-8)  (multihead_self.py:16-20: no max subtraction)
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void softmax_rows(Frag& c, int t) {
  constexpr float SCALE_LOG2E = 1.4426950408889634f / SQRT_DH;     // exp(s / sqrt 20) = 2^(s * log2(e) / sqrt 20)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = nt * 8 + 2 * t + e;
          const float v = col < S ? ex2(c[mt][nt][half * 2 + e] * SCALE_LOG2E) : 0.f;
          c[mt][nt][half * 2 + e] = v;
          sum += v;
        }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = 1.f / (sum + ATTN_EPS);
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) c[mt][nt][half * 2 + e] *= inv;
    }
}

// store the 20 x 20 block of a fragment into a tile as [row][col] (TF32-rounded); padding rows / columns are not touched
__device__ __forceinline__ void park(float* T, const Frag& c, int g, int t) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int row = mt * 16 + g + 8 * half, col = nt * 8 + 2 * t;
        if (row < S && col < S)
          *reinterpret_cast<float2*>(T + row * ST + col) =
              make_float2(tf32_round(c[mt][nt][half * 2]), tf32_round(c[mt][nt][half * 2 + 1]));
      }
}

// write the 20 x 20 block of a fragment to global rows (pitch ld floats), two floats per store
__device__ __forceinline__ void store_rows(float* dst, int64_t ld_, const Frag& c, int g, int t) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int row = mt * 16 + g + 8 * half, col = nt * 8 + 2 * t;
        if (row < S && col < S)
          *reinterpret_cast<float2*>(dst + row * ld_ + col) = make_float2(c[mt][nt][half * 2], c[mt][nt][half * 2 + 1]);
      }
}

// the same 20 x 20 block written TRANSPOSED: dstT[col * ldt + row].  The weight-gradient contraction dW = dQKV^T X reads
// dQKV^T as a K-major operand (tc_gemm.cu); writing it here replaces a 507 MB read + 507 MB write transposition pass per
// step.  The fragment is turned through a dead tile (scratch[col][row], 20 x 20 block only: the padding stays zero) and
// leaves as 16-byte stores: a column of the block is 20 consecutive floats = 80 contiguous, 16-byte aligned bytes of a
// row of dstT.  (Storing the fragments transposed directly -- 72 scalar stores per lane and matrix -- took the kernel
// from 383 to 804 us and 215 registers.)
__device__ __forceinline__ void store_rows_t(float* dstT, int64_t ldt, float* scratch, const Frag& c, int lane, int g, int t) {
  __syncwarp();
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int row = mt * 16 + g + 8 * half, col = nt * 8 + 2 * t;
        if (row < S && col < S) {
          scratch[col * ST + row] = c[mt][nt][half * 2];
          scratch[(col + 1) * ST + row] = c[mt][nt][half * 2 + 1];
        }
      }
  __syncwarp();
  for (int f = lane; f < S * (S / 4); f += 32) {
    const int col = f / (S / 4), r4 = f % (S / 4);
    *reinterpret_cast<float4*>(dstT + (int64_t)col * ldt + 4 * r4) = *reinterpret_cast<const float4*>(scratch + col * ST + 4 * r4);
  }
  __syncwarp();
}

// Warp-private staging: each warp copies the 20 x 20 blocks of its own head's NTILES tiles (tile k <- columns
// [k*300 + head*20, +20) of the title's 20 qkv rows for k < 3, of its d_ctx rows for k == 3) with 16-byte async copies,
// waits for them, and rounds them to TF32 in place (cvt.rna; a bare fp32 pattern would be truncated by the tensor core;
// the d_ctx tile also takes the dropout-2 mask of the forward's Philox stream here).  Every lane converts exactly the 16-byte
// pieces it copied itself, so there is no barrier between the copy and the conversion, and no block-wide barrier at all:
// the warps of a CTA drift apart and hide each other's copy latency.  (The first version staged cooperatively with two
// __syncthreads per title: clock64 stamps showed 4.2k cycles at the barrier, 3k waiting for the copies and 4.5k in a
// rolled conversion loop, against 5.7k of contractions.)
template <int NTILES>
__device__ __forceinline__ void warp_fetch(float* tiles, const float* __restrict__ qkv_head, const float* __restrict__ g_head,
                                           int lane, int64_t row0, int col0, float p, float scale, uint64_t seed,
                                           uint64_t offset) {
  constexpr int PT = 4;                       // 100 sixteen-byte pieces per tile over 32 lanes
  int j[PT], d4[PT];
#pragma unroll
  for (int k = 0; k < PT; ++k) {
    const int f = lane + 32 * k;
    j[k] = f / (DH / 4);
    d4[k] = f % (DH / 4);
  }
#pragma unroll
  for (int w = 0; w < NTILES; ++w)
#pragma unroll
    for (int k = 0; k < PT; ++k)
      if (lane + 32 * k < S * (DH / 4)) {
        const float* src = (w < 3) ? qkv_head + (int64_t)j[k] * D3 + w * D + d4[k] * 4 : g_head + (int64_t)j[k] * D + d4[k] * 4;
        cp_async16(tiles + w * TILE + j[k] * ST + d4[k] * 4, src);
      }
  cp_async_wait_all();
#pragma unroll
  for (int w = 0; w < NTILES; ++w) {
    float4 x[PT];
#pragma unroll
    for (int k = 0; k < PT; ++k)
      if (lane + 32 * k < S * (DH / 4)) x[k] = *reinterpret_cast<const float4*>(tiles + w * TILE + j[k] * ST + d4[k] * 4);
    if (w == 3 && p > 0.f) {
#pragma unroll
      for (int k = 0; k < PT; ++k)
        if (lane + 32 * k < S * (DH / 4)) {
          const float4 m = dropout_mask4((uint64_t)(row0 + j[k]) * D + col0 + d4[k] * 4, DROPOUT_STREAM_CTX, p, scale, seed, offset);
          x[k].x *= m.x; x[k].y *= m.y; x[k].z *= m.z; x[k].w *= m.w;
        }
    }
#pragma unroll
    for (int k = 0; k < PT; ++k)
      if (lane + 32 * k < S * (DH / 4))
        *reinterpret_cast<float4*>(tiles + w * TILE + j[k] * ST + d4[k] * 4) =
            make_float4(tf32_round(x[k].x), tf32_round(x[k].y), tf32_round(x[k].z), tf32_round(x[k].w));
  }
  __syncwarp();
}

__global__ void __launch_bounds__(THREADS)
attn_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ ctx, int64_t n_seq, float p, float scale, uint64_t seed,
                uint64_t offset) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int hc = blockIdx.y;
  for (int f = tid; f < HC * 3 * TILE; f += THREADS) smem[f] = 0.f;
  __syncthreads();
  float* Q = smem + (warp * 3 + 0) * TILE;
  const float* K = Q + TILE;
  const float* V = K + TILE;
  const int colbase = hc * (HC * DH) + warp * DH;
  for (int64_t seq = blockIdx.x; seq < n_seq; seq += gridDim.x) {
    __syncwarp();             // every lane is done with the previous title's tiles
    warp_fetch<3>(Q, qkv + seq * S * D3 + colbase, nullptr, lane, seq * S, colbase, 0.f, 1.f, 0, 0);
    Frag a;
    zero(a);
    gemm_nt(a, Q, K, g, t);
    softmax_rows(a, t);
    Frag o;
    zero(o);
    gemm_frag_n(o, a, V, g, t);
    if (p > 0.f) {
      // dropout-2 masks: 100 aligned groups of four columns = 100 Philox evaluations, 4 per lane, written over the dead
      // Q block (rows / columns < 20 only: the padding stays zero) and read back in accumulator order
      __syncwarp();
      for (int f = lane; f < S * (DH / 4); f += 32) {
        const int j = f / (DH / 4), d4 = f % (DH / 4);
        *reinterpret_cast<float4*>(Q + j * ST + d4 * 4) =
            dropout_mask4((uint64_t)(seq * S + j) * D + colbase + d4 * 4, DROPOUT_STREAM_CTX, p, scale, seed, offset);
      }
      __syncwarp();
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int row = mt * 16 + g + 8 * half, col = nt * 8 + 2 * t;
            if (row < S && col < S) {
              const float2 m = *reinterpret_cast<const float2*>(Q + row * ST + col);
              o[mt][nt][half * 2] *= m.x;
              o[mt][nt][half * 2 + 1] *= m.y;
            }
          }
    }
    store_rows(ctx + seq * S * D + colbase, D, o, g, t);
  }
}

__global__ void __launch_bounds__(THREADS)
attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ d_ctx, float* __restrict__ d_qkv,
                float* __restrict__ d_qkv_t, int64_t ldt, int64_t n_seq, float p, float scale, uint64_t seed, uint64_t offset) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int hc = blockIdx.y;
  for (int f = tid; f < HC * 4 * TILE; f += THREADS) smem[f] = 0.f;
  __syncthreads();
  constexpr float INV_SQRT_DH = 1.f / SQRT_DH;
  float* Q = smem + (warp * 4 + 0) * TILE;
  float* K = Q + TILE;
  float* V = K + TILE;
  float* G = V + TILE;
  const int colbase = hc * (HC * DH) + warp * DH;
#ifdef NRMS_AMMA_TRACE
  long long tr[12] = {0};
#endif
  for (int64_t seq = blockIdx.x; seq < n_seq; seq += gridDim.x) {
    AT(0);
    __syncwarp();             // every lane is done with the previous title's tiles
    AT(1);
    AT(2);
    warp_fetch<4>(Q, qkv + seq * S * D3 + colbase, d_ctx + seq * S * D + colbase, lane, seq * S, colbase, p, scale, seed, offset);
    AT(3);
    Frag a, dp;
    zero(a);
    gemm_nt(a, Q, K, g, t);
    AT(4);
    softmax_rows(a, t);                       // a = attn
    AT(5);
    zero(dp);
    gemm_nt(dp, G, V, g, t);                  // dp = dO V^T
    AT(6);
    // ds = attn (dp - sum_j attn dp) / sqrt 20, in place over dp
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float delta = 0.f;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) delta = fmaf(a[mt][nt][half * 2 + e], dp[mt][nt][half * 2 + e], delta);
        delta += __shfl_xor_sync(0xffffffffu, delta, 1);
        delta += __shfl_xor_sync(0xffffffffu, delta, 2);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            dp[mt][nt][half * 2 + e] = a[mt][nt][half * 2 + e] * (dp[mt][nt][half * 2 + e] - delta) * INV_SQRT_DH;
      }
    float* out = d_qkv + seq * S * D3 + colbase;
    float* out_t = d_qkv_t ? d_qkv_t + (int64_t)colbase * ldt + seq * S : nullptr;     // [900, ldt]: row = column of dQKV
    __syncwarp();
    park(V, dp, g, t);                        // V is dead: dS (TF32-rounded) parked as [i][j]
    Frag r;
    zero(r);
    gemm_frag_n(r, dp, K, g, t);              // dQ = dS K
    store_rows(out, D3, r, g, t);
    if (out_t) store_rows_t(out_t, ldt, K, r, lane, g, t);       // K is dead from here on
    AT(7);
    __syncwarp();
    // K is dead: the rounding residue dS - tf32(dS) goes there.  The rows of dS sum to ~0 (the softmax Jacobian), and
    // the K-bias gradient is exactly that sum: with dS = hi + lo in the dK contraction it cancels to fp32 accuracy
    // instead of leaving 2^-11-sized rounding noise.
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) dp[mt][nt][e] -= tf32_round(dp[mt][nt][e]);
    park(K, dp, g, t);
    __syncwarp();
    zero(r);
    gemm_tn(r, V, Q, g, t);                   // dK = (dS_hi + dS_lo)^T Q
    gemm_tn(r, K, Q, g, t);
    store_rows(out + D, D3, r, g, t);
    if (out_t) store_rows_t(out_t + (int64_t)D * ldt, ldt, K, r, lane, g, t);       // dS_lo in K is dead
    AT(8);
    __syncwarp();
    park(V, a, g, t);                         // dS is dead: attn parked as [i][j]
    __syncwarp();
    zero(r);
    gemm_tn(r, V, G, g, t);                   // dV = attn^T dO
    store_rows(out + 2 * D, D3, r, g, t);
    if (out_t) store_rows_t(out_t + (int64_t)2 * D * ldt, ldt, K, r, lane, g, t);
    AT(9);
#ifdef NRMS_AMMA_TRACE
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && seq == (int64_t)gridDim.x * 2) {
      printf("attn_bwd sync+issue/wait/round/S/softmax/dP/dS+dQ/dK/dV cycles:");
      for (int k_ = 1; k_ < 10; ++k_) printf(" %lld", tr[k_] - tr[k_ - 1]);
      printf("\n");
    }
#endif
  }
}

}  // namespace amma

int attn_mma_fwd(const float* qkv, float* ctx, int64_t n_seq, float p, uint64_t seed, uint64_t offset, cudaStream_t st) {
  const size_t smem = amma::NBUF * (size_t)amma::HC * 3 * amma::TILE * sizeof(float);
  static bool configured[64] = {false};
  if (cudaError_t e = set_max_dynamic_smem(amma::attn_fwd_kernel, (int)smem, configured)) return cuda_fail(e, "attn_mma_fwd attr");
  int64_t gx = n_seq < (int64_t)num_sms() * 8 ? n_seq : (int64_t)num_sms() * 8;
  const float scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  amma::attn_fwd_kernel<<<dim3((unsigned)gx, H / amma::HC), amma::THREADS, smem, st>>>(qkv, ctx, n_seq, p, scale, seed, offset);
  NRMS_LAUNCH_CHECK("attn_mma_fwd");
  return NRMS_OK;
}

int attn_mma_bwd(const float* qkv, const float* d_ctx, float* d_qkv, float* d_qkv_t, int64_t ldt, int64_t n_seq, float p,
                 uint64_t seed, uint64_t offset, cudaStream_t st) {
  const size_t smem = amma::NBUF * (size_t)amma::HC * 4 * amma::TILE * sizeof(float);
  static bool configured[64] = {false};
  if (cudaError_t e = set_max_dynamic_smem(amma::attn_bwd_kernel, (int)smem, configured)) return cuda_fail(e, "attn_mma_bwd attr");
  int64_t gx = n_seq < (int64_t)num_sms() * 8 ? n_seq : (int64_t)num_sms() * 8;
  const float scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  amma::attn_bwd_kernel<<<dim3((unsigned)gx, H / amma::HC), amma::THREADS, smem, st>>>(qkv, d_ctx, d_qkv, d_qkv_t, ldt, n_seq, p, scale,
                                                                                       seed, offset);
  NRMS_LAUNCH_CHECK("attn_mma_bwd");
  return NRMS_OK;
}

}  // namespace nrms
