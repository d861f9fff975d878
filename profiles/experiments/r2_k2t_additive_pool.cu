// EXPERIMENT (round 2) -- NOT part of the build.  Kept as the record of a design that was measured and lost:
//   K2t, additive pooling with W_a in tensor memory and whole context tiles in shared memory (every context row read once).
//   Parity-green (33 GPU tests) and slower than K2 on the evaluate bench: users stage 2.76 ms (8 epilogue warps), 2.79 ms
//   (16 epilogue warps), 2.90 ms (tile ingest split between the TMA unit and plain loads) against 2.57 ms with K2.
//   ncu (profiles/r2_ncu_prof_k2t.txt): 218 us per 18,944 users against 168 us for K2, tensor pipe 54 % active; one SM
//   takes a 40 KB tile in through the TMA unit in ~3,400 cycles (12 B/clk), the same per-SM ingest rate K1g's row gather
//   reaches -- K2 stays ahead because it splits its traffic between the TMA unit (GEMM operand) and plain loads (pooling).
// K2t -- additive-attention pooling over fp16 context rows with W_a RESIDENT IN TENSOR MEMORY (tensor-mode inference).
//
// Reference math: src/model/general/attention/additive.py:27-53 -- temp = tanh(Linear(c)), weights = softmax(temp . q)
// over the sequence, out = sum_i w_i c_i -- on the context rows [n_seq * S][320] (fp16, columns 300..319 zero) that the
// attention kernels K1g / K1 v6 write.
//
// The first pooling kernel (K2, tc_fused3.cu) keeps W_a (fp16, 133 KB) in SHARED memory, which leaves room for only a
// 64 KB ring of context chunks: the pool warps had to read every context row a second time from L2 once the softmax
// weights existed, and the kernel sat at the L2 bandwidth ceiling (1.2 GB per 18,944 users in 168 us).  Here the GEMM is
// transposed -- T^T = W_a C^T, A = W_a (M = hidden units, two M-tiles of 128) held in TMEM as packed fp16 for the whole
// kernel (304 of the 512 columns), B = the context tile (N = positions) in shared memory -- so shared memory holds a
// ring of FOUR WHOLE context tiles (64 rows x 5 chunks = 40 KB each) and the pooling reads the tile it needs from there:
// every context row crosses L2 -> SM once.
//
//   warp 0      TMA producer: 5 boxes (64 rows x 64 halfs, SWIZZLE_128B) per tile into the ring
//   warp 1      tcgen05 issuer (+ TMEM allocation): per tile two N = 32 halves, 2 M-tiles x 19 MMAs (kind::f16, A from
//               TMEM) each; the four D buffers [M-tile][half] of 32 columns fill the 128 columns left beside W_a, so the
//               epilogue of one half runs under the MMAs of the other
//   warps 2-17  epilogue (thread = hidden unit, warp = (half, M-tile, TMEM lane quarter)): tcgen05.ld 32 positions ->
//               tanh(. + b_a) * q_a -> butterfly reduce-scatter over the warp's 32 hidden units -> one partial logit row
//               per (M-tile, quarter)
//   warp 18     sums the seven partial rows in a fixed order, stable softmax per sequence (additive.py:37-39), leaves
//               the weights as mma.sync A fragments (fp16 head + fp16 remainder: ~22 bits)
//   warps 19-22 out = sum_i w_i c_i on mma.sync, B = the context tile read back from SHARED memory with transposing
//               ldmatrix; then the tile's ring slot goes back to the producer
// A tile = 64 context rows = one user (S = 50) or three titles (S = 20).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "tc_common.cuh"
#include "tc_api.cuh"

namespace nrms {

int make_tmap_k_major_f16(CUtensorMap* out, const void* base, int64_t rows, int cols, int64_t ld, int box_rows);

namespace k2t {

constexpr int CP = 320;                        // pitch (halfs) of the context rows and of the fp16 W_a copy
constexpr int ROWS = 64;                       // positions per tile = UMMA N (two halves of 32)
constexpr int CHUNK = ROWS * 128;              // 8,192 B: 64 rows x 64 halfs
constexpr int TILE = 5 * CHUNK;                // 40,960 B
constexpr int NSTG = 4;                        // ring of whole context tiles
constexpr int KSTEPS = 19;                     // 304 / 16
constexpr int TM_A = 0, TM_D = 304;            // W_a: M-tile m at columns 152 m; D[half][m] at 304 + 32 (2 half + m)
constexpr int N_EPI = 16, N_POOL = 4;
// A tile enters shared memory over TWO paths: chunks 0..TMA_CHUNKS-1 by bulk tensor copies, the rest by plain 16-byte
// loads + swizzled stores from the (otherwise mostly idle) epilogue warps.  Measured on this part: what one SM takes in
// through the TMA unit tops out near 15 B/clk for these access patterns (K1g's row gather, this kernel with all five
// chunks on the TMA: 3,370 cycles per 40 KB tile), and the load/store path adds its own bandwidth beside it.
constexpr int TMA_CHUNKS = 3;
constexpr int LD_PIECES = (5 - TMA_CHUNKS) * ROWS * 8;      // 16-byte pieces per tile on the load/store path (1,024)
constexpr int LD_AHEAD = 2;                                 // tiles the loaders run ahead of their epilogue work
constexpr int THREADS = 32 * (3 + N_EPI + N_POOL);      // 736
constexpr int OFF_TILES = 0;
constexpr int OFF_PART = NSTG * TILE;                   // [2][8][64] fp32 partial logits
constexpr int OFF_W = OFF_PART + 2 * 8 * 64 * 4;        // [64] fp32 weights (warp 10 only)
constexpr int OFF_AF = OFF_W + 256;                     // [2][4][32] uint4 pooling A fragments
constexpr int OFF_BAR = OFF_AF + 2 * 4 * 32 * 16;
constexpr int SMEM = OFF_BAR + 256 + 1024;
static_assert(SMEM <= 232448, "shared memory");

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
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_k16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                        uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
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

// S = sequence length (50: one user per tile, 20: three titles per tile)
template <int S>
__global__ void __launch_bounds__(THREADS, 1)
additive_pool_tmem_kernel(const __grid_constant__ CUtensorMap tmap_c, const __half* __restrict__ ctx,
                          const __half* __restrict__ wa16, const float* __restrict__ ba, const float* __restrict__ qa,
                          float* __restrict__ out, int64_t n_seq, uint32_t tile_tx_bytes) {
  constexpr int SPT = ROWS / S;                  // sequences per tile (1 / 3)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (sbase - raw);
  float* part = reinterpret_cast<float*>(sm + OFF_PART);      // [2][8][64]
  float* wsm = reinterpret_cast<float*>(sm + OFF_W);
  uint4* afrag = reinterpret_cast<uint4*>(sm + OFF_AF);       // [2][4][32]
  const uint32_t bars = sbase + OFF_BAR;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * NSTG;                 // [NSTG] each
  const uint32_t mma_done = bars + 16 * NSTG, d_free = mma_done + 32;           // [half][m] (4), [half] (2)
  const uint32_t part_full = d_free + 16, part_free = part_full + 16;           // [parity] each
  const uint32_t w_ready = part_free + 16, af_free = w_ready + 16;              // [parity] each
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(sm + OFF_BAR + 16 * NSTG + 128);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int64_t n_tiles = (n_seq + SPT - 1) / SPT;
  const int64_t n_local = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // tiles of this CTA

  if (tid == 0) {
    for (int s = 0; s < NSTG; ++s) {
      tc::mbar_init(full_bar + 8 * s, 1 + N_EPI);        // the TMA producer + one arrival per loader (epilogue) warp
      tc::mbar_init(empty_bar + 8 * s, N_POOL);
    }
    for (int i = 0; i < 4; ++i) tc::mbar_init(mma_done + 8 * i, 1);
    for (int h = 0; h < 2; ++h) tc::mbar_init(d_free + 8 * h, 7);          // 4 quarters of M-tile 0 + 3 of M-tile 1
    for (int p = 0; p < 2; ++p) {
      tc::mbar_init(part_full + 8 * p, 14);                                // 7 epilogue warps x 2 halves
      tc::mbar_init(part_free + 8 * p, 1);
      tc::mbar_init(w_ready + 8 * p, 1);
      tc::mbar_init(af_free + 8 * p, N_POOL);
    }
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc(tc::smem_u32((const void*)tmem_ptr_smem), 512);
  // ring + partial rows start as zeros: rows a short last box never writes must be finite
  for (int i = tid; i < (NSTG * TILE) / 16; i += THREADS) reinterpret_cast<uint4*>(sm + OFF_TILES)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < 2 * 8 * 64 + 64; i += THREADS) part[i] = 0.f;
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // ---- W_a -> tensor memory (once): lane = hidden unit, 152 packed-fp16x2 columns per M-tile; warps 0..7 = (M-tile, quarter) ----
  if (warp < 8) {
    const int q4 = warp & 3, m = warp >> 2;
    const int n = 128 * m + 32 * q4 + lane;
    const uint4* src = reinterpret_cast<const uint4*>(wa16 + (size_t)(n < QD ? n : 0) * CP);       // 608 B = 38 x 16 B
    const uint32_t tcol = tmem_base + TM_A + 152 * m + ((uint32_t)(32 * q4) << 16);
    uint32_t r[16];
#pragma unroll 1
    for (int c = 0; c < 9; ++c) {                      // 9 x 16 columns = 144
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (n < QD) v = __ldg(src + 4 * c + i);
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
      }
      tc::tmem_st16(tcol + 16 * c, r);
    }
    {                                                  // + 8 columns = 152
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (n < QD) v = __ldg(src + 36 + i);
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
      }
      tc::tmem_st8(tcol + 144, r);
    }
    tc::tmem_st_wait();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      for (int64_t it = 0; it < n_local; ++it) {
        const int64_t t = blockIdx.x + it * gridDim.x;
        const uint32_t s = (uint32_t)(it % NSTG);
        tc::mbar_wait(empty_bar + 8 * s, (uint32_t)((it / NSTG) & 1) ^ 1u);
        expect_tx(full_bar + 8 * s, tile_tx_bytes);
        const int r0 = (int)(t * SPT * S);
        for (int kc = 0; kc < TMA_CHUNKS; ++kc) tma_load_2d(sbase + OFF_TILES + s * TILE + kc * CHUNK, &tmap_c, kc * 64, r0, full_bar + 8 * s);
      }
    }
  } else if (warp == 1) {
    // ------------------------------ tcgen05 issuer (whole warp converged) ------------------------------
    const uint32_t el = tc::elect_one_u32();
    const uint32_t idesc = tc::umma_idesc_f16(128, 32);
    const uint64_t desc0 = tc::umma_desc_k_sw128(0);
    for (int64_t it = 0; it < n_local; ++it) {
      const uint32_t s = (uint32_t)(it % NSTG);
      tc::mbar_wait(full_bar + 8 * s, (uint32_t)((it / NSTG) & 1));
      const uint32_t cb = (sbase + OFF_TILES + s * TILE) >> 4;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (it > 0) tc::mbar_wait(d_free + 8 * h, (uint32_t)((it - 1) & 1));      // the epilogue has read this half of tile it-1
        tc::tc_fence_after();
#pragma unroll
        for (int m = 0; m < 2; ++m) {
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ++ks)
            tc::umma_f16_ts_p(tmem_base + TM_D + 32 * (2 * h + m), tmem_base + TM_A + 152 * m + 8 * ks,
                              desc0 | (uint64_t)((cb + (ks >> 2) * (CHUNK >> 4) + h * 256 + (ks & 3) * 2) & 0x3FFF), idesc,
                              ks ? 1u : 0u, el);
          tc::umma_commit_p(mma_done + 8 * (2 * h + m), el);
        }
      }
    }
  } else if (warp < 2 + N_EPI) {
    // ------------------------------ epilogue: warp = (half, M-tile, TMEM lane quarter) ------------------------------
    // One 32 x 32 unit (32 hidden units x 32 positions) per warp and tile: with eight warps doing two units each the
    // kernel was bound by their serial instruction streams (0.86 ms per 73,152 users against 0.67 ms for K2).
    const int q4 = warp & 3, e = (warp - 2) >> 2, m = e & 1, h = e >> 1;
    const bool valid = !(m == 1 && q4 == 3);                  // hidden units 224..255 do not exist
    const int en = 128 * m + 32 * q4 + lane;
    const float ba_n = en < QD ? ba[en] : 0.f, qa_n = en < QD ? qa[en] : 0.f;
    // loader role: this thread's 16-byte pieces of the chunks that do not come through the TMA unit
    const int lt = (warp - 2) * 32 + lane;                    // 0..511
    constexpr int PPT = LD_PIECES / (N_EPI * 32);             // pieces per thread and tile (2)
    const int64_t total_rows = n_seq * S;
    auto load_tile = [&](int64_t jt) {                        // tile jt of this CTA -> its ring slot
      const int64_t tile = blockIdx.x + jt * gridDim.x;
      const uint32_t sl = (uint32_t)(jt % NSTG);
      const int64_t r0 = tile * SPT * S;
      uint4 v[PPT];
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const int pc = lt + k * (N_EPI * 32);
        const int c = TMA_CHUNKS + pc / (ROWS * 8), rem = pc % (ROWS * 8), r = rem >> 3, u = rem & 7;
        v[k] = make_uint4(0u, 0u, 0u, 0u);
        if (r0 + r < total_rows) v[k] = __ldg(reinterpret_cast<const uint4*>(ctx + (r0 + r) * CP + c * 64 + u * 8));
      }
      tc::mbar_wait(empty_bar + 8 * sl, (uint32_t)((jt / NSTG) & 1) ^ 1u);     // the slot's previous tile has been pooled
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const int pc = lt + k * (N_EPI * 32);
        const int c = TMA_CHUNKS + pc / (ROWS * 8), rem = pc % (ROWS * 8), r = rem >> 3, u = rem & 7;
        *reinterpret_cast<uint4*>(sm + OFF_TILES + sl * TILE + c * CHUNK + r * 128 + ((u ^ (r & 7)) << 4)) = v[k];
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(full_bar + 8 * sl);
    };
    for (int64_t jt = 0; jt < LD_AHEAD && jt < n_local; ++jt) load_tile(jt);
    for (int64_t it = 0; it < n_local; ++it) {
      if (it + LD_AHEAD < n_local) load_tile(it + LD_AHEAD);
      if (valid) {
        const uint32_t p = (uint32_t)(it & 1);
        float* prow = part + p * 512 + (m * 4 + q4) * 64 + 32 * h + lane;
        tc::mbar_wait(mma_done + 8 * (2 * h + m), (uint32_t)(it & 1));
        tc::tc_fence_after();
        uint32_t xr[32];
        tmem_ld32_nw(tmem_base + TM_D + 32 * (2 * h + m) + ((uint32_t)(32 * q4) << 16), xr);
        tc::tmem_ld_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(d_free + 8 * h);
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fast_tanh(__uint_as_float(xr[i]) + ba_n) * qa_n;
        // butterfly reduce-scatter over the 32 hidden units of the warp: lane l ends with position 32 h + l
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
        if (it >= 2) tc::mbar_wait(part_free + 8 * p, (uint32_t)(((it >> 1) - 1) & 1));     // tile it-2's logits were consumed
        *prow = v[0];
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(part_full + 8 * p);
      }
    }
  } else if (warp == 2 + N_EPI) {
    // ------------------------------ softmax over each sequence of the tile -> pooling A fragments ------------------
    const int g = lane >> 2, t = lane & 3;
    for (int64_t it = 0; it < n_local; ++it) {
      const uint32_t p = (uint32_t)(it & 1);
      tc::mbar_wait(part_full + 8 * p, (uint32_t)((it >> 1) & 1));
      const float* pp = part + p * 512;
      float l0 = 0.f, l1 = 0.f;                       // logits of positions lane, lane + 32: seven partial rows, fixed order
#pragma unroll
      for (int un = 0; un < 7; ++un) { l0 += pp[un * 64 + lane]; l1 += pp[un * 64 + 32 + lane]; }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(part_free + 8 * p);
      float w0 = 0.f, w1 = 0.f;
#pragma unroll
      for (int s = 0; s < SPT; ++s) {
        const bool in0 = lane >= s * S && lane < (s + 1) * S;
        const bool in1 = lane + 32 >= s * S && lane + 32 < (s + 1) * S;
        const float mx = warp_max(fmaxf(in0 ? l0 : -INFINITY, in1 ? l1 : -INFINITY));
        const float e0 = in0 ? __expf(l0 - mx) : 0.f, e1 = in1 ? __expf(l1 - mx) : 0.f;
        const float inv = 1.f / warp_sum(e0 + e1);
        if (in0) w0 = e0 * inv;
        if (in1) w1 = e1 * inv;
      }
      wsm[lane] = w0;
      wsm[lane + 32] = w1;
      __syncwarp();
      if (it >= 2) tc::mbar_wait(af_free + 8 * p, (uint32_t)(((it >> 1) - 1) & 1));       // the pool warps are done with tile it-2's fragments
      uint4* af = afrag + p * 128;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[4];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int pos = 16 * ks + 8 * hf + 2 * t;
          float x0 = 0.f, x1 = 0.f;
          if (g < SPT) {
            if (pos >= g * S && pos < (g + 1) * S) x0 = wsm[pos];
            if (pos + 1 >= g * S && pos + 1 < (g + 1) * S) x1 = wsm[pos + 1];
          }
          const __half2 hi = __floats2half2_rn(x0, x1);
          const float2 hf2 = __half22float2(hi);
          a[2 * hf] = *reinterpret_cast<const uint32_t*>(&hi);
          a[2 * hf + 1] = pack_h2(x0 - hf2.x, x1 - hf2.y);
        }
        af[ks * 32 + lane] = make_uint4(a[0], a[1], a[2], a[3]);
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(w_ready + 8 * p);
    }
  } else {
    // ------------------------------ pool warps: out = sum_i w_i c_i from the tile in shared memory ------------------
    const int pw = warp - (3 + N_EPI);               // 0..3
    const int g = lane >> 2, t = lane & 3, mi = lane >> 3, rr = lane & 7;
    for (int64_t it = 0; it < n_local; ++it) {
      const int64_t tile = blockIdx.x + it * gridDim.x;
      const uint32_t p = (uint32_t)(it & 1), s = (uint32_t)(it % NSTG);
      tc::mbar_wait(w_ready + 8 * p, (uint32_t)((it >> 1) & 1));
      const uint4* af = afrag + p * 128;
      const uint4 a0 = af[lane], a1 = af[32 + lane], a2 = af[64 + lane], a3 = af[96 + lane];
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(af_free + 8 * p);
      const uint32_t rowa = sbase + OFF_TILES + s * TILE + (uint32_t)((8 * mi + rr) * 128);
      float* orow = nullptr;
      if (g < SPT && tile * SPT + g < n_seq) orow = out + (tile * SPT + g) * (int64_t)D + 2 * t;
#pragma unroll 2
      for (int j = pw; j < 38; j += N_POOL) {         // 16-byte unit = 8 context columns
        const uint32_t ua = rowa + (uint32_t)((j >> 3) * CHUNK) + (uint32_t)((((j & 7) ^ rr)) << 4);
        uint32_t b0, b1, b2, b3, b4, b5, b6, b7;
        ldsm_x4_t(ua, b0, b1, b2, b3);                // positions 0..31
        ldsm_x4_t(ua + 32 * 128, b4, b5, b6, b7);     // positions 32..63
        float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
        mma_k16(d0, a0.x, a0.y, a0.z, a0.w, b0, b1);
        mma_k16(d1, a1.x, a1.y, a1.z, a1.w, b2, b3);
        mma_k16(d0, a2.x, a2.y, a2.z, a2.w, b4, b5);
        mma_k16(d1, a3.x, a3.y, a3.z, a3.w, b6, b7);
        if (orow != nullptr && 8 * j + 2 * t < D)
          *reinterpret_cast<float2*>(orow + 8 * j) = make_float2((d0[0] + d1[0]) + (d0[2] + d1[2]), (d0[1] + d1[1]) + (d0[3] + d1[3]));
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(empty_bar + 8 * s);       // the tile's slot goes back to the producer
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace k2t

template <int S>
static int launch_k2t(const void* wa16, const void* Cbuf, int64_t n, const float* ba, const float* qa, float* out,
                      cudaStream_t st) {
  static bool configured[64] = {false};
  cudaError_t e = set_max_dynamic_smem(k2t::additive_pool_tmem_kernel<S>, k2t::SMEM, configured);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(additive_pool_tmem_kernel)");
  if (n <= 0) return NRMS_OK;
  constexpr int SPT = k2t::ROWS / S;
  alignas(64) CUtensorMap tc_;
  const int box_rows = (int)((n * S < k2t::ROWS) ? n * S : k2t::ROWS);
  if (int rc = make_tmap_k_major_f16(&tc_, Cbuf, n * S, k2t::CP, k2t::CP, box_rows)) return rc;
  const int64_t tiles = (n + SPT - 1) / SPT;
  int grid = num_sms();
  if (tiles < grid) grid = (int)tiles;
  k2t::additive_pool_tmem_kernel<S><<<grid, k2t::THREADS, k2t::SMEM, st>>>(
      tc_, reinterpret_cast<const __half*>(Cbuf), reinterpret_cast<const __half*>(wa16), ba, qa, out, n,
      (uint32_t)box_rows * 128u * (uint32_t)k2t::TMA_CHUNKS);
  NRMS_LAUNCH_CHECK("additive_pool_tmem_kernel");
  return NRMS_OK;
}

// wa16: the fp16 copy of W_a [200][320] (k2v2_prepare / pack_wa16_kernel); Cbuf: fp16 context rows [n * S][320]
int k2t_run(int S, const void* wa16, const void* Cbuf, int64_t n, const float* ba, const float* qa, float* out,
            cudaStream_t st) {
  if (S == 20) return launch_k2t<20>(wa16, Cbuf, n, ba, qa, out, st);
  if (S == 50) return launch_k2t<50>(wa16, Cbuf, n, ba, qa, out, st);
  set_error("additive_pool_tmem_kernel compiled for S = 20 or 50, got %d", S);
  return NRMS_E_UNSUPPORTED;
}

}  // namespace nrms
