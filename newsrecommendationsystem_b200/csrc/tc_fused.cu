// Fused tensor-core encoder forward (inference) for sm_100a.  Two persistent kernels per chunk of
// sequences; the only intermediate that leaves the SM is the attention context C (L2-resident chunk).
//
//  K1  encoder_attn_kernel<S,SPT>:  tile = SPT sequences (100 rows: 5 titles or 2 users)
//      workers (8 warps): gather the tile's input rows (embedding rows by token id, news-vector rows by
//        int32 index, or dense rows) with coalesced float4 loads, round to FP16 (11-bit significand, the
//        same as TF32, at half the operand bytes) and store them into the resident A tile in the UMMA
//        SWIZZLE_128B K-major layout (5 chunks of 64 halfs);
//      producer (1 thread): streams an fp16 copy of W_Q/W_K/W_V through a 4-stage TMA ring, three 80-row boxes per K
//        chunk (heads 4p..4p+3 of Q, K and V) -> B tile of 240 rows;
//      MMA (1 thread): 4 passes x 19 tcgen05.mma.kind::f16 (M=128, N=240, K=16, fp32 accumulate) into one of two
//        240-column TMEM accumulator stages;
//      workers again (two groups of 4 warps, one head each at a time): per head, tcgen05.ld q/k/v of
//        their own row (thread == token row), +bias, stage
//        K/V in shared memory, then the 15-head exp-softmax attention of the reference
//        (e = exp(qk/sqrt(20)), attn = e/(sum e + 1e-8), no max subtraction) entirely in registers with
//        warp-broadcast shared-memory reads; context rows go to C.
//  K2  additive_pool_kernel<S,SPT>: C (TMA, rounded to TF32 by the TFLOAT32 tensor map) x W_a^T on
//      tcgen05 (N=208), epilogue tanh -> . q -> stable softmax over the sequence -> weighted row sum.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <utility>
#include <vector>
#include "tc_common.cuh"
#include "tc_api.cuh"

namespace nrms {
using namespace tc;

int make_tmap_k_major(CUtensorMap* out, const float* base, int64_t rows, int cols, int64_t ld, int box_rows);
int make_tmap_k_major_f16(CUtensorMap* out, const void* base, int64_t rows, int cols, int64_t ld, int box_rows);
// K1 v2 (tc_fused2.cu): projections AND attention on tcgen05
size_t k1v2_w16_bytes();
int k1v2_prepare(const float* wqkv, void* w16, CUtensorMap* tw, cudaStream_t st);
int k1v2_run(int S, const CUtensorMap& tw, const float* src, const void* idx, int idx_kind, int64_t n,
             const float* bqkv, void* Cbuf, cudaStream_t st);
// K2 v2 (tc_fused3.cu): fp16 additive pooling with W_a resident in shared memory
int k2v2_prepare(const float* wa, void* wa16, CUtensorMap* twa, cudaStream_t st);
int k2v2_run(int S, const CUtensorMap& twa, const void* Cbuf, int64_t n, const float* ba, const float* qa, float* out,
             cudaStream_t st);
constexpr size_t WA16_SLOT_BYTES = 131072;  // fp16 copy of W_a [200][320]
// K1 v3 (tc_fused4.cu): v2 with two heads in flight and P in tensor memory (TS-form MMA)
int k1v3_prepare(const float* wqkv, void* w16, CUtensorMap* tw, cudaStream_t st);
int k1v3_run(int S, const CUtensorMap& tw, const float* src, const void* idx, int idx_kind, int64_t n,
             const float* bqkv, void* Cbuf, cudaStream_t st);
// NRMS_K1_VARIANT: 1 = first-generation K1 (CUDA-core attention, fp32 C, TF32 K2); 2 (default) = tensor-core
// attention; 3 = tensor-core attention with two heads in flight and P in tensor memory (TS-form MMA) -- measured
// within 5 % of variant 2 this round (profiles/), kept selectable for the next round's pipelining work
// K1 v4 (tc_fused5.cu): TMA-gathered fp16 source rows, bias/scale folded into the GEMM, q consumed from tensor memory
size_t k1v4_src16_bytes(int64_t n_rows);
int k1v4_prepare(const float* wqkv, const float* bqkv, void* w16, CUtensorMap* tw, cudaStream_t st);
int k1v4_pack_src(const float* src, int64_t n_rows, void* src16, CUtensorMap* ts, cudaStream_t st);
int k1v4_run(int S, const CUtensorMap& tw, const CUtensorMap& ts, const void* idx, int idx_kind, int64_t n,
             int null_row, void* Cbuf, cudaStream_t st);
// K1 v5 (tc_fused6.cu): v4 with two projection accumulators and P kept in place over the scores (default)
int k1v5_run(int S, const CUtensorMap& tw, const CUtensorMap& ts, const void* src16, const void* idx, int idx_kind,
             int64_t n, int null_row, void* Cbuf, cudaStream_t st);
// K1 v6 (tc_fused7.cu): v5 with two worker groups taking alternate passes
int k1v6_run(int S, const CUtensorMap& tw, const CUtensorMap& ts, const void* src16, const void* idx, int idx_kind,
             int64_t n, int null_row, void* Cbuf, cudaStream_t st);
// K1g (k1g_table_attn.cu): indexed user encoder over a pre-projected table (q|k|v rows gathered, no per-user GEMM)
int k1g_project_table(const float* table, int64_t n_rows, const float* wqkv, const float* bqkv, void* table16,
                      cudaStream_t st);
size_t k1g_table16_bytes(int64_t n_rows);
int k1g_run_seq(int S, int idx_kind, const void* table16, int64_t n_table_rows, const void* rows, int64_t n_seq,
                void* Cbuf, const float* qk_bound, cudaStream_t st);
int get_k1g_variant();
int k1g_run(const void* table16, int64_t n_table_rows, const int32_t* hist_rows, int64_t n_users, void* Cbuf,
            cudaStream_t st);
// "user_table_attn" option: 1 (default) = int32-indexed S=50 calls whose history rows outnumber the table rows 8:1
// project the table once and run K1g; 0 = always the per-user projection of K1 v1..v6
static int g_table_attn = -1, g_news_table_attn = -1;
static bool table_attn_enabled() {
  if (g_table_attn < 0) {
    const char* e = getenv("NRMS_USER_TABLE_ATTN");
    g_table_attn = (e && e[0] == '0') ? 0 : 1;
  }
  return g_table_attn != 0;
}
// "news_table_attn" option: the same trade for the news encoder -- project the EMBEDDING table once (70,976 rows against
// 1.3 M token rows at MIND-small shapes) and run the attention on gathered q|k|v rows
static bool news_table_attn_enabled() {
  if (g_news_table_attn < 0) {
    const char* e = getenv("NRMS_NEWS_TABLE_ATTN");
    g_news_table_attn = (e && e[0] == '0') ? 0 : 1;
  }
  return g_news_table_attn != 0;
}
static int g_fused_pool = -1;
static bool fused_pool_enabled() {
  if (g_fused_pool < 0) {
    const char* e = getenv("NRMS_FUSED_POOL");
    g_fused_pool = (e && e[0] == '1') ? 1 : 0;
  }
  return g_fused_pool != 0;
}
void set_fused_pool(bool on) { g_fused_pool = on ? 1 : 0; }
void set_table_attn(bool on) { g_table_attn = on ? 1 : 0; }
void set_news_table_attn(bool on) { g_news_table_attn = on ? 1 : 0; }
static bool use_table_attn(int S, int idx_kind, int64_t n_seq, int64_t n_src_rows) {
  // The projection costs 2.6 us per 1,000 table rows and every call (rank) pays it for the WHOLE table, and a table16
  // beyond L2 (126 MB = 58 k rows) turns the 2,160-byte row gather into DRAM traffic: measured 3.2 ms vs 4.4 ms (K1 v6)
  // at 56 history rows per table row, 4.4 vs ~4.8 ms at 28; the break-even is near 8.
  if (n_src_rows <= 0 || n_seq * S < 8 * n_src_rows) return false;
  // int32 rows = the user encoder over the news-vector table; int64 ids = the news encoder over the embedding table
  // (title length 20; a 50-token text, e.g. an abstract, takes the same kernel)
  if (idx_kind == 2) return S == 50 ? table_attn_enabled() : news_table_attn_enabled();
  if (idx_kind == 1) return news_table_attn_enabled();
  return false;
}
constexpr int K1_DEFAULT_VARIANT = 6;
static int g_k1_variant = -1;
static int k1_variant() {
  if (g_k1_variant < 0) {
    const char* e = getenv("NRMS_K1_VARIANT");
    g_k1_variant = (e && e[0] >= '1' && e[0] <= '6') ? (e[0] - '0') : K1_DEFAULT_VARIANT;
  }
  return g_k1_variant;
}
// ---- optional live timing of the K1 launches (bench.py's roofline): CUDA events around every launch, kept per kind:
//      0 = user encoder, per-user projection (K1 v1..v6), 1 = news encoder K1, 2 = user encoder, table attention (K1g),
//      3 = news encoder, table attention (K1g over the projected embedding table)
static bool g_time_k1 = false;
struct K1Record { cudaEvent_t a, b; int64_t seqs; int kind; };
static std::vector<K1Record> g_k1_records;
static std::vector<cudaEvent_t> g_k1_event_pool;     // events are reused: the first cudaEventCreate calls cost ~30 us each
void set_time_k1(bool on) {
  g_time_k1 = on;
  for (auto& r : g_k1_records) { g_k1_event_pool.push_back(r.a); g_k1_event_pool.push_back(r.b); }
  g_k1_records.clear();
}
static cudaEvent_t k1_timer_event() {
  cudaEvent_t e;
  if (!g_k1_event_pool.empty()) { e = g_k1_event_pool.back(); g_k1_event_pool.pop_back(); return e; }
  cudaEventCreate(&e);
  return e;
}
// key = 3 * kind + what; what: 0 = total ms of the timed launches, 1 = number of launches, 2 = sequences processed
double get_k1_stat(int key) {
  const int kind = key / 3, what = key % 3;
  double total = 0;
  for (auto& r : g_k1_records) {
    if (r.kind != kind) continue;
    if (what == 1) total += 1.0;
    else if (what == 2) total += (double)r.seqs;
    else {
      float ms = 0.f;
      if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) total += ms;
    }
  }
  return total;
}
struct K1Timer {
  cudaStream_t st; bool on; K1Record rec;
  K1Timer(cudaStream_t s, int64_t n, int kind) : st(s), on(g_time_k1 && g_k1_records.size() < 8192) {
    if (on) { rec.a = k1_timer_event(); rec.b = k1_timer_event(); cudaEventRecord(rec.a, st); rec.seqs = n; rec.kind = kind; }
  }
  ~K1Timer() { if (on) { cudaEventRecord(rec.b, st); g_k1_records.push_back(rec); } }
};

int set_k1_variant(int v) {
  if (v < 1 || v > 6) return NRMS_E_INVALID;
  g_k1_variant = v;
  return NRMS_OK;
}
constexpr size_t W16_SLOT_BYTES = 655360;   // >= every variant's fp16 weight copy

__device__ __forceinline__ void tma_load_2d_f(uint32_t dst_smem, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_f(uint32_t bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

constexpr int KCH = 10;            // K2: K chunks of 32 floats (300 -> 320, tail zero)
constexpr int KCH16 = 5;           // K1: K chunks of 64 halfs (128 B), 300 -> 320 with a zero tail
constexpr int K1_STAGES = 4;
constexpr int W16_LD = 320;        // fp16 copy of [W_Q;W_K;W_V]: [900][320] halfs, K zero-padded
constexpr int HP = 4;              // heads per QKV pass (the 16th "head" of the last pass is a dummy)
constexpr int NPASS = 4;
constexpr int K1_BOX = HP * DH;    // 80 weight rows per Q/K/V box (80 x 128 B = 10 KB, 1024-aligned)
constexpr int K1_N = 3 * K1_BOX;   // UMMA N of a pass = 240: [Q 80 | K 80 | V 80]
constexpr int K1_BSTAGE = K1_N * 128;     // 30,720 B
constexpr int K1_WORKERS = 256;    // 8 worker warps = 2 groups of 4 (one group per head, two heads in flight)
constexpr int K1_THREADS = 64 + K1_WORKERS;
__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }
__device__ __forceinline__ void workers_bar() { asm volatile("bar.sync 3, 256;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float LOG2E_OVER_SQRT_DH = 1.4426950408889634f / 4.47213595499957939f;

template <int S, int SPT>
struct K1 {
  static constexpr int ROWS = S * SPT;
  static constexpr int ROWS_ALLOC = (ROWS + 7) / 8 * 8;
  static constexpr int CH = ROWS_ALLOC * 128;                 // bytes of one A chunk (rows x 128 B)
  static constexpr int OFF_B = KCH16 * CH;
  static constexpr int OFF_KV = OFF_B + K1_STAGES * K1_BSTAGE;
  static constexpr int KV_BYTES = ROWS * 40 * 4;              // one group's K/V staging (one head)
  static constexpr int OFF_BIAS = OFF_KV + 2 * KV_BYTES;
  static constexpr int OFF_IDX = OFF_BIAS + 3712;
  static constexpr int OFF_BAR = OFF_IDX + 128 * 8;
  static constexpr int SMEM = OFF_BAR + 256 + 1024;
  static_assert(ROWS <= 128, "tile rows");
  static_assert(CH % 1024 == 0 && OFF_B % 1024 == 0, "swizzle alignment");
  static_assert(SMEM <= 232448, "shared memory budget");
};

// idx_kind: 0 = dense rows (row r of src), 1 = int64 ids, 2 = int32 ids
template <int S, int SPT>
__global__ void __launch_bounds__(K1_THREADS, 1)
encoder_attn_kernel(const __grid_constant__ CUtensorMap tmap_w, const float* __restrict__ src,
                    const void* __restrict__ idx, int idx_kind, int64_t n_seq, const float* __restrict__ bqkv,
                    float* __restrict__ C) {
  using Cfg = K1<S, SPT>;
  constexpr int ROWS = Cfg::ROWS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  float* kv = reinterpret_cast<float*>(sm + Cfg::OFF_KV);
  float* bias_s = reinterpret_cast<float*>(sm + Cfg::OFF_BIAS);
  int64_t* rowid = reinterpret_cast<int64_t*>(sm + Cfg::OFF_IDX);
  const uint32_t bars = base + Cfg::OFF_BAR;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * K1_STAGES, a_full = bars + 16 * K1_STAGES,
                 a_free = a_full + 8, acc_full = a_full + 16, acc_empty = a_full + 32;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(sm + Cfg::OFF_BAR + 16 * K1_STAGES + 64);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (n_seq + SPT - 1) / SPT;

  if (tid == 0) {
    for (int s = 0; s < K1_STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full + 8 * s, 1);
      mbar_init(acc_empty + 8 * s, 8);
    }
    mbar_init(a_full, K1_WORKERS);
    mbar_init(a_free, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_ptr_smem), 512);
  for (int i = tid; i < D3; i += K1_THREADS) bias_s[i] = bqkv[i];
  // K tail of the A tile (halfs 300..319 of every row: K chunk 4, bytes 88..127) is zero for good
  for (int r = tid; r < Cfg::ROWS_ALLOC; r += K1_THREADS) {
    uint8_t* rowp = sm + 4 * Cfg::CH + r * 128;
    *reinterpret_cast<uint2*>(rowp + ((5 ^ (r & 7)) << 4) + 8) = make_uint2(0u, 0u);
    *reinterpret_cast<uint4*>(rowp + ((6 ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(rowp + ((7 ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------ TMA producer: W_Q / W_K / W_V boxes ----------------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int p = 0; p < NPASS; ++p) {
          for (int kc = 0; kc < KCH16; ++kc, ++it) {
            const int s = it % K1_STAGES;
            mbar_wait(empty_bar + 8 * s, ((it / K1_STAGES) & 1) ^ 1);
            mbar_expect_tx_f(full_bar + 8 * s, K1_BSTAGE);
            const uint32_t sb = base + Cfg::OFF_B + s * K1_BSTAGE;
            tma_load_2d_f(sb, &tmap_w, kc * 64, K1_BOX * p, full_bar + 8 * s);
            tma_load_2d_f(sb + K1_BOX * 128, &tmap_w, kc * 64, D + K1_BOX * p, full_bar + 8 * s);
            tma_load_2d_f(sb + 2 * K1_BOX * 128, &tmap_w, kc * 64, 2 * D + K1_BOX * p, full_bar + 8 * s);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------------------------
    const uint32_t idesc = umma_idesc_f16(128, K1_N);
    uint32_t it = 0, pass_it = 0, tile_it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_it) {
      mbar_wait(a_full, tile_it & 1);           // the workers finished writing this tile's A rows
      tc_fence_after();
      for (int p = 0; p < NPASS; ++p, ++pass_it) {
        const uint32_t as = pass_it & 1;
        mbar_wait(acc_empty + 8 * as, ((pass_it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * K1_N;
        for (int kc = 0; kc < KCH16; ++kc, ++it) {
          const int s = it % K1_STAGES;
          mbar_wait(full_bar + 8 * s, (it / K1_STAGES) & 1);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = base + kc * Cfg::CH;
            const uint32_t sb = base + Cfg::OFF_B + s * K1_BSTAGE;
            const int ksteps = (kc == KCH16 - 1) ? 3 : 4;   // K = 300 -> 18.75 steps of 16, padded to 19
            for (int ks = 0; ks < ksteps; ++ks)
              umma_f16_ss(d_tmem, umma_desc_k_sw128(sa + ks * 32), umma_desc_k_sw128(sb + ks * 32), idesc,
                          (kc | ks) ? 1u : 0u);
            umma_commit(empty_bar + 8 * s);
            if (kc == KCH16 - 1) {
              umma_commit(acc_full + 8 * as);
              if (p == NPASS - 1) umma_commit(a_free);      // A tile may be overwritten after these MMAs
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------ workers (warps 2..9) ----------------------------------------
    const int g = (warp - 2) >> 2;               // group: handles heads hh = g, g+2 of every pass
    const int q4 = warp & 3;                     // TMEM lane quarter this warp may access
    const int wt = (warp - 2) * 32 + lane;       // 0..255 worker-thread index (gather work split)
    const int row = q4 * 32 + lane;              // tile row == TMEM lane owned by this thread
    const bool row_ok = row < ROWS;
    const int sq = row_ok ? row / S : 0;
    float* kvg = kv + g * (Cfg::KV_BYTES / 4);
    uint32_t pass_it = 0, tile_it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_it) {
      const int64_t seq0 = t * SPT;
      // ---- row ids of this tile ----
      if (wt < ROWS) {
        const int64_t seq = seq0 + wt / S;
        int64_t id = 0;
        if (seq < n_seq) {
          const int64_t e = seq * S + (wt % S);
          id = idx_kind == 0 ? e : (idx_kind == 1 ? reinterpret_cast<const int64_t*>(idx)[e]
                                                  : (int64_t) reinterpret_cast<const int32_t*>(idx)[e]);
        }
        rowid[wt] = id;
      }
      mbar_wait(a_free, (tile_it & 1) ^ 1);      // previous tile's MMAs no longer read the A tile
      workers_bar();
      // ---- gather + TF32 rounding into the swizzled A tile: ROWS x 75 float4 ----
      constexpr int TOTAL4 = ROWS * DV4;
#pragma unroll 1
      for (int f0 = 0; f0 < TOTAL4; f0 += K1_WORKERS * 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int f = f0 + u * K1_WORKERS + wt;
          if (f < TOTAL4) {
            const int r = f / DV4, c4 = f - r * DV4;
            v[u] = __ldg(reinterpret_cast<const float4*>(src + rowid[r] * D) + c4);
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int f = f0 + u * K1_WORKERS + wt;
          if (f < TOTAL4) {
            const int r = f / DV4, c4 = f - r * DV4;
            // 4 floats -> 4 halfs (round-to-nearest: same 11-bit significand as TF32) = 8 bytes
            const __half2 lo = __floats2half2_rn(v[u].x, v[u].y), hi = __floats2half2_rn(v[u].z, v[u].w);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&lo);
            pk.y = *reinterpret_cast<const uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(sm + (c4 >> 4) * Cfg::CH + r * 128 + ((((c4 & 15) >> 1) ^ (r & 7)) << 4) +
                                      ((c4 & 1) << 3)) = pk;
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(a_full);
      // ---- per pass: this group's heads 4p+g and 4p+g+2 ----
      for (int p = 0; p < NPASS; ++p, ++pass_it) {
        const uint32_t as = pass_it & 1;
        mbar_wait(acc_full + 8 * as, (pass_it >> 1) & 1);
        tc_fence_after();
        const uint32_t trow = tmem_base + as * K1_N + ((uint32_t)(q4 * 32) << 16);
#pragma unroll 1
        for (int u = 0; u < 2; ++u) {
          const int hh = g + 2 * u;
          const int h = p * HP + hh;
          if (h >= H) break;                         // dummy 16th head (last pass, group 1)
          const bool last_read = (u == 1) || (h + 2 >= H);
          float qv[DH], kk[DH], vv[DH];
          tmem_ld16(trow + hh * DH, qv);                  tmem_ld4(trow + hh * DH + 16, qv + 16);
          tmem_ld16(trow + K1_BOX + hh * DH, kk);         tmem_ld4(trow + K1_BOX + hh * DH + 16, kk + 16);
          tmem_ld16(trow + 2 * K1_BOX + hh * DH, vv);     tmem_ld4(trow + 2 * K1_BOX + hh * DH + 16, vv + 16);
          if (last_read) {                           // this warp is done with the accumulator stage
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + 8 * as);
          }
#pragma unroll
          for (int d = 0; d < DH; ++d) {
            qv[d] = (qv[d] + bias_s[h * DH + d]) * LOG2E_OVER_SQRT_DH;   // fold 1/sqrt(d) and log2(e) into q
            kk[d] += bias_s[D + h * DH + d];
            vv[d] += bias_s[2 * D + h * DH + d];
          }
          if (row_ok) {
            float4* kp = reinterpret_cast<float4*>(kvg + row * 40);
#pragma unroll
            for (int c = 0; c < 5; ++c) {
              kp[c] = make_float4(kk[4 * c], kk[4 * c + 1], kk[4 * c + 2], kk[4 * c + 3]);
              kp[5 + c] = make_float4(vv[4 * c], vv[4 * c + 1], vv[4 * c + 2], vv[4 * c + 3]);
            }
          }
          group_bar(g);
          if (row_ok) {
            float acc[DH];
#pragma unroll
            for (int d = 0; d < DH; ++d) acc[d] = 0.f;
            float Z = 0.f;
            const float4* base_kv = reinterpret_cast<const float4*>(kvg + sq * S * 40);
#pragma unroll 2
            for (int j = 0; j < S; ++j) {
              const float4* kp = base_kv + j * 10;
              float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;     // 4 independent chains
#pragma unroll
              for (int c = 0; c < 5; ++c) {
                const float4 k4 = kp[c];
                s0 = fmaf(qv[4 * c], k4.x, s0); s1 = fmaf(qv[4 * c + 1], k4.y, s1);
                s2 = fmaf(qv[4 * c + 2], k4.z, s2); s3 = fmaf(qv[4 * c + 3], k4.w, s3);
              }
              const float e = ex2_approx((s0 + s1) + (s2 + s3));   // exp(q.k / sqrt(20)), no max subtraction
              Z += e;
#pragma unroll
              for (int c = 0; c < 5; ++c) {
                const float4 v4 = kp[5 + c];
                acc[4 * c] = fmaf(e, v4.x, acc[4 * c]); acc[4 * c + 1] = fmaf(e, v4.y, acc[4 * c + 1]);
                acc[4 * c + 2] = fmaf(e, v4.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(e, v4.w, acc[4 * c + 3]);
              }
            }
            const float inv = 1.f / (Z + 1e-8f);
            if (seq0 + sq < n_seq) {
              float4* op = reinterpret_cast<float4*>(C + (seq0 * S + row) * D + h * DH);
#pragma unroll
              for (int c = 0; c < 5; ++c)
                op[c] = make_float4(acc[4 * c] * inv, acc[4 * c + 1] * inv, acc[4 * c + 2] * inv, acc[4 * c + 3] * inv);
            }
          }
          group_bar(g);       // this group's K/V staging is reused by its next head
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------
// K2: additive attention pooling.  Streaming GEMM (A = C rows, B = W_a) + fused epilogue.
// ---------------------------------------------------------------------------------------------------
constexpr int K2_STAGES = 4;
constexpr int K2_A_SLOT = 16384;           // up to 128 rows x 128 B
constexpr int K2_B_SLOT = 26624;           // 200 rows x 128 B = 25,600 -> padded to a 1024 multiple
constexpr int K2_STAGE = K2_A_SLOT + K2_B_SLOT;
constexpr int K2_N = 208;
constexpr int K2_THREADS = 192;
constexpr int K2_OFF_MISC = K2_STAGES * K2_STAGE;       // sc[128] | wv[128] | ba[200] | qa[200] | barriers
constexpr int K2_SMEM = K2_OFF_MISC + 4096 + 1024;

__device__ __forceinline__ float fast_tanh(float x) {
  // tanh(x) = 1 - 2/(exp(2x)+1); abs error ~1e-7 (ex2.approx + div.approx), saturates cleanly at +-1
  const float e = exp2f(x * 2.885390081777927f);
  return 1.f - __fdividef(2.f, e + 1.f);
}

template <int S, int SPT>
__global__ void __launch_bounds__(K2_THREADS, 1)
additive_pool_kernel(const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_wa,
                     const float* __restrict__ C, const float* __restrict__ ba, const float* __restrict__ qa,
                     float* __restrict__ out, int64_t n_seq, uint32_t stage_tx_bytes) {
  constexpr int ROWS = S * SPT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  float* sc = reinterpret_cast<float*>(sm + K2_OFF_MISC);
  float* wv = sc + 128;
  float* ba_s = wv + 128;
  float* qa_s = ba_s + 208;
  const uint32_t bars = base + K2_OFF_MISC + 3072;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * K2_STAGES, tfull_bar = bars + 16 * K2_STAGES,
                 tempty_bar = tfull_bar + 16;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(sm + K2_OFF_MISC + 3072 + 16 * K2_STAGES + 32);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (n_seq + SPT - 1) / SPT;

  if (tid == 0) {
    for (int s = 0; s < K2_STAGES; ++s) {
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
  for (int i = tid; i < 208; i += K2_THREADS) {
    ba_s[i] = i < QD ? ba[i] : 0.f;
    qa_s[i] = i < QD ? qa[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int r0 = (int)(t * ROWS);
        for (int kc = 0; kc < KCH; ++kc, ++it) {
          const int s = it % K2_STAGES;
          mbar_wait(empty_bar + 8 * s, ((it / K2_STAGES) & 1) ^ 1);
          mbar_expect_tx_f(full_bar + 8 * s, stage_tx_bytes);
          const uint32_t sa = base + s * K2_STAGE;
          tma_load_2d_f(sa, &tmap_c, kc * 32, r0, full_bar + 8 * s);
          tma_load_2d_f(sa + K2_A_SLOT, &tmap_wa, kc * 32, 0, full_bar + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_tf32(128, K2_N);
    uint32_t it = 0, tile_it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_it) {
      const uint32_t as = tile_it & 1;
      mbar_wait(tempty_bar + 8 * as, ((tile_it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * 256;
      for (int kc = 0; kc < KCH; ++kc, ++it) {
        const int s = it % K2_STAGES;
        mbar_wait(full_bar + 8 * s, (it / K2_STAGES) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = base + s * K2_STAGE;
          const uint32_t sb = sa + K2_A_SLOT;
          const int ksteps = (kc == KCH - 1) ? 2 : 4;
          for (int ks = 0; ks < ksteps; ++ks)
            umma_tf32_ss(d_tmem, umma_desc_k_sw128(sa + ks * 32), umma_desc_k_sw128(sb + ks * 32), idesc,
                         (kc | ks) ? 1u : 0u);
          umma_commit(empty_bar + 8 * s);
          if (kc == KCH - 1) umma_commit(tfull_bar + 8 * as);
        }
        __syncwarp();
      }
    }
  } else {
    const int q4 = warp & 3;
    const int wt = (warp - 2) * 32 + lane;
    const int row = q4 * 32 + lane;
    const bool row_ok = row < ROWS;
    uint32_t tile_it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_it) {
      const int64_t seq0 = t * SPT;
      const uint32_t as = tile_it & 1;
      mbar_wait(tfull_bar + 8 * as, (tile_it >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + as * 256 + ((uint32_t)(q4 * 32) << 16);
      float s = 0.f;
#pragma unroll 1
      for (int col = 0; col < 192; col += 16) {
        float v[16];
        tmem_ld16(trow + col, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) s = fmaf(fast_tanh(v[j] + ba_s[col + j]), qa_s[col + j], s);
      }
      {   // columns 192..199; accumulator columns 200..207 come from unwritten smem rows and are never read
        float v[8];
        tmem_ld4(trow + 192, v);
        tmem_ld4(trow + 196, v + 4);
#pragma unroll
        for (int j = 0; j < 8; ++j) s = fmaf(fast_tanh(v[j] + ba_s[192 + j]), qa_s[192 + j], s);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * as);
      sc[row] = s;
      worker_bar();
      if (row_ok) {
        const int sq = row / S;
        float m = -INFINITY;
        for (int j = 0; j < S; ++j) m = fmaxf(m, sc[sq * S + j]);
        float sum = 0.f;
        for (int j = 0; j < S; ++j) sum += __expf(sc[sq * S + j] - m);
        wv[row] = __fdividef(__expf(s - m), sum);
      }
      worker_bar();
      // pooled[seq, d] = sum_i w_i C[seq*S + i, d]   (C rows are L2-hot: K1 just wrote them);
      // 10 independent 16-byte loads in flight per thread
      for (int o = wt; o < SPT * DV4; o += 128) {
        const int sq = o / DV4, l = o - sq * DV4;
        if (seq0 + sq < n_seq) {
          const float4* cp = reinterpret_cast<const float4*>(C + (seq0 + sq) * S * D) + l;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
          for (int i0 = 0; i0 < S; i0 += 10) {
            float4 c4[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) c4[i] = __ldg(cp + (i0 + i) * DV4);
#pragma unroll
            for (int i = 0; i < 10; ++i) {
              const float w = wv[sq * S + i0 + i];
              acc.x = fmaf(w, c4[i].x, acc.x); acc.y = fmaf(w, c4[i].y, acc.y);
              acc.z = fmaf(w, c4[i].z, acc.z); acc.w = fmaf(w, c4[i].w, acc.w);
            }
          }
          reinterpret_cast<float4*>(out + (seq0 + sq) * D)[l] = acc;
        }
      }
      worker_bar();     // sc / wv are reused by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// fp16 copy of the packed projection weights for K1's B operand: [900][320] halfs, K tail zero
__global__ void __launch_bounds__(256) wqkv_to_f16_kernel(const float* __restrict__ w, __half* __restrict__ out) {
  const int n = D3 * W16_LD;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / W16_LD, k = i - r * W16_LD;
    out[i] = __float2half_rn(k < D ? w[r * D + k] : 0.f);
  }
}
constexpr size_t W16_BYTES = (size_t)D3 * W16_LD * 2;   // 576,000 (a multiple of 256)

// ---------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------
template <int S, int SPT>
static int64_t fused_chunk_seq(bool table_attn) {   // full waves of tiles per launch; NRMS_FUSED_WAVES to experiment
  // Measured on the evaluate bench.  K1 v6 + K2: 4 waves 5.9 ms of encoder time, 8 waves 5.3, 16 waves 4.97, 32 waves
  // 4.92 (news 1.61 vs 1.66 ms at 16; users equal): every launch pays the K2 prologue (W_a into shared memory), the
  // pipeline fill and a tail; users stay at 16 so the fp16 context chunk (152 MB) is still mostly L2-resident between
  // K1 and K2.  Table path (K1g + K2): monotonic -- users 3.28 ms at 8 waves, 2.92 at 16, 2.73 at 32, 2.63 at 64, 2.60 in
  // one launch; news 1.50 / 1.35 / 1.30 / 1.25 / 1.24 -- so 64 waves (1.1 GB of context rows).
  static int env_waves = -1;
  if (env_waves < 0) { const char* e = getenv("NRMS_FUSED_WAVES"); env_waves = e ? atoi(e) : 0; if (env_waves < 0) env_waves = 0; }
  const int waves = env_waves ? env_waves : (table_attn ? 64 : (S == 20 ? 32 : 16));
  return (int64_t)num_sms() * SPT * waves;
}

// workspace: [fp16 W_qkv copy][fp16 W_a copy][fp16 gather source (variant 4)][context rows C of one chunk]
static size_t fused_src16_bytes(int S, int idx_kind_dense, int64_t n_src_rows, int64_t chunk_rows) {
  size_t b = k1v4_src16_bytes(idx_kind_dense ? chunk_rows : n_src_rows);
  if (!idx_kind_dense) {                       // room for the projected q|k|v table of K1g (1,800 B per row)
    const size_t t = k1g_table16_bytes(n_src_rows);
    if (t > b) b = t;
  }
  return align_up(b, 1024);
}

// In-place LayerNorm(300) over the fp16 context rows [rows][320] that K1 hands to K2 (columns 300..319 stay zero):
// one warp per row, five half2 per lane, statistics in fp32.
__global__ void __launch_bounds__(256) layernorm_f16_rows_kernel(__half* __restrict__ c, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, int64_t n_rows, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n_rows; r += n_warps) {
    __half2* row = reinterpret_cast<__half2*>(c + r * 320);
    float2 v[5];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int i = lane + 32 * j;
      v[j] = (i < D / 2) ? __half22float2(row[i]) : make_float2(0.f, 0.f);
      s += v[j].x + v[j].y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j)
      if (lane + 32 * j < D / 2) q += (v[j].x - mean) * (v[j].x - mean) + (v[j].y - mean) * (v[j].y - mean);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / D) + eps);
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int i = lane + 32 * j;
      if (i < D / 2) {
        const float2 g = __ldg(reinterpret_cast<const float2*>(gamma) + i);
        const float2 b = __ldg(reinterpret_cast<const float2*>(beta) + i);
        row[i] = __floats2half2_rn((v[j].x - mean) * rstd * g.x + b.x, (v[j].y - mean) * rstd * g.y + b.y);
      }
    }
  }
}

template <int S, int SPT>
static int run_fused(const float* src, int64_t n_src_rows, const void* idx, int idx_kind, int64_t n_seq,
                     const float* wqkv, const float* bqkv, const float* wa, const float* ba, const float* qa, float* out,
                     void* workspace, size_t workspace_bytes, cudaStream_t st, const float* ln_gamma,
                     const float* ln_beta) {
  using Cfg = K1<S, SPT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(encoder_attn_kernel<S, SPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(encoder_attn_kernel)");
    e = cudaFuncSetAttribute(additive_pool_kernel<S, SPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, K2_SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(additive_pool_kernel)");
    configured = true;
  }
  int variant = k1_variant();
  if (ln_gamma && variant < 2) variant = 5;   // the LayerNorm step works on the fp16 context rows of variants 2..5
  // The LayerNorm variant keeps the per-sequence projection (K1 v6 -> LayerNorm rows -> K2): the normalisation needs the
  // whole 300-wide context row, which in the table path is spread over 15 head warps.
  const bool table_attn = !ln_gamma && use_table_attn(S, idx_kind, n_seq, n_src_rows);
  if (table_attn && variant < 2) variant = K1_DEFAULT_VARIANT;
  // every variant tiles 5 titles / 2 users, so the chunking (whole waves of tiles) is shared
  const int64_t chunk = fused_chunk_seq<S, SPT>(table_attn);
  const int64_t first = n_seq < chunk ? n_seq : chunk;
  const size_t src16_bytes = fused_src16_bytes(S, idx_kind == 0, n_src_rows, first * S);
  const size_t need = W16_SLOT_BYTES + WA16_SLOT_BYTES + src16_bytes + (size_t)first * S * D * sizeof(float);
  NRMS_CHECK_ARG(workspace && aligned16(workspace) && workspace_bytes >= need, NRMS_E_WORKSPACE,
                 "workspace too small: need %zu bytes", need);
  __half* w16 = reinterpret_cast<__half*>(workspace);
  void* wa16 = reinterpret_cast<char*>(workspace) + W16_SLOT_BYTES;
  void* src16 = reinterpret_cast<char*>(workspace) + W16_SLOT_BYTES + WA16_SLOT_BYTES;
  float* Cbuf = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + W16_SLOT_BYTES + WA16_SLOT_BYTES + src16_bytes);
  alignas(64) CUtensorMap tw, twa, tc_, ts;
  if (variant >= 2) {
    int rc = variant >= 4 ? k1v4_prepare(wqkv, bqkv, w16, &tw, st)
                          : (variant == 3 ? k1v3_prepare(wqkv, w16, &tw, st) : k1v2_prepare(wqkv, w16, &tw, st));
    if (rc) return rc;
    if (int rc2 = k2v2_prepare(wa, wa16, &twa, st)) return rc2;
    if (table_attn) {
      if (int rc3 = k1g_project_table(src, n_src_rows, wqkv, bqkv, src16, st)) return rc3;
      // bound on the attention scores over the projected table (30 floats in the 3 KB behind the 128,000-byte fp16 W_a
      // copy): the attention kernels pick the plain or the row-shifted softmax form from it
      float* bound = reinterpret_cast<float*>(reinterpret_cast<char*>(wa16) + 128000);
      if (int rc4 = k1f_qk_bound(src16, n_src_rows, bound, st)) return rc4;
      if (fused_pool_enabled()) {
        // K1f: attention + additive pooling in ONE launch over the whole call; no context buffer, no chunking.
        K1Timer timer(st, n_seq, S == 50 && idx_kind == 2 ? 2 : 3);
        return k1f_run(S, idx_kind, src16, n_src_rows, idx, n_seq, wa16, ba, qa, bound, out, st);
      }
      // K1g variants 0 / 1 write only columns 0..299 of the fp16 context rows; K2 multiplies 300..319 by zero weights, so
      // they must be finite.  (The templated kernel clears them itself: this strided memset -- 40 bytes in every 640,
      // 470 k rows -- took ~0.4 ms on the copy engine.)
      if (S == 50 && idx_kind == 2 && get_k1g_variant() != 2) {
        cudaError_t e = cudaMemset2DAsync(reinterpret_cast<char*>(Cbuf) + 600, 640, 0, 40, (size_t)first * S, st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemset2DAsync(context tail)");
      }
    } else if (variant >= 4 && idx_kind != 0) {
      NRMS_CHECK_ARG(n_src_rows > 0, NRMS_E_INVALID, "indexed input needs the row count of its source table");
      if (int rc3 = k1v4_pack_src(src, n_src_rows, src16, &ts, st)) return rc3;
    }
  } else {
    wqkv_to_f16_kernel<<<148, 256, 0, st>>>(wqkv, w16);
    NRMS_LAUNCH_CHECK("wqkv_to_f16_kernel");
    if (int rc = make_tmap_k_major_f16(&tw, w16, D3, W16_LD, W16_LD, K1_BOX)) return rc;
  }
  if (variant < 2) {
    if (int rc = make_tmap_k_major(&twa, wa, QD, D, D, QD)) return rc;
  }
  const size_t idx_elem = idx_kind == 1 ? 8 : 4;
  for (int64_t s0 = 0; s0 < n_seq; s0 += chunk) {
    const int64_t n = (n_seq - s0 < chunk) ? (n_seq - s0) : chunk;
    const int64_t tiles = (n + SPT - 1) / SPT;
    int grid = num_sms();
    if (tiles < grid) grid = (int)tiles;
    const float* src_c = idx_kind == 0 ? src + s0 * S * D : src;
    const void* idx_c = idx_kind == 0 ? nullptr : (const void*)((const char*)idx + (size_t)s0 * S * idx_elem);
    if (variant >= 2) {
      if (variant >= 4 && idx_kind == 0) {   // dense rows: this chunk's rows become the fp16 gather source
        if (int rc = k1v4_pack_src(src_c, n * S, src16, &ts, st)) return rc;
      }
      {
        K1Timer timer(st, n, idx_kind != 1 && S == 50 ? (table_attn ? 2 : 0) : (table_attn ? 3 : 1));
        int rc;
        if (table_attn && (S != 50 || idx_kind != 2 || get_k1g_variant() == 2))
          rc = k1g_run_seq(S, idx_kind, src16, n_src_rows, idx_c, n, Cbuf,
                           reinterpret_cast<const float*>(reinterpret_cast<const char*>(wa16) + 128000), st);
        else if (table_attn)
          rc = k1g_run(src16, n_src_rows, reinterpret_cast<const int32_t*>(idx_c), n, Cbuf, st);
        else if (variant == 6)
          rc = k1v6_run(S, tw, ts, src16, idx_c, idx_kind, n, (int)(idx_kind == 0 ? n * S : n_src_rows), Cbuf, st);
        else if (variant == 5)
          rc = k1v5_run(S, tw, ts, src16, idx_c, idx_kind, n, (int)(idx_kind == 0 ? n * S : n_src_rows), Cbuf, st);
        else if (variant == 4)
          rc = k1v4_run(S, tw, ts, idx_c, idx_kind, n, (int)(idx_kind == 0 ? n * S : n_src_rows), Cbuf, st);
        else
          rc = variant == 3 ? k1v3_run(S, tw, src_c, idx_c, idx_kind, n, bqkv, Cbuf, st)
                            : k1v2_run(S, tw, src_c, idx_c, idx_kind, n, bqkv, Cbuf, st);
        if (rc) return rc;
      }
      if (ln_gamma) {
        int64_t lb = (n * S + 7) / 8;
        if (lb > (int64_t)num_sms() * 8) lb = (int64_t)num_sms() * 8;
        layernorm_f16_rows_kernel<<<(unsigned)lb, 256, 0, st>>>(reinterpret_cast<__half*>(Cbuf), ln_gamma, ln_beta,
                                                                n * S, 1e-5f);
        NRMS_LAUNCH_CHECK("layernorm_f16_rows_kernel");
      }
      if (int rc = k2v2_run(S, twa, Cbuf, n, ba, qa, out + s0 * D, st)) return rc;
      continue;
    } else {
      encoder_attn_kernel<S, SPT><<<grid, K1_THREADS, Cfg::SMEM, st>>>(tw, src_c, idx_c, idx_kind, n, bqkv, Cbuf);
      NRMS_LAUNCH_CHECK("encoder_attn_kernel");
    }
    const int box_c = (int)((n * S < Cfg::ROWS) ? n * S : Cfg::ROWS);
    if (int rc = make_tmap_k_major(&tc_, Cbuf, n * S, D, D, box_c)) return rc;
    additive_pool_kernel<S, SPT><<<grid, K2_THREADS, K2_SMEM, st>>>(tc_, twa, Cbuf, ba, qa, out + s0 * D, n,
                                                                    (uint32_t)(box_c + QD) * 128u);
    NRMS_LAUNCH_CHECK("additive_pool_kernel");
  }
  return NRMS_OK;
}

// n_src_rows: rows of the gather source (vocabulary / news-vector table); 0 for dense input
size_t tc_fused_workspace_bytes(int64_t n_seq, int S, int64_t n_src_rows) {
  if (n_seq <= 0) return (size_t)-1;
  // sized for the larger launches of the table path whenever the call is big enough to take it (whatever the options say)
  const bool may_table = n_src_rows > 0 && n_seq * S >= 8 * n_src_rows;
  int64_t chunk;
  if (S == 20) chunk = fused_chunk_seq<20, 5>(may_table);
  else if (S == 50) chunk = fused_chunk_seq<50, 2>(may_table);
  else return (size_t)-1;
  const int64_t first = n_seq < chunk ? n_seq : chunk;
  return W16_SLOT_BYTES + WA16_SLOT_BYTES + fused_src16_bytes(S, n_src_rows <= 0, n_src_rows, first * S) +
         (size_t)first * S * D * sizeof(float);
}

int tc_encoder_fused(const float* src, int64_t n_src_rows, const void* idx, int idx_kind, int64_t n_seq, int S,
                     const float* wqkv, const float* bqkv, const float* wa, const float* ba, const float* qa, float* out,
                     void* workspace, size_t workspace_bytes, cudaStream_t st, const float* ln_gamma,
                     const float* ln_beta) {
  if (S == 20) return run_fused<20, 5>(src, n_src_rows, idx, idx_kind, n_seq, wqkv, bqkv, wa, ba, qa, out, workspace, workspace_bytes, st, ln_gamma, ln_beta);
  if (S == 50) return run_fused<50, 2>(src, n_src_rows, idx, idx_kind, n_seq, wqkv, bqkv, wa, ba, qa, out, workspace, workspace_bytes, st, ln_gamma, ln_beta);
  set_error("fused encoder compiled for S = 20 or 50, got %d", S);
  return NRMS_E_UNSUPPORTED;
}

}  // namespace nrms
