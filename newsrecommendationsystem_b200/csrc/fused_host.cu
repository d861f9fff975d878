// Host side of the tensor-mode inference encoders (gather -> Q/K/V -> 15-head exp-softmax attention -> additive pooling;
// reference: src/model/NRMS/news_encoder.py:27-48, src/model/NRMS/user_encoder.py:15-26 in eval mode).  Three kernel paths:
//
//   table path   (indexed input whose gathered rows outnumber the source rows 4:1): the SOURCE TABLE is projected once
//                (k1g_project_table: one kind::f16 GEMM, q|k|v rows in fp16), then
//                  K1g  k1g::seq_attn_kernel      attention over gathered q|k|v rows -> fp16 context rows -> K2   [default]
//                  K1f  k1f::attn_pool_kernel     the same attention + the additive pooling in ONE kernel ("fused_pool" 1)
//   per-sequence (dense input, small calls, the LayerNorm variant): K1 v6 k1v6::encoder_attn_tc6_kernel projects every
//                gathered row on tcgen05 -> fp16 context rows -> [LayerNorm rows] -> K2
//   K2           k2v2::additive_pool_f16_kernel   additive attention pooling of the context rows
//
// idx_kind 0: dense rows (sequence s, position i -> row s*S+i), 1: int64 ids, 2: int32 ids.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <utility>
#include <vector>
#include "tc_common.cuh"
#include "tc_api.cuh"

namespace nrms {

int make_tmap_k_major_f16(CUtensorMap* out, const void* base, int64_t rows, int cols, int64_t ld, int box_rows);
// pack.cu
size_t src16_bytes(int64_t n_rows);
int pack_weights_k1(const float* wqkv, const float* bqkv, void* w16, CUtensorMap* tw, cudaStream_t st);
int pack_rows16(const float* src, int64_t n_rows, void* src16, cudaStream_t st);
// K2 (tc_fused3.cu): fp16 additive pooling with W_a resident in shared memory
int k2v2_prepare(const float* wa, void* wa16, CUtensorMap* twa, cudaStream_t st);
int k2v2_run(int S, const CUtensorMap& twa, const void* Cbuf, int64_t n, const float* ba, const float* qa, float* out,
             cudaStream_t st);
// K1 v6 (tc_fused7.cu)
int k1v6_run(int S, const CUtensorMap& tw, const void* src16, const void* idx, int idx_kind, int64_t n, int null_row,
             void* Cbuf, cudaStream_t st);
// K1g (k1g_table_attn.cu)
int k1g_project_table(const float* table, const void* table_rows16, int64_t n_rows, const float* wqkv, const float* bqkv,
                      void* scratch, cudaStream_t st, int64_t hot_row);
size_t k1g_table16_bytes(int64_t n_rows, bool rows16_given);
const void* k1g_table16_ptr(void* scratch);
int k1g_run_seq(int S, int idx_kind, const void* table16, int64_t n_table_rows, const void* rows, int64_t n_seq,
                void* Cbuf, const float* qk_bound, cudaStream_t st, int64_t hot_row);

constexpr size_t W16_SLOT_BYTES = 655360;   // fp16 weight copy of K1 v6: [1024][320] halfs
constexpr size_t WA16_SLOT_BYTES = 131072;  // fp16 copy of W_a [200][320] (128,000 B) + the 30 score bounds behind it

// ---- options ("nrms_set_option") ------------------------------------------------------------------------------------
static int g_table_attn = -1, g_news_table_attn = -1, g_fused_pool = -1, g_table_ratio = -1;
static bool env_flag(const char* name, bool dflt) {
  const char* e = getenv(name);
  if (!e || !e[0]) return dflt;
  return e[0] != '0';
}
static bool table_attn_enabled() {
  if (g_table_attn < 0) g_table_attn = env_flag("NRMS_USER_TABLE_ATTN", true) ? 1 : 0;
  return g_table_attn != 0;
}
static bool news_table_attn_enabled() {
  if (g_news_table_attn < 0) g_news_table_attn = env_flag("NRMS_NEWS_TABLE_ATTN", true) ? 1 : 0;
  return g_news_table_attn != 0;
}
static bool fused_pool_enabled() {
  if (g_fused_pool < 0) g_fused_pool = env_flag("NRMS_FUSED_POOL", false) ? 1 : 0;
  return g_fused_pool != 0;
}
void set_fused_pool(bool on) { g_fused_pool = on ? 1 : 0; }
void set_table_attn(bool on) { g_table_attn = on ? 1 : 0; }
void set_news_table_attn(bool on) { g_news_table_attn = on ? 1 : 0; }
int set_table_ratio(int v) {
  if (v < 1 || v > 1024) return NRMS_E_INVALID;
  g_table_ratio = v;
  return NRMS_OK;
}
// The projection is paid per call (and per rank) for the WHOLE source table -- ~0.9 us per 1,000 rows since its epilogue
// leaves through bulk tensor stores (2.3 us before) -- and a projected table beyond L2 (126 MB = 58 k rows) turns the
// 2,160-byte row gather into DRAM traffic.  Against that the per-sequence projection (K1 v6) costs ~60 ns per gathered
// row.  Gathered rows >= 4 x table rows selects the table path ("table_ratio" option / NRMS_TABLE_RATIO to experiment).
static int table_ratio() {
  if (g_table_ratio < 0) {
    const char* e = getenv("NRMS_TABLE_RATIO");
    g_table_ratio = e ? atoi(e) : 4;
    if (g_table_ratio < 1) g_table_ratio = 4;
  }
  return g_table_ratio;
}
static bool may_use_table(int64_t n_seq, int S, int64_t n_src_rows) {
  return n_src_rows > 0 && n_seq * S >= (int64_t)table_ratio() * n_src_rows;
}
static bool use_table_attn(int S, int idx_kind, int64_t n_seq, int64_t n_src_rows) {
  if (!may_use_table(n_seq, S, n_src_rows)) return false;
  // int32 rows = the user encoder over the news-vector table; int64 ids = the news encoder over the embedding table
  // (title length 20; a 50-token text, e.g. an abstract, takes the same kernel)
  if (idx_kind == 2) return S == 50 ? table_attn_enabled() : news_table_attn_enabled();
  if (idx_kind == 1) return news_table_attn_enabled();
  return false;
}

// ---- optional live timing of the attention launches (bench.py's roofline): CUDA events around every launch, kept per
//      kind: 0 = user encoder K1 v6, 1 = news encoder K1 v6, 2 = user encoder table attention, 3 = news encoder table attention
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

// sequences per launch pair: full waves of tiles (5 titles / 2 users per 128-row tile)
static int64_t fused_chunk_seq(int S, bool table_attn) {
  // Measured on the evaluate bench.  K1 v6 + K2: 4 waves 5.9 ms of encoder time, 8 waves 5.3, 16 waves 4.97, 32 waves
  // 4.92 (news 1.61 vs 1.66 ms at 16; users equal): every launch pays the K2 prologue (W_a into shared memory), the
  // pipeline fill and a tail; users stay at 16 so the fp16 context chunk (152 MB) is still mostly L2-resident between
  // K1 and K2.  Table path (K1g + K2): monotonic -- users 3.28 ms at 8 waves, 2.92 at 16, 2.73 at 32, 2.63 at 64, 2.60 in
  // one launch; news 1.50 / 1.35 / 1.30 / 1.25 / 1.24 -- so 64 waves (1.1 GB of context rows).
  static int env_waves = -1;
  if (env_waves < 0) { const char* e = getenv("NRMS_FUSED_WAVES"); env_waves = e ? atoi(e) : 0; if (env_waves < 0) env_waves = 0; }
  const int waves = env_waves ? env_waves : (table_attn ? 64 : (S == 20 ? 32 : 16));
  return (int64_t)num_sms() * (S == 20 ? 5 : 2) * waves;
}

// In-place LayerNorm(300) over the fp16 context rows [rows][320] that K1 hands to K2 (columns 300..319 stay zero):
// one warp per row, five half2 per lane, statistics in fp32.  (config-5 variant, builder-defined: DESIGN.md)
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

// workspace: [fp16 W_qkv copy (K1 v6)][fp16 W_a copy + score bounds][gather source: fp16 rows, or the projected table
// with its operand copies][fp16 context rows of one chunk]
struct Plan {
  bool table;           // table path possible for this call size (whatever the options say: sizes are an upper bound)
  int64_t chunk, first;
  size_t src_bytes, ctx_bytes, total;
};
static Plan make_plan(int64_t n_seq, int S, int64_t n_src_rows, bool rows16_given) {
  Plan p;
  p.table = may_use_table(n_seq, S, n_src_rows);
  p.chunk = fused_chunk_seq(S, p.table);
  p.first = n_seq < p.chunk ? n_seq : p.chunk;
  // dense input: the chunk's rows become the fp16 gather source; indexed input: the whole source table
  size_t b = rows16_given ? 0 : src16_bytes(n_src_rows > 0 ? n_src_rows : p.first * S);
  if (p.table) {
    const size_t t = k1g_table16_bytes(n_src_rows, rows16_given);
    if (t > b) b = t;
  }
  p.src_bytes = align_up(b, 1024);
  p.ctx_bytes = align_up((size_t)p.first * S * 640, 1024);
  p.total = W16_SLOT_BYTES + WA16_SLOT_BYTES + p.src_bytes + p.ctx_bytes;
  return p;
}

size_t tc_fused_workspace_bytes(int64_t n_seq, int S, int64_t n_src_rows, bool rows16_given) {
  if (n_seq <= 0 || (S != 20 && S != 50)) return (size_t)-1;
  return make_plan(n_seq, S, n_src_rows, rows16_given).total;
}

// src: fp32 rows [*, 300] (dense input or gather source), or nullptr when src16 (the fp16 copy [n_src_rows + 1][320] in
// pack_rows16's layout, e.g. the all-gathered news-vector table of evaluate) is given instead.
int tc_encoder_fused(const float* src, const void* src16_ext, int64_t n_src_rows, const void* idx, int idx_kind, int64_t n_seq,
                     int S, const float* wqkv, const float* bqkv, const float* wa, const float* ba, const float* qa,
                     float* out, void* workspace, size_t workspace_bytes, cudaStream_t st, const float* ln_gamma,
                     const float* ln_beta, int64_t hot_row) {
  if (S != 20 && S != 50) {
    set_error("fused encoder compiled for S = 20 or 50, got %d", S);
    return NRMS_E_UNSUPPORTED;
  }
  NRMS_CHECK_ARG(src != nullptr || (src16_ext != nullptr && idx_kind != 0), NRMS_E_INVALID, "no gather source");
  NRMS_CHECK_ARG(idx_kind == 0 || n_src_rows > 0, NRMS_E_INVALID, "indexed input needs the row count of its source table");
  const Plan p = make_plan(n_seq, S, idx_kind == 0 ? 0 : n_src_rows, src16_ext != nullptr);
  NRMS_CHECK_ARG(workspace && aligned16(workspace) && workspace_bytes >= p.total, NRMS_E_WORKSPACE,
                 "workspace too small: need %zu bytes", p.total);
  char* ws = reinterpret_cast<char*>(workspace);
  void* w16 = ws;
  void* wa16 = ws + W16_SLOT_BYTES;
  float* bound = reinterpret_cast<float*>(ws + W16_SLOT_BYTES + 128000);
  void* srcbuf = ws + W16_SLOT_BYTES + WA16_SLOT_BYTES;
  void* Cbuf = ws + W16_SLOT_BYTES + WA16_SLOT_BYTES + p.src_bytes;
  // The LayerNorm variant keeps the per-sequence projection (K1 v6 -> LayerNorm rows -> K2): the normalisation needs the
  // whole 300-wide context row, which in the table path is spread over 15 head warps.
  const bool table_attn = !ln_gamma && use_table_attn(S, idx_kind, n_seq, n_src_rows);
  alignas(64) CUtensorMap tw, twa;
  if (int rc = k2v2_prepare(wa, wa16, &twa, st)) return rc;
  const int tkind_table = (S == 50 && idx_kind == 2) ? 2 : 3;

  if (table_attn) {
    if (fused_pool_enabled()) hot_row = -1;      // K1f reads the table without the replicas
    if (int rc = k1g_project_table(src, src16_ext, n_src_rows, wqkv, bqkv, srcbuf, st, hot_row)) return rc;
    const void* table16 = k1g_table16_ptr(srcbuf);
    // Bound on the attention scores over the projected table: the user-encoder attention picks the plain or the
    // row-shifted softmax form from it (the shifted form costs that kernel ~5 %).  The news-encoder attention is bound by
    // its gather and runs the shifted form for free (measured 1.032 vs 1.033 ms per 65,238 titles): no bound pass there.
    // The bound pass reads the whole projected table (2,160 B per row: 42 us at 65 k rows, 84 us at 130 k, 0.33 ms at the
    // 522 k rows of the 8-GPU weak-scaling line); the shifted form costs the user kernel ~2.8 us per 1,000 users (measured
    // at 2 GPUs: 469 vs 418 us per 18,288-user launch).  Only when the table is very large against the call -- more than
    // 5 rows per user: every rank of an 8-GPU evaluate holds ALL news but a 1/8 share of the users -- is the pass
    // skipped and the shifted form run unconditionally.
    if (S == 50 && n_src_rows <= 5 * n_seq) {
      if (int rc = k1f_qk_bound(table16, n_src_rows, bound, st)) return rc;
    } else {
      bound = nullptr;
    }
    if (fused_pool_enabled()) {
      // K1f: attention + additive pooling in ONE launch over the whole call; no context rows, no chunking
      K1Timer timer(st, n_seq, tkind_table);
      return k1f_run(S, idx_kind, table16, n_src_rows, idx, n_seq, wa16, ba, qa, bound, out, st);
    }
    const size_t idx_elem = idx_kind == 1 ? 8 : 4;
    for (int64_t s0 = 0; s0 < n_seq; s0 += p.chunk) {
      const int64_t n = (n_seq - s0 < p.chunk) ? (n_seq - s0) : p.chunk;
      const void* idx_c = (const char*)idx + (size_t)s0 * S * idx_elem;
      {
        K1Timer timer(st, n, tkind_table);
        if (int rc = k1g_run_seq(S, idx_kind, table16, n_src_rows, idx_c, n, Cbuf, bound, st, hot_row)) return rc;
      }
      if (int rc = k2v2_run(S, twa, Cbuf, n, ba, qa, out + s0 * D, st)) return rc;
    }
    return NRMS_OK;
  }

  // ---- per-sequence projection: K1 v6 -> [LayerNorm] -> K2 ----
  if (int rc = pack_weights_k1(wqkv, bqkv, w16, &tw, st)) return rc;
  const void* src16 = src16_ext;
  if (idx_kind != 0 && src16 == nullptr) {
    if (int rc = pack_rows16(src, n_src_rows, srcbuf, st)) return rc;
    src16 = srcbuf;
  }
  const size_t idx_elem = idx_kind == 1 ? 8 : 4;
  for (int64_t s0 = 0; s0 < n_seq; s0 += p.chunk) {
    const int64_t n = (n_seq - s0 < p.chunk) ? (n_seq - s0) : p.chunk;
    const void* idx_c = idx_kind == 0 ? nullptr : (const void*)((const char*)idx + (size_t)s0 * S * idx_elem);
    if (idx_kind == 0) {   // dense rows: this chunk's rows become the fp16 gather source
      if (int rc = pack_rows16(src + s0 * S * D, n * S, srcbuf, st)) return rc;
      src16 = srcbuf;
    }
    {
      K1Timer timer(st, n, (idx_kind != 1 && S == 50) ? 0 : 1);
      if (int rc = k1v6_run(S, tw, src16, idx_c, idx_kind, n, (int)(idx_kind == 0 ? n * S : n_src_rows), Cbuf, st)) return rc;
    }
    if (ln_gamma) {
      int64_t lb = (n * S + 7) / 8;
      if (lb > (int64_t)num_sms() * 8) lb = (int64_t)num_sms() * 8;
      layernorm_f16_rows_kernel<<<(unsigned)lb, 256, 0, st>>>(reinterpret_cast<__half*>(Cbuf), ln_gamma, ln_beta, n * S, 1e-5f);
      NRMS_LAUNCH_CHECK("layernorm_f16_rows_kernel");
    }
    if (int rc = k2v2_run(S, twa, Cbuf, n, ba, qa, out + s0 * D, st)) return rc;
  }
  return NRMS_OK;
}

}  // namespace nrms
