// Host side of the encoder entry points (C-ABI): argument checks, stash/workspace carving and
// the launch sequence.  Kernels live in encoder_kernels.cuh / sgemm.cuh / tc_gemm.cuh.
#include <string.h>
#include "common.cuh"
#include "encoder_kernels.cuh"
#include "sgemm.cuh"
#include "tc_api.cuh"

namespace nrms {

// ---- stash layout: X [rows,300] | QKV [rows,900] | C [rows,300] | T [rows,200] | w [rows] ----
// LayerNorm variant (config 5): + CN [rows,300] (normalised context, the additive block's input) | stats [rows,2]
struct Stash {
  float *x, *qkv, *c, *t, *w, *cn, *stats;
  size_t bytes;
};
// optional LayerNorm between the self-attention context and the additive block (nullptr = reference NRMS)
struct LnArgs {
  const float* gamma;
  const float* beta;
  float* d_gamma;
  float* d_beta;
};
constexpr float LN_EPS = 1e-5f;    // torch.nn.LayerNorm default
static Stash carve_stash(void* base, int64_t rows, bool ln = false) {
  Stash s;
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t nfloat) {
    float* r = reinterpret_cast<float*>(p + off);
    off += align_up(nfloat * sizeof(float), 256);
    return r;
  };
  s.x = take((size_t)rows * D);
  s.qkv = take((size_t)rows * D3);
  s.c = take((size_t)rows * D);
  s.t = take((size_t)rows * QD);
  s.w = take((size_t)rows);
  s.cn = ln ? take((size_t)rows * D) : nullptr;
  s.stats = ln ? take((size_t)rows * 2) : nullptr;
  s.bytes = off;
  return s;
}

static bool g_train_attn_mma = true;   // nrms_set_option("train_attn_mma"): tensor-mode title attention on mma.sync tiles
constexpr int64_t INFER_CHUNK_ROWS = 2048 * 20;  // rows processed per pass in inference (reference batch 2048 titles)
constexpr int REDUCE_BLOCKS = 296;

// Tensor mode adds the transposed operand copies of the weight-gradient GEMMs (see encoder_core_bwd):
// t1 [900, ldr] | t2 [300, ldr] (ldr = rows rounded up to 4) | wt [300, 900].
struct BwdWs {
  float *d_c, *d_u, *d_qkv, *d_x, *partial, *t1, *t2, *wt;
  int64_t ldr;
  size_t bytes;
};
static BwdWs carve_bwd(void* base, int64_t rows, bool need_dx, int mode) {
  BwdWs s;
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t nfloat) {
    float* r = reinterpret_cast<float*>(p + off);
    off += align_up(nfloat * sizeof(float), 256);
    return r;
  };
  s.d_c = take((size_t)rows * D);
  s.d_u = take((size_t)rows * QD);
  s.d_qkv = take((size_t)rows * D3);
  s.d_x = need_dx ? take((size_t)rows * D) : nullptr;
  s.partial = take((size_t)REDUCE_BLOCKS * D3);
  s.ldr = (rows + 3) / 4 * 4;
  s.t1 = s.t2 = s.wt = nullptr;
  if (mode == NRMS_MODE_TF32) {
    s.t1 = take((size_t)s.ldr * D3);
    s.t2 = take((size_t)s.ldr * D);
    s.wt = take((size_t)D3 * D);
  }
  s.bytes = off;
  return s;
}

template <int S, int HC>
static cudaError_t launch_attention_fwd(const float* qkv, float* ctx, int64_t n_seq, float p, uint64_t seed,
                                        uint64_t offset, cudaStream_t st, const int32_t* lengths = nullptr) {
  constexpr int threads = ((S * HC + 31) / 32) * 32;
  const size_t smem = 2 * S * HC * DH * sizeof(float);
  static bool configured[64] = {false};
  if (cudaError_t e = set_max_dynamic_smem(attention_fwd_kernel<S, HC>, (int)smem, configured)) return e;
  int64_t gx = n_seq < (int64_t)num_sms() * 8 ? n_seq : (int64_t)num_sms() * 8;
  dim3 grid((unsigned)gx, H / HC);
  const float scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  attention_fwd_kernel<S, HC><<<grid, threads, smem, st>>>(qkv, ctx, n_seq, p, scale, seed, offset, lengths);
  count_launch();
  return cudaGetLastError();
}

template <int S, int HC>
static cudaError_t launch_attention_bwd(const float* qkv, const float* d_ctx, float* d_qkv, int64_t n_seq, float p,
                                        uint64_t seed, uint64_t offset, cudaStream_t st) {
  constexpr int threads = ((S * HC + 31) / 32) * 32;
  const size_t smem = (4 * S * HC * DH + 2 * HC * S * (S + 1)) * sizeof(float);
  static bool configured[64] = {false};
  if (cudaError_t e = set_max_dynamic_smem(attention_bwd_kernel<S, HC>, (int)smem, configured)) return e;
  int64_t gx = n_seq < (int64_t)num_sms() * 4 ? n_seq : (int64_t)num_sms() * 4;
  dim3 grid((unsigned)gx, H / HC);
  const float scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  attention_bwd_kernel<S, HC><<<grid, threads, smem, st>>>(qkv, d_ctx, d_qkv, n_seq, p, scale, seed, offset);
  count_launch();
  return cudaGetLastError();
}

static int gemm_nt_bias(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, float* C,
                        int64_t ldc, int64_t M, int N, int K, int mode, cudaStream_t st) {
  if (mode == NRMS_MODE_TF32) return tc_gemm_nt(A, lda, B, ldb, bias, C, ldc, M, N, K, st);
  cudaError_t e = sgemm_launch<0, 0, EPI_STORE>(A, lda, B, ldb, bias, C, ldc, M, N, K, 1, st);
  if (e != cudaSuccess) return cuda_fail(e, "sgemm_nt");
  return NRMS_OK;
}

// additive pooling kernels at every compiled sequence length: 20 (titles), 50 (history / abstracts) and 2..4 (the
// final attention over the element vectors of model/Exp1, reference src/model/Exp1/news_encoder.py:106-110)
static bool additive_len_ok(int S) { return S == 20 || S == 50 || (S >= 2 && S <= 4); }
static cudaError_t launch_additive_fwd(const float* cin, float* t, const float* qa, float* wv, float* out, int64_t n_seq,
                                       int S, cudaStream_t st) {
  const unsigned gx = (unsigned)(n_seq < (int64_t)num_sms() * 8 ? n_seq : (int64_t)num_sms() * 8);
  switch (S) {
    case 20: additive_fwd_kernel<20><<<gx, 256, 0, st>>>(cin, t, qa, wv, out, n_seq); break;
    case 50: additive_fwd_kernel<50><<<gx, 256, 0, st>>>(cin, t, qa, wv, out, n_seq); break;
    case 2: additive_fwd_kernel<2><<<gx, 256, 0, st>>>(cin, t, qa, wv, out, n_seq); break;
    case 3: additive_fwd_kernel<3><<<gx, 256, 0, st>>>(cin, t, qa, wv, out, n_seq); break;
    default: additive_fwd_kernel<4><<<gx, 256, 0, st>>>(cin, t, qa, wv, out, n_seq); break;
  }
  count_launch();
  return cudaGetLastError();
}
static cudaError_t launch_additive_bwd(const float* d_out, const float* cin, const float* t, const float* wv,
                                       const float* qa, float* d_c, float* d_u, float* partial, int64_t n_seq, int S,
                                       int nb, cudaStream_t st) {
  switch (S) {
    case 20: additive_bwd_kernel<20><<<nb, 256, 0, st>>>(d_out, cin, t, wv, qa, d_c, d_u, partial, n_seq); break;
    case 50: additive_bwd_kernel<50><<<nb, 256, 0, st>>>(d_out, cin, t, wv, qa, d_c, d_u, partial, n_seq); break;
    case 2: additive_bwd_kernel<2><<<nb, 256, 0, st>>>(d_out, cin, t, wv, qa, d_c, d_u, partial, n_seq); break;
    case 3: additive_bwd_kernel<3><<<nb, 256, 0, st>>>(d_out, cin, t, wv, qa, d_c, d_u, partial, n_seq); break;
    default: additive_bwd_kernel<4><<<nb, 256, 0, st>>>(d_out, cin, t, wv, qa, d_c, d_u, partial, n_seq); break;
  }
  count_launch();
  return cudaGetLastError();
}

// Forward over `n_seq` sequences whose input rows X are already materialised in st.x.
static int encoder_core_fwd(const Stash& s, int64_t n_seq, int S, const float* wqkv, const float* bqkv,
                            const float* wa, const float* ba, const float* qa, float* out, float p2,
                            uint64_t seed, uint64_t offset, int64_t row_base, int mode, cudaStream_t st,
                            const LnArgs* ln = nullptr) {
  const int64_t rows = n_seq * S;
  int rc = gemm_nt_bias(s.x, D, wqkv, D, bqkv, s.qkv, D3, rows, D3, D, mode, st);
  if (rc) return rc;
  // dropout #2 mask indices are global row indices: offset the Philox counter by row_base*D/4
  const uint64_t off2 = offset + (uint64_t)row_base * D / 4;
  cudaError_t e;
  if (S == 20 && mode == NRMS_MODE_TF32 && g_train_attn_mma) {
    // tensor mode, titles: exp-softmax attention on mma.sync TF32 tiles (attn_mma.cu)
    if (int rc2 = attn_mma_fwd(s.qkv, s.c, n_seq, p2, seed, off2, st)) return rc2;
    e = cudaSuccess;
  } else if (S == 20) e = launch_attention_fwd<20, 15>(s.qkv, s.c, n_seq, p2, seed, off2, st);
  else e = launch_attention_fwd<50, 5>(s.qkv, s.c, n_seq, p2, seed, off2, st);
  if (e != cudaSuccess) return cuda_fail(e, "attention_fwd");
  const float* cin = s.c;
  if (ln) {
    int64_t lb = (rows + 7) / 8;
    if (lb > (int64_t)num_sms() * 8) lb = (int64_t)num_sms() * 8;
    layernorm_fwd_kernel<<<(unsigned)lb, 256, 0, st>>>(s.c, ln->gamma, ln->beta, s.cn, s.stats, rows, LN_EPS);
    NRMS_LAUNCH_CHECK("layernorm_fwd");
    cin = s.cn;
  }
  rc = gemm_nt_bias(cin, D, wa, D, ba, s.t, QD, rows, QD, D, mode, st);
  if (rc) return rc;
  if (cudaError_t e2 = launch_additive_fwd(cin, s.t, qa, s.w, out, n_seq, S, st)) return cuda_fail(e2, "additive_fwd");
  return NRMS_OK;
}

// Backward of the additive block (additive.py:27-53) over n_seq sequences of S rows: cin = its input rows, t / wv = the
// tanh activations and softmax weights its forward saved.  d_c [rows,300] is OVERWRITTEN with dL/d(cin); d_wa, d_ba,
// d_qa are accumulated into.  Tensor mode (S = 20 / 50 only) needs the transposed operand copies t1 / t2 / wt.
struct AddBwdWs {
  float *d_u, *partial, *t1, *t2, *wt;
  int64_t ldr;
};
static int additive_block_bwd(const float* d_out, const float* cin, const float* t, const float* wv, const float* wa,
                              const float* qa, float* d_c, const AddBwdWs& w, int64_t n_seq, int S, float* d_wa,
                              float* d_ba, float* d_qa, bool tc, cudaStream_t st) {
  const int64_t rows = n_seq * S;
  // 4 CTAs per SM; the partial buffer (REDUCE_BLOCKS x 900 floats) holds two 200-wide blocks per CTA
  constexpr int ADD_BWD_BLOCKS = 592;
  static_assert(2 * ADD_BWD_BLOCKS * QD <= REDUCE_BLOCKS * D3, "partial buffer too small");
  int nb = (int)(n_seq < ADD_BWD_BLOCKS ? n_seq : ADD_BWD_BLOCKS);
  if (cudaError_t e2 = launch_additive_bwd(d_out, cin, t, wv, qa, d_c, w.d_u, w.partial, n_seq, S, nb, st))
    return cuda_fail(e2, "additive_bwd");
  launch_partial_reduce_accum(w.partial, nb, QD, QD, d_qa, st);
  NRMS_LAUNCH_CHECK("dqa_reduce");
  // d_ba = colsum(dU): per-thread sums inside additive_bwd_kernel, second block of the partial buffer
  launch_partial_reduce_accum(w.partial + (size_t)nb * QD, nb, QD, QD, d_ba, st);
  NRMS_LAUNCH_CHECK("dba_reduce");
  int splits = (int)((rows + 4095) / 4096);
  if (splits > 64) splits = 64;
  if (tc) {
    // d_wa[200,300] += dU^T C
    if (int rc = transpose_f32(w.d_u, QD, w.t1, w.ldr, rows, QD, st)) return rc;
    if (int rc = transpose_f32(cin, D, w.t2, w.ldr, rows, D, st)) return rc;
    if (int rc = tc_gemm_nt_ex(w.t1, w.ldr, w.t2, w.ldr, nullptr, d_wa, D, QD, D, (int)rows,
                               tc_gemm_auto_splits(QD, D, (int)rows), TC_EPI_ATOMIC, st)) return rc;
    // d_c += dU * Wa
    if (int rc = transpose_f32(wa, D, w.wt, QD, QD, D, st)) return rc;
    if (int rc = tc_gemm_nt_ex(w.d_u, QD, w.wt, QD, nullptr, d_c, D, rows, D, QD, 1, TC_EPI_ACCUM, st)) return rc;
  } else {
    // d_wa[200,300] += dU^T C   (reduction over rows, split-K + fp32 atomics)
    cudaError_t e = sgemm_launch<1, 1, EPI_ATOMIC>(w.d_u, QD, cin, D, nullptr, d_wa, D, QD, D, rows, splits, st);
    if (e != cudaSuccess) return cuda_fail(e, "sgemm dWa");
    // d_c += dU * Wa   ([rows,200] x [200,300])
    e = sgemm_launch<0, 1, EPI_ACCUM>(w.d_u, QD, wa, D, nullptr, d_c, D, rows, D, QD, 1, st);
    if (e != cudaSuccess) return cuda_fail(e, "sgemm dC");
  }
  return NRMS_OK;
}

// Backward over the whole stash.  d_x (rows x 300) receives dL/dX.
static int encoder_core_bwd(const Stash& s, const BwdWs& w, const float* d_out, int64_t n_seq, int S,
                            const float* wqkv, const float* wa, const float* qa, float* d_x, float* d_wqkv,
                            float* d_bqkv, float* d_wa, float* d_ba, float* d_qa, float p2, uint64_t seed,
                            uint64_t offset, int mode, cudaStream_t st, const LnArgs* ln = nullptr) {
  // FP32 mode: every contraction on the CUDA cores (reference-exact up to summation order).  Tensor mode: the four
  // GEMMs run on tcgen05 (TF32 operands rounded by the TMA unit, fp32 accumulation) through the K-major NT kernel:
  //   dC  += dU   Wa        = NT(dU   [rows,200], Wa^T   [300,200])
  //   dX   = dQKV Wqkv      = NT(dQKV [rows,900], Wqkv^T [300,900])
  //   dWa   += dU^T   C     = NT(dU^T   [200,rows], C^T [300,rows])      split along K = rows, atomic epilogue
  //   dWqkv += dQKV^T X     = NT(dQKV^T [900,rows], X^T [300,rows])
  // so the row-major activations are transposed once per use (1.2 KB/row of extra traffic, a fraction of what the
  // fp32 SGEMMs cost: measured 6.3 ms -> see profiles/).
  const bool tc = (mode == NRMS_MODE_TF32);
  const int64_t rows = n_seq * S;
  const float* cin = ln ? s.cn : s.c;        // input of the additive block
  {
    const AddBwdWs aw{w.d_u, w.partial, w.t1, w.t2, w.wt, w.ldr};
    if (int rc = additive_block_bwd(d_out, cin, s.t, s.w, wa, qa, w.d_c, aw, n_seq, S, d_wa, d_ba, d_qa, tc, st)) return rc;
  }
  int cb = (int)(rows < REDUCE_BLOCKS ? rows : REDUCE_BLOCKS);
  int splits = (int)((rows + 4095) / 4096);
  if (splits > 64) splits = 64;
  cudaError_t e;
  if (ln) {   // d_c holds dL/dCN: through the LayerNorm, in place; d_gamma / d_beta via per-block partial sums
    int lb = (int)((rows + 7) / 8 < REDUCE_BLOCKS ? (rows + 7) / 8 : REDUCE_BLOCKS);
    layernorm_bwd_kernel<<<lb, 256, 0, st>>>(w.d_c, s.c, s.stats, ln->gamma, w.partial, rows);
    NRMS_LAUNCH_CHECK("layernorm_bwd");
    launch_partial_reduce_accum(w.partial, lb, 2 * D, D, ln->d_gamma, st);
    NRMS_LAUNCH_CHECK("dgamma_reduce");
    launch_partial_reduce_accum(w.partial + D, lb, 2 * D, D, ln->d_beta, st);
    NRMS_LAUNCH_CHECK("dbeta_reduce");
  }
  // attention backward (applies the dropout-2 mask to d_c on load)
#ifndef NRMS_ATTN_BWD_HC20
#define NRMS_ATTN_BWD_HC20 5        // heads per CTA of the title-length attention backward (measured: 5 -> 807 us, 15 -> 1,086 us per 7,040 titles)
#endif
  bool dqkv_t_written = false;
  if (S == 20 && tc && g_train_attn_mma) {
    // the kernel also leaves dQKV^T in t1 (the K-major operand of the dWqkv contraction below)
    if (int rc2 = attn_mma_bwd(s.qkv, w.d_c, w.d_qkv, w.t1, w.ldr, n_seq, p2, seed, offset, st)) return rc2;
    dqkv_t_written = true;
    e = cudaSuccess;
  } else if (S == 20) e = launch_attention_bwd<20, NRMS_ATTN_BWD_HC20>(s.qkv, w.d_c, w.d_qkv, n_seq, p2, seed, offset, st);
  else e = launch_attention_bwd<50, 5>(s.qkv, w.d_c, w.d_qkv, n_seq, p2, seed, offset, st);
  if (e != cudaSuccess) return cuda_fail(e, "attention_bwd");
  // d_bqkv = colsum(dQKV)
  colsum_partial_kernel<<<cb, 256, 0, st>>>(w.d_qkv, rows, D3, w.partial);
  NRMS_LAUNCH_CHECK("colsum_dqkv");
  launch_partial_reduce_accum(w.partial, cb, D3, D3, d_bqkv, st);
  NRMS_LAUNCH_CHECK("dbqkv_reduce");
  if (tc) {
    // d_wqkv[900,300] += dQKV^T X
    if (!dqkv_t_written)
      if (int rc = transpose_f32(w.d_qkv, D3, w.t1, w.ldr, rows, D3, st)) return rc;
    if (int rc = transpose_f32(s.x, D, w.t2, w.ldr, rows, D, st)) return rc;
    if (int rc = tc_gemm_nt_ex(w.t1, w.ldr, w.t2, w.ldr, nullptr, d_wqkv, D, D3, D, (int)rows,
                               tc_gemm_auto_splits(D3, D, (int)rows), TC_EPI_ATOMIC, st)) return rc;
    // d_x = dQKV * Wqkv
    if (int rc = transpose_f32(wqkv, D, w.wt, D3, D3, D, st)) return rc;
    if (int rc = tc_gemm_nt_ex(w.d_qkv, D3, w.wt, D3, nullptr, d_x, D, rows, D, D3, 1, TC_EPI_STORE, st)) return rc;
  } else {
    // d_wqkv[900,300] += dQKV^T X
    e = sgemm_launch<1, 1, EPI_ATOMIC>(w.d_qkv, D3, s.x, D, nullptr, d_wqkv, D, D3, D, rows, splits, st);
    if (e != cudaSuccess) return cuda_fail(e, "sgemm dWqkv");
    // d_x = dQKV * Wqkv   ([rows,900] x [900,300])
    e = sgemm_launch<0, 1, EPI_STORE>(w.d_qkv, D3, wqkv, D, nullptr, d_x, D, rows, D, D3, 1, st);
    if (e != cudaSuccess) return cuda_fail(e, "sgemm dX");
  }
  return NRMS_OK;
}

static int check_common(int S, int mode) {
  NRMS_CHECK_ARG(S == 20 || S == 50, NRMS_E_UNSUPPORTED, "sequence length %d unsupported (compiled: 20, 50)", S);
  NRMS_CHECK_ARG(mode == NRMS_MODE_FP32 || mode == NRMS_MODE_TF32, NRMS_E_INVALID, "bad mode %d", mode);
  return NRMS_OK;
}

}  // namespace nrms

using namespace nrms;

extern "C" {

size_t nrms_encoder_stash_bytes(int64_t n_seq, int S) {
  if (n_seq <= 0 || S <= 0) return 0;
  return carve_stash(nullptr, n_seq * S).bytes;
}

size_t nrms_encoder_ln_stash_bytes(int64_t n_seq, int S) {
  if (n_seq <= 0 || S <= 0) return 0;
  return carve_stash(nullptr, n_seq * S, true).bytes;
}

size_t nrms_encoder_fwd_workspace_bytes(int64_t n_seq, int S, int mode, int training, int64_t n_src_rows) {
  if (n_seq <= 0 || S <= 0) return 0;
  if (training) return 256;
  if (mode == NRMS_MODE_TF32) {
    size_t fused = tc_fused_workspace_bytes(n_seq, S, n_src_rows);
    if (fused != (size_t)-1) return fused + 256;
  }
  int64_t chunk_seq = INFER_CHUNK_ROWS / S;
  if (n_seq < chunk_seq) chunk_seq = n_seq;
  return carve_stash(nullptr, chunk_seq * S, true).bytes + 256;     // room for the LayerNorm variant's fields too
}

size_t nrms_encoder_bwd_workspace_bytes(int64_t n_seq, int S, int mode) {
  if (n_seq <= 0 || S <= 0) return 0;
  return carve_bwd(nullptr, n_seq * S, true, mode).bytes + 256;
}

static int check_ln(const LnArgs* ln, bool bwd) {
  if (!ln) return NRMS_OK;
  NRMS_CHECK_ARG(ln->gamma && aligned16(ln->gamma), NRMS_E_INVALID, "LayerNorm weight missing or misaligned");
  if (bwd) NRMS_CHECK_ARG(ln->d_gamma && ln->d_beta, NRMS_E_INVALID, "LayerNorm gradient pointers missing");
  else NRMS_CHECK_ARG(ln->beta && aligned16(ln->beta), NRMS_E_INVALID, "LayerNorm bias missing or misaligned");
  return NRMS_OK;
}

static int news_encoder_fwd_impl(const int64_t* tokens, int64_t n_titles, int L, const float* emb, int64_t num_words,
                                 const float* wqkv, const float* bqkv, const float* wa, const float* ba, const float* qa,
                                 float* out, void* stash, void* workspace, size_t workspace_bytes, float dropout_p,
                                 uint64_t seed, uint64_t offset, int mode, void* stream, const LnArgs* ln,
                                 const int64_t* news_rows = nullptr) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (int rc = check_ln(ln, false)) return rc;
  if (int rc = check_common(L, mode)) return rc;
  NRMS_CHECK_ARG(news_rows == nullptr || stash != nullptr, NRMS_E_UNSUPPORTED,
                 "the index-only form (token table + news rows) is the training form: pass a stash");
  // training (a stash is requested) is compiled for the title length; inference also takes 50-token texts (the
  // abstract encoder of the reference's Exp1 model is this same block, src/model/Exp1/news_encoder.py:10-34)
  NRMS_CHECK_ARG(L == 20 || (L == 50 && stash == nullptr), NRMS_E_UNSUPPORTED,
                 "news encoder compiled for title length 20 (inference: 20 or 50), got %d", L);
  NRMS_CHECK_ARG(n_titles >= 0 && num_words > 0, NRMS_E_INVALID, "bad sizes");
  if (n_titles == 0) return NRMS_OK;
  NRMS_CHECK_ARG(tokens && emb && wqkv && bqkv && wa && ba && qa && out, NRMS_E_INVALID, "null pointer");
  NRMS_CHECK_ARG(aligned16(emb) && aligned16(wqkv) && aligned16(bqkv) && aligned16(wa) && aligned16(ba) && aligned16(out),
                 NRMS_E_INVALID, "pointers must be 16-byte aligned");
  NRMS_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, NRMS_E_INVALID, "dropout_p out of range");
  const float scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;

  if (stash) {  // training: everything kept for backward
    NRMS_CHECK_ARG(aligned16(stash), NRMS_E_INVALID, "stash misaligned");
    Stash s = carve_stash(stash, n_titles * L, ln != nullptr);
    const int64_t rows = n_titles * L;
    int64_t gb = (rows + 7) / 8;
    if (gb > (int64_t)num_sms() * 16) gb = (int64_t)num_sms() * 16;
    gather_embedding_kernel<<<(unsigned)gb, 256, 0, st>>>(tokens, rows, emb, s.x, dropout_p, scale, seed, offset,
                                                          news_rows, L);
    NRMS_LAUNCH_CHECK("gather_embedding");
    return encoder_core_fwd(s, n_titles, L, wqkv, bqkv, wa, ba, qa, out, dropout_p, seed, offset, 0, mode, st, ln);
  }
  // inference
  if (mode == NRMS_MODE_TF32 && dropout_p == 0.f && tc_fused_workspace_bytes(n_titles, L, num_words) != (size_t)-1) {
    return tc_encoder_fused(emb, nullptr, num_words, tokens, 1, n_titles, L, wqkv, bqkv, wa, ba, qa, out, workspace,
                            workspace_bytes, st, ln ? ln->gamma : nullptr, ln ? ln->beta : nullptr, /*hot_row = padding_idx*/ 0);
  }
  const int64_t chunk_seq = INFER_CHUNK_ROWS / L;
  const int64_t first = n_titles < chunk_seq ? n_titles : chunk_seq;
  NRMS_CHECK_ARG(workspace && aligned16(workspace) && workspace_bytes >= carve_stash(nullptr, first * L, true).bytes,
                 NRMS_E_WORKSPACE, "workspace too small: need %zu bytes", carve_stash(nullptr, first * L, true).bytes);
  for (int64_t s0 = 0; s0 < n_titles; s0 += chunk_seq) {
    const int64_t n = (n_titles - s0 < chunk_seq) ? (n_titles - s0) : chunk_seq;
    Stash s = carve_stash(workspace, n * L, ln != nullptr);
    const int64_t rows = n * L;
    int64_t gb = (rows + 7) / 8;
    if (gb > (int64_t)num_sms() * 16) gb = (int64_t)num_sms() * 16;
    // global row index keys the dropout stream, so a chunked pass equals a single pass
    gather_embedding_kernel<<<(unsigned)gb, 256, 0, st>>>(tokens + s0 * L, rows, emb, s.x, dropout_p, scale, seed,
                                                          offset + (uint64_t)s0 * L * D / 4);
    NRMS_LAUNCH_CHECK("gather_embedding");
    int rc = encoder_core_fwd(s, n, L, wqkv, bqkv, wa, ba, qa, out + s0 * D, dropout_p, seed, offset, s0 * L, mode, st,
                              ln);
    if (rc) return rc;
  }
  return NRMS_OK;
}

int nrms_news_encoder_fwd(const int64_t* tokens, int64_t n_titles, int L, const float* emb, int64_t num_words,
                          const float* wqkv, const float* bqkv, const float* wa, const float* ba, const float* qa,
                          float* out, void* stash, void* workspace, size_t workspace_bytes, float dropout_p,
                          uint64_t seed, uint64_t offset, int mode, void* stream) {
  return news_encoder_fwd_impl(tokens, n_titles, L, emb, num_words, wqkv, bqkv, wa, ba, qa, out, stash, workspace,
                               workspace_bytes, dropout_p, seed, offset, mode, stream, nullptr);
}

int nrms_news_encoder_i32_fwd(const int32_t* tokens, int64_t n_titles, int L, const float* emb, int64_t num_words,
                              const float* wqkv, const float* bqkv, const float* wa, const float* ba, const float* qa,
                              float* out, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (int rc = check_common(L, NRMS_MODE_TF32)) return rc;
  NRMS_CHECK_ARG(n_titles >= 0 && num_words > 0, NRMS_E_INVALID, "bad sizes");
  if (n_titles == 0) return NRMS_OK;
  NRMS_CHECK_ARG(tokens && emb && wqkv && bqkv && wa && ba && qa && out, NRMS_E_INVALID, "null pointer");
  NRMS_CHECK_ARG(aligned16(emb) && aligned16(wqkv) && aligned16(bqkv) && aligned16(wa) && aligned16(ba) && aligned16(out),
                 NRMS_E_INVALID, "pointers must be 16-byte aligned");
  NRMS_CHECK_ARG(tc_fused_workspace_bytes(n_titles, L, num_words) != (size_t)-1, NRMS_E_UNSUPPORTED, "title length not compiled");
  return tc_encoder_fused(emb, nullptr, num_words, tokens, 2, n_titles, L, wqkv, bqkv, wa, ba, qa, out, workspace,
                          workspace_bytes, st, nullptr, nullptr, /*hot_row = padding_idx*/ 0);
}

int nrms_news_encoder_ln_fwd(const int64_t* tokens, int64_t n_titles, int L, const float* emb, int64_t num_words,
                             const float* wqkv, const float* bqkv, const float* ln_gamma, const float* ln_beta,
                             const float* wa, const float* ba, const float* qa, float* out, void* stash, void* workspace,
                             size_t workspace_bytes, float dropout_p, uint64_t seed, uint64_t offset, int mode,
                             void* stream) {
  const LnArgs ln{ln_gamma, ln_beta, nullptr, nullptr};
  return news_encoder_fwd_impl(tokens, n_titles, L, emb, num_words, wqkv, bqkv, wa, ba, qa, out, stash, workspace,
                               workspace_bytes, dropout_p, seed, offset, mode, stream, &ln);
}

static int news_encoder_bwd_impl(const float* d_out, const int64_t* tokens, int64_t n_titles, int L, int64_t num_words,
                                 const float* wqkv, const float* wa, const float* qa, const void* stash, float* d_emb,
                                 float* d_wqkv, float* d_bqkv, float* d_wa, float* d_ba, float* d_qa, void* workspace,
                                 size_t workspace_bytes, float dropout_p, uint64_t seed, uint64_t offset, int mode,
                                 void* stream, const LnArgs* ln, const int64_t* news_rows = nullptr) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (int rc = check_ln(ln, true)) return rc;
  if (int rc = check_common(L, mode)) return rc;
  NRMS_CHECK_ARG(L == 20, NRMS_E_UNSUPPORTED, "news encoder compiled for title length 20, got %d", L);
  if (n_titles == 0) return NRMS_OK;
  NRMS_CHECK_ARG(n_titles > 0 && num_words > 0, NRMS_E_INVALID, "bad sizes");
  NRMS_CHECK_ARG(d_out && tokens && wqkv && wa && qa && stash && d_emb && d_wqkv && d_bqkv && d_wa && d_ba && d_qa,
                 NRMS_E_INVALID, "null pointer");
  NRMS_CHECK_ARG(aligned16(d_out) && aligned16(stash) && aligned16(d_emb) && aligned16(d_wqkv) && aligned16(d_wa),
                 NRMS_E_INVALID, "pointers must be 16-byte aligned");
  const int64_t rows = n_titles * L;
  NRMS_CHECK_ARG(workspace && aligned16(workspace) && workspace_bytes >= carve_bwd(nullptr, rows, true, mode).bytes,
                 NRMS_E_WORKSPACE, "workspace too small: need %zu bytes", carve_bwd(nullptr, rows, true, mode).bytes);
  Stash s = carve_stash(const_cast<void*>(stash), rows, ln != nullptr);
  BwdWs w = carve_bwd(workspace, rows, true, mode);
  int rc = encoder_core_bwd(s, w, d_out, n_titles, L, wqkv, wa, qa, w.d_x, d_wqkv, d_bqkv, d_wa, d_ba, d_qa,
                            dropout_p, seed, offset, mode, st, ln);
  if (rc) return rc;
  const float scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  int64_t gb = (rows + 7) / 8;
  if (gb > (int64_t)num_sms() * 16) gb = (int64_t)num_sms() * 16;
  scatter_embedding_grad_kernel<<<(unsigned)gb, 256, 0, st>>>(tokens, rows, w.d_x, d_emb, dropout_p, scale, seed, offset,
                                                              news_rows, L);
  NRMS_LAUNCH_CHECK("scatter_embedding_grad");
  return NRMS_OK;
}

int nrms_news_encoder_bwd(const float* d_out, const int64_t* tokens, int64_t n_titles, int L, int64_t num_words,
                          const float* wqkv, const float* wa, const float* qa, const void* stash, float* d_emb,
                          float* d_wqkv, float* d_bqkv, float* d_wa, float* d_ba, float* d_qa, void* workspace,
                          size_t workspace_bytes, float dropout_p, uint64_t seed, uint64_t offset, int mode,
                          void* stream) {
  return news_encoder_bwd_impl(d_out, tokens, n_titles, L, num_words, wqkv, wa, qa, stash, d_emb, d_wqkv, d_bqkv, d_wa,
                               d_ba, d_qa, workspace, workspace_bytes, dropout_p, seed, offset, mode, stream, nullptr);
}

int nrms_news_encoder_ln_bwd(const float* d_out, const int64_t* tokens, int64_t n_titles, int L, int64_t num_words,
                             const float* wqkv, const float* ln_gamma, const float* wa, const float* qa,
                             const void* stash, float* d_emb, float* d_wqkv, float* d_bqkv, float* d_ln_gamma,
                             float* d_ln_beta, float* d_wa, float* d_ba, float* d_qa, void* workspace,
                             size_t workspace_bytes, float dropout_p, uint64_t seed, uint64_t offset, int mode,
                             void* stream) {
  const LnArgs ln{ln_gamma, nullptr, d_ln_gamma, d_ln_beta};
  return news_encoder_bwd_impl(d_out, tokens, n_titles, L, num_words, wqkv, wa, qa, stash, d_emb, d_wqkv, d_bqkv, d_wa,
                               d_ba, d_qa, workspace, workspace_bytes, dropout_p, seed, offset, mode, stream, &ln);
}

int nrms_news_encoder_rows_fwd(const int64_t* token_table, int64_t n_news, const int64_t* news_rows, int64_t n_titles,
                               int L, const float* emb, int64_t num_words, const float* wqkv, const float* bqkv,
                               const float* ln_gamma, const float* ln_beta, const float* wa, const float* ba,
                               const float* qa, float* out, void* stash, float dropout_p, uint64_t seed, uint64_t offset,
                               int mode, void* stream) {
  NRMS_CHECK_ARG(n_news > 0 && (n_titles == 0 || news_rows), NRMS_E_INVALID, "bad token table / news rows");
  NRMS_CHECK_ARG((ln_gamma == nullptr) == (ln_beta == nullptr), NRMS_E_INVALID, "LayerNorm needs both weight and bias");
  const LnArgs ln{ln_gamma, ln_beta, nullptr, nullptr};
  return news_encoder_fwd_impl(token_table, n_titles, L, emb, num_words, wqkv, bqkv, wa, ba, qa, out, stash, nullptr, 0,
                               dropout_p, seed, offset, mode, stream, ln_gamma ? &ln : nullptr, news_rows);
}

int nrms_news_encoder_rows_bwd(const float* d_out, const int64_t* token_table, int64_t n_news, const int64_t* news_rows,
                               int64_t n_titles, int L, int64_t num_words, const float* wqkv, const float* ln_gamma,
                               const float* wa, const float* qa, const void* stash, float* d_emb, float* d_wqkv,
                               float* d_bqkv, float* d_ln_gamma, float* d_ln_beta, float* d_wa, float* d_ba, float* d_qa,
                               void* workspace, size_t workspace_bytes, float dropout_p, uint64_t seed, uint64_t offset,
                               int mode, void* stream) {
  NRMS_CHECK_ARG(n_news > 0 && (n_titles == 0 || news_rows), NRMS_E_INVALID, "bad token table / news rows");
  const LnArgs ln{ln_gamma, nullptr, d_ln_gamma, d_ln_beta};
  return news_encoder_bwd_impl(d_out, token_table, n_titles, L, num_words, wqkv, wa, qa, stash, d_emb, d_wqkv, d_bqkv,
                               d_wa, d_ba, d_qa, workspace, workspace_bytes, dropout_p, seed, offset, mode, stream,
                               ln_gamma ? &ln : nullptr, news_rows);
}

static int user_encoder_fwd_impl(const float* x, int64_t n_rows, const int32_t* rows_idx, int64_t n_users, int S,
                                 const float* wqkv, const float* bqkv, const float* wa, const float* ba, const float* qa,
                                 float* out, void* stash, void* workspace, size_t workspace_bytes, int mode,
                                 void* stream, const LnArgs* ln) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (int rc = check_ln(ln, false)) return rc;
  if (int rc = check_common(S, mode)) return rc;
  NRMS_CHECK_ARG(n_users >= 0, NRMS_E_INVALID, "bad sizes");
  if (n_users == 0) return NRMS_OK;
  NRMS_CHECK_ARG(x && wqkv && bqkv && wa && ba && qa && out, NRMS_E_INVALID, "null pointer");
  NRMS_CHECK_ARG(aligned16(x) && aligned16(wqkv) && aligned16(bqkv) && aligned16(wa) && aligned16(ba) && aligned16(out),
                 NRMS_E_INVALID, "pointers must be 16-byte aligned");
  if (stash) {
    NRMS_CHECK_ARG(rows_idx == nullptr, NRMS_E_UNSUPPORTED, "indexed input is inference-only");
    NRMS_CHECK_ARG(aligned16(stash), NRMS_E_INVALID, "stash misaligned");
    Stash s = carve_stash(stash, n_users * S, ln != nullptr);
    NRMS_CUDA(cudaMemcpyAsync(s.x, x, (size_t)n_users * S * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return encoder_core_fwd(s, n_users, S, wqkv, bqkv, wa, ba, qa, out, 0.f, 0, 0, 0, mode, st, ln);
  }
  NRMS_CHECK_ARG(rows_idx == nullptr || n_rows > 0, NRMS_E_INVALID, "indexed input needs n_rows (rows of the table)");
  if (mode == NRMS_MODE_TF32 && tc_fused_workspace_bytes(n_users, S, rows_idx ? n_rows : 0) != (size_t)-1) {
    return tc_encoder_fused(x, nullptr, rows_idx ? n_rows : 0, rows_idx, rows_idx ? 2 : 0, n_users, S, wqkv, bqkv, wa, ba, qa, out,
                            workspace, workspace_bytes, st, ln ? ln->gamma : nullptr, ln ? ln->beta : nullptr);
  }
  const int64_t chunk_seq = INFER_CHUNK_ROWS / S;
  const int64_t first = n_users < chunk_seq ? n_users : chunk_seq;
  NRMS_CHECK_ARG(workspace && aligned16(workspace) && workspace_bytes >= carve_stash(nullptr, first * S, true).bytes,
                 NRMS_E_WORKSPACE, "workspace too small: need %zu bytes", carve_stash(nullptr, first * S, true).bytes);
  for (int64_t s0 = 0; s0 < n_users; s0 += chunk_seq) {
    const int64_t n = (n_users - s0 < chunk_seq) ? (n_users - s0) : chunk_seq;
    Stash s = carve_stash(workspace, n * S, ln != nullptr);
    const int64_t rows = n * S;
    if (rows_idx) {
      int64_t gb = (rows + 7) / 8;
      if (gb > (int64_t)num_sms() * 16) gb = (int64_t)num_sms() * 16;
      gather_rows_kernel<int32_t><<<(unsigned)gb, 256, 0, st>>>(x, rows_idx + s0 * S, rows, DV4, s.x);
      NRMS_LAUNCH_CHECK("gather_rows");
    } else {
      s.x = const_cast<float*>(x) + s0 * S * D;  // dense input is read in place
    }
    int rc = encoder_core_fwd(s, n, S, wqkv, bqkv, wa, ba, qa, out + s0 * D, 0.f, 0, 0, 0, mode, st, ln);
    if (rc) return rc;
  }
  return NRMS_OK;
}

int nrms_user_encoder_fwd(const float* x, int64_t n_rows, const int32_t* rows_idx, int64_t n_users, int S, const float* wqkv,
                          const float* bqkv, const float* wa, const float* ba, const float* qa, float* out,
                          void* stash, void* workspace, size_t workspace_bytes, int mode, void* stream) {
  return user_encoder_fwd_impl(x, n_rows, rows_idx, n_users, S, wqkv, bqkv, wa, ba, qa, out, stash, workspace,
                               workspace_bytes, mode, stream, nullptr);
}

int nrms_user_encoder_ln_fwd(const float* x, int64_t n_rows, const int32_t* rows_idx, int64_t n_users, int S,
                             const float* wqkv, const float* bqkv, const float* ln_gamma, const float* ln_beta,
                             const float* wa, const float* ba, const float* qa, float* out, void* stash, void* workspace,
                             size_t workspace_bytes, int mode, void* stream) {
  const LnArgs ln{ln_gamma, ln_beta, nullptr, nullptr};
  return user_encoder_fwd_impl(x, n_rows, rows_idx, n_users, S, wqkv, bqkv, wa, ba, qa, out, stash, workspace,
                               workspace_bytes, mode, stream, &ln);
}

size_t nrms_user_encoder_table16_workspace_bytes(int64_t n_users, int S, int64_t n_rows) {
  if (n_users <= 0 || S <= 0 || n_rows <= 0) return 0;
  size_t fused = tc_fused_workspace_bytes(n_users, S, n_rows, true);
  return fused == (size_t)-1 ? 0 : fused + 256;
}

int nrms_user_encoder_table16_fwd(const void* table16, int64_t n_rows, const int32_t* rows_idx, int64_t n_users, int S,
                                  const float* wqkv, const float* bqkv, const float* wa, const float* ba, const float* qa,
                                  float* out, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (int rc = check_common(S, NRMS_MODE_TF32)) return rc;
  NRMS_CHECK_ARG(n_users >= 0 && n_rows > 0, NRMS_E_INVALID, "bad sizes");
  if (n_users == 0) return NRMS_OK;
  NRMS_CHECK_ARG(table16 && rows_idx && wqkv && bqkv && wa && ba && qa && out, NRMS_E_INVALID, "null pointer");
  NRMS_CHECK_ARG(aligned16(table16) && aligned16(wqkv) && aligned16(bqkv) && aligned16(wa) && aligned16(ba) && aligned16(out),
                 NRMS_E_INVALID, "pointers must be 16-byte aligned");
  NRMS_CHECK_ARG(tc_fused_workspace_bytes(n_users, S, n_rows, true) != (size_t)-1, NRMS_E_UNSUPPORTED,
                 "sequence length not compiled");
  return tc_encoder_fused(nullptr, table16, n_rows, rows_idx, 2, n_users, S, wqkv, bqkv, wa, ba, qa, out, workspace,
                          workspace_bytes, st);
}

static int user_encoder_bwd_impl(const float* d_out, int64_t n_users, int S, const float* wqkv, const float* wa,
                                 const float* qa, const void* stash, float* d_x, float* d_wqkv, float* d_bqkv,
                                 float* d_wa, float* d_ba, float* d_qa, void* workspace, size_t workspace_bytes,
                                 int mode, void* stream, const LnArgs* ln) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (int rc = check_ln(ln, true)) return rc;
  if (int rc = check_common(S, mode)) return rc;
  if (n_users == 0) return NRMS_OK;
  NRMS_CHECK_ARG(n_users > 0, NRMS_E_INVALID, "bad sizes");
  NRMS_CHECK_ARG(d_out && wqkv && wa && qa && stash && d_x && d_wqkv && d_bqkv && d_wa && d_ba && d_qa, NRMS_E_INVALID,
                 "null pointer");
  NRMS_CHECK_ARG(aligned16(d_out) && aligned16(stash) && aligned16(d_x) && aligned16(d_wqkv) && aligned16(d_wa),
                 NRMS_E_INVALID, "pointers must be 16-byte aligned");
  const int64_t rows = n_users * S;
  NRMS_CHECK_ARG(workspace && aligned16(workspace) && workspace_bytes >= carve_bwd(nullptr, rows, false, mode).bytes,
                 NRMS_E_WORKSPACE, "workspace too small: need %zu bytes", carve_bwd(nullptr, rows, false, mode).bytes);
  Stash s = carve_stash(const_cast<void*>(stash), rows, ln != nullptr);
  BwdWs w = carve_bwd(workspace, rows, false, mode);
  return encoder_core_bwd(s, w, d_out, n_users, S, wqkv, wa, qa, d_x, d_wqkv, d_bqkv, d_wa, d_ba, d_qa, 0.f, 0, 0, mode,
                          st, ln);
}

int nrms_user_encoder_bwd(const float* d_out, int64_t n_users, int S, const float* wqkv, const float* wa,
                          const float* qa, const void* stash, float* d_x, float* d_wqkv, float* d_bqkv, float* d_wa,
                          float* d_ba, float* d_qa, void* workspace, size_t workspace_bytes, int mode, void* stream) {
  return user_encoder_bwd_impl(d_out, n_users, S, wqkv, wa, qa, stash, d_x, d_wqkv, d_bqkv, d_wa, d_ba, d_qa, workspace,
                               workspace_bytes, mode, stream, nullptr);
}

int nrms_user_encoder_ln_bwd(const float* d_out, int64_t n_users, int S, const float* wqkv, const float* ln_gamma,
                             const float* wa, const float* qa, const void* stash, float* d_x, float* d_wqkv,
                             float* d_bqkv, float* d_ln_gamma, float* d_ln_beta, float* d_wa, float* d_ba, float* d_qa,
                             void* workspace, size_t workspace_bytes, int mode, void* stream) {
  const LnArgs ln{ln_gamma, nullptr, d_ln_gamma, d_ln_beta};
  return user_encoder_bwd_impl(d_out, n_users, S, wqkv, wa, qa, stash, d_x, d_wqkv, d_bqkv, d_wa, d_ba, d_qa, workspace,
                               workspace_bytes, mode, stream, &ln);
}

// ---- standalone L0 blocks (inference): MultiHeadSelfAttention.forward / AdditiveAttention.forward ----
static int mhsa_fwd_impl(const float* x, const int32_t* lengths, int64_t n_seq, int S, const float* wqkv,
                         const float* bqkv, float* ctx, void* workspace, size_t workspace_bytes, int mode, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (int rc = check_common(S, mode)) return rc;
  NRMS_CHECK_ARG(n_seq >= 0, NRMS_E_INVALID, "bad sizes");
  if (n_seq == 0) return NRMS_OK;
  NRMS_CHECK_ARG(x && wqkv && bqkv && ctx, NRMS_E_INVALID, "null pointer");
  NRMS_CHECK_ARG(aligned16(x) && aligned16(wqkv) && aligned16(bqkv) && aligned16(ctx), NRMS_E_INVALID,
                 "pointers must be 16-byte aligned");
  const int64_t rows = n_seq * S;
  const size_t need = (size_t)rows * D3 * sizeof(float);
  NRMS_CHECK_ARG(workspace && aligned16(workspace) && workspace_bytes >= need, NRMS_E_WORKSPACE,
                 "workspace too small: need %zu bytes", need);
  float* qkv = reinterpret_cast<float*>(workspace);
  if (int rc = gemm_nt_bias(x, D, wqkv, D, bqkv, qkv, D3, rows, D3, D, mode, st)) return rc;
  cudaError_t e = (S == 20) ? launch_attention_fwd<20, 15>(qkv, ctx, n_seq, 0.f, 0, 0, st, lengths)
                            : launch_attention_fwd<50, 5>(qkv, ctx, n_seq, 0.f, 0, 0, st, lengths);
  if (e != cudaSuccess) return cuda_fail(e, "attention_fwd");
  return NRMS_OK;
}

int nrms_mhsa_fwd(const float* x, int64_t n_seq, int S, const float* wqkv, const float* bqkv, float* ctx,
                  void* workspace, size_t workspace_bytes, int mode, void* stream) {
  return mhsa_fwd_impl(x, nullptr, n_seq, S, wqkv, bqkv, ctx, workspace, workspace_bytes, mode, stream);
}

int nrms_mhsa_masked_fwd(const float* x, const int32_t* lengths, int64_t n_seq, int S, const float* wqkv,
                         const float* bqkv, float* ctx, void* workspace, size_t workspace_bytes, int mode, void* stream) {
  NRMS_CHECK_ARG(lengths != nullptr || n_seq == 0, NRMS_E_INVALID, "null lengths");
  return mhsa_fwd_impl(x, lengths, n_seq, S, wqkv, bqkv, ctx, workspace, workspace_bytes, mode, stream);
}

int nrms_additive_fwd(const float* c, int64_t n_seq, int S, const float* wa, const float* ba, const float* qa,
                      float* out, void* workspace, size_t workspace_bytes, int mode, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  NRMS_CHECK_ARG(additive_len_ok(S), NRMS_E_UNSUPPORTED, "candidate size %d unsupported (compiled: 2, 3, 4, 20, 50)", S);
  NRMS_CHECK_ARG(mode == NRMS_MODE_FP32 || mode == NRMS_MODE_TF32, NRMS_E_INVALID, "bad mode %d", mode);
  NRMS_CHECK_ARG(n_seq >= 0, NRMS_E_INVALID, "bad sizes");
  if (n_seq == 0) return NRMS_OK;
  NRMS_CHECK_ARG(c && wa && ba && qa && out, NRMS_E_INVALID, "null pointer");
  NRMS_CHECK_ARG(aligned16(c) && aligned16(wa) && aligned16(ba) && aligned16(out), NRMS_E_INVALID,
                 "pointers must be 16-byte aligned");
  const int64_t rows = n_seq * S;
  const size_t t_bytes = align_up((size_t)rows * QD * sizeof(float), 256);
  const size_t need = t_bytes + (size_t)rows * sizeof(float);
  NRMS_CHECK_ARG(workspace && aligned16(workspace) && workspace_bytes >= need, NRMS_E_WORKSPACE,
                 "workspace too small: need %zu bytes", need);
  float* t = reinterpret_cast<float*>(workspace);
  float* w = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + t_bytes);
  // the 2..4-row form (Exp1's final attention) is 3 % of a text encoder's work: always on the CUDA cores
  if (int rc = gemm_nt_bias(c, D, wa, D, ba, t, QD, rows, QD, D, S < 20 ? NRMS_MODE_FP32 : mode, st)) return rc;
  if (cudaError_t e = launch_additive_fwd(c, t, qa, w, out, n_seq, S, st)) return cuda_fail(e, "additive_fwd");
  return NRMS_OK;
}

size_t nrms_additive_bwd_workspace_bytes(int64_t n_seq, int S, int mode) {
  const int64_t rows = n_seq * S;
  const int64_t ldr = (rows + 3) / 4 * 4;
  size_t b = align_up((size_t)rows * QD * sizeof(float), 256) + align_up((size_t)REDUCE_BLOCKS * D3 * sizeof(float), 256);
  if (mode == NRMS_MODE_TF32 && S >= 20)
    b += align_up((size_t)ldr * QD * sizeof(float), 256) + align_up((size_t)ldr * D * sizeof(float), 256) +
         align_up((size_t)D * QD * sizeof(float), 256);
  return b;
}

int nrms_additive_bwd(const float* d_out, const float* c, int64_t n_seq, int S, const float* wa, const float* qa,
                      const void* fwd_workspace, float* d_c, float* d_wa, float* d_ba, float* d_qa, void* workspace,
                      size_t workspace_bytes, int mode, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  NRMS_CHECK_ARG(additive_len_ok(S), NRMS_E_UNSUPPORTED, "candidate size %d unsupported (compiled: 2, 3, 4, 20, 50)", S);
  NRMS_CHECK_ARG(mode == NRMS_MODE_FP32 || mode == NRMS_MODE_TF32, NRMS_E_INVALID, "bad mode %d", mode);
  NRMS_CHECK_ARG(n_seq >= 0, NRMS_E_INVALID, "bad sizes");
  if (n_seq == 0) return NRMS_OK;
  NRMS_CHECK_ARG(d_out && c && wa && qa && fwd_workspace && d_c && d_wa && d_ba && d_qa, NRMS_E_INVALID, "null pointer");
  NRMS_CHECK_ARG(aligned16(d_out) && aligned16(c) && aligned16(wa) && aligned16(d_c) && aligned16(d_wa) &&
                     aligned16(fwd_workspace), NRMS_E_INVALID, "pointers must be 16-byte aligned");
  const size_t need = nrms_additive_bwd_workspace_bytes(n_seq, S, mode);
  NRMS_CHECK_ARG(workspace && aligned16(workspace) && workspace_bytes >= need, NRMS_E_WORKSPACE,
                 "workspace too small: need %zu bytes", need);
  const int64_t rows = n_seq * S;
  const bool tc = (mode == NRMS_MODE_TF32 && S >= 20);
  const float* t = reinterpret_cast<const float*>(fwd_workspace);
  const float* wv = reinterpret_cast<const float*>(reinterpret_cast<const char*>(fwd_workspace) +
                                                   align_up((size_t)rows * QD * sizeof(float), 256));
  AddBwdWs aw{};
  char* p = reinterpret_cast<char*>(workspace);
  size_t off = 0;
  auto take = [&](size_t nfloat) {
    float* r = reinterpret_cast<float*>(p + off);
    off += align_up(nfloat * sizeof(float), 256);
    return r;
  };
  aw.d_u = take((size_t)rows * QD);
  aw.partial = take((size_t)REDUCE_BLOCKS * D3);
  aw.ldr = (rows + 3) / 4 * 4;
  if (tc) {
    aw.t1 = take((size_t)aw.ldr * QD);
    aw.t2 = take((size_t)aw.ldr * D);
    aw.wt = take((size_t)D * QD);
  }
  return additive_block_bwd(d_out, c, t, wv, wa, qa, d_c, aw, n_seq, S, d_wa, d_ba, d_qa, tc, st);
}

int nrms_set_option(const char* key, int value) {
  NRMS_CHECK_ARG(key != nullptr, NRMS_E_INVALID, "null option key");
  if (strcmp(key, "table_ratio") == 0) {
    NRMS_CHECK_ARG(set_table_ratio(value) == NRMS_OK, NRMS_E_INVALID, "table_ratio must be 1..1024");
    return NRMS_OK;
  }
  if (strcmp(key, "news_table_attn") == 0) {
    set_news_table_attn(value != 0);
    return NRMS_OK;
  }
  if (strcmp(key, "user_table_attn") == 0) {
    set_table_attn(value != 0);
    return NRMS_OK;
  }
  if (strcmp(key, "fused_pool") == 0) {
    set_fused_pool(value != 0);
    return NRMS_OK;
  }
  if (strcmp(key, "gemm_tma_epilogue") == 0) {
    set_gemm_tma_epilogue(value != 0);
    return NRMS_OK;
  }
  if (strcmp(key, "k1f_debug") == 0) {
    set_k1f_debug(value);
    return NRMS_OK;
  }
  if (strcmp(key, "attn_safe_softmax") == 0) {
    set_attn_safe_softmax(value);
    return NRMS_OK;
  }
  if (strcmp(key, "train_attn_mma") == 0) {
    g_train_attn_mma = value != 0;
    return NRMS_OK;
  }
  if (strcmp(key, "time_k1") == 0) {
    set_time_k1(value != 0);
    return NRMS_OK;
  }
  set_error("unknown option '%s'", key);
  return NRMS_E_INVALID;
}

double nrms_get_stat(const char* key) {
  if (!key) return -1.0;
  // "<kind>_ms" / "<kind>_launches" / "<kind>_sequences"; kind: k1 = user-encoder K1 (per-user projection),
  // k1n = news-encoder K1, k1g = user-encoder table attention
  static const char* kinds[4] = {"k1_", "k1n_", "k1g_", "k1gn_"};
  static const char* whats[3] = {"ms", "launches", "sequences"};
  for (int k = 3; k >= 0; --k) {          // longest prefix first (k1gn_ before k1g_, k1n_ before k1_)
    const size_t n = strlen(kinds[k]);
    if (strncmp(key, kinds[k], n) == 0)
      for (int w = 0; w < 3; ++w)
        if (strcmp(key + n, whats[w]) == 0) return get_k1_stat(3 * k + w);
  }
  return -1.0;
}

int nrms_gemm_nt(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, float* C, int64_t ldc,
                 int64_t M, int N, int K, int mode, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  NRMS_CHECK_ARG(A && B && C, NRMS_E_INVALID, "null pointer");
  NRMS_CHECK_ARG(M >= 0 && N > 0 && K > 0 && (N % 4) == 0 && (K % 4) == 0 && (lda % 4) == 0 && (ldb % 4) == 0 && (ldc % 4) == 0,
                 NRMS_E_UNSUPPORTED, "N, K and leading dimensions must be multiples of 4");
  NRMS_CHECK_ARG(aligned16(A) && aligned16(B) && aligned16(C) && (!bias || aligned16(bias)), NRMS_E_INVALID,
                 "pointers must be 16-byte aligned");
  return gemm_nt_bias(A, lda, B, ldb, bias, C, ldc, M, N, K, mode, st);
}

}  // extern "C"
