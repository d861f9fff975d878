// Building blocks of model/Exp1 that plain NRMS does not have:
//   ElementEncoder.forward   relu(linear(embedding(element)))          reference src/model/Exp1/news_encoder.py:37-44
//   UserEncoder's position embedding   x + position_embedding          reference src/model/Exp1/user_encoder.py:25-26
//   torch.stack(all_vectors, dim=1) ahead of the final attention       reference src/model/Exp1/news_encoder.py:104-110
// The text encoders and the final AdditiveAttention are the NRMS kernels (encoder.cu: nrms_news_encoder_fwd/bwd at 20
// tokens, nrms_additive_fwd/bwd at 2..4 candidates).
//
// Element encoder: the category vocabulary is tiny (275 rows, src/config.py:29) against the calls that gather from it
// (a news corpus, or 55 x 128 titles per training step), so the TABLE is pushed through the linear layer once per call
//   P[c, :] = relu(W E[c, :] + b)     [num_categories, 300]
// and the per-element work is a 1,200-byte row gather.  Backward: rows of d_out are summed per category (atomics, like
// the word-embedding scatter), masked by P > 0 (the ReLU mask depends on the category only), and the three small
// contractions run over num_categories rows instead of n.
#include "common.cuh"

namespace nrms {

constexpr int CE = NRMS_CE;    // category_embedding_dim (100)

// P[c][o] = relu(b[o] + sum_k W[o][k] E[c][k]); one block per category, one thread per output column
static __global__ void __launch_bounds__(320)
element_table_kernel(const float* __restrict__ emb, const float* __restrict__ w, const float* __restrict__ b,
                     float* __restrict__ P) {
  __shared__ __align__(16) float e[CE];
  const int c = blockIdx.x, o = threadIdx.x;
  if (o < CE) e[o] = emb[(int64_t)c * CE + o];
  __syncthreads();
  if (o >= D) return;
  const float4* wr = reinterpret_cast<const float4*>(w + (int64_t)o * CE);
  float acc = b[o];
#pragma unroll 5
  for (int k4 = 0; k4 < CE / 4; ++k4) {
    const float4 wv = __ldg(wr + k4);
    const float4 ev = *reinterpret_cast<const float4*>(e + 4 * k4);
    acc = fmaf(wv.x, ev.x, acc);
    acc = fmaf(wv.y, ev.y, acc);
    acc = fmaf(wv.z, ev.z, acc);
    acc = fmaf(wv.w, ev.w, acc);
  }
  P[(int64_t)c * D + o] = fmaxf(acc, 0.f);
}

// dst[r * dst_stride + :width] = src[rows ? rows[r] : r][:width]; one warp per row, float4 lanes
static __global__ void __launch_bounds__(256)
copy_rows_kernel(const float* __restrict__ src, int64_t src_stride, const int64_t* __restrict__ rows,
                 float* __restrict__ dst, int64_t dst_stride, int64_t n, int width4) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp; r < n; r += nwarps) {
    const float4* s = reinterpret_cast<const float4*>(src + (rows ? rows[r] : r) * src_stride);
    float4* d = reinterpret_cast<float4*>(dst + r * dst_stride);
    for (int l = lane; l < width4; l += 32) d[l] = __ldg(s + l);
  }
}

// dP[idx[r], :] += d_out[r, :]
static __global__ void __launch_bounds__(256)
scatter_rows_atomic_kernel(const int64_t* __restrict__ idx, int64_t n, const float* __restrict__ d_out,
                           float* __restrict__ dP) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp; r < n; r += nwarps) {
    const float4* s = reinterpret_cast<const float4*>(d_out + r * D);
    float4* d = reinterpret_cast<float4*>(dP + idx[r] * D);
    for (int l = lane; l < DV4; l += 32) atomicAdd(d + l, s[l]);
  }
}

// dZ[c, :] = dP[c, :] where P[c, :] > 0 else 0 (in place over dP); d_emb[c, k] += sum_o dZ[c][o] W[o][k] for c != 0
// (padding_idx = 0 of the category embedding, Exp1/news_encoder.py:75-77, never receives a gradient)
static __global__ void __launch_bounds__(320)
element_bwd_rows_kernel(const float* __restrict__ P, float* __restrict__ dP, const float* __restrict__ w,
                        float* __restrict__ d_emb) {
  __shared__ float dz[D];
  const int c = blockIdx.x, t = threadIdx.x;
  if (t < D) {
    const float v = P[(int64_t)c * D + t] > 0.f ? dP[(int64_t)c * D + t] : 0.f;
    dz[t] = v;
    dP[(int64_t)c * D + t] = v;
  }
  __syncthreads();
  if (c == 0 || t >= CE) return;
  float acc = 0.f;
#pragma unroll 4
  for (int o = 0; o < D; ++o) acc = fmaf(dz[o], __ldg(w + (int64_t)o * CE + t), acc);
  d_emb[(int64_t)c * CE + t] += acc;
}

// d_w[o][k] += sum_c dZ[c][o] E[c][k];  d_b[o] += sum_c dZ[c][o]; one block per output column o
static __global__ void __launch_bounds__(128)
element_bwd_weight_kernel(const float* __restrict__ dZ, const float* __restrict__ emb, int num_categories,
                          float* __restrict__ d_w, float* __restrict__ d_b) {
  const int o = blockIdx.x, k = threadIdx.x;
  float acc = 0.f;
  if (k < CE) {
    for (int c = 0; c < num_categories; ++c) acc = fmaf(__ldg(dZ + (int64_t)c * D + o), emb[(int64_t)c * CE + k], acc);
    d_w[(int64_t)o * CE + k] += acc;
  } else if (k == CE) {
    for (int c = 0; c < num_categories; ++c) acc += dZ[(int64_t)c * D + o];
    d_b[o] += acc;
  }
}

// out[u, i, :] = x[u, i, :] + pos[i, :]
static __global__ void __launch_bounds__(256)
add_position_kernel(const float4* __restrict__ x, const float4* __restrict__ pos, float4* __restrict__ out,
                    int64_t n4, int64_t per_user4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = x[i], b = __ldg(pos + (i % per_user4));
    out[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

// partial[b, col] = sum of x[r, col] over block b's contiguous row range (x [n_rows, N], N % 4 == 0)
static __global__ void __launch_bounds__(256)
wide_colsum_partial_kernel(const float* __restrict__ x, int64_t n_rows, int N, float* __restrict__ partial) {
  const int64_t per = (n_rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per;
  const int64_t r1 = (r0 + per < n_rows) ? r0 + per : n_rows;
  const int64_t ld4 = N / 4;
  for (int c4 = threadIdx.x; c4 < ld4; c4 += blockDim.x) {
    const float4* xp = reinterpret_cast<const float4*>(x) + c4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = r0; r < r1; ++r) {
      const float4 v = __ldg(xp + r * ld4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(partial + (int64_t)blockIdx.x * N)[c4] = acc;
  }
}
// out[col] += sum_b partial[b, col] in a fixed order
static __global__ void __launch_bounds__(256)
wide_colsum_final_kernel(const float* __restrict__ partial, int n_blocks, int N, float* __restrict__ out) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= N) return;
  float acc = 0.f;
  for (int b = 0; b < n_blocks; ++b) acc += partial[(int64_t)b * N + col];
  out[col] += acc;
}

constexpr int POS_BLOCKS = 64;

}  // namespace nrms

using namespace nrms;

extern "C" {

size_t nrms_element_encoder_table_bytes(int64_t num_categories) {
  return align_up((size_t)(num_categories > 0 ? num_categories : 0) * D * sizeof(float), 256);
}

int nrms_element_encoder_fwd(const int64_t* idx, int64_t n, const float* emb, int64_t num_categories, const float* w,
                             const float* b, float* out, void* table, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  NRMS_CHECK_ARG(n >= 0 && num_categories > 0 && num_categories < (1 << 20), NRMS_E_INVALID, "bad sizes");
  NRMS_CHECK_ARG(emb && w && b && table, NRMS_E_INVALID, "null pointer");
  NRMS_CHECK_ARG(aligned16(emb) && aligned16(w) && aligned16(table) && aligned16(out), NRMS_E_INVALID,
                 "pointers must be 16-byte aligned");
  float* P = reinterpret_cast<float*>(table);
  element_table_kernel<<<(unsigned)num_categories, 320, 0, st>>>(emb, w, b, P);
  NRMS_LAUNCH_CHECK("element_table");
  if (n == 0) return NRMS_OK;
  NRMS_CHECK_ARG(idx && out, NRMS_E_INVALID, "null pointer");
  int64_t gx = (n + 7) / 8;
  if (gx > (int64_t)num_sms() * 8) gx = (int64_t)num_sms() * 8;
  copy_rows_kernel<<<(unsigned)gx, 256, 0, st>>>(P, D, idx, out, D, n, DV4);
  NRMS_LAUNCH_CHECK("element_gather");
  return NRMS_OK;
}

int nrms_element_encoder_bwd(const float* d_out, const int64_t* idx, int64_t n, const float* emb, int64_t num_categories,
                             const float* w, const void* table, float* d_emb, float* d_w, float* d_b, void* workspace,
                             size_t workspace_bytes, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  NRMS_CHECK_ARG(n >= 0 && num_categories > 0 && num_categories < (1 << 20), NRMS_E_INVALID, "bad sizes");
  NRMS_CHECK_ARG(emb && w && table && d_emb && d_w && d_b, NRMS_E_INVALID, "null pointer");
  const size_t need = nrms_element_encoder_table_bytes(num_categories);
  NRMS_CHECK_ARG(workspace && aligned16(workspace) && workspace_bytes >= need, NRMS_E_WORKSPACE,
                 "workspace too small: need %zu bytes", need);
  NRMS_CHECK_ARG(aligned16(d_out) && aligned16(table), NRMS_E_INVALID, "pointers must be 16-byte aligned");
  float* dP = reinterpret_cast<float*>(workspace);
  NRMS_CUDA(cudaMemsetAsync(dP, 0, (size_t)num_categories * D * sizeof(float), st));
  if (n > 0) {
    NRMS_CHECK_ARG(d_out && idx, NRMS_E_INVALID, "null pointer");
    int64_t gx = (n + 7) / 8;
    if (gx > (int64_t)num_sms() * 8) gx = (int64_t)num_sms() * 8;
    scatter_rows_atomic_kernel<<<(unsigned)gx, 256, 0, st>>>(idx, n, d_out, dP);
    NRMS_LAUNCH_CHECK("element_scatter");
  }
  element_bwd_rows_kernel<<<(unsigned)num_categories, 320, 0, st>>>(reinterpret_cast<const float*>(table), dP, w, d_emb);
  NRMS_LAUNCH_CHECK("element_bwd_rows");
  element_bwd_weight_kernel<<<D, 128, 0, st>>>(dP, emb, (int)num_categories, d_w, d_b);
  NRMS_LAUNCH_CHECK("element_bwd_weight");
  return NRMS_OK;
}

int nrms_add_position_fwd(const float* x, const float* pos, int64_t n_users, int S, float* out, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  NRMS_CHECK_ARG(n_users >= 0 && S > 0, NRMS_E_INVALID, "bad sizes");
  if (n_users == 0) return NRMS_OK;
  NRMS_CHECK_ARG(x && pos && out, NRMS_E_INVALID, "null pointer");
  NRMS_CHECK_ARG(aligned16(x) && aligned16(pos) && aligned16(out), NRMS_E_INVALID, "pointers must be 16-byte aligned");
  const int64_t per4 = (int64_t)S * DV4, n4 = n_users * per4;
  int64_t gx = (n4 + 255) / 256;
  if (gx > (int64_t)num_sms() * 16) gx = (int64_t)num_sms() * 16;
  add_position_kernel<<<(unsigned)gx, 256, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(pos),
                                                     reinterpret_cast<float4*>(out), n4, per4);
  NRMS_LAUNCH_CHECK("add_position");
  return NRMS_OK;
}

size_t nrms_add_position_bwd_workspace_bytes(int S) { return (size_t)POS_BLOCKS * S * D * sizeof(float); }

int nrms_add_position_bwd(const float* d_out, int64_t n_users, int S, float* d_pos, void* workspace,
                          size_t workspace_bytes, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  NRMS_CHECK_ARG(n_users >= 0 && S > 0, NRMS_E_INVALID, "bad sizes");
  if (n_users == 0) return NRMS_OK;
  NRMS_CHECK_ARG(d_out && d_pos, NRMS_E_INVALID, "null pointer");
  NRMS_CHECK_ARG(aligned16(d_out) && workspace && aligned16(workspace) &&
                     workspace_bytes >= nrms_add_position_bwd_workspace_bytes(S), NRMS_E_WORKSPACE,
                 "workspace too small: need %zu bytes", nrms_add_position_bwd_workspace_bytes(S));
  const int N = S * D;
  const int nb = (int)(n_users < POS_BLOCKS ? n_users : POS_BLOCKS);
  float* partial = reinterpret_cast<float*>(workspace);
  wide_colsum_partial_kernel<<<nb, 256, 0, st>>>(d_out, n_users, N, partial);
  NRMS_LAUNCH_CHECK("position_colsum");
  wide_colsum_final_kernel<<<(N + 255) / 256, 256, 0, st>>>(partial, nb, N, d_pos);
  NRMS_LAUNCH_CHECK("position_colsum_final");
  return NRMS_OK;
}

int nrms_copy_rows_strided(const float* src, int64_t src_stride, float* dst, int64_t dst_stride, int64_t n, int width,
                           void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  NRMS_CHECK_ARG(n >= 0 && width > 0 && (width % 4) == 0 && (src_stride % 4) == 0 && (dst_stride % 4) == 0 &&
                     src_stride >= width && dst_stride >= width, NRMS_E_INVALID, "bad sizes / strides");
  if (n == 0) return NRMS_OK;
  NRMS_CHECK_ARG(src && dst && aligned16(src) && aligned16(dst), NRMS_E_INVALID, "null or misaligned pointer");
  int64_t gx = (n + 7) / 8;
  if (gx > (int64_t)num_sms() * 8) gx = (int64_t)num_sms() * 8;
  copy_rows_kernel<<<(unsigned)gx, 256, 0, st>>>(src, src_stride, nullptr, dst, dst_stride, n, width / 4);
  NRMS_LAUNCH_CHECK("copy_rows_strided");
  return NRMS_OK;
}

}  // extern "C"
