// Single-user latency path: what reference src/recommend.py:245-341 does for its one target user --
//   user_vector = model.get_user_vector(stack(news2vector[x] for x in clicked_news))      recommend.py:264-279
//   click_probability = model.get_prediction(stack(news2vector[c] for c in impression), user_vector)   :301-315
//   order = np.argsort(-(click_probability + 1) / 2)                                                    :338-340
// -- as TWO launches of thread-block clusters instead of ~25 batched-op launches over a batch of one.  FP32 on the
// CUDA cores throughout (reference arithmetic up to summation order): at one user the whole encoder is 16.5 M MACs
// and what matters is how many SMs share it and how few times they meet.
//
// Kernel A (15 clusters x 8 CTAs, one cluster per attention head): CTA r of cluster h projects history rows r, r+8, ..
// onto the head's 60 q|k|v columns, hands its k|v rows to the other seven CTAs through distributed shared memory, runs
// exp-softmax attention for its own rows, and writes (a) the head's slice of its context rows and (b) that slice's
// contribution to the additive layer, tpart[h][i][:] = C_h[i] Wa[:, 20h:20h+20]^T -- the additive linear map is a sum
// over heads, so no CTA ever needs a whole context row.
// Kernel B (one cluster of 8 CTAs): CTA r sums the 15 partials for 25 of the 200 additive units, adds the bias, takes
// tanh and the dot with the query; the eight partial scores per row meet in every CTA's shared memory; each CTA then
// knows the softmax weights, forms the user vector, scores every 8th candidate, and CTA 0 sorts.
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace nrms {
namespace lat {

constexpr int S = 50;           // config.num_clicked_news_a_user
constexpr int CL = 8;           // CTAs per cluster (portable maximum)
constexpr int RMAX = 7;         // history rows per CTA: ceil(50 / 8)
constexpr int XS = 308;         // padded row stride (floats): 16-byte aligned, conflict-free for strided float4 reads
constexpr int THREADS = 256;
constexpr float SQRT_DH = 4.47213595499957939f;
constexpr int MAX_SORT = 4096;

struct SmemA {
  float x[RMAX][XS];            // this CTA's history rows
  float w[3 * DH][XS];          // the head's q|k|v weight rows: j = 20 * {q,k,v} + column
  float bias[64];
  float q[RMAX][DH];
  float kv[S][2 * DH];          // k (0..19) | v (20..39) of all 50 rows; every CTA of the cluster writes its rows here
  float e[RMAX][S + 2];
  float z[RMAX + 1];
  float ctx[RMAX][DH];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS)
head_kernel(const float* __restrict__ table, int64_t n_rows, const int32_t* __restrict__ hist,
            const float* __restrict__ wqkv, const float* __restrict__ bqkv, const float* __restrict__ wa,
            float* __restrict__ ctx_g, float* __restrict__ tpart) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemA& sm = *reinterpret_cast<SmemA*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int r = (int)cluster.block_rank();
  const int h = blockIdx.x / CL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_mine = (S - r + CL - 1) / CL;      // rows r, r + 8, ... < 50

  // ---- stage the head's weight rows and this CTA's history rows (16-byte async copies) ----
  for (int f = tid; f < 3 * DH * DV4; f += THREADS) {
    const int j = f / DV4, l = f % DV4;
    const int64_t row = (int64_t)(j / DH) * D + h * DH + (j % DH);
    cp_async16(&sm.w[j][4 * l], wqkv + row * D + 4 * l);
  }
  for (int f = tid; f < RMAX * DV4; f += THREADS) {
    const int ii = f / DV4, l = f % DV4;
    if (ii < n_mine) {
      int64_t src = hist[r + CL * ii];
      if (src < 0 || src >= n_rows) src = n_rows - 1;     // callers pass the zero PADDED_NEWS row last; see the header
      cp_async16(&sm.x[ii][4 * l], table + src * D + 4 * l);
    } else {
      *reinterpret_cast<float4*>(&sm.x[ii][4 * l]) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (tid < 3 * DH) sm.bias[tid] = bqkv[(tid / DH) * D + h * DH + (tid % DH)];
  // the slice of W_a this thread needs at the end (unit q = tid, columns 20h .. 20h+19), fetched early
  float4 wa4[DH / 4];
  if (tid < QD) {
#pragma unroll
    for (int k = 0; k < DH / 4; ++k) wa4[k] = __ldg(reinterpret_cast<const float4*>(wa + (int64_t)tid * D + h * DH) + k);
  }
  cp_async_wait_all();
  __syncthreads();

  // ---- projection: warp 0 / 1 own columns 0..31 / 32..59, 7 rows each (x reads are warp broadcasts) ----
  if (warp < 2) {
    const int j = warp * 32 + lane;
    const int jc = j < 3 * DH ? j : 3 * DH - 1;
    float acc[RMAX];
#pragma unroll
    for (int ii = 0; ii < RMAX; ++ii) acc[ii] = sm.bias[jc];
    for (int l = 0; l < DV4; ++l) {
      const float4 wv = *reinterpret_cast<const float4*>(&sm.w[jc][4 * l]);
#pragma unroll
      for (int ii = 0; ii < RMAX; ++ii) {
        const float4 xv = *reinterpret_cast<const float4*>(&sm.x[ii][4 * l]);
        acc[ii] = fmaf(xv.x, wv.x, acc[ii]);
        acc[ii] = fmaf(xv.y, wv.y, acc[ii]);
        acc[ii] = fmaf(xv.z, wv.z, acc[ii]);
        acc[ii] = fmaf(xv.w, wv.w, acc[ii]);
      }
    }
    if (j < DH) {
#pragma unroll
      for (int ii = 0; ii < RMAX; ++ii) sm.q[ii][j] = acc[ii];
    } else if (j < 3 * DH) {
      // k | v of my rows go to every CTA of the cluster (distributed shared memory)
      for (int dst = 0; dst < CL; ++dst) {
        float* remote = cluster.map_shared_rank(&sm.kv[0][0], dst);
#pragma unroll
        for (int ii = 0; ii < RMAX; ++ii)
          if (ii < n_mine) remote[(r + CL * ii) * (2 * DH) + (j - DH)] = acc[ii];
      }
    }
  }
  cluster.sync();

  // ---- attention of my rows over all 50 keys (multihead_self.py:15-23: exp, no shift, eps in the denominator) ----
  for (int f = tid; f < n_mine * S; f += THREADS) {
    const int ii = f / S, j = f % S;
    float dot = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) dot = fmaf(sm.q[ii][d], sm.kv[j][d], dot);
    sm.e[ii][j] = expf(dot / SQRT_DH);
  }
  __syncthreads();
  if (warp < n_mine) {
    float zs = 0.f;
    for (int j = lane; j < S; j += 32) zs += sm.e[warp][j];
    zs = warp_sum(zs);
    if (lane == 0) sm.z[warp] = zs + 1e-8f;
  }
  __syncthreads();
  if (tid < n_mine * DH) {
    const int ii = tid / DH, d = tid % DH;
    const float zinv_den = sm.z[ii];
    float acc = 0.f;
    for (int j = 0; j < S; ++j) acc = fmaf(sm.e[ii][j] / zinv_den, sm.kv[j][DH + d], acc);
    sm.ctx[ii][d] = acc;
    ctx_g[(int64_t)(r + CL * ii) * D + h * DH + d] = acc;
  }
  __syncthreads();

  // ---- this head's share of the additive layer for my rows ----
  if (tid < QD) {
    for (int ii = 0; ii < n_mine; ++ii) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < DH / 4; ++k) {
        acc = fmaf(sm.ctx[ii][4 * k + 0], wa4[k].x, acc);
        acc = fmaf(sm.ctx[ii][4 * k + 1], wa4[k].y, acc);
        acc = fmaf(sm.ctx[ii][4 * k + 2], wa4[k].z, acc);
        acc = fmaf(sm.ctx[ii][4 * k + 3], wa4[k].w, acc);
      }
      tpart[((int64_t)h * S + (r + CL * ii)) * QD + tid] = acc;
    }
  }
  // no CTA may exit while a peer can still write into its shared memory: all remote stores precede the first
  // cluster.sync(), and none follow it
}

__device__ __forceinline__ uint32_t float_desc_key(float v) {
  uint32_t u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);    // ascending order of u == ascending order of v
  return ~u;                                          // ascending order of the key == descending score
}

struct SmemB {
  float sp[CL][S];              // partial pooling scores of the 8 unit slices (filled by all CTAs)
  float srow[S];
  float wrow[S];
  float user[D + 4];
};

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS)
pool_score_kernel(const float* __restrict__ table, int64_t n_rows, const int32_t* __restrict__ cand, int C,
                  const float* __restrict__ ba, const float* __restrict__ qa, const float* __restrict__ ctx_g,
                  const float* __restrict__ tpart, float* __restrict__ user_vec, float* __restrict__ scores,
                  int32_t* __restrict__ order, int n_sort) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemB& sm = *reinterpret_cast<SmemB*>(smem_raw);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw + sizeof(SmemB));
  cg::cluster_group cluster = cg::this_cluster();
  const int r = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int QS = QD / CL;   // 25 additive units per CTA

  if (tid < S) sm.srow[tid] = 0.f;
  __syncthreads();
  for (int f = tid; f < S * QS; f += THREADS) {
    const int i = f / QS, q = r * QS + f % QS;
    float v[H];
#pragma unroll
    for (int hh = 0; hh < H; ++hh) v[hh] = __ldg(tpart + ((int64_t)hh * S + i) * QD + q);
    float t = ba[q];
#pragma unroll
    for (int hh = 0; hh < H; ++hh) t += v[hh];
    atomicAdd(&sm.srow[i], tanhf(t) * qa[q]);
  }
  __syncthreads();
  if (tid < S) {
    const float v = sm.srow[tid];
    for (int dst = 0; dst < CL; ++dst) cluster.map_shared_rank(&sm.sp[0][0], dst)[r * S + tid] = v;
  }
  cluster.sync();

  // every CTA: softmax over the 50 rows (additive.py:37-39), then the user vector
  if (warp == 0) {
    float sv[2];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = lane + 32 * k;
      float a = 0.f;
      if (i < S) {
#pragma unroll
        for (int c = 0; c < CL; ++c) a += sm.sp[c][i];
        m = fmaxf(m, a);
      }
      sv[k] = a;
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = lane + 32 * k;
      sv[k] = (i < S) ? expf(sv[k] - m) : 0.f;
      sum += sv[k];
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = lane + 32 * k;
      if (i < S) sm.wrow[i] = sv[k] / sum;
    }
  }
  __syncthreads();
  for (int d = tid; d < D; d += THREADS) {
    float acc = 0.f;
#pragma unroll 10
    for (int i = 0; i < S; ++i) acc = fmaf(sm.wrow[i], __ldg(ctx_g + (int64_t)i * D + d), acc);
    sm.user[d] = acc;
    if (r == 0) user_vec[d] = acc;
  }
  __syncthreads();

  // scores: warp per candidate, candidates r*8 + warp, + 64, ...; two rows in flight per warp
  {
    float4 u[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int l = lane + 32 * k;
      u[k] = (l < DV4) ? *reinterpret_cast<const float4*>(&sm.user[4 * l]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int c0 = r * 8 + warp; c0 < C; c0 += 2 * CL * 8) {
      const int c1 = c0 + CL * 8;
      int64_t row0 = cand[c0], row1 = c1 < C ? cand[c1] : 0;
      if (row0 < 0 || row0 >= n_rows) row0 = n_rows - 1;
      if (row1 < 0 || row1 >= n_rows) row1 = n_rows - 1;
      const float4* p0 = reinterpret_cast<const float4*>(table + row0 * D);
      const float4* p1 = reinterpret_cast<const float4*>(table + row1 * D);
      float4 a[3], b[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int l = lane + 32 * k;
        a[k] = (l < DV4) ? __ldg(p0 + l) : make_float4(0.f, 0.f, 0.f, 0.f);
        b[k] = (l < DV4) ? __ldg(p1 + l) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        s0 = fmaf(a[k].x, u[k].x, s0); s0 = fmaf(a[k].y, u[k].y, s0); s0 = fmaf(a[k].z, u[k].z, s0); s0 = fmaf(a[k].w, u[k].w, s0);
        s1 = fmaf(b[k].x, u[k].x, s1); s1 = fmaf(b[k].y, u[k].y, s1); s1 = fmaf(b[k].z, u[k].z, s1); s1 = fmaf(b[k].w, u[k].w, s1);
      }
      s0 = warp_sum(s0);
      s1 = warp_sum(s1);
      if (lane == 0) {
        scores[c0] = s0;
        if (c1 < C) scores[c1] = s1;
      }
    }
  }
  if (order == nullptr) return;
  __threadfence();
  cluster.sync();
  if (r != 0) return;

  // CTA 0: order = argsort(-score), equal scores in ascending candidate position (bitonic sort of 64-bit keys)
  for (int i = tid; i < n_sort; i += THREADS)
    keys[i] = i < C ? (((unsigned long long)float_desc_key(__ldcg(scores + i)) << 32) | (unsigned)i) : ~0ull;
  __syncthreads();
  for (int k = 2; k <= n_sort; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n_sort; i += THREADS) {
        const int p = i ^ j;
        if (p > i) {
          const unsigned long long a = keys[i], b = keys[p];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { keys[i] = b; keys[p] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < C; i += THREADS) order[i] = (int32_t)(keys[i] & 0xffffffffu);
}

}  // namespace lat
}  // namespace nrms

using namespace nrms;

extern "C" {

size_t nrms_recommend_workspace_bytes(void) {
  return align_up((size_t)lat::S * D * sizeof(float), 256) + (size_t)H * lat::S * QD * sizeof(float);
}

int nrms_recommend_user(const float* table, int64_t n_rows, const int32_t* hist_rows, const int32_t* cand_rows, int C,
                        const float* wqkv, const float* bqkv, const float* wa, const float* ba, const float* qa,
                        float* user_vec, float* scores, int32_t* order, void* workspace, size_t workspace_bytes,
                        void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  NRMS_CHECK_ARG(n_rows > 0 && C >= 0, NRMS_E_INVALID, "bad sizes");
  NRMS_CHECK_ARG(table && hist_rows && wqkv && bqkv && wa && ba && qa && user_vec, NRMS_E_INVALID, "null pointer");
  NRMS_CHECK_ARG(C == 0 || (cand_rows && scores), NRMS_E_INVALID, "null candidate / score pointer");
  NRMS_CHECK_ARG(aligned16(table) && aligned16(wqkv) && aligned16(wa), NRMS_E_INVALID, "pointers must be 16-byte aligned");
  NRMS_CHECK_ARG(order == nullptr || C <= lat::MAX_SORT, NRMS_E_UNSUPPORTED,
                 "the in-kernel ranking holds at most %d candidates (pass order = NULL for scores only)", lat::MAX_SORT);
  NRMS_CHECK_ARG(workspace && aligned16(workspace) && workspace_bytes >= nrms_recommend_workspace_bytes(),
                 NRMS_E_WORKSPACE, "workspace too small: need %zu bytes", nrms_recommend_workspace_bytes());
  float* ctx_g = reinterpret_cast<float*>(workspace);
  float* tpart = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + align_up((size_t)lat::S * D * sizeof(float), 256));
  static bool cfg_a[64] = {false}, cfg_b[64] = {false};
  if (cudaError_t e = set_max_dynamic_smem(lat::head_kernel, (int)sizeof(lat::SmemA), cfg_a)) return cuda_fail(e, "head_kernel attr");
  lat::head_kernel<<<H * lat::CL, lat::THREADS, sizeof(lat::SmemA), st>>>(table, n_rows, hist_rows, wqkv, bqkv, wa, ctx_g, tpart);
  NRMS_LAUNCH_CHECK("recommend head_kernel");
  int n_sort = 2;
  while (order && n_sort < C) n_sort <<= 1;
  const size_t smem_b = sizeof(lat::SmemB) + (order ? (size_t)n_sort * 8 : 0);
  if (cudaError_t e = set_max_dynamic_smem(lat::pool_score_kernel, (int)(sizeof(lat::SmemB) + lat::MAX_SORT * 8), cfg_b))
    return cuda_fail(e, "pool_score_kernel attr");
  lat::pool_score_kernel<<<lat::CL, lat::THREADS, smem_b, st>>>(table, n_rows, cand_rows, C, ba, qa, ctx_g, tpart, user_vec,
                                                                scores, C > 0 ? order : nullptr, n_sort);
  NRMS_LAUNCH_CHECK("recommend pool_score_kernel");
  return NRMS_OK;
}

}  // extern "C"
