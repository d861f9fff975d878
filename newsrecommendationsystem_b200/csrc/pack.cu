// Operand packing for the tensor-mode inference kernels: fp16 copies of the projection weights (K1 v6's pass-block
// layout) and of gather-source rows (embedding rows / news vectors), each written once per encoder call.
#include <cuda.h>
#include <cuda_fp16.h>
#include "tc_common.cuh"
#include "tc_api.cuh"

namespace nrms {

int make_tmap_k_major_f16(CUtensorMap* out, const void* base, int64_t rows, int cols, int64_t ld, int box_rows);

namespace pack {

constexpr int W16_ROWS = 1024, W16_LD = 320;    // 8 pass blocks of 128 rows; 300 weights | bias | zero tail
constexpr int PN = 128;                         // projection UMMA N of K1 v6 (120 real columns per pass)
constexpr int SRC_LD = 320;                     // pitch (halfs) of the fp16 gather source
// log2(e)/sqrt(20), times (1 + 2^-11): K1 v6's S MMA reads q as tf32 by TRUNCATING the fp32 accumulator (mean relative
// error -2^-11); the pre-scale centres that error like a round-to-nearest would.
constexpr float QSCALE = 1.4426950408889634f / 4.47213595499957939f * (1.f + 1.f / 2048.f);

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
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

}  // namespace pack

size_t src16_bytes(int64_t n_rows) { return (size_t)(n_rows + 1) * pack::SRC_LD * 2; }

// fp16 weight copy of K1 v6 (once per encoder call) + its tensor map
int pack_weights_k1(const float* wqkv, const float* bqkv, void* w16, CUtensorMap* tw, cudaStream_t st) {
  pack::pack_w16_kernel<<<148, 256, 0, st>>>(wqkv, bqkv, reinterpret_cast<__half*>(w16));
  NRMS_LAUNCH_CHECK("pack_w16_kernel");
  return make_tmap_k_major_f16(tw, w16, pack::W16_ROWS, pack::W16_LD, pack::W16_LD, pack::PN);
}

// fp16 copy [n_rows + 1][320] of n_rows fp32 rows (1.0 in column 300, zero tail, all-zero last row)
int pack_rows16(const float* src, int64_t n_rows, void* src16, cudaStream_t st) {
  NRMS_CHECK_ARG((n_rows + 1) * pack::SRC_LD * 2 < (1ll << 32), NRMS_E_UNSUPPORTED,
                 "gather source too large (32-bit byte offsets in the K1 gather)");
  const int64_t total = (n_rows + 1) * 40;
  int64_t gb = (total + 255) / 256;
  if (gb > (int64_t)num_sms() * 16) gb = (int64_t)num_sms() * 16;
  pack::pack_src16_kernel<<<(unsigned)gb, 256, 0, st>>>(src, n_rows, reinterpret_cast<__half*>(src16));
  NRMS_LAUNCH_CHECK("pack_src16_kernel");
  return NRMS_OK;
}

}  // namespace nrms
