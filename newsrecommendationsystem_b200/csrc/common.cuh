// Shared helpers for libnrms_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/nrms_b200.h"

namespace nrms {

constexpr int D = NRMS_D;      // 300
constexpr int H = NRMS_H;      // 15
constexpr int DH = NRMS_DH;    // 20
constexpr int QD = NRMS_QD;    // 200
constexpr int D3 = 3 * D;      // 900
constexpr int DV4 = D / 4;     // 75 float4 per 300-wide row

// ---- error plumbing (thread-local message, int codes; nothing throws across the ABI) ----
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch();  // bumps the per-process kernel-launch counter (nrms_launch_count)

#define NRMS_CHECK_ARG(cond, code, ...)          \
  do {                                           \
    if (!(cond)) {                               \
      nrms::set_error(__VA_ARGS__);              \
      return (code);                             \
    }                                            \
  } while (0)

#define NRMS_CUDA(expr)                                        \
  do {                                                         \
    cudaError_t e__ = (expr);                                  \
    if (e__ != cudaSuccess) return nrms::cuda_fail(e__, #expr); \
  } while (0)

#define NRMS_LAUNCH_CHECK(name)                                 \
  do {                                                          \
    cudaError_t e__ = cudaGetLastError();                       \
    if (e__ != cudaSuccess) return nrms::cuda_fail(e__, name);  \
    nrms::count_launch();                                       \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
// SM count of the CURRENT device (cached per device: a process may drive several GPUs)
inline int num_sms() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (n[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: `done` is the call site's static flag array,
// one entry per device ordinal.
template <typename F>
inline cudaError_t set_max_dynamic_smem(F func, int bytes, bool (&done)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
  return e;
}

// ---- Philox4x32-10 (counter-based RNG for in-kernel dropout) -------------------------------
// mask element e (linear index in the [rows, width] activation) uses counter (e/4, stream_id,
// offset_lo, offset_hi) keyed by seed; lane e%4 of the 4 outputs.  Forward and backward
// regenerate identical masks from (seed, offset), so no mask tensor is stored.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// multiplicative dropout mask for 4 consecutive elements starting at linear index e (e%4==0)
__device__ __forceinline__ float4 dropout_mask4(uint64_t e, uint32_t stream_id, float p, float scale,
                                                uint64_t seed, uint64_t offset) {
  uint64_t c = (e >> 2) + offset;
  uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), stream_id, 0u),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  constexpr float k = 2.3283064365386963e-10f;  // 2^-32
  float4 m;
  m.x = (r.x * k >= p) ? scale : 0.f;
  m.y = (r.y * k >= p) ? scale : 0.f;
  m.z = (r.z * k >= p) ? scale : 0.f;
  m.w = (r.w * k >= p) ? scale : 0.f;
  return m;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace nrms
