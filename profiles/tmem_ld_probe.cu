// Microbenchmark 3: per-SM throughput of the worker-side primitives of K1 -- tcgen05.ld / tcgen05.st, ex2.approx,
// cvt.rna.tf32, cvt.rn.f16.f32 and the packed cvt.rn.f16x2 -- as a function of the number of warps issuing them.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I newsrecommendationsystem_b200/csrc \
//        profiles/tmem_ld_probe.cu -o profiles/_bin/tmem_ld_probe
#include <cstdio>
#include <cuda.h>
#include <cuda_fp16.h>
#include "tc_common.cuh"
using namespace nrms::tc;

constexpr int ITERS = 256;

__device__ __forceinline__ void tmem_ld32_nw(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// kind: 0 ld.x16 + wait each | 1 4 x ld.x16 then one wait | 2 ld.x32 + wait | 3 st.x16 + wait | 4 ex2 | 5 cvt.rna.tf32
//       6 cvt.rn.f16.f32 | 7 cvt.rn.f16x2.f32 | 8 FFMA (reference: full-rate pipe) | 9 ex2.approx.f16x2
__global__ void __launch_bounds__(512, 1) probe(int kind, int n_warps, long long* out, float* sink) {
  __shared__ uint32_t tmem_ptr;
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_ptr), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  const int warp = threadIdx.x >> 5;
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x * 32 + i;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.001f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  if (warp < n_warps) {
    const uint32_t tn = tm + lane_addr + (warp >> 2) * 32;     // each warp of a quarter reads its own 32 columns
    if (kind == 0) {
      for (int it = 0; it < ITERS; ++it) { tmem_ld16_nw(tn + (it & 1) * 16, r); tmem_ld_wait(); }
    } else if (kind == 1) {
      for (int it = 0; it < ITERS / 4; ++it) {
        tmem_ld16_nw(tn, r); tmem_ld16_nw(tn + 16, r + 16); tmem_ld16_nw(tn, r); tmem_ld16_nw(tn + 16, r + 16);
        tmem_ld_wait();
      }
    } else if (kind == 2) {
      for (int it = 0; it < ITERS / 2; ++it) { tmem_ld32_nw(tn, r); tmem_ld_wait(); }
    } else if (kind == 3) {
      for (int it = 0; it < ITERS; ++it) { tmem_st16(tn + (it & 1) * 16, r); tmem_st_wait(); }
    } else if (kind == 4) {
      for (int it = 0; it < ITERS * 2; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(acc[i]));
      }
    } else if (kind == 5) {
      for (int it = 0; it < ITERS * 2; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint32_t y;
          asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(acc[i]));
          acc[i] = __uint_as_float(y + 0x2001u);      // dependent chain: cannot be hoisted or merged
        }
      }
    } else if (kind == 6) {
      for (int it = 0; it < ITERS * 2; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          unsigned short h;
          asm volatile("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(acc[i] + (float)it));
          r[i] += h;
        }
      }
    } else if (kind == 7) {
      for (int it = 0; it < ITERS * 2; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint32_t y;
          asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(acc[i]), "f"(acc[i]));
          acc[i] = __uint_as_float((y & 0x03FF03FFu) | 0x3F800000u);
        }
      }
    } else if (kind == 9) {
      for (int it = 0; it < ITERS * 2; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[i]));
      }
    } else {
      for (int it = 0; it < ITERS * 2; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i]) : "f"(0.999f), "f"(0.5f));
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  float sacc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sacc += acc[i];
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) x ^= r[i];
  if (x == 0x12345678u && sacc == 1.2345f) sink[0] = 1.f;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  float* s;
  cudaMalloc(&d, 64);
  cudaMalloc(&s, 64);
  const char* names[] = {"tcgen05.ld.x16 +wait", "4x ld.x16, 1 wait", "tcgen05.ld.x32 +wait", "tcgen05.st.x16 +wait",
                         "ex2.approx", "cvt.rna.tf32", "cvt.rn.f16.f32", "cvt.rn.f16x2.f32", "fma.rn.f32", "ex2.approx.f16x2 (pairs)"};
  for (int kind = 4; kind <= 9; ++kind)
    for (int nw : {4, 8, 16}) {
      probe<<<148, 512>>>(kind, nw, d, s);
      long long h;
      cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      if (kind <= 3) {
        const double bytes = (double)nw * ITERS * 32 * 64;      // 16 columns x 4 B per lane per x16 access
        printf("%-22s warps %2d  cycles %8lld  -> %7.1f B/clk/SM\n", names[kind], nw, h, bytes / h);
      } else {
        const double ops = (double)nw * 32 * ITERS * 2 * 8;
        printf("%-22s warps %2d  cycles %8lld  -> %7.2f thread-ops/clk/SM\n", names[kind], nw, h, ops / h);
      }
    }
  return 0;
}
