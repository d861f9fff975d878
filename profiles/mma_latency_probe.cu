// Microbenchmark: issue-to-completion time of chains of tcgen05.mma (kind::f16, SS form, M=128) on one SM.
//   dependent chain (one accumulator) vs round-robin over several accumulators, for N in {32, 64, 128, 256}.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I newsrecommendationsystem_b200/csrc \
//        profiles/mma_latency_probe.cu -o profiles/_bin/mma_latency_probe
#include <cstdio>
#include <cuda.h>
#include "tc_common.cuh"
using namespace nrms::tc;

__global__ void __launch_bounds__(128, 1) probe(int N, int n_acc, int chain, int ts_form, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar_storage;
  const uint32_t bar = smem_u32(&bar_storage);
  for (int i = threadIdx.x; i < 98304 / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_ptr), 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, N);
    const uint64_t desc0 = umma_desc_k_sw128(0);
    const uint32_t sa = base >> 4, sb = (base + 32768) >> 4;
    const int acc_stride = 512 / n_acc >= N ? N : 512 / n_acc;   // keep n_acc * N <= 512 in the callers
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      for (int i = 0; i < chain; ++i) {
        const uint32_t d = tm + (i % n_acc) * acc_stride;
        if (ts_form)
          umma_f16_ts(d, tm + 448 + (i & 3) * 8, desc0 | (uint64_t)((sb + 2 * (i & 3)) & 0x3FFF), idesc, i >= n_acc);
        else
          umma_f16_ss(d, desc0 | (uint64_t)((sa + 2 * (i & 3)) & 0x3FFF), desc0 | (uint64_t)((sb + 2 * (i & 3)) & 0x3FFF),
                      idesc, i >= n_acc);
      }
      long long t1 = clock64();
      umma_commit(bar);
      mbar_wait(bar, rep & 1);
      long long t2 = clock64();
      out[2 * rep] = t1 - t0;
      out[2 * rep + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100352);
  const int chain = 64;
  printf("form  N  n_acc  issue_cyc/mma  total_cyc/mma  (chain %d, floor = N/4 cycles)\n", chain);
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {32, 64, 128, 256})
      for (int n_acc : {1, 2, 4}) {
        if (n_acc * N > (ts ? 448 : 512)) continue;
        probe<<<1, 128, 100352>>>(N, n_acc, chain, ts, d);
        long long h[6];
        cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        printf("%s  %3d  %d  %8.1f  %8.1f\n", ts ? "TS" : "SS", N, n_acc, (double)h[4] / chain, (double)h[5] / chain);
      }
  return 0;
}
