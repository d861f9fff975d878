// Microbenchmark 2: what paces a chain of tcgen05.mma (kind::f16, M=128) -- the issuing code, the tensor pipe, or
// other warps of the CTA competing for a shared resource?
//   issue modes   0: `if (threadIdx.x == 0)` branch, runtime loop
//                 1: whole warp converged, elect.sync, fully unrolled chain with compile-time operands
//                 2: mode 1 from TWO warps at once (separate accumulators): issue-bound -> same time, pipe-bound -> 2x
//                 3: mode 1 with a tcgen05.commit every 4 MMAs (as the K1 weight ring does)
//   noise kinds (8 extra warps, two per SM sub-partition, loop until the issuer is done)
//                 0 none | 1 tcgen05.ld x16 + wait | 2 tcgen05.st x16 + wait | 3 LDS.128/STS.128 | 4 mbarrier.try_wait spin
//                 5 FFMA/ex2 ALU loop | 6 one thread streaming 16 KB cp.async.bulk global->shared
// All CTAs of the grid run the same thing; block 0 reports.  Cycles are per MMA.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I newsrecommendationsystem_b200/csrc \
//        profiles/mma_issue_probe.cu -o profiles/_bin/mma_issue_probe
#include <cstdio>
#include <cuda.h>
#include "tc_common.cuh"
using namespace nrms::tc;

constexpr int CHAIN = 64;
constexpr int THREADS = 384;   // warps 0,1 issuers | 2,3 idle | 4..11 noise

template <int N, bool TS>
__device__ __forceinline__ void chain_unrolled(uint32_t tm, uint32_t ta, uint32_t sa, uint32_t sb, uint32_t bar, bool commits) {
  const uint32_t idesc = umma_idesc_f16(128, N);
  const uint64_t desc0 = umma_desc_k_sw128(0);
  if (elect_one()) {
#pragma unroll
    for (int i = 0; i < CHAIN; ++i) {
      if (TS)
        umma_f16_ts(tm, ta + (i & 3) * 8, desc0 | (uint64_t)((sb + 2 * (i & 3)) & 0x3FFF), idesc, i > 0);
      else
        umma_f16_ss(tm, desc0 | (uint64_t)((sa + 2 * (i & 3)) & 0x3FFF), desc0 | (uint64_t)((sb + 2 * (i & 3)) & 0x3FFF),
                    idesc, i > 0);
      if (commits && (i & 3) == 3 && i != CHAIN - 1) umma_commit(bar + 8);   // a barrier nobody waits on
    }
  }
  __syncwarp();
}

template <int N, bool TS>
__global__ void __launch_bounds__(THREADS, 1) probe(int mode, int noise, const uint8_t* gsrc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar_storage[8];
  __shared__ volatile int done_flag;
  const uint32_t bar = smem_u32(&bar_storage[0]);
  for (int i = threadIdx.x; i < 131072 / 16; i += THREADS) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (noise >= 100) {   // pseudo-random fp16 values in (-1, 1): sign | exponent 01110/01101 | random mantissa
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
      auto rnd = [&]() { h ^= h << 13; h ^= h >> 17; h ^= h << 5; return (h & 0x83FF83FFu) | 0x38003400u; };
      v = make_uint4(rnd(), rnd(), rnd(), rnd());
    }
    reinterpret_cast<uint4*>(sm)[i] = v;
  }
  if (noise >= 100) noise -= 100;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(bar + 8 * i, 1);
    mbar_fence_init();
    done_flag = 0;
  }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_ptr), 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sa = base >> 4, sb = (base + 32768) >> 4;
  const int n_issuers = (mode == 2) ? 2 : 1;
  if (warp < n_issuers) {
    if (mode == 0) {
      if (lane == 0) {
        const uint32_t idesc = umma_idesc_f16(128, N);
        const uint64_t desc0 = umma_desc_k_sw128(0);
        for (int rep = 0; rep < 3; ++rep) {
          long long t0 = clock64();
          for (int i = 0; i < CHAIN; ++i) {
            if (TS)
              umma_f16_ts(tm, tm + 384 + (i & 3) * 8, desc0 | (uint64_t)((sb + 2 * (i & 3)) & 0x3FFF), idesc, i > 0);
            else
              umma_f16_ss(tm, desc0 | (uint64_t)((sa + 2 * (i & 3)) & 0x3FFF),
                          desc0 | (uint64_t)((sb + 2 * (i & 3)) & 0x3FFF), idesc, i > 0);
          }
          long long t1 = clock64();
          umma_commit(bar);
          mbar_wait(bar, rep & 1);
          long long t2 = clock64();
          if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
      }
      __syncwarp();
    } else {
      const uint32_t mybar = bar + 16 * warp;
      const uint32_t mytm = tm + warp * 128;     // N <= 128 in mode 2
      for (int rep = 0; rep < 3; ++rep) {
        long long t0 = clock64();
        chain_unrolled<N, TS>(mytm, tm + 384, sa, sb + warp * 1024, mybar, mode == 3);
        long long t1 = clock64();
        if (elect_one()) umma_commit(mybar);
        __syncwarp();
        mbar_wait(mybar, rep & 1);
        long long t2 = clock64();
        if (blockIdx.x == 0 && lane == 0) { out[2 * warp] = t1 - t0; out[2 * warp + 1] = t2 - t0; }
      }
    }
    __syncwarp();
    if (warp == 0 && lane == 0) done_flag = 1;
  } else if (warp >= 4 && noise != 0) {
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tn = tm + 448 + lane_addr;     // columns [448, 464): never touched by the MMAs
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 16 + i;
    uint8_t* myrow = sm + 65536 + threadIdx.x * 64;   // private 64 bytes
    float acc = (float)threadIdx.x;
    uint32_t bulk_it = 0;
    while (!done_flag) {
      if (noise == 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { tmem_ld16_nw(tn, r); tmem_ld_wait(); }
      } else if (noise == 2) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { tmem_st16(tn, r); tmem_st_wait(); }
      } else if (noise == 3) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint4 v;
          const uint32_t a = smem_u32(myrow + 16 * k), b = smem_u32(myrow + 16 * ((k + 1) & 3));
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
          v.x += r[k];
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(b), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
      } else if (noise == 4) {
        for (int k = 0; k < 4; ++k) (void)mbar_test(bar + 56, 0);       // never completes
        uint32_t d;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(d) : "r"(bar + 56), "r"(0u) : "memory");
        r[0] += d;
      } else if (noise == 5) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          float y;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(acc * 1e-3f));
          acc = fmaf(acc, 0.999f, y);
        }
      } else if (noise == 6) {
        if (warp == 4 && lane == 0) {
          const uint32_t b6 = bar + 48;
          const uint32_t dst = base + 98304 + (bulk_it & 1) * 16384;
          asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(b6), "r"(16384u)
                       : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst), "l"(gsrc + (size_t)(bulk_it % 40) * 16384), "r"(16384u), "r"(b6) : "memory");
          mbar_wait(b6, bulk_it & 1);
          ++bulk_it;
        }
      }
    }
    if (noise == 5 && acc == 123.456f) out[7] = 1;
    if (noise != 5 && r[0] == 0xdeadbeefu && r[5] == 77u) out[7] = 2;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int N, bool TS>
void run(int grid, int mode, int noise, const uint8_t* gsrc, long long* d) {
  cudaFuncSetAttribute(probe<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 133120);
  cudaMemset(d, 0, 64);
  probe<N, TS><<<grid, THREADS, 133120>>>(mode, noise, gsrc, d);
  long long h[4];
  cudaError_t e = cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  printf("%s N=%3d grid=%3d mode=%d noise=%d  issue %7.1f total %7.1f", TS ? "TS" : "SS", N, grid, mode, noise,
         (double)h[0] / CHAIN, (double)h[1] / CHAIN);
  if (mode == 2) printf("   | warp1 issue %7.1f total %7.1f", (double)h[2] / CHAIN, (double)h[3] / CHAIN);
  printf("   (floor %d)\n", N / 2);
  fflush(stdout);
}

int main() {
  long long* d;
  uint8_t* g;
  cudaMalloc(&d, 64);
  cudaMalloc(&g, 40 * 16384);
  cudaMemset(g, 0, 40 * 16384);
  for (int grid : {148})
    for (int mode = 1; mode < 4; mode += 2) {
      run<32, false>(grid, mode, 0, g, d);
      run<128, false>(grid, mode, 0, g, d);
      if (mode != 2) run<256, false>(grid, mode, 0, g, d);
      run<32, true>(grid, mode, 0, g, d);
      run<128, true>(grid, mode, 0, g, d);
    }
  printf("--- random operand data (noise += 100)\n");
  for (int grid : {1, 148}) {
    run<128, false>(grid, 3, 100, g, d);
    run<256, false>(grid, 3, 100, g, d);
    run<32, true>(grid, 3, 100, g, d);
    run<128, true>(grid, 3, 100, g, d);
    run<128, false>(grid, 2, 100, g, d);
  }
  printf("--- noise sweep (grid 148, mode 3)\n");
  for (int noise = 0; noise <= 6; ++noise) {
    run<128, false>(148, 3, noise, g, d);
    run<32, true>(148, 3, noise, g, d);
    run<128, true>(148, 3, noise, g, d);
  }
  return 0;
}
