// Legacy mma.sync TF32 throughput on sm_100a: cycles per HMMA.1688.F32.TF32 (m16n8k8) per SM sub-partition.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o hmma_tf32_probe hmma_tf32_probe.cu && ./hmma_tf32_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(float* out, long long* cyc, int iters) {
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  for (int warps : {4, 8, 16, 32}) {
    probe<<<148, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_smsp = (double)c / ((double)iters * 8 * warps / 4);
    printf("m16n8k8.tf32 warps/SM %2d: %.2f cycles per MMA per SMSP (%.0f dense TF32 TFLOP/s at 148 SMs x 1.965 GHz)\n", warps,
           per_smsp, 2048.0 * 4 / per_smsp * 148 * 1.965e9 / 1e12);
  }
  return 0;
}
