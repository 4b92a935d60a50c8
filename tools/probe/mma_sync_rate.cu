// Probe: peak rate of the legacy warp-level tensor path (mma.sync.m16n8k8 tf32) on this GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_rate mma_sync_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = threadIdx.x ^ 5, b1 = 11;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  for (int warps_per_sm : {4, 8, 16, 32}) {
    int blocks = p.multiProcessorCount * (warps_per_sm * 32 / 256 > 0 ? warps_per_sm * 32 / 256 : 1);
    int threads = warps_per_sm >= 8 ? 256 : warps_per_sm * 32;
    float* out; cudaMalloc(&out, (size_t)blocks * threads * 4);
    int iters = 20000;
    k<<<blocks, threads>>>(out, 100);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<<<blocks, threads>>>(out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double flops = (double)blocks * (threads / 32) * iters * 8 * (16.0 * 8 * 8 * 2);
    printf("warps/SM %2d: %.1f TFLOP/s (mma.sync m16n8k8 tf32)\n", warps_per_sm, flops / (ms * 1e-3) / 1e12);
    cudaFree(out);
  }
  return 0;
}
