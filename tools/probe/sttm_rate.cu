// Probe: throughput of tcgen05.st (registers -> tensor memory) by shape and warp count, one CTA per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sttm_rate sttm_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define R32(r) "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), \
  "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), \
  "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), \
  "r"(r[30]), "r"(r[31])
#define L32 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32}"

template <int SHAPE>
__device__ __forceinline__ void st(uint32_t taddr, const uint32_t (&r)[32]) {
  if (SHAPE == 0) asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], " L32 ";" ::"r"(taddr), R32(r) : "memory");
  if (SHAPE == 1) asm volatile("tcgen05.st.sync.aligned.16x256b.x8.b32 [%0], " L32 ";" ::"r"(taddr), R32(r) : "memory");
  if (SHAPE == 2) asm volatile("tcgen05.st.sync.aligned.16x128b.x16.b32 [%0], " L32 ";" ::"r"(taddr), R32(r) : "memory");
  if (SHAPE == 3) asm volatile("tcgen05.st.sync.aligned.16x64b.x32.b32 [%0], " L32 ";" ::"r"(taddr), R32(r) : "memory");
}

// every warp stores 32 registers per thread = 4 KB per instruction
template <int SHAPE, bool WAIT_EACH>
__global__ void __launch_bounds__(512, 1) k(long long* out, int iters, int nwarps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = lane * 32 + i;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    // quarter = warp % 4; second warp of a quarter writes other columns
    const uint32_t ta = base + ((uint32_t)(32 * (warp & 3)) << 16) + 64 * (warp >> 2);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      st<SHAPE>(ta + 32 * (it & 1), r);
      if (WAIT_EACH) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    t1 = clock64();
  }
  if (lane == 0 && warp < nwarps) out[blockIdx.x * 16 + warp] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
}

template <int SHAPE, bool WAIT_EACH>
void run(const char* name, long long* d, int nwarps) {
  const int iters = 2000;
  k<SHAPE, WAIT_EACH><<<148, 512>>>(d, 10, nwarps);
  k<SHAPE, WAIT_EACH><<<148, 512>>>(d, iters, nwarps);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[148 * 16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = 0;
  for (int w = 0; w < nwarps; ++w) cyc += (double)h[w];
  cyc /= nwarps;
  printf("%-28s warps %2d wait_each %d: %7.1f cycles per 4 KB store per warp -> %6.1f B/clk per SM\n", name, nwarps, (int)WAIT_EACH,
         cyc / iters, 4096.0 * nwarps * iters / cyc);
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 16 * 8);
  for (int nw : {1, 4, 8}) {
    run<0, false>("32x32b.x32", d, nw);
    run<0, true>("32x32b.x32", d, nw);
    run<1, false>("16x256b.x8", d, nw);
    run<2, false>("16x128b.x16", d, nw);
    run<3, false>("16x64b.x32", d, nw);
  }
  return 0;
}
