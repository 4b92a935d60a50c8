// Probe: what paces a "registers -> TMEM (tcgen05.st) -> TS-form tcgen05.mma" producer / consumer pipeline?
// 4 (or 8) producer warps store one 128 x 32 k-block (hi | lo = 64 columns) per iteration into a ring of TMEM slots, one
// MMA warp consumes it with 12 MMAs (3xTF32, N = 64) and commits the slot back.  Variants switch parts off.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ts_pipeline ts_pipeline.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{ .reg .pred P; elect.sync _|P, 0xffffffff; selp.u32 %0, 1, 0, P; }" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ uint64_t make_desc_k(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p; }" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* b) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory"); }

#define R32(r) "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), \
  "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), \
  "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), \
  "r"(r[30]), "r"(r[31])
#define L32 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32}"
__device__ __forceinline__ void st32(uint32_t t, const uint32_t (&r)[32]) { asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], " L32 ";" ::"r"(t), R32(r) : "memory"); }

constexpr int SLOTS = 6;
// flags: 1 = no MMAs (wait + commit only), 2 = MMA warp spins with nanosleep back-off, 4 = two producer teams,
//        8 = producers skip tcgen05.st, 16 = N = 128 instead of 64, 32 = 4 MMAs per k-block instead of 12,
//        64 = producers also write shared memory and issue fence.proxy.async every iteration
__global__ void __launch_bounds__(416, 1) k(long long* out, int iters, int flags) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + 64 * 1024);
  uint64_t* empty = full + SLOTS;
  uint32_t* slot = (uint32_t*)(empty + SLOTS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) ((float*)smem)[i] = 1.0f;      // B tiles: anything finite
  if (threadIdx.x == 0) {
    for (int s = 0; s < SLOTS; ++s) { mbar_init(full + s, 4); mbar_init(empty + s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = *slot;
  const int teams = (flags & 4) ? 2 : 1;
  long long t0 = clock64(), busy = 0;
  if (warp < 4 * teams) {
    const int team = warp >> 2, q = warp & 3;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = 0x3f800000u + lane + i;
    for (int it = team; it < iters; it += teams) {
      const int s = it % SLOTS;
      if (lane == 0) mbar_wait(empty + s, ((it / SLOTS) & 1) ^ 1);
      __syncwarp();
      const long long b0 = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ta = base + 128 + s * 64 + ((uint32_t)(32 * q) << 16);
      if (flags & 64) {                                   // a shared-memory write that the tensor core must see
        *reinterpret_cast<float4*>(smem + 32768 + threadIdx.x * 16) = make_float4(1.f, 1.f, 1.f, 1.f);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      if (!(flags & 8)) { st32(ta, r); st32(ta + 32, r); }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(full + s);
      busy += clock64() - b0;
    }
  } else if (warp == 12) {
    const int N = (flags & 16) ? 128 : 64;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sb = smem_u32(smem);
    for (int it = 0; it < iters; ++it) {
      const int s = it % SLOTS;
      if (flags & 2) { if (lane == 0) { while (true) { uint32_t ok; asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(full + s)), "r"((uint32_t)((it / SLOTS) & 1)) : "memory"); if (ok) break; __nanosleep(40); } } __syncwarp(); }
      else mbar_wait(full + s, (it / SLOTS) & 1);
      const long long b0 = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        if (!(flags & 1)) {
          const uint64_t bh = make_desc_k(sb), bl = make_desc_k(sb + 32768);
          const int nk = (flags & 32) ? 1 : 4;
          for (int kk = 0; kk < nk; ++kk) {
            const uint32_t ta = base + 128 + s * 64 + 8 * kk;
            mma_ts(base, ta + 32, bh + 2ull * kk, idesc, (it | kk) ? 1u : 0u);
            mma_ts(base, ta, bl + 2ull * kk, idesc, 1u);
            mma_ts(base, ta, bh + 2ull * kk, idesc, 1u);
          }
        }
        commit(empty + s);
      }
      __syncwarp();
      busy += clock64() - b0;
    }
  }
  const long long t1 = clock64();
  if (lane == 0) { out[(blockIdx.x * 16 + warp) * 2] = t1 - t0; out[(blockIdx.x * 16 + warp) * 2 + 1] = busy; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 16 * 2 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  const int iters = 4000;
  const char* names[] = {"full pipeline", "no MMAs (wait + commit)", "MMA warp backs off", "two producer teams", "two teams, no MMAs", "producers skip tcgen05.st", "N = 128",
                         "4 MMAs per k-block", "two teams, 4 MMAs", "two teams, N = 128", "+ st.shared + fence.proxy.async", "same, two teams"};
  const int fl[] = {0, 1, 2, 4, 5, 8, 16, 32, 36, 20, 64, 68};
  for (int v = 0; v < 12; ++v) {
    k<<<148, 416, 128 * 1024>>>(d, 100, fl[v]);
    k<<<148, 416, 128 * 1024>>>(d, iters, fl[v]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", names[v], cudaGetErrorString(e)); return 1; }
    long long h[148 * 16 * 2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-28s: %7.0f cycles per k-block   (producer warp 0 busy %5.0f per own iteration, MMA warp busy %5.0f)\n", names[v],
           (double)h[0] / iters, (double)h[1] / (iters / ((fl[v] & 4) ? 2 : 1)), (double)h[12 * 2 + 1] / iters);
  }
  return 0;
}
