// Probe: how fast can one CTA per SM stream a row-major fp32 matrix from DRAM with TMA boxes of the engine's k-block
// shape (128 rows x 128 bytes, SWIZZLE_128B), and how does the rate change with the box shape / boxes per stage /
// CTAs per SM / stages in flight?  Nothing consumes the tiles: a consumer warp waits for the bytes and hands the
// stage straight back, so the number is the service rate of the TMA path alone.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu   (run: ./tma_stream)
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// A tile = 128 rows x (kw * 32) floats, fetched as `split` boxes of 128 / split rows (x kw boxes of 32 floats when the
// map's box is 32 floats wide).  Tiles walk along a row tile first (consecutive k-blocks), like the engine.
__global__ void __launch_bounds__(64) stream(const __grid_constant__ CUtensorMap map, int row_tiles, int kblocks, int stages,
                                             int split, int box_cols, int kw, int stage_bytes, unsigned long long* bytes_out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + stages * stage_bytes);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  const int ksteps = (kblocks + kw - 1) / kw;
  unsigned it = 0;
  if (warp == 0) {
    if (threadIdx.x == 0) {
      for (int t = blockIdx.x; t < row_tiles; t += gridDim.x)
        for (int k = 0; k < ksteps; ++k, ++it) {
          const int s = it % stages;
          mbar_wait(empty + s, ((it / stages) & 1) ^ 1);
          mbar_expect(full + s, stage_bytes);
          uint8_t* dst = smem + s * stage_bytes;
          const int rows = 128 / split;
          for (int b = 0; b < split; ++b)
            for (int c = 0; c < kw * 32 / box_cols; ++c)
              tma_2d(dst + (b * (kw * 32 / box_cols) + c) * rows * box_cols * 4, &map, full + s, k * kw * 32 + c * box_cols,
                     t * 128 + b * rows);
        }
    }
  } else {
    unsigned long long n = 0;
    if (threadIdx.x == 32) {
      for (int t = blockIdx.x; t < row_tiles; t += gridDim.x)
        for (int k = 0; k < ksteps; ++k, ++it) {
          const int s = it % stages;
          mbar_wait(full + s, (it / stages) & 1);
          n += stage_bytes;
          mbar_arrive(empty + s);
        }
      atomicAdd(bytes_out, n);
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)sym;
  const long long rows = 192LL * 896, ld = 896, cols = 883;           // the PEMS07 softmax maps (B = 64, K = 3)
  float* m;
  cudaMalloc(&m, rows * ld * 4);
  cudaMemset(m, 0, rows * ld * 4);
  unsigned long long* cnt;
  cudaMalloc(&cnt, 8);
  float* flush;
  cudaMalloc(&flush, 256 << 20);
  struct V { const char* name; int box_rows, box_cols, split, kw, stages, ctas_per_sm; CUtensorMapSwizzle sw; CUtensorMapL2promotion l2; };
  const V vs[] = {
      {"128 x 128 B box, 6 stages (engine)", 128, 32, 1, 1, 6, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
      {"same, 3 stages", 128, 32, 1, 1, 3, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
      {"same, 12 stages", 128, 32, 1, 1, 12, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
      {"same, 6 stages, L2 promotion 256 B", 128, 32, 1, 1, 6, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"same, 6 stages, no L2 promotion", 128, 32, 1, 1, 6, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE},
      {"4 boxes of 32 rows x 128 B, 6 stages", 32, 32, 4, 1, 6, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
      {"2 k-blocks per stage (2 boxes), 6 stages", 128, 32, 1, 2, 6, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
      {"2 k-blocks per stage, L2 promotion 256 B", 128, 32, 1, 2, 6, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"128 x 256 B box (no swizzle), 6 stages", 128, 64, 1, 2, 6, 1, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"128 x 128 B box, 2 CTAs / SM x 6 stages", 128, 32, 1, 1, 6, 2, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
      {"128 x 128 B box, 4 CTAs / SM x 3 stages", 128, 32, 1, 1, 3, 4, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
  };
  for (const V& v : vs) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)v.box_cols, (cuuint32_t)v.box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, m, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, v.sw,
                     v.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-48s encode failed (%d)\n", v.name, (int)r); continue; }
    const int stage_bytes = 128 * 32 * 4 * v.kw;
    const int smem = v.stages * stage_bytes + 2 * v.stages * 8 + 1024;
    cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int row_tiles = (int)(rows / 128), kblocks = 28;
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaMemset(flush, 1, 256 << 20);
      cudaMemset(cnt, 0, 8);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      stream<<<148 * v.ctas_per_sm, 64, smem>>>(map, row_tiles, kblocks, v.stages, v.split, v.box_cols, v.kw, stage_bytes, cnt);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%-48s failed: %s\n", v.name, cudaGetErrorString(e)); return 1; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    unsigned long long n;
    cudaMemcpy(&n, cnt, 8, cudaMemcpyDeviceToHost);
    printf("%-48s %8.3f ms  %7.1f GB/s  (%.0f MB landed)\n", v.name, best, n / best / 1e6, n / 1e6);
  }
  return 0;
}
