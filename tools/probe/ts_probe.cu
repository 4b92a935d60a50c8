// Hardware probe (build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o ts_probe ts_probe.cu): does tcgen05.mma (kind::tf32) with the A operand in TMEM expect "lane = row m, column = k"?
// One CTA, M = 128, N = 128, K = 32.  A goes registers -> TMEM (tcgen05.st 32x32b), B sits K-major SWIZZLE_128B
// in shared memory.  Prints the max error against a CPU product of the tf32-truncated inputs.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_k(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 16384 + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // B tile: row n, 8 chunks of 4 k
  for (int e = tid; e < 128 * 8; e += 128) {
    const int n = e >> 3, c = e & 7;
    *reinterpret_cast<float4*>(smem + swz(n, c)) = *reinterpret_cast<const float4*>(B + n * 32 + 4 * c);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *slot;
  // A row of this thread -> TMEM lane (32 * warp + lane), columns 128 .. 159
  uint32_t a[32];
  for (int k = 0; k < 32; ++k) a[k] = __float_as_uint(A[tid * 32 + k]);
  const uint32_t ta = tbase + ((uint32_t)(warp * 32) << 16) + 128;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(ta),
      "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]),
      "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]), "r"(a[16]), "r"(a[17]), "r"(a[18]),
      "r"(a[19]), "r"(a[20]), "r"(a[21]), "r"(a[22]), "r"(a[23]), "r"(a[24]), "r"(a[25]), "r"(a[26]), "r"(a[27]),
      "r"(a[28]), "r"(a[29]), "r"(a[30]), "r"(a[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t bdesc = make_desc_k(smem_u32(smem));
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t acc = kk ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tbase),
          "r"(tbase + 128 + 8 * kk), "l"(bdesc + 2ull * kk), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  // wait for the MMAs
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int cc = 0; cc < 4; ++cc) {
    uint32_t r[32];
    const uint32_t td = tbase + ((uint32_t)(warp * 32) << 16) + cc * 32;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(td));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[tid * 128 + cc * 32 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tbase) : "memory");
}

static float trunc_tf32(float v) { uint32_t b; memcpy(&b, &v, 4); b &= 0xFFFFE000u; memcpy(&v, &b, 4); return v; }
#include <cstring>
int main() {
  std::vector<float> A(128 * 32), B(128 * 32), D(128 * 128), R(128 * 128);
  srand(1);
  for (auto& v : A) v = (float)rand() / RAND_MAX - 0.5f;
  for (auto& v : B) v = (float)rand() / RAND_MAX - 0.5f;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      double s = 0;
      for (int k = 0; k < 32; ++k) s += (double)trunc_tf32(A[m * 32 + k]) * trunc_tf32(B[n * 32 + k]);
      R[m * 128 + n] = (float)s;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  probe<<<1, 128, 16384 + 1024 + 256>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < D.size(); ++i) { maxerr = fmax(maxerr, fabs(D[i] - R[i])); maxref = fmax(maxref, fabs(R[i])); }
  printf("TS probe: max|D - ref| = %.3e (max|ref| = %.3e) -> %s\n", maxerr, maxref, maxerr < 1e-4 * maxref ? "LAYOUT OK" : "MISMATCH");
  printf("D[0][0..3] = %f %f %f %f | ref %f %f %f %f\n", D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3]);
  printf("D[5][7] = %f ref %f ; D[100][64] = %f ref %f\n", D[5 * 128 + 7], R[5 * 128 + 7], D[100 * 128 + 64], R[100 * 128 + 64]);
  return 0;
}
