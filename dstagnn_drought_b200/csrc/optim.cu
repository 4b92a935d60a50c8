// The step either side of the block stack: SmoothL1 loss forward / backward
// (train_DSTAGNN_my.py:132,156), multi-tensor Adam over the flat per-block parameter / gradient buffers
// (train_DSTAGNN_my.py:133,158), the device-resident step counters that make a captured step replayable,
// and the batch-independent (adj_pa . mask_k)^T bias of cheb_conv_withSAt (model/DSTAGNN_my.py:126).
// All of them are single-pass streaming kernels: HBM-bound by construction.
#include <cstring>

#include "common.cuh"

namespace dstag {

constexpr int kOptThreads = 256;

// ---------------------------------------------------------------------------------------
// SmoothL1 (beta = 1, mean reduction): loss = mean(0.5 d^2 if |d| < 1 else |d| - 0.5), d = pred - target.
// Per-CTA partial sums land in `partials`; the last CTA to finish adds them in index order, so the loss
// is reproducible run to run (no float atomics).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kOptThreads)
smooth_l1_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ target, float* __restrict__ loss,
                     float* __restrict__ partials, unsigned int* __restrict__ done, long long n, float inv_n) {
  __shared__ float red[33];
  __shared__ bool last;
  float s = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = fabsf(__ldg(pred + i) - __ldg(target + i));
    s += d < 1.f ? 0.5f * d * d : d - 0.5f;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    float t = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(partials + i);
    t = block_sum(t, red);
    if (threadIdx.x == 0) {
      *loss = t * inv_n;
      *done = 0u;                 // self-resetting: the next launch (or graph replay) starts from zero
    }
  }
}

// dpred = gout * (d if |d| < 1 else sign(d)) / n, gout a DEVICE scalar (the gradient arriving at the loss)
__global__ void __launch_bounds__(kOptThreads)
smooth_l1_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ gout,
                     float* __restrict__ dpred, long long n, float inv_n) {
  const float g = (gout ? __ldg(gout) : 1.f) * inv_n;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = __ldg(pred + i) - __ldg(target + i);
    dpred[i] = g * fminf(fmaxf(d, -1.f), 1.f);
  }
}

// ---------------------------------------------------------------------------------------
// Adam (torch.optim.Adam defaults of the reference: no weight decay, no amsgrad) on one flat buffer:
//   m = b1 m + (1 - b1) g ; v = b2 v + (1 - b2) g^2
//   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// t is read from a device counter (1-based after dstag_tick), gscale pre-multiplies the gradient
// (1 / world_size when the all-reduce delivered a SUM).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kOptThreads)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, float lr, float b1, float b2, float eps, float gscale,
            const unsigned long long* __restrict__ step_dev, unsigned long long step_host) {
  const float t = (float)(step_dev ? *step_dev : step_host);
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x; const float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = gp[j] * gscale;
      mp[j] = b1 * mp[j] + (1.f - b1) * gj;
      vp[j] = b2 * vp[j] + (1.f - b2) * gj * gj;
      pp[j] -= step_size * mp[j] / (sqrtf(vp[j]) * inv_sqrt_bc2 + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gj = g[i] * gscale;
    const float mj = b1 * m[i] + (1.f - b1) * gj;
    const float vj = b2 * v[i] + (1.f - b2) * gj * gj;
    m[i] = mj;
    v[i] = vj;
    p[i] -= step_size * mj / (sqrtf(vj) * inv_sqrt_bc2 + eps);
  }
}

// flat[off[t] + i] = src[t][i]: the step's parameter gradients (separate tensors, wherever autograd left them)
// collected into the flat bucket that the all-reduce and the flat Adam consume -- one launch per <= 64 tensors
// instead of one accumulation kernel per parameter.
struct GatherArgs {
  const float* src[64];
  long long off[64];
  long long size[64];
};
__global__ void __launch_bounds__(256)
gather_flat_kernel(GatherArgs g, float* __restrict__ flat) {
  const int t = blockIdx.y;
  const float* __restrict__ s = g.src[t];
  float* __restrict__ d = flat + g.off[t];
  const long long n = g.size[t];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
    const long long n4 = n >> 2;
    for (long long i = i0; i < n4; i += stride) reinterpret_cast<float4*>(d)[i] = __ldg(reinterpret_cast<const float4*>(s) + i);
    for (long long i = (n4 << 2) + i0; i < n; i += stride) d[i] = __ldg(s + i);
  } else {
    for (long long i = i0; i < n; i += stride) d[i] = __ldg(s + i);
  }
}

// Up to 16 strided copies of <= 4-D fp32 blocks in ONE launch (signed element strides: a flipped axis is a
// negative stride): the per-step re-layouts of small parameter tensors -- [W_Q | W_K | W_V] packed into
// 16-byte-aligned rows, the stacked Theta_k, the (o, (d, c)) conv matrices of the GTUs and their flipped
// transposes, and the matching gradient un-packs -- that used to be ~25 torch cat / pad / flip / copy kernels
// per block.  blockIdx.y = copy, blockIdx.x strides over its elements (innermost index fastest).
struct CopyDesc {
  const float* src;
  float* dst;
  int n[4];
  long long ss[4], ds[4];
};
struct MultiCopyArgs { CopyDesc d[16]; };
__global__ void __launch_bounds__(256)
multi_copy_kernel(MultiCopyArgs a) {
  const CopyDesc& c = a.d[blockIdx.y];
  const long long total = (long long)c.n[0] * c.n[1] * c.n[2] * c.n[3];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    long long q = i;
    const int i3 = (int)(q % c.n[3]); q /= c.n[3];
    const int i2 = (int)(q % c.n[2]); q /= c.n[2];
    const int i1 = (int)(q % c.n[1]);
    const int i0 = (int)(q / c.n[1]);
    c.dst[i0 * c.ds[0] + i1 * c.ds[1] + i2 * c.ds[2] + i3 * c.ds[3]] =
        __ldg(c.src + i0 * c.ss[0] + i1 * c.ss[1] + i2 * c.ss[2] + i3 * c.ss[3]);
  }
}

__global__ void tick_kernel(unsigned long long* counters, int n) {
  if ((int)threadIdx.x < n) counters[threadIdx.x] += 1ull;
}

// biasT[k, j, i] = adj_pa[i, j] * mask_k[i, j]  (i < N), 0 for N <= i < ld.  K mask pointers arrive by value.
struct MaskPtrs { const float* m[8]; };
__global__ void __launch_bounds__(256)
bias_transpose_kernel(const float* __restrict__ adj_pa, MaskPtrs masks, float* __restrict__ biasT, int N, int ld) {
  __shared__ float tile[32][33];
  const int k = blockIdx.z;
  const float* mk = masks.m[k];
  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;       // input tile rows i, cols j
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    const int i = i0 + ty + r, j = j0 + tx;
    tile[ty + r][tx] = (i < N && j < N) ? __ldg(adj_pa + (long long)i * N + j) * __ldg(mk + (long long)i * N + j) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    const int j = j0 + ty + r, i = i0 + tx;
    if (j < N && i < ld) biasT[((long long)k * N + j) * ld + i] = tile[tx][ty + r];
  }
}

// dmask_k[i, j] = adj_pa[i, j] * dbiasT[k, j, i]: the transpose back (gradient of the kernel above)
__global__ void __launch_bounds__(256)
bias_transpose_bwd_kernel(const float* __restrict__ adj_pa, const float* __restrict__ dbiasT, float* __restrict__ dmask,
                          int N, int ld) {
  __shared__ float tile[32][33];
  const int k = blockIdx.z;
  const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;       // input tile rows j, cols i
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    const int j = j0 + ty + r, i = i0 + tx;
    tile[ty + r][tx] = (j < N && i < N) ? __ldg(dbiasT + ((long long)k * N + j) * ld + i) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    const int i = i0 + ty + r, j = j0 + tx;
    if (i < N && j < N) dmask[((long long)k * N + i) * N + j] = __ldg(adj_pa + (long long)i * N + j) * tile[tx][ty + r];
  }
}

}  // namespace dstag

using namespace dstag;

extern "C" {

int dstag_smooth_l1_fwd(const float* pred, const float* target, float* loss, float* workspace, int64_t n,
                        void* stream) {
  DSTAG_CHECK_ARG(pred && target && loss && workspace && n > 0, "dstag_smooth_l1_fwd: bad arguments");
  long long blocks = (n + kOptThreads * 4 - 1) / (kOptThreads * 4);
  if (blocks > 1024) blocks = 1024;                            // workspace: 1024 partials + 1 counter
  unsigned int* done = reinterpret_cast<unsigned int*>(workspace + 1024);
  smooth_l1_fwd_kernel<<<(unsigned)blocks, kOptThreads, 0, (cudaStream_t)stream>>>(pred, target, loss, workspace, done, n,
                                                                                  1.f / (float)n);
  DSTAG_RETURN_LAUNCH("dstag_smooth_l1_fwd");
}

int dstag_smooth_l1_bwd(const float* pred, const float* target, const float* gout, float* dpred, int64_t n,
                        void* stream) {
  DSTAG_CHECK_ARG(pred && target && dpred && n > 0, "dstag_smooth_l1_bwd: bad arguments");
  long long blocks = (n + kOptThreads - 1) / kOptThreads;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  smooth_l1_bwd_kernel<<<(unsigned)blocks, kOptThreads, 0, (cudaStream_t)stream>>>(pred, target, gout, dpred, n,
                                                                                  1.f / (float)n);
  DSTAG_RETURN_LAUNCH("dstag_smooth_l1_bwd");
}

int dstag_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                    float eps, float grad_scale, const uint64_t* step_dev, uint64_t step_host, void* stream) {
  DSTAG_CHECK_ARG(p && g && m && v && n > 0, "dstag_adam_step: bad arguments");
  DSTAG_CHECK_ARG(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                    reinterpret_cast<uintptr_t>(v)) & 15) == 0, "dstag_adam_step: buffers must be 16-byte aligned");
  DSTAG_CHECK_ARG(step_dev || step_host >= 1, "dstag_adam_step: step count starts at 1");
  long long blocks = ((n >> 2) + kOptThreads - 1) / kOptThreads;
  if (blocks < 1) blocks = 1;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  adam_kernel<<<(unsigned)blocks, kOptThreads, 0, (cudaStream_t)stream>>>(
      p, g, m, v, n, lr, beta1, beta2, eps, grad_scale, reinterpret_cast<const unsigned long long*>(step_dev),
      (unsigned long long)step_host);
  DSTAG_RETURN_LAUNCH("dstag_adam_step");
}

int dstag_gather_flat(const float* const* srcs, const int64_t* offsets, const int64_t* sizes, int32_t n, float* flat,
                      void* stream) {
  DSTAG_CHECK_ARG(srcs && offsets && sizes && flat && n > 0, "dstag_gather_flat: bad arguments");
  for (int base = 0; base < n; base += 64) {
    GatherArgs g;
    const int m = n - base < 64 ? n - base : 64;
    long long biggest = 0;
    for (int t = 0; t < 64; ++t) {
      const bool on = t < m;
      g.src[t] = on ? srcs[base + t] : nullptr;
      g.off[t] = on ? offsets[base + t] : 0;
      g.size[t] = on ? sizes[base + t] : 0;
      DSTAG_CHECK_ARG(!on || (g.src[t] && g.size[t] >= 0 && g.off[t] >= 0), "dstag_gather_flat: bad entry %d", base + t);
      if (g.size[t] > biggest) biggest = g.size[t];
    }
    long long bx = (biggest / 4 + 255) / 256;
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;
    gather_flat_kernel<<<dim3((unsigned)bx, (unsigned)m), 256, 0, (cudaStream_t)stream>>>(g, flat);
    DSTAG_CHECK_LAUNCH("dstag_gather_flat");
  }
  return 0;
}

int dstag_multi_copy(const dstag_copy_t* copies, int32_t n, void* stream) {
  DSTAG_CHECK_ARG(copies && n > 0, "dstag_multi_copy: bad arguments");
  for (int base = 0; base < n; base += 16) {
    MultiCopyArgs a;
    memset(&a, 0, sizeof(a));
    const int m = n - base < 16 ? n - base : 16;
    long long biggest = 0;
    for (int t = 0; t < m; ++t) {
      const dstag_copy_t& c = copies[base + t];
      DSTAG_CHECK_ARG(c.src && c.dst, "dstag_multi_copy: null pointer in copy %d", base + t);
      a.d[t].src = (const float*)c.src;
      a.d[t].dst = (float*)c.dst;
      long long tot = 1;
      for (int k = 0; k < 4; ++k) {
        DSTAG_CHECK_ARG(c.n[k] >= 1, "dstag_multi_copy: extents must be >= 1");
        a.d[t].n[k] = c.n[k];
        a.d[t].ss[k] = c.src_stride[k];
        a.d[t].ds[k] = c.dst_stride[k];
        tot *= c.n[k];
      }
      if (tot > biggest) biggest = tot;
    }
    long long bx = (biggest + 255) / 256;
    if (bx > 128) bx = 128;
    multi_copy_kernel<<<dim3((unsigned)bx, (unsigned)m), 256, 0, (cudaStream_t)stream>>>(a);
    DSTAG_CHECK_LAUNCH("dstag_multi_copy");
  }
  return 0;
}

int dstag_tick(uint64_t* counters, int32_t n, void* stream) {
  DSTAG_CHECK_ARG(counters && n > 0 && n <= 32, "dstag_tick: bad arguments");
  tick_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(counters), n);
  DSTAG_RETURN_LAUNCH("dstag_tick");
}

int dstag_bias_transpose(const float* adj_pa, const float* const* masks, int32_t K, float* biasT, int32_t N,
                         int32_t ld, void* stream) {
  DSTAG_CHECK_ARG(adj_pa && masks && biasT && K > 0 && K <= 8 && N > 0 && ld >= N, "dstag_bias_transpose: bad arguments");
  MaskPtrs mp;
  for (int k = 0; k < 8; ++k) mp.m[k] = k < K ? masks[k] : nullptr;
  for (int k = 0; k < K; ++k) DSTAG_CHECK_ARG(mp.m[k] != nullptr, "dstag_bias_transpose: null mask pointer");
  const dim3 grid((ld + 31) / 32, (N + 31) / 32, K);
  bias_transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(adj_pa, mp, biasT, N, ld);
  DSTAG_RETURN_LAUNCH("dstag_bias_transpose");
}

int dstag_bias_transpose_bwd(const float* adj_pa, const float* dbiasT, float* dmask, int32_t K, int32_t N, int32_t ld,
                             void* stream) {
  DSTAG_CHECK_ARG(adj_pa && dbiasT && dmask && K > 0 && N > 0 && ld >= N, "dstag_bias_transpose_bwd: bad arguments");
  const dim3 grid((N + 31) / 32, (N + 31) / 32, K);
  bias_transpose_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(adj_pa, dbiasT, dmask, N, ld);
  DSTAG_RETURN_LAUNCH("dstag_bias_transpose_bwd");
}

}  // extern "C"
