// HBM-bound row kernels: LayerNorm (+embedding/residual add, +dropout), row softmax,
// column sums, batched transpose, ReLU backward.  One pass over the data each; rows are
// staged in shared memory so every global element is read once and written once.
#include "common.cuh"

namespace dstag {

constexpr int kRowThreads = 256;
constexpr int kMaxDynSmem = 200 * 1024;

// ======================================================================================
// LayerNorm forward
// ======================================================================================
__global__ void __launch_bounds__(kRowThreads)
ln_fwd_kernel(const float* __restrict__ a, long long a_s0, long long a_s1, int a_div, long long a_sc,
              const float* __restrict__ tab, int tab_mod, long long tab_ld, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, float* __restrict__ y, long long y_ld,
              float* __restrict__ mean, float* __restrict__ rstd, int rows, int L, float drop_p,
              unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
  seed = fold_seed(seed, seed_dev);
  extern __shared__ float sm[];
  float* row = sm;        // L floats; slot c is only ever touched by thread (c % blockDim)
  float* red = sm + L;    // 33 floats
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const float invL = 1.f / (float)L;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* ap = a + (long long)(r / a_div) * a_s0 + (long long)(r % a_div) * a_s1;
    const float* tp = tab ? tab + (long long)(r % tab_mod) * tab_ld : nullptr;
    float s = 0.f;
    for (int c = threadIdx.x; c < L; c += blockDim.x) {
      float v = __ldg(ap + (long long)c * a_sc);
      if (tp) v += __ldg(tp + c);
      row[c] = v;
      s += v;
    }
    const float mu = block_sum(s, red) * invL;
    float q = 0.f;
    for (int c = threadIdx.x; c < L; c += blockDim.x) {
      const float d = row[c] - mu;
      q = fmaf(d, d, q);
    }
    const float rs = rsqrtf(block_sum(q, red) * invL + eps);
    float* yp = y + (long long)r * y_ld;
    for (int c = threadIdx.x; c < L; c += blockDim.x) {
      float o = (row[c] - mu) * rs * __ldg(gamma + c) + __ldg(beta + c);
      if (drop_p > 0.f) o *= dropout_scale((unsigned long long)r * L + c, seed, drop_p, inv_keep);
      yp[c] = o;
    }
    if (threadIdx.x == 0) { mean[r] = mu; rstd[r] = rs; }
  }
}

// ======================================================================================
// LayerNorm backward: dpre = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
// dgamma / dbeta are accumulated per CTA in shared memory, then one atomicAdd per column.
// ======================================================================================
__global__ void __launch_bounds__(kRowThreads)
ln_bwd_kernel(const float* __restrict__ dy, long long dy_ld, const float* __restrict__ a, long long a_s0,
              long long a_s1, int a_div, long long a_sc, const float* __restrict__ tab, int tab_mod,
              long long tab_ld, const float* __restrict__ gamma, const float* __restrict__ mean,
              const float* __restrict__ rstd, float* __restrict__ dpre, long long dpre_ld,
              float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int L, float drop_p,
              unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
  seed = fold_seed(seed, seed_dev);
  extern __shared__ float sm[];
  float* xh = sm;            // xhat
  float* gg = sm + L;        // dy * gamma
  float* ag = sm + 2 * L;    // dgamma partial
  float* ab = sm + 3 * L;    // dbeta partial
  float* red = sm + 4 * L;
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const float invL = 1.f / (float)L;
  for (int c = threadIdx.x; c < L; c += blockDim.x) { ag[c] = 0.f; ab[c] = 0.f; }
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* ap = a + (long long)(r / a_div) * a_s0 + (long long)(r % a_div) * a_s1;
    const float* tp = tab ? tab + (long long)(r % tab_mod) * tab_ld : nullptr;
    const float* dyp = dy + (long long)r * dy_ld;
    const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
    float s1 = 0.f, s2 = 0.f;
    for (int c = threadIdx.x; c < L; c += blockDim.x) {
      float v = __ldg(ap + (long long)c * a_sc);
      if (tp) v += __ldg(tp + c);
      const float x = (v - mu) * rs;
      float d = __ldg(dyp + c);
      if (drop_p > 0.f) d *= dropout_scale((unsigned long long)r * L + c, seed, drop_p, inv_keep);
      const float g = d * __ldg(gamma + c);
      xh[c] = x;
      gg[c] = g;
      ag[c] = fmaf(d, x, ag[c]);
      ab[c] += d;
      s1 += g;
      s2 = fmaf(g, x, s2);
    }
    const float m1 = block_sum(s1, red) * invL;
    const float m2 = block_sum(s2, red) * invL;
    float* op = dpre + (long long)r * dpre_ld;
    for (int c = threadIdx.x; c < L; c += blockDim.x) op[c] = rs * (gg[c] - m1 - xh[c] * m2);
  }
  for (int c = threadIdx.x; c < L; c += blockDim.x) {
    atomicAdd(dgamma + c, ag[c]);
    atomicAdd(dbeta + c, ab[c]);
  }
}

// ======================================================================================
// Column sums.  Two regimes:
//   wide   (cols >= 2048, e.g. the batch reduction of the N x N map gradients): one thread per 4
//          columns, float4 loads, rows walked with 4 independent accumulators; row chunks combine
//          with atomicAdd;
//   narrow (bias gradients: 12..512 columns, 10^5..10^6 rows): a CTA owns a row range, lanes run
//          along the columns (coalesced), 8-row unrolled accumulation, shared-memory combine,
//          one atomicAdd per column per CTA.
// ======================================================================================
__global__ void __launch_bounds__(kRowThreads)
colsum_wide_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, long long cols,
                   long long ld, long long rows_per_chunk) {
  const long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= cols) return;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = min(rows, r0 + rows_per_chunk);
  float4 acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  long long r = r0;
  for (; r + 3 < r1; r += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(in + (r + u) * ld + c));
      acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
    }
  }
  for (; r < r1; ++r) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(in + r * ld + c));
    acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
  }
  const float4 t = make_float4(acc[0].x + acc[1].x + acc[2].x + acc[3].x, acc[0].y + acc[1].y + acc[2].y + acc[3].y,
                               acc[0].z + acc[1].z + acc[2].z + acc[3].z, acc[0].w + acc[1].w + acc[2].w + acc[3].w);
  if (gridDim.y == 1) {
    float4* o = reinterpret_cast<float4*>(out + c);
    const float4 old = *o;
    *o = make_float4(old.x + t.x, old.y + t.y, old.z + t.z, old.w + t.w);
  } else {
    atomicAdd(out + c, t.x); atomicAdd(out + c + 1, t.y); atomicAdd(out + c + 2, t.z); atomicAdd(out + c + 3, t.w);
  }
}

__global__ void __launch_bounds__(kRowThreads)
colsum_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, long long cols,
              long long ld, int lanes_c, long long rows_per_chunk) {
  __shared__ float part[kRowThreads];
  const int lc = threadIdx.x % lanes_c, lr = threadIdx.x / lanes_c, lanes_r = kRowThreads / lanes_c;
  const long long c = (long long)blockIdx.x * lanes_c + lc;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = min(rows, r0 + rows_per_chunk);
  float a[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) a[u] = 0.f;
  if (c < cols) {
    long long r = r0 + lr;
    for (; r + 7LL * lanes_r < r1; r += 8LL * lanes_r) {
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] += __ldg(in + (r + (long long)u * lanes_r) * ld + c);
    }
    for (; r < r1; r += lanes_r) a[0] += __ldg(in + r * ld + c);
  }
  float acc = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  part[threadIdx.x] = acc;
  __syncthreads();
  if (lr == 0 && c < cols) {
    for (int j = 1; j < lanes_r; ++j) acc += part[j * lanes_c + lc];
    atomicAdd(out + c, acc);
  }
}

// ======================================================================================
// Batched transpose through a padded shared tile.
// ======================================================================================
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ in, long long in_ld, float* __restrict__ out, long long out_ld, int R,
                 int C, int r_d, int c_d) {
  // out[b, pc(c), pr(r)] = in[b, r, c] with p(i) = (i % d) * (n / d) + i / d  (d = 1: identity): the index
  // pair (t, f) <-> (f, t) swap between the channels-last block layout and the (b, f, t) row order of
  // the temporal attention is folded into the transpose
  __shared__ float tile[32][33];
  const long long ibase = (long long)blockIdx.z * R * in_ld, obase = (long long)blockIdx.z * C * out_ld;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int r = r0 + ty + j, c = c0 + tx;
    if (r < R && c < C) tile[ty + j][tx] = __ldg(in + ibase + (long long)r * in_ld + c);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int c = c0 + ty + j, r = r0 + tx;
    if (r < R && c < C) {
      const int pc = c_d > 1 ? (c % c_d) * (C / c_d) + c / c_d : c;
      const int pr = r_d > 1 ? (r % r_d) * (R / r_d) + r / r_d : r;
      out[obase + (long long)pc * out_ld + pr] = tile[tx][ty + j];
    }
  }
}

// 64 x 64 tiles moved as float4 on both sides.  The tile's rows are enumerated by OUTPUT position p (r = the input
// row that lands there), so the stores of a row-permuting transpose (r_d > 1: the gradient's way back from the
// (b, (f, t), n) order of the temporal attention) are contiguous too -- walking input rows, consecutive lanes wrote
// R / r_d floats apart.  Needs in_ld, out_ld multiples of 4 with in_ld >= C rounded up to 4, 16-byte aligned bases.
__global__ void __launch_bounds__(256)
transpose64_kernel(const float* __restrict__ in, long long in_ld, float* __restrict__ out, long long out_ld, int R,
                   int C, int r_d, int c_d) {
  __shared__ float tile[64][65];
  const long long ibase = (long long)blockIdx.z * R * in_ld, obase = (long long)blockIdx.z * C * out_ld;
  const int c0 = blockIdx.x * 64, p0 = blockIdx.y * 64;
  const int lo = threadIdx.x & 15, hi = threadIdx.x >> 4;
  const int Rd = R / r_d, Cd = C / c_d;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int pl = hi + 16 * j, p = p0 + pl, c = c0 + 4 * lo;
    if (p < R && c < C) {
      const int r = r_d > 1 ? (p % Rd) * r_d + p / Rd : p;
      const float4 v = __ldg(reinterpret_cast<const float4*>(in + ibase + (long long)r * in_ld + c));
      tile[pl][4 * lo] = v.x; tile[pl][4 * lo + 1] = v.y; tile[pl][4 * lo + 2] = v.z; tile[pl][4 * lo + 3] = v.w;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int cl = hi + 16 * j, c = c0 + cl, p = p0 + 4 * lo;
    if (c < C && p < R) {
      const int pc = c_d > 1 ? (c % c_d) * Cd + c / c_d : c;
      float* dst = out + obase + (long long)pc * out_ld + p;
      const float4 v = make_float4(tile[4 * lo][cl], tile[4 * lo + 1][cl], tile[4 * lo + 2][cl], tile[4 * lo + 3][cl]);
      if (p + 3 < R) {
        *reinterpret_cast<float4*>(dst) = v;
      } else {
        dst[0] = v.x;
        if (p + 1 < R) dst[1] = v.y;
        if (p + 2 < R) dst[2] = v.z;
      }
    }
  }
}

// ======================================================================================
// Row softmax (in place), padded columns zeroed.
// ======================================================================================
__global__ void __launch_bounds__(kRowThreads)
softmax_fwd_kernel(float* __restrict__ x, long long rows, int L, int ld) {
  extern __shared__ float sm[];
  float* row = sm;
  float* red = sm + L;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    float* xp = x + r * ld;
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < L; c += blockDim.x) {
      const float v = xp[c];
      row[c] = v;
      mx = fmaxf(mx, v);
    }
    mx = block_max(mx, red);
    float s = 0.f;
    for (int c = threadIdx.x; c < L; c += blockDim.x) {
      const float e = __expf(row[c] - mx);
      row[c] = e;
      s += e;
    }
    const float inv = 1.f / block_sum(s, red);
    for (int c = threadIdx.x; c < ld; c += blockDim.x) xp[c] = c < L ? row[c] * inv : 0.f;
  }
}

__global__ void __launch_bounds__(kRowThreads)
softmax_bwd_kernel(const float* __restrict__ p, float* __restrict__ dp, const float* __restrict__ mul,
                   long long mul_rows, long long rows, int L, int ld) {
  extern __shared__ float sm[];
  float* pr = sm;
  float* dr = sm + L;
  float* red = sm + 2 * L;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* pp = p + r * ld;
    float* dpp = dp + r * ld;
    const float* mp = mul ? mul + (r % mul_rows) * ld : nullptr;
    float s = 0.f;
    for (int c = threadIdx.x; c < L; c += blockDim.x) {
      const float a = __ldg(pp + c);
      float b = dpp[c];
      if (mp) b *= __ldg(mp + c);
      pr[c] = a;
      dr[c] = b;
      s = fmaf(a, b, s);
    }
    s = block_sum(s, red);
    for (int c = threadIdx.x; c < ld; c += blockDim.x) dpp[c] = c < L ? pr[c] * (dr[c] - s) : 0.f;
  }
}

// Fixed-order sum of nslices partial results (the second stage of a reproducible split-K contraction).
template <bool VEC>
__global__ void __launch_bounds__(256)
sum_slices_kernel(const float* __restrict__ ws, long long slice_stride, int nslices, float* __restrict__ out,
                  long long rows, long long cols, long long out_ld) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (VEC) {
    const long long c4 = cols >> 2, n4 = rows * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const long long r = i / c4, c = (i - r * c4) << 2;
      const float* src = ws + r * cols + c;
      float4 a = __ldg(reinterpret_cast<const float4*>(src));
      for (int s = 1; s < nslices; ++s) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + s * slice_stride));
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
      *reinterpret_cast<float4*>(out + r * out_ld + c) = a;
    }
  } else {
    const long long n = rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const long long r = i / cols, c = i - r * cols;
      float a = __ldg(ws + i);
      for (int s = 1; s < nslices; ++s) a += __ldg(ws + s * slice_stride + i);
      out[r * out_ld + c] = a;
    }
  }
}

// Small results with many slices (a 128 x 32 Gram result summed over 148 per-CTA partials, a weight gradient over 64
// batch slices): one thread per output walked the slices one dependent-latency step at a time on a handful of CTAs
// (25-33 us for a few MB).  Here 8 lanes share one float4 of the result: lane u adds the slices u, u + 8, ... in
// ascending order and the 8 partial sums are combined by a fixed xor tree -- still a run-to-run reproducible order.
__global__ void __launch_bounds__(256)
sum_slices_par_kernel(const float* __restrict__ ws, long long slice_stride, int nslices, float* __restrict__ out,
                      long long rows, long long cols, long long out_ld) {
  const long long c4 = cols >> 2, n4 = rows * c4;
  const int sub = threadIdx.x & 7;
  const unsigned gmask = 0xFFu << (threadIdx.x & 24);          // the 8 lanes of this group (groups may leave the loop apart)
  const long long stride = ((long long)gridDim.x * blockDim.x) >> 3;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3; i < n4; i += stride) {
    const long long r = i / c4, c = (i - r * c4) << 2;
    const float* src = ws + r * cols + c;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int s = sub; s < nslices; s += 8) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + s * slice_stride));
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      a.x += __shfl_xor_sync(gmask, a.x, o);
      a.y += __shfl_xor_sync(gmask, a.y, o);
      a.z += __shfl_xor_sync(gmask, a.z, o);
      a.w += __shfl_xor_sync(gmask, a.w, o);
    }
    if (sub == 0) *reinterpret_cast<float4*>(out + r * out_ld + c) = a;
  }
}

__global__ void relu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                float* __restrict__ dx, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dx[i] = __ldg(y + i) > 0.f ? __ldg(dy + i) : 0.f;
}

// ======================================================================================
// Warp-per-row variants (rows of up to 1024 elements: every shape of the PEMS configurations).
// A warp keeps its row in registers, reduces with shuffles and never touches shared memory or
// a block barrier, so 8 rows per CTA are in flight and the kernels run at HBM speed.
// ======================================================================================

// Warp-per-row LayerNorm for rows of up to 128 * NV floats whose storage is 16-byte aligned (row-padded
// tensors): a lane keeps NV float4 of the row in registers, statistics by warp shuffles, no shared
// memory and no block barriers, so many rows are in flight per SM.  Columns >= L (the padding of the
// last float4) are masked out of the statistics; whatever is written there is never read.
__device__ __forceinline__ void ld4_guard(float (&o)[4], const float* __restrict__ p, int c, int L) {
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = (c + j < L) ? __ldg(p + c + j) : 0.f;
}

// gamma / beta sit in shared memory, not in registers: with them the NV = 7 instantiation (rows of 883 nodes) held
// 202 registers, one 8-warp CTA per SM was resident (12 % of the warp slots) and the kernel ran at 1.8 TB/s
template <int NV>
__global__ void __launch_bounds__(256, NV <= 4 ? 4 : 3)
ln_fwd_warp_kernel(const float* __restrict__ a, long long a_ld, const float* __restrict__ tab, int tab_mod,
                   long long tab_ld, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                   float* __restrict__ y, long long y_ld, float* __restrict__ mean, float* __restrict__ rstd, int rows,
                   int L, float drop_p, unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
  seed = fold_seed(seed, seed_dev);
  const int lane = threadIdx.x & 31;
  const int nw = gridDim.x * (blockDim.x >> 5);
  const float invL = 1.f / (float)L;
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  __shared__ __align__(16) float gs[128 * NV], bs[128 * NV];
  for (int c = threadIdx.x; c < 128 * NV; c += blockDim.x) {
    gs[c] = c < L ? __ldg(gamma + c) : 0.f;
    bs[c] = c < L ? __ldg(beta + c) : 0.f;
  }
  __syncthreads();
  for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += nw) {
    const float* ap = a + (long long)r * a_ld;
    const float* tp = tab ? tab + (long long)(r % tab_mod) * tab_ld : nullptr;
    float v[NV][4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = 4 * (lane + 32 * i);
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < L) {
        x = __ldg(reinterpret_cast<const float4*>(ap + c));
        if (tp) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(tp + c));
          x.x += t4.x; x.y += t4.y; x.z += t4.z; x.w += t4.w;
        }
      }
      v[i][0] = x.x; v[i][1] = (c + 1 < L) ? x.y : 0.f; v[i][2] = (c + 2 < L) ? x.z : 0.f; v[i][3] = (c + 3 < L) ? x.w : 0.f;
      s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    }
    const float mu = warp_sum(s) * invL;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = 4 * (lane + 32 * i);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d = (c + j < L) ? v[i][j] - mu : 0.f;
        q = fmaf(d, d, q);
      }
    }
    const float rs = rsqrtf(warp_sum(q) * invL + eps);
    float* yp = y + (long long)r * y_ld;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (c < L) {
        float o[4], dm[4] = {1.f, 1.f, 1.f, 1.f};
        if (drop_p > 0.f) {
          const unsigned long long e0 = (unsigned long long)r * L + c;
          if ((e0 & 3) == 0) {                                     // warp-uniform: one Philox per float4
            dropout_scale4(e0, seed, drop_p, inv_keep, dm);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) dm[j] = dropout_scale(e0 + j, seed, drop_p, inv_keep);
          }
        }
        const float4 g4 = *reinterpret_cast<const float4*>(gs + c), b4 = *reinterpret_cast<const float4*>(bs + c);
        const float gm[4] = {g4.x, g4.y, g4.z, g4.w}, bt[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fmaf((v[i][j] - mu) * rs, gm[j], bt[j]) * dm[j];
        *reinterpret_cast<float4*>(yp + c) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
  }
}

// 128-thread CTAs: the per-lane d gamma / d beta accumulators make this kernel register-heavy (162 registers at
// NV = 4), and 4-warp CTAs pack 3 per SM (12 warps) where 8-warp CTAs fit once.
constexpr int kLnBwdThreads = 128;
template <int NV>
__global__ void __launch_bounds__(kLnBwdThreads, NV <= 4 ? 4 : 3)
ln_bwd_warp_kernel(const float* __restrict__ dy, long long dy_ld, const float* __restrict__ a, long long a_ld,
                   const float* __restrict__ tab, int tab_mod, long long tab_ld, const float* __restrict__ gamma,
                   const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dpre,
                   long long dpre_ld, float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int L,
                   float drop_p, unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
  seed = fold_seed(seed, seed_dev);
  extern __shared__ __align__(16) float sm[];   // [3][128 * NV]: CTA-level d gamma / d beta accumulators, gamma
  constexpr int W = 128 * NV;
  float* sg = sm;
  float* sb = sm + W;
  float* gs = sm + 2 * W;
  for (int c = threadIdx.x; c < 2 * W; c += blockDim.x) sm[c] = 0.f;
  for (int c = threadIdx.x; c < W; c += blockDim.x) gs[c] = c < L ? __ldg(gamma + c) : 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nw = gridDim.x * (blockDim.x >> 5);
  const float invL = 1.f / (float)L;
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  float ag[NV][4], ab[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { ag[i][j] = 0.f; ab[i][j] = 0.f; }
  }
  for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += nw) {
    const float* ap = a + (long long)r * a_ld;
    const float* tp = tab ? tab + (long long)(r % tab_mod) * tab_ld : nullptr;
    const float* dyp = dy + (long long)r * dy_ld;
    const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
    float xh[NV][4], gg[NV][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = 4 * (lane + 32 * i);
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f), d4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < L) {
        x = __ldg(reinterpret_cast<const float4*>(ap + c));
        if (tp) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(tp + c));
          x.x += t4.x; x.y += t4.y; x.z += t4.z; x.w += t4.w;
        }
        d4 = __ldg(reinterpret_cast<const float4*>(dyp + c));
      }
      const float xv[4] = {x.x, x.y, x.z, x.w};
      const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
      const float4 g4 = *reinterpret_cast<const float4*>(gs + c);
      const float gmv[4] = {g4.x, g4.y, g4.z, g4.w};
      float dm[4] = {1.f, 1.f, 1.f, 1.f};
      if (drop_p > 0.f && c < L) {
        const unsigned long long e0 = (unsigned long long)r * L + c;
        if ((e0 & 3) == 0) {                                       // warp-uniform: one Philox per float4
          dropout_scale4(e0, seed, drop_p, inv_keep, dm);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) dm[j] = dropout_scale(e0 + j, seed, drop_p, inv_keep);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = c + j < L;
        const float d = ok ? dv[j] * dm[j] : 0.f;
        const float xn = ok ? (xv[j] - mu) * rs : 0.f;
        const float g = d * gmv[j];
        xh[i][j] = xn;
        gg[i][j] = g;
        ag[i][j] = fmaf(d, xn, ag[i][j]);
        ab[i][j] += d;
        s1 += g;
        s2 = fmaf(g, xn, s2);
      }
    }
    const float m1 = warp_sum(s1) * invL, m2 = warp_sum(s2) * invL;
    float* op = dpre + (long long)r * dpre_ld;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (c < L)
        *reinterpret_cast<float4*>(op + c) =
            make_float4(rs * (gg[i][0] - m1 - xh[i][0] * m2), rs * (gg[i][1] - m1 - xh[i][1] * m2),
                        rs * (gg[i][2] - m1 - xh[i][2] * m2), rs * (gg[i][3] - m1 - xh[i][3] * m2));
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = 4 * (lane + 32 * i);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (c + j < L) { atomicAdd(sg + c + j, ag[i][j]); atomicAdd(sb + c + j, ab[i][j]); }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < L; c += blockDim.x) {
    atomicAdd(dgamma + c, sg[c]);
    atomicAdd(dbeta + c, sb[c]);
  }
}

// dp <- p (.) (g - sum(p (.) g)), g = dp (.) mul: warp per row, NV float4 per lane (ld <= 128 NV, ld % 4 == 0,
// 16-byte aligned rows; columns [L, ld) of p are zero, so they drop out of the sum and are written as zero)
template <int NV>
__global__ void __launch_bounds__(256)
softmax_bwd_warp_kernel(const float* __restrict__ p, float* __restrict__ dp, const float* __restrict__ mul,
                        long long mul_rows, long long rows, int L, int ld) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += nw) {
    const float4* pp = reinterpret_cast<const float4*>(p + r * ld);
    float4* dpp = reinterpret_cast<float4*>(dp + r * ld);
    const float4* mp = mul ? reinterpret_cast<const float4*>(mul + (r % mul_rows) * ld) : nullptr;
    float4 pv[NV], dv[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = lane + 32 * i;
      pv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      dv[i] = pv[i];
      if (4 * c4 < ld) {
        pv[i] = __ldg(pp + c4);
        dv[i] = dpp[c4];
        if (mp) {
          const float4 m = __ldg(mp + c4);
          dv[i].x *= m.x; dv[i].y *= m.y; dv[i].z *= m.z; dv[i].w *= m.w;
        }
        const int c = 4 * c4;
        if (c + 3 >= L) {                         // ragged end: the padding of dp may hold anything
          if (c >= L) dv[i].x = 0.f;
          if (c + 1 >= L) dv[i].y = 0.f;
          if (c + 2 >= L) dv[i].z = 0.f;
          dv[i].w = 0.f;
        }
        s = fmaf(pv[i].x, dv[i].x, fmaf(pv[i].y, dv[i].y, fmaf(pv[i].z, dv[i].z, fmaf(pv[i].w, dv[i].w, s))));
      }
    }
    s = warp_sum(s);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = lane + 32 * i;
      if (4 * c4 < ld)
        dpp[c4] = make_float4(pv[i].x * (dv[i].x - s), pv[i].y * (dv[i].y - s), pv[i].z * (dv[i].z - s),
                              pv[i].w * (dv[i].w - s));
    }
  }
}

static int ensure_smem(const void* fn, size_t bytes, const char* name) {
  if (bytes > (size_t)kMaxDynSmem) {
    set_error("%s: row too long (%zu bytes of shared memory needed, %d available)", name, bytes,
              kMaxDynSmem);
    return DSTAG_ERR_SMEM;
  }
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute(%zu) failed: %s", name, bytes, cudaGetErrorString(e));
      return (int)e;
    }
  }
  return 0;
}

}  // namespace dstag

using namespace dstag;

extern "C" {

int dstag_layernorm_fwd(const float* a, int64_t a_s0, int64_t a_s1, int32_t a_div, int64_t a_sc,
                        const float* tab, int32_t tab_mod, int64_t tab_ld, const float* gamma, const float* beta,
                        float eps, float* y, int64_t y_ld, float* mean, float* rstd, int32_t rows, int32_t L,
                        float drop_p, uint64_t seed, const uint64_t* seed_dev, void* stream) {
  const unsigned long long* sd = reinterpret_cast<const unsigned long long*>(seed_dev);
  DSTAG_CHECK_ARG(a && gamma && beta && y && mean && rstd, "dstag_layernorm_fwd: null pointer");
  DSTAG_CHECK_ARG(y_ld >= L && (tab == nullptr || tab_ld >= L), "dstag_layernorm_fwd: leading dimension < L");
  DSTAG_CHECK_ARG(rows > 0 && L > 0 && a_div > 0 && (tab == nullptr || tab_mod > 0),
                  "dstag_layernorm_fwd: bad sizes rows=%d L=%d", rows, L);
  DSTAG_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "dstag_layernorm_fwd: bad dropout p");
  // rows in 16-byte aligned storage of up to 1024 floats: warp-per-row kernel
  const int nv = (L + 127) / 128;
  const bool al = a_sc == 1 && a_div == 1 && a_s0 % 4 == 0 && y_ld % 4 == 0 && (tab == nullptr || tab_ld % 4 == 0) &&
                  ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(tab)) & 15) == 0 &&
                  a_s0 >= 4 * ((L + 3) / 4) && y_ld >= 4 * ((L + 3) / 4) && (tab == nullptr || tab_ld >= 4 * ((L + 3) / 4));
  if (al && nv <= 8) {
    const int blocks = (rows + 7) / 8;
    const int grid = blocks < kNumSMs * 8 ? blocks : kNumSMs * 8;
#define DSTAG_LN_FWD(NV_)                                                                                       \
  case NV_:                                                                                                     \
    ln_fwd_warp_kernel<NV_><<<grid, 256, 0, (cudaStream_t)stream>>>(a, a_s0, tab, tab_mod, tab_ld, gamma, beta, eps, y, \
                                                                    y_ld, mean, rstd, rows, L, drop_p, seed, sd); \
    break
    switch (nv) {
      DSTAG_LN_FWD(1); DSTAG_LN_FWD(2); DSTAG_LN_FWD(3); DSTAG_LN_FWD(4);
      DSTAG_LN_FWD(5); DSTAG_LN_FWD(6); DSTAG_LN_FWD(7); DSTAG_LN_FWD(8);
    }
#undef DSTAG_LN_FWD
    DSTAG_RETURN_LAUNCH("dstag_layernorm_fwd(warp)");
  }
  const size_t smem = (size_t)(L + 40) * sizeof(float);
  int rc = ensure_smem((const void*)ln_fwd_kernel, smem, "dstag_layernorm_fwd");
  if (rc) return rc;
  const int grid = rows < kNumSMs * 8 ? rows : kNumSMs * 8;
  ln_fwd_kernel<<<grid, kRowThreads, smem, (cudaStream_t)stream>>>(
      a, a_s0, a_s1, a_div, a_sc, tab, tab_mod, tab_ld, gamma, beta, eps, y, y_ld, mean, rstd, rows, L, drop_p, seed, sd);
  DSTAG_RETURN_LAUNCH("dstag_layernorm_fwd");
}

int dstag_layernorm_bwd(const float* dy, int64_t dy_ld, const float* a, int64_t a_s0, int64_t a_s1, int32_t a_div,
                        int64_t a_sc, const float* tab, int32_t tab_mod, int64_t tab_ld, const float* gamma,
                        const float* mean, const float* rstd, float* dpre, int64_t dpre_ld, float* dgamma,
                        float* dbeta, int32_t rows, int32_t L, float drop_p, uint64_t seed,
                        const uint64_t* seed_dev, void* stream) {
  const unsigned long long* sd = reinterpret_cast<const unsigned long long*>(seed_dev);
  DSTAG_CHECK_ARG(dy_ld >= L && dpre_ld >= L && (tab == nullptr || tab_ld >= L),
                  "dstag_layernorm_bwd: leading dimension < L");
  DSTAG_CHECK_ARG(dy && a && gamma && mean && rstd && dpre && dgamma && dbeta,
                  "dstag_layernorm_bwd: null pointer");
  DSTAG_CHECK_ARG(rows > 0 && L > 0 && a_div > 0 && (tab == nullptr || tab_mod > 0),
                  "dstag_layernorm_bwd: bad sizes rows=%d L=%d", rows, L);
  const int nv = (L + 127) / 128;
  const long long lq = 4 * ((L + 3) / 4);
  const bool al = a_sc == 1 && a_div == 1 && a_s0 % 4 == 0 && dy_ld % 4 == 0 && dpre_ld % 4 == 0 &&
                  (tab == nullptr || tab_ld % 4 == 0) &&
                  ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dpre) |
                    reinterpret_cast<uintptr_t>(tab)) & 15) == 0 &&
                  a_s0 >= lq && dy_ld >= lq && dpre_ld >= lq && (tab == nullptr || tab_ld >= lq);
  if (al && nv <= 8) {
    const int wpc = kLnBwdThreads / 32;
    const int blocks = (rows + wpc - 1) / wpc;
    const int resident = kNumSMs * (nv <= 4 ? 4 : 3);
    const int grid = blocks < resident ? blocks : resident;
    const size_t sm_bytes = (size_t)3 * 128 * nv * sizeof(float);
#define DSTAG_LN_BWD(NV_)                                                                                        \
  case NV_:                                                                                                      \
    ln_bwd_warp_kernel<NV_><<<grid, kLnBwdThreads, sm_bytes, (cudaStream_t)stream>>>(                                      \
        dy, dy_ld, a, a_s0, tab, tab_mod, tab_ld, gamma, mean, rstd, dpre, dpre_ld, dgamma, dbeta, rows, L, drop_p, seed, sd); \
    break
    switch (nv) {
      DSTAG_LN_BWD(1); DSTAG_LN_BWD(2); DSTAG_LN_BWD(3); DSTAG_LN_BWD(4);
      DSTAG_LN_BWD(5); DSTAG_LN_BWD(6); DSTAG_LN_BWD(7); DSTAG_LN_BWD(8);
    }
#undef DSTAG_LN_BWD
    DSTAG_RETURN_LAUNCH("dstag_layernorm_bwd(warp)");
  }
  const size_t smem = (size_t)(4 * L + 40) * sizeof(float);
  int rc = ensure_smem((const void*)ln_bwd_kernel, smem, "dstag_layernorm_bwd");
  if (rc) return rc;
  const int grid = rows < kNumSMs * 4 ? rows : kNumSMs * 4;
  ln_bwd_kernel<<<grid, kRowThreads, smem, (cudaStream_t)stream>>>(
      dy, dy_ld, a, a_s0, a_s1, a_div, a_sc, tab, tab_mod, tab_ld, gamma, mean, rstd, dpre, dpre_ld, dgamma, dbeta,
      rows, L, drop_p, seed, sd);
  DSTAG_RETURN_LAUNCH("dstag_layernorm_bwd");
}

int dstag_sum_slices(const float* ws, int64_t slice_stride, int32_t nslices, float* out, int64_t rows, int64_t cols,
                     int64_t out_ld, void* stream) {
  DSTAG_CHECK_ARG(ws && out && nslices > 0 && rows > 0 && cols > 0 && out_ld >= cols && slice_stride >= rows * cols,
                  "dstag_sum_slices: bad arguments");
  const bool vec = cols % 4 == 0 && out_ld % 4 == 0 && slice_stride % 4 == 0 &&
                   ((reinterpret_cast<uintptr_t>(ws) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  const long long n = vec ? rows * (cols / 4) : rows * cols;
  long long blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  if (vec && nslices >= 16 && n * 8 <= (long long)kNumSMs * 8 * 256 * 4) {
    long long pb = (n * 8 + 255) / 256;
    if (pb > kNumSMs * 8) pb = kNumSMs * 8;
    sum_slices_par_kernel<<<(unsigned)pb, 256, 0, (cudaStream_t)stream>>>(ws, slice_stride, nslices, out, rows, cols, out_ld);
  } else if (vec)
    sum_slices_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ws, slice_stride, nslices, out, rows, cols, out_ld);
  else
    sum_slices_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ws, slice_stride, nslices, out, rows, cols, out_ld);
  DSTAG_RETURN_LAUNCH("dstag_sum_slices");
}

int dstag_colsum(const float* in, float* out, int64_t rows, int64_t cols, int64_t ld, void* stream) {
  DSTAG_CHECK_ARG(in && out && rows > 0 && cols > 0 && ld >= cols, "dstag_colsum: bad arguments");
  const bool wide = cols >= 2048 && cols % 4 == 0 && ld % 4 == 0 &&
                    (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (wide) {
    const long long gx = (cols / 4 + kRowThreads - 1) / kRowThreads;
    long long chunks = (kNumSMs * 8 + gx - 1) / gx;                 // enough CTAs to fill the machine
    const long long max_chunks = (rows + 15) / 16;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    const long long rpc = (rows + chunks - 1) / chunks;
    const dim3 grid((unsigned)gx, (unsigned)((rows + rpc - 1) / rpc));
    colsum_wide_kernel<<<grid, kRowThreads, 0, (cudaStream_t)stream>>>(in, out, rows, cols, ld, rpc);
    DSTAG_RETURN_LAUNCH("dstag_colsum(wide)");
  }
  int lanes_c = 256;
  while (lanes_c > 1 && lanes_c / 2 >= cols) lanes_c /= 2;   // smallest power of two >= cols, <= 256
  const int lanes_r = kRowThreads / lanes_c;
  const long long gx = (cols + lanes_c - 1) / lanes_c;
  long long chunks = (kNumSMs * 8 + gx - 1) / gx;
  const long long max_chunks = (rows + (long long)lanes_r * 16 - 1) / ((long long)lanes_r * 16);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  if (chunks > 65535) chunks = 65535;
  const long long rpc = (rows + chunks - 1) / chunks;
  const dim3 grid((unsigned)gx, (unsigned)((rows + rpc - 1) / rpc));
  colsum_kernel<<<grid, kRowThreads, 0, (cudaStream_t)stream>>>(in, out, rows, cols, ld, lanes_c, rpc);
  DSTAG_RETURN_LAUNCH("dstag_colsum");
}

int dstag_transpose(const float* in, int64_t in_ld, float* out, int64_t out_ld, int32_t batch, int32_t R, int32_t C,
                    int32_t r_d, int32_t c_d, void* stream) {
  DSTAG_CHECK_ARG(in && out && batch > 0 && R > 0 && C > 0 && in_ld >= C && out_ld >= R,
                  "dstag_transpose: bad arguments");
  DSTAG_CHECK_ARG(r_d >= 1 && c_d >= 1 && R % r_d == 0 && C % c_d == 0, "dstag_transpose: bad permutation factors");
  DSTAG_CHECK_ARG(batch <= 65535 && (R + 31) / 32 <= 65535, "dstag_transpose: grid too large");
  if (in_ld % 4 == 0 && out_ld % 4 == 0 && in_ld >= (C + 3) / 4 * 4 &&
      ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 && (R + 63) / 64 <= 65535) {
    const dim3 grid64((C + 63) / 64, (R + 63) / 64, batch);
    transpose64_kernel<<<grid64, 256, 0, (cudaStream_t)stream>>>(in, in_ld, out, out_ld, R, C, r_d, c_d);
    DSTAG_RETURN_LAUNCH("dstag_transpose");
  }
  const dim3 grid((C + 31) / 32, (R + 31) / 32, batch);
  transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, in_ld, out, out_ld, R, C, r_d, c_d);
  DSTAG_RETURN_LAUNCH("dstag_transpose");
}

int dstag_softmax_fwd(float* x, int64_t rows, int32_t L, int32_t ld, void* stream) {
  DSTAG_CHECK_ARG(x && rows > 0 && L > 0 && ld >= L, "dstag_softmax_fwd: bad arguments");
  const size_t smem = (size_t)(L + 40) * sizeof(float);
  int rc = ensure_smem((const void*)softmax_fwd_kernel, smem, "dstag_softmax_fwd");
  if (rc) return rc;
  const int grid = (int)(rows < kNumSMs * 8 ? rows : kNumSMs * 8);
  softmax_fwd_kernel<<<grid, kRowThreads, smem, (cudaStream_t)stream>>>(x, rows, L, ld);
  DSTAG_RETURN_LAUNCH("dstag_softmax_fwd");
}

int dstag_softmax_bwd(const float* p, float* dp, const float* mul, int64_t mul_rows, int64_t rows, int32_t L,
                      int32_t ld, void* stream) {
  DSTAG_CHECK_ARG(p && dp && rows > 0 && L > 0 && ld >= L && (mul == nullptr || mul_rows > 0),
                  "dstag_softmax_bwd: bad arguments");
  const int nv = (ld + 127) / 128;
  const bool al = ld % 4 == 0 && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(dp) |
                                   reinterpret_cast<uintptr_t>(mul)) & 15) == 0;
  if (al && nv <= 8) {
    const long long blocks = (rows + 7) / 8;
    const int grid = (int)(blocks < kNumSMs * 8 ? blocks : kNumSMs * 8);
#define DSTAG_SMB(NV_)                                                                                               \
  case NV_:                                                                                                          \
    softmax_bwd_warp_kernel<NV_><<<grid, 256, 0, (cudaStream_t)stream>>>(p, dp, mul, mul_rows, rows, L, ld);         \
    break
    switch (nv) {
      DSTAG_SMB(1); DSTAG_SMB(2); DSTAG_SMB(3); DSTAG_SMB(4); DSTAG_SMB(5); DSTAG_SMB(6); DSTAG_SMB(7); DSTAG_SMB(8);
    }
#undef DSTAG_SMB
    DSTAG_RETURN_LAUNCH("dstag_softmax_bwd(warp)");
  }
  const size_t smem = (size_t)(2 * L + 40) * sizeof(float);
  int rc = ensure_smem((const void*)softmax_bwd_kernel, smem, "dstag_softmax_bwd");
  if (rc) return rc;
  const int grid = (int)(rows < kNumSMs * 8 ? rows : kNumSMs * 8);
  softmax_bwd_kernel<<<grid, kRowThreads, smem, (cudaStream_t)stream>>>(p, dp, mul, mul_rows, rows, L, ld);
  DSTAG_RETURN_LAUNCH("dstag_softmax_bwd");
}

int dstag_relu_bwd(const float* y, const float* dy, float* dx, int64_t n, void* stream) {
  DSTAG_CHECK_ARG(y && dy && dx && n > 0, "dstag_relu_bwd: bad arguments");
  long long blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  relu_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(y, dy, dx, n);
  DSTAG_RETURN_LAUNCH("dstag_relu_bwd");
}

}  // extern "C"
