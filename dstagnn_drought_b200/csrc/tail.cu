// Block tail (model/DSTAGNN_my.py:194-196, 243-252): GTU gate, fcmy, dropout -> relu(tc [+ g]) ->
// relu(r + .) -> LayerNorm over the channels.
//
// EVERY tensor here is channels-last: (X = B*N, time, C) with the C channels innermost -- the graph
// convolution output g, the GTU conv outputs u, the concatenated gate outputs cat (X, 3T-12, C), the
// residual r, the block output and all of their gradients.  A warp owns the C channels of one
// (node, time) row or of one node, so every global access is one coalesced line and the LayerNorm over
// the channels is a warp-shuffle reduction; nothing is transposed or staged in shared memory.
// The block returns its output to the caller as a permuted VIEW (B, N, C, T) of this layout.
#include "common.cuh"

namespace dstag {

// exp-based gate functions on the fast-math units (abs error ~1e-7: far inside the parity budget;
// the libm tanhf/expf versions made the gate kernels instruction-bound, not HBM-bound)
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

// One warp per node x, lanes over the channels (c, c + 32, ...): every access is a coalesced line and
// there is no index arithmetic beyond adds.
__global__ void __launch_bounds__(256)
gtu_gate_fwd_kernel(const float* __restrict__ u, float* __restrict__ out, long long X, int C, int Tp, int S, int off) {
  // u: (X, Tp, 2C) conv output;  out[x, off + t, c] = tanh(u[x,t,c]) * sigmoid(u[x,t,C+c]), out (X, S, C)
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long x = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); x < X; x += nw) {
    const float* up = u + x * Tp * 2 * C;
    float* op = out + (x * S + off) * C;
    for (int t = 0; t < Tp; ++t, up += 2 * C, op += C)
      for (int c = lane; c < C; c += 32) op[c] = tanhf_(__ldg(up + c)) * sigmoidf_(__ldg(up + C + c));
  }
}

constexpr int kGateNC = 8;       // channels per lane kept in registers for the bias gradient (C <= 256)
// FROM_SV: `u` holds the saved gate factors [tanh(p) | sigmoid(q)] of the fused forward kernel (csrc/gtu_tc.cu)
// instead of the pre-activations.
template <bool FROM_SV>
__global__ void __launch_bounds__(256)
gtu_gate_bwd_kernel(const float* __restrict__ u, const float* __restrict__ dout, float* __restrict__ dup,
                    float* __restrict__ db, long long X, int C, int Tp, int pad, int S, int off) {
  // dup: (X, Tp + 2 pad, 2C) gradient of the conv output, zero in the pad rows.
  // db (2C, optional): the conv bias gradient = column sums of dup, kept per lane in registers, folded
  // over the CTA's warps in shared memory, one atomic per channel and CTA.
  extern __shared__ float red[];           // [8][2C] when db
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  float sp[kGateNC], sq[kGateNC];
#pragma unroll
  for (int i = 0; i < kGateNC; ++i) { sp[i] = 0.f; sq[i] = 0.f; }
  for (long long x = (long long)blockIdx.x * (blockDim.x >> 5) + wid; x < X; x += nw) {
    const float* up = u + x * Tp * 2 * C;
    const float* dp = dout + (x * S + off) * C;
    float* o = dup + x * (Tp + 2 * pad) * 2 * C;
    for (int i = lane; i < pad * 2 * C; i += 32) o[i] = 0.f;
    o += pad * 2 * C;
    for (int t = 0; t < Tp; ++t, up += 2 * C, dp += C, o += 2 * C) {
#pragma unroll
      for (int i = 0; i < kGateNC; ++i) {
        const int c = lane + 32 * i;
        if (c < C) {
          const float th = FROM_SV ? __ldg(up + c) : tanhf_(__ldg(up + c));
          const float sg = FROM_SV ? __ldg(up + C + c) : sigmoidf_(__ldg(up + C + c));
          const float d = __ldg(dp + c);
          const float rp = d * sg * (1.f - th * th), rq = d * th * sg * (1.f - sg);
          o[c] = rp;
          o[C + c] = rq;
          sp[i] += rp;
          sq[i] += rq;
        }
      }
    }
    for (int i = lane; i < pad * 2 * C; i += 32) o[i] = 0.f;
  }
  if (db) {
#pragma unroll
    for (int i = 0; i < kGateNC; ++i) {
      const int c = lane + 32 * i;
      if (c < C) { red[wid * 2 * C + c] = sp[i]; red[wid * 2 * C + C + c] = sq[i]; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
      float a = 0.f;
      for (int w = 0; w < 8; ++w) a += red[w * 2 * C + c];
      atomicAdd(db + c, a);
    }
  }
}

// ---------------------------------------------------------------------------------------
// out[m, n] += sum_x sum_c a[x, m, c] * b[x, n, c]   (a: X x m x C, b: X x n x C, channels-last) and,
// optionally, asum[m] += sum_x sum_c a[x, m, c]: the fcmy weight and bias gradients
// (dW[t, s] = sum dtc[x,t,c] cat[x,s,c]; T x (3T-12) outputs reduced over B*N*C).  A tensor-core tile
// would be 98 % padding; this is a streaming reduction: a CTA stages slabs of nodes in shared memory,
// thread (phase, m, n quad) keeps 4 accumulators over float4 channel groups, the phases are folded
// through shared memory at the end.
// ---------------------------------------------------------------------------------------
constexpr int kOsNodes = 8;            // nodes per slab
// MR rows of `a` per thread: with one row a thread issues 5 shared-memory float4 loads per 16 FMAs and the kernel is
// bound by the shared-memory pipe (ncu: issue active 68 %, 2.1 TB/s); three rows share the four `b` loads (7 per 48).
template <int MR>
__global__ void __launch_bounds__(256)
outer_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                 float* __restrict__ asum, long long X, int m, int n, int C, int phases) {
  extern __shared__ __align__(16) float sm[];
  const int Cp = C + 4;                    // padded row: rows that differ in m / n hit different banks
  float* as = sm;                          // kOsNodes x m x Cp
  float* bs = sm + kOsNodes * m * Cp;      // kOsNodes x n x Cp
  const int nq = n >> 2, per = (m / MR) * nq;     // threads per phase: (row group, output quad)
  const int ph = threadIdx.x / per, id = threadIdx.x - ph * per;
  const bool active = ph < phases;
  const int mi = (id / nq) * MR, q = id - (id / nq) * nq;
  const int c4n = C >> 2;                  // float4 groups per row
  float acc[MR][4];
  float accs[MR];
#pragma unroll
  for (int r = 0; r < MR; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; accs[r] = 0.f; }
  const long long nslab = (X + kOsNodes - 1) / kOsNodes;
  for (long long sl = blockIdx.x; sl < nslab; sl += gridDim.x) {
    const long long x0 = sl * kOsNodes;
    const int nx = (int)min((long long)kOsNodes, X - x0);
    __syncthreads();
    {
      const float4* a4 = reinterpret_cast<const float4*>(a + x0 * m * C);
      const float4* b4 = reinterpret_cast<const float4*>(b + x0 * n * C);
      const int na = nx * m * c4n, nb = nx * n * c4n;
      for (int i = threadIdx.x; i < na; i += 256)
        *reinterpret_cast<float4*>(as + (i / c4n) * Cp + 4 * (i % c4n)) = __ldg(a4 + i);
      for (int i = threadIdx.x; i < nb; i += 256)
        *reinterpret_cast<float4*>(bs + (i / c4n) * Cp + 4 * (i % c4n)) = __ldg(b4 + i);
    }
    __syncthreads();
    if (active) {
      // phase ph takes the (node, channel group) pairs ph, ph + phases, ...
      for (int e = ph; e < nx * c4n; e += phases) {
        const int xl = e / c4n, cg = e - xl * c4n;
        float4 av[MR];
#pragma unroll
        for (int r = 0; r < MR; ++r) av[r] = *reinterpret_cast<const float4*>(as + (xl * m + mi + r) * Cp + 4 * cg);
        // the thread's four outputs are n = q, q + nq, q + 2 nq, q + 3 nq: for a fixed j the lanes of a
        // quarter warp read consecutive padded rows (stride 36 floats = 4 banks): conflict-free float4 loads
        const float* bp = bs + (xl * n + q) * Cp + 4 * cg;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 bv = *reinterpret_cast<const float4*>(bp + j * nq * Cp);
#pragma unroll
          for (int r = 0; r < MR; ++r)
            acc[r][j] = fmaf(av[r].x, bv.x, fmaf(av[r].y, bv.y, fmaf(av[r].z, bv.z, fmaf(av[r].w, bv.w, acc[r][j]))));
        }
        if (q == 0) {
#pragma unroll
          for (int r = 0; r < MR; ++r) accs[r] += (av[r].x + av[r].y) + (av[r].z + av[r].w);
        }
      }
    }
  }
  __syncthreads();
  constexpr int RW = 5 * MR;               // floats per thread in the fold buffer
  float* red = sm;                         // [256][RW]
  if (active) {
#pragma unroll
    for (int r = 0; r < MR; ++r) {
#pragma unroll
      for (int j = 0; j < 4; ++j) red[threadIdx.x * RW + 5 * r + j] = acc[r][j];
      red[threadIdx.x * RW + 5 * r + 4] = accs[r];
    }
  }
  __syncthreads();
  if (threadIdx.x < per) {
    float t[RW];
#pragma unroll
    for (int j = 0; j < RW; ++j) t[j] = 0.f;
    for (int p2 = 0; p2 < phases; ++p2)
#pragma unroll
      for (int j = 0; j < RW; ++j) t[j] += red[(p2 * per + threadIdx.x) * RW + j];
#pragma unroll
    for (int r = 0; r < MR; ++r) {
      float* dst = out + (mi + r) * n + q;
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(dst + j * nq, t[5 * r + j]);
      if (asum && q == 0) atomicAdd(asum + mi + r, t[5 * r + 4]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// unfused tail (any T, C <= 256): one warp per (node, time) row, lane c (+32, +64, ...) = channel.
// dropout element index = (x * C + c) * T + t, i.e. the index of the reference's (B, N, C, T) tensor,
// so the fused and the unfused kernels drop the same elements for the same seed.
// ---------------------------------------------------------------------------------------
constexpr int kTailNC = 8;                 // channel registers per lane (C <= 256)

__global__ void __launch_bounds__(256)
tail_rows_fwd_kernel(const float* __restrict__ tc, const float* __restrict__ g, const float* __restrict__ r,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                     float* __restrict__ out, float* __restrict__ mean, float* __restrict__ rstd, long long rows,
                     int C, int T, float drop_p, unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
  seed = fold_seed(seed, seed_dev);
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const float invC = 1.f / (float)C;
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += nw) {
    const long long x = row / T;
    const int t = (int)(row - x * T);
    float v[kTailNC];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kTailNC; ++i) {
      const int c = lane + 32 * i;
      v[i] = 0.f;
      if (c < C) {
        float a = __ldg(tc + row * C + c);
        if (drop_p > 0.f) a *= dropout_scale((unsigned long long)((x * C + c) * T + t), seed, drop_p, inv_keep);
        if (g) a += __ldg(g + row * C + c);
        v[i] = fmaxf(__ldg(r + row * C + c) + fmaxf(a, 0.f), 0.f);
        s += v[i];
      }
    }
    const float mu = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kTailNC; ++i) {
      const float d = (lane + 32 * i < C) ? v[i] - mu : 0.f;
      q = fmaf(d, d, q);
    }
    const float rs = rsqrtf(warp_sum(q) * invC + eps);
#pragma unroll
    for (int i = 0; i < kTailNC; ++i) {
      const int c = lane + 32 * i;
      if (c < C) out[row * C + c] = (v[i] - mu) * rs * __ldg(gamma + c) + __ldg(beta + c);
    }
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
  }
}

__global__ void __launch_bounds__(256)
tail_rows_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ tc, const float* __restrict__ g,
                     const float* __restrict__ r, const float* __restrict__ gamma, const float* __restrict__ mean,
                     const float* __restrict__ rstd, float* __restrict__ dtc, float* __restrict__ dg,
                     float* __restrict__ dr, float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows,
                     int C, int T, float drop_p, unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
  seed = fold_seed(seed, seed_dev);
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const float invC = 1.f / (float)C;
  float ag[kTailNC], ab[kTailNC];
#pragma unroll
  for (int i = 0; i < kTailNC; ++i) { ag[i] = 0.f; ab[i] = 0.f; }
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += nw) {
    const long long x = row / T;
    const int t = (int)(row - x * T);
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
    float av[kTailNC], s2v[kTailNC], dsc[kTailNC], gy[kTailNC], xh[kTailNC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kTailNC; ++i) {
      const int c = lane + 32 * i;
      av[i] = s2v[i] = gy[i] = xh[i] = 0.f;
      dsc[i] = 1.f;
      if (c < C) {
        float a = __ldg(tc + row * C + c);
        if (drop_p > 0.f) {
          dsc[i] = dropout_scale((unsigned long long)((x * C + c) * T + t), seed, drop_p, inv_keep);
          a *= dsc[i];
        }
        if (g) a += __ldg(g + row * C + c);
        const float sv = __ldg(r + row * C + c) + fmaxf(a, 0.f);
        const float vv = fmaxf(sv, 0.f);
        const float d = __ldg(dout + row * C + c);
        av[i] = a;
        s2v[i] = sv;
        xh[i] = (vv - mu) * rs;
        gy[i] = d * __ldg(gamma + c);
        ag[i] = fmaf(d, xh[i], ag[i]);
        ab[i] += d;
        s1 += gy[i];
        s2 = fmaf(gy[i], xh[i], s2);
      }
    }
    const float m1 = warp_sum(s1) * invC, m2 = warp_sum(s2) * invC;
#pragma unroll
    for (int i = 0; i < kTailNC; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        const float ds2 = s2v[i] > 0.f ? rs * (gy[i] - m1 - xh[i] * m2) : 0.f;
        const float ds1 = av[i] > 0.f ? ds2 : 0.f;
        dr[row * C + c] = ds2;
        if (dg) dg[row * C + c] = ds1;
        dtc[row * C + c] = ds1 * dsc[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kTailNC; ++i) {
    const int c = lane + 32 * i;
    if (c < C) { atomicAdd(dgamma + c, ag[i]); atomicAdd(dbeta + c, ab[i]); }
  }
}

static int opt_in_smem(const void* fn, size_t bytes, const char* name) {
  if (bytes <= 48 * 1024) return 0;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%zu) failed: %s", name, bytes, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

// C = 32 (every shipped configuration; the unfused tail serves the long series, T = 144): eight lanes own one row as
// float4, a warp takes 4 rows per step and two steps per iteration -- 1.5 KB of loads in flight per warp instead of
// 384 bytes with one lane per channel (ncu before: 1.0 TB/s).  Same arithmetic and dropout indices as the generic kernels.
__device__ __forceinline__ float quad8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ float4 ld4_or_zero(const float* p, bool ok) {
  return ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(256)
tail_rows32_fwd_kernel(const float* __restrict__ tc, const float* __restrict__ g, const float* __restrict__ r,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                       float* __restrict__ out, float* __restrict__ mean, float* __restrict__ rstd, long long rows, int T,
                       float drop_p, unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
  constexpr int C = 32, U = 2;
  seed = fold_seed(seed, seed_dev);
  const int lane = threadIdx.x & 31, q = lane & 7, sub = lane >> 3;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + 4 * q));
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta + 4 * q));
  for (long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (4 * U); row0 < rows;
       row0 += nw * (4 * U)) {
    float4 a[U], gg[U], rr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = row0 + 4 * u + sub;
      const bool ok = row < rows;
      a[u] = ld4_or_zero(tc + row * C + 4 * q, ok);
      gg[u] = ld4_or_zero(g + row * C + 4 * q, ok && g);
      rr[u] = ld4_or_zero(r + row * C + 4 * q, ok);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = row0 + 4 * u + sub;
      const bool ok = row < rows;
      float av[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
      const float gv[4] = {gg[u].x, gg[u].y, gg[u].z, gg[u].w};
      const float rv[4] = {rr[u].x, rr[u].y, rr[u].z, rr[u].w};
      float v[4];
      if (drop_p > 0.f && ok) {
        const long long x = row / T;
        const int t = (int)(row - x * T);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          av[j] *= dropout_scale((unsigned long long)((x * C + 4 * q + j) * T + t), seed, drop_p, inv_keep);
      }
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = fmaxf(rv[j] + fmaxf(av[j] + gv[j], 0.f), 0.f);
        s += v[j];
      }
      const float mu = quad8_sum(s) * (1.f / C);
      float qq = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) qq = fmaf(v[j] - mu, v[j] - mu, qq);
      const float rs = rsqrtf(quad8_sum(qq) * (1.f / C) + eps);
      if (ok) {
        *reinterpret_cast<float4*>(out + row * C + 4 * q) =
            make_float4((v[0] - mu) * rs * ga.x + be.x, (v[1] - mu) * rs * ga.y + be.y, (v[2] - mu) * rs * ga.z + be.z,
                        (v[3] - mu) * rs * ga.w + be.w);
        if (q == 0) { mean[row] = mu; rstd[row] = rs; }
      }
    }
  }
}

__global__ void __launch_bounds__(256)
tail_rows32_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ tc, const float* __restrict__ g,
                       const float* __restrict__ r, const float* __restrict__ gamma, const float* __restrict__ mean,
                       const float* __restrict__ rstd, float* __restrict__ dtc, float* __restrict__ dg,
                       float* __restrict__ dr, float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows,
                       int T, float drop_p, unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
  constexpr int C = 32, U = 2;
  seed = fold_seed(seed, seed_dev);
  const int lane = threadIdx.x & 31, q = lane & 7, sub = lane >> 3;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const float4 ga4 = __ldg(reinterpret_cast<const float4*>(gamma + 4 * q));
  const float ga[4] = {ga4.x, ga4.y, ga4.z, ga4.w};
  float ag[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (4 * U); row0 < rows;
       row0 += nw * (4 * U)) {
    float4 a[U], gg[U], rr[U], dd[U];
    float mus[U], rss[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = row0 + 4 * u + sub;
      const bool ok = row < rows;
      a[u] = ld4_or_zero(tc + row * C + 4 * q, ok);
      gg[u] = ld4_or_zero(g + row * C + 4 * q, ok && g);
      rr[u] = ld4_or_zero(r + row * C + 4 * q, ok);
      dd[u] = ld4_or_zero(dout + row * C + 4 * q, ok);
      mus[u] = ok ? __ldg(mean + row) : 0.f;
      rss[u] = ok ? __ldg(rstd + row) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = row0 + 4 * u + sub;
      const bool ok = row < rows;
      float av[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
      const float gv[4] = {gg[u].x, gg[u].y, gg[u].z, gg[u].w};
      const float rv[4] = {rr[u].x, rr[u].y, rr[u].z, rr[u].w};
      const float dv[4] = {dd[u].x, dd[u].y, dd[u].z, dd[u].w};
      float dsc[4] = {1.f, 1.f, 1.f, 1.f};
      if (drop_p > 0.f && ok) {
        const long long x = row / T;
        const int t = (int)(row - x * T);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dsc[j] = dropout_scale((unsigned long long)((x * C + 4 * q + j) * T + t), seed, drop_p, inv_keep);
          av[j] *= dsc[j];
        }
      }
      const float mu = mus[u], rs = rss[u];
      float s2v[4], xh[4], gy[4];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        av[j] += gv[j];
        s2v[j] = rv[j] + fmaxf(av[j], 0.f);
        xh[j] = ok ? (fmaxf(s2v[j], 0.f) - mu) * rs : 0.f;
        gy[j] = dv[j] * ga[j];
        ag[j] = fmaf(dv[j], xh[j], ag[j]);
        ab[j] += dv[j];
        s1 += gy[j];
        s2 = fmaf(gy[j], xh[j], s2);
      }
      const float m1 = quad8_sum(s1) * (1.f / C), m2 = quad8_sum(s2) * (1.f / C);
      if (ok) {
        float ds2[4], ds1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ds2[j] = s2v[j] > 0.f ? rs * (gy[j] - m1 - xh[j] * m2) : 0.f;
          ds1[j] = av[j] > 0.f ? ds2[j] : 0.f;
        }
        *reinterpret_cast<float4*>(dr + row * C + 4 * q) = make_float4(ds2[0], ds2[1], ds2[2], ds2[3]);
        if (dg) *reinterpret_cast<float4*>(dg + row * C + 4 * q) = make_float4(ds1[0], ds1[1], ds1[2], ds1[3]);
        *reinterpret_cast<float4*>(dtc + row * C + 4 * q) =
            make_float4(ds1[0] * dsc[0], ds1[1] * dsc[1], ds1[2] * dsc[2], ds1[3] * dsc[3]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    ag[j] += __shfl_xor_sync(0xffffffffu, ag[j], 8);
    ag[j] += __shfl_xor_sync(0xffffffffu, ag[j], 16);
    ab[j] += __shfl_xor_sync(0xffffffffu, ab[j], 8);
    ab[j] += __shfl_xor_sync(0xffffffffu, ab[j], 16);
  }
  if (sub == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { atomicAdd(dgamma + 4 * q + j, ag[j]); atomicAdd(dbeta + 4 * q + j, ab[j]); }
  }
}

static bool aligned16(const void* a, const void* b = nullptr, const void* c = nullptr, const void* d = nullptr,
                      const void* e = nullptr, const void* f = nullptr, const void* h = nullptr) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(f) |
           reinterpret_cast<uintptr_t>(h)) & 15) == 0;
}

}  // namespace dstag

using namespace dstag;

extern "C" {

int dstag_gtu_gate_fwd(const float* u, float* out, int64_t X, int32_t C, int32_t Tp, int32_t S, int32_t off,
                       void* stream) {
  DSTAG_CHECK_ARG(u && out && X > 0 && C > 0 && Tp > 0 && off >= 0 && off + Tp <= S,
                  "dstag_gtu_gate_fwd: bad arguments");
  long long blocks = (X + 7) / 8;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  gtu_gate_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(u, out, X, C, Tp, S, off);
  DSTAG_RETURN_LAUNCH("dstag_gtu_gate_fwd");
}

static int gate_bwd(const char* name, bool from_sv, const float* u, const float* dout, float* dup, float* db, int64_t X,
                    int32_t C, int32_t Tp, int32_t pad, int32_t S, int32_t off, void* stream) {
  DSTAG_CHECK_ARG(u && dout && dup && X > 0 && C > 0 && Tp > 0 && pad >= 0 && off >= 0 && off + Tp <= S,
                  "%s: bad arguments", name);
  DSTAG_CHECK_ARG(C <= 32 * kGateNC, "%s: at most %d channels", name, 32 * kGateNC);
  long long blocks = (X + 7) / 8;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  const size_t smem = db ? (size_t)8 * 2 * C * sizeof(float) : 0;
  if (from_sv)
    gtu_gate_bwd_kernel<true><<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(u, dout, dup, db, X, C, Tp, pad, S, off);
  else
    gtu_gate_bwd_kernel<false><<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(u, dout, dup, db, X, C, Tp, pad, S, off);
  DSTAG_RETURN_LAUNCH(name);
}

int dstag_gtu_gate_bwd(const float* u, const float* dout, float* dup, float* db, int64_t X, int32_t C, int32_t Tp,
                       int32_t pad, int32_t S, int32_t off, void* stream) {
  return gate_bwd("dstag_gtu_gate_bwd", false, u, dout, dup, db, X, C, Tp, pad, S, off, stream);
}

int dstag_gtu_gate_bwd_sv(const float* sv, const float* dout, float* dup, float* db, int64_t X, int32_t C, int32_t Tp,
                          int32_t pad, int32_t S, int32_t off, void* stream) {
  return gate_bwd("dstag_gtu_gate_bwd_sv", true, sv, dout, dup, db, X, C, Tp, pad, S, off, stream);
}

int dstag_outer_sum_supported(int32_t m, int32_t n, int32_t C) {
  return (m > 0 && n > 0 && C > 0 && n % 4 == 0 && C % 4 == 0 && m * (n / 4) <= 256 &&
          (size_t)kOsNodes * (m + n) * (C + 4) * 4 <= 160 * 1024 && (size_t)kOsNodes * (m + n) * C >= 256 * 5)
             ? 1 : 0;
}

int dstag_outer_sum(const float* a, const float* b, float* out, float* asum, int64_t X, int32_t m, int32_t n,
                    int32_t C, void* stream) {
  DSTAG_CHECK_ARG(a && b && out && X > 0, "dstag_outer_sum: bad arguments");
  DSTAG_CHECK_ARG(dstag_outer_sum_supported(m, n, C), "dstag_outer_sum: unsupported shape m=%d n=%d C=%d", m, n, C);
  DSTAG_CHECK_ARG(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0,
                  "dstag_outer_sum: operands must be 16-byte aligned");
  // three rows per thread when the fold buffer (15 floats per thread) fits in the slab's shared memory
  const bool three = m % 3 == 0 && (size_t)kOsNodes * (m + n) * (C + 4) >= 256 * 15;
  const int per = (three ? m / 3 : m) * (n / 4);
  const int phases = 256 / per;
  const size_t smem = (size_t)kOsNodes * (m + n) * (C + 4) * sizeof(float);
  int rc = three ? opt_in_smem((const void*)outer_sum_kernel<3>, smem, "dstag_outer_sum")
                 : opt_in_smem((const void*)outer_sum_kernel<1>, smem, "dstag_outer_sum");
  if (rc) return rc;
  const long long nslab = (X + kOsNodes - 1) / kOsNodes;
  const int grid = (int)(nslab < kNumSMs * 4 ? nslab : kNumSMs * 4);
  if (three)
    outer_sum_kernel<3><<<grid, 256, smem, (cudaStream_t)stream>>>(a, b, out, asum, X, m, n, C, phases);
  else
    outer_sum_kernel<1><<<grid, 256, smem, (cudaStream_t)stream>>>(a, b, out, asum, X, m, n, C, phases);
  DSTAG_RETURN_LAUNCH("dstag_outer_sum");
}

int dstag_tail_fwd(const float* tc, const float* g, const float* r, const float* gamma,
                   const float* beta, float eps, float* out, float* mean, float* rstd, int64_t X,
                   int32_t C, int32_t T, float drop_p, uint64_t seed, const uint64_t* seed_dev, void* stream) {
  const unsigned long long* sd = reinterpret_cast<const unsigned long long*>(seed_dev);
  DSTAG_CHECK_ARG(tc && r && gamma && beta && out && mean && rstd, "dstag_tail_fwd: null pointer");
  DSTAG_CHECK_ARG(X > 0 && C > 0 && T > 0 && drop_p >= 0.f && drop_p < 1.f, "dstag_tail_fwd: bad sizes");
  DSTAG_CHECK_ARG(C <= 32 * kTailNC, "dstag_tail_fwd: at most %d channels", 32 * kTailNC);
  const long long rows = (long long)X * T;
  if (C == 32 && aligned16(tc, g, r, gamma, beta, out)) {
    const long long blocks = (rows + 63) / 64;               // 8 warps x 8 rows per iteration
    const int grid = (int)(blocks < kNumSMs * 8 ? blocks : kNumSMs * 8);
    tail_rows32_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tc, g, r, gamma, beta, eps, out, mean, rstd, rows, T,
                                                                   drop_p, seed, sd);
    DSTAG_RETURN_LAUNCH("dstag_tail_fwd");
  }
  const long long blocks = (rows + 7) / 8;
  const int grid = (int)(blocks < kNumSMs * 8 ? blocks : kNumSMs * 8);
  tail_rows_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tc, g, r, gamma, beta, eps, out, mean, rstd, rows, C, T,
                                                               drop_p, seed, sd);
  DSTAG_RETURN_LAUNCH("dstag_tail_fwd");
}

int dstag_tail_bwd(const float* dout, const float* tc, const float* g, const float* r,
                   const float* gamma, const float* mean, const float* rstd, float* dtc, float* dg,
                   float* dr, float* dgamma, float* dbeta, int64_t X, int32_t C, int32_t T,
                   float drop_p, uint64_t seed, const uint64_t* seed_dev, void* stream) {
  const unsigned long long* sd = reinterpret_cast<const unsigned long long*>(seed_dev);
  DSTAG_CHECK_ARG(dout && tc && r && gamma && mean && rstd && dtc && dr && dgamma && dbeta,
                  "dstag_tail_bwd: null pointer");
  DSTAG_CHECK_ARG((g == nullptr) == (dg == nullptr), "dstag_tail_bwd: g and dg must both be given or both be null");
  DSTAG_CHECK_ARG(X > 0 && C > 0 && T > 0 && drop_p >= 0.f && drop_p < 1.f, "dstag_tail_bwd: bad sizes");
  DSTAG_CHECK_ARG(C <= 32 * kTailNC, "dstag_tail_bwd: at most %d channels", 32 * kTailNC);
  const long long rows = (long long)X * T;
  if (C == 32 && aligned16(dout, tc, g, r, gamma, dtc, dr) && aligned16(dg)) {
    const long long blocks = (rows + 63) / 64;
    const int grid = (int)(blocks < kNumSMs * 4 ? blocks : kNumSMs * 4);
    tail_rows32_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dout, tc, g, r, gamma, mean, rstd, dtc, dg, dr, dgamma,
                                                                   dbeta, rows, T, drop_p, seed, sd);
    DSTAG_RETURN_LAUNCH("dstag_tail_bwd");
  }
  const long long blocks = (rows + 7) / 8;
  const int grid = (int)(blocks < kNumSMs * 4 ? blocks : kNumSMs * 4);
  tail_rows_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dout, tc, g, r, gamma, mean, rstd, dtc, dg, dr, dgamma,
                                                               dbeta, rows, C, T, drop_p, seed, sd);
  DSTAG_RETURN_LAUNCH("dstag_tail_bwd");
}

}  // extern "C"

// =======================================================================================
// Fused fcmy + tail for the PEMS shape (C = 32 channels = one lane each, T = 12): a warp owns one node.
// Lane c reads its channel of the concatenated GTU outputs cat[x, s, c] (3T-12 coalesced lines), evaluates
// the Linear over the concatenated time axis in registers and the LayerNorm over the 32 channels is a
// warp-shuffle reduction per time step.  tc never touches HBM; the backward emits d cat and d tc (for the
// weight gradient) directly.
// The fcmy weights are staged per CTA in shared memory (288 + 12 floats, read back as warp-uniform
// float4 broadcasts: one LDS.128 per four FMAs).  They used to live in __constant__ symbols re-uploaded
// per call, which two streams running different blocks could race on; with per-CTA staging the entry
// points are re-entrant per stream and need no memcpy nodes when the step is captured in a CUDA graph.
// =======================================================================================
namespace dstag {

// dropout multipliers of one (node, channel): element e = (x*C + c) * T + t draws word (e & 3) of
// philox(e >> 2) exactly like dropout_scale(), but one Philox evaluation serves 4 time steps
template <int T>
__device__ __forceinline__ void row_dropout(float (&ds)[T], long long row, unsigned long long seed, float p,
                                            float inv_keep) {
  static_assert(T % 4 == 0, "row_dropout needs T % 4 == 0");
#pragma unroll
  for (int i = 0; i < T / 4; ++i) {
    const uint4 q = philox4x32((unsigned long long)(row * (T / 4) + i), seed);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      ds[4 * i + j] = ((float)(w[j] >> 8) * (1.0f / 16777216.0f)) < p ? 0.f : inv_keep;
  }
}

template <int T>
__global__ void __launch_bounds__(256)
tail_fused_fwd_kernel(const float* __restrict__ cat, const float* __restrict__ wf, const float* __restrict__ bf,
                      const float* __restrict__ g, const float* __restrict__ r,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                      float* __restrict__ out, float* __restrict__ mean, float* __restrict__ rstd, long long X,
                      float drop_p, unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
  seed = fold_seed(seed, seed_dev);
  constexpr int S = 3 * T - 12, C = 32;
  static_assert(S % 4 == 0 && T % 4 == 0 && T <= 32, "fused tail tiling");
  __shared__ __align__(16) float s_w[T * S];     // fcmy weight (T, S) row-major
  __shared__ float s_b[T];
  for (int i = threadIdx.x; i < T * S; i += blockDim.x) s_w[i] = __ldg(wf + i);
  if (threadIdx.x < T) s_b[threadIdx.x] = __ldg(bf + threadIdx.x);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const float gam = __ldg(gamma + lane), bet = __ldg(beta + lane);
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (long long x = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); x < X; x += nwarps) {
    float cv[S];
#pragma unroll
    for (int s = 0; s < S; ++s) cv[s] = __ldg(cat + (x * S + s) * C + lane);
    float ds[T];
#pragma unroll
    for (int t = 0; t < T; ++t) ds[t] = 1.f;
    if (drop_p > 0.f) row_dropout<T>(ds, x * C + lane, seed, drop_p, inv_keep);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      float a = s_b[t];
#pragma unroll
      for (int s = 0; s < S; s += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(s_w + t * S + s);      // warp-uniform broadcast
        a = fmaf(cv[s], w4.x, a);
        a = fmaf(cv[s + 1], w4.y, a);
        a = fmaf(cv[s + 2], w4.z, a);
        a = fmaf(cv[s + 3], w4.w, a);
      }
      a *= ds[t];
      const long long e = (x * T + t) * C + lane;
      if (g) a += __ldg(g + e);
      const float v = fmaxf(__ldg(r + e) + fmaxf(a, 0.f), 0.f);
      const float mu = warp_sum(v) * (1.f / C);
      const float d = v - mu;
      const float rs = rsqrtf(warp_sum(d * d) * (1.f / C) + eps);
      out[e] = d * rs * gam + bet;
      if (lane == 0) { mean[x * T + t] = mu; rstd[x * T + t] = rs; }
    }
  }
}

template <int T>
__global__ void __launch_bounds__(256, 2)
tail_fused_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ cat, const float* __restrict__ wf,
                      const float* __restrict__ bf, const float* __restrict__ g,
                      const float* __restrict__ r, const float* __restrict__ gamma, const float* __restrict__ mean,
                      const float* __restrict__ rstd, float* __restrict__ dcat, float* __restrict__ dtc_out,
                      float* __restrict__ dg, float* __restrict__ dr, float* __restrict__ dgamma,
                      float* __restrict__ dbeta, long long X, float drop_p, unsigned long long seed,
                      const unsigned long long* __restrict__ seed_dev) {
  seed = fold_seed(seed, seed_dev);
  constexpr int S = 3 * T - 12, C = 32, NW = 8;
  __shared__ __align__(16) float s_w[T * S];     // (T, S): rows for the forward recompute
  __shared__ __align__(16) float s_wt[S * T];    // (S, T): rows for d cat[s] = sum_t dtc[t] w[t][s]
  __shared__ float s_b[T];
  for (int i = threadIdx.x; i < T * S; i += blockDim.x) {
    const float w = __ldg(wf + i);
    s_w[i] = w;
    s_wt[(i % S) * T + i / S] = w;
  }
  if (threadIdx.x < T) s_b[threadIdx.x] = __ldg(bf + threadIdx.x);
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long nwarps = (long long)gridDim.x * NW;
  const float gam = __ldg(gamma + lane);
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  float acc_g = 0.f, acc_be = 0.f;     // d gamma[c], d beta[c] partials (lane = c)
  for (long long x = (long long)blockIdx.x * NW + wid; x < X; x += nwarps) {
    float cv[S];
#pragma unroll
    for (int s = 0; s < S; ++s) cv[s] = __ldg(cat + (x * S + s) * C + lane);
    float ds[T];
#pragma unroll
    for (int t = 0; t < T; ++t) ds[t] = 1.f;
    if (drop_p > 0.f) row_dropout<T>(ds, x * C + lane, seed, drop_p, inv_keep);
    float dtc[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      // ---- recompute the forward of this element
      float a = s_b[t];
#pragma unroll
      for (int s = 0; s < S; s += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(s_w + t * S + s);
        a = fmaf(cv[s], w4.x, a);
        a = fmaf(cv[s + 1], w4.y, a);
        a = fmaf(cv[s + 2], w4.z, a);
        a = fmaf(cv[s + 3], w4.w, a);
      }
      a *= ds[t];
      const long long e = (x * T + t) * C + lane;
      if (g) a += __ldg(g + e);
      const float s2 = __ldg(r + e) + fmaxf(a, 0.f);
      const float vv = fmaxf(s2, 0.f);
      // ---- LayerNorm backward over the 32 lanes
      const float mu = __ldg(mean + x * T + t), rs = __ldg(rstd + x * T + t);
      const float xh = (vv - mu) * rs;
      const float d = __ldg(dout + e);
      const float gy = d * gam;
      const float m1 = warp_sum(gy) * (1.f / C);
      const float m2 = warp_sum(gy * xh) * (1.f / C);
      acc_g = fmaf(d, xh, acc_g);
      acc_be += d;
      const float ds2 = s2 > 0.f ? rs * (gy - m1 - xh * m2) : 0.f;
      const float ds1 = a > 0.f ? ds2 : 0.f;
      dr[e] = ds2;
      if (dg) dg[e] = ds1;
      dtc[t] = ds1 * ds[t];
      dtc_out[e] = dtc[t];           // d tc feeds the weight / bias gradient reduction (dstag_outer_sum)
    }
    // ---- d cat[s] = sum_t dtc[t] Wf[t][s]
#pragma unroll
    for (int s = 0; s < S; ++s) {
      float a = 0.f;
#pragma unroll
      for (int t = 0; t < T; t += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(s_wt + s * T + t);
        a = fmaf(dtc[t], w4.x, a);
        a = fmaf(dtc[t + 1], w4.y, a);
        a = fmaf(dtc[t + 2], w4.z, a);
        a = fmaf(dtc[t + 3], w4.w, a);
      }
      dcat[(x * S + s) * C + lane] = a;
    }
  }
  atomicAdd(dgamma + lane, acc_g);
  atomicAdd(dbeta + lane, acc_be);
}

}  // namespace dstag

extern "C" {

int dstag_tail_fused_supported(int32_t C, int32_t T) { return (C == 32 && T == 12) ? 1 : 0; }

int dstag_tail_fused_fwd(const float* cat, const float* wf, const float* bf, const float* g, const float* r,
                         const float* gamma, const float* beta, float eps, float* out, float* mean, float* rstd,
                         int64_t X, int32_t C, int32_t T, float drop_p, uint64_t seed, const uint64_t* seed_dev,
                         void* stream) {
  const unsigned long long* sd = reinterpret_cast<const unsigned long long*>(seed_dev);
  DSTAG_CHECK_ARG(cat && wf && bf && r && gamma && beta && out && mean && rstd, "dstag_tail_fused_fwd: null pointer");
  DSTAG_CHECK_ARG(X > 0 && drop_p >= 0.f && drop_p < 1.f, "dstag_tail_fused_fwd: bad sizes");
  if (!dstag_tail_fused_supported(C, T)) {
    set_error("dstag_tail_fused_fwd: only C = 32, T = 12 is fused (got C=%d T=%d); use the unfused path", C, T);
    return DSTAG_ERR_UNSUPPORTED;
  }
  const long long blocks = (X + 7) / 8;
  const int grid = (int)(blocks < kNumSMs * 8 ? blocks : kNumSMs * 8);
  tail_fused_fwd_kernel<12><<<grid, 256, 0, (cudaStream_t)stream>>>(cat, wf, bf, g, r, gamma, beta, eps, out, mean,
                                                                    rstd, X, drop_p, seed, sd);
  DSTAG_RETURN_LAUNCH("dstag_tail_fused_fwd");
}

int dstag_tail_fused_bwd(const float* dout, const float* cat, const float* wf, const float* bf, const float* g,
                         const float* r, const float* gamma, const float* mean, const float* rstd, float* dcat,
                         float* dtc, float* dg, float* dr, float* dgamma, float* dbeta, int64_t X, int32_t C,
                         int32_t T, float drop_p, uint64_t seed, const uint64_t* seed_dev, void* stream) {
  const unsigned long long* sd = reinterpret_cast<const unsigned long long*>(seed_dev);
  DSTAG_CHECK_ARG(dout && cat && wf && bf && r && gamma && mean && rstd && dcat && dtc && dr && dgamma && dbeta,
                  "dstag_tail_fused_bwd: null pointer");
  DSTAG_CHECK_ARG((g == nullptr) == (dg == nullptr), "dstag_tail_fused_bwd: g and dg must both be given or both be null");
  DSTAG_CHECK_ARG(X > 0 && drop_p >= 0.f && drop_p < 1.f, "dstag_tail_fused_bwd: bad sizes");
  if (!dstag_tail_fused_supported(C, T)) {
    set_error("dstag_tail_fused_bwd: only C = 32, T = 12 is fused (got C=%d T=%d)", C, T);
    return DSTAG_ERR_UNSUPPORTED;
  }
  const long long blocks = (X + 7) / 8;
  const int grid = (int)(blocks < kNumSMs * 8 ? blocks : kNumSMs * 8);
  tail_fused_bwd_kernel<12><<<grid, 256, 0, (cudaStream_t)stream>>>(dout, cat, wf, bf, g, r, gamma, mean, rstd, dcat,
                                                                    dtc, dg, dr, dgamma, dbeta, X, drop_p, seed, sd);
  DSTAG_RETURN_LAUNCH("dstag_tail_fused_bwd");
}

}  // extern "C"
