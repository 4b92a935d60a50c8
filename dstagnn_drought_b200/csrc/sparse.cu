// Sparse Chebyshev path of cheb_conv_withSAt (model/DSTAGNN_my.py:117-133).
//
// The Chebyshev polynomials of this model are Hadamard powers of the scaled Laplacian (lib/utils.py:198-203,
// T_k = 2 L~ (.) T_{k-1} - T_{k-2} ELEMENT-wise): T_0 = I and every T_k has the sparsity pattern of the
// adjacency matrix (the STAG generator keeps the top 1 % of every row, data/STAG_gen.py:108-116).  The operator
// T_k (.) softmax(S_k) that multiplies the graph signal is therefore ~1 % dense, and so is the part of the map
// gradient that survives the Hadamard product.  The dense formulation (gemm_tc.cu) multiplies N x N x F T
// zeros; here only the edges are touched:
//
//   forward   a[b,e]  = T_k[i,j] * P[b,k,j,i]                       edge e = (k, j, i), P = column softmax (row j here)
//             Z[k,b,j,:] = sum_{e in row (k,j)} a[b,e] * X[b,i_e,:]          (then the dense Theta_k mix + ReLU)
//   backward  dX[b,i,:] = sum_k sum_{e in column (k,i)} a[b,e] * dZ[k,b,j_e,:]
//             h[b,e]  = a[b,e] * < dZ[k,b,j,:], X[b,i,:] >           (= P (.) dP on the edge)
//             r[b,k,j] = sum_{e in row (k,j)} h[b,e]
//             dS[b,k,j,i] = h[b,e] (on edges) - r[b,k,j] * P[b,k,j,i]   -- never materialised:
//             dK_j = alpha ( sum_e h_e Q_{i_e} - r_j * (P Q)_j ),   dQ_i = alpha ( sum_e h_e K_{j_e} - (P^T (r K))_i )
//             d biasT[k,j,i] = sum_b dS[b,k,j,i]                     only where adj_pa != 0 (the mask gradient)
//   The two dense products P Q and P^T (r K) run on the tensor-core engine; P >= 0, so they are free of the
//   cancellation that a product with the zero-sum rows of dS has.
// All kernels here are exact fp32 (FMA) gather / reduce kernels, one warp per graph row.
#include "common.cuh"

namespace dstag {

// a[b, e] = tval[e] * P[b, k, j, i]   (thread per (b, e); erow[e] = k * N + j)
__global__ void __launch_bounds__(256)
edge_probs_kernel(const float* __restrict__ P, const int* __restrict__ erow, const int* __restrict__ col,
                  const float* __restrict__ tval, float* __restrict__ a, int B, int KN, int ld, int E) {
  const long long n = (long long)B * E;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const int b = (int)(t / E), e = (int)(t - (long long)b * E);
    a[t] = __ldg(tval + e) * __ldg(P + ((long long)b * KN + __ldg(erow + e)) * ld + __ldg(col + e));
  }
}

// Every kernel below walks the edges of a CSR row the same way: the 32 lanes fetch up to 32 column indices
// and edge weights at once, then the warp broadcasts them one edge at a time (no dependent index loads in
// the gather loop, which the compiler unrolls so that several row gathers are in flight).
struct EdgeBatch {
  int c;        // this lane's column index
  float w;      // this lane's edge weight
  int n;        // edges in the batch
  __device__ __forceinline__ void load(int e0, int e1, int lane, const int* __restrict__ col,
                                       const float* __restrict__ a, const int* __restrict__ eid) {
    const int e = e0 + lane;
    n = min(32, e1 - e0);
    c = e < e1 ? __ldg(col + e) : 0;
    w = e < e1 ? __ldg(a + (eid ? __ldg(eid + e) : e)) : 0.f;
  }
};

// dst[k, b, n, chunk] = sum_{e in row (k, n)} a[b, eid ? eid[e] : e] * src[b, col[e], chunk]
// (per-order outputs from an order-free source: the forward product Z_k = A_k^T X with the row CSR)
// warp per (b, n, 128-float chunk of the row) looping over the orders k: T_1, T_2, ... share the adjacency's
// pattern, so the rows gathered for one order are L1 hits for the next.  lane = one float4.
__global__ void __launch_bounds__(256, 6)
spmm_expand_kernel(const float* __restrict__ src, const float* __restrict__ a, const int* __restrict__ rowptr,
                   const int* __restrict__ col, const int* __restrict__ eid, float* __restrict__ dst, int B, int K,
                   int N, int L, int E, int chunks) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  const long long total = (long long)B * N * chunks;
  for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total; w += nw) {
    const int ch = (int)(w % chunks);
    const long long bn = w / chunks;
    const int n = (int)(bn % N), b = (int)(bn / N);
    const int c = ch * 128 + 4 * lane;
    const bool on = c < L;
    const float* ab = a + (long long)b * E;
    const float* sb = src + (long long)b * N * L + (on ? c : 0);
    for (int k = 0; k < K; ++k) {
      const int e0 = __ldg(rowptr + k * N + n), e1 = __ldg(rowptr + k * N + n + 1);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int eb = e0; eb < e1; eb += 32) {
        EdgeBatch eb_;
        eb_.load(eb, e1, lane, col, ab, eid);
#pragma unroll 4
        for (int t = 0; t < eb_.n; ++t) {
          const int ci = __shfl_sync(0xffffffffu, eb_.c, t);
          const float w_e = __shfl_sync(0xffffffffu, eb_.w, t);
          if (on) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(sb + (long long)ci * L));
            acc.x = fmaf(w_e, v.x, acc.x); acc.y = fmaf(w_e, v.y, acc.y);
            acc.z = fmaf(w_e, v.z, acc.z); acc.w = fmaf(w_e, v.w, acc.w);
          }
        }
      }
      if (on) *reinterpret_cast<float4*>(dst + (((long long)k * B + b) * N + n) * L + c) = acc;
    }
  }
}

// dst[b, n, chunk] = sum_k sum_{e in row (k, n)} a[b, eid ? eid[e] : e] * src[k, b, col[e], chunk]
// (an order-free result from per-order sources: the input gradient dX = sum_k A_k dZ_k with the column CSR)
__global__ void __launch_bounds__(256, 6)
spmm_reduce_kernel(const float* __restrict__ src, const float* __restrict__ a, const int* __restrict__ rowptr,
                   const int* __restrict__ col, const int* __restrict__ eid, float* __restrict__ dst, int B, int K,
                   int N, int L, int E, int chunks) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  const long long total = (long long)B * N * chunks;
  for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total; w += nw) {
    const int ch = (int)(w % chunks);
    const long long bn = w / chunks;
    const int n = (int)(bn % N), b = (int)(bn / N);
    const int c = ch * 128 + 4 * lane;
    const bool on = c < L;
    const float* ab = a + (long long)b * E;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < K; ++k) {
      const int e0 = __ldg(rowptr + k * N + n), e1 = __ldg(rowptr + k * N + n + 1);
      const float* sk = src + ((long long)k * B + b) * N * L + (on ? c : 0);
      for (int eb = e0; eb < e1; eb += 32) {
        EdgeBatch eb_;
        eb_.load(eb, e1, lane, col, ab, eid);
#pragma unroll 4
        for (int t = 0; t < eb_.n; ++t) {
          const int ci = __shfl_sync(0xffffffffu, eb_.c, t);
          const float w_e = __shfl_sync(0xffffffffu, eb_.w, t);
          if (on) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(sk + (long long)ci * L));
            acc.x = fmaf(w_e, v.x, acc.x); acc.y = fmaf(w_e, v.y, acc.y);
            acc.z = fmaf(w_e, v.z, acc.z); acc.w = fmaf(w_e, v.w, acc.w);
          }
        }
      }
    }
    if (on) *reinterpret_cast<float4*>(dst + bn * L + c) = acc;
  }
}

// h[b, e] = a[b, e] * < dZ[k, b, j, :], X[b, col[e], :] >,  r[b, k, j] = sum_{e in row (k, j)} h[b, e]
// warp per (b, j) looping over the orders (the X rows of one order are L1 hits for the next); the row of dZ
// stays in registers when L <= 32 * 4 * kDotRegs
constexpr int kDotRegs = 4;     // float4 per lane kept in registers: rows of up to 512 floats
__global__ void __launch_bounds__(256, 4)
edge_dots_kernel(const float* __restrict__ dZ, const float* __restrict__ X, const float* __restrict__ a,
                 const int* __restrict__ rowptr, const int* __restrict__ col, float* __restrict__ h,
                 float* __restrict__ r, int B, int K, int N, int L, int E) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  const long long total = (long long)B * N;
  const int n4 = L >> 2;                                            // L % 4 == 0
  const bool cached = n4 <= 32 * kDotRegs;
  for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total; w += nw) {
    const int j = (int)(w % N), b = (int)(w / N);
    const float* Xb = X + (long long)b * N * L;
    const float* ab = a + (long long)b * E;
    for (int k = 0; k < K; ++k) {
      const int e0 = __ldg(rowptr + k * N + j), e1 = __ldg(rowptr + k * N + j + 1);
      const float4* dg = reinterpret_cast<const float4*>(dZ + (((long long)k * B + b) * N + j) * L);
      float4 g[kDotRegs];
      if (cached && e0 < e1) {
#pragma unroll
        for (int u = 0; u < kDotRegs; ++u) {
          const int c4 = lane + 32 * u;
          g[u] = c4 < n4 ? __ldg(dg + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      float rsum = 0.f;
      for (int eb = e0; eb < e1; eb += 32) {
        EdgeBatch eb_;
        eb_.load(eb, e1, lane, col, ab, nullptr);
        float hmine = 0.f;                                            // lane t keeps h of edge eb + t
#pragma unroll 2
        for (int t = 0; t < eb_.n; ++t) {
          const int ci = __shfl_sync(0xffffffffu, eb_.c, t);
          const float4* v = reinterpret_cast<const float4*>(Xb + (long long)ci * L);
          float s = 0.f;
          if (cached) {
#pragma unroll
            for (int u = 0; u < kDotRegs; ++u) {
              const int c4 = lane + 32 * u;
              if (c4 < n4) {
                const float4 x = __ldg(v + c4);
                s = fmaf(g[u].x, x.x, fmaf(g[u].y, x.y, fmaf(g[u].z, x.z, fmaf(g[u].w, x.w, s))));
              }
            }
          } else {
            for (int c4 = lane; c4 < n4; c4 += 32) {
              const float4 x = __ldg(v + c4), y = __ldg(dg + c4);
              s = fmaf(y.x, x.x, fmaf(y.y, x.y, fmaf(y.z, x.z, fmaf(y.w, x.w, s))));
            }
          }
          s = warp_sum(s);
          if (lane == t) hmine = eb_.w * s;
        }
        if (lane < eb_.n) h[(long long)b * E + eb + lane] = hmine;   // one coalesced store per batch
        rsum += warp_sum(lane < eb_.n ? hmine : 0.f);
      }
      if (lane == 0) r[((long long)b * K + k) * N + j] = rsum;        // r is (B, K, N)
    }
  }
}

// Qc[b, n, k, :] = Qs[b, n, k, :] - mean_n Qs[b, n, k, :]     (qk: (B, N, 2, K, dk) -> Qc: (B, N, K, dk))
// The key gradient is alpha * sum_i dS[j, i] Qs[i] with rows of dS that sum to ZERO: any vector common to all
// Qs[i] drops out.  Removing the node mean up front keeps the two halves of that sum (the edge part and
// r_j * (P Qs)_j) from cancelling a large common component -- the tensor core's truncating fp32 accumulation
// of P Qs otherwise shows up at 2e-4 of the gradient on the real PEMS04 graph.
// Two passes: column sums over node slices (one atomic per column and slice), then the subtraction.
constexpr int kCenterSlices = 16;
__global__ void __launch_bounds__(256)
query_sums_kernel(const float* __restrict__ qk, float* __restrict__ sums, int N, int W) {
  // grid (B, kCenterSlices); warp w of 8 walks rows slice_start + w, + 8, ...; lane covers columns lane + 32 u
  __shared__ float red[8][128];
  const int b = blockIdx.x, sl = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int per = (N + kCenterSlices - 1) / kCenterSlices, n0 = sl * per, n1 = min(N, n0 + per);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};                              // W <= 128
  for (int n = n0 + w; n < n1; n += 8) {
    const float* row = qk + ((long long)b * N + n) * 2 * W;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (lane + 32 * u < W) acc[u] += __ldg(row + lane + 32 * u);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) red[w][lane + 32 * u] = acc[u];
  __syncthreads();
  if (threadIdx.x < W) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    atomicAdd(sums + (long long)b * W + threadIdx.x, s);
  }
}
__global__ void __launch_bounds__(256)
center_queries_kernel(const float* __restrict__ qk, const float* __restrict__ sums, float* __restrict__ Qc, int B, int N,
                      int W) {
  const long long n = (long long)B * N * W;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float inv = 1.f / (float)N;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const int c = (int)(t % W);
    const long long bn = t / W;
    Qc[t] = __ldg(qk + bn * 2 * W + c) - __ldg(sums + (bn / N) * W + c) * inv;
  }
}

// rK[b, j, k, :] = r[b, k, j] * Ks[b, j, k, :]     (qk: (B, N, 2, K, dk); rK: (B, N, K, dk))
__global__ void __launch_bounds__(256)
scale_keys_kernel(const float* __restrict__ qk, const float* __restrict__ r, float* __restrict__ rK, int B, int N, int K,
                  int dk) {
  const long long n = (long long)B * N * K * dk;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const int d = (int)(t % dk);
    long long q = t / dk;
    const int k = (int)(q % K); q /= K;
    const int j = (int)(q % N), b = (int)(q / N);
    rK[t] = __ldg(r + ((long long)b * K + k) * N + j) * __ldg(qk + ((((long long)b * N + j) * 2 + 1) * K + k) * dk + d);
  }
}

// dqk[b, n, 0, k, :] = alpha ( sum_{e in column (k, n)} h[b, eid] * Ks[b, j_e, k, :] - W[b, n, k, :] )     (d Qs)
// dqk[b, n, 1, k, :] = alpha ( sum_{e in row (k, n)}    h[b, e]   * Qc[b, i_e, k, :] - r[b, k, n] * O[b, n, k, :] )  (d Ks)
// with Qc the node-centred queries and O = P Qc
// warp per (b, n, k), lane = d (dk <= 32)
__global__ void __launch_bounds__(256)
dqk_combine_kernel(const float* __restrict__ qk, const float* __restrict__ Qc, const float* __restrict__ h,
                   const float* __restrict__ r,
                   const float* __restrict__ O, const float* __restrict__ W, const int* __restrict__ rowptr,
                   const int* __restrict__ col, const int* __restrict__ t_rowptr, const int* __restrict__ t_col,
                   const int* __restrict__ t_eid, float* __restrict__ dqk, int B, int N, int K, int dk, int E,
                   float alpha) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  const long long total = (long long)B * N * K;
  const long long nstride = 2LL * K * dk;
  for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total; w += nw) {
    const int k = (int)(w % K);
    const long long bn = w / K;
    const int n = (int)(bn % N), b = (int)(bn / N);
    const int lane32 = lane;
    const bool act = lane < dk;                                      // (all 32 lanes take part in the shuffles)
    const int ld_ = act ? lane : 0;
    const float* kb = qk + (long long)b * N * nstride + (long long)(K + k) * dk + ld_;  // Ks[b, ., k, lane]
    const float* qb = Qc + (long long)b * N * K * dk + (long long)k * dk + ld_;         // centred Qs[b, ., k, lane]
    const long long qstride = (long long)K * dk;
    const float* hb = h + (long long)b * E;
    float dq = 0.f, dkk = 0.f;
    {
      const int e0 = __ldg(t_rowptr + k * N + n), e1 = __ldg(t_rowptr + k * N + n + 1);
      for (int eb = e0; eb < e1; eb += 32) {
        EdgeBatch eb_;
        eb_.load(eb, e1, lane32, t_col, hb, t_eid);
#pragma unroll 4
        for (int t = 0; t < eb_.n; ++t) {
          const int ci = __shfl_sync(0xffffffffu, eb_.c, t);
          const float he = __shfl_sync(0xffffffffu, eb_.w, t);
          if (act) dq = fmaf(he, __ldg(kb + (long long)ci * nstride), dq);
        }
      }
    }
    {
      const int e0 = __ldg(rowptr + k * N + n), e1 = __ldg(rowptr + k * N + n + 1);
      for (int eb = e0; eb < e1; eb += 32) {
        EdgeBatch eb_;
        eb_.load(eb, e1, lane32, col, hb, nullptr);
#pragma unroll 4
        for (int t = 0; t < eb_.n; ++t) {
          const int ci = __shfl_sync(0xffffffffu, eb_.c, t);
          const float he = __shfl_sync(0xffffffffu, eb_.w, t);
          if (act) dkk = fmaf(he, __ldg(qb + (long long)ci * qstride), dkk);
        }
      }
    }
    if (!act) continue;
    const long long o = (bn * K + k) * dk + lane;                                       // O, W: (B, N, K, dk)
    dq -= __ldg(W + o);
    dkk -= __ldg(r + ((long long)b * K + k) * N + n) * __ldg(O + o);
    float* out = dqk + bn * nstride + (long long)k * dk + lane;
    out[0] = alpha * dq;
    out[(long long)K * dk] = alpha * dkk;
  }
}

// d biasT[k, pa_j[p], pa_i[p]] = sum_b ( h[b, pa_eid[k, p]] (if >= 0) - r[b, k, j] * P[b, k, j, i] )
// thread per (k, p); everything else of d biasT stays zero (the caller zero-fills it)
__global__ void __launch_bounds__(256)
dbias_sparse_kernel(const float* __restrict__ P, const float* __restrict__ h, const float* __restrict__ r,
                    const int* __restrict__ pa_j, const int* __restrict__ pa_i, const int* __restrict__ pa_eid,
                    float* __restrict__ dbiasT, int B, int K, int N, int ld, int E, int npa) {
  const long long n = (long long)K * npa;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const int k = (int)(t / npa), p = (int)(t - (long long)k * npa);
    const int j = __ldg(pa_j + p), i = __ldg(pa_i + p), eid = __ldg(pa_eid + t);
    float s = 0.f;
    for (int b = 0; b < B; ++b) {
      const long long row = ((long long)b * K + k) * N + j;
      float v = -__ldg(r + row) * __ldg(P + row * ld + i);
      if (eid >= 0) v += __ldg(h + (long long)b * E + eid);
      s += v;
    }
    dbiasT[((long long)k * N + j) * ld + i] = s;
  }
}

static unsigned grid_for(long long items, int per_block, int max_blocks_per_sm = 8) {
  long long blocks = (items + per_block - 1) / per_block;
  const long long cap = (long long)kNumSMs * max_blocks_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace dstag

using namespace dstag;

extern "C" {

int dstag_edge_probs(const float* P, const int32_t* erow, const int32_t* col, const float* tval, float* a, int32_t B,
                     int32_t K, int32_t N, int32_t ld, int32_t E, void* stream) {
  DSTAG_CHECK_ARG(P && erow && col && tval && a && B > 0 && K > 0 && N > 0 && ld >= N && E > 0,
                  "dstag_edge_probs: bad arguments");
  edge_probs_kernel<<<grid_for((long long)B * E, 256), 256, 0, (cudaStream_t)stream>>>(P, erow, col, tval, a, B, K * N, ld, E);
  DSTAG_RETURN_LAUNCH("dstag_edge_probs");
}

int dstag_spmm_expand(const float* src, const float* a, const int32_t* rowptr, const int32_t* col, const int32_t* eid,
                      float* dst, int32_t B, int32_t K, int32_t N, int32_t L, int32_t E, void* stream) {
  DSTAG_CHECK_ARG(src && a && rowptr && col && dst && B > 0 && K > 0 && N > 0 && L > 0 && L % 4 == 0,
                  "dstag_spmm_expand: bad arguments (row length must be a multiple of 4)");
  DSTAG_CHECK_ARG(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0,
                  "dstag_spmm_expand: 16-byte aligned operands");
  const int chunks = (L + 127) / 128;
  spmm_expand_kernel<<<grid_for((long long)B * N * chunks, 8, 16), 256, 0, (cudaStream_t)stream>>>(
      src, a, rowptr, col, eid, dst, B, K, N, L, E, chunks);
  DSTAG_RETURN_LAUNCH("dstag_spmm_expand");
}

int dstag_spmm_reduce(const float* src, const float* a, const int32_t* rowptr, const int32_t* col, const int32_t* eid,
                      float* dst, int32_t B, int32_t K, int32_t N, int32_t L, int32_t E, void* stream) {
  DSTAG_CHECK_ARG(src && a && rowptr && col && dst && B > 0 && K > 0 && N > 0 && L > 0 && L % 4 == 0,
                  "dstag_spmm_reduce: bad arguments (row length must be a multiple of 4)");
  DSTAG_CHECK_ARG(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0,
                  "dstag_spmm_reduce: 16-byte aligned operands");
  const int chunks = (L + 127) / 128;
  spmm_reduce_kernel<<<grid_for((long long)B * N * chunks, 8, 16), 256, 0, (cudaStream_t)stream>>>(
      src, a, rowptr, col, eid, dst, B, K, N, L, E, chunks);
  DSTAG_RETURN_LAUNCH("dstag_spmm_reduce");
}

int dstag_edge_dots(const float* dZ, const float* X, const float* a, const int32_t* rowptr, const int32_t* col, float* h,
                    float* r, int32_t B, int32_t K, int32_t N, int32_t L, int32_t E, void* stream) {
  DSTAG_CHECK_ARG(dZ && X && a && rowptr && col && h && r && B > 0 && K > 0 && N > 0 && L > 0 && L % 4 == 0,
                  "dstag_edge_dots: bad arguments (row length must be a multiple of 4)");
  DSTAG_CHECK_ARG(((reinterpret_cast<uintptr_t>(dZ) | reinterpret_cast<uintptr_t>(X)) & 15) == 0,
                  "dstag_edge_dots: 16-byte aligned operands");
  edge_dots_kernel<<<grid_for((long long)B * N, 8, 16), 256, 0, (cudaStream_t)stream>>>(dZ, X, a, rowptr, col, h, r, B, K, N,
                                                                                            L, E);
  DSTAG_RETURN_LAUNCH("dstag_edge_dots");
}

int dstag_scale_keys(const float* qk, const float* r, float* rK, int32_t B, int32_t N, int32_t K, int32_t dk,
                     void* stream) {
  DSTAG_CHECK_ARG(qk && r && rK && B > 0 && N > 0 && K > 0 && dk > 0, "dstag_scale_keys: bad arguments");
  scale_keys_kernel<<<grid_for((long long)B * N * K * dk, 256), 256, 0, (cudaStream_t)stream>>>(qk, r, rK, B, N, K, dk);
  DSTAG_RETURN_LAUNCH("dstag_scale_keys");
}

int dstag_center_queries(const float* qk, float* sums, float* Qc, int32_t B, int32_t N, int32_t K, int32_t dk,
                         void* stream) {
  DSTAG_CHECK_ARG(qk && sums && Qc && B > 0 && N > 0 && K > 0 && dk > 0 && K * dk <= 128 && B <= 65535,
                  "dstag_center_queries: bad arguments (K * d_k <= 128)");
  const int W = K * dk;
  cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(float) * (size_t)B * W, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_error("dstag_center_queries: memset failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  query_sums_kernel<<<dim3(B, kCenterSlices), 256, 0, (cudaStream_t)stream>>>(qk, sums, N, W);
  DSTAG_CHECK_LAUNCH("dstag_center_queries(sums)");
  center_queries_kernel<<<grid_for((long long)B * N * W, 256), 256, 0, (cudaStream_t)stream>>>(qk, sums, Qc, B, N, W);
  DSTAG_RETURN_LAUNCH("dstag_center_queries");
}

int dstag_dqk_combine(const float* qk, const float* Qc, const float* h, const float* r, const float* O, const float* W,
                      const int32_t* rowptr, const int32_t* col, const int32_t* t_rowptr, const int32_t* t_col,
                      const int32_t* t_eid, float* dqk, int32_t B, int32_t N, int32_t K, int32_t dk, int32_t E,
                      float alpha, void* stream) {
  DSTAG_CHECK_ARG(qk && Qc && h && r && O && W && rowptr && col && t_rowptr && t_col && t_eid && dqk,
                  "dstag_dqk_combine: null pointer");
  DSTAG_CHECK_ARG(B > 0 && N > 0 && K > 0 && dk > 0 && dk <= 32, "dstag_dqk_combine: bad sizes (d_k <= 32)");
  dqk_combine_kernel<<<grid_for((long long)B * N * K, 8, 16), 256, 0, (cudaStream_t)stream>>>(
      qk, Qc, h, r, O, W, rowptr, col, t_rowptr, t_col, t_eid, dqk, B, N, K, dk, E, alpha);
  DSTAG_RETURN_LAUNCH("dstag_dqk_combine");
}

int dstag_dbias_sparse(const float* P, const float* h, const float* r, const int32_t* pa_j, const int32_t* pa_i,
                       const int32_t* pa_eid, float* dbiasT, int32_t B, int32_t K, int32_t N, int32_t ld, int32_t E,
                       int32_t npa, void* stream) {
  DSTAG_CHECK_ARG(P && h && r && pa_j && pa_i && pa_eid && dbiasT && B > 0 && K > 0 && N > 0 && ld >= N && npa > 0,
                  "dstag_dbias_sparse: bad arguments");
  dbias_sparse_kernel<<<grid_for((long long)K * npa, 256), 256, 0, (cudaStream_t)stream>>>(P, h, r, pa_j, pa_i, pa_eid,
                                                                                          dbiasT, B, K, N, ld, E, npa);
  DSTAG_RETURN_LAUNCH("dstag_dbias_sparse");
}

}  // extern "C"
