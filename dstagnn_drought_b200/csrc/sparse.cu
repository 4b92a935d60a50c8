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

// dst[k, b, n, chunk] = sum_{e in row (k, n)} a[b, eid ? eid[e] : e] * src[b, col[e], chunk]
// (per-order outputs from an order-free source: the forward product Z_k = A_k^T X with the row CSR)
// warp per (k, b, n, 128-float chunk of the row), lane = one float4.
__global__ void __launch_bounds__(256)
spmm_expand_kernel(const float* __restrict__ src, const float* __restrict__ a, const int* __restrict__ rowptr,
                   const int* __restrict__ col, const int* __restrict__ eid, float* __restrict__ dst, int B, int K,
                   int N, int L, int E, int chunks) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  const long long total = (long long)K * B * N * chunks;
  for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total; w += nw) {
    const int ch = (int)(w % chunks);
    long long q = w / chunks;
    const int n = (int)(q % N); q /= N;
    const int b = (int)(q % B), k = (int)(q / B);
    const int c = ch * 128 + 4 * lane;
    if (c >= L) continue;
    const int e0 = __ldg(rowptr + k * N + n), e1 = __ldg(rowptr + k * N + n + 1);
    const float* ab = a + (long long)b * E;
    const float* sb = src + (long long)b * N * L + c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = e0; e < e1; ++e) {
      const float w_e = __ldg(ab + (eid ? __ldg(eid + e) : e));
      const float4 v = __ldg(reinterpret_cast<const float4*>(sb + (long long)__ldg(col + e) * L));
      acc.x = fmaf(w_e, v.x, acc.x); acc.y = fmaf(w_e, v.y, acc.y);
      acc.z = fmaf(w_e, v.z, acc.z); acc.w = fmaf(w_e, v.w, acc.w);
    }
    *reinterpret_cast<float4*>(dst + (((long long)k * B + b) * N + n) * L + c) = acc;
  }
}

// dst[b, n, chunk] = sum_k sum_{e in row (k, n)} a[b, eid ? eid[e] : e] * src[k, b, col[e], chunk]
// (an order-free result from per-order sources: the input gradient dX = sum_k A_k dZ_k with the column CSR)
__global__ void __launch_bounds__(256)
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
    if (c >= L) continue;
    const float* ab = a + (long long)b * E;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < K; ++k) {
      const int e0 = __ldg(rowptr + k * N + n), e1 = __ldg(rowptr + k * N + n + 1);
      const float* sk = src + ((long long)k * B + b) * N * L + c;
      for (int e = e0; e < e1; ++e) {
        const float w_e = __ldg(ab + (eid ? __ldg(eid + e) : e));
        const float4 v = __ldg(reinterpret_cast<const float4*>(sk + (long long)__ldg(col + e) * L));
        acc.x = fmaf(w_e, v.x, acc.x); acc.y = fmaf(w_e, v.y, acc.y);
        acc.z = fmaf(w_e, v.z, acc.z); acc.w = fmaf(w_e, v.w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(dst + bn * L + c) = acc;
  }
}

// h[b, e] = a[b, e] * < dZ[k, b, j, :], X[b, col[e], :] >,  r[b, k, j] = sum_{e in row (k, j)} h[b, e]
// warp per (b, row (k, j)); the row of dZ stays in registers when L <= 32 * 4 * kDotRegs
constexpr int kDotRegs = 4;     // float4 per lane kept in registers: rows of up to 512 floats
__global__ void __launch_bounds__(256)
edge_dots_kernel(const float* __restrict__ dZ, const float* __restrict__ X, const float* __restrict__ a,
                 const int* __restrict__ rowptr, const int* __restrict__ col, float* __restrict__ h,
                 float* __restrict__ r, int B, int K, int N, int L, int E) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  const long long total = (long long)B * K * N;
  const int n4 = L >> 2;                                            // L % 4 == 0
  for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total; w += nw) {
    const int j = (int)(w % N);
    const long long bk = w / N;
    const int k = (int)(bk % K), b = (int)(bk / K);
    const int e0 = __ldg(rowptr + k * N + j), e1 = __ldg(rowptr + k * N + j + 1);
    const float4* dg = reinterpret_cast<const float4*>(dZ + (((long long)k * B + b) * N + j) * L);
    const float* Xb = X + (long long)b * N * L;
    float4 g[kDotRegs];
    const bool cached = n4 <= 32 * kDotRegs;
    if (cached && e0 < e1) {
#pragma unroll
      for (int u = 0; u < kDotRegs; ++u) {
        const int c4 = lane + 32 * u;
        g[u] = c4 < n4 ? __ldg(dg + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float rsum = 0.f;
    for (int e = e0; e < e1; ++e) {
      const float4* v = reinterpret_cast<const float4*>(Xb + (long long)__ldg(col + e) * L);
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
      const float he = __ldg(a + (long long)b * E + e) * s;
      if (lane == 0) h[(long long)b * E + e] = he;
      rsum += he;
    }
    if (lane == 0) r[w] = rsum;                                     // r is (B, K, N)
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
// dqk[b, n, 1, k, :] = alpha ( sum_{e in row (k, n)}    h[b, e]   * Qs[b, i_e, k, :] - r[b, k, n] * O[b, n, k, :] )  (d Ks)
// warp per (b, n, k), lane = d (dk <= 32)
__global__ void __launch_bounds__(256)
dqk_combine_kernel(const float* __restrict__ qk, const float* __restrict__ h, const float* __restrict__ r,
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
    if (lane >= dk) continue;
    const float* qb = qk + (long long)b * N * nstride + (long long)k * dk + lane;     // Qs[b, ., k, lane]
    const float* kb = qb + (long long)K * dk;                                           // Ks[b, ., k, lane]
    const float* hb = h + (long long)b * E;
    float dq = 0.f, dkk = 0.f;
    for (int e = __ldg(t_rowptr + k * N + n); e < __ldg(t_rowptr + k * N + n + 1); ++e)
      dq = fmaf(__ldg(hb + __ldg(t_eid + e)), __ldg(kb + (long long)__ldg(t_col + e) * nstride), dq);
    for (int e = __ldg(rowptr + k * N + n); e < __ldg(rowptr + k * N + n + 1); ++e)
      dkk = fmaf(__ldg(hb + e), __ldg(qb + (long long)__ldg(col + e) * nstride), dkk);
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
  spmm_expand_kernel<<<grid_for((long long)K * B * N * chunks, 8, 16), 256, 0, (cudaStream_t)stream>>>(
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
  edge_dots_kernel<<<grid_for((long long)B * K * N, 8, 16), 256, 0, (cudaStream_t)stream>>>(dZ, X, a, rowptr, col, h, r, B, K, N,
                                                                                            L, E);
  DSTAG_RETURN_LAUNCH("dstag_edge_dots");
}

int dstag_scale_keys(const float* qk, const float* r, float* rK, int32_t B, int32_t N, int32_t K, int32_t dk,
                     void* stream) {
  DSTAG_CHECK_ARG(qk && r && rK && B > 0 && N > 0 && K > 0 && dk > 0, "dstag_scale_keys: bad arguments");
  scale_keys_kernel<<<grid_for((long long)B * N * K * dk, 256), 256, 0, (cudaStream_t)stream>>>(qk, r, rK, B, N, K, dk);
  DSTAG_RETURN_LAUNCH("dstag_scale_keys");
}

int dstag_dqk_combine(const float* qk, const float* h, const float* r, const float* O, const float* W,
                      const int32_t* rowptr, const int32_t* col, const int32_t* t_rowptr, const int32_t* t_col,
                      const int32_t* t_eid, float* dqk, int32_t B, int32_t N, int32_t K, int32_t dk, int32_t E,
                      float alpha, void* stream) {
  DSTAG_CHECK_ARG(qk && h && r && O && W && rowptr && col && t_rowptr && t_col && t_eid && dqk,
                  "dstag_dqk_combine: null pointer");
  DSTAG_CHECK_ARG(B > 0 && N > 0 && K > 0 && dk > 0 && dk <= 32, "dstag_dqk_combine: bad sizes (d_k <= 32)");
  dqk_combine_kernel<<<grid_for((long long)B * N * K, 8, 16), 256, 0, (cudaStream_t)stream>>>(
      qk, h, r, O, W, rowptr, col, t_rowptr, t_col, t_eid, dqk, B, N, K, dk, E, alpha);
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
