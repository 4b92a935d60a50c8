// Block tail: GTU gate and the fused  dropout -> relu(tc [+ g]) -> relu(r + .) -> LayerNorm_C
// kernel (model/DSTAGNN_my.py:194-196, 243-252).  tc / r / out are (X = B*N, C, T) with T
// innermost (the block's boundary layout); the graph-conv output g and the GTU conv tensors are
// channels-last (X, T, C) so that the temporal convolutions are plain K-contiguous contractions.
// A CTA stages a group of nodes in shared memory with fully coalesced loads,
// normalises every (node, t) column over the C channels out of shared memory (padded so
// the column walk is bank-conflict free), and stores coalesced.
#include "common.cuh"

namespace dstag {

// exp-based gate functions on the fast-math units (abs error ~1e-7: far inside the parity budget;
// the libm tanhf/expf versions made the gate kernels instruction-bound, not HBM-bound)
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

__global__ void gtu_gate_fwd_kernel(const float* __restrict__ u, float* __restrict__ out, long long total,
                                    int C, int Tp, int out_ld, int out_off) {
  // u: (X, Tp, 2C) channels-last conv output;  out[x, c, off + t] = tanh(u[x,t,c]) * sigmoid(u[x,t,C+c])
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int ct = C * Tp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long x = i / ct;
    const int rem = (int)(i - x * ct);
    const int c = rem / Tp, t = rem - c * Tp;          // t fastest: coalesced writes; reads hit L1 lines
    const float* up = u + (x * Tp + t) * 2 * C;
    out[(x * C + c) * out_ld + out_off + t] = tanhf_(__ldg(up + c)) * sigmoidf_(__ldg(up + C + c));
  }
}

__global__ void __launch_bounds__(256)
gtu_gate_bwd_kernel(const float* __restrict__ u, const float* __restrict__ dout, float* __restrict__ dup,
                    float* __restrict__ db, long long total, int C, int Tp, int pad, int out_ld, int out_off) {
  // dup: (X, Tp + 2 pad, 2C) channels-last gradient of the conv output, zero in the pad rows.
  // One thread per (x, padded t, c): evaluates the gate once and writes both halves (p and q).
  // db (2C, optional): the conv bias gradient = column sums of dup.  When the grid stride is a
  // multiple of C every thread keeps one channel, so the sums stay in registers until the end.
  __shared__ float red[2][256];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int Tpad = Tp + 2 * pad;
  const bool fixed_c = db && (256 % C == 0);
  float sp = 0.f, sq = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long xt = i / C;                        // c fastest: coalesced reads of u, writes of dup
    const int c = (int)(i - xt * C);
    const long long x = xt / Tpad;
    const int tp = (int)(xt - x * Tpad), t = tp - pad;
    float rp = 0.f, rq = 0.f;
    if (t >= 0 && t < Tp) {
      const float* up = u + (x * Tp + t) * 2 * C;
      const float th = tanhf_(__ldg(up + c)), sg = sigmoidf_(__ldg(up + C + c));
      const float d = __ldg(dout + (x * C + c) * out_ld + out_off + t);
      rp = d * sg * (1.f - th * th);
      rq = d * th * sg * (1.f - sg);
      if (db && !fixed_c) {
        atomicAdd(db + c, rp);
        atomicAdd(db + C + c, rq);
      }
    }
    sp += rp;
    sq += rq;
    dup[xt * 2 * C + c] = rp;
    dup[xt * 2 * C + C + c] = rq;
  }
  if (fixed_c) {
    red[0][threadIdx.x] = sp;
    red[1][threadIdx.x] = sq;
    __syncthreads();
    if (threadIdx.x < C) {                             // thread c sums the 256 / C threads of its channel
      float a = 0.f, b = 0.f;
      for (int j = threadIdx.x; j < 256; j += C) { a += red[0][j]; b += red[1][j]; }
      atomicAdd(db + threadIdx.x, a);
      atomicAdd(db + C + threadIdx.x, b);
    }
  }
}

// ---------------------------------------------------------------------------------------
// out[m, n] += sum_r a[r, m] * b[r, n] for two tall, narrow row-major matrices (m * n / 4 <= 256):
// the fcmy weight gradient (T x (3T-12) outputs, B*N*C rows).  A tensor-core tile would be 98% padding;
// this is a streaming reduction: a CTA stages 256-row slabs in shared memory, thread (row phase, m,
// n quad) keeps 4 accumulators, the row phases are folded through shared memory at the end.
// ---------------------------------------------------------------------------------------
constexpr int kOsRows = 256;
__global__ void __launch_bounds__(256)
outer_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long rows,
                 int m, int n, int phases) {
  extern __shared__ float sm[];
  float* as = sm;                     // kOsRows x m
  float* bs = sm + kOsRows * m;       // kOsRows x n
  const int nq = n >> 2, per = m * nq;               // outputs quads
  const int ph = threadIdx.x / per, id = threadIdx.x - ph * per;
  const bool active = ph < phases;
  const int mi = id / nq, q = id - mi * nq;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long nslab = (rows + kOsRows - 1) / kOsRows;
  const bool vec = (m % 4 == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0);   // n % 4 == 0, b aligned: checked by the host
  for (long long sl = blockIdx.x; sl < nslab; sl += gridDim.x) {
    const long long r0 = sl * kOsRows;
    const int nr = (int)min((long long)kOsRows, rows - r0);
    __syncthreads();
    if (vec) {                                            // whole slabs are 16-byte aligned: float4 streams
      const float4* a4 = reinterpret_cast<const float4*>(a + r0 * m);
      const float4* b4 = reinterpret_cast<const float4*>(b + r0 * n);
      const int na = nr * m / 4, nb = nr * n / 4;
#pragma unroll 4
      for (int i = threadIdx.x; i < na; i += 256) reinterpret_cast<float4*>(as)[i] = __ldg(a4 + i);
#pragma unroll 8
      for (int i = threadIdx.x; i < nb; i += 256) reinterpret_cast<float4*>(bs)[i] = __ldg(b4 + i);
    } else {
      for (int i = threadIdx.x; i < nr * m; i += blockDim.x) as[i] = __ldg(a + r0 * m + i);
      for (int i = threadIdx.x; i < nr * n; i += blockDim.x) bs[i] = __ldg(b + r0 * n + i);
    }
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int r = ph; r < nr; r += phases) {
        const float av = as[r * m + mi];
        const float4 bv = *reinterpret_cast<const float4*>(bs + r * n + 4 * q);
        acc.x = fmaf(av, bv.x, acc.x); acc.y = fmaf(av, bv.y, acc.y);
        acc.z = fmaf(av, bv.z, acc.z); acc.w = fmaf(av, bv.w, acc.w);
      }
    }
  }
  __syncthreads();
  float4* red = reinterpret_cast<float4*>(sm);
  if (active) red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < per) {
    float4 t = red[threadIdx.x];
    for (int p2 = 1; p2 < phases; ++p2) {
      const float4 o = red[p2 * per + threadIdx.x];
      t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
    }
    float* dst = out + mi * n + 4 * q;
    atomicAdd(dst, t.x); atomicAdd(dst + 1, t.y); atomicAdd(dst + 2, t.z); atomicAdd(dst + 3, t.w);
  }
}

// ---------------------------------------------------------------------------------------
// tail forward
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tail_fwd_kernel(const float* __restrict__ tc, const float* __restrict__ g, const float* __restrict__ r,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                float* __restrict__ out, float* __restrict__ mean, float* __restrict__ rstd, long long X,
                int C, int T, int npc, float drop_p, unsigned long long seed) {
  extern __shared__ float sm[];
  const int ct = C * T, nst = ct + T;   // padded node stride: column walks hit distinct banks
  float* V = sm;
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const float invC = 1.f / (float)C;
  const long long nchunks = (X + npc - 1) / npc;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long long x0 = ch * npc;
    const int nn = (int)min((long long)npc, X - x0);
    __syncthreads();
    for (int i = threadIdx.x; i < nn * ct; i += blockDim.x) {
      const int xl = i / ct, rem = i - xl * ct;
      const long long e = (x0 + xl) * ct + rem;
      float a = __ldg(tc + e);
      if (drop_p > 0.f) a *= dropout_scale((unsigned long long)e, seed, drop_p, inv_keep);
      if (g) {                                   // g is channels-last (X, T, C)
        const int c = rem / T, t = rem - c * T;
        a += __ldg(g + ((x0 + xl) * T + t) * C + c);
      }
      a = fmaxf(a, 0.f);
      V[xl * nst + rem] = fmaxf(__ldg(r + e) + a, 0.f);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nn * T; i += blockDim.x) {
      const int xl = i / T, t = i - xl * T;
      float* col = V + xl * nst + t;
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += col[c * T];
      const float mu = s * invC;
      float q = 0.f;
      for (int c = 0; c < C; ++c) { const float d = col[c * T] - mu; q = fmaf(d, d, q); }
      const float rs = rsqrtf(q * invC + eps);
      for (int c = 0; c < C; ++c) col[c * T] = (col[c * T] - mu) * rs * __ldg(gamma + c) + __ldg(beta + c);
      mean[(x0 + xl) * T + t] = mu;
      rstd[(x0 + xl) * T + t] = rs;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nn * ct; i += blockDim.x) {
      const int xl = i / ct, rem = i - xl * ct;
      out[(x0 + xl) * ct + rem] = V[xl * nst + rem];
    }
  }
}

// ---------------------------------------------------------------------------------------
// tail backward
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tail_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ tc, const float* __restrict__ g,
                const float* __restrict__ r, const float* __restrict__ gamma,
                const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dtc,
                float* __restrict__ dg, float* __restrict__ dr, float* __restrict__ dgamma,
                float* __restrict__ dbeta, long long X, int C, int T, int npc, float drop_p,
                unsigned long long seed) {
  extern __shared__ float sm[];
  const int ct = C * T, nst = ct + T;
  float* V = sm;                       // v, then unused
  float* D = V + (size_t)npc * nst;    // dout, then dv
  float* Fl = D + (size_t)npc * nst;   // flags: bit0 = (s1 > 0), bit1 = (s2 > 0)
  float* Ag = Fl + (size_t)npc * nst;  // C floats: dgamma partial
  float* Ab = Ag + C;                  // C floats: dbeta partial
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const float invC = 1.f / (float)C;
  const int lane = threadIdx.x & 31;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { Ag[c] = 0.f; Ab[c] = 0.f; }
  const long long nchunks = (X + npc - 1) / npc;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long long x0 = ch * npc;
    const int nn = (int)min((long long)npc, X - x0);
    __syncthreads();
    for (int i = threadIdx.x; i < nn * ct; i += blockDim.x) {
      const int xl = i / ct, rem = i - xl * ct;
      const long long e = (x0 + xl) * ct + rem;
      float a = __ldg(tc + e);
      if (drop_p > 0.f) a *= dropout_scale((unsigned long long)e, seed, drop_p, inv_keep);
      if (g) {
        const int c = rem / T, t = rem - c * T;
        a += __ldg(g + ((x0 + xl) * T + t) * C + c);
      }
      const float s2 = __ldg(r + e) + fmaxf(a, 0.f);
      V[xl * nst + rem] = fmaxf(s2, 0.f);
      D[xl * nst + rem] = __ldg(dout + e);
      Fl[xl * nst + rem] = (a > 0.f ? 1.f : 0.f) + (s2 > 0.f ? 2.f : 0.f);
    }
    __syncthreads();
    const int ncol = nn * T;
    const int ncol_pad = (ncol + 31) & ~31;   // whole warps take part in the shuffles
    for (int i = threadIdx.x; i < ncol_pad; i += blockDim.x) {
      const bool act = i < ncol;
      const int xl = act ? i / T : 0, t = act ? i - xl * T : 0;
      float* vcol = V + xl * nst + t;
      float* dcol = D + xl * nst + t;
      const float mu = act ? __ldg(mean + (x0 + xl) * T + t) : 0.f;
      const float rs = act ? __ldg(rstd + (x0 + xl) * T + t) : 0.f;
      float s1 = 0.f, s2 = 0.f;
      if (act)
        for (int c = 0; c < C; ++c) {
          const float xh = (vcol[c * T] - mu) * rs;
          const float gy = dcol[c * T] * __ldg(gamma + c);
          s1 += gy;
          s2 = fmaf(gy, xh, s2);
        }
      const float m1 = s1 * invC, m2 = s2 * invC;
      for (int c = 0; c < C; ++c) {
        float pa = 0.f, pb = 0.f;
        if (act) {
          const float xh = (vcol[c * T] - mu) * rs;
          const float d = dcol[c * T];
          dcol[c * T] = rs * (d * __ldg(gamma + c) - m1 - xh * m2);
          pa = d * xh;
          pb = d;
        }
        pa = warp_sum(pa);
        pb = warp_sum(pb);
        if (lane == 0) { atomicAdd(Ag + c, pa); atomicAdd(Ab + c, pb); }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nn * ct; i += blockDim.x) {
      const int xl = i / ct, rem = i - xl * ct;
      const long long e = (x0 + xl) * ct + rem;
      const int fl = (int)Fl[xl * nst + rem];
      const float ds2 = (fl & 2) ? D[xl * nst + rem] : 0.f;
      const float ds1 = (fl & 1) ? ds2 : 0.f;
      dr[e] = ds2;
      if (dg) {                                  // stage d(tco) transposed: V is free after the LN pass
        const int c = rem / T, t = rem - c * T;
        V[xl * nst + t * C + c] = ds1;
      }
      float o = ds1;
      if (drop_p > 0.f) o *= dropout_scale((unsigned long long)e, seed, drop_p, inv_keep);
      dtc[e] = o;
    }
    if (dg) {
      __syncthreads();
      for (int i = threadIdx.x; i < nn * ct; i += blockDim.x) {   // coalesced channels-last store
        const int xl = i / ct, rem = i - xl * ct;
        dg[(x0 + xl) * ct + rem] = V[xl * nst + rem];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(dgamma + c, Ag[c]);
    atomicAdd(dbeta + c, Ab[c]);
  }
}

static int pick_npc(size_t per_node_bytes, size_t extra, size_t budget, int cap) {
  long long n = ((long long)budget - (long long)extra) / (long long)per_node_bytes;
  if (n > cap) n = cap;
  return (int)n;
}

static int opt_in_smem(const void* fn, size_t bytes, const char* name) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute failed: %s", name, cudaGetErrorString(e));
      return (int)e;
    }
  }
  return 0;
}

}  // namespace dstag

using namespace dstag;

extern "C" {

int dstag_gtu_gate_fwd(const float* u, float* out, int64_t X, int32_t C, int32_t Tp, int32_t out_ld,
                       int32_t out_off, void* stream) {
  DSTAG_CHECK_ARG(u && out && X > 0 && C > 0 && Tp > 0 && out_off >= 0 && out_off + Tp <= out_ld,
                  "dstag_gtu_gate_fwd: bad arguments");
  const long long total = (long long)X * C * Tp;
  long long blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  gtu_gate_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(u, out, total, C, Tp, out_ld,
                                                                          out_off);
  DSTAG_RETURN_LAUNCH("dstag_gtu_gate_fwd");
}

int dstag_gtu_gate_bwd(const float* u, const float* dout, float* dup, float* db, int64_t X, int32_t C, int32_t Tp,
                       int32_t pad, int32_t out_ld, int32_t out_off, void* stream) {
  DSTAG_CHECK_ARG(u && dout && dup && X > 0 && C > 0 && Tp > 0 && pad >= 0 && out_off >= 0 &&
                      out_off + Tp <= out_ld,
                  "dstag_gtu_gate_bwd: bad arguments");
  const long long total = (long long)X * (Tp + 2 * pad) * C;
  long long blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  gtu_gate_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(u, dout, dup, db, total, C, Tp, pad,
                                                                          out_ld, out_off);
  DSTAG_RETURN_LAUNCH("dstag_gtu_gate_bwd");
}

int dstag_outer_sum_supported(int32_t m, int32_t n) {
  return (m > 0 && n > 0 && n % 4 == 0 && m * (n / 4) <= 256 && (size_t)kOsRows * (m + n) * 4 <= 96 * 1024) ? 1 : 0;
}

int dstag_outer_sum(const float* a, const float* b, float* out, int64_t rows, int32_t m, int32_t n, void* stream) {
  DSTAG_CHECK_ARG(a && b && out && rows > 0, "dstag_outer_sum: bad arguments");
  DSTAG_CHECK_ARG(dstag_outer_sum_supported(m, n), "dstag_outer_sum: unsupported shape m=%d n=%d", m, n);
  DSTAG_CHECK_ARG((reinterpret_cast<uintptr_t>(b) & 15) == 0, "dstag_outer_sum: b must be 16-byte aligned");
  const int per = m * (n / 4);
  const int phases = 256 / per;
  const size_t smem = (size_t)kOsRows * (m + n) * sizeof(float);
  int rc = opt_in_smem((const void*)outer_sum_kernel, smem, "dstag_outer_sum");
  if (rc) return rc;
  const long long nslab = (rows + kOsRows - 1) / kOsRows;
  const int grid = (int)(nslab < kNumSMs * 4 ? nslab : kNumSMs * 4);
  outer_sum_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(a, b, out, rows, m, n, phases);
  DSTAG_RETURN_LAUNCH("dstag_outer_sum");
}

int dstag_tail_fwd(const float* tc, const float* g, const float* r, const float* gamma,
                   const float* beta, float eps, float* out, float* mean, float* rstd, int64_t X,
                   int32_t C, int32_t T, float drop_p, uint64_t seed, void* stream) {
  DSTAG_CHECK_ARG(tc && r && gamma && beta && out && mean && rstd, "dstag_tail_fwd: null pointer");
  DSTAG_CHECK_ARG(X > 0 && C > 0 && T > 0 && drop_p >= 0.f && drop_p < 1.f, "dstag_tail_fwd: bad sizes");
  const size_t per_node = (size_t)(C + 1) * T * sizeof(float);
  const int npc = pick_npc(per_node, 0, 96 * 1024, 16);
  if (npc < 1) { set_error("dstag_tail_fwd: C*T too large for shared memory"); return DSTAG_ERR_SMEM; }
  const size_t smem = per_node * npc;
  int rc = opt_in_smem((const void*)tail_fwd_kernel, smem, "dstag_tail_fwd");
  if (rc) return rc;
  long long chunks = (X + npc - 1) / npc;
  const int grid = (int)(chunks < kNumSMs * 4 ? chunks : kNumSMs * 4);
  tail_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(tc, g, r, gamma, beta, eps, out, mean, rstd,
                                                            X, C, T, npc, drop_p, seed);
  DSTAG_RETURN_LAUNCH("dstag_tail_fwd");
}

int dstag_tail_bwd(const float* dout, const float* tc, const float* g, const float* r,
                   const float* gamma, const float* mean, const float* rstd, float* dtc, float* dg,
                   float* dr, float* dgamma, float* dbeta, int64_t X, int32_t C, int32_t T,
                   float drop_p, uint64_t seed, void* stream) {
  DSTAG_CHECK_ARG(dout && tc && r && gamma && mean && rstd && dtc && dr && dgamma && dbeta,
                  "dstag_tail_bwd: null pointer");
  DSTAG_CHECK_ARG((g == nullptr) == (dg == nullptr), "dstag_tail_bwd: g and dg must both be given or both be null");
  DSTAG_CHECK_ARG(X > 0 && C > 0 && T > 0 && drop_p >= 0.f && drop_p < 1.f, "dstag_tail_bwd: bad sizes");
  const size_t per_node = (size_t)3 * (C + 1) * T * sizeof(float);
  const size_t extra = (size_t)2 * C * sizeof(float);
  const int npc = pick_npc(per_node, extra, 160 * 1024, 16);
  if (npc < 1) { set_error("dstag_tail_bwd: C*T too large for shared memory"); return DSTAG_ERR_SMEM; }
  const size_t smem = per_node * npc + extra;
  int rc = opt_in_smem((const void*)tail_bwd_kernel, smem, "dstag_tail_bwd");
  if (rc) return rc;
  long long chunks = (X + npc - 1) / npc;
  const int grid = (int)(chunks < kNumSMs * 2 ? chunks : kNumSMs * 2);
  tail_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(dout, tc, g, r, gamma, mean, rstd, dtc, dg, dr,
                                                            dgamma, dbeta, X, C, T, npc, drop_p, seed);
  DSTAG_RETURN_LAUNCH("dstag_tail_bwd");
}

}  // extern "C"

// =======================================================================================
// Fused block tail for the PEMS shapes (T = 12, C = 32):
//   tc = fcmy(cat) ; tc' = dropout(tc) ; tco = relu(tc' [+ g]) ; out = LN_C(relu(r + tco))
// One WARP per node (lane = channel): the 24 -> 12 fcmy Linear of model/DSTAGNN_my.py:243 runs on
// register-resident rows against the weight matrix in the constant bank (FFMA with an immediate
// constant operand -- no loads), and the LayerNorm over the 32 channels is a warp-shuffle
// reduction per time step.  tc / dtc never touch HBM.
// =======================================================================================
namespace dstag {

constexpr int kFusedMaxW = 4096;
__constant__ float c_wf[kFusedMaxW];   // fcmy weight (T, S) row-major, re-uploaded per call (stream-ordered)
__constant__ float c_wf2[kFusedMaxW];  // second copy: the backward uses every weight twice (recompute + d cat);
                                       // with ONE symbol the compiler keeps all T*S weights in registers
                                       // (and spills) instead of using constant-bank FFMA operands
__constant__ float c_bf[64];           // fcmy bias (T)

// dropout multipliers of one (node, channel) row: element e = row * T + t draws word (e & 3) of
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
tail_fused_fwd_kernel(const float* __restrict__ cat, const float* __restrict__ g, const float* __restrict__ r,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                      float* __restrict__ out, float* __restrict__ mean, float* __restrict__ rstd, long long X,
                      float drop_p, unsigned long long seed) {
  constexpr int S = 3 * T - 12, C = 32;
  static_assert(S % 4 == 0 && T % 4 == 0 && T <= 32, "fused tail tiling");
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const float gam = __ldg(gamma + lane), bet = __ldg(beta + lane);
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (long long x = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); x < X; x += nwarps) {
    const long long row = x * C + lane;
    float cv[S];
#pragma unroll
    for (int i = 0; i < S / 4; ++i) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(cat + row * S) + i);
      cv[4 * i] = q.x; cv[4 * i + 1] = q.y; cv[4 * i + 2] = q.z; cv[4 * i + 3] = q.w;
    }
    float rv[T];
#pragma unroll
    for (int i = 0; i < T / 4; ++i) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(r + row * T) + i);
      rv[4 * i] = q.x; rv[4 * i + 1] = q.y; rv[4 * i + 2] = q.z; rv[4 * i + 3] = q.w;
    }
    float v[T], ds[T];
#pragma unroll
    for (int t = 0; t < T; ++t) ds[t] = 1.f;
    if (drop_p > 0.f) row_dropout<T>(ds, row, seed, drop_p, inv_keep);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      float a = c_bf[t];
#pragma unroll
      for (int s = 0; s < S; ++s) a = fmaf(cv[s], c_wf[t * S + s], a);
      a *= ds[t];
      if (g) a += __ldg(g + (x * T + t) * C + lane);          // channels-last: coalesced over lanes
      v[t] = fmaxf(rv[t] + fmaxf(a, 0.f), 0.f);
    }
    float o[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float mu = warp_sum(v[t]) * (1.f / C);
      const float d = v[t] - mu;
      const float rs = rsqrtf(warp_sum(d * d) * (1.f / C) + eps);
      o[t] = d * rs * gam + bet;
      if (lane == t) { mean[x * T + t] = mu; rstd[x * T + t] = rs; }
    }
#pragma unroll
    for (int i = 0; i < T / 4; ++i)
      reinterpret_cast<float4*>(out + row * T)[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
  }
}

template <int T>
__global__ void __launch_bounds__(256, 2)
tail_fused_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ cat, const float* __restrict__ g,
                      const float* __restrict__ r, const float* __restrict__ gamma, const float* __restrict__ mean,
                      const float* __restrict__ rstd, float* __restrict__ dcat, float* __restrict__ dtc_out,
                      float* __restrict__ dg, float* __restrict__ dr, float* __restrict__ dgamma,
                      float* __restrict__ dbeta, long long X, float drop_p, unsigned long long seed) {
  constexpr int S = 3 * T - 12, C = 32, NW = 8;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long nwarps = (long long)gridDim.x * NW;
  const float gam = __ldg(gamma + lane);
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  float acc_g = 0.f, acc_be = 0.f;     // d gamma[c], d beta[c] partials (lane = c)
  for (long long x = (long long)blockIdx.x * NW + wid; x < X; x += nwarps) {
    const long long row = x * C + lane;
    float cv[S];
#pragma unroll
    for (int i = 0; i < S / 4; ++i) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(cat + row * S) + i);
      cv[4 * i] = q.x; cv[4 * i + 1] = q.y; cv[4 * i + 2] = q.z; cv[4 * i + 3] = q.w;
    }
    float ds[T];
#pragma unroll
    for (int t = 0; t < T; ++t) ds[t] = 1.f;
    if (drop_p > 0.f) row_dropout<T>(ds, row, seed, drop_p, inv_keep);
    float dtc[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      // ---- recompute the forward of this element
      float a = c_bf[t];
#pragma unroll
      for (int s = 0; s < S; ++s) a = fmaf(cv[s], c_wf[t * S + s], a);
      a *= ds[t];
      if (g) a += __ldg(g + (x * T + t) * C + lane);
      const float s2 = __ldg(r + row * T + t) + fmaxf(a, 0.f);
      const float vv = fmaxf(s2, 0.f);
      // ---- LayerNorm backward over the 32 lanes
      const float mu = __ldg(mean + x * T + t), rs = __ldg(rstd + x * T + t);
      const float xh = (vv - mu) * rs;
      const float d = __ldg(dout + row * T + t);
      const float gy = d * gam;
      const float m1 = warp_sum(gy) * (1.f / C);
      const float m2 = warp_sum(gy * xh) * (1.f / C);
      acc_g = fmaf(d, xh, acc_g);
      acc_be += d;
      const float ds2 = s2 > 0.f ? rs * (gy - m1 - xh * m2) : 0.f;
      const float ds1 = a > 0.f ? ds2 : 0.f;
      dr[row * T + t] = ds2;
      if (dg) dg[(x * T + t) * C + lane] = ds1;                 // channels-last, coalesced over lanes
      dtc[t] = ds1 * ds[t];
    }
#pragma unroll
    for (int i = 0; i < T / 4; ++i)    // d tc feeds the (tensor-core) weight-gradient contraction and d bias
      reinterpret_cast<float4*>(dtc_out + row * T)[i] = make_float4(dtc[4 * i], dtc[4 * i + 1], dtc[4 * i + 2], dtc[4 * i + 3]);
    // ---- d cat[s] = sum_t dtc[t] Wf[t][s]
#pragma unroll
    for (int i = 0; i < S / 4; ++i) {
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = 0.f;
#pragma unroll
        for (int t = 0; t < T; ++t) a = fmaf(dtc[t], c_wf2[t * S + 4 * i + j], a);
        o[j] = a;
      }
      reinterpret_cast<float4*>(dcat + row * S)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  atomicAdd(dgamma + lane, acc_g);
  atomicAdd(dbeta + lane, acc_be);
}

static int upload_fcmy(const float* wf, const float* bf, int T, int S, cudaStream_t stream, const char* name) {
  cudaError_t e = cudaMemcpyToSymbolAsync(c_wf, wf, sizeof(float) * T * S, 0, cudaMemcpyDeviceToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_wf2, wf, sizeof(float) * T * S, 0, cudaMemcpyDeviceToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_bf, bf, sizeof(float) * T, 0, cudaMemcpyDeviceToDevice, stream);
  if (e != cudaSuccess) {
    set_error("%s: constant upload failed: %s", name, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

}  // namespace dstag

extern "C" {

int dstag_tail_fused_supported(int32_t C, int32_t T) { return (C == 32 && T == 12) ? 1 : 0; }

int dstag_tail_fused_fwd(const float* cat, const float* wf, const float* bf, const float* g, const float* r,
                         const float* gamma, const float* beta, float eps, float* out, float* mean, float* rstd,
                         int64_t X, int32_t C, int32_t T, float drop_p, uint64_t seed, void* stream) {
  DSTAG_CHECK_ARG(cat && wf && bf && r && gamma && beta && out && mean && rstd, "dstag_tail_fused_fwd: null pointer");
  DSTAG_CHECK_ARG(X > 0 && drop_p >= 0.f && drop_p < 1.f, "dstag_tail_fused_fwd: bad sizes");
  if (!dstag_tail_fused_supported(C, T)) {
    set_error("dstag_tail_fused_fwd: only C = 32, T = 12 is fused (got C=%d T=%d); use the unfused path", C, T);
    return DSTAG_ERR_UNSUPPORTED;
  }
  int rc = upload_fcmy(wf, bf, T, 3 * T - 12, (cudaStream_t)stream, "dstag_tail_fused_fwd");
  if (rc) return rc;
  const long long blocks = (X + 7) / 8;
  const int grid = (int)(blocks < kNumSMs * 8 ? blocks : kNumSMs * 8);
  tail_fused_fwd_kernel<12><<<grid, 256, 0, (cudaStream_t)stream>>>(cat, g, r, gamma, beta, eps, out, mean, rstd, X,
                                                                    drop_p, seed);
  DSTAG_RETURN_LAUNCH("dstag_tail_fused_fwd");
}

int dstag_tail_fused_bwd(const float* dout, const float* cat, const float* wf, const float* bf, const float* g,
                         const float* r, const float* gamma, const float* mean, const float* rstd, float* dcat,
                         float* dtc, float* dg, float* dr, float* dgamma, float* dbeta, int64_t X, int32_t C,
                         int32_t T, float drop_p, uint64_t seed, void* stream) {
  DSTAG_CHECK_ARG(dout && cat && wf && bf && r && gamma && mean && rstd && dcat && dtc && dr && dgamma && dbeta,
                  "dstag_tail_fused_bwd: null pointer");
  DSTAG_CHECK_ARG((g == nullptr) == (dg == nullptr), "dstag_tail_fused_bwd: g and dg must both be given or both be null");
  DSTAG_CHECK_ARG(X > 0 && drop_p >= 0.f && drop_p < 1.f, "dstag_tail_fused_bwd: bad sizes");
  if (!dstag_tail_fused_supported(C, T)) {
    set_error("dstag_tail_fused_bwd: only C = 32, T = 12 is fused (got C=%d T=%d)", C, T);
    return DSTAG_ERR_UNSUPPORTED;
  }
  int rc = upload_fcmy(wf, bf, T, 3 * T - 12, (cudaStream_t)stream, "dstag_tail_fused_bwd");
  if (rc) return rc;
  const long long blocks = (X + 7) / 8;
  const int grid = (int)(blocks < kNumSMs * 8 ? blocks : kNumSMs * 8);
  tail_fused_bwd_kernel<12><<<grid, 256, 0, (cudaStream_t)stream>>>(dout, cat, g, r, gamma, mean, rstd, dcat, dtc, dg,
                                                                    dr, dgamma, dbeta, X, drop_p, seed);
  DSTAG_RETURN_LAUNCH("dstag_tail_fused_bwd");
}

}  // extern "C"
