// Temporal attention core: fused  Q K^T / sqrt(dk) + res  ->  softmax over the QUERY axis
// -> P V, one CTA per (batch, feature, head) tile.  T is 12 (PEMS) or 144 (GAMBIA), dk = 32,
// so the whole tile lives in shared memory and the logits / probabilities never round-trip
// through HBM except for the logits the reference returns (model/DSTAGNN_my.py:37-42).
#include "common.cuh"

namespace dstag {

// qkv (B,F,T,3,H,dk): offset of (t, s, d) for a fixed (b,f,h)
__device__ __forceinline__ long long qkv_off(long long bf, int t, int s, int h, int d, int T, int H,
                                             int dk) {
  return (((bf * T + t) * 3 + s) * H + h) * (long long)dk + d;
}

__global__ void tattn_fwd_kernel(const float* __restrict__ qkv, const float* __restrict__ res,
                                 int res_F, float* __restrict__ scores, float* __restrict__ ctx, int F,
                                 int H, int T, int dk) {
  extern __shared__ float sm[];
  const int ldk = dk + 1, ldt = T + 1;
  float* Qs = sm;
  float* Ks = Qs + T * ldk;
  float* Vs = Ks + T * ldk;
  float* Ss = Vs + T * ldk;  // T x ldt
  const int h = blockIdx.x % H;
  const long long bf = blockIdx.x / H;  // b * F + f
  const int b = (int)(bf / F), f = (int)(bf % F);
  const int tid = threadIdx.x, nt = blockDim.x;

  for (int i = tid; i < T * dk; i += nt) {
    const int t = i / dk, d = i - t * dk;
    Qs[t * ldk + d] = __ldg(qkv + qkv_off(bf, t, 0, h, d, T, H, dk));
    Ks[t * ldk + d] = __ldg(qkv + qkv_off(bf, t, 1, h, d, T, H, dk));
    Vs[t * ldk + d] = __ldg(qkv + qkv_off(bf, t, 2, h, d, T, H, dk));
  }
  __syncthreads();

  const float scale = rsqrtf((float)dk);
  const float* rp = nullptr;
  if (res_F > 0) rp = res + (((long long)b * res_F + (res_F == 1 ? 0 : f)) * H + h) * (long long)T * T;
  float* sp = scores + (bf * H + h) * (long long)T * T;
  for (int i = tid; i < T * T; i += nt) {
    const int q = i / T, k = i - q * T;
    float acc = 0.f;
    for (int d = 0; d < dk; ++d) acc = fmaf(Qs[q * ldk + d], Ks[k * ldk + d], acc);
    acc *= scale;
    if (rp) acc += __ldg(rp + i);
    Ss[q * ldt + k] = acc;
    sp[i] = acc;
  }
  __syncthreads();

  // softmax over q for every key column k (reference dim=3 of (B,F,H,Tq,Tk))
  for (int k = tid; k < T; k += nt) {
    float mx = -INFINITY;
    for (int q = 0; q < T; ++q) mx = fmaxf(mx, Ss[q * ldt + k]);
    float s = 0.f;
    for (int q = 0; q < T; ++q) {
      const float e = __expf(Ss[q * ldt + k] - mx);
      Ss[q * ldt + k] = e;
      s += e;
    }
    const float inv = 1.f / s;
    for (int q = 0; q < T; ++q) Ss[q * ldt + k] *= inv;
  }
  __syncthreads();

  // ctx (B,F,T,H,dk)
  for (int i = tid; i < T * dk; i += nt) {
    const int q = i / dk, d = i - q * dk;
    float acc = 0.f;
    for (int k = 0; k < T; ++k) acc = fmaf(Ss[q * ldt + k], Vs[k * ldk + d], acc);
    ctx[((bf * T + q) * H + h) * (long long)dk + d] = acc;
  }
}

__global__ void tattn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ scores,
                                 const float* __restrict__ dctx, const float* __restrict__ dsext,
                                 float* __restrict__ dqkv, float* __restrict__ dsout, int F, int H, int T,
                                 int dk) {
  extern __shared__ float sm[];
  const int ldk = dk + 1, ldt = T + 1;
  float* Xa = sm;                // phase 1: V      phase 2: Q
  float* Xb = Xa + T * ldk;      // phase 1: dctx   phase 2: K
  float* Ps = Xb + T * ldk;      // probabilities
  float* Ds = Ps + T * ldt;      // dP, then dS
  const int h = blockIdx.x % H;
  const long long bf = blockIdx.x / H;
  const int tid = threadIdx.x, nt = blockDim.x;
  const long long sbase = (bf * H + h) * (long long)T * T;

  for (int i = tid; i < T * dk; i += nt) {
    const int t = i / dk, d = i - t * dk;
    Xa[t * ldk + d] = __ldg(qkv + qkv_off(bf, t, 2, h, d, T, H, dk));
    Xb[t * ldk + d] = __ldg(dctx + ((bf * T + t) * H + h) * (long long)dk + d);
  }
  for (int i = tid; i < T * T; i += nt) {
    const int q = i / T, k = i - q * T;
    Ps[q * ldt + k] = __ldg(scores + sbase + i);
  }
  __syncthreads();
  for (int k = tid; k < T; k += nt) {
    float mx = -INFINITY;
    for (int q = 0; q < T; ++q) mx = fmaxf(mx, Ps[q * ldt + k]);
    float s = 0.f;
    for (int q = 0; q < T; ++q) {
      const float e = __expf(Ps[q * ldt + k] - mx);
      Ps[q * ldt + k] = e;
      s += e;
    }
    const float inv = 1.f / s;
    for (int q = 0; q < T; ++q) Ps[q * ldt + k] *= inv;
  }
  __syncthreads();

  // dV[k,d] = sum_q P[q,k] dctx[q,d]
  for (int i = tid; i < T * dk; i += nt) {
    const int k = i / dk, d = i - k * dk;
    float acc = 0.f;
    for (int q = 0; q < T; ++q) acc = fmaf(Ps[q * ldt + k], Xb[q * ldk + d], acc);
    dqkv[qkv_off(bf, k, 2, h, d, T, H, dk)] = acc;
  }
  // dP[q,k] = sum_d dctx[q,d] V[k,d]
  for (int i = tid; i < T * T; i += nt) {
    const int q = i / T, k = i - q * T;
    float acc = 0.f;
    for (int d = 0; d < dk; ++d) acc = fmaf(Xb[q * ldk + d], Xa[k * ldk + d], acc);
    Ds[q * ldt + k] = acc;
  }
  __syncthreads();
  // softmax backward along q per column k, plus the gradient arriving through the logits
  for (int k = tid; k < T; k += nt) {
    float s = 0.f;
    for (int q = 0; q < T; ++q) s = fmaf(Ps[q * ldt + k], Ds[q * ldt + k], s);
    for (int q = 0; q < T; ++q) {
      float g = Ps[q * ldt + k] * (Ds[q * ldt + k] - s);
      if (dsext) g += __ldg(dsext + sbase + (long long)q * T + k);
      Ds[q * ldt + k] = g;
    }
  }
  __syncthreads();
  for (int i = tid; i < T * T; i += nt) {
    const int q = i / T, k = i - q * T;
    dsout[sbase + i] = Ds[q * ldt + k];
  }
  // phase 2: Q, K replace V, dctx
  for (int i = tid; i < T * dk; i += nt) {
    const int t = i / dk, d = i - t * dk;
    Xa[t * ldk + d] = __ldg(qkv + qkv_off(bf, t, 0, h, d, T, H, dk));
    Xb[t * ldk + d] = __ldg(qkv + qkv_off(bf, t, 1, h, d, T, H, dk));
  }
  __syncthreads();
  const float scale = rsqrtf((float)dk);
  for (int i = tid; i < T * dk; i += nt) {
    const int t = i / dk, d = i - t * dk;
    float aq = 0.f, ak = 0.f;
    for (int j = 0; j < T; ++j) {
      aq = fmaf(Ds[t * ldt + j], Xb[j * ldk + d], aq);   // dQ[t] = sum_k dS[t,k] K[k]
      ak = fmaf(Ds[j * ldt + t], Xa[j * ldk + d], ak);   // dK[t] = sum_q dS[q,t] Q[q]
    }
    dqkv[qkv_off(bf, t, 0, h, d, T, H, dk)] = aq * scale;
    dqkv[qkv_off(bf, t, 1, h, d, T, H, dk)] = ak * scale;
  }
}

static int set_smem(const void* fn, size_t bytes, const char* name) {
  if (bytes > 227 * 1024) {
    set_error("%s: T too large for the shared-memory tile (%zu bytes)", name, bytes);
    return DSTAG_ERR_SMEM;
  }
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

int dstag_tattn_fwd(const float* qkv, const float* res, int32_t res_F, float* scores, float* ctx,
                    int32_t B, int32_t F, int32_t H, int32_t T, int32_t dk, void* stream) {
  DSTAG_CHECK_ARG(qkv && scores && ctx, "dstag_tattn_fwd: null pointer");
  DSTAG_CHECK_ARG(B > 0 && F > 0 && H > 0 && T > 0 && dk > 0, "dstag_tattn_fwd: bad sizes");
  DSTAG_CHECK_ARG(res_F == 0 || res_F == 1 || res_F == F, "dstag_tattn_fwd: res_F must be 0, 1 or F");
  DSTAG_CHECK_ARG(res_F == 0 || res != nullptr, "dstag_tattn_fwd: res missing");
  const size_t smem = ((size_t)3 * T * (dk + 1) + (size_t)T * (T + 1)) * sizeof(float);
  int rc = set_smem((const void*)tattn_fwd_kernel, smem, "dstag_tattn_fwd");
  if (rc) return rc;
  const int threads = T <= 16 ? 128 : 256;
  tattn_fwd_kernel<<<(unsigned)((long long)B * F * H), threads, smem, (cudaStream_t)stream>>>(
      qkv, res_F ? res : nullptr, res_F, scores, ctx, F, H, T, dk);
  DSTAG_RETURN_LAUNCH("dstag_tattn_fwd");
}

int dstag_tattn_bwd(const float* qkv, const float* scores, const float* dctx,
                    const float* dscores_ext, float* dqkv, float* dscores_out, int32_t B, int32_t F,
                    int32_t H, int32_t T, int32_t dk, void* stream) {
  DSTAG_CHECK_ARG(qkv && scores && dctx && dqkv && dscores_out, "dstag_tattn_bwd: null pointer");
  DSTAG_CHECK_ARG(B > 0 && F > 0 && H > 0 && T > 0 && dk > 0, "dstag_tattn_bwd: bad sizes");
  const size_t smem = ((size_t)2 * T * (dk + 1) + (size_t)2 * T * (T + 1)) * sizeof(float);
  int rc = set_smem((const void*)tattn_bwd_kernel, smem, "dstag_tattn_bwd");
  if (rc) return rc;
  const int threads = T <= 16 ? 128 : 256;
  tattn_bwd_kernel<<<(unsigned)((long long)B * F * H), threads, smem, (cudaStream_t)stream>>>(
      qkv, scores, dctx, dscores_ext, dqkv, dscores_out, F, H, T, dk);
  DSTAG_RETURN_LAUNCH("dstag_tattn_bwd");
}

}  // extern "C"
