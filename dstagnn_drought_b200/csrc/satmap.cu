// Spatial attention maps, forward: scores + bias + softmax fused (SMultiHeadAttention /
// SScaledDotProductAttention, model/DSTAGNN_my.py:13-22, 62-66, and the softmax of cheb_conv_withSAt,
// lines 126-127), producing the K per-order N x N probability maps directly in the transposed,
// padded layout the Chebyshev contraction consumes:
//
//   P[b, k, j, i] = softmax_i( Qs[b,i,k,:] . Ks[b,j,k,:] / sqrt(dk) + biasT[k, j, i] )
//
// A CTA owns a strip of R output nodes j of one (b, k): the R x N logits live in shared memory
// (never in HBM), the Qs tiles stream through a second buffer, and every row is normalised and
// written once -- N x N x 4 bytes of HBM traffic per map instead of 3x that for a GEMM + softmax pair.
#include "common.cuh"

namespace dstag {

constexpr int kSatThreads = 256;
constexpr int kSatTileI = 128;

constexpr int kSatQld = kSatTileI + 4;   // padded row of the transposed Qs tile (keeps float4 alignment)

template <int R>
__global__ void __launch_bounds__(kSatThreads)
sat_softmax_fwd_kernel(const float* __restrict__ qk, const float* __restrict__ biasT, float* __restrict__ P,
                       int N, int K, int dk, int ld, float scale) {
  static_assert(R % 8 == 0 && R <= 32, "strip height");
  constexpr int MR = R / 8;             // rows per thread: 8 row groups (one per warp) x 32 column quads
  extern __shared__ __align__(16) float sm[];
  float* KsT = sm;                      // [dk][R]       transposed W_K rows of the strip
  float* QsT = KsT + dk * R;            // [dk][kSatQld] transposed W_Q rows of the current column tile
  float* S = QsT + dk * kSatQld;        // [R][ld]       logits, then exponentials
  const int bk = blockIdx.y, b = bk / K, k = bk - b * K;
  const int j0 = blockIdx.x * R;
  const int tid = threadIdx.x;
  const long long nstride = 2LL * K * dk;                        // qk: (B, N, 2, K, dk)
  const float* qbase = qk + (long long)b * N * nstride + (long long)k * dk;          // W_Q projection
  const float* kbase = qbase + (long long)K * dk;                                     // W_K projection

  for (int e = tid; e < R * dk; e += kSatThreads) {
    const int r = e / dk, d = e - r * dk;
    KsT[d * R + r] = (j0 + r < N) ? __ldg(kbase + (long long)(j0 + r) * nstride + d) : 0.f;
  }
  const int iq = tid & 31, rgp = tid >> 5;                        // column quad, row group (= warp)
  for (int i0 = 0; i0 < N; i0 += kSatTileI) {
    __syncthreads();
    for (int e = tid; e < kSatTileI * dk; e += kSatThreads) {
      const int il = e / dk, d = e - il * dk;
      QsT[d * kSatQld + il] = (i0 + il < N) ? __ldg(qbase + (long long)(i0 + il) * nstride + d) : 0.f;
    }
    __syncthreads();
    float acc[MR][4];
#pragma unroll
    for (int rr = 0; rr < MR; ++rr) acc[rr][0] = acc[rr][1] = acc[rr][2] = acc[rr][3] = 0.f;
#pragma unroll 8
    for (int d = 0; d < dk; ++d) {
      const float4 q = *reinterpret_cast<const float4*>(QsT + d * kSatQld + 4 * iq);
      float kvv[MR];                                              // warp-uniform rows: one broadcast load
      if (MR == 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(KsT + d * R + rgp * MR);
        kvv[0] = t4.x; kvv[1 % MR] = t4.y; kvv[2 % MR] = t4.z; kvv[3 % MR] = t4.w;
      } else if (MR == 2) {
        const float2 t2 = *reinterpret_cast<const float2*>(KsT + d * R + rgp * MR);
        kvv[0] = t2.x; kvv[1 % MR] = t2.y;
      } else {
        kvv[0] = KsT[d * R + rgp * MR];
      }
#pragma unroll
      for (int rr = 0; rr < MR; ++rr) {
        const float kv = kvv[rr];
        acc[rr][0] = fmaf(q.x, kv, acc[rr][0]);
        acc[rr][1] = fmaf(q.y, kv, acc[rr][1]);
        acc[rr][2] = fmaf(q.z, kv, acc[rr][2]);
        acc[rr][3] = fmaf(q.w, kv, acc[rr][3]);
      }
    }
    const int i = i0 + 4 * iq;
    if (i < ld) {
#pragma unroll
      for (int rr = 0; rr < MR; ++rr) {
        const int r = rgp * MR + rr, j = j0 + r;
        float4 v = make_float4(acc[rr][0] * scale, acc[rr][1] * scale, acc[rr][2] * scale, acc[rr][3] * scale);
        if (j < N) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(biasT + ((long long)k * N + j) * ld + i));
          v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        }
        *reinterpret_cast<float4*>(S + r * ld + i) = v;
      }
    }
  }
  __syncthreads();
  // row softmax: each warp normalises R / 8 rows
  const int lane = tid & 31, w = tid >> 5;
  for (int r = w; r < R; r += kSatThreads / 32) {
    const int j = j0 + r;
    if (j >= N) continue;
    float* row = S + r * ld;
    float mx = -INFINITY;
    for (int c = lane; c < N; c += 32) mx = fmaxf(mx, row[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < N; c += 32) {
      const float e = __expf(row[c] - mx);
      row[c] = e;
      s += e;
    }
    const float inv = 1.f / warp_sum(s);
    float* out = P + (((long long)b * K + k) * N + j) * ld;
    for (int c = lane; c < ld; c += 32) out[c] = c < N ? row[c] * inv : 0.f;
  }
}

}  // namespace dstag

using namespace dstag;

extern "C" int dstag_sat_softmax_fwd(const float* qk, const float* biasT, float* P, int32_t B, int32_t N, int32_t K,
                                     int32_t dk, int32_t ld, void* stream) {
  DSTAG_CHECK_ARG(qk && biasT && P, "dstag_sat_softmax_fwd: null pointer");
  DSTAG_CHECK_ARG(B > 0 && N > 0 && K > 0 && dk > 0 && ld >= N && (long long)B * K <= 65535,
                  "dstag_sat_softmax_fwd: bad sizes");
  DSTAG_CHECK_ARG(ld % 4 == 0 && (reinterpret_cast<uintptr_t>(biasT) & 15) == 0,
                  "dstag_sat_softmax_fwd: ld must be a multiple of 4 and biasT 16-byte aligned");
  const size_t fixed = (size_t)(dk * kSatQld) * sizeof(float);
  auto need = [&](int R) { return fixed + (size_t)R * (dk + ld) * sizeof(float); };
  const float scale = 1.0f / sqrtf((float)dk);
  const size_t budget = 200 * 1024;
#define DSTAG_SAT_LAUNCH(R_)                                                                                     \
  do {                                                                                                           \
    const size_t smem = need(R_);                                                                                \
    cudaError_t e = cudaFuncSetAttribute(sat_softmax_fwd_kernel<R_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem);                                                             \
    if (e != cudaSuccess) {                                                                                      \
      set_error("dstag_sat_softmax_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));                \
      return (int)e;                                                                                             \
    }                                                                                                            \
    const dim3 grid((N + R_ - 1) / R_, B * K);                                                                   \
    sat_softmax_fwd_kernel<R_><<<grid, kSatThreads, smem, (cudaStream_t)stream>>>(qk, biasT, P, N, K, dk, ld, scale); \
    DSTAG_RETURN_LAUNCH("dstag_sat_softmax_fwd");                                                                \
  } while (0)
  if (2 * need(32) <= budget + 24 * 1024) DSTAG_SAT_LAUNCH(32);   // two CTAs per SM preferred
  if (2 * need(16) <= budget + 24 * 1024) DSTAG_SAT_LAUNCH(16);
  if (need(32) <= budget) DSTAG_SAT_LAUNCH(32);
  if (need(16) <= budget) DSTAG_SAT_LAUNCH(16);
  // Strips of 8 rows would fit up to N ~ 6000 but leave one FMA per shared-memory load: the
  // tensor-core logits GEMM + row softmax pair is faster there (measured at N = 4096).
  set_error("dstag_sat_softmax_fwd: N = %d too large for a 16-row shared-memory strip; use dstag_gemm + dstag_softmax_fwd", N);
  return DSTAG_ERR_SMEM;
#undef DSTAG_SAT_LAUNCH
}
