// Thin-feature contractions.  The first DSTAGNN block sees F_in = 1 input feature (model/DSTAGNN_my.py:131 with
// in_channels = 1, and the 1x1 residual_conv of line 228): its Theta mix and residual convolution are contractions
// over J = K * F_in <= 4 values per row with 32 output channels.  A 128-row tensor-core tile spends its time on
// pipeline set-up for one k-block there; as streaming kernels they move the (R, C) tensor once:
//   expand:  out[r, c] = act(bias[c] + sum_j a_j[r] * w[j, c])                       -- one write of (R, C)
//   bwd:     d' = d * [y > 0];  dz_j[r] = sum_c d'[r, c] w[j, c];  dw[j, c] = sum_r a_j[r] d'[r, c];  db[c] = sum_r d'
//                                                                                     -- one read of (R, C) (+ y)
// a_j[r] = a[(j / F) * a_plane + r * F + j % F]: `taps` planes of (R, F) rows.  Thread = (row, channel quad).
#include "common.cuh"

namespace dstag {

constexpr int kThinMaxJ = 4;
constexpr int kThinSlices = kNumSMs * 4;     // CTAs of the backward kernel = partial results it leaves

template <int J>
__global__ void __launch_bounds__(256)
thin_expand_kernel(const float* __restrict__ a, long long a_plane, int F, const float* __restrict__ w, long long w_js,
                   long long w_cs, const float* __restrict__ bias, float* __restrict__ out, long long R, int Q,
                   int relu) {
  constexpr int U = 4;
  const int q = threadIdx.x % Q, rl = threadIdx.x / Q, RP = 256 / Q;
  float4 wv[J];
  long long aoff[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const float* wp = w + j * w_js + 4 * q * w_cs;
    wv[j] = make_float4(wp[0], wp[w_cs], wp[2 * w_cs], wp[3 * w_cs]);
    aoff[j] = (long long)(j / F) * a_plane + j % F;
  }
  const float4 bv = bias ? *reinterpret_cast<const float4*>(bias + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const long long C = 4LL * Q;
  for (long long r0 = (long long)blockIdx.x * RP * U; r0 < R; r0 += (long long)gridDim.x * RP * U) {
    float av[U][J];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * RP + rl;
#pragma unroll
      for (int j = 0; j < J; ++j) av[u][j] = r < R ? __ldg(a + aoff[j] + r * F) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * RP + rl;
      float4 o = bv;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        o.x = fmaf(av[u][j], wv[j].x, o.x);
        o.y = fmaf(av[u][j], wv[j].y, o.y);
        o.z = fmaf(av[u][j], wv[j].z, o.z);
        o.w = fmaf(av[u][j], wv[j].w, o.w);
      }
      if (relu) o = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
      if (r < R) *reinterpret_cast<float4*>(out + r * C + 4 * q) = o;
    }
  }
}

template <int J>
__global__ void __launch_bounds__(256)
thin_bwd_kernel(const float* __restrict__ a, long long a_plane, int F, const float* __restrict__ w, long long w_js,
                long long w_cs, const float* __restrict__ d, const float* __restrict__ y, float* __restrict__ dz,
                float* __restrict__ partial, long long R, int Q) {
  constexpr int U = 4;
  __shared__ float4 red[(kThinMaxJ + 1) * 256];
  const int q = threadIdx.x % Q, rl = threadIdx.x / Q, RP = 256 / Q;
  float4 wv[J];
  long long aoff[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const float* wp = w + j * w_js + 4 * q * w_cs;
    wv[j] = make_float4(wp[0], wp[w_cs], wp[2 * w_cs], wp[3 * w_cs]);
    aoff[j] = (long long)(j / F) * a_plane + j % F;
  }
  float4 acc[J + 1];
#pragma unroll
  for (int j = 0; j <= J; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long C = 4LL * Q;
  for (long long r0 = (long long)blockIdx.x * RP * U; r0 < R; r0 += (long long)gridDim.x * RP * U) {
    float4 dv[U], yv[U];
    float av[U][J];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * RP + rl;
      const bool in = r < R;
      dv[u] = in ? __ldg(reinterpret_cast<const float4*>(d + r * C + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
      yv[u] = (in && y) ? __ldg(reinterpret_cast<const float4*>(y + r * C + 4 * q)) : make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
      for (int j = 0; j < J; ++j) av[u][j] = (in && partial) ? __ldg(a + aoff[j] + r * F) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * RP + rl;
      float4 v = dv[u];
      v.x = yv[u].x > 0.f ? v.x : 0.f;
      v.y = yv[u].y > 0.f ? v.y : 0.f;
      v.z = yv[u].z > 0.f ? v.z : 0.f;
      v.w = yv[u].w > 0.f ? v.w : 0.f;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        acc[j].x = fmaf(av[u][j], v.x, acc[j].x);
        acc[j].y = fmaf(av[u][j], v.y, acc[j].y);
        acc[j].z = fmaf(av[u][j], v.z, acc[j].z);
        acc[j].w = fmaf(av[u][j], v.w, acc[j].w);
      }
      acc[J].x += v.x; acc[J].y += v.y; acc[J].z += v.z; acc[J].w += v.w;
      if (dz) {                                   // uniform: every lane takes the shuffles, rows past R carry zeros
#pragma unroll
        for (int j = 0; j < J; ++j) {
          float p = fmaf(v.x, wv[j].x, fmaf(v.y, wv[j].y, fmaf(v.z, wv[j].z, v.w * wv[j].w)));
          for (int o = Q >> 1; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
          if (q == 0 && r < R) dz[aoff[j] + r * F] = p;
        }
      }
    }
  }
  if (!partial) return;
  // fold the RP rows of the CTA in a fixed order: red[j][rl][q]
#pragma unroll
  for (int j = 0; j <= J; ++j) red[(j * RP + rl) * Q + q] = acc[j];
  __syncthreads();
  if (threadIdx.x < (J + 1) * Q) {
    const int j = threadIdx.x / Q, qq = threadIdx.x % Q;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < RP; ++i) {
      const float4 v = red[(j * RP + i) * Q + qq];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    *reinterpret_cast<float4*>(partial + ((long long)blockIdx.x * (J + 1) + j) * C + 4 * qq) = t;
  }
}

static bool thin_shape_ok(int J, int C) {
  const int Q = C / 4;
  return J >= 1 && J <= kThinMaxJ && C >= 4 && C % 4 == 0 && Q <= 32 && (Q & (Q - 1)) == 0;
}

}  // namespace dstag

using namespace dstag;

extern "C" {

int dstag_thin_supported(int32_t J, int32_t C) { return thin_shape_ok(J, C) ? 1 : 0; }
int dstag_thin_slices(void) { return kThinSlices; }

int dstag_thin_expand(const float* a, int64_t a_plane, int32_t taps, int32_t F, const float* w, int64_t w_js,
                      int64_t w_cs, const float* bias, float* out, int64_t R, int32_t C, int32_t relu, void* stream) {
  DSTAG_CHECK_ARG(a && w && out && R > 0 && taps > 0 && F > 0, "dstag_thin_expand: bad arguments");
  const int J = taps * F;
  DSTAG_CHECK_ARG(thin_shape_ok(J, C), "dstag_thin_expand: unsupported shape J=%d C=%d", J, C);
  DSTAG_CHECK_ARG(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0,
                  "dstag_thin_expand: out / bias must be 16-byte aligned");
  const int Q = C / 4, RP = 256 / Q;
  const long long passes = (R + RP * 4 - 1) / (RP * 4);
  const int grid = (int)(passes < kNumSMs * 8 ? passes : kNumSMs * 8);
  cudaStream_t st = (cudaStream_t)stream;
#define DSTAG_THIN_EXPAND(JJ) \
  case JJ: thin_expand_kernel<JJ><<<grid, 256, 0, st>>>(a, a_plane, F, w, w_js, w_cs, bias, out, R, Q, relu); break;
  switch (J) {
    DSTAG_THIN_EXPAND(1) DSTAG_THIN_EXPAND(2) DSTAG_THIN_EXPAND(3) DSTAG_THIN_EXPAND(4)
  }
#undef DSTAG_THIN_EXPAND
  DSTAG_RETURN_LAUNCH("dstag_thin_expand");
}

int dstag_thin_bwd(const float* a, int64_t a_plane, int32_t taps, int32_t F, const float* w, int64_t w_js, int64_t w_cs,
                   const float* d, const float* y, float* dz, float* partial, int64_t R, int32_t C, void* stream) {
  DSTAG_CHECK_ARG(a && w && d && R > 0 && taps > 0 && F > 0 && (dz || partial), "dstag_thin_bwd: bad arguments");
  const int J = taps * F;
  DSTAG_CHECK_ARG(thin_shape_ok(J, C), "dstag_thin_bwd: unsupported shape J=%d C=%d", J, C);
  DSTAG_CHECK_ARG(((reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(y) |
                    reinterpret_cast<uintptr_t>(partial)) & 15) == 0,
                  "dstag_thin_bwd: d / y / partial must be 16-byte aligned");
  const int Q = C / 4;
  cudaStream_t st = (cudaStream_t)stream;
#define DSTAG_THIN_BWD(JJ) \
  case JJ: thin_bwd_kernel<JJ><<<kThinSlices, 256, 0, st>>>(a, a_plane, F, w, w_js, w_cs, d, y, dz, partial, R, Q); break;
  switch (J) {
    DSTAG_THIN_BWD(1) DSTAG_THIN_BWD(2) DSTAG_THIN_BWD(3) DSTAG_THIN_BWD(4)
  }
#undef DSTAG_THIN_BWD
  DSTAG_RETURN_LAUNCH("dstag_thin_bwd");
}

}  // extern "C"
