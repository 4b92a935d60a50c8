// Generic strided tensor contraction on CUDA cores (exact fp32 FMA).
//
// This is the shape-agnostic engine behind dstag_gemm(): any operand layout that the
// two-level axis description can express (permuted views, unfolded conv windows,
// padded leading dimensions, broadcast batch) is read in place.  The tcgen05 engine in
// gemm_tc.cu takes over the large contractions; this kernel is the exact-fp32 path and
// serves every shape the tensor-core engine declines.
#include "common.cuh"

namespace dstag {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

__device__ __forceinline__ long long axoff(const dstag_axis_t& a, int idx) {
  if (a.d1 == 1) return (long long)idx * a.s0;
  const int hi = idx / a.d1;
  const int lo = idx - hi * a.d1;
  return (long long)hi * a.s0 + (long long)lo * a.s1;
}

template <bool A_FAST_M, bool B_FAST_N>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const dstag_gemm_t p) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int batch = blockIdx.z / p.split_k;
  const int ks = blockIdx.z - batch * p.split_k;

  const int ktiles = (p.K + BK - 1) / BK;
  const int per_split = (ktiles + p.split_k - 1) / p.split_k;
  const int kt0 = ks * per_split;
  const int kt1 = min(ktiles, kt0 + per_split);
  if (kt0 >= kt1 && p.split_k > 1 && p.split_stride == 0) return;  // empty K slice of an atomic accumulation
                                                                   // (a slice of a two-stage reduction is written as zeros)

  const float* __restrict__ A = (const float*)p.A.ptr + axoff(p.A.batch, batch);
  const float* __restrict__ Bp = (const float*)p.B.ptr + axoff(p.B.batch, batch);
  const float* __restrict__ Am =
      p.Amul.ptr ? (const float*)p.Amul.ptr + axoff(p.Amul.batch, batch) : nullptr;

  // ---- global -> register staging maps -------------------------------------------------
  int a_ml[4], a_kl[4], b_nl[4], b_kl[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (A_FAST_M) { a_ml[r] = tid & 63; a_kl[r] = (tid >> 6) + 4 * r; }
    else          { a_kl[r] = tid & 15; a_ml[r] = (tid >> 4) + 16 * r; }
    if (B_FAST_N) { b_nl[r] = tid & 63; b_kl[r] = (tid >> 6) + 4 * r; }
    else          { b_kl[r] = tid & 15; b_nl[r] = (tid >> 4) + 16 * r; }
  }
  long long a_moff[4], am_moff[4], b_noff[4];
  bool a_mok[4], b_nok[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int m = m0 + a_ml[r], n = n0 + b_nl[r];
    a_mok[r] = m < p.M;
    b_nok[r] = n < p.N;
    a_moff[r] = a_mok[r] ? axoff(p.A.row, m) : 0;
    am_moff[r] = (Am && a_mok[r]) ? axoff(p.Amul.row, m) : 0;
    b_noff[r] = b_nok[r] ? axoff(p.B.col, n) : 0;
  }

  float ra[4], rb[4];
  auto load_tile = [&](int kt) {
    const int kb = kt * BK;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ka = kb + a_kl[r];
      float v = 0.f;
      if (a_mok[r] && ka < p.K) {
        v = __ldg(A + a_moff[r] + axoff(p.A.col, ka));
        if (Am) v *= __ldg(Am + am_moff[r] + axoff(p.Amul.col, ka));
      }
      ra[r] = v;
      const int kk = kb + b_kl[r];
      float w = 0.f;
      if (b_nok[r] && kk < p.K) w = __ldg(Bp + b_noff[r] + axoff(p.B.row, kk));
      rb[r] = w;
    }
  };

  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  if (kt0 < kt1) load_tile(kt0);
  for (int kt = kt0; kt < kt1; ++kt) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      As[a_kl[r]][a_ml[r]] = ra[r];
      Bs[b_kl[r]][b_nl[r]] = rb[r];
    }
    __syncthreads();
    if (kt + 1 < kt1) load_tile(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ------------------------------------------------------------------------
  float* __restrict__ C = (float*)p.C.ptr + axoff(p.C.batch, batch) + (long long)ks * p.split_stride;
  const float* __restrict__ Cm =
      p.Cmul.ptr ? (const float*)p.Cmul.ptr + axoff(p.Cmul.batch, batch) : nullptr;
  const float* __restrict__ Ca =
      (p.Cadd.ptr && ks == 0) ? (const float*)p.Cadd.ptr + axoff(p.Cadd.batch, batch) : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    const long long cmo = axoff(p.C.row, m);
    const long long mmo = Cm ? axoff(p.Cmul.row, m) : 0;
    const long long amo = Ca ? axoff(p.Cadd.row, m) : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = p.alpha * acc[i][j];
      if (Ca) v += __ldg(Ca + amo + axoff(p.Cadd.col, n));
      if (Cm) v *= __ldg(Cm + mmo + axoff(p.Cmul.col, n));
      if (p.relu) v = fmaxf(v, 0.f);
      float* dst = C + cmo + axoff(p.C.col, n);
      if (p.accumulate == DSTAG_ACC_STORE) *dst = v;
      else if (p.accumulate == DSTAG_ACC_ADD) *dst += v;
      else atomicAdd(dst, v);
    }
  }
}

static inline long long inner_stride(const dstag_axis_t& a) { return a.d1 == 1 ? a.s0 : a.s1; }

int gemm_simt(const dstag_gemm_t* g, cudaStream_t stream) {
  const dim3 grid((g->M + BM - 1) / BM, (g->N + BN - 1) / BN, g->batch * g->split_k);
  DSTAG_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535,
                  "dstag_gemm: grid too large (N tiles %u, batch*split %u)", grid.y, grid.z);
  const bool a_fast_m = inner_stride(g->A.row) == 1;
  const bool b_fast_n = inner_stride(g->B.col) == 1;
  if (a_fast_m && b_fast_n) gemm_simt_kernel<true, true><<<grid, 256, 0, stream>>>(*g);
  else if (a_fast_m) gemm_simt_kernel<true, false><<<grid, 256, 0, stream>>>(*g);
  else if (b_fast_n) gemm_simt_kernel<false, true><<<grid, 256, 0, stream>>>(*g);
  else gemm_simt_kernel<false, false><<<grid, 256, 0, stream>>>(*g);
  DSTAG_RETURN_LAUNCH("dstag_gemm(simt)");
}

}  // namespace dstag
