// Spatial attention maps, forward: scores + bias + softmax fused (SMultiHeadAttention /
// SScaledDotProductAttention, model/DSTAGNN_my.py:13-22, 62-66, and the softmax of cheb_conv_withSAt,
// lines 126-127), producing the K per-order N x N probability maps directly in the transposed,
// padded layout the Chebyshev contraction consumes:
//
//   P[b, k, j, i] = softmax_i( Qs[b,i,k,:] . Ks[b,j,k,:] / sqrt(dk) + biasT[k, j, i] )
//
// A CTA owns a strip of output nodes j of one (b, k); the logits of the strip never reach HBM and every
// row is normalised and written once -- N x N x 4 bytes of HBM traffic per map instead of 3x that for a
// GEMM + softmax pair.  Three variants, chosen by the launcher at the bottom:
//   sat_softmax_mma_kernel   d_k = 32, ld <= 1024 (the PEMS graphs; GAMBIA's N = 2139 takes the strip kernel):
//                            logits as warp-level tensor-core MMA
//                            fragments (3xTF32), bias / T_k rows and W_Q pieces by cp.async
//   sat_softmax_rows_kernel  other d_k, ld <= 1024: a warp keeps 4 complete logit rows in registers (SIMT)
//   sat_softmax_fwd_kernel   ld > 1024: R x N logits of the strip in shared memory (SIMT)
#include "common.cuh"

namespace dstag {

constexpr int kSatThreads = 256;
constexpr int kSatTileI = 128;

constexpr int kSatQld = kSatTileI + 4;   // padded row of the transposed Qs tile (keeps float4 alignment)

template <int R>
__global__ void __launch_bounds__(kSatThreads)
sat_softmax_fwd_kernel(const float* __restrict__ qk, const float* __restrict__ biasT, const float* __restrict__ tkT,
                       float* __restrict__ P, float* __restrict__ TP, int N, int K, int dk, int ld, float scale) {
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
    const long long orow = (((long long)b * K + k) * N + j) * ld;
    const float* trow = tkT + ((long long)k * N + j) * ld;
    for (int c = lane; c < ld; c += 32) {
      const float pv = c < N ? row[c] * inv : 0.f;
      P[orow + c] = pv;
      if (TP) TP[orow + c] = pv * __ldg(trow + c);
    }
  }
}


// ---------------------------------------------------------------------------------------
// Register-resident variant for ld <= 1024 (all PEMS graph sizes; not the N = 2139 drought graph): a warp owns 4 output
// rows j and keeps their complete logit rows in registers (4 x 4*NT values per lane, NT column tiles
// of 128), so the softmax needs no shared-memory round trip at all: 16 FMAs per pair of
// shared-memory loads in the product loop, row max / sum by warp shuffles, one float4 store per value.
// CTA = 4 warps = 16 rows of one (b, k); the W_Q projection streams through shared memory in
// transposed 128-column tiles.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float4 mul4s(float4 a, float4 b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
constexpr int kSatRowsCta = 16;
constexpr int kSatRowThreads = 128;

template <int NT>
__global__ void __launch_bounds__(kSatRowThreads)
sat_softmax_rows_kernel(const float* __restrict__ qk, const float* __restrict__ biasT, const float* __restrict__ tkT,
                        float* __restrict__ P, float* __restrict__ TP, int N, int K, int dk, int ld, float scale) {
  extern __shared__ __align__(16) float sm[];
  float* KsT = sm;                               // [dk][16]
  float* QsT = sm + dk * kSatRowsCta;            // [dk][kSatQld]
  const int bk = blockIdx.y, b = bk / K, k = bk - b * K;
  const int j0 = blockIdx.x * kSatRowsCta;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const long long nstride = 2LL * K * dk;                        // qk: (B, N, 2, K, dk)
  const float* qbase = qk + (long long)b * N * nstride + (long long)k * dk;
  const float* kbase = qbase + (long long)K * dk;
  const int dq = dk >> 2;                                         // float4 per projection row (dk % 4 == 0)

  for (int e = tid; e < kSatRowsCta * dq; e += kSatRowThreads) {
    const int r = e / dq, d4 = e - r * dq;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j0 + r < N) v = __ldg(reinterpret_cast<const float4*>(kbase + (long long)(j0 + r) * nstride) + d4);
    KsT[(4 * d4 + 0) * kSatRowsCta + r] = v.x;
    KsT[(4 * d4 + 1) * kSatRowsCta + r] = v.y;
    KsT[(4 * d4 + 2) * kSatRowsCta + r] = v.z;
    KsT[(4 * d4 + 3) * kSatRowsCta + r] = v.w;
  }
  float acc[4][4 * NT];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4 * NT; ++c) acc[r][c] = 0.f;

#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int i0 = t * kSatTileI;
    __syncthreads();
    for (int e = tid; e < kSatTileI * dq; e += kSatRowThreads) {
      const int il = e / dq, d4 = e - il * dq;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i0 + il < N) v = __ldg(reinterpret_cast<const float4*>(qbase + (long long)(i0 + il) * nstride) + d4);
      QsT[(4 * d4 + 0) * kSatQld + il] = v.x;
      QsT[(4 * d4 + 1) * kSatQld + il] = v.y;
      QsT[(4 * d4 + 2) * kSatQld + il] = v.z;
      QsT[(4 * d4 + 3) * kSatQld + il] = v.w;
    }
    __syncthreads();
#pragma unroll 4
    for (int d = 0; d < dk; ++d) {
      const float4 q = *reinterpret_cast<const float4*>(QsT + d * kSatQld + 4 * lane);
      const float4 kv = *reinterpret_cast<const float4*>(KsT + d * kSatRowsCta + 4 * w);   // warp-uniform: broadcast
      const float kr[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        acc[r][4 * t + 0] = fmaf(q.x, kr[r], acc[r][4 * t + 0]);
        acc[r][4 * t + 1] = fmaf(q.y, kr[r], acc[r][4 * t + 1]);
        acc[r][4 * t + 2] = fmaf(q.z, kr[r], acc[r][4 * t + 2]);
        acc[r][4 * t + 3] = fmaf(q.w, kr[r], acc[r][4 * t + 3]);
      }
    }
  }
  // logits = acc * scale + bias; softmax over the row; columns [N, ld) are written as zeros
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int j = j0 + 4 * w + r;
    if (j >= N) continue;                                          // warp-uniform
    const float* brow = biasT + ((long long)k * N + j) * ld;
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int i = t * kSatTileI + 4 * lane;
      float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < ld) bb = __ldg(reinterpret_cast<const float4*>(brow + i));
      const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float v = (i + c < N) ? fmaf(acc[r][4 * t + c], scale, bv[c]) : -INFINITY;
        acc[r][4 * t + c] = v;
        mx = fmaxf(mx, v);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 4 * NT; ++c) {
      const float e = __expf(acc[r][c] - mx);                      // exp(-inf) = 0 for the masked columns
      acc[r][c] = e;
      sum += e;
    }
    const float inv = 1.f / warp_sum(sum);
    const long long orow = (((long long)b * K + k) * N + j) * ld;
    const float* trow = tkT + ((long long)k * N + j) * ld;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int i = t * kSatTileI + 4 * lane;
      if (i < ld) {
        const float4 pv = make_float4(acc[r][4 * t] * inv, acc[r][4 * t + 1] * inv, acc[r][4 * t + 2] * inv,
                                      acc[r][4 * t + 3] * inv);
        *reinterpret_cast<float4*>(P + orow + i) = pv;
        if (TP) *reinterpret_cast<float4*>(TP + orow + i) = mul4s(pv, __ldg(reinterpret_cast<const float4*>(trow + i)));
      }
    }
  }
}

template <int NT>
static int launch_sat_rows(const float* qk, const float* biasT, const float* tkT, float* P, float* TP, int B, int N,
                           int K, int dk, int ld, float scale, cudaStream_t stream) {
  const size_t smem = (size_t)dk * (kSatRowsCta + kSatQld) * sizeof(float);
  const dim3 grid((N + kSatRowsCta - 1) / kSatRowsCta, B * K);
  sat_softmax_rows_kernel<NT><<<grid, kSatRowThreads, smem, stream>>>(qk, biasT, tkT, P, TP, N, K, dk, ld, scale);
  DSTAG_RETURN_LAUNCH("dstag_sat_softmax_fwd(rows)");
}


// ---------------------------------------------------------------------------------------
// Tensor-core variant for d_k = 32 and ld <= 1024 (every PEMS graph size; the drought graph, N = 2139, does not fit).  The K = 32 logit
// products are far too thin for the tcgen05 engine (one k-block, 883 accumulator columns per row against
// 512 TMEM columns) but they are what bounds the SIMT variant above (one shared-memory float per 4 FMAs =
// the shared-memory roofline).  Here a CTA of 8 warps owns 16 output rows j of one (b, k); warp w owns the
// 8-column tiles {w, w+8, ...} of the complete logit rows as m16n8k8 accumulator fragments (4*NT registers
// per thread) and computes them as 3xTF32 (hi.hi + hi.lo + lo.hi with truncation splits: fp32-level
// accuracy).  No shared-memory staging at all: the fragments' K index is permuted (thread tg takes the
// physical components {4tg..4tg+3, 16+4tg..16+4tg+3}) so both operands are plain float4 loads of whole
// 128-byte projection rows, and the W_Q rows stay L1-resident across the row groups a CTA walks.
// Row max / sum: quad shuffles + an 8-way cross-warp reduction through 1 KB of shared memory; the fragment
// layout gives every quad one full 32-byte sector of P, TP, biasT and tkT per row.
// ---------------------------------------------------------------------------------------
constexpr int kSatMmaThreads = 256;
constexpr int kSatMmaRows = 16;

__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v) & 0xFFFFE000u;                     // the 19 bits the tensor core reads
  lo = __float_as_uint(v - __uint_as_float(hi));             // exact remainder
}

constexpr int kSatRing = 4;
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;                                    // 0 source bytes: the slot is zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(PENDING) : "memory"); }

constexpr int kSatRingBytes = (kSatMmaThreads / 32) * kSatRing * 2 * 32 * 16;   // lane-private W_Q slots
constexpr int kSatRedBytes = 2 * kSatMmaRows * 8 * 4;
__host__ __device__ constexpr int sat_mma_pitch(int nt) { return 64 * nt + 8; }   // = 8 (mod 32): conflict-free LDS.64
__host__ __device__ constexpr size_t sat_mma_smem(int nt) {
  return (size_t)kSatRingBytes + kSatRedBytes + (size_t)kSatMmaRows * sat_mma_pitch(nt) * 4;
}

template <int NT>
__global__ void __launch_bounds__(kSatMmaThreads, 2)
sat_softmax_mma_kernel(const float* __restrict__ qk, const float* __restrict__ biasT, const float* __restrict__ tkT,
                       float* __restrict__ P, float* __restrict__ TP, int N, int K, int ld, float scale) {
  constexpr int DK = 32;
  constexpr int PITCH = sat_mma_pitch(NT);
  extern __shared__ __align__(16) unsigned char sat_smem[];
  auto qring = reinterpret_cast<float4(*)[kSatRing][2][32]>(sat_smem);
  auto red = reinterpret_cast<float(*)[kSatMmaRows][8]>(sat_smem + kSatRingBytes);
  float* rows_s = reinterpret_cast<float*>(sat_smem + kSatRingBytes + kSatRedBytes);   // [16][PITCH]: bias, then T_k
  const int bk = blockIdx.y, b = bk / K, k = bk - b * K;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, tg = lane & 3;
  const long long nstride = 2LL * K * DK;                         // qk: (B, N, 2, K, dk)
  const float* qbase = qk + (long long)b * N * nstride + (long long)k * DK + 4 * tg;
  const float* kbase = qbase + (long long)K * DK;
  const int j0 = blockIdx.x * kSatMmaRows;

  // The 16 bias rows (and later the 16 T_k rows) of the strip travel to shared memory asynchronously: warp w
  // copies rows w and w + 8 in 16-byte pieces.  Fetched through registers these loads sat behind their uses
  // at 16 warps per SM and cost a full L2 round trip each -- 3/4 of the kernel's time in the first version.
  auto copy_rows = [&](const float* src) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = w + 8 * h;
      const float* srow = src + ((long long)k * N + min(j0 + r, N - 1)) * ld;
      float* drow = rows_s + r * PITCH;
      for (int c = 4 * lane; c < ld; c += 128) cp_async16_zfill(drow + c, srow + c, true);
    }
  };
  copy_rows(biasT);
  cp_async_commit();

  // A fragments (W_K rows j0+g, j0+g+8), split once
  uint32_t ah[4][4], al[4][4];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = j0 + g + 8 * h;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
    if (j < N) {
      const float* p = kbase + (long long)j * nstride;
      x = __ldg(reinterpret_cast<const float4*>(p));
      y = __ldg(reinterpret_cast<const float4*>(p + 16));
    }
    const float e[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      split_tf32(e[2 * s], ah[s][h], al[s][h]);
      split_tf32(e[2 * s + 1], ah[s][2 + h], al[s][2 + h]);
    }
  }
  float acc[NT][4];
#pragma unroll
  for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;

  // W_Q fragments: every lane fetches exactly the two 16-byte pieces it consumes itself, kSatRing - 1 column
  // tiles ahead, with cp.async into lane-private shared-memory slots -- an asynchronous register prefetch that
  // costs no registers and needs no barrier (cp.async.ca keeps the rows in L1 for the other warps and CTAs).
  auto issue_q = [&](int t) {
    const int i = 8 * (t * 8 + w) + g;
    const bool ok = i < N;
    const float* p = qbase + (long long)(ok ? i : 0) * nstride;
    float4* slot = &qring[w][t % kSatRing][0][lane];
    cp_async16_zfill(slot, p, ok);
    cp_async16_zfill(slot + 32, p + 16, ok);
  };
#pragma unroll
  for (int t = 0; t < kSatRing - 1; ++t) {
    if (t < NT) issue_q(t);
    cp_async_commit();
  }
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    if (t + kSatRing - 1 < NT) issue_q(t + kSatRing - 1);
    cp_async_commit();
    cp_async_wait<kSatRing - 1>();                                  // tile t (and the older bias group) landed
    const float4 qa = qring[w][t % kSatRing][0][lane], qb = qring[w][t % kSatRing][1][lane];
    const float e[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      uint32_t b0h, b0l, b1h, b1l;
      split_tf32(e[2 * s], b0h, b0l);
      split_tf32(e[2 * s + 1], b1h, b1l);
      mma_tf32_16x8x8(acc[t], al[s], b0h, b1h);                   // small terms first
      mma_tf32_16x8x8(acc[t], ah[s], b0l, b1l);
      mma_tf32_16x8x8(acc[t], ah[s], b0h, b1h);
    }
  }
  cp_async_wait<0>();
  __syncthreads();                                                  // every warp's share of the bias rows is visible

  // logits = acc * scale + bias, masked beyond N; row max
  const int cw = 8 * w + 2 * tg;                                    // this thread's column within every 64-column group
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float* brow = rows_s + (g + 8 * h) * PITCH + cw;
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int c = cw + 64 * t;
      const float2 bb = *reinterpret_cast<const float2*>(brow + 64 * t);
      const float v0 = c < N ? fmaf(acc[t][2 * h], scale, bb.x) : -INFINITY;
      const float v1 = c + 1 < N ? fmaf(acc[t][2 * h + 1], scale, bb.y) : -INFINITY;
      acc[t][2 * h] = v0;
      acc[t][2 * h + 1] = v1;
      mx = fmaxf(mx, fmaxf(v0, v1));
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    if (tg == 0) red[0][g + 8 * h][w] = mx;
  }
  __syncthreads();                                                  // maxima published, bias rows consumed
  if (TP) copy_rows(tkT);                                           // T_k rows land behind the exponentials
  cp_async_commit();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float4 m0 = *reinterpret_cast<const float4*>(&red[0][g + 8 * h][0]);
    const float4 m1 = *reinterpret_cast<const float4*>(&red[0][g + 8 * h][4]);
    const float mx = fmaxf(fmaxf(fmaxf(m0.x, m0.y), fmaxf(m0.z, m0.w)), fmaxf(fmaxf(m1.x, m1.y), fmaxf(m1.z, m1.w)));
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const float e0 = __expf(acc[t][2 * h] - mx), e1 = __expf(acc[t][2 * h + 1] - mx);   // exp(-inf) = 0
      acc[t][2 * h] = e0;
      acc[t][2 * h + 1] = e1;
      sum += e0 + e1;
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    if (tg == 0) red[1][g + 8 * h][w] = sum;
  }
  cp_async_wait<0>();
  __syncthreads();
  if (TP == nullptr) {
    // P only (the sparse Chebyshev path): the strip's 16 rows are assembled in the row buffer (free since the bias
    // rows were consumed) and leave as ONE bulk copy per row (cp.async.bulk shared -> global): the fragment layout
    // gives a store instruction eight 32-byte sectors of eight different rows, which kept the LSU pipe at 72 %
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 s0 = *reinterpret_cast<const float4*>(&red[1][g + 8 * h][0]);
      const float4 s1 = *reinterpret_cast<const float4*>(&red[1][g + 8 * h][4]);
      const float inv = 1.f / (((s0.x + s0.y) + (s0.z + s0.w)) + ((s1.x + s1.y) + (s1.z + s1.w)));
      float* srow = rows_s + (g + 8 * h) * PITCH + cw;
#pragma unroll
      for (int t = 0; t < NT; ++t)
        if (cw + 64 * t < ld) *reinterpret_cast<float2*>(srow + 64 * t) = make_float2(acc[t][2 * h] * inv, acc[t][2 * h + 1] * inv);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the bulk copy
    __syncthreads();
    if (tid < kSatMmaRows && j0 + tid < N) {
      float* dst = P + (((long long)b * K + k) * N + j0 + tid) * ld;
      const uint32_t src = (uint32_t)__cvta_generic_to_shared(rows_s + tid * PITCH);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(ld * 4) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the CTA's shared memory must outlive the read
    }
    return;
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = j0 + g + 8 * h;
    if (j >= N) continue;
    const float4 s0 = *reinterpret_cast<const float4*>(&red[1][g + 8 * h][0]);
    const float4 s1 = *reinterpret_cast<const float4*>(&red[1][g + 8 * h][4]);
    const float inv = 1.f / (((s0.x + s0.y) + (s0.z + s0.w)) + ((s1.x + s1.y) + (s1.z + s1.w)));
    float* prow = P + (((long long)b * K + k) * N + j) * ld + cw;
    float* tprow = TP + (((long long)b * K + k) * N + j) * ld + cw;
    const float* trow = rows_s + (g + 8 * h) * PITCH + cw;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      if (cw + 64 * t < ld) {
        const float2 pv = make_float2(acc[t][2 * h] * inv, acc[t][2 * h + 1] * inv);
        *reinterpret_cast<float2*>(prow + 64 * t) = pv;
        if (TP) {
          const float2 tv = *reinterpret_cast<const float2*>(trow + 64 * t);
          *reinterpret_cast<float2*>(tprow + 64 * t) = make_float2(pv.x * tv.x, pv.y * tv.y);
        }
      }
    }
  }
}

template <int NT>
static int launch_sat_mma(const float* qk, const float* biasT, const float* tkT, float* P, float* TP, int B, int N,
                          int K, int ld, float scale, cudaStream_t stream) {
  static int configured_dev = -1;                                   // per instantiation; the attribute is per device
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t smem = sat_mma_smem(NT);
  if (dev != configured_dev) {
    cudaError_t e = cudaFuncSetAttribute(sat_softmax_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("dstag_sat_softmax_fwd(mma): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured_dev = dev;
  }
  const dim3 grid((N + kSatMmaRows - 1) / kSatMmaRows, B * K);
  sat_softmax_mma_kernel<NT><<<grid, kSatMmaThreads, smem, stream>>>(qk, biasT, tkT, P, TP, N, K, ld, scale);
  DSTAG_RETURN_LAUNCH("dstag_sat_softmax_fwd(mma)");
}

}  // namespace dstag

using namespace dstag;

extern "C" int dstag_sat_softmax_fwd(const float* qk, const float* biasT, const float* tkT, float* P, float* TP,
                                     int32_t B, int32_t N, int32_t K,
                                     int32_t dk, int32_t ld, void* stream) {
  DSTAG_CHECK_ARG(qk && biasT && P, "dstag_sat_softmax_fwd: null pointer");
  DSTAG_CHECK_ARG((TP == nullptr) || (tkT != nullptr && (reinterpret_cast<uintptr_t>(tkT) & 15) == 0 &&
                                      (reinterpret_cast<uintptr_t>(TP) & 15) == 0),
                  "dstag_sat_softmax_fwd: TP needs a 16-byte aligned tkT");
  DSTAG_CHECK_ARG(B > 0 && N > 0 && K > 0 && dk > 0 && ld >= N && (long long)B * K <= 65535,
                  "dstag_sat_softmax_fwd: bad sizes");
  DSTAG_CHECK_ARG(ld % 4 == 0 && (reinterpret_cast<uintptr_t>(biasT) & 15) == 0,
                  "dstag_sat_softmax_fwd: ld must be a multiple of 4 and biasT 16-byte aligned");
  const float scale = 1.0f / sqrtf((float)dk);
  const int nt = (ld + kSatTileI - 1) / kSatTileI;
  const bool rows_ok = nt <= 8 && dk % 4 == 0 && (size_t)dk * (kSatRowsCta + kSatQld) * 4 <= 48 * 1024 &&
                       (reinterpret_cast<uintptr_t>(qk) & 15) == 0 && (reinterpret_cast<uintptr_t>(P) & 15) == 0;
  // DSTAG_SAT_MMA=0 keeps the SIMT variants (A/B measurements)
  static const int mma_on = [] { const char* e = getenv("DSTAG_SAT_MMA"); return e ? atoi(e) : 1; }();
  if (mma_on && rows_ok && dk == 32) {
    cudaStream_t st = (cudaStream_t)stream;
    const int nt64 = (ld + 63) / 64;
#define DSTAG_SAT_MMA_CASE(NT_) \
  case NT_: return launch_sat_mma<NT_>(qk, biasT, tkT, P, TP, B, N, K, ld, scale, st)
    switch (nt64) {
      DSTAG_SAT_MMA_CASE(1); DSTAG_SAT_MMA_CASE(2); DSTAG_SAT_MMA_CASE(3); DSTAG_SAT_MMA_CASE(4);
      DSTAG_SAT_MMA_CASE(5); DSTAG_SAT_MMA_CASE(6); DSTAG_SAT_MMA_CASE(7); DSTAG_SAT_MMA_CASE(8);
      DSTAG_SAT_MMA_CASE(9); DSTAG_SAT_MMA_CASE(10); DSTAG_SAT_MMA_CASE(11); DSTAG_SAT_MMA_CASE(12);
      DSTAG_SAT_MMA_CASE(13); DSTAG_SAT_MMA_CASE(14); DSTAG_SAT_MMA_CASE(15);
      default: DSTAG_SAT_MMA_CASE(16);
    }
#undef DSTAG_SAT_MMA_CASE
  }
  if (rows_ok) {
    cudaStream_t st = (cudaStream_t)stream;
    switch (nt) {
      case 1: return launch_sat_rows<1>(qk, biasT, tkT, P, TP, B, N, K, dk, ld, scale, st);
      case 2: return launch_sat_rows<2>(qk, biasT, tkT, P, TP, B, N, K, dk, ld, scale, st);
      case 3: return launch_sat_rows<3>(qk, biasT, tkT, P, TP, B, N, K, dk, ld, scale, st);
      case 4: return launch_sat_rows<4>(qk, biasT, tkT, P, TP, B, N, K, dk, ld, scale, st);
      case 5: return launch_sat_rows<5>(qk, biasT, tkT, P, TP, B, N, K, dk, ld, scale, st);
      case 6: return launch_sat_rows<6>(qk, biasT, tkT, P, TP, B, N, K, dk, ld, scale, st);
      case 7: return launch_sat_rows<7>(qk, biasT, tkT, P, TP, B, N, K, dk, ld, scale, st);
      default: return launch_sat_rows<8>(qk, biasT, tkT, P, TP, B, N, K, dk, ld, scale, st);
    }
  }
  const size_t fixed = (size_t)(dk * kSatQld) * sizeof(float);
  auto need = [&](int R) { return fixed + (size_t)R * (dk + ld) * sizeof(float); };
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
    sat_softmax_fwd_kernel<R_><<<grid, kSatThreads, smem, (cudaStream_t)stream>>>(qk, biasT, tkT, P, TP, N, K, dk, ld, \
                                                                                  scale);                           \
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
