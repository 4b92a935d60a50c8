// Gated temporal convolutions (GTU, model/DSTAGNN_my.py:184-197, 237-242) as two fused tcgen05 kernels.
//
// Every tensor is channels-last: g (X, T, C) with C = 32 channels innermost, so the window of ks time
// steps that a Conv2d(C -> 2C, (1, ks)) output reads is ks consecutive 128-byte rows of g.
//
//   gtu_fwd_tc_kernel<KS>   cat[x, off + t', :] = tanh(p) * sigmoid(q),  [p | q] = bias + sum_j g[x, t'+j, :] W_j^T
//                           One GEMM row per (x, t'): M = 128 rows per tile, N = 64, K = KS k-blocks of 32
//                           channels.  The A operand never exists in memory: a producer thread owns one
//                           row, reads the 128-byte line of time step t'+j straight from global memory (the
//                           KS lines of a window are re-used through L1), splits it into the 3xTF32 hi / lo
//                           terms in registers and stores them into tensor memory (tcgen05.st); the MMAs
//                           take A from TMEM (TS form) and the weights (hi and lo, all KS k-blocks) stay
//                           resident in shared memory for the life of the CTA.  The epilogue applies bias
//                           and gate to the TMEM accumulator and writes cat and -- for the backward --
//                           [tanh(p) | sigmoid(q)]: the conv output u never reaches HBM.
//   gtu_bwd_tc_kernel<KS>   du = gate'(saved) * dcat (written once, unpadded, for the weight-gradient
//                           contraction), then Y[(x,t'), (j,c)] = sum_o du[(x,t'), o] W[o, c, j]
//                           (M = 128, N = 32 KS, K = 64: two k-blocks, A again through TMEM) and the
//                           overlap-add dg[x, t'+j, c] += Y[(x,t'), (j,c)] in the epilogue through a
//                           shared-memory tile of whole nodes -- no zero-padded copy of du, no
//                           KS-fold re-read, no atomics.
//
// Roles: warps 0-7 two producer teams (warp % 4 = TMEM lane quarter; the teams take alternate k-blocks / tiles),
// then the epilogue warps (8 in the forward kernel: two per lane quarter with 16 channels each; 4 in the
// backward kernel, whose producers need 128 registers), then the MMA issuer (whole warp in the loop, one lane
// chosen by elect.sync issues).  Persistent: one CTA per SM, TMEM fully allocated.  All global traffic is
// coalesced: the rows of g (forward) and of the gate factors / dcat (backward) are staged with 16-byte cp.async
// into chunk-swizzled shared-memory tiles one tile ahead, and the outputs leave through swizzled tiles as well --
// a thread reading or writing "its" 128-byte row directly touched 32 lines per instruction and ran at a third
// of the speed.  profiles/r02_gtu_fused.md has the per-role cycle traces behind these choices.
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace dstag {

namespace {

constexpr int GT_C = 32;
constexpr int GT_TEAMS = 2;
constexpr int GT_PROD_WARPS = 4 * GT_TEAMS;
constexpr int GT_EPI_WARPS = 8;                            // two per TMEM lane quarter (channel halves)
constexpr int GT_MMA_WARP = GT_PROD_WARPS + GT_EPI_WARPS;
constexpr int GT_THREADS = (GT_MMA_WARP + 1) * 32;        // 544
// backward: 4 epilogue warps (13 warps = 416 threads: 128 registers for the producers' three row blocks)
constexpr int GB_EPI_WARPS = 4;
constexpr int GB_MMA_WARP = GT_PROD_WARPS + GB_EPI_WARPS;
constexpr int GB_THREADS = (GB_MMA_WARP + 1) * 32;        // 416
constexpr int GT_MIN_SMEM = 120 * 1024;                   // > half of an SM's shared memory: one CTA per SM (TMEM)

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// one row of a k-block: 32 floats -> hi (columns 0-31) | lo (columns 32-63) of a TMEM A slot, two x32 stores
// (the raw fp32 value IS the hi operand: the tensor core ignores the low 13 mantissa bits)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void split_store_row(uint32_t taddr, const float4* v) {
  uint32_t r[32];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    r[4 * c] = __float_as_uint(v[c].x); r[4 * c + 1] = __float_as_uint(v[c].y);
    r[4 * c + 2] = __float_as_uint(v[c].z); r[4 * c + 3] = __float_as_uint(v[c].w);
  }
  tmem_st32(taddr, r);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    r[4 * c] = __float_as_uint(tf32_lo(v[c].x)); r[4 * c + 1] = __float_as_uint(tf32_lo(v[c].y));
    r[4 * c + 2] = __float_as_uint(tf32_lo(v[c].z)); r[4 * c + 3] = __float_as_uint(tf32_lo(v[c].w));
  }
  tmem_st32(taddr + 32, r);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct GtuFwdParams {
  const float* g;      // (X, T, 32)
  const float* w;      // (64, 32, 1, KS) conv weight, PyTorch layout
  const float* bias;   // (64)
  float* cat;          // (X, S, 32): rows off .. off + Tp - 1 of every node are written
  float* sv;           // (X, Tp, 64) [tanh(p) | sigmoid(q)] or nullptr
  int R;               // X * Tp GEMM rows
  long long GR;        // X * T rows of g
  int T, Tp, S, off, ntiles;
  long long* trace;    // optional per-warp cycle counters (tools/gtu_microbench.py --trace), normally NULL
};

constexpr int GT_STAGE_ROWS = 288;       // rows of g staged per tile: 127 + KS + (KS - 1) * ceil(128 / Tp) must fit

template <int KS>
struct FwdCfg {
  static constexpr int SLOTS = 6;                  // TMEM A slots of 64 columns (hi 32 | lo 32)
  static constexpr int ACC = 64;                   // accumulator columns (double buffered)
  static constexpr int A_TM = 2 * ACC;
  static constexpr int B_TILE = 64 * 128;          // one k-block of the weights: 64 rows (o) x 32 floats (c)
  static constexpr int B_PLANE = KS * B_TILE;
  static constexpr int G_OFF = 2 * B_PLANE;        // two stages of GT_STAGE_ROWS rows of g (128 B each, chunk-swizzled)
  static constexpr int G_STAGE = GT_STAGE_ROWS * 128;
  static constexpr int E_OFF = G_OFF + 2 * G_STAGE;   // per epilogue warp: 32 rows x 128 B transposing buffer
  static constexpr int BAR_OFF = E_OFF + GT_EPI_WARPS * 2048;
  static constexpr int SMEM_USED = BAR_OFF + 1024 + 1024 /*align slack*/;
  static constexpr int SMEM = SMEM_USED > GT_MIN_SMEM ? SMEM_USED : GT_MIN_SMEM;
};

// Rows of g a tile of 128 GEMM rows (x, t') needs: from the first row's window start to the last row's window end.
__device__ __forceinline__ void tile_span(const GtuFwdParams& p, int tile, int ks, long long& first, int& count) {
  const int r0 = tile * 128;
  const int r1 = min(p.R, r0 + 128) - 1;
  const int x0 = r0 / p.Tp, x1 = r1 / p.Tp;
  first = (long long)x0 * p.T + (r0 - x0 * p.Tp);
  count = (int)((long long)x1 * p.T + (r1 - x1 * p.Tp) + ks - first);
}

template <int KS>
__global__ void __launch_bounds__(GT_THREADS, 1) gtu_fwd_tc_kernel(const __grid_constant__ GtuFwdParams p) {
  using Cfg = FwdCfg<KS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* b_hi = smem;
  uint8_t* b_lo = smem + Cfg::B_PLANE;
  uint8_t* gst = smem + Cfg::G_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* full_bar = bars;                          // [SLOTS] producers -> MMA
  uint64_t* empty_bar = bars + Cfg::SLOTS;            // [SLOTS] MMA (commit) -> producers
  uint64_t* afull_bar = bars + 2 * Cfg::SLOTS;        // [2] MMA (commit) -> epilogue
  uint64_t* aempty_bar = afull_bar + 2;               // [2] epilogue -> MMA
  uint64_t* gfull_bar = aempty_bar + 2;               // [2] cp.async of a g stage landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gfull_bar + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_begin = p.trace ? clock64() : 0;
  long long w_wait = 0;

  // ---- one-time setup: weights (hi / lo, K-major SWIZZLE_128B tiles, one per tap j), barriers, TMEM ----
  for (int idx = threadIdx.x; idx < KS * 64 * 32; idx += GT_THREADS) {
    const int j = idx / 2048, o = (idx >> 5) & 63, c = idx & 31;
    const float v = __ldg(p.w + (o * GT_C + c) * KS + j);
    const uint32_t off = (uint32_t)(j * Cfg::B_TILE) + swz(o, c >> 2) + ((c & 3) << 2);
    *reinterpret_cast<float*>(b_hi + off) = v;
    *reinterpret_cast<float*>(b_lo + off) = tf32_lo(v);
  }
  if (threadIdx.x < 64) bias_s[threadIdx.x] = __ldg(p.bias + threadIdx.x);
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::SLOTS; ++s) {
      mbar_init(full_bar + s, 4);
      mbar_init(empty_bar + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(afull_bar + a, 1);
      mbar_init(aempty_bar + a, GT_EPI_WARPS);
      mbar_init(gfull_bar + a, GT_PROD_WARPS * 32);
    }
    fence_barrier_init();
  }
  fence_proxy_async_smem();                 // generic-proxy writes of the weights -> visible to the tensor core
  if (warp == GT_MMA_WARP) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long t_setup = p.trace ? clock64() : 0;

  const int my_tiles = p.ntiles > (int)blockIdx.x ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp < GT_PROD_WARPS) {
    // =============================== producers =========================================
    // All 256 producer threads stage the rows of g a tile needs with coalesced 16-byte cp.async copies (8 lanes
    // per 128-byte row, chunks XOR-swizzled with the row so that a thread reading ITS row later is conflict
    // free), one tile ahead.  Then the two teams take alternate k-blocks: thread = GEMM row (x, t'), k-block j =
    // the 32 channels of time step t' + j -- the same staged row serves up to KS windows.
    const int team = warp >> 2, q = warp & 3;
    const int rl = 32 * q + lane;
    const int pt = threadIdx.x;                       // 0 .. 255
    auto stage_tile = [&](int ts) {
      const int tile = (int)blockIdx.x + ts * (int)gridDim.x;
      long long first;
      int count;
      tile_span(p, tile, KS, first, count);
      const uint32_t dst0 = smem_u32(gst + (ts & 1) * Cfg::G_STAGE);
      const float* src0 = p.g + first * GT_C;
      for (int idx = pt; idx < count * 8; idx += GT_PROD_WARPS * 32) {
        const int m = idx >> 3, c = idx & 7;
        const bool ok = first + m < p.GR;
        cp_async16(dst0 + m * 128 + ((c ^ (m & 7)) << 4), ok ? (const void*)(src0 + m * GT_C + c * 4) : (const void*)p.g,
                   ok ? 16u : 0u);
      }
      cp_async_arrive(gfull_bar + (ts & 1));
    };
    if (my_tiles > 0) stage_tile(0);
    uint32_t it = 0;
    for (int ts = 0; ts < my_tiles; ++ts) {
      named_bar_sync(2, GT_PROD_WARPS * 32);           // everyone has finished reading tile ts - 1: its stage is free
      if (ts + 1 < my_tiles) stage_tile(ts + 1);
      const int tile = (int)blockIdx.x + ts * (int)gridDim.x;
      const int r = tile * 128 + rl;
      const bool valid = r < p.R;
      long long first;
      int count;
      tile_span(p, tile, KS, first, count);
      int m0 = 0;
      if (valid) {
        const int x = r / p.Tp, tp = r - x * p.Tp;
        m0 = (int)((long long)x * p.T + tp - first);
      }
      const uint8_t* gs = gst + (ts & 1) * Cfg::G_STAGE;
      {
        const long long t0 = p.trace ? clock64() : 0;
        if (lane == 0) mbar_wait(gfull_bar + (ts & 1), (ts >> 1) & 1, 0);
        __syncwarp();
        if (p.trace) w_wait += clock64() - t0;
      }
#pragma unroll 1
      for (int j = 0; j < KS; ++j, ++it) {
        if ((int)(it & 1) != team) continue;
        const int s = (int)(it % Cfg::SLOTS);
        const uint32_t ph = (it / Cfg::SLOTS) & 1;
        float4 v[8];
        if (valid) {
          const int m = m0 + j;
          const uint8_t* row = gs + m * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4*>(row + ((c ^ (m & 7)) << 4));
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        {
          const long long t0 = p.trace ? clock64() : 0;
          if (lane == 0) mbar_wait(empty_bar + s, ph ^ 1, 0);
          __syncwarp();
          if (p.trace) w_wait += clock64() - t0;
        }
        tc_fence_after();
        split_store_row(tmem_base + Cfg::A_TM + s * 64 + ((uint32_t)(32 * q) << 16), v);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar + s);
      }
    }
  } else if (warp == GT_MMA_WARP) {
    // =============================== MMA issuer ========================================
    // (the whole warp walks the loop -- uniform control flow -- and one elected lane issues)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sb_hi = smem_u32(b_hi), sb_lo = smem_u32(b_lo);
    uint32_t it = 0;
    for (int ts = 0; ts < my_tiles; ++ts) {
      const uint32_t acc = ts & 1;
      mbar_wait(aempty_bar + acc, ((ts >> 1) & 1) ^ 1, 0);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * Cfg::ACC;
#pragma unroll 1
      for (int j = 0; j < KS; ++j, ++it) {
        const int s = (int)(it % Cfg::SLOTS);
        {
          const long long t0 = p.trace ? clock64() : 0;
          mbar_wait(full_bar + s, (it / Cfg::SLOTS) & 1, 0);
          if (p.trace) w_wait += clock64() - t0;
        }
        tc_fence_after();
        if (elect_one()) {
          const uint64_t bh = make_desc_k(sb_hi + j * Cfg::B_TILE);
          const uint64_t bl = make_desc_k(sb_lo + j * Cfg::B_TILE);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t ta = tmem_base + Cfg::A_TM + s * 64 + 8 * kk;      // hi columns; lo at +32
            umma_tf32_ts(d_tmem, ta + 32, bh + 2ull * kk, idesc, (j == 0 && kk == 0) ? 0u : 1u);
            umma_tf32_ts(d_tmem, ta, bl + 2ull * kk, idesc, 1u);
            umma_tf32_ts(d_tmem, ta, bh + 2ull * kk, idesc, 1u);
          }
          umma_commit(empty_bar + s);
          if (j == KS - 1) umma_commit(afull_bar + acc);
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== epilogue ==========================================
    // Two warps per TMEM lane quarter, each with 16 of the 32 channels; thread = GEMM row: bias + gate on the TMEM
    // accumulator.  The three 32 x 16 blocks of a warp (cat, tanh(p), sigmoid(q)) go through a swizzled
    // shared-memory tile so that every global store instruction writes whole 64-byte row halves of 8 rows (a
    // thread storing its own row touched 32 lines per instruction).
    const int e = warp - GT_PROD_WARPS, q = e & 3, half = e >> 2;
    long long e_ld = 0, e_math = 0, e_st = 0;
    uint8_t* ebuf = smem + Cfg::E_OFF + e * 2048;
    const int rsub = lane >> 2, cc = lane & 3;
    for (int ts = 0; ts < my_tiles; ++ts) {
      const int tile = (int)blockIdx.x + ts * (int)gridDim.x;
      const uint32_t acc = ts & 1;
      {
        const long long t0 = p.trace ? clock64() : 0;
        if (lane == 0) mbar_wait(afull_bar + acc, (ts >> 1) & 1, 64);
        __syncwarp();
        if (p.trace) w_wait += clock64() - t0;
      }
      tc_fence_after();
      const long long e0 = p.trace ? clock64() : 0;
      float pv[16], qv[16];
      const uint32_t ta = tmem_base + acc * Cfg::ACC + 16 * half + ((uint32_t)(32 * q) << 16);
      tmem_ld16(ta, pv);
      tmem_ld16(ta + 32, qv);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(aempty_bar + acc);
      const long long e1 = p.trace ? clock64() : 0;
      const int r = tile * 128 + 32 * q + lane;
      long long cat_off = -1;                               // element offset of this thread's row of cat (-1: no row)
      if (r < p.R) {
        const int x = r / p.Tp, tp = r - x * p.Tp;
        cat_off = ((long long)x * p.S + p.off + tp) * GT_C;
      }
      long long co[4];                                      // rows this lane stores: 8 i + lane / 4
#pragma unroll
      for (int i = 0; i < 4; ++i) co[i] = __shfl_sync(0xffffffffu, cat_off, 8 * i + rsub);
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        pv[c] = tanh_fast(pv[c] + bias_s[16 * half + c]);
        qv[c] = sigmoid_fast(qv[c] + bias_s[32 + 16 * half + c]);
      }
      const long long e2 = p.trace ? clock64() : 0;
#pragma unroll 1
      for (int blk = 0; blk < (p.sv ? 3 : 1); ++blk) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float4 o;
          if (blk == 0) o = make_float4(pv[4 * c] * qv[4 * c], pv[4 * c + 1] * qv[4 * c + 1], pv[4 * c + 2] * qv[4 * c + 2], pv[4 * c + 3] * qv[4 * c + 3]);
          else if (blk == 1) o = make_float4(pv[4 * c], pv[4 * c + 1], pv[4 * c + 2], pv[4 * c + 3]);
          else o = make_float4(qv[4 * c], qv[4 * c + 1], qv[4 * c + 2], qv[4 * c + 3]);
          *reinterpret_cast<float4*>(ebuf + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) = o;
        }
        __syncwarp();
        float4 o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = 8 * i + rsub;
          o[i] = *reinterpret_cast<const float4*>(ebuf + rr * 64 + ((cc ^ ((rr >> 1) & 3)) << 4));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (co[i] >= 0) {
            const int rr = 8 * i + rsub;
            float* dst = blk == 0 ? p.cat + co[i] + 16 * half + 4 * cc
                                  : p.sv + (long long)(tile * 128 + 32 * q + rr) * 64 + (blk - 1) * 32 + 16 * half + 4 * cc;
            *reinterpret_cast<float4*>(dst) = o[i];
          }
        }
        __syncwarp();
      }
      if (p.trace) { const long long e3 = clock64(); e_ld += e1 - e0; e_math += e2 - e1; e_st += e3 - e2; }
    }
    if (p.trace && lane == 0 && e == 0) {
      long long* tr = p.trace + ((long long)blockIdx.x * 20 + 18) * 4;
      tr[0] = e_ld; tr[1] = e_math; tr[2] = e_st;
    }
  }

  if (p.trace && lane == 0) {
    long long* tr = p.trace + ((long long)blockIdx.x * 20 + warp) * 4;
    tr[0] = t_setup - t_begin; tr[1] = clock64() - t_setup; tr[2] = w_wait; tr[3] = my_tiles;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == GT_MMA_WARP) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------
// backward: gate' -> du (global), Y = du W (tensor cores), overlap-add -> dg
// ---------------------------------------------------------------------------------------------------------
struct GtuBwdParams {
  const float* sv;     // (X, Tp, 64)
  const float* dcat;   // (X, S, 32)
  const float* w;      // (64, 32, 1, KS)
  float* du;           // (X, Tp, 64)
  float* dg;           // (X, T, 32): stored, or added to when `accumulate`
  float* db_ws;        // (gridDim.x * 8, 64): per producer warp partial column sums of du (the bias gradient)
  int X, T, Tp, S, off, npt, ntiles, accumulate;
  long long* trace;
};

template <int KS>
struct BwdCfg {
  static constexpr int N = 32 * KS;                // accumulator columns: (tap j, channel c)
  // the accumulator is used in two halves (taps [0, TAPS0) and [TAPS0, KS)): the epilogue drains one while the
  // tensor core fills the other -- 2 x N columns for a double buffer do not fit next to the A slots
  static constexpr int TAPS0 = (KS + 1) / 2, TAPS1 = KS / 2;
  static constexpr int A_TM = 256;                 // two A slot pairs of 128 columns: [p hi | p lo | q hi | q lo]
  static constexpr int B_TILE = N * 128;           // one k-block (32 of the 64 conv outputs o) x N rows
  static constexpr int B_PLANE = 2 * B_TILE;
  // per producer warp: NB tiles of 32 rows x 128 B (tanh, sigmoid and -- when they fit next to the weights -- dcat)
  static constexpr int NB = KS <= 5 ? 3 : 2;
  static constexpr int P_OFF = 2 * B_PLANE;
  static constexpr int BAR_OFF = P_OFF + GT_PROD_WARPS * NB * 4096;
  static constexpr int DG_OFF = BAR_OFF + 1024;
  static constexpr int DG_PITCH = 33;
};

// The warp's 32 rows of a (rows x 128 B) block, copied with coalesced 16-byte cp.async (8 lanes per row, 4 rows
// per instruction) into the warp's swizzled shared-memory tile -- no registers are held while the copy is in
// flight -- where every thread later reads ITS row (conflict free).  rowptr(r) -> address of row r, or nullptr
// (the row is zero-filled).
template <class RowPtr>
__device__ __forceinline__ void issue_rows(uint8_t* buf, int lane, const float* dummy, RowPtr rowptr) {
  const int rsub = lane >> 3, cc = lane & 7;
  const uint32_t b0 = smem_u32(buf);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = 4 * i + rsub;
    const float* rp = rowptr(rr);
    cp_async16(b0 + rr * 128 + ((cc ^ (rr & 7)) << 4), rp ? (const void*)(rp + 4 * cc) : (const void*)dummy, rp ? 16u : 0u);
  }
}
__device__ __forceinline__ void read_own_row(const uint8_t* buf, int lane, float4 (&out)[8]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) out[c] = *reinterpret_cast<const float4*>(buf + lane * 128 + ((c ^ (lane & 7)) << 4));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// the reverse: one row per thread -> coalesced stores
// Returns the sums over the warp's 32 rows of the four columns 4 (lane % 8) .. + 3 (the same value in the four
// lanes that share lane % 8): the conv bias gradient is the column sum of du.
template <class RowPtr>
__device__ __forceinline__ float4 store_rows_transposed(uint8_t* buf, int lane, RowPtr rowptr, const float4 (&in)[8]) {
  const int rsub = lane >> 3, cc = lane & 7;
#pragma unroll
  for (int c = 0; c < 8; ++c) *reinterpret_cast<float4*>(buf + lane * 128 + ((c ^ (lane & 7)) << 4)) = in[c];
  __syncwarp();
  float4 t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = 4 * i + rsub;
    t[i] = *reinterpret_cast<const float4*>(buf + rr * 128 + ((cc ^ (rr & 7)) << 4));
  }
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float* rp = rowptr(4 * i + rsub);
    if (rp) *(reinterpret_cast<float4*>(rp) + cc) = t[i];
    cs.x += t[i].x; cs.y += t[i].y; cs.z += t[i].z; cs.w += t[i].w;      // (rows beyond the tile hold zeros)
  }
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o);
    cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
    cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o);
    cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
  }
  __syncwarp();
  return cs;
}

template <int KS>
__global__ void __launch_bounds__(GB_THREADS, 1) gtu_bwd_tc_kernel(const __grid_constant__ GtuBwdParams p) {
  using Cfg = BwdCfg<KS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* b_hi = smem;
  uint8_t* b_lo = smem + Cfg::B_PLANE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* full_bar = bars;            // [2] producers -> MMA (one per A slot pair)
  uint64_t* empty_bar = bars + 2;       // [2] MMA (commit) -> producers
  uint64_t* afull_bar = bars + 4;       // [2] MMA (commit) -> epilogue, per accumulator half
  uint64_t* aempty_bar = bars + 6;      // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* dgs = reinterpret_cast<float*>(smem + Cfg::DG_OFF);       // (npt * T) x 33: dg of the tile's nodes

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_begin = p.trace ? clock64() : 0;
  long long w_wait = 0;

  // weights as B[n = (j, c)][k = o - 32 h], K-major SWIZZLE_128B, one tile per half h of the conv outputs
  for (int idx = threadIdx.x; idx < 2 * Cfg::N * 32; idx += GB_THREADS) {
    const int h = idx / (Cfg::N * 32), rem = idx - h * (Cfg::N * 32);
    const int n = rem >> 5, k = rem & 31;
    const int j = n >> 5, c = n & 31, o = 32 * h + k;
    const float v = __ldg(p.w + (o * GT_C + c) * KS + j);
    const uint32_t off = (uint32_t)(h * Cfg::B_TILE) + swz(n, k >> 2) + ((k & 3) << 2);
    *reinterpret_cast<float*>(b_hi + off) = v;
    *reinterpret_cast<float*>(b_lo + off) = tf32_lo(v);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(full_bar + s, 4);
      mbar_init(empty_bar + s, 1);
      mbar_init(afull_bar + s, 1);
      mbar_init(aempty_bar + s, GB_EPI_WARPS);
    }
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (warp == GB_MMA_WARP) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long t_setup = p.trace ? clock64() : 0;

  const int my_tiles = p.ntiles > (int)blockIdx.x ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp < GT_PROD_WARPS) {
    // =============================== producers =========================================
    // thread = GEMM row (x, t') of the tile's whole nodes; the two teams take alternate tiles
    const int team = warp >> 2, q = warp & 3;
    const int rl = 32 * q + lane;
    uint8_t* pbuf = smem + Cfg::P_OFF + warp * (Cfg::NB * 4096);
    float4 bsum_p = make_float4(0.f, 0.f, 0.f, 0.f), bsum_q = bsum_p;
    // copies of tile ts: tanh -> tile 0, sigmoid -> tile 1, dcat -> tile 2 (if there is one)
    auto issue_tile = [&](int ts, bool with_d) {
      const int tile = (int)blockIdx.x + ts * (int)gridDim.x;
      const int x0 = tile * p.npt;
      const int nrows = min(p.npt, p.X - x0) * p.Tp;
      const long long r0 = (long long)x0 * p.Tp;
      if (!with_d || Cfg::NB == 3) {
        issue_rows(pbuf, lane, p.sv, [&](int r) -> const float* {
          const int row = 32 * q + r;
          return row < nrows ? p.sv + (r0 + row) * 64 : nullptr; });
        issue_rows(pbuf + 4096, lane, p.sv, [&](int r) -> const float* {
          const int row = 32 * q + r;
          return row < nrows ? p.sv + (r0 + row) * 64 + 32 : nullptr; });
      }
      if (with_d) {
        issue_rows(pbuf + (Cfg::NB == 3 ? 8192 : 0), lane, p.sv, [&](int r) -> const float* {
          const int row = 32 * q + r;
          if (row >= nrows) return nullptr;
          const int xl = row / p.Tp, tp = row - xl * p.Tp;
          return p.dcat + ((long long)(x0 + xl) * p.S + p.off + tp) * GT_C; });
      }
    };
    if (team < my_tiles) issue_tile(team, Cfg::NB == 3);
    for (int ts = team; ts < my_tiles; ts += GT_TEAMS) {
      const int tile = (int)blockIdx.x + ts * (int)gridDim.x;
      const int sp = ts & 1;
      const uint32_t ph = (ts >> 1) & 1;
      const int x0 = tile * p.npt;
      const int nn = min(p.npt, p.X - x0);
      const int nrows = nn * p.Tp;
      const long long r0 = (long long)x0 * p.Tp;          // first row of the tile in sv / du (whole nodes: contiguous)
      float4 th[8], sg[8];
      cp_async_wait_all();
      __syncwarp();
      read_own_row(pbuf, lane, th);
      read_own_row(pbuf + 4096, lane, sg);
      if (Cfg::NB == 2) {
        __syncwarp();
        issue_tile(ts, true);                             // dcat into tile 0 while the gate derivative is computed
      }
      // gate': th <- sigmoid(q) (1 - tanh(p)^2), sg <- tanh(p) sigmoid(q) (1 - sigmoid(q)); then du = . * dcat
      // (rows beyond the tile hold zeros)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float t[4] = {th[c].x, th[c].y, th[c].z, th[c].w}, s[4] = {sg[c].x, sg[c].y, sg[c].z, sg[c].w};
        float rp[4], rq[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          rp[e] = s[e] * (1.f - t[e] * t[e]);
          rq[e] = t[e] * s[e] * (1.f - s[e]);
        }
        th[c] = make_float4(rp[0], rp[1], rp[2], rp[3]);
        sg[c] = make_float4(rq[0], rq[1], rq[2], rq[3]);
      }
      {
        float4 d[8];
        if (Cfg::NB == 2) {
          cp_async_wait_all();
          __syncwarp();
        }
        read_own_row(pbuf + (Cfg::NB == 3 ? 8192 : 0), lane, d);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          th[c] = make_float4(th[c].x * d[c].x, th[c].y * d[c].y, th[c].z * d[c].z, th[c].w * d[c].w);
          sg[c] = make_float4(sg[c].x * d[c].x, sg[c].y * d[c].y, sg[c].z * d[c].z, sg[c].w * d[c].w);
        }
      }
      __syncwarp();                                       // every lane has read its rows: the tiles can be reused
      {
        const float4 cp = store_rows_transposed(pbuf, lane, [&](int r) -> float* {
          const int row = 32 * q + r;
          return row < nrows ? p.du + (r0 + row) * 64 : nullptr; }, th);
        const float4 cq = store_rows_transposed(pbuf + 4096, lane, [&](int r) -> float* {
          const int row = 32 * q + r;
          return row < nrows ? p.du + (r0 + row) * 64 + 32 : nullptr; }, sg);
        bsum_p.x += cp.x; bsum_p.y += cp.y; bsum_p.z += cp.z; bsum_p.w += cp.w;
        bsum_q.x += cq.x; bsum_q.y += cq.y; bsum_q.z += cq.z; bsum_q.w += cq.w;
      }
      if (ts + GT_TEAMS < my_tiles) issue_tile(ts + GT_TEAMS, Cfg::NB == 3);      // the team's next tile, in flight from here
      {
        const long long t0 = p.trace ? clock64() : 0;
        if (lane == 0) mbar_wait(empty_bar + sp, ph ^ 1, 0);
        __syncwarp();
        if (p.trace) w_wait += clock64() - t0;
      }
      tc_fence_after();
      const uint32_t ta = tmem_base + Cfg::A_TM + sp * 128 + ((uint32_t)(32 * q) << 16);
      split_store_row(ta, th);
      split_store_row(ta + 64, sg);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar + sp);
      (void)rl;
    }
    if (p.db_ws && lane < 8) {
      float* wsr = p.db_ws + ((long long)blockIdx.x * GT_PROD_WARPS + warp) * 64;
      *reinterpret_cast<float4*>(wsr + 4 * lane) = bsum_p;
      *reinterpret_cast<float4*>(wsr + 32 + 4 * lane) = bsum_q;
    }
  } else if (warp == GB_MMA_WARP) {
    // =============================== MMA issuer ========================================
    const uint32_t sb_hi = smem_u32(b_hi), sb_lo = smem_u32(b_lo);
    for (int ts = 0; ts < my_tiles; ++ts) {
      const int sp = ts & 1;
      {
        const long long t0 = p.trace ? clock64() : 0;
        mbar_wait(full_bar + sp, (ts >> 1) & 1, 0);
        if (p.trace) w_wait += clock64() - t0;
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int nh = half == 0 ? 32 * Cfg::TAPS0 : 32 * Cfg::TAPS1;      // columns of this accumulator half
        const int row0 = half == 0 ? 0 : 32 * Cfg::TAPS0;                  // first B row = first accumulator column
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nh >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        {
          const long long t0 = p.trace ? clock64() : 0;
          mbar_wait(aempty_bar + half, (ts & 1) ^ 1, 0);
          if (p.trace) w_wait += clock64() - t0;
        }
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t bh = make_desc_k(sb_hi + h * Cfg::B_TILE + row0 * 128);
            const uint64_t bl = make_desc_k(sb_lo + h * Cfg::B_TILE + row0 * 128);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t ta = tmem_base + Cfg::A_TM + sp * 128 + h * 64 + 8 * kk;
              umma_tf32_ts(tmem_base + row0, ta + 32, bh + 2ull * kk, idesc, (h == 0 && kk == 0) ? 0u : 1u);
              umma_tf32_ts(tmem_base + row0, ta, bl + 2ull * kk, idesc, 1u);
              umma_tf32_ts(tmem_base + row0, ta, bh + 2ull * kk, idesc, 1u);
            }
          }
          if (half == 1) umma_commit(empty_bar + sp);
          umma_commit(afull_bar + half);
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== epilogue: overlap-add ================================
    // Two warps per TMEM lane quarter (16 of a tap's 32 channels each); thread = GEMM row (x, t'): tap j of its
    // accumulator row is the contribution to dg[x, t' + j, :].  The taps are added one at a time into a
    // shared-memory tile of the tile's whole nodes (within a tap every row of the tile has ONE writer), which
    // then goes out with coalesced stores.
    const int q = warp & 3;
    const int et = threadIdx.x - GT_PROD_WARPS * 32;      // 0 .. 127
    constexpr int NE = GB_EPI_WARPS * 32;
    const int rl = 32 * q + lane;
    for (int ts = 0; ts < my_tiles; ++ts) {
      const int tile = (int)blockIdx.x + ts * (int)gridDim.x;
      const int x0 = tile * p.npt;
      const int nn = min(p.npt, p.X - x0);
      const int nrows = nn * p.T;
      // dg of the previous kernel sizes (accumulate mode): fetched now, used after the taps
      constexpr int PF = 16;
      float* gout = p.dg + (long long)x0 * p.T * GT_C;
      const int total = nrows * 8;
      float4 old[PF];
#pragma unroll
      for (int k = 0; k < PF; ++k) {
        const int i = et + k * NE;
        old[k] = (p.accumulate && i < total) ? __ldcs(reinterpret_cast<const float4*>(gout + (i >> 3) * GT_C + ((i & 7) << 2)))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      for (int i = et; i < nrows * Cfg::DG_PITCH; i += NE) dgs[i] = 0.f;
      named_bar_sync(1, NE);
      const bool valid = rl < nn * p.Tp;
      const int xl = valid ? rl / p.Tp : 0, tp = valid ? rl - xl * p.Tp : 0;
      float* drow = dgs + (xl * p.T + tp) * Cfg::DG_PITCH;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        {
          const long long t0 = p.trace ? clock64() : 0;
          if (lane == 0) mbar_wait(afull_bar + half, ts & 1, 32);
          __syncwarp();
          if (p.trace) w_wait += clock64() - t0;
        }
        tc_fence_after();
        const int j0 = half == 0 ? 0 : Cfg::TAPS0, j1 = half == 0 ? Cfg::TAPS0 : KS;
#pragma unroll 1
        for (int j = j0; j < j1; ++j) {
          float v[32];
          tmem_ld32(tmem_base + 32 * j + ((uint32_t)(32 * q) << 16), v);
          if (valid) {
            float* dd = drow + j * Cfg::DG_PITCH;
#pragma unroll
            for (int c = 0; c < 32; ++c) dd[c] += v[c];
          }
          named_bar_sync(1, NE);       // rows of tap j and tap j+1 of one node belong to different threads
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(aempty_bar + half);
      }
      // write-out: coalesced float4 stores; the first PF chunks of every thread were prefetched (accumulate mode)
#pragma unroll
      for (int k = 0; k < PF; ++k) {
        const int i = et + k * NE;
        if (i < total) {
          const float* s = dgs + (i >> 3) * Cfg::DG_PITCH + ((i & 7) << 2);
          *reinterpret_cast<float4*>(gout + (i >> 3) * GT_C + ((i & 7) << 2)) =
              make_float4(old[k].x + s[0], old[k].y + s[1], old[k].z + s[2], old[k].w + s[3]);
        }
      }
      for (int i = et + PF * NE; i < total; i += NE) {
        const float* s = dgs + (i >> 3) * Cfg::DG_PITCH + ((i & 7) << 2);
        float4* dst = reinterpret_cast<float4*>(gout + (i >> 3) * GT_C + ((i & 7) << 2));
        float4 o = make_float4(s[0], s[1], s[2], s[3]);
        if (p.accumulate) {
          const float4 prev = *dst;
          o.x += prev.x; o.y += prev.y; o.z += prev.z; o.w += prev.w;
        }
        *dst = o;
      }
      named_bar_sync(1, NE);
    }
  }

  if (p.trace && lane == 0) {
    long long* tr = p.trace + ((long long)blockIdx.x * 20 + warp) * 4;
    tr[0] = t_setup - t_begin; tr[1] = clock64() - t_setup; tr[2] = w_wait; tr[3] = my_tiles;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == GB_MMA_WARP) tmem_dealloc(tmem_base, 512);
}

constexpr int kMaxDev = 64;

int device_sms(int dev) {
  static int sms[kMaxDev] = {0};
  if (sms[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms[dev] = n > 0 ? n : kNumSMs;
  }
  return sms[dev];
}

template <int KS>
int launch_fwd(const GtuFwdParams& p, cudaStream_t stream) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDev) dev = 0;
  static bool configured[kMaxDev] = {false};
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gtu_fwd_tc_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdCfg<KS>::SMEM);
    if (e != cudaSuccess) {
      set_error("dstag_gtu_conv_gate_fwd: cudaFuncSetAttribute(%d) failed: %s", FwdCfg<KS>::SMEM, cudaGetErrorString(e));
      return (int)e;
    }
    configured[dev] = true;
  }
  const int grid = p.ntiles < device_sms(dev) ? p.ntiles : device_sms(dev);
  gtu_fwd_tc_kernel<KS><<<grid, GT_THREADS, FwdCfg<KS>::SMEM, stream>>>(p);
  DSTAG_RETURN_LAUNCH("dstag_gtu_conv_gate_fwd");
}

int bwd_nodes_per_tile(int T, int Tp) {
  int npt = 128 / Tp;
  const int cap = (48 * 1024) / (T * BwdCfg<3>::DG_PITCH * 4);      // shared-memory tile of whole nodes (what the weights and the producer tiles leave)
  if (npt > cap) npt = cap;
  return npt;
}

template <int KS>
int launch_bwd(const GtuBwdParams& p, cudaStream_t stream) {
  using Cfg = BwdCfg<KS>;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDev) dev = 0;
  int smem = Cfg::DG_OFF + p.npt * p.T * Cfg::DG_PITCH * 4 + 1024;   // weights + producer tiles + barriers + dg tile
  if (smem < GT_MIN_SMEM) smem = GT_MIN_SMEM;
  static int configured[kMaxDev] = {0};
  if (configured[dev] < smem) {
    cudaError_t e = cudaFuncSetAttribute(gtu_bwd_tc_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("dstag_gtu_gate_conv_bwd: cudaFuncSetAttribute(%d) failed: %s", smem, cudaGetErrorString(e));
      return (int)e;
    }
    configured[dev] = smem;
  }
  const int grid = p.ntiles < device_sms(dev) ? p.ntiles : device_sms(dev);
  gtu_bwd_tc_kernel<KS><<<grid, GB_THREADS, smem, stream>>>(p);
  DSTAG_RETURN_LAUNCH("dstag_gtu_gate_conv_bwd");
}

}  // namespace

}  // namespace dstag

extern "C" {

int dstag_gtu_tc_supported(int32_t C, int32_t T, int32_t ks, int32_t backward) {
  if (C != dstag::GT_C || (ks != 3 && ks != 5 && ks != 7) || T < ks) return 0;
  {
    // the rows of g one tile of 128 (x, t') rows reads must fit the staging buffer
    const int Tp = T - ks + 1;
    if (127 + ks + (ks - 1) * ((128 + Tp - 1) / Tp) > dstag::GT_STAGE_ROWS) return 0;
  }
  if (backward) {
    const int Tp = T - ks + 1;
    if (Tp > 128 || dstag::bwd_nodes_per_tile(T, Tp) < 1) return 0;
  }
  return 1;
}

int dstag_gtu_conv_gate_fwd(const float* g, const float* w, const float* bias, float* cat, float* sv, int64_t X,
                            int32_t C, int32_t T, int32_t ks, int32_t S, int32_t off, void* stream) {
  DSTAG_CHECK_ARG(g && w && bias && cat && X > 0, "dstag_gtu_conv_gate_fwd: null pointer or empty problem");
  DSTAG_CHECK_ARG(dstag_gtu_tc_supported(C, T, ks, 0), "dstag_gtu_conv_gate_fwd: unsupported shape C=%d T=%d ks=%d", C, T, ks);
  const int Tp = T - ks + 1;
  DSTAG_CHECK_ARG(off >= 0 && off + Tp <= S, "dstag_gtu_conv_gate_fwd: rows %d..%d outside the %d rows of cat", off, off + Tp, S);
  DSTAG_CHECK_ARG(X * Tp < (1ll << 31) - 256, "dstag_gtu_conv_gate_fwd: too many rows");
  dstag::GtuFwdParams p;
  p.g = g; p.w = w; p.bias = bias; p.cat = cat; p.sv = sv;
  p.R = (int)(X * Tp); p.T = T; p.Tp = Tp; p.S = S; p.off = off;
  p.GR = (long long)X * T;
  p.ntiles = (p.R + 127) / 128;
  { const char* e = getenv("DSTAG_GTU_TRACE"); p.trace = e ? (long long*)strtoull(e, nullptr, 0) : nullptr; }
  cudaStream_t st = (cudaStream_t)stream;
  if (ks == 3) return dstag::launch_fwd<3>(p, st);
  if (ks == 5) return dstag::launch_fwd<5>(p, st);
  return dstag::launch_fwd<7>(p, st);
}

int dstag_gtu_bwd_ws_rows(void) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= dstag::kMaxDev) dev = 0;
  return dstag::device_sms(dev) * dstag::GT_PROD_WARPS;
}

int dstag_gtu_gate_conv_bwd(const float* sv, const float* dcat, const float* w, float* du, float* dg, float* db_ws,
                            int64_t X, int32_t C, int32_t T, int32_t ks, int32_t S, int32_t off, int32_t accumulate,
                            void* stream) {
  DSTAG_CHECK_ARG(sv && dcat && w && du && dg && X > 0, "dstag_gtu_gate_conv_bwd: null pointer or empty problem");
  DSTAG_CHECK_ARG(dstag_gtu_tc_supported(C, T, ks, 1), "dstag_gtu_gate_conv_bwd: unsupported shape C=%d T=%d ks=%d", C, T, ks);
  const int Tp = T - ks + 1;
  DSTAG_CHECK_ARG(off >= 0 && off + Tp <= S, "dstag_gtu_gate_conv_bwd: rows %d..%d outside the %d rows of dcat", off, off + Tp, S);
  DSTAG_CHECK_ARG(X < (1ll << 31) / T, "dstag_gtu_gate_conv_bwd: too many rows");
  dstag::GtuBwdParams p;
  p.sv = sv; p.dcat = dcat; p.w = w; p.du = du; p.dg = dg; p.db_ws = db_ws;
  p.X = (int)X; p.T = T; p.Tp = Tp; p.S = S; p.off = off;
  p.npt = dstag::bwd_nodes_per_tile(T, Tp);
  p.ntiles = (int)((X + p.npt - 1) / p.npt);
  p.accumulate = accumulate;
  { const char* e = getenv("DSTAG_GTU_TRACE"); p.trace = e ? (long long*)strtoull(e, nullptr, 0) : nullptr; }
  cudaStream_t st = (cudaStream_t)stream;
  if (ks == 3) return dstag::launch_bwd<3>(p, st);
  if (ks == 5) return dstag::launch_bwd<5>(p, st);
  return dstag::launch_bwd<7>(p, st);
}

}  // extern "C"
