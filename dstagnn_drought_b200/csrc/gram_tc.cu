// Tall-skinny "Gram" products on tcgen05: the weight-gradient shapes of the block.
//
//   D[(j, c), n] = sum_r A_j[r, c] * B[r, n]       r over 10^5 .. 10^6 rows, j < taps <= 7, c < 32, n < NB (32 or 64)
//
// with A_j[r, :] = the 128-byte row `arow(r) + j * tap_rows` (+ j * tap_plane floats) of a channels-last activation:
//   * GTU conv weight gradient  dW_ks[o, c, j] = sum_{x,t'} du[(x,t'), o] g[x, t'+j, c]     (model/DSTAGNN_my.py:193,
//     autograd): A = g, arow(x,t') = x T + t', tap j = the next time step; B = du (R, 64)
//   * Theta gradient            dTheta_k[f, c] = sum_{x,t} Z_k[(x,t), f] dG[(x,t), c]       (line 131): A = Z, tap k = the
//     next polynomial order (a plane stride); B = dG (R, 32)
// Both operands are row-contiguous, i.e. the reduction index is the SLOW index of both: the contraction engine
// stages such operands MN-major through shared memory and splits the reduction 150-300 ways with a second pass; these
// were its slowest launches (19-60 TFLOP/s).  Here:
//   * rows are the K dimension: a k-block is G = floor(32 / Tp) whole nodes (up to 32 rows, the rest zero).  A (M = (tap, channel), 128 lanes per 4 taps) goes through TENSOR
//     MEMORY: producer thread (tap j, channel c) reads its column of the k-block with 32 coalesced 4-byte loads (a warp
//     = one tap reads one whole 128-byte row per instruction; the taps' windows overlap in L1), splits it into the
//     3xTF32 hi / lo terms and stores them with tcgen05.st -- no shared memory, no transposition;
//   * B tiles (32 rows x NB) are copied by 16-byte cp.async straight into the MN-major SWIZZLE_128B_BASE32B atom layout
//     the tensor core reads (the layout gemm_tc.cu established), lo plane derived in place, 4 stages;
//   * the accumulators (128 lanes x NB columns per 4 taps) stay in TMEM for the CTA's WHOLE share of the rows; one
//     partial result per CTA goes to a workspace and dstag_sum_slices adds the partials in a fixed order
//     (reproducible, no atomics).
// 13 warps: 0-7 A producers (two M tiles, or two teams on alternate k-blocks when there is one), 8-11 B producers,
// 12 MMA issuer (elect.sync).  One CTA per SM.
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace dstag {

namespace {

constexpr int GR_A_WARPS = 8, GR_B_WARPS = 4;
constexpr int GR_MMA_WARP = GR_A_WARPS + GR_B_WARPS;
constexpr int GR_THREADS = (GR_MMA_WARP + 1) * 32;        // 416
constexpr int GR_MIN_SMEM = 120 * 1024;                   // one CTA per SM (TMEM is allocated whole)
constexpr int GR_SLOTS = 6;                               // TMEM A slots of 64 columns (hi 32 | lo 32)
constexpr int GR_BSTAGES = 4;

struct GramParams {
  const float* a;           // rows of 32 floats
  const float* b;           // (R, NB)
  float* partial;           // (gridDim.x, MT * 128, NB)
  int X, T, Tp;             // X nodes of Tp GEMM rows each; GEMM row (x, tp) reads A row x * T + tp
  int G;                    // nodes per k-block: G * Tp <= 32 rows (the other rows of the block are zero rows)
  int taps;
  long long tap_stride;     // floats from tap j to tap j + 1 of the same GEMM row
  int nkb;                  // k-blocks = ceil(X / G)
  long long* trace;         // optional per-warp cycle counters (DSTAG_GRAM_TRACE), normally NULL
};

__device__ __forceinline__ void tmem_st32g(uint32_t taddr, const uint32_t (&r)[32]) {
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

template <int MT, int NB>
struct GramCfg {
  static constexpr int RB = NB / 32;                   // 32-column blocks of the B tile
  static constexpr int B_PLANE = 32 * NB * 4;          // one k-block of B: 32 rows x NB floats
  static constexpr int B_STAGE = 2 * B_PLANE;          // hi | lo
  static constexpr int BAR_OFF = GR_BSTAGES * B_STAGE;
  static constexpr int SMEM_USED = BAR_OFF + 1024 + 1024;
  static constexpr int SMEM = SMEM_USED > GR_MIN_SMEM ? SMEM_USED : GR_MIN_SMEM;
  static constexpr int ACC_COLS = MT * NB;             // accumulators: M tile mt at columns [mt * NB, (mt + 1) * NB)
  static constexpr int A_TM = 128;                     // first column of the A slots
};

// TPC: Tp known at compile time (the GTU kernel sizes at T = 12: 10, 8, 6; the Theta gradient: 1) or 0 = run time.
// A k-block is G whole nodes, so with a constant Tp every load of the block is `LDG [node base + immediate]`: ONE
// instruction.  The producers are instruction-issue bound (ncu: 55 % issue active, 5 500 warp instructions per
// k-block); walking the rows with a running (node, time) pair cost 13 instructions per load.
template <int MT, int NB, int TPC>
__global__ void __launch_bounds__(GR_THREADS, 1) gram_tc_kernel(const __grid_constant__ GramParams p) {
  using Cfg = GramCfg<MT, NB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* afull_bar = bars;                          // [SLOTS]   A producers -> MMA
  uint64_t* aempty_bar = bars + GR_SLOTS;              // [SLOTS]   MMA (commit) -> A producers
  uint64_t* bfull_bar = bars + 2 * GR_SLOTS;           // [BSTAGES] B producers -> MMA
  uint64_t* bempty_bar = bfull_bar + GR_BSTAGES;       // [BSTAGES] MMA (commit) -> B producers
  uint64_t* done_bar = bempty_bar + GR_BSTAGES;        // MMA (commit) -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_begin = p.trace ? clock64() : 0;
  long long w_wait = 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < GR_SLOTS; ++s) {
      mbar_init(afull_bar + s, 4);
      mbar_init(aempty_bar + s, 1);
    }
    for (int s = 0; s < GR_BSTAGES; ++s) {
      mbar_init(bfull_bar + s, GR_B_WARPS);
      mbar_init(bempty_bar + s, 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == GR_MMA_WARP) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // this CTA's contiguous share of the k-blocks
  const int per = (p.nkb + (int)gridDim.x - 1) / (int)gridDim.x;
  const int kb0 = (int)blockIdx.x * per;
  const int kb1 = min(p.nkb, kb0 + per);
  const int nk = max(0, kb1 - kb0);

  if (warp < GR_A_WARPS) {
    // =============================== A producers =========================================
    // MT == 2: warps 0-3 build M tile 0 (taps 0-3), warps 4-7 M tile 1 (taps 4-7), every k-block.
    // MT == 1: warps 0-3 and 4-7 are two teams that take alternate k-blocks.
    const int q = warp & 3, half = warp >> 2;
    const int mt = MT == 2 ? half : 0;
    const int tap = 4 * mt + q;
    const bool live = tap < p.taps;
    const float* abase = p.a + (long long)tap * p.tap_stride + lane;
    // the 32 loads of the NEXT k-block are in flight while the current one is split and stored (one exposed global
    // latency per k-block otherwise: the whole CTA works on the same k-block)
    const int Tp = TPC > 0 ? TPC : p.Tp;
    const int G = TPC > 0 ? 32 / TPC : p.G;
    const long long node_stride = (long long)p.T * 32;
    auto load_block = [&](int i, uint32_t (&v)[32]) {
      if (live) {
        const int x0 = (kb0 + i) * G;
        const int nlast = min(G, p.X - x0) - 1;            // nodes beyond it re-read the last one: their B rows are zero
        const float* bp = abase + (long long)x0 * node_stride;
        if (TPC == 1 && nlast == 31) {
          // (this instantiation is dispatched for T == Tp == 1 only: the 32 rows of the block are consecutive)
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(__ldg(bp + k * 32));
        } else if (TPC > 0) {
          constexpr int GC = TPC > 0 ? 32 / TPC : 1, TC = TPC > 0 ? TPC : 1;
          const bool whole = nlast == GC - 1;               // (a partial last block clamps the node index)
#pragma unroll
          for (int n = 0; n < GC; ++n) {
            const float* np = bp + (long long)(whole ? n : min(n, nlast)) * node_stride;
#pragma unroll
            for (int t = 0; t < TC; ++t) v[n * TC + t] = __float_as_uint(__ldg(np + t * 32));
          }
#pragma unroll
          for (int k = GC * TC; k < 32; ++k) v[k] = 0u;
        } else {
          int n = 0, t = 0;
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const bool in = n < G;
            v[k] = in ? __float_as_uint(__ldg(bp + (long long)min(n, nlast) * node_stride + t * 32)) : 0u;
            ++t;
            const bool wrap = t == Tp;
            t = wrap ? 0 : t;
            n += wrap ? 1 : 0;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0u;
      }
    };
    constexpr int STEP = MT == 2 ? 1 : 2;
    const int i0 = MT == 2 ? 0 : half;
    uint32_t v[32], vn[32];
    if (i0 < nk) load_block(i0, v);
    for (int i = i0; i < nk; i += STEP) {
      const int slot = MT == 2 ? (i % (GR_SLOTS / 2)) * 2 + mt : i % GR_SLOTS;
      const uint32_t ph = MT == 2 ? (i / (GR_SLOTS / 2)) & 1 : (i / GR_SLOTS) & 1;
      if (i + STEP < nk) load_block(i + STEP, vn);
      { const long long t0 = p.trace ? clock64() : 0;
      if (lane == 0) mbar_wait(aempty_bar + slot, ph ^ 1, 0);
      __syncwarp();
      if (p.trace) w_wait += clock64() - t0; }
      tc_fence_after();
      const uint32_t ta = tmem_base + Cfg::A_TM + slot * 64 + ((uint32_t)(32 * q) << 16);
      tmem_st32g(ta, v);                                 // hi: the raw value (the tensor core ignores the low 13 bits)
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(tf32_lo(__uint_as_float(v[k])));
      tmem_st32g(ta + 32, v);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(afull_bar + slot);
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = vn[k];
    }
    // =============================== epilogue (same warps) ================================
    if (MT == 2 || half == 0) {
      if (nk > 0) {
        if (lane == 0) mbar_wait(done_bar, 0, 64);
        __syncwarp();
        tc_fence_after();
      }
      float* out = p.partial + ((long long)blockIdx.x * (MT * 128) + mt * 128 + 32 * q + lane) * NB;
#pragma unroll
      for (int cb = 0; cb < NB / 32; ++cb) {
        float acc[32];
        if (nk > 0) {
          tmem_ld32(tmem_base + mt * NB + 32 * cb + ((uint32_t)(32 * q) << 16), acc);
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) acc[k] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < 32; k += 4)
          *reinterpret_cast<float4*>(out + 32 * cb + k) = make_float4(acc[k], acc[k + 1], acc[k + 2], acc[k + 3]);
      }
    }
  } else if (warp < GR_A_WARPS + GR_B_WARPS) {
    // =============================== B producers =========================================
    // 32 rows x NB floats per k-block = 8 NB 16-byte chunks, chunk ch of row k -> MN-major atom layout:
    //   atom (ka = k / 4, nb = ch / 8) at (ka * RB + nb) * 512, k-row kr = k % 4 at + 128 kr,
    //   chunk mc = ch % 8 of the 32-column block at + (((mc >> 1) ^ kr) << 5) + ((mc & 1) << 4)
    const int bt = threadIdx.x - GR_A_WARPS * 32;          // 0 .. 127
    constexpr int CHUNKS = 8 * NB;                         // per k-block
    constexpr int PER = CHUNKS / (GR_B_WARPS * 32);        // 4 (NB = 64) or 2 (NB = 32)
    constexpr int CPR = NB / 4;                            // chunks per row
    uint32_t doff[PER];
    int krow[PER], kch[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int idx = bt + u * (GR_B_WARPS * 32);
      const int k = idx / CPR, ch = idx - k * CPR;
      const int ka = k >> 2, kr = k & 3, nb = ch >> 3, mc = ch & 7;
      doff[u] = (uint32_t)(((ka * Cfg::RB + nb) << 9) + (kr << 7) + (((mc >> 1) ^ kr) << 5) + ((mc & 1) << 4));
      krow[u] = k;
      kch[u] = ch;
    }
    auto issue = [&](int i) {
      const int s = i % GR_BSTAGES;
      { const long long t0 = p.trace ? clock64() : 0;
      if (lane == 0) mbar_wait(bempty_bar + s, ((i / GR_BSTAGES) & 1) ^ 1, 0);
      __syncwarp();
      if (p.trace) w_wait += clock64() - t0; }
      const uint32_t dst = smem_u32(smem + s * Cfg::B_STAGE);
      const int x0 = (kb0 + i) * p.G;
      const int nvalid = min(p.G, p.X - x0) * p.Tp;          // the block's rows are contiguous in B; the rest are zero rows
      const long long r0 = (long long)x0 * p.Tp;
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const bool ok = krow[u] < nvalid;
        cp_async16(dst + doff[u], ok ? (const void*)(p.b + (r0 + krow[u]) * NB + 4 * kch[u]) : (const void*)p.b, ok ? 16u : 0u);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (nk > 0) issue(0);
    if (nk > 1) issue(1);
    for (int i = 0; i < nk; ++i) {
      if (i + 2 < nk) {
        issue(i + 2);
        asm volatile("cp.async.wait_group 2;" ::: "memory");
      } else if (i + 1 < nk) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      // this thread's own chunks of k-block i have landed: derive the lo plane
      uint8_t* st = smem + (i % GR_BSTAGES) * Cfg::B_STAGE;
#pragma unroll
      for (int u = 0; u < PER; ++u)
        *reinterpret_cast<float4*>(st + Cfg::B_PLANE + doff[u]) = tf32_lo4(*reinterpret_cast<const float4*>(st + doff[u]));
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bfull_bar + (i % GR_BSTAGES));
    }
  } else {
    // =============================== MMA issuer ========================================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) /* B MN-major */ |
                           ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint64_t B_ADV = (uint64_t)(Cfg::RB * 1024) >> 4;      // one k-step of 8 = two k-atoms
    long long w_bw = 0;
    for (int i = 0; i < nk; ++i) {
      const int sb = i % GR_BSTAGES;
      { const long long t0 = p.trace ? clock64() : 0;
      mbar_wait(bfull_bar + sb, (i / GR_BSTAGES) & 1, 0);
      if (p.trace) w_bw += clock64() - t0; }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int slot = MT == 2 ? (i % (GR_SLOTS / 2)) * 2 + mt : i % GR_SLOTS;
        const uint32_t ph = MT == 2 ? (i / (GR_SLOTS / 2)) & 1 : (i / GR_SLOTS) & 1;
        { const long long t0 = p.trace ? clock64() : 0;
        mbar_wait(afull_bar + slot, ph, 0);
        if (p.trace) w_wait += clock64() - t0; }
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sbase = smem_u32(smem + sb * Cfg::B_STAGE);
          const uint64_t bh = make_desc_mn(sbase, Cfg::RB * 512);
          const uint64_t bl = make_desc_mn(sbase + Cfg::B_PLANE, Cfg::RB * 512);
          const uint32_t d_tmem = tmem_base + mt * NB;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t ta = tmem_base + Cfg::A_TM + slot * 64 + 8 * kk;      // hi columns; lo at +32
            umma_tf32_ts(d_tmem, ta + 32, bh + B_ADV * kk, idesc, (i == 0 && kk == 0) ? 0u : 1u);
            umma_tf32_ts(d_tmem, ta, bl + B_ADV * kk, idesc, 1u);
            umma_tf32_ts(d_tmem, ta, bh + B_ADV * kk, idesc, 1u);
          }
          umma_commit(aempty_bar + slot);
          if (mt == MT - 1) {
            umma_commit(bempty_bar + sb);
            if (i == nk - 1) umma_commit(done_bar);
          }
        }
        __syncwarp();
      }
    }
    if (p.trace && lane == 0) p.trace[((long long)blockIdx.x * 16 + 13) * 4] = w_bw;
  }

  if (p.trace && lane == 0) {
    long long* tr = p.trace + ((long long)blockIdx.x * 16 + warp) * 4;
    tr[1] = clock64() - t_begin; tr[2] = w_wait; tr[3] = nk;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == GR_MMA_WARP) tmem_dealloc(tmem_base, 512);
}

constexpr int kMaxDevG = 64;

int gram_sms(int dev) {
  static int sms[kMaxDevG] = {0};
  if (sms[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms[dev] = n > 0 ? n : kNumSMs;
  }
  return sms[dev];
}

template <int MT, int NB, int TPC>
int launch_gram(const GramParams& p, cudaStream_t stream) {
  using Cfg = GramCfg<MT, NB>;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevG) dev = 0;
  static bool configured[kMaxDevG] = {false};
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gram_tc_kernel<MT, NB, TPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) {
      set_error("dstag_gram_rows: cudaFuncSetAttribute(%d) failed: %s", Cfg::SMEM, cudaGetErrorString(e));
      return (int)e;
    }
    configured[dev] = true;
  }
  gram_tc_kernel<MT, NB, TPC><<<gram_sms(dev), GR_THREADS, Cfg::SMEM, stream>>>(p);
  DSTAG_RETURN_LAUNCH("dstag_gram_rows");
}

}  // namespace

}  // namespace dstag

extern "C" {

int dstag_gram_slices(void) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= dstag::kMaxDevG) dev = 0;
  return dstag::gram_sms(dev);
}

int dstag_gram_supported(int32_t taps, int32_t NB, int32_t T, int32_t Tp) {
  return (taps >= 1 && taps <= 8 && (NB == 32 || NB == 64) && Tp >= 1 && Tp <= T && Tp <= 32) ? 1 : 0;
}

int dstag_gram_rows(const float* a, const float* b, float* partial, int64_t X, int32_t T, int32_t Tp, int32_t taps,
                    int64_t tap_stride, int32_t NB, void* stream) {
  DSTAG_CHECK_ARG(a && b && partial && X > 0, "dstag_gram_rows: null pointer or empty problem");
  DSTAG_CHECK_ARG(dstag_gram_supported(taps, NB, T, Tp), "dstag_gram_rows: unsupported shape taps=%d NB=%d T=%d Tp=%d", taps,
                  NB, T, Tp);
  DSTAG_CHECK_ARG(X * T * 32 < (1ll << 31) - 64, "dstag_gram_rows: too many rows");
  DSTAG_CHECK_ARG(tap_stride != 32 || Tp + taps - 1 <= T,
                  "dstag_gram_rows: taps that are consecutive rows need Tp + taps - 1 <= T (got Tp=%d taps=%d T=%d)", Tp, taps, T);
  DSTAG_CHECK_ARG(((reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(partial)) & 15) == 0,
                  "dstag_gram_rows: b and partial must be 16-byte aligned");
  dstag::GramParams p;
  p.a = a; p.b = b; p.partial = partial;
  p.X = (int)X; p.T = T; p.Tp = Tp; p.G = 32 / Tp; p.taps = taps; p.tap_stride = tap_stride;
  p.nkb = (int)((X + p.G - 1) / p.G);
  { const char* e = getenv("DSTAG_GRAM_TRACE"); p.trace = e ? (long long*)strtoull(e, nullptr, 0) : nullptr; }
  cudaStream_t st = (cudaStream_t)stream;
  const bool two = taps > 4;
  // instantiations with a compile-time Tp for the shapes of the shipped configurations (T = 12, ks = 3 / 5 / 7; the
  // Theta gradient), the run-time walk for everything else
  if (NB == 64 && !two && Tp == 10) return dstag::launch_gram<1, 64, 10>(p, st);
  if (NB == 64 && two && Tp == 8) return dstag::launch_gram<2, 64, 8>(p, st);
  if (NB == 64 && two && Tp == 6) return dstag::launch_gram<2, 64, 6>(p, st);
  if (NB == 32 && !two && Tp == 1 && T == 1) return dstag::launch_gram<1, 32, 1>(p, st);
  if (!two) return NB == 64 ? dstag::launch_gram<1, 64, 0>(p, st) : dstag::launch_gram<1, 32, 0>(p, st);
  return NB == 64 ? dstag::launch_gram<2, 64, 0>(p, st) : dstag::launch_gram<2, 32, 0>(p, st);
}

}  // extern "C"
