// tcgen05 tensor-core contraction engine for sm_100a.
//
//   C[b,m,n] (+)= alpha * sum_k (A . Amul)[b,m,k] * B[b,k,n]   (+Cadd) (*Cmul) (relu)
//
// Persistent, warp-specialised kernel (one CTA per SM, 18 warps = 576 threads):
//   warps 0-7   A producers   cp.async (16-byte, zero-filled at ragged edges) from any 2-level strided view
//                             straight into the stage's swizzled tile, STAGES-1 k-blocks ahead; then
//                             EITHER A through tensor memory (K-major A, 3xTF32): each thread splits its
//                             row of the raw tile in registers and tcgen05.st's hi / lo into TMEM,
//                             OR in place in shared memory: [x Hadamard factor] -> lo = v - trunc_tf32(v)
//   warps 8-11  B producers   the same copies for the second operand; split in place in shared memory
//                             (with a 128-row B tile all 12 producer warps share that split)
//   warps 12-15 epilogue      TMEM -> registers (tcgen05.ld) -> alpha / Cadd / Cmul / relu ->
//                             global (store, add or atomic), accumulator double-buffered in TMEM
//   warp  16    MMA issuer    one thread issues tcgen05.mma.kind::tf32 (128 x BN x 8), A from TMEM (TS form)
//                             or from a shared-memory descriptor; tcgen05.commit releases the stages
//   warp  17    TMA issuer    cp.async.bulk.tensor.4d for K-major, 16-byte aligned, single-level operands
//
// The operands never need a packed / transposed copy in HBM: the producers read the caller's
// strides (permuted views, unfolded conv windows, padded leading dimensions) and the Hadamard
// factor (T_k of cheb_conv_withSAt, model/DSTAGNN_my.py:128) is fused into the operand load.
// Row-contiguous operands are staged MN-major (tcgen05 kind::tf32 takes either major), so no
// transpose is ever materialised either.
//
// Precision: kind::tf32 keeps 10 mantissa bits of each operand.  DSTAG_PREC_TF32X3 restores
// fp32-grade accuracy with the classic 3-pass split  a = a_hi + a_lo,  b = b_hi + b_lo,
// a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi  (error ~2^-21 relative), all accumulated in the
// same fp32 TMEM accumulator.  DSTAG_PREC_TF32 issues the single a_hi*b_hi pass.
#include <cuda.h>   // CUtensorMap types only; cuTensorMapEncodeTiled is fetched at run time (no libcuda link)
#include <cstdlib>
#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace dstag {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                  // 32 x 4 B = one 128-byte swizzle row
constexpr int TC_A_WARPS = 8, TC_B_WARPS = 4, TC_EPI_WARPS = 4;   // epilogue warps 12-15: warp % 4 = TMEM lane quarter
constexpr int TC_MMA_WARP = TC_A_WARPS + TC_B_WARPS + TC_EPI_WARPS;
constexpr int TC_TMA_WARP = TC_MMA_WARP + 1;         // one thread issues the TMA bulk-tensor loads
constexpr int TC_THREADS = (TC_TMA_WARP + 1) * 32;   // 576
#ifndef TC_DBG_SKIP
#define TC_DBG_SKIP 0                                 // timing experiments only (results are wrong): 1 = no A split,
#endif                                                // 2 = no B split, 4 = one MMA per k-step instead of three
#ifndef TC_STAGE_LIMIT
#define TC_STAGE_LIMIT 6                              // (-DTC_STAGE_LIMIT=n: pipeline-depth experiments)
#endif
constexpr int TC_SMEM_BUDGET = 196 * 1024;            // operand stages
constexpr int TC_EPI_SMEM = TC_EPI_WARPS * 32 * 36 * 4;   // per-warp 32 x 32 transpose buffers (row stride 33 or 36)

enum { MODE_KVEC = 0, MODE_MVEC = 1, MODE_ROWG = 2, MODE_KSC = 3 };

struct TcParams {
  dstag_gemm_t g;
  int ntm, ntn, kblocks, kb_per_split, nsplit, total_items;
  int a_mode, b_mode;      // MODE_* (see Producer)
  int a_tma, b_tma;        // operand tile arrives by TMA (cp.async.bulk.tensor) instead of LDG
  int a_bdim[2], b_bdim[2];  // which of the two batch levels are real tensor-map dimensions
  int chunk_kb;              // > 0: accumulate at most this many k-blocks per TMEM pass (see "short chains")
  int a_rt;                  // output rows per M tile: 128, or fewer when the A tile is a ragged TMA box
  int a_rag_d1;              // > 0: A rows are (x, t) windows loaded by a rank-3 box (32 k, d1 t, a_rt / d1 x)
  int epi_coalesce;          // row-contiguous outputs leave through the 32 x 32 transposing path (whole lines per store)
};

struct alignas(64) TcMaps {
  CUtensorMap a, b;
};

__device__ __forceinline__ int axoff32(const dstag_axis_t& a, int idx) {
  if (a.d1 == 1) return idx * (int)a.s0;
  const int hi = idx / a.d1;
  return hi * (int)a.s0 + (idx - hi * a.d1) * (int)a.s1;
}
__device__ __forceinline__ long long axoff64(const dstag_axis_t& a, int idx) {
  if (a.d1 == 1) return (long long)idx * a.s0;
  const int hi = idx / a.d1;
  return (long long)hi * a.s0 + (long long)(idx - hi * a.d1) * a.s1;
}

// Offsets of idx0, idx0 + step, idx0 + 2 step, ... along a (possibly two-level) axis with one
// division at init() and none per step.
struct Walk {
  int off, lo, d1, st_off, st_lo, wrap;
  __device__ __forceinline__ void init(const dstag_axis_t& a, int idx0, int step) {
    d1 = a.d1;
    if (d1 == 1) {
      off = idx0 * (int)a.s0;
      st_off = step * (int)a.s0;
      lo = st_lo = wrap = 0;
    } else {
      const int hi = idx0 / d1;
      lo = idx0 - hi * d1;
      off = hi * (int)a.s0 + lo * (int)a.s1;
      const int sh = step / d1;
      st_lo = step - sh * d1;
      st_off = sh * (int)a.s0 + st_lo * (int)a.s1;
      wrap = (int)a.s0 - d1 * (int)a.s1;
    }
  }
  __device__ __forceinline__ void next() {
    off += st_off;
    if (d1 != 1) {
      lo += st_lo;
      const bool c = lo >= d1;
      lo -= c ? d1 : 0;
      off += c ? wrap : 0;
    }
  }
};

__device__ __forceinline__ float4 mul4(float4 a, float4 b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}

// ---------------------------------------------------------------------------------------
// Operand producer.  NT threads copy one ROWS x 32 operand tile per k-block from global memory
// straight into its final (swizzled) place in shared memory with cp.async, STAGES-1 k-blocks ahead;
// no registers are held across the copy.  Load modes, chosen per launch:
//   KVEC  K contiguous, 16B aligned   : 16-byte copies along K   -> K-major tile,  thread = (chunk, rows)
//   MVEC  rows contiguous, aligned    : 16-byte copies along M/N -> MN-major tile, thread = (row quad, k's)
//   KSC   K contiguous, unaligned     : 4-byte copies, lanes along K (coalesced)  -> K-major tile
//   ROWG  rows contiguous / anything  : 4-byte copies, lanes along rows           -> K-major tile
// Row offsets are computed once per output tile (begin_item) and kept in registers, so a k-block
// costs a handful of instructions per copy.
// ---------------------------------------------------------------------------------------
template <int ROWS, int NT, bool MUL>
struct Producer {
  static constexpr int NV = ROWS * 8 / NT;     // 16-byte chunks per thread per k-block
  static constexpr int NM = MUL ? NV : 1;
  static constexpr int RB = NT / 8;            // KVEC: rows per pass
  static constexpr int QPK = ROWS / 4;         // MVEC: row quads per k
  static constexpr int KPTM = TC_BK / (NT / QPK);   // MVEC: k's per thread (== NV)
  static constexpr int KPTR = TC_BK / (NT / ROWS);  // ROWG: scalars per thread (== 4 NV)
  static constexpr int NW = NT / 32;           // KSC : warps; rows per thread = ROWS / NW (== 4 NV)
  static_assert(KPTM == NV && KPTR == 4 * NV && ROWS / RB == NV && ROWS / NW == 4 * NV, "producer tiling mismatch");
  static_assert(RB % 8 == 0, "KVEC swizzle phase must not depend on the pass");

  const float* base;
  const float* mbase;
  const dstag_axis_t* rowax;
  const dstag_axis_t* colax;
  const dstag_axis_t* mrowax;
  const dstag_axis_t* mcolax;
  int K, t, mode, row0, limit;
  int r_off, r_lo, m_off, m_lo;   // row walk start (this thread's first row) for A and Amul
  int ro[NV], mo[NM];             // KVEC: element offset of this thread's row in pass i (-1: out of range)

  __device__ __forceinline__ int first_row() const {
    if (mode == MODE_KVEC) return row0 + (t >> 3);
    if (mode == MODE_MVEC) return row0 + 4 * ((((t >> 5) % (ROWS / 32)) << 3) + (t & 7));
    if (mode == MODE_KSC) return row0 + (t >> 5);
    return row0 + (t % ROWS);
  }
  __device__ __forceinline__ int row_step() const {
    return mode == MODE_KVEC ? RB : (mode == MODE_KSC ? NW : 0);
  }

  __device__ __forceinline__ void begin_item(const float* b, const float* mb, int r0, int lim) {
    base = b;
    mbase = MUL ? mb : nullptr;
    row0 = r0;
    limit = lim;
    const int fr = first_row();
    Walk w;
    w.init(*rowax, fr < lim ? fr : 0, mode == MODE_KVEC ? RB : 0);
    r_off = w.off; r_lo = w.lo;
    if (mode == MODE_KVEC) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        ro[i] = (fr + RB * i < lim) ? w.off : -1;
        w.next();
      }
    }
    if (MUL && mb) {
      w.init(*mrowax, fr < lim ? fr : 0, mode == MODE_KVEC ? RB : 0);
      m_off = w.off; m_lo = w.lo;
      if (mode == MODE_KVEC) {
#pragma unroll
        for (int i = 0; i < NM; ++i) {
          mo[i] = (fr + RB * i < lim) ? w.off : -1;
          w.next();
        }
      }
    }
  }
  __device__ __forceinline__ void row_walk(Walk& w, const dstag_axis_t& ax, int off, int lo) const {
    w.init(ax, 0, row_step());
    w.off = off;
    w.lo = lo;
  }

  // Issue the asynchronous copies of k-block kb of one operand stream (the tile itself, or its
  // Hadamard factor) into `plane`.
  template <int N>
  __device__ __forceinline__ void issue_stream(int kb, uint8_t* plane, const float* bp, const dstag_axis_t& rax,
                                               const dstag_axis_t& cax, int roff, int rlo, const int (&rows)[N]) const {
    const int k0 = kb * TC_BK;
    const uint32_t pbase = smem_u32(plane);
    if (mode == MODE_KVEC) {
      const int c = t & 7;
      const int k = k0 + 4 * c;
      int kbytes = 4 * (K - k);
      kbytes = kbytes < 0 ? 0 : (kbytes > 16 ? 16 : kbytes);
      const uint32_t d0 = pbase + swz(t >> 3, c);
      const float* kp = bp + (cax.d1 == 1 ? k : (k < K ? axoff32(cax, k) : 0));   // two-level K: runs of d1 % 4 == 0
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const bool ok = rows[i] >= 0 && kbytes > 0;
        cp_async16(d0 + i * (RB * 128), ok ? (const void*)(kp + rows[i]) : (const void*)bp, ok ? (uint32_t)kbytes : 0u);
      }
      return;
    }
    const int fr = first_row();
    if (mode == MODE_MVEC) {
      // MN-major tile: 512-byte atoms of 32 rows x 4 k, atom (ka, mb) at (ka * ROWS/32 + mb) * 512; inside an
      // atom 128-byte k-rows with a 32-byte swizzle on the row quad.  One warp instruction fills one atom
      // (lane = (k in atom, row quad)): 512 contiguous bytes of shared memory -- no bank conflicts -- from
      // four 128-byte lines of global memory.  A thread keeps its row quad and steps through the k atoms.
      constexpr int RBLK = ROWS / 32, KSTEP = NW / RBLK;      // atoms: a = ka * RBLK + mb = warp + NW * j
      static_assert(NW % RBLK == 0 && NV * KSTEP == TC_BK / 4, "MVEC atom tiling");
      const int lane = t & 31, wq = t >> 5;
      const int mc = lane & 7, kr = lane >> 3;
      const int mb = wq % RBLK, ka0 = wq / RBLK;
      int nv = limit - fr;
      nv = nv < 0 ? 0 : (nv > 4 ? 4 : nv);
      const uint32_t d0 = pbase + (uint32_t)(((ka0 * RBLK + mb) << 9) + (kr << 7) + (((mc >> 1) ^ kr) << 5) + ((mc & 1) << 4));
      const int kf = k0 + 4 * ka0 + kr;                       // this thread's k in pass j: kf + 4 * KSTEP * j
      const float* rp = bp + roff;
      if (cax.d1 == 1) {
        const int ks = (int)cax.s0;
        const float* kp = rp + kf * ks;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const uint32_t bytes = (kf + 4 * KSTEP * j < K) ? (uint32_t)(4 * nv) : 0u;
          cp_async16(d0 + j * (NW << 9), bytes ? (const void*)(kp + j * (4 * KSTEP) * ks) : (const void*)bp, bytes);
        }
      } else {
        Walk w;
        w.init(cax, kf < K ? kf : 0, 4 * KSTEP);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const uint32_t bytes = (kf + 4 * KSTEP * j < K) ? (uint32_t)(4 * nv) : 0u;
          cp_async16(d0 + j * (NW << 9), bytes ? (const void*)(rp + w.off) : (const void*)bp, bytes);
          w.next();
        }
      }
    } else if (mode == MODE_KSC) {
      // unaligned K-contiguous rows: 4-byte copies, lane <-> k (one coalesced line per row)
      const int kl = t & 31, wq = t >> 5;
      const int k = k0 + kl;
      const bool kok = k < K;
      const int ko = kok ? axoff32(cax, k) : 0;
      Walk rw;
      row_walk(rw, rax, roff, rlo);
#pragma unroll
      for (int i = 0; i < 4 * NV; ++i) {
        const int r = wq + NW * i;
        const uint32_t bytes = (kok && fr + NW * i < limit) ? 4u : 0u;
        cp_async4(pbase + swz(r, kl >> 2) + ((kl & 3) << 2), bytes ? (const void*)(bp + rw.off + ko) : (const void*)bp,
                  bytes);
        rw.next();
      }
    } else {
      // unaligned row-contiguous operand: 4-byte copies, lane <-> row, transposed into the K-major tile
      const int r = t % ROWS, c0 = (t / ROWS) * NV;
      const int kf = k0 + (t / ROWS) * KPTR;
      Walk w;
      w.init(cax, kf, 1);
      const bool rok = fr < limit;
      const float* rp = bp + roff;
#pragma unroll
      for (int c = 0; c < NV; ++c) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t bytes = (rok && kf + 4 * c + j < K) ? 4u : 0u;
          cp_async4(pbase + swz(r, c0 + c) + (j << 2), bytes ? (const void*)(rp + w.off) : (const void*)bp, bytes);
          w.next();
        }
      }
    }
  }
  __device__ __forceinline__ void issue(int kb, uint8_t* plane_hi, uint8_t* plane_lo) const {
    issue_stream<NV>(kb, plane_hi, base, *rowax, *colax, r_off, r_lo, ro);
    if constexpr (MUL) {
      if (mbase) issue_stream<NM>(kb, plane_lo, mbase, *mrowax, *mcolax, m_off, m_lo, mo);   // factor parks in the lo plane
    }
  }
};

template <int BN, bool X3, bool AMUL, bool ATM_>
struct TcCfg {
  static constexpr int A_PLANE = TC_BM * 128;
  static constexpr int B_PLANE = BN * 128;
  static constexpr int PLANES = X3 ? 2 : 1;
  // A through tensor memory (3xTF32 without a Hadamard factor): the raw A tile lands in shared memory, the
  // A producer warps read their row of it, split it in registers and store hi / lo into TMEM columns
  // (tcgen05.st), and the MMAs take A from TMEM (TS form).  A then needs ONE shared-memory plane and no
  // tensor-core reads of shared memory: the 3-pass product is bound by shared-memory bandwidth
  // (operand reads 3 x (A + B) per k-block), so this removes a third of the traffic and buys a 4th stage.
  static constexpr bool ATM = ATM_;
  static_assert(!ATM_ || (X3 && !AMUL), "A through tensor memory: 3xTF32 without a Hadamard factor");
  static constexpr int A_PLANES = ATM ? 1 : ((X3 || AMUL) ? 2 : 1);   // (a Hadamard factor lands in a second A plane)
  static constexpr int B_OFF = A_PLANES * A_PLANE;        // B planes follow the A planes in a stage
  static constexpr int STAGE = A_PLANES * A_PLANE + PLANES * B_PLANE;
  static constexpr int STAGES_RAW = TC_SMEM_BUDGET / STAGE;
  static constexpr int STAGES_TM = ATM ? (512 - 2 * BN) / 64 : 6;    // 64 TMEM columns (hi + lo) per stage
  static constexpr int STAGES_CAP = STAGES_TM < TC_STAGE_LIMIT ? STAGES_TM : TC_STAGE_LIMIT;
  static constexpr int STAGES = STAGES_RAW > STAGES_CAP ? STAGES_CAP : STAGES_RAW;
  static constexpr int BAR_OFF = STAGES * STAGE;
  static constexpr int EPI_OFF = BAR_OFF + 512;
  static constexpr int SMEM = EPI_OFF + TC_EPI_SMEM + 1024 /*align slack*/;
  static constexpr int TMEM_COLS = ATM ? 512 : (2 * BN < 32 ? 32 : 2 * BN);
  static constexpr int A_TM = 2 * BN;                     // first TMEM column of the A slots
};

// The CTA's flat (output tile, k-block) sequence, walked identically by every role.
struct PosIter {
  int w, kb, kb1, tm, tn, batch;
  __device__ __forceinline__ void enter(const TcParams& p) {
    int q = w;
    const int ks = q % p.nsplit; q /= p.nsplit;
    tn = q % p.ntn; q /= p.ntn;
    tm = q % p.ntm;
    batch = q / p.ntm;
    kb = ks * p.kb_per_split;
    kb1 = min(p.kblocks, kb + p.kb_per_split);
  }
  __device__ __forceinline__ bool init(const TcParams& p) {
    w = blockIdx.x;
    if (w >= p.total_items) return false;
    enter(p);
    return true;
  }
  __device__ __forceinline__ bool next(const TcParams& p) {
    if (++kb < kb1) return true;
    w += gridDim.x;
    if (w >= p.total_items) return false;
    enter(p);
    return true;
  }
};

// TMA-fed operand: the raw fp32 tile lands (already swizzled) in the hi plane -- the tensor core ignores
// the low 13 mantissa bits, so the raw tile IS the hi operand -- and these warps only derive the lo plane
// (v - trunc_tf32(v)) for the 3-pass mode.  Up to STAGES k-blocks are in flight per operand.
// Finish a landed stage in place:
//   plain   : lo = v - trunc_tf32(v)               (the raw tile IS the hi operand)
//   Hadamard: v = p * t ; hi = v ; lo = v - trunc_tf32(v)     (A only; the factor was copied to the lo plane)
// With a 128-row B tile the 4 B warps would have twice the split work of the 8 A warps; in that case
// (JOINT) all 12 producer warps wait for both tiles and take equal shares of the A and B chunks.
// Measured on B200: weight-gradient shapes 7-20 % faster; not used with a Hadamard factor or narrower
// B tiles, where the coupling of the two groups costs more than the balance gains.
// A through tensor memory: thread (warp w, lane l) owns row 32 (w % 4) + l of the raw tile and the k half w / 4.
// In team mode (TEAMS: B tiles narrower than 128 rows) the A warps form two teams of four that take alternate
// k-blocks and split BOTH k halves of their row: the chain wait -> LDS -> split -> tcgen05.st -> wait::st -> arrive
// is latency-bound per warp (ncu on sat_PQ: 1400 cycles per k-block with the TMA data two stages ahead), and two
// teams keep two k-blocks in that chain at once.
template <class Cfg, bool TEAMS>
__device__ __forceinline__ void finish_a_tmem(const uint8_t* a_raw, bool mn_major, bool plain, uint32_t tmem_slot, int t) {
  const int lane = t & 31, w = t >> 5;
  const int r = 32 * (w & 3) + lane;
#pragma unroll
  for (int half = 0; half < (TEAMS ? 2 : 1); ++half) {
  const int kh = TEAMS ? half : (w >> 2);
  uint32_t hi[16], lo[16];
  if (!mn_major) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(a_raw + swz(r, 4 * kh + c));
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hi[4 * c + j] = __float_as_uint(e[j]);
        lo[4 * c + j] = __float_as_uint(tf32_lo(e[j]));
      }
    }
  } else {
    // MN-major raw tile: atom (ka, mb) at (ka * 4 + mb) * 512, k-row kr at +128 kr, row quad q at 32-byte unit
    // (q >> 1) ^ kr, half (q & 1) * 16, element (r & 3) * 4
    const int mb = r >> 5, q = (r & 31) >> 2, e = r & 3;
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const int k = 16 * kh + kk, ka = k >> 2, kr = k & 3;
      const uint32_t off = plain ? (uint32_t)((mb << 12) + (k << 7) + ((r & 31) << 2))      // TMA boxes: [mb][k][32 rows]
                                 : (uint32_t)(((ka * (TC_BM / 32) + mb) << 9) + (kr << 7) + (((q >> 1) ^ kr) << 5) +
                                              ((q & 1) << 4) + (e << 2));
      const float v = *reinterpret_cast<const float*>(a_raw + off);
      hi[kk] = __float_as_uint(v);
      lo[kk] = __float_as_uint(tf32_lo(v));
    }
  }
  const uint32_t ta = tmem_slot + ((uint32_t)(32 * (w & 3)) << 16) + 16 * kh;
  tmem_st16(ta, hi);
  tmem_st16(ta + 32, lo);
  }
  tmem_st_wait();
  tc_fence_before();
}

// Two A-warp teams on alternate k-blocks: A through tensor memory and the B split left to the B warps.
template <class Cfg, int BN, bool IS_A>
struct TcTeams { static constexpr bool value = Cfg::ATM && BN < 128 && IS_A; };

template <class Cfg, int BN, bool X3, bool AMUL, bool IS_A>
__device__ __forceinline__ void finish_stage(uint8_t* stage, uint64_t* rawa, uint64_t* rawb, uint32_t ph, bool has_mul,
                                             int t, int pt, int lane, uint32_t tmem_slot = 0, bool a_mn = false,
                                             bool a_plain = false) {
  if constexpr (Cfg::ATM) {
    // A warps: row split into tensor memory.  With a 128-row B tile all 12 warps then share B's split
    // (the 4 B warps alone would be the critical path); narrower B tiles stay with the B warps.
    constexpr bool SHARE_B = BN == 128;
    if (IS_A) {
      if (lane == 0) mbar_wait(rawa, ph, 0);
      __syncwarp();
#if !(TC_DBG_SKIP & 1)
      finish_a_tmem<Cfg, TcTeams<Cfg, BN, IS_A>::value>(stage, a_mn, a_plain, tmem_slot, t);
#endif
      if (!SHARE_B) return;
    }
    if (lane == 0) mbar_wait(rawb, ph, 0);
    __syncwarp();
#if TC_DBG_SKIP & 2
    return;
#endif
    uint8_t* b_hi = stage + Cfg::B_OFF;
    constexpr int B_CH = BN * 8;
    constexpr int NTB = SHARE_B ? (TC_A_WARPS + TC_B_WARPS) * 32 : TC_B_WARPS * 32;
    const int tb = SHARE_B ? pt : t;
#pragma unroll
    for (int i = 0; i < (B_CH + NTB - 1) / NTB; ++i) {
      const int c = tb + i * NTB;
      if (c < B_CH) {
        const uint32_t off = (uint32_t)c * 16u;
        *reinterpret_cast<float4*>(b_hi + Cfg::B_PLANE + off) = tf32_lo4(*reinterpret_cast<const float4*>(b_hi + off));
      }
    }
    fence_proxy_async_smem();
    return;
  }
  constexpr bool JOINT = (BN == 128) && !AMUL && X3 && !Cfg::ATM;
  constexpr int A_CH = TC_BM * 8, B_CH = BN * 8;
  if (lane == 0) {
    if (JOINT || IS_A) mbar_wait(rawa, ph, 0);
    if (JOINT || !IS_A) mbar_wait(rawb, ph, 0);
  }
  __syncwarp();
  uint8_t* a_hi = stage;
  uint8_t* b_hi = stage + Cfg::B_OFF;
  if constexpr (JOINT) {
    constexpr int NTOT = (TC_A_WARPS + TC_B_WARPS) * 32;
#pragma unroll
    for (int i = 0; i < (A_CH + B_CH + NTOT - 1) / NTOT; ++i) {
      const int c = pt + i * NTOT;
      if (c < A_CH) {
        const uint32_t off = (uint32_t)c * 16u;
        *reinterpret_cast<float4*>(a_hi + Cfg::A_PLANE + off) = tf32_lo4(*reinterpret_cast<const float4*>(a_hi + off));
      } else if (c < A_CH + B_CH) {
        const uint32_t off = (uint32_t)(c - A_CH) * 16u;
        *reinterpret_cast<float4*>(b_hi + Cfg::B_PLANE + off) = tf32_lo4(*reinterpret_cast<const float4*>(b_hi + off));
      }
    }
  } else if (IS_A) {
    constexpr int NT = TC_A_WARPS * 32;
    if (AMUL && has_mul) {
#pragma unroll
      for (int i = 0; i < A_CH / NT; ++i) {
        const uint32_t off = (uint32_t)(t + i * NT) * 16u;
        const float4 v = mul4(*reinterpret_cast<const float4*>(a_hi + off),
                              *reinterpret_cast<const float4*>(a_hi + Cfg::A_PLANE + off));
        *reinterpret_cast<float4*>(a_hi + off) = v;
        if (X3) *reinterpret_cast<float4*>(a_hi + Cfg::A_PLANE + off) = tf32_lo4(v);
      }
    } else if (X3) {
#pragma unroll
      for (int i = 0; i < A_CH / NT; ++i) {
        const uint32_t off = (uint32_t)(t + i * NT) * 16u;
        *reinterpret_cast<float4*>(a_hi + Cfg::A_PLANE + off) = tf32_lo4(*reinterpret_cast<const float4*>(a_hi + off));
      }
    }
  } else if (X3) {
    constexpr int NT = TC_B_WARPS * 32;
#pragma unroll
    for (int i = 0; i < B_CH / NT; ++i) {
      const uint32_t off = (uint32_t)(t + i * NT) * 16u;
      *reinterpret_cast<float4*>(b_hi + Cfg::B_PLANE + off) = tf32_lo4(*reinterpret_cast<const float4*>(b_hi + off));
    }
  }
  fence_proxy_async_smem();          // generic-proxy writes (cp.async, st.shared) -> visible to the tensor core
}

// TMA-fed operand: the raw fp32 tile lands (already swizzled) in the hi plane; these warps only take
// part in finishing the stages.
template <class Cfg, int BN, bool X3, bool AMUL, bool IS_A>
__device__ __forceinline__ void transform_all(const TcParams& p, uint8_t* smem, uint64_t* full_bar, uint64_t* rawa_bar,
                                              uint64_t* rawb_bar, int t, int pt, int lane, uint32_t tmem_base) {
  const bool has_mul = AMUL && p.g.Amul.ptr != nullptr;
  const bool a_mn = p.a_mode == MODE_MVEC;
  PosIter pos;
  uint32_t it = 0;
  constexpr bool TEAMS = TcTeams<Cfg, BN, IS_A>::value;
  const uint32_t team = (uint32_t)(t >> 7);             // A warps 0-3 / 4-7
  for (bool ok = pos.init(p); ok; ok = pos.next(p), ++it) {
    if (TEAMS && (it & 1) != team) continue;
    const int s = it % Cfg::STAGES;
    finish_stage<Cfg, BN, X3, AMUL, IS_A>(smem + s * Cfg::STAGE, rawa_bar + s, rawb_bar + s, (it / Cfg::STAGES) & 1,
                                          has_mul, t, pt, lane, tmem_base + Cfg::A_TM + s * 64, a_mn, a_mn && p.a_tma);
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(full_bar + s);
      if (TEAMS) mbar_arrive(full_bar + s);             // (the barrier counts all eight A warps)
    }
  }
}

// Operand fed by cp.async (any load mode; an optional Hadamard factor lands in the lo plane): this
// group's copies run STAGES-1 k-blocks ahead of the stage being finished.
template <class Cfg, int BN, int ROWS, int NT, bool X3, bool AMUL, bool IS_A>
__device__ __forceinline__ void produce_async(const TcParams& p, uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar,
                                              uint64_t* rawa_bar, uint64_t* rawb_bar, int t, int pt, int lane,
                                              uint32_t tmem_base) {
  using P = Producer<ROWS, NT, IS_A && AMUL>;
  constexpr int STAGES = Cfg::STAGES;
  const dstag_gemm_t& g = p.g;
  P pr;
  pr.t = t;
  pr.K = g.K;
  if (IS_A) {
    pr.mode = p.a_mode;
    pr.rowax = &g.A.row; pr.colax = &g.A.col;
    pr.mrowax = &g.Amul.row; pr.mcolax = &g.Amul.col;
  } else {
    pr.mode = p.b_mode;
    pr.rowax = &g.B.col; pr.colax = &g.B.row;
    pr.mrowax = &g.B.col; pr.mcolax = &g.B.row;
  }
  const bool has_mul = AMUL && g.Amul.ptr != nullptr;
  uint64_t* raw_bar = IS_A ? rawa_bar : rawb_bar;
  const uint32_t plane_off = IS_A ? 0u : (uint32_t)Cfg::B_OFF;
  const uint32_t plane_bytes = IS_A ? (uint32_t)Cfg::A_PLANE : (uint32_t)Cfg::B_PLANE;
  PosIter pi, px;
  bool oki = pi.init(p), okx = px.init(p);
  if (!okx) return;
  int cur_w = -1;
  uint32_t ni = 0, nx = 0;
  auto issue_one = [&]() {
    const int s = ni % STAGES;
    const uint32_t ph = (ni / STAGES) & 1;
    if (lane == 0) mbar_wait(empty_bar + s, ph ^ 1, 0);
    __syncwarp();
    if (pi.w != cur_w) {
      cur_w = pi.w;
      if (IS_A) {
        const float* ab = (const float*)g.A.ptr + axoff64(g.A.batch, pi.batch);
        const float* mb = g.Amul.ptr ? (const float*)g.Amul.ptr + axoff64(g.Amul.batch, pi.batch) : nullptr;
        pr.begin_item(ab, mb, pi.tm * p.a_rt, g.M);
      } else {
        pr.begin_item((const float*)g.B.ptr + axoff64(g.B.batch, pi.batch), nullptr, pi.tn * ROWS, g.N);
      }
    }
    uint8_t* hi = smem + s * Cfg::STAGE + plane_off;
    pr.issue(pi.kb, hi, hi + plane_bytes);
    cp_async_arrive(raw_bar + s);
    oki = pi.next(p);
    ++ni;
  };
  for (int d = 0; d < STAGES - 1 && oki; ++d) issue_one();
  constexpr bool TEAMS = TcTeams<Cfg, BN, IS_A>::value;
  const uint32_t team = (uint32_t)(t >> 7);
  while (okx) {
    const int s = nx % STAGES;
    if (!TEAMS || (nx & 1) == team) {
      finish_stage<Cfg, BN, X3, AMUL, IS_A>(smem + s * Cfg::STAGE, rawa_bar + s, rawb_bar + s, (nx / STAGES) & 1, has_mul,
                                            t, pt, lane, tmem_base + Cfg::A_TM + s * 64, p.a_mode == MODE_MVEC, false);
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(full_bar + s);
        if (TEAMS) mbar_arrive(full_bar + s);
      }
    }
    okx = px.next(p);
    ++nx;
    if (oki) issue_one();
  }
}

template <int BN, bool X3, bool AMUL, bool ATM>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TcParams p,
                                                                const __grid_constant__ TcMaps maps) {
  using Cfg = TcCfg<BN, X3, AMUL, ATM>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* full_bar = bars;                          // [STAGES] producers -> MMA
  uint64_t* empty_bar = bars + Cfg::STAGES;           // [STAGES] MMA (commit) -> producers
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;       // [2] MMA (commit) -> epilogue
  uint64_t* tempty_bar = bars + 2 * Cfg::STAGES + 2;  // [2] epilogue -> MMA
  uint64_t* rawa_bar = bars + 2 * Cfg::STAGES + 4;    // [STAGES] TMA bytes of the A tile landed
  uint64_t* rawb_bar = bars + 3 * Cfg::STAGES + 4;    // [STAGES] TMA bytes of the B tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * Cfg::STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const dstag_gemm_t& g = p.g;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar + s, TC_A_WARPS + TC_B_WARPS);
      mbar_init(empty_bar + s, 1);
      mbar_init(rawa_bar + s, (!AMUL && p.a_tma) ? 1 : TC_A_WARPS * 32);
      mbar_init(rawb_bar + s, p.b_tma ? 1 : TC_B_WARPS * 32);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + a, 1);
      mbar_init(tempty_bar + a, TC_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == TC_MMA_WARP) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < TC_A_WARPS + TC_B_WARPS) {
    // =============================== producers =========================================
    const int pt = threadIdx.x;                          // producer thread index over both groups
    if (warp < TC_A_WARPS) {
      if (!AMUL && p.a_tma)
        transform_all<Cfg, BN, X3, AMUL, true>(p, smem, full_bar, rawa_bar, rawb_bar, threadIdx.x, pt, lane, tmem_base);
      else
        produce_async<Cfg, BN, TC_BM, TC_A_WARPS * 32, X3, AMUL, true>(p, smem, full_bar, empty_bar, rawa_bar, rawb_bar,
                                                                       threadIdx.x, pt, lane, tmem_base);
    } else {
      if (p.b_tma)
        transform_all<Cfg, BN, X3, AMUL, false>(p, smem, full_bar, rawa_bar, rawb_bar, threadIdx.x - TC_A_WARPS * 32, pt,
                                                lane, tmem_base);
      else
        produce_async<Cfg, BN, BN, TC_B_WARPS * 32, X3, AMUL, false>(p, smem, full_bar, empty_bar, rawa_bar, rawb_bar,
                                                                     threadIdx.x - TC_A_WARPS * 32, pt, lane, tmem_base);
    }
  } else if (warp == TC_TMA_WARP) {
    // =============================== TMA issuer ========================================
    // (whole warp in the loop, one elected lane issues: no waterfall loops around UTMALDG, see the MMA issuer)
    if (p.a_tma || p.b_tma) {
      const bool a_mn = p.a_mode == MODE_MVEC, b_mn = p.b_mode == MODE_MVEC;
      PosIter pos;
      uint32_t it = 0;
      for (bool ok = pos.init(p); ok; ok = pos.next(p), ++it) {
        const int s = it % Cfg::STAGES;
        const uint32_t ph = (it / Cfg::STAGES) & 1;
        mbar_wait(empty_bar + s, ph ^ 1, 0);            // the MMAs that read this stage have retired
        if (elect_one()) {
        uint8_t* st = smem + s * Cfg::STAGE;
        const int k0 = pos.kb * TC_BK;
        if (!AMUL && p.a_tma) {
          const int bl = p.a_bdim[1] ? pos.batch % g.A.batch.d1 : 0, bh = p.a_bdim[0] ? pos.batch / g.A.batch.d1 : 0;
          mbar_expect_tx(rawa_bar + s, p.a_rt * 128);
          if (a_mn) {
#pragma unroll
            for (int mb = 0; mb < TC_BM / 32; ++mb)
              tma_load_4d(st + mb * 4096, &maps.a, rawa_bar + s, pos.tm * TC_BM + mb * 32, k0, bl, bh);
          } else if (p.a_rag_d1) {
            // ragged tile: whole x's only -- box (32 k, d1 t, a_rt / d1 x) lands as a_rt consecutive tile rows
            tma_load_4d(st, &maps.a, rawa_bar + s, k0, 0, pos.tm * (p.a_rt / p.a_rag_d1), bl);
          } else {
            tma_load_4d(st, &maps.a, rawa_bar + s, k0, pos.tm * TC_BM, bl, bh);
          }
        }
        if (p.b_tma) {
          const int bl = p.b_bdim[1] ? pos.batch % g.B.batch.d1 : 0, bh = p.b_bdim[0] ? pos.batch / g.B.batch.d1 : 0;
          uint8_t* sb = st + Cfg::B_OFF;
          mbar_expect_tx(rawb_bar + s, Cfg::B_PLANE);
          if (b_mn) {
#pragma unroll
            for (int mb = 0; mb < BN / 32; ++mb)
              tma_load_4d(sb + mb * 4096, &maps.b, rawb_bar + s, pos.tn * BN + mb * 32, k0, bl, bh);
          } else {
            tma_load_4d(sb, &maps.b, rawb_bar + s, k0, pos.tn * BN, bl, bh);
          }
        }
        }
        __syncwarp();
      }
    }
  } else if (warp == TC_MMA_WARP) {
    // =============================== MMA issuer ========================================
    // The whole warp walks the loops (uniform control flow) and ONE lane chosen by elect.sync issues: inside an
    // elect.sync region the compiler knows a single thread is active and feeds tcgen05.mma / commit from uniform
    // registers directly; under `if (lane == 0)` it wrapped every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY
    // waterfall loop (~60 issue cycles per MMA against a 32-cycle tensor-pipe floor at BN = 64).
    {
      const bool a_mn = p.a_mode == MODE_MVEC, b_mn = p.b_mode == MODE_MVEC;
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      // descriptor advance per K step of 8 (16-byte units): +32 B inside the swizzle row (K-major),
      // two 4-k atoms = +1024 B (MN-major)
      const uint32_t idesc_ts = idesc & ~(1u << 15);     // A from tensor memory is always K-major
      const uint64_t a_adv = a_mn ? (uint64_t)((TC_BM / 32) * 1024 >> 4) : 2ull;
      const uint64_t b_adv = b_mn ? (uint64_t)((BN / 32) * 1024 >> 4) : 2ull;
      // Short chains: the tensor core's fp32 accumulator TRUNCATES at every k-step, a bias that grows with
      // the chain and shows up in products whose result is much smaller than its terms.  With chunk_kb > 0 a
      // tile is accumulated in passes of chunk_kb k-blocks that alternate between the two TMEM accumulators;
      // the epilogue warps add the passes in registers (round to nearest).  `pass` counts accumulator passes
      // the same way in both roles; without chunking a pass is a whole tile.
      uint32_t it = 0, pass = 0;
      for (int w = blockIdx.x; w < p.total_items; w += gridDim.x) {
        const int ks = w % p.nsplit;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
        const int chunk = p.chunk_kb > 0 ? p.chunk_kb : kb1 - kb0;
        for (int c0 = kb0; c0 < kb1; c0 += chunk, ++pass) {
        const int c1 = min(kb1, c0 + chunk);
        const uint32_t acc = pass & 1;
        mbar_wait(tempty_bar + acc, ((pass >> 1) & 1) ^ 1, 0);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = c0; kb < c1; ++kb, ++it) {
          const int s = it % Cfg::STAGES;
          const uint32_t ph = (it / Cfg::STAGES) & 1;
          mbar_wait(full_bar + s, ph, 0);
          tc_fence_after();
          if (elect_one()) {
          const uint32_t sa = smem_u32(smem + s * Cfg::STAGE);
          const uint32_t sb = sa + Cfg::B_OFF;
          const uint64_t a_hi = a_mn ? make_desc_mn(sa, (TC_BM / 32) * 512) : make_desc_k(sa);
          const uint64_t a_lo = a_mn ? make_desc_mn(sa + Cfg::A_PLANE, (TC_BM / 32) * 512) : make_desc_k(sa + Cfg::A_PLANE);
          const uint64_t b_hi = b_mn ? make_desc_mn(sb, (BN / 32) * 512) : make_desc_k(sb);
          const uint64_t b_lo = b_mn ? make_desc_mn(sb + Cfg::B_PLANE, (BN / 32) * 512) : make_desc_k(sb + Cfg::B_PLANE);
#pragma unroll
          for (int kk = 0; kk < TC_BK / 8; ++kk) {
            const uint64_t aa = a_adv * kk, ba = b_adv * kk;
            const uint32_t first = (kb == c0 && kk == 0) ? 0u : 1u;
            if constexpr (Cfg::ATM) {
              const uint32_t ta = tmem_base + Cfg::A_TM + s * 64 + 8 * kk;     // hi columns; lo at +32
#if TC_DBG_SKIP & 4
              umma_tf32_ts(d_tmem, ta, b_hi + ba, idesc_ts, first);
#else
              umma_tf32_ts(d_tmem, ta + 32, b_hi + ba, idesc_ts, first);
              umma_tf32_ts(d_tmem, ta, b_lo + ba, idesc_ts, 1u);
              umma_tf32_ts(d_tmem, ta, b_hi + ba, idesc_ts, 1u);
#endif
            } else if (X3) {
              umma_tf32(d_tmem, a_lo + aa, b_hi + ba, idesc, first);
              umma_tf32(d_tmem, a_hi + aa, b_lo + ba, idesc, 1u);
              umma_tf32(d_tmem, a_hi + aa, b_hi + ba, idesc, 1u);
            } else {
              umma_tf32(d_tmem, a_hi + aa, b_hi + ba, idesc, first);
            }
          }
          umma_commit(empty_bar + s);          // stage reusable once these MMAs have read it
          if (kb == c1 - 1) umma_commit(tfull_bar + acc);          // accumulator pass complete -> epilogue
          }
          __syncwarp();
        }
        }
      }
    }
  } else {
    // =============================== epilogue ==========================================
    const int q4 = warp & 3;                   // TMEM lane quarter this warp may access
    float* tbuf = reinterpret_cast<float*>(smem + Cfg::EPI_OFF) + q4 * (32 * 36);
    const bool c_unit = g.C.col.d1 == 1 && g.C.col.s0 == 1;
    const bool ca_unit = !g.Cadd.ptr || (g.Cadd.col.d1 == 1 && g.Cadd.col.s0 == 1);
    const bool cm_unit = !g.Cmul.ptr || (g.Cmul.col.d1 == 1 && g.Cmul.col.s0 == 1);
    auto inner1 = [](const dstag_axis_t& a) { return (a.d1 == 1 ? a.s0 : a.s1) == 1; };
    const bool c_inner1 = inner1(g.C.col);
    const bool ca_inner1 = !g.Cadd.ptr || inner1(g.Cadd.col) || (g.Cadd.col.s0 == 0 && g.Cadd.col.d1 == 1);
    const bool cm_inner1 = !g.Cmul.ptr || inner1(g.Cmul.col) || (g.Cmul.col.s0 == 0 && g.Cmul.col.d1 == 1);
    // 4 consecutive columns (starting at a multiple of 4) are 4 consecutive, 16-byte aligned floats
    auto quad = [](const dstag_operand_t& o) {
      const dstag_axis_t& a = o.col;
      const bool runs = a.d1 == 1 ? a.s0 == 1 : (a.s1 == 1 && a.d1 % 4 == 0 && a.s0 % 4 == 0);
      return runs && (reinterpret_cast<uintptr_t>(o.ptr) & 15) == 0 && o.batch.s0 % 4 == 0 &&
             (o.batch.d1 == 1 || o.batch.s1 % 4 == 0);
    };
    const bool c_quad = g.accumulate != DSTAG_ACC_ATOMIC && g.split_stride % 4 == 0 && quad(g.C) && (!g.Cadd.ptr || quad(g.Cadd)) &&
                        (!g.Cmul.ptr || quad(g.Cmul));
    uint32_t pass = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x) {
      int q = w;
      const int ks = q % p.nsplit; q /= p.nsplit;
      const int tn = q % p.ntn; q /= p.ntn;
      const int tm = q % p.ntm;
      const int batch = q / p.ntm;
      const int kb0e = ks * p.kb_per_split, kb1e = min(p.kblocks, kb0e + p.kb_per_split);
      const int npass = p.chunk_kb > 0 ? (kb1e - kb0e + p.chunk_kb - 1) / p.chunk_kb : 1;
      const int m = tm * p.a_rt + q4 * 32 + lane;               // tile rows beyond a_rt are padding of a ragged tile
      const bool mok = q4 * 32 + lane < p.a_rt && m < g.M;
      float* Cb = (float*)g.C.ptr + axoff64(g.C.batch, batch) + (long long)ks * g.split_stride;
      const float* Cab = (g.Cadd.ptr && ks == 0) ? (const float*)g.Cadd.ptr + axoff64(g.Cadd.batch, batch) : nullptr;
      const float* Cmb = g.Cmul.ptr ? (const float*)g.Cmul.ptr + axoff64(g.Cmul.batch, batch) : nullptr;
      const int c_ro = mok ? axoff32(g.C.row, m) : 0;
      const int ca_ro = (Cab && mok) ? axoff32(g.Cadd.row, m) : 0;
      const int cm_ro = (Cmb && mok) ? axoff32(g.Cmul.row, m) : 0;
      float* C = Cb + c_ro;
      const float* Ca = Cab ? Cab + ca_ro : nullptr;
      const float* Cm = Cmb ? Cmb + cm_ro : nullptr;
      // all passes but the last (short-chain mode, BN == 32 only): fold them into registers and hand the
      // accumulator back at once
      // (partial sums parked in this warp's 32 x 32 shared-memory tile, pitch 33: the store paths below need
      //  the registers, and the kernel is capped at 96 per thread)
      if constexpr (BN == 32) {
        for (int ps = 0; ps + 1 < npass; ++ps, ++pass) {
          const uint32_t a0 = pass & 1;
          if (lane == 0) mbar_wait(tfull_bar + a0, (pass >> 1) & 1, 200);
          __syncwarp();
          tc_fence_after();
          float t[32];
          tmem_ld32(tmem_base + a0 * BN + ((uint32_t)(q4 * 32) << 16), t);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar + a0);          // the accumulator is free again: MMAs go on
#pragma unroll
          for (int j = 0; j < 32; ++j) tbuf[lane * 33 + j] = ps == 0 ? t[j] : tbuf[lane * 33 + j] + t[j];
        }
      }
      const uint32_t acc = pass & 1;
      if (lane == 0) mbar_wait(tfull_bar + acc, (pass >> 1) & 1, 1000);
      __syncwarp();                 // tcgen05.ld is .sync.aligned: the warp must be converged
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        float v[32];
        tmem_ld32(tmem_base + acc * BN + cc * 32 + ((uint32_t)(q4 * 32) << 16), v);
        if constexpr (BN == 32) {
          if (npass > 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += tbuf[lane * 33 + j];
            __syncwarp();                                        // tbuf is reused by the transposing store paths
          }
        }
        const int n0 = tn * BN + cc * 32;
        if (n0 >= g.N) continue;                       // warp-uniform
        const bool full = n0 + 32 <= g.N;
        // (a) row-contiguous, 16B-aligned output: every thread streams its own row with float4
        bool vec = c_unit && ca_unit && cm_unit && g.accumulate != DSTAG_ACC_ATOMIC && full &&
                   ((reinterpret_cast<uintptr_t>(C + n0) & 15) == 0) &&
                   (!Ca || (reinterpret_cast<uintptr_t>(Ca + n0) & 15) == 0) &&
                   (!Cm || (reinterpret_cast<uintptr_t>(Cm + n0) & 15) == 0);
        vec = __all_sync(0xffffffffu, vec || !mok);
        const bool b4ok = c_quad && full && __all_sync(0xffffffffu, ((c_ro | ca_ro | cm_ro) & 3) == 0);
        if (vec && !(p.epi_coalesce && b4ok)) {
          if (mok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o = make_float4(g.alpha * v[j], g.alpha * v[j + 1], g.alpha * v[j + 2], g.alpha * v[j + 3]);
              if (Ca) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(Ca + n0 + j));
                o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
              }
              if (Cm) o = mul4(o, __ldg(reinterpret_cast<const float4*>(Cm + n0 + j)));
              if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              float4* dst = reinterpret_cast<float4*>(C + n0 + j);
              if (g.accumulate == DSTAG_ACC_ADD) {
                const float4 old = *dst;
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
              }
              *dst = o;
            }
          }
        } else if (b4ok) {
          // (b4) columns contiguous in 16-byte aligned runs of a multiple of 4 (e.g. n = (x, t), t innermost):
          //      transpose the 32 x 32 block through shared memory; every lane then stores float4s,
          //      8 lanes covering the 32 columns of one row, 4 rows per instruction
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(tbuf + lane * 36 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          __syncwarp();
          const int q = lane & 7, rsub = lane >> 3;
          const int n = n0 + 4 * q;
          const int co = axoff32(g.C.col, n);
          const int cao = Cab ? axoff32(g.Cadd.col, n) : 0;
          const int cmo = Cmb ? axoff32(g.Cmul.col, n) : 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = rsub + 4 * i;
            const int mr = tm * p.a_rt + q4 * 32 + rr;
            const int ro = __shfl_sync(0xffffffffu, c_ro, rr);
            const int rao = __shfl_sync(0xffffffffu, ca_ro, rr);
            const int rmo = __shfl_sync(0xffffffffu, cm_ro, rr);
            if (q4 * 32 + rr < p.a_rt && mr < g.M) {
              const float4 tv = *reinterpret_cast<const float4*>(tbuf + rr * 36 + 4 * q);
              float4 o = make_float4(g.alpha * tv.x, g.alpha * tv.y, g.alpha * tv.z, g.alpha * tv.w);
              if (Cab) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(Cab + rao + cao));
                o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
              }
              if (Cmb) o = mul4(o, __ldg(reinterpret_cast<const float4*>(Cmb + rmo + cmo)));
              if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              float4* dst = reinterpret_cast<float4*>(Cb + ro + co);
              if (g.accumulate == DSTAG_ACC_ADD) {
                const float4 old = *dst;
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
              }
              *dst = o;
            }
          }
          __syncwarp();
        } else if (c_inner1 && ca_inner1 && cm_inner1) {
          // (b) columns contiguous in memory but unaligned / narrow / atomic / two-level (e.g. n = (x, t)):
          //     transpose the 32 x 32 block through shared memory so that consecutive lanes touch
          //     consecutive columns (coalesced)
#pragma unroll
          for (int j = 0; j < 32; ++j) tbuf[lane * 33 + j] = v[j];
          __syncwarp();
          const int n = n0 + lane;
          const bool nok = n < g.N;
          const int co = nok ? axoff32(g.C.col, n) : 0;
          const int cao = (Cab && nok) ? axoff32(g.Cadd.col, n) : 0;
          const int cmo = (Cmb && nok) ? axoff32(g.Cmul.col, n) : 0;
#pragma unroll 4
          for (int rr = 0; rr < 32; ++rr) {
            const int mr = tm * p.a_rt + q4 * 32 + rr;
            const int ro = __shfl_sync(0xffffffffu, c_ro, rr);
            const int rao = __shfl_sync(0xffffffffu, ca_ro, rr);
            const int rmo = __shfl_sync(0xffffffffu, cm_ro, rr);
            if (q4 * 32 + rr < p.a_rt && mr < g.M && nok) {
              float o = g.alpha * tbuf[rr * 33 + lane];
              if (Cab) o += __ldg(Cab + rao + cao);
              if (Cmb) o *= __ldg(Cmb + rmo + cmo);
              if (g.relu) o = fmaxf(o, 0.f);
              float* dst = Cb + ro + co;
              if (g.accumulate == DSTAG_ACC_STORE) *dst = o;
              else if (g.accumulate == DSTAG_ACC_ADD) *dst += o;
              else atomicAdd(dst, o);
            }
          }
          __syncwarp();
        } else if (mok) {
          // (c) column-strided output (rows contiguous in memory: consecutive lanes = consecutive rows, so
          //     every store instruction still writes whole lines): lanes walk rows, column offsets stepped
          //     without divisions
          Walk wc, wa, wm;
          wc.init(g.C.col, n0, 1);
          if (Ca) wa.init(g.Cadd.col, n0, 1);
          if (Cm) wm.init(g.Cmul.col, n0, 1);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + j;
            const int co = wc.off, cao = Ca ? wa.off : 0, cmo = Cm ? wm.off : 0;
            wc.next();
            if (Ca) wa.next();
            if (Cm) wm.next();
            if (n < g.N) {
              float o = g.alpha * v[j];
              if (Ca) o += __ldg(Ca + cao);
              if (Cm) o *= __ldg(Cm + cmo);
              if (g.relu) o = fmaxf(o, 0.f);
              float* dst = C + co;
              if (g.accumulate == DSTAG_ACC_STORE) *dst = o;
              else if (g.accumulate == DSTAG_ACC_ADD) *dst += o;
              else atomicAdd(dst, o);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + acc);
      ++pass;
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == TC_MMA_WARP) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static long long axis_extent(const dstag_axis_t& a, long long n) {   // max offset reached by index < n
  if (n <= 0) return 0;
  if (a.d1 == 1) return (n - 1) * a.s0;
  return ((n - 1) / a.d1) * a.s0 + (long long)(a.d1 - 1) * a.s1;
}
static bool axis_nonneg(const dstag_axis_t& a) { return a.s0 >= 0 && (a.d1 == 1 || a.s1 >= 0); }
static bool axis_mult4(const dstag_axis_t& a) { return a.s0 % 4 == 0 && (a.d1 == 1 || a.s1 % 4 == 0); }
static long long inner_stride(const dstag_axis_t& a) { return a.d1 == 1 ? a.s0 : a.s1; }

static bool operand_fits32(const dstag_operand_t& o, long long nrow, long long ncol) {
  if (!o.ptr) return true;
  if (!axis_nonneg(o.batch) || !axis_nonneg(o.row) || !axis_nonneg(o.col)) return false;
  // Walk keeps s0 - d1 * s1 and step products in 32 bits as well
  return axis_extent(o.row, nrow) + axis_extent(o.col, ncol) < (1ll << 30);
}

bool gemm_tc_supported(const dstag_gemm_t* g) {
  if (g->N < 8 || g->K < 8 || g->M < 8) return false;
  if ((long long)g->M * g->N * g->K * g->batch < (1ll << 19)) return false;   // tiny: not worth a persistent launch
  return operand_fits32(g->A, g->M, g->K) && operand_fits32(g->Amul, g->M, g->K) &&
         operand_fits32(g->B, g->K, g->N) && operand_fits32(g->C, g->M, g->N) &&
         operand_fits32(g->Cadd, g->M, g->N) && operand_fits32(g->Cmul, g->M, g->N);
}

static bool aligned16(const dstag_operand_t& o) {
  return (reinterpret_cast<uintptr_t>(o.ptr) & 15) == 0 && axis_mult4(o.batch);
}
// pick the cheapest load mode an operand (and its optional Hadamard factor) admits
static int pick_mode(const dstag_operand_t& o, const dstag_axis_t& kax, const dstag_axis_t& rowax,
                     const dstag_operand_t* m, const dstag_axis_t* mkax, const dstag_axis_t* mrowax) {
  const bool has_m = m && m->ptr;
  // "contiguous in aligned runs": single-level stride 1, or two-level with a contiguous inner level whose
  // length is a multiple of 4 (16-byte chunks never straddle a run), e.g. K = (k, f) with f innermost
  auto runs4 = [](const dstag_axis_t& a) {
    return a.d1 == 1 ? a.s0 == 1 : (a.s1 == 1 && a.d1 % 4 == 0 && a.s0 % 4 == 0);
  };
  auto kvec = [&](const dstag_operand_t& x, const dstag_axis_t& k, const dstag_axis_t& r) {
    return runs4(k) && axis_mult4(r) && aligned16(x);
  };
  auto mvec = [&](const dstag_operand_t& x, const dstag_axis_t& k, const dstag_axis_t& r) {
    return runs4(r) && axis_mult4(k) && aligned16(x);
  };
  if (kvec(o, kax, rowax) && (!has_m || kvec(*m, *mkax, *mrowax))) return MODE_KVEC;
  if (mvec(o, kax, rowax) && (!has_m || mvec(*m, *mkax, *mrowax))) return MODE_MVEC;
  // scalar modes: lanes along whichever index is contiguous in memory
  if (inner_stride(kax) == 1 && inner_stride(rowax) != 1) return MODE_KSC;
  return MODE_ROWG;
}

// ---- TMA tensor maps -------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* env = getenv("DSTAG_TC_TMA");
    if (env && env[0] == '0') return nullptr;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// Rank-4 map (inner, outer, batch_lo, batch_hi) of one operand in its KVEC (inner = K) or MVEC (inner = rows)
// orientation.  Returns false when the view is not expressible (two-level row / K axis, odd strides).
static bool make_tmap(CUtensorMap* map, int* bdim, const dstag_operand_t& o, const dstag_axis_t& kax,
                      const dstag_axis_t& rowax, int mode, long long K, long long rows, long long batch,
                      int box_rows, bool plain_mn = false) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc || !o.ptr) return false;
  // MN-major tiles for the tensor core need 512-byte atoms per (k-atom, row block), which a tiled copy cannot
  // produce; an A tile that goes through registers into tensor memory can take plain [k][32 rows] boxes
  if (mode != MODE_KVEC && !(mode == MODE_MVEC && plain_mn)) return false;
  if (kax.d1 != 1 || rowax.d1 != 1) return false;
  const long long inner_n = mode == MODE_KVEC ? K : rows, outer_n = mode == MODE_KVEC ? rows : K;
  const long long outer_stride = mode == MODE_KVEC ? rowax.s0 : kax.s0;     // elements
  if (outer_stride <= 0 || outer_stride % 4 != 0) return false;
  const long long nb_lo = o.batch.d1 > 1 ? o.batch.d1 : 1;
  const long long nb_hi = (batch + nb_lo - 1) / nb_lo;
  const long long s_lo = o.batch.d1 > 1 ? o.batch.s1 : 0, s_hi = o.batch.s0;
  bdim[1] = (nb_lo > 1 && s_lo > 0) ? 1 : 0;
  bdim[0] = (nb_hi > 1 && s_hi > 0) ? 1 : 0;
  if ((bdim[1] && s_lo % 4 != 0) || (bdim[0] && s_hi % 4 != 0)) return false;
  cuuint64_t dims[4] = {(cuuint64_t)inner_n, (cuuint64_t)outer_n, (cuuint64_t)(bdim[1] ? nb_lo : 1),
                        (cuuint64_t)(bdim[0] ? nb_hi : 1)};
  const cuuint64_t dummy = (cuuint64_t)outer_stride * 4;                    // stride of a size-1 dimension
  cuuint64_t strides[3] = {(cuuint64_t)outer_stride * 4, bdim[1] ? (cuuint64_t)s_lo * 4 : dummy,
                           bdim[0] ? (cuuint64_t)s_hi * 4 : dummy};
  cuuint32_t box[4] = {32u, (cuuint32_t)(mode == MODE_KVEC ? box_rows : 32), 1u, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUtensorMapSwizzle sw = mode == MODE_KVEC ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(o.ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// A row axis (b, n) whose inner level is contiguous (an activation stored (..., N) with the batch index
// further out) defeats the row-contiguous vector path, because 4-row groups would straddle the level
// boundary.  Such a problem is the same as a batched one over the outer level: M' = d1 rows per batch,
// batch' = batch * M / d1.  Rewrites the descriptor in place; false = left untouched.
static bool split_rows_to_batch(dstag_gemm_t& g) {
  const int d1 = g.A.row.d1;
  // (inner level contiguous: restores the row-contiguous vector path; inner level strided, e.g. rows (k, j)
  //  of the Chebyshev output gradient: each batch member then has single-level rows and can use TMA)
  if (d1 < 128 || g.M % d1 != 0 || g.M == d1) return false;
  const int q = g.M / d1;
  dstag_operand_t* rowops[] = {&g.A, &g.Amul, &g.C, &g.Cadd, &g.Cmul};
  for (dstag_operand_t* o : rowops) {
    if (!o->ptr) continue;
    if (o->row.d1 != d1 && o->row.d1 != 1) return false;
    if (g.batch > 1 && o->batch.d1 != 1) return false;
  }
  if (g.batch > 1 && g.B.batch.d1 != 1) return false;
  for (dstag_operand_t* o : rowops) {
    if (!o->ptr) continue;
    const long long outer = o->row.d1 == d1 ? o->row.s0 : o->row.s0 * d1;
    const long long inner = o->row.d1 == d1 ? o->row.s1 : o->row.s0;
    o->batch.s0 = g.batch > 1 ? o->batch.s0 : 0;
    o->batch.s1 = outer;
    o->batch.d1 = q;
    o->row.s0 = inner;
    o->row.s1 = 0;
    o->row.d1 = 1;
  }
  g.B.batch.s0 = g.batch > 1 ? g.B.batch.s0 : 0;
  g.B.batch.s1 = 0;
  g.B.batch.d1 = q;
  g.batch *= q;
  g.M = d1;
  return true;
}

// The same for a column index (b, n) of the output / second operand (the pre_conv input gradient writes
// rows of N = 883 nodes per (b, f, t)): as a batch over b the output columns are one contiguous,
// 16-byte aligned run, so the epilogue streams float4 rows instead of transposing scalar stores.
static bool split_cols_to_batch(dstag_gemm_t& g) {
  const int d1 = g.C.col.d1;
  if (d1 < 128 || g.C.col.s1 != 1 || g.N % d1 != 0 || g.N == d1) return false;
  const int q = g.N / d1;
  if (g.batch > 1) return false;                       // keep it simple: only un-batched problems
  dstag_operand_t* colops[] = {&g.C, &g.Cadd, &g.Cmul};
  for (dstag_operand_t* o : colops) {
    if (!o->ptr) continue;
    if (o->col.d1 != d1 && o->col.d1 != 1) return false;
  }
  if (g.B.col.d1 != d1 && g.B.col.d1 != 1) return false;
  auto split = [&](dstag_operand_t* o) {
    const long long outer = o->col.d1 == d1 ? o->col.s0 : o->col.s0 * d1;
    const long long inner = o->col.d1 == d1 ? o->col.s1 : o->col.s0;
    o->batch.s0 = outer; o->batch.s1 = 0; o->batch.d1 = 1;
    o->col.s0 = inner; o->col.s1 = 0; o->col.d1 = 1;
  };
  for (dstag_operand_t* o : colops)
    if (o->ptr) split(o);
  split(&g.B);
  g.A.batch.s0 = 0; g.A.batch.s1 = 0; g.A.batch.d1 = 1;
  if (g.Amul.ptr) { g.Amul.batch.s0 = 0; g.Amul.batch.s1 = 0; g.Amul.batch.d1 = 1; }
  g.batch = q;
  g.N = d1;
  return true;
}

// Ragged-tile map for an A operand whose rows are (x, t) windows of a channels-last activation (the GTU
// convolutions): rows two-level with d1 = windows per node, K contiguous.  A tile takes whole nodes only:
// a_rt = floor(128 / d1) * d1 rows, box (32 k, d1 t, a_rt / d1 x); the unused tile rows are never stored.
static bool make_tmap_ragged(CUtensorMap* map, const dstag_gemm_t& g, int* a_rt) {
  EncodeTiledFn enc = get_encode_tiled();
  const dstag_operand_t& o = g.A;
  if (!enc || !o.ptr || g.batch != 1) return false;
  const dstag_axis_t& r = o.row;
  const dstag_axis_t& k = o.col;
  if (r.d1 < 2 || r.d1 > 128 || k.d1 != 1 || k.s0 != 1 || g.M % r.d1 != 0) return false;
  if (r.s0 <= 0 || r.s1 <= 0 || r.s0 % 4 != 0 || r.s1 % 4 != 0 || (reinterpret_cast<uintptr_t>(o.ptr) & 15) != 0) return false;
  const int xb = 128 / r.d1;
  if (xb * r.d1 < 96) return false;                       // more than a quarter of the tile would be padding
  cuuint64_t dims[4] = {(cuuint64_t)g.K, (cuuint64_t)r.d1, (cuuint64_t)(g.M / r.d1), 1};
  cuuint64_t strides[3] = {(cuuint64_t)r.s1 * 4, (cuuint64_t)r.s0 * 4, (cuuint64_t)r.s0 * 4};
  cuuint32_t box[4] = {32u, (cuuint32_t)r.d1, (cuuint32_t)xb, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(o.ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return false;
  *a_rt = xb * r.d1;
  return true;
}

// Both the opt-in shared-memory limit of a kernel and the SM count are properties of a DEVICE: a process
// that drives several GPUs (one replica per device in one process) must set / query them on each.
constexpr int kMaxDevices = 64;
static int g_num_sms[kMaxDevices] = {0};

static int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
static int num_sms(int dev) {
  if (g_num_sms[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    g_num_sms[dev] = n > 0 ? n : kNumSMs;
  }
  return g_num_sms[dev];
}

template <int BN, bool X3, bool AMUL, bool ATM>
static int launch_tc(const TcParams& p, const TcMaps& maps, cudaStream_t stream) {
  using Cfg = TcCfg<BN, X3, AMUL, ATM>;
  static bool configured[kMaxDevices] = {false};      // per instantiation AND per device
  const int dev = current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, X3, AMUL, ATM>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) {
      set_error("dstag_gemm(tc): cudaFuncSetAttribute(%d) failed: %s", Cfg::SMEM, cudaGetErrorString(e));
      return (int)e;
    }
    configured[dev] = true;
  }
  const int nsm = num_sms(dev);
  const int grid = p.total_items < nsm ? p.total_items : nsm;
  gemm_tc_kernel<BN, X3, AMUL, ATM><<<grid, TC_THREADS, Cfg::SMEM, stream>>>(p, maps);
  DSTAG_RETURN_LAUNCH("dstag_gemm(tcgen05)");
}

int gemm_tc(const dstag_gemm_t* g_in, cudaStream_t stream) {
  TcParams p;
  p.g = *g_in;
  if (!split_rows_to_batch(p.g)) split_cols_to_batch(p.g);
  const dstag_gemm_t* g = &p.g;
  const int bn = g->N > 64 ? 128 : (g->N > 32 ? 64 : 32);
  p.a_mode = pick_mode(g->A, g->A.col, g->A.row, &g->Amul, &g->Amul.col, &g->Amul.row);
  p.b_mode = pick_mode(g->B, g->B.row, g->B.col, nullptr, nullptr, nullptr);
  const bool x3 = g->precision == DSTAG_PREC_TF32X3;
  const bool am = g->Amul.ptr != nullptr;
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  p.a_bdim[0] = p.a_bdim[1] = p.b_bdim[0] = p.b_bdim[1] = 0;
  p.a_rt = TC_BM;
  p.a_rag_d1 = 0;
  {
    // long reductions (the epilogue hides behind the next tile's main loop): whole-line stores through the transposing
    // path are 4-12 % faster than a thread streaming its own row (32 lines per store instruction); short-K launches are
    // epilogue bound and keep the cheaper per-thread path.  DSTAG_TC_EPI = 0 / 1 forces either.
    static int epi = -1;
    if (epi < 0) { const char* e = getenv("DSTAG_TC_EPI"); epi = e ? (e[0] == '1' ? 1 : 0) : 2; }
    p.epi_coalesce = epi == 2 ? (g->K >= 256 ? 1 : 0) : epi;
  }
  p.a_tma = (!am && make_tmap(&maps.a, p.a_bdim, g->A, g->A.col, g->A.row, p.a_mode, g->K, g->M, g->batch, TC_BM)) ? 1 : 0;
  static int mn_tma = -1;
  if (mn_tma < 0) { const char* e = getenv("DSTAG_TC_TMA_MN"); mn_tma = (e && e[0] == '0') ? 0 : 1; }
  if (!p.a_tma && !am && x3 && mn_tma && p.a_mode == MODE_MVEC &&
      make_tmap(&maps.a, p.a_bdim, g->A, g->A.col, g->A.row, p.a_mode, g->K, g->M, g->batch, TC_BM, true))
    p.a_tma = 1;                         // MN-major A by TMA: only valid together with A through tensor memory
  if (!p.a_tma && !am && p.a_mode == MODE_KVEC) {
    static int special = -1;
    if (special < 0) { const char* e = getenv("DSTAG_TC_TMA_SPECIAL"); special = (e && e[0] == '0') ? 0 : 1; }
    int rt = 0;
    if (special && make_tmap_ragged(&maps.a, *g, &rt)) {
      p.a_tma = 1;
      p.a_rt = rt;
      p.a_rag_d1 = g->A.row.d1;
    }
  }
  p.b_tma = make_tmap(&maps.b, p.b_bdim, g->B, g->B.row, g->B.col, p.b_mode, g->K, g->N, g->batch, bn) ? 1 : 0;
  // short accumulation chains: only with a 32-column tile (the partial sums live in 32 registers per thread)
  p.chunk_kb = (g->chain > 0 && bn == 32) ? (g->chain + TC_BK - 1) / TC_BK : 0;
  p.ntm = (g->M + p.a_rt - 1) / p.a_rt;
  p.ntn = (g->N + bn - 1) / bn;
  p.kblocks = (g->K + TC_BK - 1) / TC_BK;
  int split = g->split_k < 1 ? 1 : g->split_k;
  if (split > p.kblocks) split = p.kblocks;
  p.kb_per_split = (p.kblocks + split - 1) / split;
  p.nsplit = (p.kblocks + p.kb_per_split - 1) / p.kb_per_split;
  DSTAG_CHECK_ARG(g->split_stride == 0 || p.nsplit == g->split_k,
                  "dstag_gemm(tc): split_k = %d leaves empty K slices (%d k-blocks); the caller of a two-stage "
                  "reduction must pick a split count that divides the k-blocks evenly", g->split_k, p.kblocks);
  const long long items = (long long)g->batch * p.ntm * p.ntn * p.nsplit;
  DSTAG_CHECK_ARG(items < (1ll << 31), "dstag_gemm(tc): too many tiles");
  p.total_items = (int)items;
  {
    static int trace = -1;
    if (trace < 0) { const char* e = getenv("DSTAG_TC_TRACE"); trace = (e && e[0] == '1') ? 1 : 0; }
    if (trace) {
      static const char* mn[] = {"KVEC", "MVEC", "ROWG", "KSC"};
      fprintf(stderr, "[dstag_gemm tc] M=%d N=%d K=%d batch=%d bn=%d A=%s%s%s B=%s%s split=%d acc=%d\n", g->M, g->N, g->K,
              g->batch, bn, mn[p.a_mode], p.a_tma ? (p.a_rag_d1 ? "+TMA(ragged)" : "+TMA") : "",
              am ? "+mul" : "", mn[p.b_mode], p.b_tma ? "+TMA" : "",
              p.nsplit, g->accumulate);
    }
  }
  // A through tensor memory pays when the A tile is K-major (a row is four 16-byte loads); an MN-major
  // tile would need 16 scalar loads per thread and stays on the shared-memory operand path
  static int atm_on = -1;
  if (atm_on < 0) { const char* e = getenv("DSTAG_TC_ATM"); atm_on = e ? atoi(e) : 1; }
  const bool atm = (atm_on && x3 && !am && (p.a_mode != MODE_MVEC || atm_on == 2)) || (p.a_mode == MODE_MVEC && p.a_tma);
#define DSTAG_TC_DISPATCH(BN_)                                                                     \
  return x3 ? (am ? launch_tc<BN_, true, true, false>(p, maps, stream)                                            \
                  : (atm ? launch_tc<BN_, true, false, true>(p, maps, stream)                                    \
                         : launch_tc<BN_, true, false, false>(p, maps, stream)))                                 \
            : (am ? launch_tc<BN_, false, true, false>(p, maps, stream) : launch_tc<BN_, false, false, false>(p, maps, stream))
  if (bn == 128) DSTAG_TC_DISPATCH(128);
  if (bn == 64) DSTAG_TC_DISPATCH(64);
  DSTAG_TC_DISPATCH(32);
#undef DSTAG_TC_DISPATCH
}

}  // namespace dstag
