// tcgen05 tensor-core contraction engine for sm_100a.
//
//   C[b,m,n] (+)= alpha * sum_k (A . Amul)[b,m,k] * B[b,k,n]   (+Cadd) (*Cmul) (relu)
//
// Persistent, warp-specialised kernel (one CTA per SM, 13 warps):
//   warps 0-3   A producers   global (any 2-level strided view) -> registers -> [x Amul] ->
//                             [hi/lo TF32 split] -> 128B-swizzled K-major shared-memory tile
//   warps 4-7   B producers   same for the second operand
//   warps 8-11  epilogue      TMEM -> registers (tcgen05.ld) -> alpha / Cadd / Cmul / relu ->
//                             global (store, add or atomic), accumulator double-buffered in TMEM
//   warp  12    MMA issuer    one thread issues tcgen05.mma.kind::tf32 (128 x BN x 8) from the
//                             shared-memory descriptors; tcgen05.commit releases the stages
//
// The operands never need a packed / transposed copy in HBM: the producers read the caller's
// strides (permuted views, unfolded conv windows, padded leading dimensions) and the Hadamard
// factor (T_k of cheb_conv_withSAt, model/DSTAGNN_my.py:128) is fused into the operand load.
//
// Precision: kind::tf32 keeps 10 mantissa bits of each operand.  DSTAG_PREC_TF32X3 restores
// fp32-grade accuracy with the classic 3-pass split  a = a_hi + a_lo,  b = b_hi + b_lo,
// a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi  (error ~2^-21 relative), all accumulated in the
// same fp32 TMEM accumulator.  DSTAG_PREC_TF32 issues the single a_hi*b_hi pass.
#include "common.cuh"

namespace dstag {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                  // 32 x 4 B = one 128-byte swizzle row
constexpr int TC_A_WARPS = 4, TC_B_WARPS = 4, TC_EPI_WARPS = 4;   // epilogue warps 8-11: warp % 4 = TMEM lane quarter
constexpr int TC_MMA_WARP = TC_A_WARPS + TC_B_WARPS + TC_EPI_WARPS;
constexpr int TC_THREADS = (TC_MMA_WARP + 1) * 32;   // 416
constexpr int TC_SMEM_BUDGET = 196 * 1024;

struct TcParams {
  dstag_gemm_t g;
  int ntm, ntn, kblocks, kb_per_split, nsplit, total_items;
  int a_mode, b_mode;      // MODE_KVEC / MODE_MVEC / MODE_ROW (see Producer)
};

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(addr), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, 128 x N x 8, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t swz(int r, int c) {   // byte offset of 16B chunk c of row r
  return (uint32_t)(((r >> 3) << 10) + ((r & 7) << 7) + ((c ^ (r & 7)) << 4));
}

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

template <bool X3>
__device__ __forceinline__ void store_chunk(uint8_t* plane_hi, uint8_t* plane_lo, uint32_t off, float4 v) {
  if (X3) {
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    *reinterpret_cast<float4*>(plane_hi + off) = h;
    *reinterpret_cast<float4*>(plane_lo + off) = l;
  } else {
    *reinterpret_cast<float4*>(plane_hi + off) = v;
  }
}

// ---------------------------------------------------------------------------------------
// Operand producer.  NT threads fill one ROWS x 32 operand tile (hi [+ lo] plane) per k-block,
// one k-block ahead in registers (the next block's global loads are in flight while the
// current block is split / swizzled into shared memory).  Three load modes, chosen per launch:
//   KVEC  K contiguous   : float4 along K  -> K-major  swizzled tile
//   MVEC  row contiguous : float4 along M/N -> MN-major swizzled tile (no transpose needed:
//                          tcgen05 kind::tf32 accepts MN-major operands)
//   ROW   anything else  : scalar gather   -> K-major  swizzled tile
// ---------------------------------------------------------------------------------------
enum { MODE_KVEC = 0, MODE_MVEC = 1, MODE_ROW = 2 };

struct KWalk {   // offset of consecutive k indices along a (possibly two-level) axis, no division per step
  int off, lo, s1, wrap, d1;
  __device__ __forceinline__ void init(const dstag_axis_t& a, int k) {
    d1 = a.d1;
    if (d1 == 1) {
      off = k * (int)a.s0; s1 = (int)a.s0; wrap = 0; lo = 0;
    } else {
      const int hi = k / d1;
      lo = k - hi * d1;
      off = hi * (int)a.s0 + lo * (int)a.s1;
      s1 = (int)a.s1;
      wrap = (int)a.s0 - (d1 - 1) * (int)a.s1;
    }
  }
  __device__ __forceinline__ void next() {
    if (d1 == 1) {
      off += s1;
    } else {
      ++lo;
      const bool w = lo == d1;
      off += w ? wrap : s1;
      lo = w ? 0 : lo;
    }
  }
};

__device__ __forceinline__ float4 mul4(float4 a, float4 b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}

template <int ROWS, int NT, bool X3>
struct Producer {
  static constexpr int NV = ROWS * 8 / NT;     // float4 per thread per k-block
  static constexpr int RB = NT / 8;            // KVEC: rows per pass
  static constexpr int QPK = ROWS / 4;         // MVEC: row quads per k
  static constexpr int KGM = NT / QPK;         // MVEC: k groups
  static constexpr int KPTM = TC_BK / KGM;     // MVEC: k's per thread (== NV)
  static constexpr int KGR = NT / ROWS;        // ROW : k groups
  static constexpr int KPTR = TC_BK / KGR;     // ROW : scalars per thread (== 4 NV)
  static_assert(KPTM == NV && KPTR == 4 * NV && ROWS / RB == NV, "producer tiling mismatch");

  const float* base;
  const float* mbase;
  const dstag_axis_t* colax;
  const dstag_axis_t* mcolax;
  int K, t, mode;
  int ro[NV], mo[NV];      // KVEC: one per row;  MVEC / ROW: only [0] used
  int nvalid[NV];          // KVEC: row valid (0/1);  MVEC [0]: valid rows of the quad (0..4);  ROW [0]: 0/1

  __device__ __forceinline__ void begin_item(const float* b, const float* mb, const dstag_axis_t& rowax,
                                             const dstag_axis_t& mrowax, int row0, int limit) {
    base = b;
    mbase = mb;
    if (mode == MODE_KVEC) {
      const int rb = t >> 3;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int r = row0 + rb + RB * i;
        nvalid[i] = r < limit ? 1 : 0;
        ro[i] = nvalid[i] ? axoff32(rowax, r) : 0;
        mo[i] = (mb && nvalid[i]) ? axoff32(mrowax, r) : 0;
      }
    } else if (mode == MODE_MVEC) {
      const int r = row0 + 4 * (t % QPK);
      const int nv = limit - r;
      nvalid[0] = nv < 0 ? 0 : (nv > 4 ? 4 : nv);
      ro[0] = nvalid[0] ? axoff32(rowax, r) : 0;
      mo[0] = (mb && nvalid[0]) ? axoff32(mrowax, r) : 0;
    } else {
      const int r = row0 + (t % ROWS);
      nvalid[0] = r < limit ? 1 : 0;
      ro[0] = nvalid[0] ? axoff32(rowax, r) : 0;
      mo[0] = (mb && nvalid[0]) ? axoff32(mrowax, r) : 0;
    }
  }

  // issue the global loads of k-block kb into v[] (values are consumed one k-block later)
  __device__ __forceinline__ void load(int kb, float4 (&v)[NV]) {
    const int k0 = kb * TC_BK;
    if (mode == MODE_KVEC) {
      const int k = k0 + 4 * (t & 7);
      if (k + 3 < K) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (nvalid[i]) v[i] = __ldg(reinterpret_cast<const float4*>(base + ro[i] + k));
        }
        if (mbase) {
#pragma unroll
          for (int i = 0; i < NV; ++i)
            if (nvalid[i]) v[i] = mul4(v[i], __ldg(reinterpret_cast<const float4*>(mbase + mo[i] + k)));
        }
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          float e[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            e[j] = 0.f;
            if (nvalid[i] && k + j < K) {
              e[j] = __ldg(base + ro[i] + k + j);
              if (mbase) e[j] *= __ldg(mbase + mo[i] + k + j);
            }
          }
          v[i] = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
    } else if (mode == MODE_MVEC) {
      const int kf = k0 + (t / QPK) * KPTM;
      KWalk w, mw;
      w.init(*colax, kf);
      if (mbase) mw.init(*mcolax, kf);
      const int nv = nvalid[0];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kf + j < K) {
          const float* ptr = base + ro[0] + w.off;
          if (nv == 4) {
            v[j] = __ldg(reinterpret_cast<const float4*>(ptr));
          } else {
            if (nv > 0) v[j].x = __ldg(ptr);
            if (nv > 1) v[j].y = __ldg(ptr + 1);
            if (nv > 2) v[j].z = __ldg(ptr + 2);
          }
        }
        w.next();
      }
      if (mbase) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          if (kf + j < K) {
            const float* ptr = mbase + mo[0] + mw.off;
            if (nv == 4) {
              v[j] = mul4(v[j], __ldg(reinterpret_cast<const float4*>(ptr)));
            } else {
              if (nv > 0) v[j].x *= __ldg(ptr);
              if (nv > 1) v[j].y *= __ldg(ptr + 1);
              if (nv > 2) v[j].z *= __ldg(ptr + 2);
            }
          }
          mw.next();
        }
      }
    } else {
      const int kf = k0 + (t / ROWS) * KPTR;
      KWalk w, mw;
      w.init(*colax, kf);
      if (mbase) mw.init(*mcolax, kf);
      const bool rok = nvalid[0] != 0;
      const float* rp = base + ro[0];
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rok && kf + 4 * c + 0 < K) q.x = __ldg(rp + w.off);
        w.next();
        if (rok && kf + 4 * c + 1 < K) q.y = __ldg(rp + w.off);
        w.next();
        if (rok && kf + 4 * c + 2 < K) q.z = __ldg(rp + w.off);
        w.next();
        if (rok && kf + 4 * c + 3 < K) q.w = __ldg(rp + w.off);
        w.next();
        v[c] = q;
      }
      if (mbase) {
        const float* mp = mbase + mo[0];
#pragma unroll
        for (int c = 0; c < NV; ++c) {
          if (rok && kf + 4 * c + 0 < K) v[c].x *= __ldg(mp + mw.off);
          mw.next();
          if (rok && kf + 4 * c + 1 < K) v[c].y *= __ldg(mp + mw.off);
          mw.next();
          if (rok && kf + 4 * c + 2 < K) v[c].z *= __ldg(mp + mw.off);
          mw.next();
          if (rok && kf + 4 * c + 3 < K) v[c].w *= __ldg(mp + mw.off);
          mw.next();
        }
      }
    }
  }

  // split / swizzle the staged registers into the shared-memory operand tile
  __device__ __forceinline__ void store(const float4 (&v)[NV], uint8_t* plane_hi, uint8_t* plane_lo) {
    if (mode == MODE_KVEC) {
      const int c = t & 7, rb = t >> 3;
#pragma unroll
      for (int i = 0; i < NV; ++i) store_chunk<X3>(plane_hi, plane_lo, swz(rb + RB * i, c), v[i]);
    } else if (mode == MODE_MVEC) {
      // tf32 MN-major uses the SWIZZLE_128B_BASE32B layout: atoms of 32 rows (one 128 B line) x 4 k
      // (512 B); inside an atom the 32-byte units of line k are XOR-ed with k % 4.  Atom (ka, mb)
      // lives at (ka * ROWS/32 + mb) * 512.
      const int mq = t % QPK, kl0 = (t / QPK) * KPTM;
      const int mb = mq >> 3, mc = mq & 7;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int kl = kl0 + j, ka = kl >> 2, kr = kl & 3;
        const uint32_t off = (uint32_t)(((ka * (ROWS / 32) + mb) << 9) + (kr << 7) + (((mc >> 1) ^ kr) << 5) +
                                        ((mc & 1) << 4));
        store_chunk<X3>(plane_hi, plane_lo, off, v[j]);
      }
    } else {
      const int r = t % ROWS, c0 = (t / ROWS) * NV;
#pragma unroll
      for (int c = 0; c < NV; ++c) store_chunk<X3>(plane_hi, plane_lo, swz(r, c0 + c), v[c]);
    }
  }
};

template <int BN, bool X3>
struct TcCfg {
  static constexpr int A_PLANE = TC_BM * 128;
  static constexpr int B_PLANE = BN * 128;
  static constexpr int PLANES = X3 ? 2 : 1;
  static constexpr int STAGE = PLANES * (A_PLANE + B_PLANE);
  static constexpr int STAGES_RAW = TC_SMEM_BUDGET / STAGE;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM = STAGES * STAGE + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
};

// shared-memory matrix descriptors (sm_100 "version 1", SWIZZLE_128B)
//   K-major : rows 128 B apart, 8-row groups SBO = 1024 B apart; K step of 8 tf32 = +32 B
//   MN-major (tf32 => SWIZZLE_128B_BASE32B): 32-row x 4-k atoms of 512 B; LBO = 512 B between
//             row blocks, SBO between k atoms; one K = 8 MMA step spans two atoms
__device__ __forceinline__ uint64_t make_desc_k(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (32ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
         (1ull << 61);
}

template <int ROWS, int NT, bool X3, int STAGES, int STAGE_BYTES>
__device__ __forceinline__ void produce_item(Producer<ROWS, NT, X3>& pr, int kb0, int kb1, uint32_t& it,
                                             uint8_t* smem, uint32_t plane_off, uint32_t plane_bytes,
                                             uint64_t* full_bar, uint64_t* empty_bar, int lane) {
  constexpr int NV = Producer<ROWS, NT, X3>::NV;
  float4 cur[NV], nxt[NV];
  pr.load(kb0, cur);
  for (int kb = kb0; kb < kb1; ++kb, ++it) {
    if (kb + 1 < kb1) pr.load(kb + 1, nxt);          // next block's loads fly while this one is stored
    const int s = it % STAGES;
    const uint32_t ph = (it / STAGES) & 1;
    if (lane == 0) mbar_wait(empty_bar + s, ph ^ 1);
    __syncwarp();
    uint8_t* st = smem + s * STAGE_BYTES + plane_off;
    pr.store(cur, st, st + plane_bytes);
    fence_proxy_async_smem();      // generic-proxy stores -> visible to the tensor core (async proxy)
    __syncwarp();
    if (lane == 0) mbar_arrive(full_bar + s);
#pragma unroll
    for (int i = 0; i < NV; ++i) cur[i] = nxt[i];
  }
}

template <int BN, bool X3>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TcParams p) {
  using Cfg = TcCfg<BN, X3>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE);
  uint64_t* full_bar = bars;                          // [STAGES] producers -> MMA
  uint64_t* empty_bar = bars + Cfg::STAGES;           // [STAGES] MMA (commit) -> producers
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;       // [2] MMA (commit) -> epilogue
  uint64_t* tempty_bar = bars + 2 * Cfg::STAGES + 2;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const dstag_gemm_t& g = p.g;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar + s, TC_A_WARPS + TC_B_WARPS);
      mbar_init(empty_bar + s, 1);
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
    const bool is_a = warp < TC_A_WARPS;
    uint32_t it = 0;   // running k-block counter across items (stage ring position)
    Producer<TC_BM, TC_A_WARPS * 32, X3> pa;
    Producer<BN, TC_B_WARPS * 32, X3> pb;
    pa.t = threadIdx.x;
    pb.t = threadIdx.x - TC_A_WARPS * 32;
    pa.K = pb.K = g.K;
    pa.mode = p.a_mode;
    pb.mode = p.b_mode;
    pa.colax = &g.A.col;
    pa.mcolax = &g.Amul.col;
    pb.colax = &g.B.row;      // B operand: tile rows = n (B.col axis), tile columns = k (B.row axis)
    pb.mcolax = &g.B.row;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x) {
      int q = w;
      const int ks = q % p.nsplit; q /= p.nsplit;
      const int tn = q % p.ntn; q /= p.ntn;
      const int tm = q % p.ntm;
      const int batch = q / p.ntm;
      const int kb0 = ks * p.kb_per_split;
      const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
      if (is_a) {
        const float* ab = (const float*)g.A.ptr + axoff64(g.A.batch, batch);
        const float* mb = g.Amul.ptr ? (const float*)g.Amul.ptr + axoff64(g.Amul.batch, batch) : nullptr;
        pa.begin_item(ab, mb, g.A.row, g.Amul.row, tm * TC_BM, g.M);
        produce_item<TC_BM, TC_A_WARPS * 32, X3, Cfg::STAGES, Cfg::STAGE>(pa, kb0, kb1, it, smem, 0, Cfg::A_PLANE,
                                                                          full_bar, empty_bar, lane);
      } else {
        const float* bb = (const float*)g.B.ptr + axoff64(g.B.batch, batch);
        pb.begin_item(bb, nullptr, g.B.col, g.B.col, tn * BN, g.N);
        produce_item<BN, TC_B_WARPS * 32, X3, Cfg::STAGES, Cfg::STAGE>(pb, kb0, kb1, it, smem,
                                                                       Cfg::PLANES * Cfg::A_PLANE, Cfg::B_PLANE,
                                                                       full_bar, empty_bar, lane);
      }
    }
  } else if (warp == TC_MMA_WARP) {
    // =============================== MMA issuer ========================================
    if (lane == 0) {
      const bool a_mn = p.a_mode == MODE_MVEC, b_mn = p.b_mode == MODE_MVEC;
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      // descriptor advance per K step of 8 (in 16-byte units): +32 B inside the swizzle row (K-major),
      // one 8-k atom row (MN-major)
      const uint64_t a_adv = a_mn ? (uint64_t)((TC_BM / 32) * 1024 >> 4) : 2ull;   // 2 atoms of SBO = ROWS/32 * 512 B
      const uint64_t b_adv = b_mn ? (uint64_t)((BN / 32) * 1024 >> 4) : 2ull;
      uint32_t it = 0, item = 0;
      for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item) {
        const int ks = w % p.nsplit;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
        const uint32_t acc = item & 1;
        mbar_wait(tempty_bar + acc, ((item >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % Cfg::STAGES;
          const uint32_t ph = (it / Cfg::STAGES) & 1;
          mbar_wait(full_bar + s, ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::STAGE);
          const uint32_t sb = sa + Cfg::PLANES * Cfg::A_PLANE;
          const uint64_t a_hi = a_mn ? make_desc_mn(sa, (TC_BM / 32) * 512) : make_desc_k(sa);
          const uint64_t a_lo = a_mn ? make_desc_mn(sa + Cfg::A_PLANE, (TC_BM / 32) * 512) : make_desc_k(sa + Cfg::A_PLANE);
          const uint64_t b_hi = b_mn ? make_desc_mn(sb, (BN / 32) * 512) : make_desc_k(sb);
          const uint64_t b_lo = b_mn ? make_desc_mn(sb + Cfg::B_PLANE, (BN / 32) * 512) : make_desc_k(sb + Cfg::B_PLANE);
#pragma unroll
          for (int kk = 0; kk < TC_BK / 8; ++kk) {
            const uint64_t aa = a_adv * kk, ba = b_adv * kk;
            const uint32_t first = (kb == kb0 && kk == 0) ? 0u : 1u;
            if (X3) {
              umma_tf32(d_tmem, a_lo + aa, b_hi + ba, idesc, first);
              umma_tf32(d_tmem, a_hi + aa, b_lo + ba, idesc, 1u);
              umma_tf32(d_tmem, a_hi + aa, b_hi + ba, idesc, 1u);
            } else {
              umma_tf32(d_tmem, a_hi + aa, b_hi + ba, idesc, first);
            }
          }
          umma_commit(empty_bar + s);          // stage reusable once these MMAs have read it
        }
        umma_commit(tfull_bar + acc);          // accumulator complete -> epilogue
      }
    }
  } else {
    // =============================== epilogue ==========================================
    const int q4 = warp & 3;                   // TMEM lane quarter this warp may access
    uint32_t item = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item) {
      int q = w;
      const int ks = q % p.nsplit; q /= p.nsplit;
      const int tn = q % p.ntn; q /= p.ntn;
      const int tm = q % p.ntm;
      const int batch = q / p.ntm;
      const uint32_t acc = item & 1;
      const int m = tm * TC_BM + q4 * 32 + lane;
      const bool mok = m < g.M;
      float* C = (float*)g.C.ptr + axoff64(g.C.batch, batch) + (mok ? axoff64(g.C.row, m) : 0);
      const float* Ca = (g.Cadd.ptr && ks == 0)
                            ? (const float*)g.Cadd.ptr + axoff64(g.Cadd.batch, batch) + (mok ? axoff64(g.Cadd.row, m) : 0)
                            : nullptr;
      const float* Cm = g.Cmul.ptr
                            ? (const float*)g.Cmul.ptr + axoff64(g.Cmul.batch, batch) + (mok ? axoff64(g.Cmul.row, m) : 0)
                            : nullptr;
      const bool c_unit = g.C.col.d1 == 1 && g.C.col.s0 == 1;
      const bool ca_unit = !Ca || (g.Cadd.col.d1 == 1 && g.Cadd.col.s0 == 1);
      const bool cm_unit = !Cm || (g.Cmul.col.d1 == 1 && g.Cmul.col.s0 == 1);
      if (lane == 0) mbar_wait(tfull_bar + acc, (item >> 1) & 1);
      __syncwarp();                 // tcgen05.ld is .sync.aligned: the warp must be converged
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        float v[32];
        tmem_ld32(tmem_base + acc * BN + cc * 32 + ((uint32_t)(q4 * 32) << 16), v);
        const int n0 = tn * BN + cc * 32;
        if (!mok || n0 >= g.N) continue;
        const bool vec = c_unit && ca_unit && cm_unit && g.accumulate != DSTAG_ACC_ATOMIC && (n0 + 32 <= g.N) &&
                         ((reinterpret_cast<uintptr_t>(C + n0) & 15) == 0) &&
                         (!Ca || (reinterpret_cast<uintptr_t>(Ca + n0) & 15) == 0) &&
                         (!Cm || (reinterpret_cast<uintptr_t>(Cm + n0) & 15) == 0);
        if (vec) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(g.alpha * v[j], g.alpha * v[j + 1], g.alpha * v[j + 2], g.alpha * v[j + 3]);
            if (Ca) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(Ca + n0 + j));
              o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
            }
            if (Cm) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(Cm + n0 + j));
              o.x *= a.x; o.y *= a.y; o.z *= a.z; o.w *= a.w;
            }
            if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            float4* dst = reinterpret_cast<float4*>(C + n0 + j);
            if (g.accumulate == DSTAG_ACC_ADD) {
              const float4 old = *dst;
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *dst = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + j;
            if (n < g.N) {
              float o = g.alpha * v[j];
              if (Ca) o += __ldg(Ca + axoff64(g.Cadd.col, n));
              if (Cm) o *= __ldg(Cm + axoff64(g.Cmul.col, n));
              if (g.relu) o = fmaxf(o, 0.f);
              float* dst = C + axoff64(g.C.col, n);
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

static bool operand_fits32(const dstag_operand_t& o, long long nrow, long long ncol) {
  if (!o.ptr) return true;
  if (!axis_nonneg(o.batch) || !axis_nonneg(o.row) || !axis_nonneg(o.col)) return false;
  return axis_extent(o.row, nrow) + axis_extent(o.col, ncol) < (1ll << 31) - 64;
}

bool gemm_tc_supported(const dstag_gemm_t* g) {
  if (g->N < 8 || g->K < 8 || g->M < 32) return false;
  if ((long long)g->M * g->N * g->K * g->batch < (1ll << 19)) return false;   // tiny: not worth a persistent launch
  return operand_fits32(g->A, g->M, g->K) && operand_fits32(g->Amul, g->M, g->K) &&
         operand_fits32(g->B, g->K, g->N) && operand_fits32(g->C, g->M, g->N) &&
         operand_fits32(g->Cadd, g->M, g->N) && operand_fits32(g->Cmul, g->M, g->N);
}

// K contiguous and every other stride a multiple of 4 elements on a 16-byte aligned base
static bool kvec_ok(const dstag_operand_t& o, const dstag_axis_t& kax, const dstag_axis_t& rowax) {
  if (!o.ptr) return true;
  return kax.d1 == 1 && kax.s0 == 1 && axis_mult4(rowax) && axis_mult4(o.batch) &&
         (reinterpret_cast<uintptr_t>(o.ptr) & 15) == 0;
}

// rows contiguous (single-level, stride 1), every K / batch stride a multiple of 4, aligned base
static bool mvec_ok(const dstag_operand_t& o, const dstag_axis_t& kax, const dstag_axis_t& rowax) {
  if (!o.ptr) return true;
  return rowax.d1 == 1 && rowax.s0 == 1 && axis_mult4(kax) && axis_mult4(o.batch) &&
         (reinterpret_cast<uintptr_t>(o.ptr) & 15) == 0;
}

static int g_num_sms = 0;

template <int BN, bool X3>
static int launch_tc(const TcParams& p, cudaStream_t stream) {
  using Cfg = TcCfg<BN, X3>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) {
      set_error("dstag_gemm(tc): cudaFuncSetAttribute(%d) failed: %s", Cfg::SMEM, cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  const int grid = p.total_items < g_num_sms ? p.total_items : g_num_sms;
  gemm_tc_kernel<BN, X3><<<grid, TC_THREADS, Cfg::SMEM, stream>>>(p);
  DSTAG_RETURN_LAUNCH("dstag_gemm(tcgen05)");
}

int gemm_tc(const dstag_gemm_t* g, cudaStream_t stream) {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = kNumSMs;
  }
  TcParams p;
  p.g = *g;
  const int bn = g->N > 64 ? 128 : (g->N > 32 ? 64 : 32);
  p.ntm = (g->M + TC_BM - 1) / TC_BM;
  p.ntn = (g->N + bn - 1) / bn;
  p.kblocks = (g->K + TC_BK - 1) / TC_BK;
  int split = g->split_k < 1 ? 1 : g->split_k;
  if (split > p.kblocks) split = p.kblocks;
  p.kb_per_split = (p.kblocks + split - 1) / split;
  p.nsplit = (p.kblocks + p.kb_per_split - 1) / p.kb_per_split;
  const long long items = (long long)g->batch * p.ntm * p.ntn * p.nsplit;
  DSTAG_CHECK_ARG(items < (1ll << 31), "dstag_gemm(tc): too many tiles");
  p.total_items = (int)items;
  if (kvec_ok(g->A, g->A.col, g->A.row) && kvec_ok(g->Amul, g->Amul.col, g->Amul.row)) p.a_mode = MODE_KVEC;
  else if (mvec_ok(g->A, g->A.col, g->A.row) && mvec_ok(g->Amul, g->Amul.col, g->Amul.row)) p.a_mode = MODE_MVEC;
  else p.a_mode = MODE_ROW;
  if (kvec_ok(g->B, g->B.row, g->B.col)) p.b_mode = MODE_KVEC;
  else if (mvec_ok(g->B, g->B.row, g->B.col)) p.b_mode = MODE_MVEC;
  else p.b_mode = MODE_ROW;
  const bool x3 = g->precision == DSTAG_PREC_TF32X3;
  if (bn == 128) return x3 ? launch_tc<128, true>(p, stream) : launch_tc<128, false>(p, stream);
  if (bn == 64) return x3 ? launch_tc<64, true>(p, stream) : launch_tc<64, false>(p, stream);
  return x3 ? launch_tc<32, true>(p, stream) : launch_tc<32, false>(p, stream);
}

}  // namespace dstag
