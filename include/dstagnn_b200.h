/*
 * dstagnn_b200.h -- C ABI of libdstagnn_b200.so (hand-written sm_100a kernels for the
 * DSTAGNN block forward/backward path).
 *
 * The reference (Ghoul-tn/DSTAGNN_Drought) has no FFI: every operation of the hot path
 * is a stock PyTorch call inside model/DSTAGNN_my.py.  Each entry point below therefore
 * cites the reference lines whose arithmetic it replaces.  Conventions:
 *   - all tensors are fp32 device pointers owned by the caller, borrowed for the
 *     duration of the enqueue; nothing is allocated inside the library;
 *   - `stream` is a cudaStream_t (CUstream) passed as void*; work is stream-ordered and
 *     the calls are re-entrant per stream (CUDA-graph capturable);
 *   - return value 0 = enqueued, negative = DSTAG_ERR_* (no kernel launched),
 *     positive = cudaError_t of the failed launch.
 *   - there is no CPU path: every function requires device pointers.
 */
#ifndef DSTAGNN_B200_H
#define DSTAGNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSTAG_ERR_BAD_ARG (-1)
#define DSTAG_ERR_UNSUPPORTED (-2)
#define DSTAG_ERR_SMEM (-3)

/* ABI version, bumped whenever a signature below changes. */
int dstag_abi_version(void);
/* Last error text of the calling thread ("" if none). */
const char* dstag_last_error(void);
/* Number of kernels this library has launched since load / since the last reset
 * (process-wide counter; used by bench.py for "gpu_launches"). */
int64_t dstag_launch_count(void);
void dstag_reset_launch_count(void);

/* ------------------------------------------------------------------------------------
 * Strided tensor contraction  C[b,m,n] (+)= alpha * sum_k (A*Amul)[b,m,k] * B[b,k,n]
 *                             then  (+ Cadd) (* Cmul) (relu)
 *
 * Each logical index (batch, m, n, k) may be the flattening of up to TWO tensor
 * dimensions: idx -> (idx / d1) * s0 + (idx % d1) * s1  (element strides; d1 == 1 means a
 * single dimension).  This one primitive covers every dense product of the block:
 *   TAt W_Q/W_K/W_V/fc       model/DSTAGNN_my.py:92-99     (x permuted view, no copy)
 *   pre_conv, final_conv     model/DSTAGNN_my.py:232,278   (Conv2d == per-node Linear)
 *   SAt W_Q/W_K, Q K^T       model/DSTAGNN_my.py:62-63,19
 *   cheb: (T_k . A)^T x      model/DSTAGNN_my.py:128-131   (Amul = T_k fused into operand load)
 *   Theta_k mix, fcmy        model/DSTAGNN_my.py:131,243
 *   GTU Conv2d(1,k)          model/DSTAGNN_my.py:193       (im2col folded into the strides)
 * and, by swapping operand roles, all of their input/weight gradients.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int64_t s0;   /* stride of the outer sub-index (elements) */
  int64_t s1;   /* stride of the inner sub-index (elements); ignored when d1 == 1 */
  int32_t d1;   /* extent of the inner sub-index; 1 = not split */
  int32_t pad_;
} dstag_axis_t;

typedef struct {
  const void* ptr;       /* fp32 device pointer; NULL = operand absent (Amul/Cmul/Cadd) */
  dstag_axis_t batch;
  dstag_axis_t row;      /* A: m   B: k   C/Cmul/Cadd: m   Amul: m */
  dstag_axis_t col;      /* A: k   B: n   C/Cmul/Cadd: n   Amul: k */
} dstag_operand_t;

enum { DSTAG_ACC_STORE = 0, DSTAG_ACC_ADD = 1, DSTAG_ACC_ATOMIC = 2 };
enum {
  DSTAG_PREC_FP32 = 0,     /* fp32 FMA on CUDA cores (exact fp32 semantics) */
  DSTAG_PREC_TF32X3 = 1,   /* tcgen05 kind::tf32, 3-pass split: fp32-grade accuracy */
  DSTAG_PREC_TF32 = 2      /* tcgen05 kind::tf32 single pass (reported separately) */
};

typedef struct {
  int32_t M, N, K, batch;
  dstag_operand_t A, B, C, Amul, Cmul, Cadd;
  float alpha;
  int32_t relu;          /* apply max(.,0) last */
  int32_t accumulate;    /* DSTAG_ACC_*; with split_k > 1 (or C aliasing over batch) either ATOMIC, or STORE
                            together with split_stride != 0 */
  int32_t split_k;       /* >= 1 */
  int32_t precision;     /* DSTAG_PREC_* */
  int32_t chain;         /* 0, or the longest accumulation chain (in k) the tensor-core engine may run in its
                            fp32 accumulator before folding the partial result into registers: the accumulator
                            truncates at every k-step, a bias that matters for products whose result is far
                            smaller than its terms.  Honoured for N <= 32 (tile of 32 columns). */
  int64_t split_stride;  /* 0, or (accumulate == STORE): K-split s writes its partial product to
                            C + s * split_stride (elements) instead of adding atomically; the caller sums the
                            split_k slices in a fixed order with dstag_sum_slices -- a reproducible reduction */
} dstag_gemm_t;

int dstag_gemm(const dstag_gemm_t* g, void* stream);
/* 1 if this shape/precision is served by the tcgen05 engine, 0 if by the SIMT kernel. */
int dstag_gemm_uses_tensor_cores(const dstag_gemm_t* g);

/* ------------------------------------------------------------------------------------
 * Row LayerNorm with fused pre-add:  y[r,:] = LN(a[r,:] + tab[r % tab_mod,:]) * gamma + beta
 * followed by optional inverted dropout (Philox, keyed by seed; p == 0 disables).  seed_dev (nullable, all
 * dropout entry points): a device-resident step counter folded into the key, so that a step captured in a CUDA
 * graph draws fresh masks on every replay (dstag_tick advances it inside the graph).
 *   a is gathered as a[(r / a_div) * a_s0 + (r % a_div) * a_s1 + c * a_sc].
 * Replaces Embedding.forward (model/DSTAGNN_my.py:172-182: pos_embed + LayerNorm, both
 * 'T' and 'S' types, tab = the embedding table), the residual LayerNorm of
 * MultiHeadAttention (line 100: tab = residual, tab_mod = rows) and the dropout of
 * line 234.  mean/rstd (rows each) are saved for the backward.  tab rows and y rows have leading
 * dimensions tab_ld / y_ld >= L (rows over the node axis are padded to 16-byte multiples so the
 * contractions that consume them take the vector / TMA load paths).
 * ---------------------------------------------------------------------------------- */
int dstag_layernorm_fwd(const float* a, int64_t a_s0, int64_t a_s1, int32_t a_div, int64_t a_sc,
                        const float* tab, int32_t tab_mod, int64_t tab_ld, const float* gamma, const float* beta,
                        float eps, float* y, int64_t y_ld, float* mean, float* rstd, int32_t rows, int32_t L,
                        float drop_p, uint64_t seed, const uint64_t* seed_dev, void* stream);
/* dpre[r,:] = d(loss)/d(a + tab) (contiguous rows x L); dgamma/dbeta are ACCUMULATED
 * (atomicAdd) into caller-zeroed buffers. */
int dstag_layernorm_bwd(const float* dy, int64_t dy_ld, const float* a, int64_t a_s0, int64_t a_s1, int32_t a_div,
                        int64_t a_sc, const float* tab, int32_t tab_mod, int64_t tab_ld, const float* gamma,
                        const float* mean, const float* rstd, float* dpre, int64_t dpre_ld, float* dgamma,
                        float* dbeta, int32_t rows, int32_t L, float drop_p, uint64_t seed,
                        const uint64_t* seed_dev, void* stream);

/* out[r * out_ld + c] = sum_{s < nslices} ws[s * slice_stride + r * cols + c]  (s ascending: a fixed-order,
 * run-to-run reproducible reduction).  Second stage of the split-K / K-outer-fold contractions that must
 * not use atomics (every forward product: an activation within rounding of a ReLU kink would otherwise
 * land on either side of it from run to run). */
int dstag_sum_slices(const float* ws, int64_t slice_stride, int32_t nslices, float* out, int64_t rows, int64_t cols,
                     int64_t out_ld, void* stream);

/* out[c] (+)= sum_r in[r * ld + c], r < rows, c < cols.  Bias / embedding-table /
 * mask gradients (autograd of model/DSTAGNN_my.py:127,175,180,193,207,221).  `out` must be
 * zeroed by the caller (partial sums are combined with atomicAdd). */
int dstag_colsum(const float* in, float* out, int64_t rows, int64_t cols, int64_t ld, void* stream);

/* out[(b * C + pc(c)) * out_ld + pr(r)] = in[(b * R + r) * in_ld + c], p(i) = (i % d) * (n / d) + i / d
 * with d = r_d (n = R) for rows and d = c_d (n = C) for columns; d = 1 is the identity.
 * x.permute(0,2,3,1) of model/DSTAGNN_my.py:230 made contiguous: the block keeps its activations
 * channels-last (B, N, T, F), the temporal attention wants rows (b, f, t) -- c_d = F swaps the pair while
 * transposing (and r_d = T swaps it back in the backward). */
int dstag_transpose(const float* in, int64_t in_ld, float* out, int64_t out_ld, int32_t batch, int32_t R, int32_t C,
                    int32_t r_d, int32_t c_d, void* stream);

/* In-place row softmax over L entries of rows with leading dimension ld; entries
 * [L, ld) are set to 0.  F.softmax(..., dim=1) of model/DSTAGNN_my.py:127 on the
 * transposed map layout (row = output node j, column = i). */
int dstag_softmax_fwd(float* x, int64_t rows, int32_t L, int32_t ld, void* stream);
/* In place on dp: g = dp (.) mul ; dp <- p * (g - sum(p * g)).  mul (nullable): mul_rows rows of
 * leading dimension ld, row r uses mul[r % mul_rows] -- the transposed Chebyshev polynomial T_k that
 * follows the softmax in model/DSTAGNN_my.py:128, so the contraction that produces dp needs no
 * multiplier epilogue. */
int dstag_softmax_bwd(const float* p, float* dp, const float* mul, int64_t mul_rows, int64_t rows, int32_t L,
                      int32_t ld, void* stream);

/* Spatial attention maps, forward, fused: logits + bias + softmax in one kernel
 * (SScaledDotProductAttention, model/DSTAGNN_my.py:19, and lines 126-127 of cheb_conv_withSAt):
 *   P[b,k,j,i] = softmax_i( Qs[b,i,k,:] . Ks[b,j,k,:] / sqrt(dk) + biasT[k,j,i] ),  pad columns = 0
 *   qk (B, N, 2, K, dk) packed [W_Q | W_K] projections; biasT (K, N, ld); P (B, K, N, ld); ld % 4 == 0.
 * TP (nullable, same shape as P): second output TP = tkT (.) P with tkT (K, N, ld) the transposed
 * Chebyshev polynomials -- the masked operator of line 128, so that the Chebyshev contraction and its
 * input gradient read ONE operand stream instead of two.
 * Returns DSTAG_ERR_SMEM when a strip of 16 rows of N logits does not fit in shared memory (N > ~2900)
 * (then use dstag_gemm + dstag_softmax_fwd). */
int dstag_sat_softmax_fwd(const float* qk, const float* biasT, const float* tkT, float* P, float* TP, int32_t B,
                          int32_t N, int32_t K, int32_t dk, int32_t ld, void* stream);

/* dx = dy * (y > 0)  (backward of the ReLUs at model/DSTAGNN_my.py:133,245-247). */
int dstag_relu_bwd(const float* y, const float* dy, float* dx, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------
 * Temporal attention core (ScaledDotProductAttention, model/DSTAGNN_my.py:30-42), one
 * CTA per (b, f, h):  S = Q K^T / sqrt(dk) + res ; P = softmax over the QUERY index
 * (dim=3 of (B,F,H,Tq,Tk)) ; ctx = P V.
 *   qkv    (B,F,T,3,H,dk)   packed projections
 *   res    (B,res_F,H,T,T)  res_F in {0 (absent), 1 (broadcast over F), F}
 *   scores (B,F,H,T,T)      returned logits (line 42)
 *   ctx    (B,F,T,H,dk)     == context.transpose(2,3).reshape(...) of line 98
 * ---------------------------------------------------------------------------------- */
int dstag_tattn_fwd(const float* qkv, const float* res, int32_t res_F, float* scores, float* ctx,
                    int32_t B, int32_t F, int32_t H, int32_t T, int32_t dk, void* stream);
/* dscores_ext (nullable): gradient arriving through the returned logits (next block's
 * res_att).  dscores_out (B,F,H,T,T): total gradient w.r.t. the logits (= grad of res). */
int dstag_tattn_bwd(const float* qkv, const float* scores, const float* dctx,
                    const float* dscores_ext, float* dqkv, float* dscores_out, int32_t B, int32_t F,
                    int32_t H, int32_t T, int32_t dk, void* stream);

/* ------------------------------------------------------------------------------------
 * Block tail.  ALL tensors are channels-last (X = B*N, time, C): the graph-conv output g, the GTU conv
 * outputs u, the concatenated gate outputs cat (X, S = 3T-12, C), the residual r, the block output
 * and their gradients.  (The block hands its output back as a permuted (B, N, C, T) view.)
 *
 * GTU gate (model/DSTAGNN_my.py:194-196) + torch.cat (line 242):
 *   out[x, off + t, c] = tanh(u[x, t, c]) * sigmoid(u[x, t, C + c]),  u (X, Tp, 2C), out (X, S, C).
 * The temporal convolutions themselves run through dstag_gemm on an overlapping strided view
 * of the channels-last activation (K = ks*C contiguous), see ops.GtuStack.
 * ---------------------------------------------------------------------------------- */
int dstag_gtu_gate_fwd(const float* u, float* out, int64_t X, int32_t C, int32_t Tp, int32_t S, int32_t off,
                       void* stream);
/* dup (X, Tp + 2*pad, 2C): gradient w.r.t. u, written with `pad` zero rows before and after the
 * Tp valid time steps, so that the conv input gradient is a gather-only correlation over an
 * unfolded view of dup (no atomics).  db (2C, nullable, zeroed by the caller): the conv bias
 * gradient (column sums of dup) is ACCUMULATED into it. */
int dstag_gtu_gate_bwd(const float* u, const float* dout, float* dup, float* db, int64_t X, int32_t C,
                       int32_t Tp, int32_t pad, int32_t S, int32_t off, void* stream);
/* the same from the gate factors sv = [tanh(p) | sigmoid(q)] (X, Tp, 2C) that dstag_gtu_conv_gate_fwd saves: shapes whose
 * forward runs fused but whose backward does not (more than 128 output steps per node: the drought configuration) */
int dstag_gtu_gate_bwd_sv(const float* sv, const float* dout, float* dup, float* db, int64_t X, int32_t C,
                          int32_t Tp, int32_t pad, int32_t S, int32_t off, void* stream);

/* Fused GTU kernels on the tcgen05 tensor cores (csrc/gtu_tc.cu), C = 32 channels, ks in {3, 5, 7}:
 *   dstag_gtu_conv_gate_fwd: Conv2d(C -> 2C, (1, ks)) over time + bias + tanh * sigmoid gate + torch.cat
 *     (model/DSTAGNN_my.py:192-196, 238-242) in one kernel.  g (X, T, C) channels-last, w (2C, C, 1, ks) and
 *     bias (2C) as nn.Conv2d stores them; writes rows off .. off + T - ks of cat (X, S, C) and, when sv is not
 *     NULL, sv (X, T - ks + 1, 2C) = [tanh(p) | sigmoid(q)] for the backward.  The conv output never reaches HBM.
 *   dstag_gtu_gate_conv_bwd: autograd of the above w.r.t. g, plus the gradient of the conv output:
 *     du (X, T - ks + 1, 2C) = gate'(sv) * dcat[:, off .., :] (the operand of the weight gradient; its column
 *     sums -- partial sums per warp in db_ws, see below -- are the bias gradient) and dg (X, T, C) (+)= the transposed convolution of du with w (`accumulate`
 *     != 0 adds to dg: the three kernel sizes share one input).  Deterministic (no atomics).
 *   dstag_gtu_tc_supported(C, T, ks, backward) tells whether a shape is covered (T - ks + 1 <= 128 rows per
 *     node in the backward); other shapes take dstag_gemm + the gate kernels above. */
int dstag_gtu_tc_supported(int32_t C, int32_t T, int32_t ks, int32_t backward);
int dstag_gtu_conv_gate_fwd(const float* g, const float* w, const float* bias, float* cat, float* sv, int64_t X,
                            int32_t C, int32_t T, int32_t ks, int32_t S, int32_t off, void* stream);
int dstag_gtu_gate_conv_bwd(const float* sv, const float* dcat, const float* w, float* du, float* dg, float* db_ws,
                            int64_t X, int32_t C, int32_t T, int32_t ks, int32_t S, int32_t off, int32_t accumulate,
                            void* stream);
/* rows of the (rows, 2C) workspace db_ws (nullable; ZEROED by the caller): every producer warp of the kernel leaves
 * its partial column sums of du there; their column sum (dstag_colsum) is the conv bias gradient */
int dstag_gtu_bwd_ws_rows(void);

/* Tall-skinny Gram product on the tcgen05 tensor cores (csrc/gram_tc.cu) -- the weight-gradient shapes:
 *   D[(j, c), n] = sum_r A_j[r, c] * B[r, n],   j < taps <= 8, c < 32, n < NB in {32, 64}, r < X * Tp,
 *   A_j[r, :] = the 32 floats at  a + ((r / Tp) * T + r % Tp) * 32 + j * tap_stride,   B (X * Tp, NB) contiguous.
 * GTU conv weight gradient (autograd of model/DSTAGNN_my.py:193): a = g (X, T, 32), tap_stride = 32 (the next time
 * step), b = du (X, Tp, 64), D[(j, c), o] = dW[o, c, 0, j].
 * Every CTA (dstag_gram_slices() of them) writes ONE partial result: partial (slices, 128 * ceil(taps / 4), NB); the
 * caller adds them in a fixed order with dstag_sum_slices (rows (j, c) with j >= taps are zero).  Rows are taken in
 * k-blocks of floor(32 / Tp) whole nodes. */
int dstag_gram_slices(void);
int dstag_gram_supported(int32_t taps, int32_t NB, int32_t T, int32_t Tp);   /* Tp <= 32: a k-block is whole nodes */
int dstag_gram_rows(const float* a, const float* b, float* partial, int64_t X, int32_t T, int32_t Tp, int32_t taps,
                    int64_t tap_stride, int32_t NB, void* stream);

/* out[m, n] += sum_x sum_c a[x, m, c] * b[x, n, c];  asum[m] += sum_x sum_c a[x, m, c] (nullable).
 * a (X, m, C), b (X, n, C) channels-last, 16-byte aligned; out / asum zeroed by the caller.
 * The fcmy weight and bias gradients (autograd of model/DSTAGNN_my.py:243): a T x (3T-12) result
 * reduced over B*N*C -- a streaming reduction, not a tensor-core shape. */
int dstag_outer_sum_supported(int32_t m, int32_t n, int32_t C);
int dstag_outer_sum(const float* a, const float* b, float* out, float* asum, int64_t X, int32_t m, int32_t n,
                    int32_t C, void* stream);

/* Thin-feature contractions (csrc/thin.cu): J = taps * F <= 4 input values per row, C output channels (C / 4 a power
 * of two <= 32).  The first block's Theta mix (model/DSTAGNN_my.py:131 with in_channels = 1) and 1x1 residual_conv
 * (line 228) -- streaming kernels that move the (R, C) tensor once instead of one-k-block tensor-core tiles.
 *   a_j[r] = a[(j / F) * a_plane + r * F + j % F]  (taps planes of (R, F) rows);   w[j, c] = w[j * w_js + c * w_cs]
 *   expand: out[r, c] = act(bias[c] + sum_j a_j[r] w[j, c])          out (R, C) contiguous; bias nullable; relu flag
 *   bwd:    d' = d * [y > 0] (y nullable: the forward output of a relu expand)
 *           dz (nullable, laid out like a):   dz_j[r] = sum_c d'[r, c] w[j, c]
 *           partial (nullable): (dstag_thin_slices(), J + 1, C) per-CTA partials of dw[j, c] = sum_r a_j[r] d'[r, c]
 *           and (row J) db[c] = sum_r d'[r, c]; the caller adds them in a fixed order with dstag_sum_slices. */
int dstag_thin_supported(int32_t J, int32_t C);
int dstag_thin_slices(void);
int dstag_thin_expand(const float* a, int64_t a_plane, int32_t taps, int32_t F, const float* w, int64_t w_js,
                      int64_t w_cs, const float* bias, float* out, int64_t R, int32_t C, int32_t relu, void* stream);
int dstag_thin_bwd(const float* a, int64_t a_plane, int32_t taps, int32_t F, const float* w, int64_t w_js, int64_t w_cs,
                   const float* d, const float* y, float* dz, float* partial, int64_t R, int32_t C, void* stream);

/* Unfused tail (model/DSTAGNN_my.py:243-252) for any T and C <= 256:
 *   tc' = dropout(tc) ; tco = relu(tc' [+ g]) ; out = LN_C(relu(r + tco)) over the channel axis
 *   tc, g (NULL in block 0, line 245), r, out: (X, T, C);  mean/rstd: (X, T) saved statistics.
 * Dropout element index = (x*C + c)*T + t (the reference's (B,N,C,T) order). */
int dstag_tail_fwd(const float* tc, const float* g, const float* r, const float* gamma,
                   const float* beta, float eps, float* out, float* mean, float* rstd, int64_t X,
                   int32_t C, int32_t T, float drop_p, uint64_t seed, const uint64_t* seed_dev, void* stream);
/* dtc, dr, dg (when g != NULL): (X, T, C); accumulates dgamma / dbeta (caller-zeroed). */
int dstag_tail_bwd(const float* dout, const float* tc, const float* g, const float* r,
                   const float* gamma, const float* mean, const float* rstd, float* dtc, float* dg,
                   float* dr, float* dgamma, float* dbeta, int64_t X, int32_t C, int32_t T,
                   float drop_p, uint64_t seed, const uint64_t* seed_dev, void* stream);

/* fcmy (Linear over the concatenated time axis, line 243) fused with the tail, for C = 32, T = 12:
 *   tc[x, t, c] = bf[t] + sum_s wf[t, s] * cat[x, s, c]   evaluated in registers, then the tail above.
 * cat (X, S, C); wf (T, S); bf (T).  dstag_tail_fused_supported(C, T) tells whether the shape is covered. */
int dstag_tail_fused_supported(int32_t C, int32_t T);
int dstag_tail_fused_fwd(const float* cat, const float* wf, const float* bf, const float* g, const float* r,
                         const float* gamma, const float* beta, float eps, float* out, float* mean, float* rstd,
                         int64_t X, int32_t C, int32_t T, float drop_p, uint64_t seed, const uint64_t* seed_dev,
                         void* stream);
/* Backward: dcat (X, S, C), dtc (X, T, C) (feeds dstag_outer_sum for the fcmy weight / bias gradients),
 * dr, dg (when g != NULL) (X, T, C); accumulates dgamma / dbeta. */
int dstag_tail_fused_bwd(const float* dout, const float* cat, const float* wf, const float* bf, const float* g,
                         const float* r, const float* gamma, const float* mean, const float* rstd, float* dcat,
                         float* dtc, float* dg, float* dr, float* dgamma, float* dbeta, int64_t X, int32_t C,
                         int32_t T, float drop_p, uint64_t seed, const uint64_t* seed_dev, void* stream);

/* ------------------------------------------------------------------------------------
 * The step either side of the block stack.
 *
 * SmoothL1Loss (beta = 1, reduction = mean; train_DSTAGNN_my.py:132,156):
 *   loss = mean(0.5 d^2 if |d| < 1 else |d| - 0.5), d = pred - target, n elements.
 *   workspace: 1025 floats of caller-owned scratch, ZEROED ONCE when allocated (partial sums + a
 *   self-resetting arrival counter); the partials are added in index order: reproducible.
 * Backward: dpred = gout * clamp(d, -1, 1) / n with gout a DEVICE scalar (nullable = 1).
 * ---------------------------------------------------------------------------------- */
int dstag_smooth_l1_fwd(const float* pred, const float* target, float* loss, float* workspace, int64_t n,
                        void* stream);
int dstag_smooth_l1_bwd(const float* pred, const float* target, const float* gout, float* dpred, int64_t n,
                        void* stream);

/* Adam on one flat fp32 buffer (torch.optim.Adam with the reference's defaults, train_DSTAGNN_my.py:133,158:
 * betas given by the caller, no weight decay, no amsgrad):
 *   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 * with g = grad_scale * g (1/world_size after a summing all-reduce).  The step count t (1-based) is read from
 * the device counter step_dev when given (CUDA-graph replays), else step_host.  16-byte aligned buffers. */
int dstag_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                    float eps, float grad_scale, const uint64_t* step_dev, uint64_t step_host, void* stream);

/* flat[offsets[t] + i] = srcs[t][i], i < sizes[t], for t < n: collects the step's parameter gradients (separate
 * device tensors; srcs / offsets / sizes are HOST arrays, passed to the kernel by value) into the flat bucket that
 * the gradient all-reduce (train_DSTAGNN_my.py:148,158) and dstag_adam_step consume. */
int dstag_gather_flat(const float* const* srcs, const int64_t* offsets, const int64_t* sizes, int32_t n, float* flat,
                      void* stream);

/* n strided copies of <= 4-D fp32 blocks in one launch per 16 (element strides, may be negative: a flipped
 * axis): the per-step re-layouts of small parameter tensors -- packed [W_Q | W_K | W_V] rows, stacked Theta_k, the
 * GTU conv matrices and their flipped transposes (model/DSTAGNN_my.py:62-63, 92-94, 112-113, 190) -- and the
 * un-packing of their gradients.  copies: HOST array (passed to the kernel by value). */
typedef struct {
  const void* src;
  void* dst;
  int32_t n[4];            /* extents, outermost first; unused leading extents = 1 */
  int64_t src_stride[4];
  int64_t dst_stride[4];
} dstag_copy_t;
int dstag_multi_copy(const dstag_copy_t* copies, int32_t n, void* stream);

/* counters[i] += 1 for i < n (n <= 32): the device-resident step counters (dropout key, Adam step). */
int dstag_tick(uint64_t* counters, int32_t n, void* stream);

/* biasT[k, j, i] = adj_pa[i, j] * mask_k[i, j] for i < N, 0 for N <= i < ld: the batch-independent additive
 * term of model/DSTAGNN_my.py:126 in the transposed, padded map layout.  masks: HOST array of K (<= 8) device
 * pointers to the (N, N) mask parameters.  Backward: dmask[k, i, j] = adj_pa[i, j] * dbiasT[k, j, i]. */
int dstag_bias_transpose(const float* adj_pa, const float* const* masks, int32_t K, float* biasT, int32_t N,
                         int32_t ld, void* stream);
int dstag_bias_transpose_bwd(const float* adj_pa, const float* dbiasT, float* dmask, int32_t K, int32_t N, int32_t ld,
                             void* stream);

/* ------------------------------------------------------------------------------------
 * Sparse Chebyshev path of cheb_conv_withSAt (model/DSTAGNN_my.py:117-133) -- csrc/sparse.cu.
 *
 * The polynomials T_k of this model are element-wise (Hadamard) powers of the scaled Laplacian
 * (lib/utils.py:198-203): T_0 = I and every T_k carries the adjacency's sparsity pattern (~1 % with the STAG
 * graphs), so the operator T_k (.) softmax(S_k) and the surviving part of its gradient live on the graph's
 * EDGES.  Edge e of order k joins output node j to input node i with weight tval[e] = T_k[i, j]; the
 * E edges are stored as a CSR over rows r = k * N + j (rowptr (K N + 1), col = i, erow = r) and as the
 * transposed CSR over rows k * N + i (t_rowptr, t_col = j, t_eid = index of the same edge in the row CSR).
 * All index arrays are int32 device pointers; a, h are (B, E); r is (B, K, N); row length L % 4 == 0.
 *
 *   dstag_edge_probs   a[b,e] = tval[e] * P[b,k,j,i]                                  (line 128 on the edges)
 *   dstag_spmm_expand  dst[k,b,n,:] = sum_{e in row (k,n)} a[b,eid?eid[e]:e] * src[b,col[e],:]   Z_k = A_k^T X (l.130)
 *   dstag_spmm_reduce  dst[b,n,:] = sum_k sum_{e in row (k,n)} a[b,..] * src[k,b,col[e],:]      dX = sum_k A_k dZ_k
 *   dstag_edge_dots    h[b,e] = a[b,e] * <dZ[k,b,j,:], X[b,col[e],:]>, r[b,k,j] = sum_{e in row} h[b,e]
 *   dstag_scale_keys   rK[b,j,k,:] = r[b,k,j] * Ks[b,j,k,:]          qk (B,N,2,K,dk) packed [Qs | Ks], rK (B,N,K,dk)
 *   dstag_center_queries Qc[b,n,k,:] = Qs[b,n,k,:] - mean_n Qs[b,n,k,:]  (rows of dS sum to zero: a vector common
 *                      to all queries drops out of the key gradient; removing it first avoids its cancellation)
 *   dstag_dqk_combine  dqk[b,n,0,k,:] = alpha (sum_{e in col (k,n)} h[b,e] Ks[b,j_e,k,:] - W[b,n,k,:])
 *                      dqk[b,n,1,k,:] = alpha (sum_{e in row (k,n)} h[b,e] Qc[b,i_e,k,:] - r[b,k,n] O[b,n,k,:])
 *                      with O = P Qc and W = P^T (r Ks) from dstag_gemm: the softmax backward
 *                      dS = h (on edges) - r P is never materialised
 *   dstag_dbias_sparse dbiasT[k, pa_j[p], pa_i[p]] = sum_b (h[b, pa_eid[k,p]] if >= 0) - r[b,k,j] P[b,k,j,i]
 *                      over the npa non-zeros of adj_pa (the only entries the mask gradient keeps)
 * ---------------------------------------------------------------------------------- */
int dstag_edge_probs(const float* P, const int32_t* erow, const int32_t* col, const float* tval, float* a, int32_t B,
                     int32_t K, int32_t N, int32_t ld, int32_t E, void* stream);
int dstag_spmm_expand(const float* src, const float* a, const int32_t* rowptr, const int32_t* col, const int32_t* eid,
                      float* dst, int32_t B, int32_t K, int32_t N, int32_t L, int32_t E, void* stream);
int dstag_spmm_reduce(const float* src, const float* a, const int32_t* rowptr, const int32_t* col, const int32_t* eid,
                      float* dst, int32_t B, int32_t K, int32_t N, int32_t L, int32_t E, void* stream);
int dstag_edge_dots(const float* dZ, const float* X, const float* a, const int32_t* rowptr, const int32_t* col, float* h,
                    float* r, int32_t B, int32_t K, int32_t N, int32_t L, int32_t E, void* stream);
int dstag_scale_keys(const float* qk, const float* r, float* rK, int32_t B, int32_t N, int32_t K, int32_t dk,
                     void* stream);
int dstag_center_queries(const float* qk, float* sums /* (B, K dk) scratch */, float* Qc, int32_t B, int32_t N,
                         int32_t K, int32_t dk, void* stream);
int dstag_dqk_combine(const float* qk, const float* Qc, const float* h, const float* r, const float* O, const float* W,
                      const int32_t* rowptr, const int32_t* col, const int32_t* t_rowptr, const int32_t* t_col,
                      const int32_t* t_eid, float* dqk, int32_t B, int32_t N, int32_t K, int32_t dk, int32_t E,
                      float alpha, void* stream);
int dstag_dbias_sparse(const float* P, const float* h, const float* r, const int32_t* pa_j, const int32_t* pa_i,
                       const int32_t* pa_eid, float* dbiasT, int32_t B, int32_t K, int32_t N, int32_t ld, int32_t E,
                       int32_t npa, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DSTAGNN_B200_H */
