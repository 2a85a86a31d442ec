/*
 * b200ad.h — C ABI of libb200ad.so: the B200 (sm_100a) replacement for the array-library calls
 * (NumPy / CuPy) that the reference Jianqoq/Auto-Differentiation makes on its autodiff hot path.
 *
 * Boundary: every tape op of the reference is a subclass of `Function`
 * implementing `result()` (auto_grad/tgrad.py:48-52) and `get_grad(grad)` (:54-58); every
 * hand-written layer is a `forward()/backward()` pair over `cupy as np` (layers/layers.py:6).
 * Each entry point below replaces the NumPy/CuPy call(s) inside one of those bodies; the
 * reference-side binding a maintainer would add is a ctypes stub.
 *
 * Conventions
 *   - plain C, no torch types. All tensor pointers are DEVICE pointers to contiguous row-major
 *     fp32 unless a stride/ld argument says otherwise; labels/indices are int64 (NumPy default).
 *   - the caller owns every buffer (including workspaces); the library never allocates.
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued and the call returns.
 *   - return 0 on success, negative B200AD_ERR_* otherwise; b200ad_last_error() gives the text.
 *     Shape validation happens before any launch.  The Python shim re-raises the reference's
 *     exception types (tgrad.py:2179 RuntimeError, :271 ValueError, :375/:499/:1296 AssertionError).
 *   - no global mutable state, no CPU fallback: without a CUDA device every compute entry fails.
 */
#ifndef B200AD_H
#define B200AD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200AD_OK 0
#define B200AD_ERR_INVALID (-1)     /* bad argument / shape */
#define B200AD_ERR_CUDA (-2)        /* launch or runtime error */
#define B200AD_ERR_UNSUPPORTED (-3) /* shape/alignment not supported by this entry */

#define B200AD_ABI_VERSION 1

int b200ad_abi_version(void);
const char* b200ad_last_error(void);
/* number of kernels launched by this library in the calling process since load (bench `gpu_launches`). */
int64_t b200ad_launch_count(void);
/* Fills `name` (cap bytes) with the device name, returns SM count, or <0 when no CUDA device. */
int b200ad_device_info(char* name, int cap, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * Elementwise unary ops.  fwd: y = f(x, param).  bwd: gx = g * f'(.)
 *   replaces np.exp/np.tanh/... in  exp.result tgrad.py:1073-1075, tanh :1613-1615,
 *   Sigmoid :901-903, ln :1294-1297, log :1338-1341, sqrt :1784-1786, square :1741-1743,
 *   pow :1103-1105, Abs :1846-1848, sin/cos/tan/sec/sinh/cosh/arcsin/arctan/arcot/arcsec/cot/csc/arcos
 *   (:1115-1713), layers.ReLU layers.py:1284-1290;  and their get_grad bodies
 *   (:1058-1060, :1596-1598, :887-888, :1280-1282, :1324-1326, :1769-1772, :1725-1728, :1091-1093,
 *    :1850-1853).
 *   bwd reads `x` or `y` depending on the op (b200ad_unary_bwd_uses): pass the one it uses,
 *   the other may be NULL.
 * ------------------------------------------------------------------------------------------ */
enum b200ad_unary_op {
  B200AD_U_EXP = 0,
  B200AD_U_TANH = 1,
  B200AD_U_SIGMOID = 2,
  B200AD_U_RELU = 3,
  B200AD_U_LN = 4,
  B200AD_U_LOGB = 5,   /* param = base:  ln(x)/ln(base) */
  B200AD_U_SQRT = 6,
  B200AD_U_SQUARE = 7,
  B200AD_U_POW = 8,    /* param = exponent */
  B200AD_U_ABS = 9,
  B200AD_U_NEG = 10,
  B200AD_U_SIN = 11,
  B200AD_U_COS = 12,
  B200AD_U_TAN = 13,
  B200AD_U_SEC = 14,
  B200AD_U_SINH = 15,
  B200AD_U_COSH = 16,
  B200AD_U_ARCSIN = 17,
  B200AD_U_ARCTAN = 18,
  B200AD_U_ARCOT = 19,
  B200AD_U_ARCSEC = 20,
  B200AD_U_SCALE = 21, /* y = param * x   (Multi with a 0-d operand, tgrad.py:147-148) */
  B200AD_U_ADDC = 22,  /* y = x + param   (Add/sub with a host-scalar operand) */
  B200AD_U_DIVC = 23,  /* y = x / param   (Divide by a host scalar, e.g. / ad.sqrt(d_k) Test:76);
                          bwd g/(param+1e-8) keeps the epsilon of tgrad.py:1004 */
  B200AD_U_RSUBC = 24, /* y = param - x */
  B200AD_U_RDIVC = 25, /* y = param / x ; bwd g*(-param/(x*x+1e-8)) tgrad.py:1005 */
  B200AD_U_COT = 26,   /* 1/tan(x); bwd +g*csc(x)^2: the sign of tgrad.py:1639-1641 (the derivative's sign is dropped there) */
  B200AD_U_CSC = 27,   /* 1/sin(x); bwd -g*csc(x)*cot(x): the intent of :1476-1478 (as shipped it calls np.csc and raises) */
  B200AD_U_ARCCOS = 28,/* arccos(x); bwd g*(-1/sqrt(x+1)) as written at :1438-1440 */
  B200AD_U_COUNT = 29
};
int b200ad_unary_fwd(int op, const float* x, float* y, int64_t n, float param, void* stream);
int b200ad_unary_bwd(int op, const float* g, const float* x, const float* y, float* gx, int64_t n,
                     float param, void* stream);
/* 0 → bwd reads x, 1 → bwd reads y, 2 → neither */
int b200ad_unary_bwd_uses(int op);

/* ------------------------------------------------------------------------------------------
 * Elementwise binary ops with NumPy broadcasting collapsed to a 2-D (rows, period) pattern.
 *   operand mode: FULL  index i            (same shape as the output)
 *                 ROW   index i % period   ((1,C) / (C,) operand of a (..., C) output)
 *                 COL   index i / period   ((R,1) operand of an (R, period) output)
 *                 SCALAR index 0
 *   replaces np.add Add.result tgrad.py:682-684, np.subtract sub.result :769-771,
 *   np.multiply Multi.result :878-880 / starmulti :585-587, Divide.result :1026-1028.
 * bwd writes the PRE-reduction operand gradients at the full output shape (either may be NULL):
 *   add: (g, g)   sub: REF_COMPAT (-g, +g) — sign-inverted exactly like tgrad.py:753-756, or
 *   the mathematically right (+g, -g) when ref_compat == 0;  mul: (g*b, g*a) :850-852;
 *   div: (g/(b+1e-8), g*(-a/(b*b+1e-8))) :1003-1006 (the epsilons are part of the contract).
 * The caller then applies b200ad_reduce (check_shape, tgrad.py:2163-2187) to broadcast operands.
 * b200ad_binary_bwd_rowfused does both in one pass for the dominant pattern
 *   a FULL (R,C), b ROW (C):  ga[R,C] full, gb[C] = column sums, partials in `ws`
 *   (ws floats >= b200ad_colsum_ws_floats(R, C)).
 * ------------------------------------------------------------------------------------------ */
enum b200ad_binary_op { B200AD_B_ADD = 0, B200AD_B_SUB = 1, B200AD_B_MUL = 2, B200AD_B_DIV = 3 };
enum b200ad_bcast_mode { B200AD_FULL = 0, B200AD_ROW = 1, B200AD_COL = 2, B200AD_SCALAR = 3 };
int b200ad_binary_fwd(int op, const float* a, const float* b, float* out, int64_t n, int a_mode,
                      int b_mode, int64_t period, void* stream);
int b200ad_binary_bwd(int op, const float* g, const float* a, const float* b, float* ga, float* gb,
                      int64_t n, int a_mode, int b_mode, int64_t period, int ref_compat,
                      void* stream);
int b200ad_binary_bwd_rowfused(int op, const float* g, const float* a, const float* b, float* ga,
                               float* gb, int64_t rows, int64_t cols, int ref_compat, float* ws,
                               void* stream);

/* ------------------------------------------------------------------------------------------
 * Reductions.  b200ad_reduce sums `in` viewed as (outer, red, inner) over `red`:
 *   out[o, i] = sum_r in[o, r, i]   — the kernel behind check_shape (tgrad.py:2163-2187),
 *   sum.result (:720-721), repeat.get_grad (:781-783), TimeAffine db (layers.py:306),
 *   LSTM db (:674), FeedForward db (:1264).  Deterministic two-stage (no atomics);
 *   ws floats >= b200ad_reduce_ws_floats(outer, red, inner).
 * b200ad_repeat_axis: out[o, r, i] = in[o, i]  (sum.get_grad :699-704, repeat.result :795-797).
 * b200ad_sumsq: out[0] (+)= sum(x^2)   (clip_grads layers.py:132-141);  accumulate != 0 adds.
 * ------------------------------------------------------------------------------------------ */
int64_t b200ad_reduce_ws_floats(int64_t outer, int64_t red, int64_t inner);
int64_t b200ad_colsum_ws_floats(int64_t rows, int64_t cols);
int b200ad_reduce(const float* in, float* out, int64_t outer, int64_t red, int64_t inner,
                  float scale, float* ws, void* stream);
int b200ad_repeat_axis(const float* in, float* out, int64_t outer, int64_t rep, int64_t inner,
                       float scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Softmax family over the last axis of a (rows, cols) view.
 *   softmax fwd: stable form  softmax.result tgrad.py:964-970, SoftmaxWithLoss.result :926-932,
 *                layers.softmax layers.py:38-43
 *   softmax bwd: gx = y*g - y*sum(y*g)            softmax.get_grad :947-951
 *   diag bwd:    gx = g*y*(1-y)                   SoftmaxWithLoss.get_grad :918-919 (as shipped)
 *   xent fwd:    labels int64 (ignore == ignore_label), probs out, per-row -log p[t]*mask into
 *                row_loss, then loss[0] = sum(row_loss)/count(mask)   TimeSoftmaxWithLoss.forward
 *                layers.py:514-535; eps is added inside the log for multi_cross_entropy_error
 *                (utils/activation_func.py:4-13, eps=1e-7, denom = rows).
 *   xent bwd:    dx = (p - onehot) * dout / denom * mask            layers.py:538-549
 *   denom_mode 0: count of non-ignored rows; 1: rows.
 * ------------------------------------------------------------------------------------------ */
int b200ad_softmax_fwd(const float* x, float* y, int64_t rows, int64_t cols, float in_scale,
                       void* stream);
int b200ad_softmax_bwd(const float* g, const float* y, float* gx, int64_t rows, int64_t cols,
                       float out_scale, int diag_only, void* stream);
int b200ad_xent_fwd(const float* logits, const int64_t* labels, float* probs, float* row_loss,
                    float* loss_and_count /* [2] */, int64_t rows, int64_t cols,
                    int64_t ignore_label, float eps, int denom_mode, int input_is_probs,
                    void* stream);
int b200ad_xent_bwd(const float* probs, const int64_t* labels, const float* loss_and_count,
                    float* dx, int64_t rows, int64_t cols, int64_t ignore_label, float dout,
                    int denom_mode, void* stream);

/* ------------------------------------------------------------------------------------------
 * Integer / index work (bit-exact).
 *   argmax/argmin over the flattened array, first occurrence:  argmax.result tgrad.py:1810-1812,
 *   argmin :1827-1829.  NaN handling follows NumPy (first NaN wins).  ws >= 2*1184 int64-slots.
 *   argmax_rows: per-row argmax (one-hot → label, `ts.argmax(axis=2)` layers.py:519-520).
 *   embedding gather  W[idx]  layers.py:451;  scatter-add  dW[...]=0; np.add.at(dW, idx, dout)
 *   :455-457 — deterministic: rows are accumulated in ascending position order per target row.
 *   slice_scatter_rows: zeros(x.shape)[index] = grad for a leading-axis slice (Slice.get_grad :370-378).
 * ------------------------------------------------------------------------------------------ */
int b200ad_argext_flat(const float* x, int64_t n, int is_min, int64_t* out, void* ws, void* stream);
int b200ad_argmax_rows(const float* x, int64_t rows, int64_t cols, int64_t* out, void* stream);
int b200ad_embedding_gather(const float* w, const int64_t* idx, float* out, int64_t n_idx,
                            int64_t vocab, int64_t dim, void* stream);
int b200ad_embedding_scatter_add(const float* dout, const int64_t* idx, float* dw, int64_t n_idx,
                                 int64_t vocab, int64_t dim, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layout: strided → contiguous (or strided → strided) copy of a rank<=6 view: transpose.result
 * tgrad.py:611-613 when a consumer needs contiguous data; tensordot's permute+reshape :294-303;
 * `zeros[index] = grad` of Slice.get_grad :376-378.  Strides in elements.
 * ------------------------------------------------------------------------------------------ */
int b200ad_strided_copy(const float* in, float* out, int rank, const int64_t* shape,
                        const int64_t* in_strides, const int64_t* out_strides /* NULL: contiguous */,
                        void* stream);
int b200ad_fill(float* out, int64_t n, float value, void* stream);
int b200ad_axpby(const float* x, float* y, int64_t n, float a, float b, void* stream); /* y = a*x + b*y */
/* TimeDropout.forward layers.py:685-692: mask = (u > ratio) * scale, y = x * mask, u = caller's uniform draws.
 * Backward (:696-697) is binary_fwd(MUL, dout, mask). */
int b200ad_dropout_fwd(const float* x, const float* u, float* mask, float* y, int64_t n, float ratio,
                       float scale, void* stream);
/* U[0,1) draws on the device for the dropout masks (reference: np.random.rand, layers/layers.py:687): Philox4x32-10,
 * element i comes from counter offset + i/4; callers advance `offset` by ceil(n/4) per call.  24-bit resolution. */
int b200ad_uniform(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream);

/* ------------------------------------------------------------------------------------------
 * Dense contractions.   C[b0,b1] = alpha * op(A[b0,b1]) @ op(B[b0,b1]) + beta * C[b0,b1]
 *   op(A) is M x K: element (m,k) at A[m*lda + k] (transA=0) or A[k*lda + m] (transA=1);
 *   op(B) is K x N likewise; C is M x N with leading dimension ldc.  Two batch levels with
 *   independent element strides (0 = broadcast) cover the 4-D head-split views without copies.
 *   Optional fused epilogue: + bias[n] (may be NULL), then activation (0 none, 1 tanh,
 *   2 sigmoid, 3 relu); `colsum_out` is not fused (use b200ad_reduce).
 *   replaces np.matmul Matmul.result tgrad.py:535-539 and get_grad :530-533 (2-D and 4-D),
 *   tensordot's single GEMM :304 and its two backward GEMMs :222-228, np.dot TimeAffine
 *   layers.py:280,307,308, LSTM gate GEMMs :632,:666-669, FeedForward :1252,:1263,:1266,
 *   MultiHeadAttention :1347-1360.
 *   precision: B200AD_GEMM_FP32   SIMT fp32 FMA (exact-fp32 accumulate; any shape/alignment)
 *              B200AD_GEMM_TF32X3 tcgen05 kind::tf32 with 3-term error-compensated split
 *                                 (fp32-faithful: |err| ~ 2^-21 relative; meets rtol 1e-5)
 *              B200AD_GEMM_BF16   tcgen05 kind::f16 bf16 inputs, fp32 accumulate in TMEM
 *                                 (looser tolerance: rtol 2e-2 of |A||B| row-norm bound)
 *              B200AD_GEMM_F16X3  tcgen05 kind::f16 on fp16 hi/lo pairs with one power-of-two scale per operand
 *                                 row, 3-term split (fp32-faithful like TF32X3 at twice its MMA rate; the operands
 *                                 are split by a pre-pass into the workspace, or — a row-major op(A) — converted inside
 *                                 the kernel after a read-only row-scale pass, csrc/gemm_h16.cu)
 *              B200AD_GEMM_AUTO   fp32-faithful: F16X3 for the large row-major-A products where it measures faster
 *                                 (b200ad_gemm_set_f16x3_auto), TF32X3 when the shape/alignment allows the TMA
 *                                 path, else FP32
 *   b200ad_gemm_supported tells whether `precision` can run this problem (1) or not (0).
 * ------------------------------------------------------------------------------------------ */
enum b200ad_gemm_precision {
  B200AD_GEMM_FP32 = 0,
  B200AD_GEMM_TF32X3 = 1,
  B200AD_GEMM_BF16 = 2,
  B200AD_GEMM_AUTO = 3,
  B200AD_GEMM_TF32 = 4, /* single-pass tf32 (diagnostic / roofline probe) */
  B200AD_GEMM_F16X3 = 5
};
typedef struct b200ad_gemm_desc {
  int32_t precision, trans_a, trans_b, act;
  int64_t m, n, k;
  int64_t lda, ldb, ldc;
  int64_t batch0, batch1;
  int64_t a_stride0, a_stride1, b_stride0, b_stride1, c_stride0, c_stride1;
  float alpha, beta;
} b200ad_gemm_desc;
int b200ad_gemm(const b200ad_gemm_desc* d, const float* a, const float* b, float* c,
                const float* bias, void* workspace, int64_t workspace_bytes, void* stream);
int b200ad_gemm_supported(const b200ad_gemm_desc* d, const float* a, const float* b, const float* c);
int64_t b200ad_gemm_workspace_bytes(const b200ad_gemm_desc* d);
/* B200AD_GEMM_AUTO has two implementations of the M <= 64 recurrent-step product (h_{t-1} @ Wh, dA_t @ Wh^T;
 * reference time loops layers/layers.py:203-207, :573-576, :588-596): tcgen05 64-column tiles + split-K, and a
 * single-launch exact-fp32 cluster kernel.  mode 0 (default) picks by whether `stream` is being captured into a
 * CUDA graph (device time vs host launch cost), 1 pins the first, 2 the second.  Returns the previous mode. */
int b200ad_gemm_set_recurrent_dispatch(int mode);
/* The B200AD_TC_* tuning / debugging environment knobs of the tcgen05 GEMM are read once; this re-reads them. */
int b200ad_gemm_reload_env(void);
/* When B200AD_GEMM_AUTO resolves to F16X3: 0 = never, 1 (default; environment B200AD_GEMM_F16X3_AUTO) = where a device-time
 * model built from measurements says it beats TF32X3 by >= 10 % (the operand split pre-pass must pay for itself: large tiles
 * with K >= ~2048, or one operand shared by several products), 2 = whenever F16X3 supports the problem.  Returns the
 * previous mode. */
int b200ad_gemm_set_f16x3_auto(int mode);
/* C_i = alpha * op(A) @ op(B_i) + beta * C_i (+ bias_i, activation) for i < count with ONE shared A operand: the
 * Q / K / V projections of self-attention multiply the same input (layers/layers.py:1347-1349).  With precision
 * F16X3 (or AUTO resolving to it) A is split once for all products; any other precision loops b200ad_gemm.
 * All B_i / C_i share the descriptor's geometry; `biases` may be NULL.  Workspace: b200ad_gemm_shared_a_workspace_bytes
 * (AUTO decides with the shared split amortised over `count` products). */
int64_t b200ad_gemm_shared_a_workspace_bytes(const b200ad_gemm_desc* d, int32_t count);
int b200ad_gemm_shared_a(const b200ad_gemm_desc* d, const float* a, int32_t count, const float* const* bs,
                         float* const* cs, const float* const* biases, void* workspace, int64_t workspace_bytes,
                         void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused multi-head attention core (tcgen05, fp32-faithful three-product split; scores never leave the SM).
 *   Two carriers of the same split: tf32 hi/lo pairs (csrc/attn_tc.cu, variant 1) and fp16 hi/lo pairs at the
 *   kind::f16 MMA rate with one power-of-two scale per (batch, head) operand slab (csrc/attn_h16.cu, variant 0).
 *   b200ad_attn_set_variant(-1 | 0 | 1 | 2 | 3) selects one and returns the previous choice; -1 (default) = the measured
 *   best mix: forward on variant 1, backward on variant 0.  The fp16 backward computes dK and dV in ONE pass (S and P
 *   once for both); variant 2 = variant 0 with dV and dK as two separate passes (A/B measurements, second witness);
 *   variant 3 = the LOOSER-TOLERANCE variant: variant 0 issuing only the hi . hi product of every contraction (fp16 operands
 *   with 11 significant bits, fp32 accumulation): relative error ~1e-3 of the tensor scale, a third of the MMA work — the
 *   attention counterpart of B200AD_GEMM_BF16.
 *   fwd:  O = softmax(scale * Q K^T) V  per (batch, head);  lse[b,h,s] = log2-domain log-sum-exp
 *         of the scaled scores of query row s (kept for the backward instead of the S x S matrix).
 *         replaces MultiHeadAttention.forward layers.py:1351-1357 (split heads are strided views:
 *         element (b, h, s, d) of a tensor sits at ptr + b*batch + h*head + s*row + d).
 *   bwd:  dQ, dK, dV of the same expression given dO (the tape gradients of
 *         Test/MultiHeadAttentionTest.py:70-85: Matmul :530-533, Divide :1004-1005 without the 1e-8
 *         that only guards a zero denominator, exp :1058-1060, sum :699-704).
 *         dO shares `o`'s strides, dQ/dK/dV share q/k/v's.  ws: b200ad_attn_{fwd,bwd}_ws_floats floats of scratch
 *         (the forward's may be NULL on the default variant; when given, it receives the slab maxima of Q / K / V).
 *   Built for head_dim <= 64 and a multiple of 4 (tiles are 64 wide: smaller heads are zero-filled on the way in by the
 *   tensor maps and clipped on the way out), any seq_q / seq_kv, strides multiples of 4 elements, 16 B aligned
 *   bases; b200ad_attn_supported says whether the fused path applies (otherwise — e.g. the reference test's
 *   d_k = 5, whose rows are not 16 B aligned — the caller composes the same math from b200ad_gemm + b200ad_softmax_*).
 * ------------------------------------------------------------------------------------------ */
typedef struct b200ad_attn_strides {
  int64_t batch, head, row; /* element strides */
} b200ad_attn_strides;
typedef struct b200ad_attn_desc {
  int64_t batch, heads, seq_q, seq_kv, head_dim;
  float scale;
  int32_t reserved;
  b200ad_attn_strides q, k, v, o;
} b200ad_attn_desc;
int b200ad_attn_supported(const b200ad_attn_desc* d);
int b200ad_attn_set_variant(int variant);
/* b200ad_attn_bwd_maxima: b200ad_attn_bwd with the workspace of the b200ad_attn_fwd call on the SAME q / k / v (fwd_ws; may
 * be NULL): the forward leaves max |Q|, |K|, |V| per (batch, head) in it, which saves the fp16-split backward three of the
 * five tensor reads of its pre-pass. */
int64_t b200ad_attn_fwd_ws_floats(const b200ad_attn_desc* d);
int b200ad_attn_fwd(const b200ad_attn_desc* d, const float* q, const float* k, const float* v,
                    float* o, float* lse, float* ws, void* stream);
int64_t b200ad_attn_bwd_ws_floats(const b200ad_attn_desc* d);
int b200ad_attn_bwd(const b200ad_attn_desc* d, const float* q, const float* k, const float* v,
                    const float* o, const float* dout, const float* lse, float* dq, float* dk,
                    float* dv, float* ws, void* stream);
int b200ad_attn_bwd_maxima(const b200ad_attn_desc* d, const float* q, const float* k, const float* v,
                           const float* o, const float* dout, const float* lse, const float* fwd_ws,
                           float* dq, float* dk, float* dv, float* ws, void* stream);

/* ------------------------------------------------------------------------------------------
 * Recurrent cells (elementwise halves; the gate GEMMs go through b200ad_gemm).
 *   lstm fwd: A (N,4H) pre-activations with column blocks [f,g,i,o] = [sigmoid,tanh,sigmoid,sigmoid]
 *             c' = f*c + g*i ; h' = o*tanh(c')     LSTM.forward layers.py:624-646.
 *             gates (N,4H) receives the activated gates (cache for backward).
 *   lstm bwd: hand-derived gate gradients dA (N,4H) [stack0..3] and dc   LSTM.backward :648-664,670.
 *             dh_a + dh_b is the incoming dh (upstream dhs[:,t,:] plus recurrent dh; dh_b may be NULL).
 *   rnn  fwd: h' = tanh(pre)  RNN.forward layers.py:166-173;  bwd: dt = (dh_a+dh_b)*(1-h'^2) :176-177.
 *   ld* are row strides in elements so a time slice [:, t, :] of an (N,T,X) tensor is addressed in place.
 * ------------------------------------------------------------------------------------------ */
int b200ad_lstm_cell_fwd(const float* a, int64_t lda, const float* c_prev, int64_t ldc_prev,
                         float* gates, int64_t ldg, float* c_next, int64_t ldc_next, float* h_next,
                         int64_t ldh, int64_t n, int64_t h, void* stream);
int b200ad_lstm_cell_bwd(const float* dh_a, int64_t ld_dha, const float* dh_b, int64_t ld_dhb,
                         const float* dc_next, const float* gates, int64_t ldg, const float* c_prev,
                         int64_t ldc_prev, const float* c_next, int64_t ldc_next, const float* h_next,
                         int64_t ldh, float* da, int64_t ldda, float* dc_prev, int64_t n, int64_t h,
                         void* stream);
int b200ad_rnn_cell_bwd(const float* dh_a, int64_t ld_dha, const float* dh_b, int64_t ld_dhb,
                        const float* h_next, int64_t ldh, float* dt, int64_t lddt, int64_t n,
                        int64_t h, void* stream);
/* out[r, c] = act(x[r, c] + bias[c]) with row strides (bias may be NULL) */
/* Persistent tanh-RNN time loop: ONE launch runs all T dependent steps (forward) / all T+1 BPTT phases (backward) with
 * the recurrent weight resident in tensor memory (csrc/rnn_seq.cu).  Replaces the Python time loops
 * TimeRNN.forward layers/layers.py:203-207 and TimeRNN.backward :210-225 (intended math).
 *   fwd: hs_ext is (T+1, N, H) time-major with hs_ext[0] = h_0 on entry; pre is (T, N, H) = x_t Wx + b (one big GEMM by
 *        the caller); on exit hs_ext[t+1] = tanh(pre[t] + hs_ext[t] @ Wh).  hs_ext[1:] is overwritten.
 *   bwd: dA[t] = (dhs[n,t,:] + dA[t+1] @ Wh^T) * (1 - hs_ext[t+1]^2) for t = T-1..0 (dA is (T, N, H) time-major),
 *        dh0 = dA[0] @ Wh^T (may be NULL).  dhs element (n, t, c) sits at dhs[n*stride_n + t*stride_t + c].
 *   b200ad_rnn_seq_supported(batch, hidden) is 1 when this device can keep all clusters co-resident for the shape
 *   (batch <= 64, hidden in {256, 512, 1024}); otherwise callers run the per-step loop over b200ad_gemm. */
int b200ad_rnn_seq_supported(int64_t batch, int64_t hidden);
int b200ad_rnn_tanh_seq_fwd(float* hs_ext, const float* pre, const float* wh, int64_t T, int64_t N, int64_t H,
                            void* stream);
int b200ad_rnn_tanh_seq_bwd(const float* dhs, int64_t dhs_stride_n, int64_t dhs_stride_t, const float* hs_ext,
                            const float* wh, float* dA, float* dh0, int64_t T, int64_t N, int64_t H, void* stream);
/* LSTM time loops from ONE call (csrc/lstm_seq.cu): the T dependent steps of TimeLSTM.forward (layers/layers.py:573-576,
 * cell :624-646) and of its BPTT (:588-596, cell :648-676), two launches per step (tcgen05 split-K recurrent product +
 * a cell kernel that sums the split-K partials) issued by a C loop instead of three per step from the host language.
 *   fwd: hs_ext / cs_ext are (T+1, N, H) time-major with [0] = h_0 / c_0 on entry, [t+1] written; pre is (T, N, 4H) =
 *        x_t Wx + b (gate order f | g | i | o, layers.py:636-639) and is scratch afterwards; gates (T, N, 4H) receives the
 *        activated gates the backward needs; wh is (H, 4H) row-major.
 *   bwd: dhs is element (n, t, :) at dhs + n*stride_n + t*stride_t (the (N, T, H) upstream gradient); dA (T, N, 4H) and
 *        dh0 (N, H) = dL/dh_0 are written; dc_scratch: 2*N*H floats.
 *   workspace: b200ad_lstm_seq_workspace_bytes(N, H) bytes.  b200ad_lstm_seq_supported says whether the shape runs on
 *   this path (otherwise the caller loops over b200ad_gemm + b200ad_lstm_cell_*). */
int b200ad_lstm_seq_supported(int64_t batch, int64_t hidden);
int64_t b200ad_lstm_seq_workspace_bytes(int64_t batch, int64_t hidden);
int b200ad_lstm_seq_fwd(float* hs_ext, float* cs_ext, float* gates, const float* pre, const float* wh, int64_t T,
                        int64_t N, int64_t H, void* workspace, int64_t workspace_bytes, void* stream);
int b200ad_lstm_seq_bwd(const float* dhs, int64_t dhs_stride_n, int64_t dhs_stride_t, const float* hs_ext,
                        const float* cs_ext, const float* gates, const float* wh, float* dA, float* dh0,
                        float* dc_scratch, int64_t T, int64_t N, int64_t H, void* workspace, int64_t workspace_bytes,
                        void* stream);
int b200ad_bias_act_fwd(const float* x, int64_t ldx, const float* bias, float* out, int64_t ldo,
                        int64_t rows, int64_t cols, int act, void* stream);
/* gx = g * act'(y) (act: 1 tanh, 2 sigmoid, 3 relu on y) with row strides */
int b200ad_act_bwd(const float* g, int64_t ldg, const float* y, int64_t ldy, float* gx, int64_t ldgx,
                   int64_t rows, int64_t cols, int act, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimizers: ONE launch updates every parameter tensor (multi-tensor, chunked).
 *   table: device array of n_tensors entries {param*, grad*, m*, v*, numel}; built by the host once.
 *   Adam   m += (1-b1)(g-m); v += (1-b2)(g*g-v); p -= lr_t*m/(sqrt(v)+1e-7)
 *          utils/optimizer.py:14-29 == layers/layers.py:59-73 (lr_t computed by the host :22-23;
 *          NOT torch-Adam: eps is added to the un-corrected sqrt(v)).
 *   SGD    p -= lr*g                                layers/layers.py:829-831
 *   grad_scale multiplies g on the fly (1/world_size after a sum all-reduce, or the clip rate).
 *   clip:  rate = max_norm/(sqrt(sum g^2)+1e-6); grads *= rate if rate < 1    layers.py:132-141
 * ------------------------------------------------------------------------------------------ */
typedef struct b200ad_tensor_entry {
  float* param;
  float* grad;
  float* m;
  float* v;
  int64_t numel;
} b200ad_tensor_entry;
int b200ad_adam_multi(const b200ad_tensor_entry* table_dev, int n_tensors, int64_t total_chunks,
                      const int32_t* chunk_tensor_dev, const int32_t* chunk_index_dev, float lr_t,
                      float beta1, float beta2, float eps, float grad_scale,
                      const float* lr_t_dev /* NULL: use lr_t; else read the step size from the device */,
                      void* stream);
/* CUDA-graph friendly step scalar: iter_dev[0] += 1; lr_t_dev[0] = lr*sqrt(1-beta2^iter)/(1-beta1^iter) in fp64
 * (utils/optimizer.py:21-23 evaluated on the device). */
int b200ad_adam_step_scalar(int64_t* iter_dev, double lr, double beta1, double beta2, float* lr_t_dev,
                            void* stream);
int b200ad_sgd_multi(const b200ad_tensor_entry* table_dev, int n_tensors, int64_t total_chunks,
                     const int32_t* chunk_tensor_dev, const int32_t* chunk_index_dev, float lr,
                     float grad_scale, void* stream);
/* norm_out[0] = sum over all tensors of g^2 (deterministic two-stage); ws >= total_chunks floats */
int b200ad_sumsq_multi(const b200ad_tensor_entry* table_dev, int n_tensors, int64_t total_chunks,
                       const int32_t* chunk_tensor_dev, const int32_t* chunk_index_dev, float* ws,
                       float* norm_out, void* stream);
/* grads *= rate where rate = max_norm/(sqrt(norm_sq[0])+1e-6) if that is < 1 (decided on device) */
int b200ad_clip_scale_multi(const b200ad_tensor_entry* table_dev, int n_tensors, int64_t total_chunks,
                            const int32_t* chunk_tensor_dev, const int32_t* chunk_index_dev,
                            const float* norm_sq, float max_norm, void* stream);
#define B200AD_OPT_CHUNK 4096 /* elements per chunk in the multi-tensor tables */

#ifdef __cplusplus
}
#endif
#endif /* B200AD_H */
