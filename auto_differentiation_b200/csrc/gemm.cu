// Public GEMM entry: validation + dispatch between the fp32 SIMT kernel and the tcgen05/TMA kernels.
// Reference calls replaced: np.matmul Matmul.result tgrad.py:535-539 / get_grad :530-533, tensordot's
// GEMMs :304,:222-228, np.dot TimeAffine layers.py:280,307,308, LSTM :632,:666-669, FeedForward
// :1252,:1263,:1266, MultiHeadAttention :1347-1360.
#include "common.cuh"
#include "gemm_internal.cuh"
#include <atomic>
#include <stdlib.h>

using namespace b200ad;

static int fill_args(const b200ad_gemm_desc* d, const float* a, const float* b, float* c, const float* bias,
                     GemmArgs* g) {
  B200AD_REQUIRE(d, "gemm: null descriptor");
  B200AD_REQUIRE(d->m >= 0 && d->n >= 0 && d->k >= 0, "gemm: negative dimension");
  B200AD_REQUIRE(d->batch0 >= 1 && d->batch1 >= 1, "gemm: batch extents must be >= 1");
  B200AD_REQUIRE(d->act >= 0 && d->act <= 3, "gemm: bad activation %d", d->act);
  B200AD_REQUIRE(d->precision >= 0 && d->precision <= 5, "gemm: bad precision %d", d->precision);
  const int64_t a_inner = d->trans_a ? d->m : d->k, b_inner = d->trans_b ? d->k : d->n;
  B200AD_REQUIRE(d->lda >= a_inner && d->ldb >= b_inner && d->ldc >= d->n,
                 "gemm: leading dimension smaller than the row (lda=%lld ldb=%lld ldc=%lld)", (long long)d->lda,
                 (long long)d->ldb, (long long)d->ldc);
  g->a = a; g->b = b; g->c = c; g->bias = bias;
  g->m = d->m; g->n = d->n; g->k = d->k;
  g->lda = d->lda; g->ldb = d->ldb; g->ldc = d->ldc;
  g->batch0 = d->batch0; g->batch1 = d->batch1;
  g->a_stride0 = d->a_stride0; g->a_stride1 = d->a_stride1;
  g->b_stride0 = d->b_stride0; g->b_stride1 = d->b_stride1;
  g->c_stride0 = d->c_stride0; g->c_stride1 = d->c_stride1;
  g->alpha = d->alpha; g->beta = d->beta;
  g->trans_a = d->trans_a != 0; g->trans_b = d->trans_b != 0; g->act = d->act;
  return B200AD_OK;
}

// 0 = by capture state (default), 1 = device time decides (the captured behaviour), 2 = single launch (the eager behaviour)
static std::atomic<int> g_recurrent_dispatch{0};

// AUTO -> F16X3 (gemm_h16.cu) where a device-time model built from measurements (profiles/r2_gemm_h16_probe.log) says it
// beats the in-kernel 3xTF32 split by >= 10 %: the F16X3 kernel runs a 256 x 256 tile in max(1536 cycles per 64-deep k-block,
// ~11 k cycles of output stores), the operand split costs ~2 ps per element (row-major) or ~4 ps (transposing), generation 3
// sustains ~270 TFLOP/s.  In practice: large tiles with K >= ~2048, or one operand shared by several products.
// 0 = never, 1 = by the model (default), 2 = whenever the path supports the problem (tests).
static std::atomic<int> g_f16x3_auto{-1};
static int f16x3_auto_mode() {
  int v = g_f16x3_auto.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("B200AD_GEMM_F16X3_AUTO");
    v = e ? atoi(e) : 1;
    if (v < 0 || v > 2) v = 1;
    g_f16x3_auto.store(v, std::memory_order_relaxed);
  }
  return v;
}
static bool f16x3_auto_pick(const GemmArgs& g, int count = 1) {
  const int mode = f16x3_auto_mode();
  if (mode == 0 || g.m < 1 || g.n < 1 || g.k < 1) return false;
  const char* why = nullptr;
  if (!gemm_h16_supported(g, &why)) return false;
  if (mode == 2) return true;
  if (g.m < 512 || g.n < 256 || g.k < 128) return false;
  const double batches = (double)(g.batch0 * g.batch1);
  const double tiles = (double)(ceil_div(g.m, 256) * ceil_div(g.n, 256)) * batches;
  const double pairs = sm_count() / 2;
  // less than two waves of tiles: launch / ramp costs the model does not carry decide (config 3's M = 8192 projections
  // measured 39 against 28 us), unless K is long enough for split-K to fill the machine
  if (tiles < 2 * pairs && g.k < 2048) return false;
  const double waves = (double)ceil_div((int64_t)tiles, (int64_t)pairs);
  const double kb = (double)ceil_div(g.k, 64);
  // row-major single-matrix A: converted inside the kernel (measured ~2300 cycles per k-block instead of the 1536 of pure MMA
  // time), the pre-pass only reads it once for the row scales
  const bool a_raw = !g.trans_a && batches == 1 && (g.lda % 4) == 0 && g.n <= 2048;
  double cyc = kb * (a_raw ? 2300.0 : 1536.0);
  if (cyc < 11000.0) cyc = 11000.0;
  if (tiles < pairs && kb >= 16) cyc = cyc * tiles / pairs + 11000.0;      // split-K spreads the k-blocks over the idle pairs
  const double t_kernel = waves * cyc / 1.7e9 + 2e-6;
  const double t_a = (double)g.m * g.k * batches * (g.trans_a ? 4e-12 : (a_raw ? 0.7e-12 : 2e-12)) + 3e-6;
  const double t_b = (double)g.n * g.k * batches * (g.trans_b ? 2e-12 : 4e-12) + 3e-6;
  const double t_h16 = t_a + count * (t_b + t_kernel);
  const double t_tc = count * (2.0 * g.m * g.n * g.k * batches / 2.7e14 + 3e-6);
  return t_h16 < 0.9 * t_tc;
}
// what AUTO resolves to for the tensor-core shapes (the skinny recurrent kernels are decided at launch time)
static int auto_precision(const GemmArgs& g) {
  if (f16x3_auto_pick(g)) return B200AD_GEMM_F16X3;
  const char* why = nullptr;
  return gemm_tc_supported(g, B200AD_GEMM_TF32X3, &why) ? B200AD_GEMM_TF32X3 : B200AD_GEMM_FP32;
}

extern "C" {

int b200ad_gemm_reload_env(void) {
  gemm_tc_reload_env();
  return B200AD_OK;
}

int b200ad_gemm_set_recurrent_dispatch(int mode) {
  if (mode < 0 || mode > 2) return set_error(B200AD_ERR_INVALID, "gemm_set_recurrent_dispatch: mode %d not in 0..2", mode);
  return g_recurrent_dispatch.exchange(mode);
}

int b200ad_gemm_set_f16x3_auto(int mode) {
  const int prev = f16x3_auto_mode();
  if (mode < 0 || mode > 2) return set_error(B200AD_ERR_INVALID, "gemm_set_f16x3_auto: mode %d not in 0..2", mode);
  g_f16x3_auto.store(mode, std::memory_order_relaxed);
  return prev;
}

int64_t b200ad_gemm_shared_a_workspace_bytes(const b200ad_gemm_desc* d, int32_t count) {
  GemmArgs g;
  if (fill_args(d, nullptr, nullptr, nullptr, nullptr, &g)) return 0;
  if (d->precision == B200AD_GEMM_AUTO && count > 1 && f16x3_auto_pick(g, count)) return gemm_h16_workspace_bytes(g);
  return b200ad_gemm_workspace_bytes(d);
}

int b200ad_gemm_supported(const b200ad_gemm_desc* d, const float* a, const float* b, const float* c) {
  GemmArgs g;
  if (fill_args(d, a, b, const_cast<float*>(c), nullptr, &g)) return 0;
  if (d->precision == B200AD_GEMM_FP32 || d->precision == B200AD_GEMM_AUTO) return 1;
  const char* why = nullptr;
  if (d->precision == B200AD_GEMM_BF16) return gemm_bf16_supported(g, &why) ? 1 : 0;
  if (d->precision == B200AD_GEMM_F16X3) return gemm_h16_supported(g, &why) ? 1 : 0;
  return gemm_tc_supported(g, d->precision, &why) ? 1 : 0;
}

int64_t b200ad_gemm_workspace_bytes(const b200ad_gemm_desc* d) {
  GemmArgs g;
  if (fill_args(d, nullptr, nullptr, nullptr, nullptr, &g)) return 0;
  if (d->precision == B200AD_GEMM_FP32) return 0;
  if (d->precision == B200AD_GEMM_BF16) return gemm_bf16_workspace_bytes(g);
  if (d->precision == B200AD_GEMM_F16X3) return gemm_h16_workspace_bytes(g);
  if (d->precision == B200AD_GEMM_AUTO && g.k > 0 && f16x3_auto_pick(g)) return gemm_h16_workspace_bytes(g);
  if (d->precision == B200AD_GEMM_AUTO && gemm_skinny_supported(g, false)) {
    // skinny recurrent-step kernels need no workspace.  (Pointer alignment is part of the check; with the null pointers
    // of this query it passes, and a misaligned operand at launch time falls back to an un-split tensor-core launch.)
    return 0;
  }
  const int prec = d->precision == B200AD_GEMM_AUTO ? B200AD_GEMM_TF32X3 : d->precision;
  return gemm_tc_workspace_bytes(g, prec);
}

int b200ad_gemm(const b200ad_gemm_desc* d, const float* a, const float* b, float* c, const float* bias,
                void* workspace, int64_t workspace_bytes, void* stream) {
  GemmArgs g;
  int rc = fill_args(d, a, b, c, bias, &g);
  if (rc) return rc;
  if (g.m == 0 || g.n == 0) return B200AD_OK;
  B200AD_REQUIRE(c && (g.k == 0 || (a && b)), "gemm: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  int prec = d->precision;
  if (prec == B200AD_GEMM_AUTO && g.k > 0) {
    // inside a stream capture the host cost is paid once and device time decides; eagerly the recurrent step is host-bound.
    // The two branches differ by ~1e-6 relative (both inside the parity tolerance, tests/test_config_parity.py pins each);
    // b200ad_gemm_set_recurrent_dispatch(1 | 2) fixes one of them regardless of the capture state.
    const int mode = g_recurrent_dispatch.load(std::memory_order_relaxed);
    bool device_time = (mode == 1);
    if (mode == 0) {
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      device_time = cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusActive;
    }
    if (gemm_skinny_supported(g, !device_time)) return gemm_skinny_launch(g, s);
  }
  if (prec == B200AD_GEMM_AUTO) {
    prec = g.k > 0 ? auto_precision(g) : B200AD_GEMM_FP32;
    // a caller that sized its workspace before the F16X3 switch was flipped still gets a correct product
    if (prec == B200AD_GEMM_F16X3 && (!workspace || workspace_bytes < gemm_h16_workspace_bytes(g))) {
      const char* w2 = nullptr;
      prec = gemm_tc_supported(g, B200AD_GEMM_TF32X3, &w2) ? B200AD_GEMM_TF32X3 : B200AD_GEMM_FP32;
    }
  }
  if (prec == B200AD_GEMM_FP32 || g.k == 0) return gemm_simt_launch(g, s);
  const char* why = "unsupported";
  if (prec == B200AD_GEMM_F16X3) {
    if (!gemm_h16_supported(g, &why)) return set_error(B200AD_ERR_UNSUPPORTED, "gemm(f16x3): %s", why);
    return gemm_h16_launch(g, workspace, workspace_bytes, s);
  }
  if (prec == B200AD_GEMM_BF16) {
    if (!gemm_bf16_supported(g, &why)) return set_error(B200AD_ERR_UNSUPPORTED, "gemm(bf16): %s", why);
    return gemm_bf16_launch(g, workspace, workspace_bytes, s);
  }
  if (!gemm_tc_supported(g, prec, &why)) return set_error(B200AD_ERR_UNSUPPORTED, "gemm(tcgen05): %s", why);
  return gemm_tc_launch(g, prec, workspace, workspace_bytes, s);
}

int b200ad_gemm_shared_a(const b200ad_gemm_desc* d, const float* a, int32_t count, const float* const* bs, float* const* cs,
                         const float* const* biases, void* workspace, int64_t workspace_bytes, void* stream) {
  B200AD_REQUIRE(count >= 0 && (count == 0 || (bs && cs)), "gemm_shared_a: null operand list");
  GemmArgs g;
  int rc = fill_args(d, a, count ? bs[0] : nullptr, count ? cs[0] : nullptr, nullptr, &g);
  if (rc) return rc;
  int prec = d->precision;
  if (prec == B200AD_GEMM_AUTO && count > 0 && f16x3_auto_pick(g, count)) prec = B200AD_GEMM_F16X3;
  if (prec == B200AD_GEMM_F16X3 && g.m > 0 && g.n > 0 && g.k > 0 && workspace && workspace_bytes >= gemm_h16_workspace_bytes(g)) {
    const char* why = "unsupported";
    if (!gemm_h16_supported(g, &why)) return set_error(B200AD_ERR_UNSUPPORTED, "gemm_shared_a(f16x3): %s", why);
    for (int i = 0; i < count; ++i) B200AD_REQUIRE(bs[i] && cs[i] && a, "gemm_shared_a: null pointer");
    return gemm_h16_shared_a_launch(g, count, bs, cs, biases, workspace, workspace_bytes, (cudaStream_t)stream);
  }
  for (int i = 0; i < count; ++i)
    if ((rc = b200ad_gemm(d, a, bs[i], cs[i], biases ? biases[i] : nullptr, workspace, workspace_bytes, stream))) return rc;
  return B200AD_OK;
}
}
