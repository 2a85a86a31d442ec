// LSTM time loops issued from ONE C call: the T dependent steps of TimeLSTM.forward (reference layers/layers.py:573-576,
// cell :624-646) and of its BPTT (:588-596, cell :648-676) as two launches per step — the tcgen05 split-K recurrent product
// and a cell kernel that sums the split-K partials itself — instead of three launches per step driven from Python.
//
// Why not one persistent kernel like the tanh cell (rnn_seq.cu): W_h is (H, 4H) = 16 MB at H = 1024, 32 MB as tf32 hi/lo
// pairs; the co-resident CTAs of one launch hold at most 148 x 224 KB of tensor memory, and the backward product
// (K = 4H) does not tile onto 148 SMs at all with tf32 pairs.  What this loop removes is the host side (the
// reference-API time loop cost 19 ms per step eagerly against 12.9 ms of device time) and one third of the launches.
#include "common.cuh"
#include <stdlib.h>
#include "gemm_internal.cuh"

namespace b200ad {

// Every launch of the loops is a programmatic dependent of the one before: its set-up (barriers, tensor memory, descriptor
// prefetch, block scheduling) overlaps the predecessor's tail, and it waits for the predecessor's data itself.
// B200AD_LSTM_NO_PDL=1 turns that off (A/B measurements).
static bool lstm_pdl() {
  static const bool v = getenv("B200AD_LSTM_NO_PDL") == nullptr;
  return v;
}

static GemmArgs lstm_step_gemm(const float* a, int64_t lda, const float* b, int64_t ldb, float* c, int64_t ldc, int64_t m,
                               int64_t n, int64_t k, int trans_b, float beta) {
  GemmArgs g{};
  g.a = a; g.b = b; g.c = c; g.bias = nullptr;
  g.m = m; g.n = n; g.k = k;
  g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  g.batch0 = g.batch1 = 1;
  g.alpha = 1.f; g.beta = beta;
  g.trans_a = 0; g.trans_b = trans_b; g.act = 0;
  g.pdl = lstm_pdl() ? 1 : 0;
  return g;
}

// forward: A_t (N x 4H) += h_{t-1} (N x H) Wh (H x 4H);  backward: dh (N x H) = dA_t (N x 4H) Wh^T
static GemmArgs lstm_fwd_gemm(const float* h_prev, const float* wh, float* a_t, int64_t N, int64_t H) {
  return lstm_step_gemm(h_prev, H, wh, 4 * H, a_t, 4 * H, N, 4 * H, H, 0, 1.f);
}
static GemmArgs lstm_bwd_gemm(const float* da_t, const float* wh, float* dh, int64_t N, int64_t H) {
  return lstm_step_gemm(da_t, 4 * H, wh, 4 * H, dh, H, N, H, 4 * H, 1, 0.f);
}

}  // namespace b200ad

using namespace b200ad;

extern "C" {

int b200ad_lstm_seq_supported(int64_t batch, int64_t hidden) {
  if (batch < 1 || hidden < 4 || (hidden & 3)) return 0;
  // placeholder operands with the alignment the real ones must have (checked again at launch)
  const float* aligned = reinterpret_cast<const float*>(uintptr_t(256));
  const char* why = nullptr;
  GemmArgs f = lstm_fwd_gemm(aligned, aligned, const_cast<float*>(aligned), batch, hidden);
  GemmArgs b = lstm_bwd_gemm(aligned, aligned, const_cast<float*>(aligned), batch, hidden);
  return gemm_tc_supported(f, B200AD_GEMM_TF32X3, &why) && gemm_tc_supported(b, B200AD_GEMM_TF32X3, &why) ? 1 : 0;
}

int64_t b200ad_lstm_seq_workspace_bytes(int64_t batch, int64_t hidden) {
  if (!b200ad_lstm_seq_supported(batch, hidden)) return 0;
  const float* aligned = reinterpret_cast<const float*>(uintptr_t(256));
  const int64_t f = gemm_tc_workspace_bytes(lstm_fwd_gemm(aligned, aligned, const_cast<float*>(aligned), batch, hidden), B200AD_GEMM_TF32X3);
  const int64_t b = gemm_tc_workspace_bytes(lstm_bwd_gemm(aligned, aligned, const_cast<float*>(aligned), batch, hidden), B200AD_GEMM_TF32X3);
  return f > b ? f : b;
}

int b200ad_lstm_seq_fwd(float* hs_ext, float* cs_ext, float* gates, const float* pre, const float* wh, int64_t T, int64_t N,
                        int64_t H, void* workspace, int64_t workspace_bytes, void* stream) {
  B200AD_REQUIRE(T >= 0 && N >= 1 && H >= 1, "lstm_seq_fwd: bad shape");
  B200AD_REQUIRE(hs_ext && cs_ext && gates && pre && wh, "lstm_seq_fwd: null pointer");
  if (!b200ad_lstm_seq_supported(N, H)) return set_error(B200AD_ERR_UNSUPPORTED, "lstm_seq_fwd: shape not on the tcgen05 path");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t nh = N * H;
  for (int64_t t = 0; t < T; ++t) {
    float* a_t = const_cast<float*>(pre) + t * 4 * nh;
    GemmArgs g = lstm_fwd_gemm(hs_ext + t * nh, wh, a_t, N, H);
    const char* why = "unsupported";
    if (!gemm_tc_supported(g, B200AD_GEMM_TF32X3, &why)) return set_error(B200AD_ERR_UNSUPPORTED, "lstm_seq_fwd: %s", why);
    TcSplitInfo sp;
    int rc = gemm_tc_launch(g, B200AD_GEMM_TF32X3, workspace, workspace_bytes, s, &sp);
    if (rc) return rc;
    // splits > 1: `pre` is left untouched and the cell sums x_t Wx + b + partials; otherwise the product already sits in a_t
    rc = lstm_cell_fwd_launch(a_t, 4 * H, sp.splits > 1 ? sp.partial : nullptr, sp.splits, sp.stride, cs_ext + t * nh, H,
                              gates + t * 4 * nh, 4 * H, cs_ext + (t + 1) * nh, H, hs_ext + (t + 1) * nh, H, N, H, s, lstm_pdl());
    if (rc) return rc;
  }
  return B200AD_OK;
}

int b200ad_lstm_seq_bwd(const float* dhs, int64_t dhs_stride_n, int64_t dhs_stride_t, const float* hs_ext, const float* cs_ext,
                        const float* gates, const float* wh, float* da, float* dh0, float* dc_scratch, int64_t T, int64_t N,
                        int64_t H, void* workspace, int64_t workspace_bytes, void* stream) {
  (void)hs_ext;
  B200AD_REQUIRE(T >= 1 && N >= 1 && H >= 1, "lstm_seq_bwd: bad shape");
  B200AD_REQUIRE(dhs && cs_ext && gates && wh && da && dh0 && dc_scratch, "lstm_seq_bwd: null pointer");
  B200AD_REQUIRE(dhs_stride_n >= H || N == 1, "lstm_seq_bwd: bad row stride");
  if (!b200ad_lstm_seq_supported(N, H)) return set_error(B200AD_ERR_UNSUPPORTED, "lstm_seq_bwd: shape not on the tcgen05 path");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t nh = N * H;
  TcSplitInfo sp{nullptr, 1, 0};
  bool have_dh = false;          // dh of the step after t: in split-K partial form (sp) or, un-split, in dh0
  for (int64_t t = T - 1; t >= 0; --t) {
    float* dc_out = dc_scratch + (t & 1) * nh;
    const float* dc_in = (t == T - 1) ? nullptr : dc_scratch + ((t + 1) & 1) * nh;
    const bool partial = have_dh && sp.splits > 1;
    int rc = lstm_cell_bwd_launch(dhs + t * dhs_stride_t, dhs_stride_n, (have_dh && !partial) ? dh0 : nullptr, H,
                                  partial ? sp.partial : nullptr, sp.splits, sp.stride, dc_in, gates + t * 4 * nh, 4 * H,
                                  cs_ext + t * nh, H, cs_ext + (t + 1) * nh, H, da + t * 4 * nh, 4 * H, dc_out, N, H, s, lstm_pdl());
    if (rc) return rc;
    GemmArgs g = lstm_bwd_gemm(da + t * 4 * nh, wh, dh0, N, H);
    const char* why = "unsupported";
    if (!gemm_tc_supported(g, B200AD_GEMM_TF32X3, &why)) return set_error(B200AD_ERR_UNSUPPORTED, "lstm_seq_bwd: %s", why);
    // the last product (t = 0) is reduced into dh0 = dL/dh_0; the others stay in partial form for the next cell kernel
    rc = gemm_tc_launch(g, B200AD_GEMM_TF32X3, workspace, workspace_bytes, s, t > 0 ? &sp : nullptr);
    if (rc) return rc;
    have_dh = true;
  }
  return B200AD_OK;
}

}  // extern "C"
