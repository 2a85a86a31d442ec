// Internal GEMM plumbing shared by the SIMT and tcgen05 paths.
#pragma once
#include "common.cuh"

namespace b200ad {

struct GemmArgs {
  const float* a;
  const float* b;
  float* c;
  const float* bias;
  int64_t m, n, k;
  int64_t lda, ldb, ldc;
  int64_t batch0, batch1;
  int64_t a_stride0, a_stride1, b_stride0, b_stride1, c_stride0, c_stride1;
  float alpha, beta;
  int trans_a, trans_b, act;
  // programmatic dependent launch (the LSTM time loop, lstm_seq.cu): the kernel is launched with
  // cudaLaunchAttributeProgrammaticStreamSerialization, sets itself up while its predecessor drains and waits
  // (griddepcontrol.wait) before it touches global memory.  Only the 1-CTA tcgen05 kernels honour it.
  int pdl = 0;
};

int gemm_simt_launch(const GemmArgs& p, cudaStream_t s);

// tcgen05 / TMA path (gemm_tc.cu)
bool gemm_tc_supported(const GemmArgs& p, int precision, const char** why);
int64_t gemm_tc_workspace_bytes(const GemmArgs& p, int precision);
void gemm_tc_reload_env();      // re-read the B200AD_TC_* knobs (cached at first use)
// `split_out` non-null: a split-K launch stops after the partial products ([split][M][N] fp32 in `ws`) and reports them, so
// that the caller's next kernel sums them itself (the LSTM time loop folds the sum into its cell kernels, lstm_seq.cu);
// splits == 1 means C was written (with alpha / beta / bias / act) as usual.
struct TcSplitInfo {
  const float* partial;
  int splits;
  int64_t stride;      // elements between two splits' partial matrices (= batches * M * N)
};
int gemm_tc_launch(const GemmArgs& p, int precision, void* ws, int64_t ws_bytes, cudaStream_t s, TcSplitInfo* split_out = nullptr);

// LSTM cell kernels with an optional split-K partial input (cells.cu): the forward adds sum_s part[s] to A, the backward
// takes sum_s part[s] as the recurrent gradient dh_b.
int lstm_cell_fwd_launch(const float* a, int64_t lda, const float* part, int splits, int64_t part_stride, const float* c_prev,
                         int64_t ldc_prev, float* gates, int64_t ldg, float* c_next, int64_t ldc_next, float* h_next, int64_t ldh,
                         int64_t n, int64_t h, cudaStream_t s, bool pdl = false);
int lstm_cell_bwd_launch(const float* dh_a, int64_t ld_dha, const float* dh_b, int64_t ld_dhb, const float* part, int splits,
                         int64_t part_stride, const float* dc_next, const float* gates, int64_t ldg, const float* c_prev,
                         int64_t ldc_prev, const float* c_next, int64_t ldc_next, float* da, int64_t ldda, float* dc_prev,
                         int64_t n, int64_t h, cudaStream_t s, bool pdl = false);

// skinny fp32 kernel for the recurrent time steps, M <= 64 (gemm_skinny.cu)
bool gemm_skinny_supported(const GemmArgs& p, bool host_bound);
int gemm_skinny_launch(const GemmArgs& p, cudaStream_t s);

// bf16 variant (gemm_bf16.cu): conversion pre-pass + cta_group::2 kind::f16 kernel
bool gemm_bf16_supported(const GemmArgs& p, const char** why);
int64_t gemm_bf16_workspace_bytes(const GemmArgs& p);
int gemm_bf16_launch(const GemmArgs& p, void* ws, int64_t ws_bytes, cudaStream_t s);

// fp16 hi/lo three-product variant (gemm_h16.cu): operand split pre-pass + pure TMA / MMA cta_group::2 kernel
bool gemm_h16_supported(const GemmArgs& p, const char** why);
int64_t gemm_h16_workspace_bytes(const GemmArgs& p);
int gemm_h16_launch(const GemmArgs& p, void* ws, int64_t ws_bytes, cudaStream_t s);
int gemm_h16_shared_a_launch(const GemmArgs& p, int count, const float* const* bs, float* const* cs, const float* const* biases,
                             void* ws, int64_t ws_bytes, cudaStream_t s);

}  // namespace b200ad
