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
};

int gemm_simt_launch(const GemmArgs& p, cudaStream_t s);

// tcgen05 / TMA path (gemm_tc.cu)
bool gemm_tc_supported(const GemmArgs& p, int precision, const char** why);
int64_t gemm_tc_workspace_bytes(const GemmArgs& p, int precision);
int gemm_tc_launch(const GemmArgs& p, int precision, void* ws, int64_t ws_bytes, cudaStream_t s);

// skinny fp32 kernel for the recurrent time steps, M <= 64 (gemm_skinny.cu)
bool gemm_skinny_supported(const GemmArgs& p, bool host_bound);
int gemm_skinny_launch(const GemmArgs& p, cudaStream_t s);

// bf16 variant (gemm_bf16.cu): conversion pre-pass + cta_group::2 kind::f16 kernel
bool gemm_bf16_supported(const GemmArgs& p, const char** why);
int64_t gemm_bf16_workspace_bytes(const GemmArgs& p);
int gemm_bf16_launch(const GemmArgs& p, void* ws, int64_t ws_bytes, cudaStream_t s);

}  // namespace b200ad
