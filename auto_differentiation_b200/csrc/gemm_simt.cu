// fp32 SIMT GEMM (exact fp32 FMA accumulate) — the any-shape / any-alignment contraction path and the
// bit-faithful fallback for shapes the TMA/tcgen05 path cannot take (ld not 16B aligned, tiny N).
//   C[b0,b1] = alpha * op(A) @ op(B) + beta * C  (+ bias[n], activation)   row-major, strided batches.
// Tiling: BMxBNx16 CTA tile, 256 threads, (TMxTN) register tile split in 4-wide groups so that the
// shared-memory reads are conflict-free float4; register-staged double buffering (global loads of
// tile k+1 are in flight while tile k is multiplied).
#include "common.cuh"
#include "gemm_internal.cuh"

namespace b200ad {

template <int ACT>
__device__ __forceinline__ float gemm_act(float x) {
  if constexpr (ACT == 1) return tanhf(x);
  else if constexpr (ACT == 2) return 1.f / (1.f + expf(-x));
  else if constexpr (ACT == 3) return (x > 0.f || x != x) ? x : 0.f;
  else return x;
}

template <int BM, int BN, int TM, int TN, bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs p) {
  constexpr int BK = 16;
  constexpr int RG = TM / 4, CG = TN / 4;  // 4-wide groups per thread
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  constexpr int LA = BM * BK / 256, LB = BN * BK / 256;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  const int64_t bz = blockIdx.z;
  const int64_t b0 = bz / p.batch1, b1 = bz - b0 * p.batch1;
  const float* __restrict__ A = p.a + b0 * p.a_stride0 + b1 * p.a_stride1;
  const float* __restrict__ B = p.b + b0 * p.b_stride0 + b1 * p.b_stride1;
  float* __restrict__ C = p.c + b0 * p.c_stride0 + b1 * p.c_stride1;
  const int64_t M = p.m, N = p.n, K = p.k;

  float ra[LA], rb[LB];
  auto load_tiles = [&](int64_t k0) {
#pragma unroll
    for (int j = 0; j < LA; ++j) {
      const int idx = tid + j * 256;
      int m, k;
      if (TA) { m = idx % BM; k = idx / BM; } else { k = idx % BK; m = idx / BK; }
      const int64_t gm = m0 + m, gk = k0 + k;
      ra[j] = (gm < M && gk < K) ? __ldg(A + (TA ? gk * p.lda + gm : gm * p.lda + gk)) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < LB; ++j) {
      const int idx = tid + j * 256;
      int n, k;
      if (TB) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
      const int64_t gn = n0 + n, gk = k0 + k;
      rb[j] = (gn < N && gk < K) ? __ldg(B + (TB ? gn * p.ldb + gk : gk * p.ldb + gn)) : 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int j = 0; j < LA; ++j) {
      const int idx = tid + j * 256;
      int m, k;
      if (TA) { m = idx % BM; k = idx / BM; } else { k = idx % BK; m = idx / BK; }
      As[buf][k][m] = ra[j];
    }
#pragma unroll
    for (int j = 0; j < LB; ++j) {
      const int idx = tid + j * 256;
      int n, k;
      if (TB) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
      Bs[buf][k][n] = rb[j];
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int64_t nk = (K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int64_t kt = 0; kt < nk; ++kt) {
    const int buf = (int)(kt & 1);
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      __align__(16) float a[TM];
      __align__(16) float b[TN];
#pragma unroll
      for (int g = 0; g < RG; ++g)
        *reinterpret_cast<float4*>(&a[g * 4]) = *reinterpret_cast<const float4*>(&As[buf][k][g * (BM / RG) + ty * 4]);
#pragma unroll
      for (int g = 0; g < CG; ++g)
        *reinterpret_cast<float4*>(&b[g * 4]) = *reinterpret_cast<const float4*>(&Bs[buf][k][g * (BN / CG) + tx * 4]);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int gi = 0; gi < RG; ++gi)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t gm = m0 + gi * (BM / RG) + ty * 4 + i;
      if (gm >= M) continue;
#pragma unroll
      for (int gj = 0; gj < CG; ++gj)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t gn = n0 + gj * (BN / CG) + tx * 4 + j;
          if (gn >= N) continue;
          float v = p.alpha * acc[gi * 4 + i][gj * 4 + j];
          if (p.beta != 0.f) v += p.beta * C[gm * p.ldc + gn];
          if (p.bias) v += __ldg(p.bias + gn);
          switch (p.act) {
            case 1: v = gemm_act<1>(v); break;
            case 2: v = gemm_act<2>(v); break;
            case 3: v = gemm_act<3>(v); break;
            default: break;
          }
          C[gm * p.ldc + gn] = v;
        }
    }
}

template <int BM, int BN, int TM, int TN>
static void launch_cfg(const GemmArgs& p, cudaStream_t s) {
  dim3 grid((unsigned)ceil_div(p.m, BM), (unsigned)ceil_div(p.n, BN), (unsigned)(p.batch0 * p.batch1));
  if (p.trans_a) {
    if (p.trans_b) sgemm_kernel<BM, BN, TM, TN, true, true><<<grid, 256, 0, s>>>(p);
    else sgemm_kernel<BM, BN, TM, TN, true, false><<<grid, 256, 0, s>>>(p);
  } else {
    if (p.trans_b) sgemm_kernel<BM, BN, TM, TN, false, true><<<grid, 256, 0, s>>>(p);
    else sgemm_kernel<BM, BN, TM, TN, false, false><<<grid, 256, 0, s>>>(p);
  }
}

int gemm_simt_launch(const GemmArgs& p, cudaStream_t s) {
  const int64_t batches = p.batch0 * p.batch1;
  if (ceil_div(p.n, 64) > 65535 || batches > 65535)
    return set_error(B200AD_ERR_UNSUPPORTED, "gemm(simt): grid too large (n=%lld, batches=%lld)", (long long)p.n,
                     (long long)batches);
  // 128x128 tiles when they give at least ~one wave of CTAs, else 64x64.
  const int64_t big_tiles = ceil_div(p.m, 128) * ceil_div(p.n, 128) * batches;
  if (p.m >= 96 && p.n >= 96 && big_tiles >= sm_count()) launch_cfg<128, 128, 8, 8>(p, s);
  else launch_cfg<64, 64, 4, 4>(p, s);
  count_launch();
  return check_launch("gemm(simt)");
}

}  // namespace b200ad
