// Skinny GEMM for the recurrent time steps: C (M x N) = act(alpha * A (M x K) @ op(B) + beta * C + bias), M <= 64.
// Replaces np.dot(h_prev, Wh) in RNN.forward layers.py:171 / LSTM.forward :632 and np.dot(dA, Wh.T) in their
// backward (:180, :667) — 2 x T strictly sequential products of a 64-row state with a resident weight matrix.
// These are latency-bound: a 134 MFLOP product cannot fill 128-row tensor-core tiles, and the tcgen05 path spent
// ~48 us per step (split-K + reduce, two launches).  Here one launch spreads the N columns over the whole chip
// (one CTA per 8 columns: 128 CTAs for H = 1024, 512 for the LSTM's 4H), the 8 warps of a CTA split K, every lane
// owns 2 rows x 8 columns in registers, and the epilogue (beta * C, bias, tanh / sigmoid / relu) is fused, so a tanh
// RNN step is ONE launch.  Exact fp32 FMA accumulation, deterministic reduction order.
#include "common.cuh"
#include "gemm_internal.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>

namespace b200ad {

constexpr int SK_BN = 8;
constexpr int SK_WARPS = 8;
constexpr int SK_THREADS = SK_WARPS * 32;

template <int ACT>
__device__ __forceinline__ float sk_act(float x) {
  if constexpr (ACT == 1) return tanhf(x);
  else if constexpr (ACT == 2) return 1.f / (1.f + expf(-x));
  else if constexpr (ACT == 3) return (x > 0.f || x != x) ? x : 0.f;
  else return x;
}

template <bool TRANS_B>
__global__ void __launch_bounds__(SK_THREADS) gemm_skinny_kernel(const GemmArgs g) {
  __shared__ __align__(16) float bsm[SK_WARPS][32][SK_BN];     // per warp: a 32 (k) x 8 (n) chunk of op(B)
  __shared__ __align__(16) float red[SK_WARPS][64][SK_BN];     // per warp partial sums
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n0 = (int64_t)blockIdx.x * SK_BN;
  const int M = (int)g.m;
  const int64_t K = g.k;
  const int chunks = (int)(K / 32);
  const int per_warp = (chunks + SK_WARPS - 1) / SK_WARPS;
  const int c_begin = warp * per_warp, c_end = min(chunks, c_begin + per_warp);
  const int r0 = 2 * lane, r1 = r0 + 1;
  const float* a0p = g.a + (int64_t)(r0 < M ? r0 : 0) * g.lda;
  const float* a1p = g.a + (int64_t)(r1 < M ? r1 : 0) * g.lda;
  float acc0[SK_BN], acc1[SK_BN];
#pragma unroll
  for (int j = 0; j < SK_BN; ++j) acc0[j] = acc1[j] = 0.f;

  for (int c = c_begin; c < c_end; ++c) {
    const int64_t k0 = (int64_t)c * 32;
    // stage the B chunk: bsm[warp][k][n] = op(B)[k0 + k][n0 + n]
    if (!TRANS_B) {            // B stored (K x N), n contiguous: lane = k row, two float4
      const float4* src = reinterpret_cast<const float4*>(g.b + (k0 + lane) * g.ldb + n0);
      const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
      *reinterpret_cast<float4*>(&bsm[warp][lane][0]) = v0;
      *reinterpret_cast<float4*>(&bsm[warp][lane][4]) = v1;
    } else {                   // B stored (N x K), k contiguous: lane -> (n = lane >> 2, 8 consecutive k)
      const int n = lane >> 2, kq = (lane & 3) * 8;
      const float4* src = reinterpret_cast<const float4*>(g.b + (n0 + n) * g.ldb + k0 + kq);
      const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
      bsm[warp][kq + 0][n] = v0.x; bsm[warp][kq + 1][n] = v0.y; bsm[warp][kq + 2][n] = v0.z; bsm[warp][kq + 3][n] = v0.w;
      bsm[warp][kq + 4][n] = v1.x; bsm[warp][kq + 5][n] = v1.y; bsm[warp][kq + 6][n] = v1.z; bsm[warp][kq + 7][n] = v1.w;
    }
    // this lane's two A rows, 32 k values each
    float a0[32], a1[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(a0p + k0) + q);
      const float4 w = __ldg(reinterpret_cast<const float4*>(a1p + k0) + q);
      a0[4 * q] = u.x; a0[4 * q + 1] = u.y; a0[4 * q + 2] = u.z; a0[4 * q + 3] = u.w;
      a1[4 * q] = w.x; a1[4 * q + 1] = w.y; a1[4 * q + 2] = w.z; a1[4 * q + 3] = w.w;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float4 b0 = *reinterpret_cast<const float4*>(&bsm[warp][k][0]);     // broadcast reads
      const float4 b1 = *reinterpret_cast<const float4*>(&bsm[warp][k][4]);
      const float bb[SK_BN] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < SK_BN; ++j) {
        acc0[j] = fmaf(a0[k], bb[j], acc0[j]);
        acc1[j] = fmaf(a1[k], bb[j], acc1[j]);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int j = 0; j < SK_BN; ++j) {
    red[warp][r0][j] = acc0[j];
    red[warp][r1][j] = acc1[j];
  }
  __syncthreads();
  // 512 outputs, 2 per thread, warps summed in a fixed order
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    const int idx = threadIdx.x + o * SK_THREADS;
    const int row = idx >> 3, col = idx & 7;
    if (row < M) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < SK_WARPS; ++w) s += red[w][row][col];
      float* dst = g.c + (int64_t)row * g.ldc + n0 + col;
      float v = g.alpha * s;
      if (g.beta != 0.f) v += g.beta * *dst;
      if (g.bias) v += __ldg(g.bias + n0 + col);
      if (g.act == 1) v = sk_act<1>(v);
      else if (g.act == 2) v = sk_act<2>(v);
      else if (g.act == 3) v = sk_act<3>(v);
      *dst = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Cluster variant: a 2-D decomposition (64 output columns x one eighth of K per CTA) with the eight K-partials of a
// column group reduced through DISTRIBUTED SHARED MEMORY inside a thread-block cluster of 8.  Every operand element
// is read from L2 exactly once per column group (A: N/64 x, B: once), where the kernel above re-reads the whole A
// panel in every CTA; the reduction order over the cluster ranks is fixed (deterministic).
//   grid = 8 * (N / 64) CTAs, cluster (8,1,1): rank = K-slice, blockIdx.x / 8 = column group.  256 threads: 4 x 4 outputs each.
// ---------------------------------------------------------------------------------------------------------
constexpr int SC_BN = 64, SC_KC = 32, SC_CL = 8;

template <bool TRANS_B>
__global__ void __cluster_dims__(SC_CL, 1, 1) __launch_bounds__(256) gemm_skinny_cluster_kernel(const GemmArgs g) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ __align__(16) float As[SC_KC][64];          // [k][row]
  __shared__ __align__(16) float Bs[SC_KC][SC_BN];       // [k][col]
  __shared__ __align__(16) float part[64][SC_BN];        // this CTA's K-partial of the 64 x 64 output block
  const int tid = threadIdx.x;
  const int rank = (int)cluster.block_rank();
  const int64_t n0 = (int64_t)(blockIdx.x / SC_CL) * SC_BN;
  const int M = (int)g.m;
  const int64_t kslice = g.k / SC_CL, kbeg = (int64_t)rank * kslice;
  const int ty = tid >> 4, tx = tid & 15;                 // 16 x 16 threads, 4 rows x 4 columns each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // The K-slice is walked in groups of up to four 32-deep chunks.  ALL global loads of a group are issued first (16
  // independent float4 per thread), so one L2 round trip is exposed per group instead of one per chunk: at K = 1024
  // (k-slice 128 = one group) the recurrent step pays a single load latency.
  const int64_t kend = kbeg + kslice;
  for (int64_t kg = kbeg; kg < kend; kg += 4 * SC_KC) {
    float4 ra[4][2], rb[4][2];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int64_t k0 = kg + c * SC_KC;
      if (k0 < kend) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int v = tid + i * 256;
          {   // A chunk: 64 rows x 32 k (row-major source, k contiguous): 512 float4, 2 per thread
            const int row = v >> 3, kq = (v & 7) * 4;
            ra[c][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < M) ra[c][i] = __ldg(reinterpret_cast<const float4*>(g.a + (int64_t)row * g.lda + k0 + kq));
          }
          if (!TRANS_B) {   // B stored (K x N), columns contiguous
            const int kk = v >> 4, cq = (v & 15) * 4;
            rb[c][i] = __ldg(reinterpret_cast<const float4*>(g.b + (k0 + kk) * g.ldb + n0 + cq));
          } else {          // B stored (N x K), k contiguous
            const int col = v >> 3, kq = (v & 7) * 4;
            rb[c][i] = __ldg(reinterpret_cast<const float4*>(g.b + (n0 + col) * g.ldb + k0 + kq));
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (kg + c * SC_KC < kend) {        // block-uniform
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int v = tid + i * 256;
          {
            const int row = v >> 3, kq = (v & 7) * 4;
            const float4 a4 = ra[c][i];
            As[kq][row] = a4.x; As[kq + 1][row] = a4.y; As[kq + 2][row] = a4.z; As[kq + 3][row] = a4.w;
          }
          if (!TRANS_B) {
            const int kk = v >> 4, cq = (v & 15) * 4;
            *reinterpret_cast<float4*>(&Bs[kk][cq]) = rb[c][i];
          } else {
            const int col = v >> 3, kq = (v & 7) * 4;
            const float4 b4 = rb[c][i];
            Bs[kq][col] = b4.x; Bs[kq + 1][col] = b4.y; Bs[kq + 2][col] = b4.z; Bs[kq + 3][col] = b4.w;
          }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SC_KC; ++kk) {
          const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
          const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
          const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(&part[ty * 4 + i][tx * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  // The partials live in SHARED memory, so CTA-scope completion of the stores is enough before the cluster barrier;
  // `cluster.sync()` (arrive.release) would prepend MEMBAR.ALL.GPU + ERRBAR, ~1000 cycles of this ~5 us kernel.
  __threadfence_block();
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // every CTA's partial is visible cluster-wide
  // CTA `rank` reduces rows [8 rank, 8 rank + 8) x 64 columns over the 8 partials (fixed order) and applies the epilogue
  {
    const int row = rank * 8 + (tid >> 5), col = (tid & 31) * 2;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int r = 0; r < SC_CL; ++r) {
      const float* peer = cluster.map_shared_rank(&part[0][0], r);
      const float2 v = *reinterpret_cast<const float2*>(peer + row * SC_BN + col);
      s0 += v.x;
      s1 += v.y;
    }
    if (row < M) {
      float* dst = g.c + (int64_t)row * g.ldc + n0 + col;
      float v[2] = {g.alpha * s0, g.alpha * s1};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (g.beta != 0.f) v[j] += g.beta * dst[j];
        if (g.bias) v[j] += __ldg(g.bias + n0 + col + j);
        if (g.act == 1) v[j] = sk_act<1>(v[j]);
        else if (g.act == 2) v[j] = sk_act<2>(v[j]);
        else if (g.act == 3) v[j] = sk_act<3>(v[j]);
        dst[j] = v[j];
      }
    }
  }
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");   // peers may still be reading this CTA's partial
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Measured crossover against the tcgen05 path with 64-column tiles + split-K (profiles/recurrent_shapes_probe.py, us per
// dependent step in a CUDA graph, M = 64): 256 x 256: 3.8 vs 16.0, 512 x 512: 5.1 vs 9.2, 1024 x 256: 5.5 vs 16.0,
// 1024 x 1024: 12.7 vs 9.6, 2048 x 2048: 23.2 vs 18.1 -> the exact-fp32 kernels keep the small weight matrices when DEVICE
// time decides (the launch stream is being captured into a CUDA graph).  Launched eagerly from Python the same step is
// host-bound (one kernel here against two kernels + two tensor-map encodes there: config 4 measured 15.5 vs 21.3 ms per
// step), so outside a capture the single-launch kernels keep every shape up to N, K <= 2048.
constexpr int64_t SKINNY_MAX_NK = 512 * 1024;

// environment switches are read once (the dispatcher runs per GEMM call)
static bool env_no_skinny() { static const bool v = getenv("B200AD_NO_SKINNY") != nullptr; return v; }
static bool env_no_skinny_cluster() { static const bool v = getenv("B200AD_NO_SKINNY_CLUSTER") != nullptr; return v; }

static bool skinny_cluster_ok(const GemmArgs& g, bool host_bound) {
  if (env_no_skinny_cluster()) return false;
  // measured crossover against the tensor-core path (profiles/skinny_probe.py): wins for the H = 1024 tanh-RNN step
  // (one wave of 128 CTAs), loses beyond ~2 waves or K > 2048 (the LSTM's 4H-wide / 4H-deep products)
  if (!host_bound && g.n * g.k > SKINNY_MAX_NK) return false;
  return g.m >= 1 && g.m <= 64 && g.batch0 * g.batch1 == 1 && !g.trans_a && g.k >= 256 && g.k <= 2048 &&
         (g.k % (SC_CL * SC_KC)) == 0 && g.n >= 64 && g.n <= 2048 && (g.n % SC_BN) == 0 && (g.lda & 3) == 0 && (g.ldb & 3) == 0 && aligned16(g.a) && aligned16(g.b);
}

bool gemm_skinny_supported(const GemmArgs& g, bool host_bound) {
  if (env_no_skinny()) return false;
  if (skinny_cluster_ok(g, host_bound)) return true;
  if (g.m < 1 || g.m > 64 || g.batch0 * g.batch1 != 1 || g.trans_a || (!host_bound && g.n * g.k > SKINNY_MAX_NK)) return false;
  // N <= 2048: every CTA re-reads the whole A panel (M x K) from L2, which stops paying beyond ~256 CTAs (the LSTM's
  // 4H = 4096 columns run faster on the tensor-core path: measured 14.7 vs 30.7 ms per graph-captured step)
  // ... and K <= 2048: the A panel (M x K) is re-read by every CTA (measured: 64 x 1024 x 4096 takes 52 us here, 23 us on
  // the tensor-core path; 64 x 1024 x 1024: 17 us here (host-bound), 37 us there)
  if (g.k < 256 || g.k > 2048 || (g.k & 31) || g.n < 256 || g.n > 2048 || (g.n & 7)) return false;
  if ((g.lda & 3) || (g.ldb & 3) || !aligned16(g.a) || !aligned16(g.b)) return false;
  return true;
}

int gemm_skinny_launch(const GemmArgs& g, cudaStream_t s) {
  if (!env_no_skinny() && skinny_cluster_ok(g, true)) {
    const int grid_c = (int)(g.n / SC_BN) * SC_CL;
    if (g.trans_b) gemm_skinny_cluster_kernel<true><<<grid_c, 256, 0, s>>>(g);
    else gemm_skinny_cluster_kernel<false><<<grid_c, 256, 0, s>>>(g);
    count_launch();
    return check_launch("gemm(skinny, cluster)");
  }
  const int grid = (int)(g.n / SK_BN);
  if (g.trans_b) gemm_skinny_kernel<true><<<grid, SK_THREADS, 0, s>>>(g);
  else gemm_skinny_kernel<false><<<grid, SK_THREADS, 0, s>>>(g);
  count_launch();
  return check_launch("gemm(skinny)");
}

}  // namespace b200ad
