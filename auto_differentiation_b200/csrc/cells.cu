// Elementwise halves of the recurrent cells and the affine epilogues, with row strides so that a time
// slice [:, t, :] of an (N, T, X) tensor is processed in place (no per-step copies).
// Reference bodies replaced: LSTM.forward layers.py:624-646, LSTM.backward :648-676 (gate part),
// RNN.forward :166-173 / backward :175-177, FeedForward/ReLU :1242-1290, TimeAffine bias add :280.
#include "common.cuh"
#include "gemm_internal.cuh"

namespace b200ad {

// V consecutive floats of a split-K partial matrix (16 B load when V = 4: the partial rows are 16 B aligned then)
template <int V>
__device__ __forceinline__ void add_partial(float (&acc)[V], const float* __restrict__ p) {
  if constexpr (V == 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
  } else {
#pragma unroll
    for (int u = 0; u < V; ++u) acc[u] += __ldg(p + u);
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <int ACT>
__device__ __forceinline__ float act_fwd(float x) {
  if constexpr (ACT == 1) return tanhf(x);
  else if constexpr (ACT == 2) return sigmoidf_(x);
  else if constexpr (ACT == 3) return (x > 0.f || x != x) ? x : 0.f;
  else return x;
}
template <int ACT>
__device__ __forceinline__ float act_bwd(float g, float y) {
  if constexpr (ACT == 1) return g * (1.f - y * y);
  else if constexpr (ACT == 2) return g * y * (1.f - y);
  else if constexpr (ACT == 3) return y > 0.f ? g : 0.f * g;
  else return g;
}

// One thread per (row, 4 hidden units) when VEC, else per (row, unit).
template <bool VEC>
__global__ void __launch_bounds__(256) lstm_cell_fwd_kernel(const float* __restrict__ a, int64_t lda,
                                                            const float* __restrict__ part, int splits, int64_t pstride,
                                                            const float* __restrict__ c_prev, int64_t ldcp,
                                                            float* __restrict__ gates, int64_t ldg,
                                                            float* __restrict__ c_next, int64_t ldcn,
                                                            float* __restrict__ h_next, int64_t ldh,
                                                            int64_t n, int64_t h, int pdl) {
  if (pdl) {      // programmatic dependent launch (lstm_seq.cu): release the successor's set-up, wait for the predecessor's data
    pdl_launch_dependents();
    pdl_wait();
  }
  constexpr int V = VEC ? 4 : 1;
  const int64_t hv = h / V, total = n * hv;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int64_t r = idx / hv, j = (idx - r * hv) * V;
    __align__(16) float af[V]; __align__(16) float ag[V]; __align__(16) float ai[V]; __align__(16) float ao[V]; __align__(16) float cp[V]; __align__(16) float f[V]; __align__(16) float g[V]; __align__(16) float ii[V]; __align__(16) float o[V]; __align__(16) float cn[V]; __align__(16) float hn[V];
    if constexpr (VEC) {
      *reinterpret_cast<float4*>(af) = ldg_stream(reinterpret_cast<const float4*>(a + r * lda + j));
      *reinterpret_cast<float4*>(ag) = ldg_stream(reinterpret_cast<const float4*>(a + r * lda + h + j));
      *reinterpret_cast<float4*>(ai) = ldg_stream(reinterpret_cast<const float4*>(a + r * lda + 2 * h + j));
      *reinterpret_cast<float4*>(ao) = ldg_stream(reinterpret_cast<const float4*>(a + r * lda + 3 * h + j));
      *reinterpret_cast<float4*>(cp) = ldg_stream(reinterpret_cast<const float4*>(c_prev + r * ldcp + j));
    } else {
      af[0] = a[r * lda + j]; ag[0] = a[r * lda + h + j]; ai[0] = a[r * lda + 2 * h + j];
      ao[0] = a[r * lda + 3 * h + j]; cp[0] = c_prev[r * ldcp + j];
    }
    if (part) {     // A += h_{t-1} Wh, still in split-K partial form ([split][n][4h]): summed here instead of in a pass of its own
      const float* pr = part + r * 4 * h + j;
      for (int sp = 0; sp < splits; ++sp, pr += pstride) {
        add_partial<V>(af, pr);
        add_partial<V>(ag, pr + h);
        add_partial<V>(ai, pr + 2 * h);
        add_partial<V>(ao, pr + 3 * h);
      }
    }
#pragma unroll
    for (int u = 0; u < V; ++u) {
      f[u] = sigmoidf_(af[u]);   // slice1  layers.py:636
      g[u] = tanhf(ag[u]);       // slice2  :637
      ii[u] = sigmoidf_(ai[u]);  // slice3  :638
      o[u] = sigmoidf_(ao[u]);   // slice4  :639
      cn[u] = f[u] * cp[u] + g[u] * ii[u];  // :641
      hn[u] = o[u] * tanhf(cn[u]);          // :642
    }
    if constexpr (VEC) {
      *reinterpret_cast<float4*>(gates + r * ldg + j) = *reinterpret_cast<float4*>(f);
      *reinterpret_cast<float4*>(gates + r * ldg + h + j) = *reinterpret_cast<float4*>(g);
      *reinterpret_cast<float4*>(gates + r * ldg + 2 * h + j) = *reinterpret_cast<float4*>(ii);
      *reinterpret_cast<float4*>(gates + r * ldg + 3 * h + j) = *reinterpret_cast<float4*>(o);
      *reinterpret_cast<float4*>(c_next + r * ldcn + j) = *reinterpret_cast<float4*>(cn);
      *reinterpret_cast<float4*>(h_next + r * ldh + j) = *reinterpret_cast<float4*>(hn);
    } else {
      gates[r * ldg + j] = f[0]; gates[r * ldg + h + j] = g[0]; gates[r * ldg + 2 * h + j] = ii[0];
      gates[r * ldg + 3 * h + j] = o[0]; c_next[r * ldcn + j] = cn[0]; h_next[r * ldh + j] = hn[0];
    }
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256) lstm_cell_bwd_kernel(
    const float* __restrict__ dh_a, int64_t ld_dha, const float* __restrict__ dh_b, int64_t ld_dhb,
    const float* __restrict__ part, int splits, int64_t pstride,
    const float* __restrict__ dc_next, const float* __restrict__ gates, int64_t ldg,
    const float* __restrict__ c_prev, int64_t ldcp, const float* __restrict__ c_next, int64_t ldcn,
    float* __restrict__ da, int64_t ldda, float* __restrict__ dc_prev, int64_t n, int64_t h, int pdl) {
  if (pdl) {
    pdl_launch_dependents();
    pdl_wait();
  }
  constexpr int V = VEC ? 4 : 1;
  const int64_t hv = h / V, total = n * hv;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int64_t r = idx / hv, j = (idx - r * hv) * V;
    __align__(16) float dh[V]; __align__(16) float dhb[V]; __align__(16) float dcn[V]; __align__(16) float f[V]; __align__(16) float g[V]; __align__(16) float ii[V]; __align__(16) float o[V]; __align__(16) float cp[V]; __align__(16) float cn[V];
    __align__(16) float s0[V]; __align__(16) float s1[V]; __align__(16) float s2[V]; __align__(16) float s3[V]; __align__(16) float dcp[V];
#pragma unroll
    for (int u = 0; u < V; ++u) { dhb[u] = 0.f; dcn[u] = 0.f; }
    if constexpr (VEC) {
      *reinterpret_cast<float4*>(dh) = ldg_stream(reinterpret_cast<const float4*>(dh_a + r * ld_dha + j));
      if (dh_b) *reinterpret_cast<float4*>(dhb) = ldg_stream(reinterpret_cast<const float4*>(dh_b + r * ld_dhb + j));
      if (dc_next) *reinterpret_cast<float4*>(dcn) = ldg_stream(reinterpret_cast<const float4*>(dc_next + r * h + j));
      *reinterpret_cast<float4*>(f) = ldg_stream(reinterpret_cast<const float4*>(gates + r * ldg + j));
      *reinterpret_cast<float4*>(g) = ldg_stream(reinterpret_cast<const float4*>(gates + r * ldg + h + j));
      *reinterpret_cast<float4*>(ii) = ldg_stream(reinterpret_cast<const float4*>(gates + r * ldg + 2 * h + j));
      *reinterpret_cast<float4*>(o) = ldg_stream(reinterpret_cast<const float4*>(gates + r * ldg + 3 * h + j));
      *reinterpret_cast<float4*>(cp) = ldg_stream(reinterpret_cast<const float4*>(c_prev + r * ldcp + j));
      *reinterpret_cast<float4*>(cn) = ldg_stream(reinterpret_cast<const float4*>(c_next + r * ldcn + j));
    } else {
      dh[0] = dh_a[r * ld_dha + j];
      if (dh_b) dhb[0] = dh_b[r * ld_dhb + j];
      if (dc_next) dcn[0] = dc_next[r * h + j];
      f[0] = gates[r * ldg + j]; g[0] = gates[r * ldg + h + j]; ii[0] = gates[r * ldg + 2 * h + j];
      o[0] = gates[r * ldg + 3 * h + j]; cp[0] = c_prev[r * ldcp + j]; cn[0] = c_next[r * ldcn + j];
    }
    if (part) {     // the recurrent gradient dA_{t+1} Wh^T in split-K partial form ([split][n][h])
      const float* pr = part + r * h + j;
      for (int sp = 0; sp < splits; ++sp, pr += pstride) add_partial<V>(dhb, pr);
    }
#pragma unroll
    for (int u = 0; u < V; ++u) {
      const float d = dh[u] + dhb[u];                  // dhs[:, t, :] + dh          layers.py:590
      const float tc = tanhf(cn[u]);
      const float one_m_t2 = 1.f - tc * tc;            // tang_c_2_1                 :655
      const float dho = d * o[u];                      // dh_next_o                  :652
      const float ds = dho * one_m_t2 + dcn[u];        // d loss / d c_t
      s0[u] = ds * f[u] * cp[u] * (1.f - f[u]);        // stack0 (f gate)            :659
      s1[u] = ds * ii[u] * (1.f - g[u] * g[u]);        // stack1 (g gate)            :660
      s2[u] = ds * g[u] * ii[u] * (1.f - ii[u]);       // stack2 (i gate)            :661
      s3[u] = d * (o[u] * tc) * (1.f - o[u]);          // stack3: dh*ht*(1-o)        :662
      dcp[u] = ds * f[u];                              // dc                         :670
    }
    if constexpr (VEC) {
      *reinterpret_cast<float4*>(da + r * ldda + j) = *reinterpret_cast<float4*>(s0);
      *reinterpret_cast<float4*>(da + r * ldda + h + j) = *reinterpret_cast<float4*>(s1);
      *reinterpret_cast<float4*>(da + r * ldda + 2 * h + j) = *reinterpret_cast<float4*>(s2);
      *reinterpret_cast<float4*>(da + r * ldda + 3 * h + j) = *reinterpret_cast<float4*>(s3);
      *reinterpret_cast<float4*>(dc_prev + r * h + j) = *reinterpret_cast<float4*>(dcp);
    } else {
      da[r * ldda + j] = s0[0]; da[r * ldda + h + j] = s1[0]; da[r * ldda + 2 * h + j] = s2[0];
      da[r * ldda + 3 * h + j] = s3[0]; dc_prev[r * h + j] = dcp[0];
    }
  }
}

// gx = (g_a + g_b) * act'(y)  — rnn cell backward (ACT=1) and generic activation backward.
template <int ACT, bool VEC>
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ ga, int64_t ldga,
                                                      const float* __restrict__ gb, int64_t ldgb,
                                                      const float* __restrict__ y, int64_t ldy,
                                                      float* __restrict__ gx, int64_t ldgx, int64_t rows,
                                                      int64_t cols) {
  constexpr int V = VEC ? 4 : 1;
  const int64_t cv = cols / V, total = rows * cv;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int64_t r = idx / cv, c = (idx - r * cv) * V;
    __align__(16) float a[V]; __align__(16) float b[V]; __align__(16) float yy[V]; __align__(16) float o[V];
#pragma unroll
    for (int u = 0; u < V; ++u) b[u] = 0.f;
    if constexpr (VEC) {
      *reinterpret_cast<float4*>(a) = ldg_stream(reinterpret_cast<const float4*>(ga + r * ldga + c));
      if (gb) *reinterpret_cast<float4*>(b) = ldg_stream(reinterpret_cast<const float4*>(gb + r * ldgb + c));
      *reinterpret_cast<float4*>(yy) = ldg_stream(reinterpret_cast<const float4*>(y + r * ldy + c));
    } else {
      a[0] = ga[r * ldga + c];
      if (gb) b[0] = gb[r * ldgb + c];
      yy[0] = y[r * ldy + c];
    }
#pragma unroll
    for (int u = 0; u < V; ++u) o[u] = act_bwd<ACT>(a[u] + b[u], yy[u]);
    if constexpr (VEC) *reinterpret_cast<float4*>(gx + r * ldgx + c) = *reinterpret_cast<float4*>(o);
    else gx[r * ldgx + c] = o[0];
  }
}

template <int ACT, bool VEC>
__global__ void __launch_bounds__(256) bias_act_fwd_kernel(const float* __restrict__ x, int64_t ldx,
                                                           const float* __restrict__ bias, float* __restrict__ out,
                                                           int64_t ldo, int64_t rows, int64_t cols) {
  constexpr int V = VEC ? 4 : 1;
  const int64_t cv = cols / V, total = rows * cv;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int64_t r = idx / cv, c = (idx - r * cv) * V;
    __align__(16) float a[V]; __align__(16) float b[V]; __align__(16) float o[V];
#pragma unroll
    for (int u = 0; u < V; ++u) b[u] = 0.f;
    if constexpr (VEC) {
      *reinterpret_cast<float4*>(a) = ldg_stream(reinterpret_cast<const float4*>(x + r * ldx + c));
      if (bias) *reinterpret_cast<float4*>(b) = __ldg(reinterpret_cast<const float4*>(bias + c));
    } else {
      a[0] = x[r * ldx + c];
      if (bias) b[0] = __ldg(bias + c);
    }
#pragma unroll
    for (int u = 0; u < V; ++u) o[u] = act_fwd<ACT>(a[u] + b[u]);
    if constexpr (VEC) *reinterpret_cast<float4*>(out + r * ldo + c) = *reinterpret_cast<float4*>(o);
    else out[r * ldo + c] = o[0];
  }
}

static inline bool ld_ok(int64_t ld) { return (ld & 3) == 0; }

static cudaLaunchConfig_t pdl_config(int grid, int block, cudaStream_t s, bool pdl) {
  static thread_local cudaLaunchAttribute attr[1];
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.stream = s;
  if (pdl) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  return cfg;
}

int lstm_cell_fwd_launch(const float* a, int64_t lda, const float* part, int splits, int64_t part_stride, const float* c_prev,
                         int64_t ldc_prev, float* gates, int64_t ldg, float* c_next, int64_t ldc_next, float* h_next, int64_t ldh,
                         int64_t n, int64_t h, cudaStream_t s, bool pdl) {
  B200AD_REQUIRE(n >= 0 && h >= 0, "lstm_cell_fwd: negative shape");
  if (n == 0 || h == 0) return B200AD_OK;
  B200AD_REQUIRE(a && c_prev && gates && c_next && h_next, "lstm_cell_fwd: null pointer");
  B200AD_REQUIRE(lda >= 4 * h && ldg >= 4 * h && ldc_prev >= h && ldc_next >= h && ldh >= h, "lstm_cell_fwd: bad ld");
  // (16 B accesses only when that still leaves every SM several warps: a recurrent step of 64 x 1024 units is latency-bound,
  // 64 blocks of float4 threads measured slower than 256 blocks of scalar ones)
  const bool vec = (h % 4 == 0) && n * h >= (int64_t)sm_count() * 2048 && ld_ok(lda) && ld_ok(ldc_prev) && ld_ok(ldg) && ld_ok(ldc_next) && ld_ok(ldh) &&
                   aligned16(a) && aligned16(c_prev) && aligned16(gates) && aligned16(c_next) && aligned16(h_next) &&
                   (!part || (aligned16(part) && (part_stride & 3) == 0));
  const int grid = stream_grid(n * h / (vec ? 4 : 1), 256);
  cudaLaunchConfig_t cfg = pdl_config(grid, 256, s, pdl);
  const int ipdl = pdl ? 1 : 0;
  cudaError_t e = vec ? cudaLaunchKernelEx(&cfg, lstm_cell_fwd_kernel<true>, a, lda, part, splits, part_stride, c_prev, ldc_prev, gates, ldg, c_next, ldc_next, h_next, ldh, n, h, ipdl)
                      : cudaLaunchKernelEx(&cfg, lstm_cell_fwd_kernel<false>, a, lda, part, splits, part_stride, c_prev, ldc_prev, gates, ldg, c_next, ldc_next, h_next, ldh, n, h, ipdl);
  if (e != cudaSuccess) return set_error(B200AD_ERR_CUDA, "lstm_cell_fwd: launch: %s", cudaGetErrorString(e));
  count_launch();
  return check_launch("lstm_cell_fwd");
}

int lstm_cell_bwd_launch(const float* dh_a, int64_t ld_dha, const float* dh_b, int64_t ld_dhb, const float* part, int splits,
                         int64_t part_stride, const float* dc_next, const float* gates, int64_t ldg, const float* c_prev,
                         int64_t ldc_prev, const float* c_next, int64_t ldc_next, float* da, int64_t ldda, float* dc_prev,
                         int64_t n, int64_t h, cudaStream_t s, bool pdl) {
  B200AD_REQUIRE(n >= 0 && h >= 0, "lstm_cell_bwd: negative shape");
  if (n == 0 || h == 0) return B200AD_OK;
  B200AD_REQUIRE(dh_a && gates && c_prev && c_next && da && dc_prev, "lstm_cell_bwd: null pointer");
  const bool vec = (h % 4 == 0) && n * h >= (int64_t)sm_count() * 2048 && ld_ok(ld_dha) && (!dh_b || ld_ok(ld_dhb)) && ld_ok(ldg) && ld_ok(ldc_prev) &&
                   ld_ok(ldc_next) && ld_ok(ldda) && aligned16(dh_a) && (!dh_b || aligned16(dh_b)) &&
                   (!dc_next || aligned16(dc_next)) && aligned16(gates) && aligned16(c_prev) && aligned16(c_next) &&
                   aligned16(da) && aligned16(dc_prev) && (!part || (aligned16(part) && (part_stride & 3) == 0));
  const int grid = stream_grid(n * h / (vec ? 4 : 1), 256);
  cudaLaunchConfig_t cfg = pdl_config(grid, 256, s, pdl);
  const int ipdl = pdl ? 1 : 0;
  cudaError_t e = vec ? cudaLaunchKernelEx(&cfg, lstm_cell_bwd_kernel<true>, dh_a, ld_dha, dh_b, ld_dhb, part, splits, part_stride, dc_next, gates, ldg, c_prev, ldc_prev, c_next, ldc_next, da, ldda, dc_prev, n, h, ipdl)
                      : cudaLaunchKernelEx(&cfg, lstm_cell_bwd_kernel<false>, dh_a, ld_dha, dh_b, ld_dhb, part, splits, part_stride, dc_next, gates, ldg, c_prev, ldc_prev, c_next, ldc_next, da, ldda, dc_prev, n, h, ipdl);
  if (e != cudaSuccess) return set_error(B200AD_ERR_CUDA, "lstm_cell_bwd: launch: %s", cudaGetErrorString(e));
  count_launch();
  return check_launch("lstm_cell_bwd");
}

}  // namespace b200ad

using namespace b200ad;

extern "C" {

int b200ad_lstm_cell_fwd(const float* a, int64_t lda, const float* c_prev, int64_t ldc_prev, float* gates,
                         int64_t ldg, float* c_next, int64_t ldc_next, float* h_next, int64_t ldh, int64_t n,
                         int64_t h, void* stream) {
  return lstm_cell_fwd_launch(a, lda, nullptr, 0, 0, c_prev, ldc_prev, gates, ldg, c_next, ldc_next, h_next, ldh, n, h,
                              (cudaStream_t)stream);
}

int b200ad_lstm_cell_bwd(const float* dh_a, int64_t ld_dha, const float* dh_b, int64_t ld_dhb, const float* dc_next,
                         const float* gates, int64_t ldg, const float* c_prev, int64_t ldc_prev, const float* c_next,
                         int64_t ldc_next, const float* h_next, int64_t ldh, float* da, int64_t ldda, float* dc_prev,
                         int64_t n, int64_t h, void* stream) {
  (void)h_next; (void)ldh;  // h' = o*tanh(c') is recomputed from the cached gates
  return lstm_cell_bwd_launch(dh_a, ld_dha, dh_b, ld_dhb, nullptr, 0, 0, dc_next, gates, ldg, c_prev, ldc_prev, c_next,
                              ldc_next, da, ldda, dc_prev, n, h, (cudaStream_t)stream);
}

static int launch_act_bwd(const float* ga, int64_t ldga, const float* gb, int64_t ldgb, const float* y, int64_t ldy,
                          float* gx, int64_t ldgx, int64_t rows, int64_t cols, int act, void* stream, const char* who) {
  B200AD_REQUIRE(rows >= 0 && cols >= 0, "%s: negative shape", who);
  if (rows == 0 || cols == 0) return B200AD_OK;
  B200AD_REQUIRE(ga && y && gx, "%s: null pointer", who);
  B200AD_REQUIRE(act >= 0 && act <= 3, "%s: bad activation %d", who, act);
  const bool vec = (cols % 4 == 0) && ld_ok(ldga) && (!gb || ld_ok(ldgb)) && ld_ok(ldy) && ld_ok(ldgx) &&
                   aligned16(ga) && (!gb || aligned16(gb)) && aligned16(y) && aligned16(gx);
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = stream_grid(rows * cols / (vec ? 4 : 1), 256);
#define L(ACT)                                                                                                   \
  if (vec) act_bwd_kernel<ACT, true><<<grid, 256, 0, s>>>(ga, ldga, gb, ldgb, y, ldy, gx, ldgx, rows, cols);     \
  else act_bwd_kernel<ACT, false><<<grid, 256, 0, s>>>(ga, ldga, gb, ldgb, y, ldy, gx, ldgx, rows, cols);
  switch (act) {
    case 0: L(0) break;
    case 1: L(1) break;
    case 2: L(2) break;
    default: L(3) break;
  }
#undef L
  count_launch();
  return check_launch(who);
}

int b200ad_rnn_cell_bwd(const float* dh_a, int64_t ld_dha, const float* dh_b, int64_t ld_dhb, const float* h_next,
                        int64_t ldh, float* dt, int64_t lddt, int64_t n, int64_t h, void* stream) {
  return launch_act_bwd(dh_a, ld_dha, dh_b, ld_dhb, h_next, ldh, dt, lddt, n, h, 1, stream, "rnn_cell_bwd");
}

int b200ad_act_bwd(const float* g, int64_t ldg, const float* y, int64_t ldy, float* gx, int64_t ldgx, int64_t rows,
                   int64_t cols, int act, void* stream) {
  return launch_act_bwd(g, ldg, nullptr, 0, y, ldy, gx, ldgx, rows, cols, act, stream, "act_bwd");
}

int b200ad_bias_act_fwd(const float* x, int64_t ldx, const float* bias, float* out, int64_t ldo, int64_t rows,
                        int64_t cols, int act, void* stream) {
  B200AD_REQUIRE(rows >= 0 && cols >= 0, "bias_act_fwd: negative shape");
  if (rows == 0 || cols == 0) return B200AD_OK;
  B200AD_REQUIRE(x && out, "bias_act_fwd: null pointer");
  B200AD_REQUIRE(act >= 0 && act <= 3, "bias_act_fwd: bad activation %d", act);
  const bool vec = (cols % 4 == 0) && ld_ok(ldx) && ld_ok(ldo) && aligned16(x) && aligned16(out) &&
                   (!bias || aligned16(bias));
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = stream_grid(rows * cols / (vec ? 4 : 1), 256);
#define L(ACT)                                                                                  \
  if (vec) bias_act_fwd_kernel<ACT, true><<<grid, 256, 0, s>>>(x, ldx, bias, out, ldo, rows, cols); \
  else bias_act_fwd_kernel<ACT, false><<<grid, 256, 0, s>>>(x, ldx, bias, out, ldo, rows, cols);
  switch (act) {
    case 0: L(0) break;
    case 1: L(1) break;
    case 2: L(2) break;
    default: L(3) break;
  }
#undef L
  count_launch();
  return check_launch("bias_act_fwd");
}
}
