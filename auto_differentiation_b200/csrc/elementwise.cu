// Elementwise unary / binary forward and backward kernels (HBM-bound: 128-bit streaming access,
// grid-stride over whole waves of 148 SMs x 8 CTAs, four independent float4 loads in flight).
//
// Reference bodies replaced (auto_grad/tgrad.py): exp :1053-1075, tanh :1591-1615, Sigmoid :883-903,
// ln :1272-1297, log :1314-1341, sqrt :1760-1786, square :1716-1743, pow :1085-1105, Abs :1838-1858,
// trig family :1115-1713, Add :666-689, sub :747-771, Multi :844-880, Divide :996-1028;
// layers/layers.py ReLU :1280-1290.
#include "common.cuh"
#include "colsum.cuh"

namespace b200ad {

// ---------------------------------------------------------------------------------------------
// Unary functors
// ---------------------------------------------------------------------------------------------
template <int OP>
__device__ __forceinline__ float u_fwd(float x, float p) {
  if constexpr (OP == B200AD_U_EXP) return expf(x);
  else if constexpr (OP == B200AD_U_TANH) return tanhf(x);
  else if constexpr (OP == B200AD_U_SIGMOID) return 1.f / (1.f + expf(-x));
  else if constexpr (OP == B200AD_U_RELU) return (x > 0.f || x != x) ? x : 0.f;  // np.maximum keeps NaN
  else if constexpr (OP == B200AD_U_LN) return logf(x);
  else if constexpr (OP == B200AD_U_LOGB) return logf(x) / logf(p);
  else if constexpr (OP == B200AD_U_SQRT) return sqrtf(x);
  else if constexpr (OP == B200AD_U_SQUARE) return x * x;
  else if constexpr (OP == B200AD_U_POW) return powf(x, p);
  else if constexpr (OP == B200AD_U_ABS) return fabsf(x);
  else if constexpr (OP == B200AD_U_NEG) return -x;
  else if constexpr (OP == B200AD_U_SIN) return sinf(x);
  else if constexpr (OP == B200AD_U_COS) return cosf(x);
  else if constexpr (OP == B200AD_U_TAN) return tanf(x);
  else if constexpr (OP == B200AD_U_SEC) return 1.f / cosf(x);
  else if constexpr (OP == B200AD_U_SINH) return sinhf(x);
  else if constexpr (OP == B200AD_U_COSH) return coshf(x);
  else if constexpr (OP == B200AD_U_ARCSIN) return asinf(x);
  else if constexpr (OP == B200AD_U_ARCTAN) return atanf(x);
  else if constexpr (OP == B200AD_U_ARCOT) return atanf(1.f / x);
  else if constexpr (OP == B200AD_U_ARCSEC) return acosf(1.f / x);
  else if constexpr (OP == B200AD_U_COT) return 1.f / tanf(x);
  else if constexpr (OP == B200AD_U_CSC) return 1.f / sinf(x);
  else if constexpr (OP == B200AD_U_ARCCOS) return acosf(x);
  else if constexpr (OP == B200AD_U_ADDC) return x + p;
  else if constexpr (OP == B200AD_U_DIVC) return x / p;
  else if constexpr (OP == B200AD_U_RSUBC) return p - x;
  else if constexpr (OP == B200AD_U_RDIVC) return p / x;
  else return p * x;  // SCALE
}

// 0: reads x, 1: reads y, 2: neither.
__host__ __device__ constexpr int u_bwd_uses(int op) {
  switch (op) {
    case B200AD_U_EXP: case B200AD_U_TANH: case B200AD_U_SIGMOID: case B200AD_U_RELU:
    case B200AD_U_SQRT:
      return 1;
    case B200AD_U_NEG: case B200AD_U_SCALE: case B200AD_U_ADDC: case B200AD_U_DIVC: case B200AD_U_RSUBC:
      return 2;
    default:
      return 0;
  }
}

// s is x or y according to u_bwd_uses(OP).
template <int OP>
__device__ __forceinline__ float u_bwd(float g, float s, float p) {
  if constexpr (OP == B200AD_U_EXP) return g * s;                       // g*exp(x)       :1059
  else if constexpr (OP == B200AD_U_TANH) return g * (1.f - s * s);     // g*(1-y^2)      :1597
  else if constexpr (OP == B200AD_U_SIGMOID) return g * s * (1.f - s);  // g*y*(1-y)      :888
  else if constexpr (OP == B200AD_U_RELU) return s > 0.f ? g : 0.f * g; // dout*(x>0)     layers.py:1290
  else if constexpr (OP == B200AD_U_LN) return g * (1.f / s);           // :1281
  else if constexpr (OP == B200AD_U_LOGB) return g * (1.f / (s * logf(p)));  // :1325
  else if constexpr (OP == B200AD_U_SQRT) return g * 0.5f / s;          // g*0.5*x^-0.5   :1771
  else if constexpr (OP == B200AD_U_SQUARE) return g * 2.f * s;         // :1727
  else if constexpr (OP == B200AD_U_POW) return g * p * powf(s, p - 1.f);  // :1092
  else if constexpr (OP == B200AD_U_ABS) return s < 0.f ? -g : g;       // :1851-1853 (+1 at 0)
  else if constexpr (OP == B200AD_U_NEG) return -g;
  else if constexpr (OP == B200AD_U_SIN) return g * cosf(s);            // :1121
  else if constexpr (OP == B200AD_U_COS) return -g * sinf(s);           // :1364
  else if constexpr (OP == B200AD_U_TAN) { float c = cosf(s); return g * (1.f / (c * c)); }  // :1518
  else if constexpr (OP == B200AD_U_SEC) return g * tanf(s) * (1.f / cosf(s));  // :1152
  else if constexpr (OP == B200AD_U_SINH) return g * coshf(s);          // :1185
  else if constexpr (OP == B200AD_U_COSH) return g * sinhf(s);          // :1402
  else if constexpr (OP == B200AD_U_ARCSIN) return g * (1.f / sqrtf(1.f - s * s));  // :1216
  else if constexpr (OP == B200AD_U_ARCTAN) return g * (1.f / (1.f + s * s));       // :1559
  else if constexpr (OP == B200AD_U_ARCOT) return g * (-1.f / (1.f + s * s));       // :1683
  else if constexpr (OP == B200AD_U_ARCSEC) return g * (1.f / (fabsf(s) * sqrtf(s * s - 1.f)));  // :1247
  else if constexpr (OP == B200AD_U_COT) { float sn = sinf(s); return g * (1.f / (sn * sn)); }   // :1639 (sign as written)
  else if constexpr (OP == B200AD_U_CSC) { float sn = sinf(s); return -g * (1.f / sn) * (cosf(s) / sn); }   // :1476 (intent)
  else if constexpr (OP == B200AD_U_ARCCOS) return g * (-1.f / sqrtf(s + 1.f));                 // :1438 (as written)
  else if constexpr (OP == B200AD_U_ADDC) return g;
  else if constexpr (OP == B200AD_U_DIVC) return g / (p + 1e-8f);                   // :1004
  else if constexpr (OP == B200AD_U_RSUBC) return -g;
  else if constexpr (OP == B200AD_U_RDIVC) return g * (-p / (s * s + 1e-8f));      // :1005
  else return g * p;  // SCALE
}

template <int OP>
__global__ void __launch_bounds__(256) unary_fwd_kernel(const float* __restrict__ x,
                                                        float* __restrict__ y, int64_t n, float p,
                                                        int vec_ok) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    int64_t i = tid;
    for (; i + 3 * stride < n4; i += 4 * stride) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ldg_stream(x4 + i + u * stride);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u].x = u_fwd<OP>(v[u].x, p); v[u].y = u_fwd<OP>(v[u].y, p);
        v[u].z = u_fwd<OP>(v[u].z, p); v[u].w = u_fwd<OP>(v[u].w, p);
        stg_stream(y4 + i + u * stride, v[u]);
      }
    }
    for (; i < n4; i += stride) {
      float4 v = ldg_stream(x4 + i);
      v.x = u_fwd<OP>(v.x, p); v.y = u_fwd<OP>(v.y, p);
      v.z = u_fwd<OP>(v.z, p); v.w = u_fwd<OP>(v.w, p);
      stg_stream(y4 + i, v);
    }
    const int64_t t = (n4 << 2) + tid;  // tail (< 4 elements)
    if (t < n) y[t] = u_fwd<OP>(x[t], p);
  } else {
    for (int64_t i = tid; i < n; i += stride) y[i] = u_fwd<OP>(__ldg(x + i), p);
  }
}

template <int OP>
__global__ void __launch_bounds__(256) unary_bwd_kernel(const float* __restrict__ g,
                                                        const float* __restrict__ s,
                                                        float* __restrict__ gx, int64_t n, float p,
                                                        int vec_ok) {
  constexpr bool kReads = u_bwd_uses(OP) != 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const float4* s4 = reinterpret_cast<const float4*>(s);
    float4* o4 = reinterpret_cast<float4*>(gx);
    int64_t i = tid;
    for (; i + stride < n4; i += 2 * stride) {
      float4 a[2], b[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        a[u] = ldg_stream(g4 + i + u * stride);
        b[u] = kReads ? ldg_stream(s4 + i + u * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float4 r;
        r.x = u_bwd<OP>(a[u].x, b[u].x, p); r.y = u_bwd<OP>(a[u].y, b[u].y, p);
        r.z = u_bwd<OP>(a[u].z, b[u].z, p); r.w = u_bwd<OP>(a[u].w, b[u].w, p);
        stg_stream(o4 + i + u * stride, r);
      }
    }
    for (; i < n4; i += stride) {
      float4 a = ldg_stream(g4 + i);
      float4 b = kReads ? ldg_stream(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 r;
      r.x = u_bwd<OP>(a.x, b.x, p); r.y = u_bwd<OP>(a.y, b.y, p);
      r.z = u_bwd<OP>(a.z, b.z, p); r.w = u_bwd<OP>(a.w, b.w, p);
      stg_stream(o4 + i, r);
    }
    const int64_t t = (n4 << 2) + tid;
    if (t < n) gx[t] = u_bwd<OP>(g[t], kReads ? s[t] : 0.f, p);
  } else {
    for (int64_t i = tid; i < n; i += stride)
      gx[i] = u_bwd<OP>(__ldg(g + i), kReads ? __ldg(s + i) : 0.f, p);
  }
}

#define B200AD_FOR_EACH_UNARY(X)                                                                  \
  X(B200AD_U_EXP) X(B200AD_U_TANH) X(B200AD_U_SIGMOID) X(B200AD_U_RELU) X(B200AD_U_LN)            \
  X(B200AD_U_LOGB) X(B200AD_U_SQRT) X(B200AD_U_SQUARE) X(B200AD_U_POW) X(B200AD_U_ABS)            \
  X(B200AD_U_NEG) X(B200AD_U_SIN) X(B200AD_U_COS) X(B200AD_U_TAN) X(B200AD_U_SEC) X(B200AD_U_SINH) \
  X(B200AD_U_COSH) X(B200AD_U_ARCSIN) X(B200AD_U_ARCTAN) X(B200AD_U_ARCOT) X(B200AD_U_ARCSEC)     \
  X(B200AD_U_SCALE) X(B200AD_U_ADDC) X(B200AD_U_DIVC) X(B200AD_U_RSUBC) X(B200AD_U_RDIVC)                \
  X(B200AD_U_COT) X(B200AD_U_CSC) X(B200AD_U_ARCCOS)

// ---------------------------------------------------------------------------------------------
// Binary ops with collapsed broadcasting
// ---------------------------------------------------------------------------------------------
template <int OP>
__device__ __forceinline__ float b_fwd(float a, float b) {
  if constexpr (OP == B200AD_B_ADD) return a + b;
  else if constexpr (OP == B200AD_B_SUB) return a - b;
  else if constexpr (OP == B200AD_B_MUL) return a * b;
  else return a / b;
}
// d out / d a  and  d out / d b  times g, with the reference's quirks.
template <int OP>
__device__ __forceinline__ float b_bwd_a(float g, float a, float b, int ref_compat) {
  if constexpr (OP == B200AD_B_ADD) return g;
  else if constexpr (OP == B200AD_B_SUB) return ref_compat ? -g : g;       // tgrad.py:754 (-1*grad, grad)
  else if constexpr (OP == B200AD_B_MUL) return g * b;                      // :851
  else return g / (b + 1e-8f);                                              // :1004
}
template <int OP>
__device__ __forceinline__ float b_bwd_b(float g, float a, float b, int ref_compat) {
  if constexpr (OP == B200AD_B_ADD) return g;
  else if constexpr (OP == B200AD_B_SUB) return ref_compat ? g : -g;
  else if constexpr (OP == B200AD_B_MUL) return g * a;
  else return g * (-a / (b * b + 1e-8f));                                   // :1005
}

template <typename I>
__device__ __forceinline__ float4 load4_mode(const float* __restrict__ p, int mode, I i4, I period) {
  if (mode == B200AD_FULL) return ldg_stream(reinterpret_cast<const float4*>(p) + i4);
  if (mode == B200AD_ROW) return __ldg(reinterpret_cast<const float4*>(p + (I)(i4 * 4) % period));
  float s = (mode == B200AD_COL) ? __ldg(p + (I)(i4 * 4) / period) : __ldg(p);
  return make_float4(s, s, s, s);
}
template <typename I>
__device__ __forceinline__ float load1_mode(const float* __restrict__ p, int mode, I i, I period) {
  if (mode == B200AD_FULL) return __ldg(p + i);
  if (mode == B200AD_ROW) return __ldg(p + i % period);
  if (mode == B200AD_COL) return __ldg(p + i / period);
  return __ldg(p);
}

template <int OP, typename I>
__global__ void __launch_bounds__(256) binary_fwd_kernel(const float* __restrict__ a,
                                                         const float* __restrict__ b,
                                                         float* __restrict__ out, I n, int a_mode,
                                                         int b_mode, I period, int vec_ok) {
  const I stride = (I)gridDim.x * blockDim.x;
  const I tid = (I)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec_ok) {
    const I n4 = n >> 2;
    float4* o4 = reinterpret_cast<float4*>(out);
    I i = tid;
    for (; i + stride < n4; i += 2 * stride) {
      float4 x[2], y[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        x[u] = load4_mode<I>(a, a_mode, i + u * stride, period);
        y[u] = load4_mode<I>(b, b_mode, i + u * stride, period);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float4 r;
        r.x = b_fwd<OP>(x[u].x, y[u].x); r.y = b_fwd<OP>(x[u].y, y[u].y);
        r.z = b_fwd<OP>(x[u].z, y[u].z); r.w = b_fwd<OP>(x[u].w, y[u].w);
        stg_stream(o4 + i + u * stride, r);
      }
    }
    for (; i < n4; i += stride) {
      float4 x = load4_mode<I>(a, a_mode, i, period), y = load4_mode<I>(b, b_mode, i, period), r;
      r.x = b_fwd<OP>(x.x, y.x); r.y = b_fwd<OP>(x.y, y.y);
      r.z = b_fwd<OP>(x.z, y.z); r.w = b_fwd<OP>(x.w, y.w);
      stg_stream(o4 + i, r);
    }
    const I t = (n4 << 2) + tid;
    if (t < n) out[t] = b_fwd<OP>(load1_mode<I>(a, a_mode, t, period), load1_mode<I>(b, b_mode, t, period));
  } else {
    for (I i = tid; i < n; i += stride)
      out[i] = b_fwd<OP>(load1_mode<I>(a, a_mode, i, period), load1_mode<I>(b, b_mode, i, period));
  }
}

template <int OP, typename I>
__global__ void __launch_bounds__(256) binary_bwd_kernel(const float* __restrict__ g,
                                                         const float* __restrict__ a,
                                                         const float* __restrict__ b,
                                                         float* __restrict__ ga, float* __restrict__ gb,
                                                         I n, int a_mode, int b_mode, I period,
                                                         int ref_compat, int vec_ok) {
  const I stride = (I)gridDim.x * blockDim.x;
  const I tid = (I)blockIdx.x * blockDim.x + threadIdx.x;
  constexpr bool kNeedsOperands = (OP == B200AD_B_MUL || OP == B200AD_B_DIV);
  if (vec_ok) {
    const I n4 = n >> 2;
    for (I i = tid; i < n4; i += stride) {
      float4 gg = ldg_stream(reinterpret_cast<const float4*>(g) + i);
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
      if (kNeedsOperands) {
        x = load4_mode<I>(a, a_mode, i, period);
        y = load4_mode<I>(b, b_mode, i, period);
      }
      if (ga) {
        float4 r;
        r.x = b_bwd_a<OP>(gg.x, x.x, y.x, ref_compat); r.y = b_bwd_a<OP>(gg.y, x.y, y.y, ref_compat);
        r.z = b_bwd_a<OP>(gg.z, x.z, y.z, ref_compat); r.w = b_bwd_a<OP>(gg.w, x.w, y.w, ref_compat);
        stg_stream(reinterpret_cast<float4*>(ga) + i, r);
      }
      if (gb) {
        float4 r;
        r.x = b_bwd_b<OP>(gg.x, x.x, y.x, ref_compat); r.y = b_bwd_b<OP>(gg.y, x.y, y.y, ref_compat);
        r.z = b_bwd_b<OP>(gg.z, x.z, y.z, ref_compat); r.w = b_bwd_b<OP>(gg.w, x.w, y.w, ref_compat);
        stg_stream(reinterpret_cast<float4*>(gb) + i, r);
      }
    }
    const I t = (n4 << 2) + tid;
    if (t < n) {
      float gg = g[t], x = 0.f, y = 0.f;
      if (kNeedsOperands) { x = load1_mode<I>(a, a_mode, t, period); y = load1_mode<I>(b, b_mode, t, period); }
      if (ga) ga[t] = b_bwd_a<OP>(gg, x, y, ref_compat);
      if (gb) gb[t] = b_bwd_b<OP>(gg, x, y, ref_compat);
    }
  } else {
    for (I i = tid; i < n; i += stride) {
      float gg = __ldg(g + i), x = 0.f, y = 0.f;
      if (kNeedsOperands) { x = load1_mode<I>(a, a_mode, i, period); y = load1_mode<I>(b, b_mode, i, period); }
      if (ga) ga[i] = b_bwd_a<OP>(gg, x, y, ref_compat);
      if (gb) gb[i] = b_bwd_b<OP>(gg, x, y, ref_compat);
    }
  }
}

// Producer for the fused "a FULL (R,C), b ROW (C)" backward: writes ga, yields the gb addend.
template <int OP>
struct RowFusedProducer {
  const float* g; const float* a; const float* b; float* ga; int ref_compat;
  __device__ __forceinline__ float4 get4(int64_t off, int64_t col) const {
    constexpr bool kNeedsOperands = (OP == B200AD_B_MUL || OP == B200AD_B_DIV);
    float4 gg = ldg_stream(reinterpret_cast<const float4*>(g + off));
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
    if (kNeedsOperands) {
      x = ldg_stream(reinterpret_cast<const float4*>(a + off));
      y = __ldg(reinterpret_cast<const float4*>(b + col));
    }
    if (ga) {
      float4 r;
      r.x = b_bwd_a<OP>(gg.x, x.x, y.x, ref_compat); r.y = b_bwd_a<OP>(gg.y, x.y, y.y, ref_compat);
      r.z = b_bwd_a<OP>(gg.z, x.z, y.z, ref_compat); r.w = b_bwd_a<OP>(gg.w, x.w, y.w, ref_compat);
      stg_stream(reinterpret_cast<float4*>(ga + off), r);
    }
    float4 r;
    r.x = b_bwd_b<OP>(gg.x, x.x, y.x, ref_compat); r.y = b_bwd_b<OP>(gg.y, x.y, y.y, ref_compat);
    r.z = b_bwd_b<OP>(gg.z, x.z, y.z, ref_compat); r.w = b_bwd_b<OP>(gg.w, x.w, y.w, ref_compat);
    return r;
  }
  __device__ __forceinline__ float get1(int64_t off, int64_t col) const {
    constexpr bool kNeedsOperands = (OP == B200AD_B_MUL || OP == B200AD_B_DIV);
    float gg = __ldg(g + off), x = 0.f, y = 0.f;
    if (kNeedsOperands) { x = __ldg(a + off); y = __ldg(b + col); }
    if (ga) ga[off] = b_bwd_a<OP>(gg, x, y, ref_compat);
    return b_bwd_b<OP>(gg, x, y, ref_compat);
  }
};

}  // namespace b200ad

using namespace b200ad;

extern "C" {

int b200ad_unary_bwd_uses(int op) { return u_bwd_uses(op); }

int b200ad_unary_fwd(int op, const float* x, float* y, int64_t n, float param, void* stream) {
  B200AD_REQUIRE(op >= 0 && op < B200AD_U_COUNT, "unary_fwd: bad op %d", op);
  B200AD_REQUIRE(n >= 0, "unary_fwd: negative n");
  if (n == 0) return B200AD_OK;
  B200AD_REQUIRE(x && y, "unary_fwd: null pointer");
  const int vec_ok = aligned16(x) && aligned16(y);
  const int grid = stream_grid(vec_ok ? (n + 3) / 4 : n, 256 * (vec_ok ? 4 : 1));
  cudaStream_t s = (cudaStream_t)stream;
  switch (op) {
#define X(OP) case OP: unary_fwd_kernel<OP><<<grid, 256, 0, s>>>(x, y, n, param, vec_ok); break;
    B200AD_FOR_EACH_UNARY(X)
#undef X
  }
  count_launch();
  return check_launch("unary_fwd");
}

int b200ad_unary_bwd(int op, const float* g, const float* x, const float* y, float* gx, int64_t n,
                     float param, void* stream) {
  B200AD_REQUIRE(op >= 0 && op < B200AD_U_COUNT, "unary_bwd: bad op %d", op);
  B200AD_REQUIRE(n >= 0, "unary_bwd: negative n");
  if (n == 0) return B200AD_OK;
  const int uses = u_bwd_uses(op);
  const float* s_in = uses == 0 ? x : (uses == 1 ? y : g);
  B200AD_REQUIRE(g && gx && s_in, "unary_bwd: null pointer (op %d reads %s)", op, uses == 0 ? "x" : "y");
  const int vec_ok = aligned16(g) && aligned16(gx) && aligned16(s_in);
  const int grid = stream_grid(vec_ok ? (n + 3) / 4 : n, 256 * (vec_ok ? 2 : 1));
  cudaStream_t s = (cudaStream_t)stream;
  switch (op) {
#define X(OP) case OP: unary_bwd_kernel<OP><<<grid, 256, 0, s>>>(g, s_in, gx, n, param, vec_ok); break;
    B200AD_FOR_EACH_UNARY(X)
#undef X
  }
  count_launch();
  return check_launch("unary_bwd");
}

static int check_modes(const char* who, int a_mode, int b_mode, int64_t n, int64_t period) {
  B200AD_REQUIRE(a_mode >= 0 && a_mode <= 3 && b_mode >= 0 && b_mode <= 3, "%s: bad broadcast mode", who);
  const bool needs_period = a_mode == B200AD_ROW || a_mode == B200AD_COL || b_mode == B200AD_ROW ||
                            b_mode == B200AD_COL;
  if (needs_period) B200AD_REQUIRE(period > 0 && n % period == 0, "%s: period %lld does not divide n %lld", who,
                                   (long long)period, (long long)n);
  return B200AD_OK;
}

static bool modes_vec_ok(const float* p, int mode, int64_t period) {
  if (mode == B200AD_FULL) return aligned16(p);
  if (mode == B200AD_ROW) return aligned16(p) && (period % 4 == 0);
  if (mode == B200AD_COL) return period % 4 == 0;
  return true;
}

int b200ad_binary_fwd(int op, const float* a, const float* b, float* out, int64_t n, int a_mode,
                      int b_mode, int64_t period, void* stream) {
  B200AD_REQUIRE(op >= 0 && op <= 3, "binary_fwd: bad op %d", op);
  B200AD_REQUIRE(n >= 0, "binary_fwd: negative n");
  if (n == 0) return B200AD_OK;
  B200AD_REQUIRE(a && b && out, "binary_fwd: null pointer");
  int rc = check_modes("binary_fwd", a_mode, b_mode, n, period);
  if (rc) return rc;
  if (period <= 0) period = 1;
  const int vec_ok = aligned16(out) && modes_vec_ok(a, a_mode, period) && modes_vec_ok(b, b_mode, period);
  const int grid = stream_grid(vec_ok ? (n + 3) / 4 : n, 256 * (vec_ok ? 2 : 1));
  cudaStream_t s = (cudaStream_t)stream;
  const bool small = n < (int64_t)0x7fffffff;
#define LAUNCH(OP)                                                                                   \
  if (small) binary_fwd_kernel<OP, uint32_t><<<grid, 256, 0, s>>>(a, b, out, (uint32_t)n, a_mode, b_mode, (uint32_t)period, vec_ok); \
  else binary_fwd_kernel<OP, int64_t><<<grid, 256, 0, s>>>(a, b, out, n, a_mode, b_mode, period, vec_ok);
  switch (op) {
    case B200AD_B_ADD: LAUNCH(B200AD_B_ADD) break;
    case B200AD_B_SUB: LAUNCH(B200AD_B_SUB) break;
    case B200AD_B_MUL: LAUNCH(B200AD_B_MUL) break;
    default: LAUNCH(B200AD_B_DIV) break;
  }
#undef LAUNCH
  count_launch();
  return check_launch("binary_fwd");
}

int b200ad_binary_bwd(int op, const float* g, const float* a, const float* b, float* ga, float* gb,
                      int64_t n, int a_mode, int b_mode, int64_t period, int ref_compat,
                      void* stream) {
  B200AD_REQUIRE(op >= 0 && op <= 3, "binary_bwd: bad op %d", op);
  B200AD_REQUIRE(n >= 0, "binary_bwd: negative n");
  if (n == 0 || (!ga && !gb)) return B200AD_OK;
  const bool needs = (op == B200AD_B_MUL || op == B200AD_B_DIV);
  B200AD_REQUIRE(g && (!needs || (a && b)), "binary_bwd: null pointer");
  int rc = check_modes("binary_bwd", a_mode, b_mode, n, period);
  if (rc) return rc;
  if (period <= 0) period = 1;
  int vec_ok = aligned16(g) && (!ga || aligned16(ga)) && (!gb || aligned16(gb));
  if (needs) vec_ok = vec_ok && modes_vec_ok(a, a_mode, period) && modes_vec_ok(b, b_mode, period);
  const int grid = stream_grid(vec_ok ? (n + 3) / 4 : n, 256);
  cudaStream_t s = (cudaStream_t)stream;
  const bool small = n < (int64_t)0x7fffffff;
#define LAUNCH(OP)                                                                                   \
  if (small) binary_bwd_kernel<OP, uint32_t><<<grid, 256, 0, s>>>(g, a, b, ga, gb, (uint32_t)n, a_mode, b_mode, (uint32_t)period, ref_compat, vec_ok); \
  else binary_bwd_kernel<OP, int64_t><<<grid, 256, 0, s>>>(g, a, b, ga, gb, n, a_mode, b_mode, period, ref_compat, vec_ok);
  switch (op) {
    case B200AD_B_ADD: LAUNCH(B200AD_B_ADD) break;
    case B200AD_B_SUB: LAUNCH(B200AD_B_SUB) break;
    case B200AD_B_MUL: LAUNCH(B200AD_B_MUL) break;
    default: LAUNCH(B200AD_B_DIV) break;
  }
#undef LAUNCH
  count_launch();
  return check_launch("binary_bwd");
}

int b200ad_binary_bwd_rowfused(int op, const float* g, const float* a, const float* b, float* ga,
                               float* gb, int64_t rows, int64_t cols, int ref_compat, float* ws,
                               void* stream) {
  B200AD_REQUIRE(op >= 0 && op <= 3, "binary_bwd_rowfused: bad op %d", op);
  B200AD_REQUIRE(rows >= 0 && cols >= 0, "binary_bwd_rowfused: negative shape");
  if (rows == 0 || cols == 0) return B200AD_OK;
  const bool needs = (op == B200AD_B_MUL || op == B200AD_B_DIV);
  B200AD_REQUIRE(g && gb && ws && (!needs || (a && b)), "binary_bwd_rowfused: null pointer");
  const bool vec = (cols % 4 == 0) && aligned16(g) && (!ga || aligned16(ga)) &&
                   (!needs || (aligned16(a) && aligned16(b)));
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  switch (op) {
#define CASE(OP)                                                         \
  case OP: {                                                             \
    RowFusedProducer<OP> p{g, a, b, ga, ref_compat};                     \
    rc = colsum_launch(p, gb, 1, rows, cols, 1.f, ws, vec, s);           \
  } break;
    CASE(B200AD_B_ADD) CASE(B200AD_B_SUB) CASE(B200AD_B_MUL)
    default: {
      RowFusedProducer<B200AD_B_DIV> p{g, a, b, ga, ref_compat};
      rc = colsum_launch(p, gb, 1, rows, cols, 1.f, ws, vec, s);
    } break;
#undef CASE
  }
  return rc;
}
}
