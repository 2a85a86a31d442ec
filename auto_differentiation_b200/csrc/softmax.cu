// Row softmax forward/backward, fused softmax + cross-entropy, argmax (bit-exact index work).
// Reference bodies replaced: softmax tgrad.py:942-970, SoftmaxWithLoss :913-932, layers.softmax
// layers.py:38-43, TimeSoftmaxWithLoss :510-549, multi_cross_entropy_error
// utils/activation_func.py:4-13, argmax/argmin tgrad.py:1803-1835.
//
// Layout: (rows, cols) row-major, reduction over cols.  cols <= 1024: one warp per row, the row
// lives in registers (one HBM read, one write), max/sum by warp shuffles.  Larger rows: one CTA per
// row, three passes (the re-reads hit L1/L2).
#include "common.cuh"

namespace b200ad {

// ---- register-resident row held by a warp ------------------------------------------------------
// VEC: lane holds N float4 (cols % 4 == 0, 16B-aligned rows, cols <= 128*N)
// !VEC: lane holds N scalars (cols <= 32*N)
template <int N, bool VEC>
struct WarpRow {
  float v[VEC ? 4 * N : N];
  static constexpr int kElems = VEC ? 4 * N : N;
  __device__ __forceinline__ void load(const float* __restrict__ row, int cols, int lane, float fill) {
    if constexpr (VEC) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const int c = (lane + 32 * j) * 4;
        if (c < cols) {
          float4 t = ldg_stream(reinterpret_cast<const float4*>(row + c));
          v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
        } else {
          v[4 * j] = v[4 * j + 1] = v[4 * j + 2] = v[4 * j + 3] = fill;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const int c = lane + 32 * j;
        v[j] = c < cols ? __ldg(row + c) : fill;
      }
    }
  }
  // same as load(), through coherent loads and without __restrict__: `row` may alias the kernel's output
  __device__ __forceinline__ void load_aliasable(const float* row, int cols, int lane, float fill) {
    if constexpr (VEC) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const int c = (lane + 32 * j) * 4;
        if (c < cols) {
          float4 t = ld_stream_coherent(reinterpret_cast<const float4*>(row + c));
          v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
        } else {
          v[4 * j] = v[4 * j + 1] = v[4 * j + 2] = v[4 * j + 3] = fill;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const int c = lane + 32 * j;
        v[j] = c < cols ? row[c] : fill;
      }
    }
  }
  __device__ __forceinline__ void store(float* row, int cols, int lane) const {
    if constexpr (VEC) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const int c = (lane + 32 * j) * 4;
        if (c < cols) stg_stream(reinterpret_cast<float4*>(row + c), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const int c = lane + 32 * j;
        if (c < cols) row[c] = v[j];
      }
    }
  }
  // column index of element e of this lane
  __device__ __forceinline__ int col(int e, int lane) const {
    if constexpr (VEC) return (lane + 32 * (e >> 2)) * 4 + (e & 3);
    else return lane + 32 * e;
  }
};

template <int N, bool VEC>
__global__ void __launch_bounds__(256) softmax_fwd_warp(const float* __restrict__ x, float* __restrict__ y,
                                                        int64_t rows, int cols, float in_scale) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    WarpRow<N, VEC> w;
    w.load(x + r * cols, cols, lane, -INFINITY);
    float m = -INFINITY;
#pragma unroll
    for (int e = 0; e < w.kElems; ++e) { w.v[e] *= in_scale; m = fmaxf(m, w.v[e]); }
    m = warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < w.kElems; ++e) { w.v[e] = expf(w.v[e] - m); s += w.v[e]; }
    s = warp_sum(s);
    const float inv = 1.f / s;
#pragma unroll
    for (int e = 0; e < w.kElems; ++e) w.v[e] *= inv;
    w.store(y + r * cols, cols, lane);
  }
}

// gx may alias g or y (MultiHeadAttention.backward writes dS over dP): a warp holds its whole row in registers before
// it stores, rows are private to a warp, and the loads are coherent and not __restrict__.
template <int N, bool VEC>
__global__ void __launch_bounds__(256) softmax_bwd_warp(const float* g, const float* y, float* gx, int64_t rows, int cols,
                                                        float out_scale, int diag_only) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    WarpRow<N, VEC> wy, wg;
    wy.load_aliasable(y + r * cols, cols, lane, 0.f);
    wg.load_aliasable(g + r * cols, cols, lane, 0.f);
    if (diag_only) {  // SoftmaxWithLoss.get_grad as shipped: g*y*(1-y)   tgrad.py:919
#pragma unroll
      for (int e = 0; e < wy.kElems; ++e) wg.v[e] = out_scale * wg.v[e] * wy.v[e] * (1.f - wy.v[e]);
    } else {          // dx = y*g; dx -= y*sum(dx)                         tgrad.py:948-950
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < wy.kElems; ++e) { wg.v[e] *= wy.v[e]; d += wg.v[e]; }
      d = warp_sum(d);
#pragma unroll
      for (int e = 0; e < wy.kElems; ++e) wg.v[e] = out_scale * (wg.v[e] - wy.v[e] * d);
    }
    wg.store(gx + r * cols, cols, lane);
  }
}

// ---- long rows: one CTA per row ------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_fwd_block(const float* __restrict__ x, float* __restrict__ y,
                                                         int64_t rows, int64_t cols, float in_scale) {
  __shared__ float red[32];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* xr = x + r * cols;
    float* yr = y + r * cols;
    float m = -INFINITY;
    for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) m = fmaxf(m, __ldg(xr + c) * in_scale);
    m = block_max(m, red);
    float s = 0.f;
    for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) {
      const float e = expf(__ldg(xr + c) * in_scale - m);
      yr[c] = e;
      s += e;
    }
    s = block_sum(s, red);
    const float inv = 1.f / s;
    for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) yr[c] *= inv;
  }
}

// gx may alias g or y: pass 1 only reads, the block-wide reduction orders it before pass 2, where every thread reads
// exactly the elements it then writes (coherent loads, no __restrict__).
__global__ void __launch_bounds__(256) softmax_bwd_block(const float* g, const float* y, float* gx, int64_t rows,
                                                         int64_t cols, float out_scale, int diag_only) {
  __shared__ float red[32];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* gr = g + r * cols;
    const float* yr = y + r * cols;
    float* o = gx + r * cols;
    float d = 0.f;
    if (!diag_only) {
      for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) d += gr[c] * yr[c];
      d = block_sum(d, red);
    }
    for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) {
      const float yy = yr[c], gg = gr[c];
      o[c] = diag_only ? out_scale * gg * yy * (1.f - yy) : out_scale * (gg * yy - yy * d);
    }
  }
}

// ---- fused softmax + cross entropy ------------------------------------------------------------------
// label handling follows NumPy fancy indexing on ys[arange, ts] (layers.py:530): -cols <= t < 0 wraps;
// t == ignore_label is masked out.  Out-of-range labels yield a NaN loss (NumPy would raise).
__device__ __forceinline__ int64_t wrap_label(int64_t t, int64_t cols) { return t < 0 ? t + cols : t; }

template <int N, bool VEC>
__global__ void __launch_bounds__(256) xent_fwd_warp(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                                     float* __restrict__ probs, float* __restrict__ row_loss,
                                                     int64_t rows, int cols, int64_t ignore_label, float eps,
                                                     int input_is_probs) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    WarpRow<N, VEC> w;
    w.load(logits + r * cols, cols, lane, input_is_probs ? 0.f : -INFINITY);
    const int64_t t_raw = labels[r];
    const bool masked = (t_raw == ignore_label);
    const int64_t t = wrap_label(t_raw, cols);
    float logp_t;
    if (!input_is_probs) {
      float m = -INFINITY;
#pragma unroll
      for (int e = 0; e < w.kElems; ++e) m = fmaxf(m, w.v[e]);
      m = warp_max(m);
      float s = 0.f, xt = 0.f;
#pragma unroll
      for (int e = 0; e < w.kElems; ++e) {
        if (w.col(e, lane) == t) xt = w.v[e] - m;
        w.v[e] = expf(w.v[e] - m);
        s += w.v[e];
      }
      s = warp_sum(s);
      xt = warp_sum(xt);  // exactly one lane holds it
      const float inv = 1.f / s;
#pragma unroll
      for (int e = 0; e < w.kElems; ++e) w.v[e] *= inv;
      if (probs) w.store(probs + r * cols, cols, lane);
      logp_t = (eps == 0.f) ? (xt - logf(s)) : logf(expf(xt) * inv + eps);
    } else {
      float pt = 0.f;
#pragma unroll
      for (int e = 0; e < w.kElems; ++e)
        if (w.col(e, lane) == t) pt = w.v[e];
      pt = warp_sum(pt);
      if (probs && probs != logits) w.store(probs + r * cols, cols, lane);
      logp_t = logf(pt + eps);
    }
    if (lane == 0) {
      float l = masked ? 0.f : -logp_t;
      if (!masked && (t < 0 || t >= cols)) l = NAN;
      row_loss[r] = l;
    }
  }
}

__global__ void __launch_bounds__(256) xent_fwd_block(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                                      float* __restrict__ probs, float* __restrict__ row_loss,
                                                      int64_t rows, int64_t cols, int64_t ignore_label, float eps,
                                                      int input_is_probs) {
  __shared__ float red[32];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* xr = logits + r * cols;
    float* pr = probs ? probs + r * cols : nullptr;
    const int64_t t_raw = labels[r];
    const bool masked = (t_raw == ignore_label);
    const int64_t t = wrap_label(t_raw, cols);
    const bool in_range = (t >= 0 && t < cols);
    float logp_t;
    if (!input_is_probs) {
      float m = -INFINITY;
      for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) m = fmaxf(m, __ldg(xr + c));
      m = block_max(m, red);
      float s = 0.f;
      for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) s += expf(__ldg(xr + c) - m);
      s = block_sum(s, red);
      const float inv = 1.f / s;
      if (pr)
        for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) pr[c] = expf(__ldg(xr + c) - m) * inv;
      const float xt = in_range ? __ldg(xr + t) - m : 0.f;
      logp_t = (eps == 0.f) ? (xt - logf(s)) : logf(expf(xt) * inv + eps);
    } else {
      if (pr && pr != xr)
        for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) pr[c] = __ldg(xr + c);
      logp_t = logf((in_range ? __ldg(xr + t) : 0.f) + eps);
    }
    if (threadIdx.x == 0) {
      float l = masked ? 0.f : -logp_t;
      if (!masked && !in_range) l = NAN;
      row_loss[r] = l;
    }
    __syncthreads();
  }
}

// loss_and_count[0] = sum(row_loss)/denom, [1] = denom  (single CTA, fixed order → reproducible)
__global__ void __launch_bounds__(1024) xent_finalize(const float* __restrict__ row_loss, const int64_t* __restrict__ labels,
                                                      float* __restrict__ loss_and_count, int64_t rows,
                                                      int64_t ignore_label, int denom_mode) {
  __shared__ float red[32];
  float s = 0.f, cnt = 0.f;
  for (int64_t r = threadIdx.x; r < rows; r += blockDim.x) {
    s += row_loss[r];
    cnt += (labels[r] != ignore_label) ? 1.f : 0.f;
  }
  s = block_sum(s, red);
  cnt = block_sum(cnt, red);
  if (threadIdx.x == 0) {
    const float denom = denom_mode == 0 ? cnt : (float)rows;
    loss_and_count[0] = s / denom;
    loss_and_count[1] = denom;
  }
}

// dx = (p - onehot) * dout / denom * mask     (TimeSoftmaxWithLoss.backward layers.py:538-549)
// dx may alias probs (TimeSoftmaxWithLoss.backward works in place on the cached probabilities, layers.py:541-545):
// every thread reads an element (coherent load, no __restrict__ on the pair) and then writes the same element.
__global__ void __launch_bounds__(256) xent_bwd_kernel(const float* probs, const int64_t* __restrict__ labels,
                                                       const float* __restrict__ loss_and_count, float* dx,
                                                       int64_t rows, int64_t cols, int64_t ignore_label, float dout,
                                                       int vec_ok) {
  const float k = dout / loss_and_count[1];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec_ok) {  // cols % 4 == 0
    // four float4 per thread and iteration, all loads before the first store: an in-place pass (dx == probs) otherwise
    // alternates single reads and writes to the same DRAM pages (ncu: 4.1 TB/s against 5.7+ for the out-of-place kernels)
    const int64_t n4 = (rows * cols) >> 2, c4 = cols >> 2;
    for (int64_t i0 = tid; i0 < n4; i0 += 4 * stride) {
      float4 p[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < n4) p[u] = ld_stream_coherent(reinterpret_cast<const float4*>(probs) + i);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = i0 + u * stride;
        if (i >= n4) break;
        const int64_t r = i / c4, c = (i - r * c4) << 2;
        const int64_t t_raw = labels[r];
        const int64_t t = wrap_label(t_raw, cols);
        float4 q = p[u];
        if (t_raw == ignore_label) {
          q = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          if (t == c) q.x -= 1.f;
          if (t == c + 1) q.y -= 1.f;
          if (t == c + 2) q.z -= 1.f;
          if (t == c + 3) q.w -= 1.f;
          q.x *= k; q.y *= k; q.z *= k; q.w *= k;
        }
        stg_stream(reinterpret_cast<float4*>(dx) + i, q);
      }
    }
  } else {
    const int64_t n = rows * cols;
    for (int64_t i = tid; i < n; i += stride) {
      const int64_t r = i / cols, c = i - r * cols;
      const int64_t t_raw = labels[r];
      const int64_t t = wrap_label(t_raw, cols);
      float p = probs[i];
      dx[i] = (t_raw == ignore_label) ? 0.f : (p - (t == c ? 1.f : 0.f)) * k;
    }
  }
}

// ---- argmax / argmin ------------------------------------------------------------------------------
// NumPy semantics: first occurrence of the extremum; a NaN beats everything (first NaN wins).
struct ArgPair { float v; int64_t i; };
template <bool MIN>
__device__ __forceinline__ bool arg_better(float v, int64_t i, float bv, int64_t bi) {
  if (bi < 0) return true;
  const bool vn = v != v, bn = bv != bv;
  if (vn || bn) return vn && (!bn || i < bi);
  if (v == bv) return i < bi;
  return MIN ? (v < bv) : (v > bv);
}
template <bool MIN>
__device__ __forceinline__ void warp_arg_reduce(float& v, int64_t& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int64_t oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (oi >= 0 && arg_better<MIN>(ov, oi, v, i)) { v = ov; i = oi; }
  }
}
template <bool MIN>
__device__ __forceinline__ void block_arg_reduce(float& v, int64_t& i, float* sv, int64_t* si) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  warp_arg_reduce<MIN>(v, i);
  __syncthreads();
  if (lane == 0) { sv[warp] = v; si[warp] = i; }
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = lane < nw ? sv[lane] : 0.f;
  i = lane < nw ? si[lane] : -1;
  warp_arg_reduce<MIN>(v, i);
}

template <bool MIN>
__global__ void __launch_bounds__(256) argext_stage1(const float* __restrict__ x, int64_t n, int64_t per_block,
                                                     int64_t* __restrict__ pidx, float* __restrict__ pval) {
  __shared__ float sv[32];
  __shared__ int64_t si[32];
  const int64_t b = (int64_t)blockIdx.x * per_block;
  const int64_t e = (b + per_block < n) ? b + per_block : n;
  float bv = 0.f;
  int64_t bi = -1;
  for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
    const float v = __ldg(x + i);
    if (arg_better<MIN>(v, i, bv, bi)) { bv = v; bi = i; }
  }
  block_arg_reduce<MIN>(bv, bi, sv, si);
  if (threadIdx.x == 0) { pidx[blockIdx.x] = bi; pval[blockIdx.x] = bv; }
}
template <bool MIN>
__global__ void __launch_bounds__(256) argext_stage2(const int64_t* __restrict__ pidx, const float* __restrict__ pval,
                                                     int nparts, int64_t* __restrict__ out) {
  __shared__ float sv[32];
  __shared__ int64_t si[32];
  float bv = 0.f;
  int64_t bi = -1;
  for (int p = threadIdx.x; p < nparts; p += blockDim.x) {
    const int64_t i = pidx[p];
    const float v = pval[p];
    if (i >= 0 && arg_better<MIN>(v, i, bv, bi)) { bv = v; bi = i; }
  }
  block_arg_reduce<MIN>(bv, bi, sv, si);
  if (threadIdx.x == 0) out[0] = bi < 0 ? 0 : bi;
}

__global__ void __launch_bounds__(256) argmax_rows_kernel(const float* __restrict__ x, int64_t rows, int64_t cols,
                                                          int64_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const float* xr = x + r * cols;
    float bv = 0.f;
    int64_t bi = -1;
    for (int64_t c = lane; c < cols; c += 32) {
      const float v = __ldg(xr + c);
      if (arg_better<false>(v, c, bv, bi)) { bv = v; bi = c; }
    }
    warp_arg_reduce<false>(bv, bi);
    if (lane == 0) out[r] = bi < 0 ? 0 : bi;
  }
}

}  // namespace b200ad

using namespace b200ad;

#define B200AD_ARG_PARTS 1184

// Picks the register-resident warp kernel for a given row length.
#define DISPATCH_WARP_ROWS(KERNEL, cols, vec, grid, s, ...)                                   \
  do {                                                                                        \
    if (vec) {                                                                                \
      if (cols <= 128) KERNEL<1, true><<<grid, 256, 0, s>>>(__VA_ARGS__);                     \
      else if (cols <= 256) KERNEL<2, true><<<grid, 256, 0, s>>>(__VA_ARGS__);                \
      else if (cols <= 512) KERNEL<4, true><<<grid, 256, 0, s>>>(__VA_ARGS__);                \
      else KERNEL<8, true><<<grid, 256, 0, s>>>(__VA_ARGS__);                                 \
    } else {                                                                                  \
      if (cols <= 32) KERNEL<1, false><<<grid, 256, 0, s>>>(__VA_ARGS__);                     \
      else if (cols <= 64) KERNEL<2, false><<<grid, 256, 0, s>>>(__VA_ARGS__);                \
      else if (cols <= 128) KERNEL<4, false><<<grid, 256, 0, s>>>(__VA_ARGS__);               \
      else if (cols <= 256) KERNEL<8, false><<<grid, 256, 0, s>>>(__VA_ARGS__);               \
      else if (cols <= 512) KERNEL<16, false><<<grid, 256, 0, s>>>(__VA_ARGS__);              \
      else KERNEL<32, false><<<grid, 256, 0, s>>>(__VA_ARGS__);                               \
    }                                                                                         \
  } while (0)

extern "C" {

int b200ad_softmax_fwd(const float* x, float* y, int64_t rows, int64_t cols, float in_scale, void* stream) {
  B200AD_REQUIRE(rows >= 0 && cols >= 0, "softmax_fwd: negative shape");
  if (rows == 0 || cols == 0) return B200AD_OK;
  B200AD_REQUIRE(x && y, "softmax_fwd: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  if (cols <= 1024) {
    const bool vec = (cols % 4 == 0) && aligned16(x) && aligned16(y);
    const int grid = stream_grid(rows, 8);
    const int c = (int)cols;
    DISPATCH_WARP_ROWS(softmax_fwd_warp, c, vec, grid, s, x, y, rows, c, in_scale);
  } else {
    softmax_fwd_block<<<stream_grid(rows, 1), 256, 0, s>>>(x, y, rows, cols, in_scale);
  }
  count_launch();
  return check_launch("softmax_fwd");
}

int b200ad_softmax_bwd(const float* g, const float* y, float* gx, int64_t rows, int64_t cols, float out_scale,
                       int diag_only, void* stream) {
  B200AD_REQUIRE(rows >= 0 && cols >= 0, "softmax_bwd: negative shape");
  if (rows == 0 || cols == 0) return B200AD_OK;
  B200AD_REQUIRE(g && y && gx, "softmax_bwd: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  if (cols <= 1024) {
    const bool vec = (cols % 4 == 0) && aligned16(g) && aligned16(y) && aligned16(gx);
    const int grid = stream_grid(rows, 8);
    const int c = (int)cols;
    DISPATCH_WARP_ROWS(softmax_bwd_warp, c, vec, grid, s, g, y, gx, rows, c, out_scale, diag_only);
  } else {
    softmax_bwd_block<<<stream_grid(rows, 1), 256, 0, s>>>(g, y, gx, rows, cols, out_scale, diag_only);
  }
  count_launch();
  return check_launch("softmax_bwd");
}

int b200ad_xent_fwd(const float* logits, const int64_t* labels, float* probs, float* row_loss, float* loss_and_count,
                    int64_t rows, int64_t cols, int64_t ignore_label, float eps, int denom_mode, int input_is_probs,
                    void* stream) {
  B200AD_REQUIRE(rows > 0 && cols > 0, "xent_fwd: empty input (%lld x %lld)", (long long)rows, (long long)cols);
  B200AD_REQUIRE(logits && labels && row_loss && loss_and_count, "xent_fwd: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  if (cols <= 1024) {
    const bool vec = (cols % 4 == 0) && aligned16(logits) && (!probs || aligned16(probs));
    const int grid = stream_grid(rows, 8);
    const int c = (int)cols;
    DISPATCH_WARP_ROWS(xent_fwd_warp, c, vec, grid, s, logits, labels, probs, row_loss, rows, c, ignore_label, eps,
                       input_is_probs);
  } else {
    xent_fwd_block<<<stream_grid(rows, 1), 256, 0, s>>>(logits, labels, probs, row_loss, rows, cols, ignore_label,
                                                         eps, input_is_probs);
  }
  xent_finalize<<<1, 1024, 0, s>>>(row_loss, labels, loss_and_count, rows, ignore_label, denom_mode);
  count_launch(2);
  return check_launch("xent_fwd");
}

int b200ad_xent_bwd(const float* probs, const int64_t* labels, const float* loss_and_count, float* dx, int64_t rows,
                    int64_t cols, int64_t ignore_label, float dout, int denom_mode, void* stream) {
  (void)denom_mode;  // the denominator chosen in the forward is read back from loss_and_count[1]
  B200AD_REQUIRE(rows > 0 && cols > 0, "xent_bwd: empty input");
  B200AD_REQUIRE(probs && labels && loss_and_count && dx, "xent_bwd: null pointer");
  const int vec_ok = (cols % 4 == 0) && aligned16(probs) && aligned16(dx);
  const int64_t n = rows * cols;
  xent_bwd_kernel<<<stream_grid(vec_ok ? n / 4 : n, 256), 256, 0, (cudaStream_t)stream>>>(
      probs, labels, loss_and_count, dx, rows, cols, ignore_label, dout, vec_ok);
  count_launch();
  return check_launch("xent_bwd");
}

int b200ad_argext_flat(const float* x, int64_t n, int is_min, int64_t* out, void* ws, void* stream) {
  B200AD_REQUIRE(n > 0, "argmax: attempt to get argmax of an empty sequence");
  B200AD_REQUIRE(x && out && ws, "argext_flat: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  int64_t* pidx = reinterpret_cast<int64_t*>(ws);
  float* pval = reinterpret_cast<float*>(pidx + B200AD_ARG_PARTS);
  int64_t per_block = ceil_div(n, B200AD_ARG_PARTS);
  if (per_block < 1024) per_block = 1024;
  const int nparts = (int)ceil_div(n, per_block);
  if (is_min) {
    argext_stage1<true><<<nparts, 256, 0, s>>>(x, n, per_block, pidx, pval);
    argext_stage2<true><<<1, 256, 0, s>>>(pidx, pval, nparts, out);
  } else {
    argext_stage1<false><<<nparts, 256, 0, s>>>(x, n, per_block, pidx, pval);
    argext_stage2<false><<<1, 256, 0, s>>>(pidx, pval, nparts, out);
  }
  count_launch(2);
  return check_launch("argext_flat");
}

int b200ad_argmax_rows(const float* x, int64_t rows, int64_t cols, int64_t* out, void* stream) {
  B200AD_REQUIRE(rows >= 0 && cols > 0, "argmax_rows: bad shape");
  if (rows == 0) return B200AD_OK;
  B200AD_REQUIRE(x && out, "argmax_rows: null pointer");
  argmax_rows_kernel<<<stream_grid(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, rows, cols, out);
  count_launch();
  return check_launch("argmax_rows");
}
}
