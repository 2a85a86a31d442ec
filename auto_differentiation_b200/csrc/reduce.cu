// Reductions over the middle axis of an (outer, red, inner) view, axis repeat, fill, axpby, and the
// strided->contiguous copy.  Reference bodies replaced: check_shape tgrad.py:2163-2187,
// sum.result/get_grad :692-721, repeat :774-797, transpose.result :611-613, tensordot's
// transpose+reshape :294-303, TimeAffine db layers.py:306, LSTM db :674.
#include "common.cuh"
#include "colsum.cuh"

namespace b200ad {

// ---- inner == 1: one warp per (virtual) row, shuffle reduction ------------------------------------
// A row of length `len` is cut into `nchunks` chunks of `chunk` elements (chunk % 4 == 0) so that a
// few very long rows still fill the machine; virtual row v = row * nchunks + chunk_id.
template <int VEC>
__global__ void __launch_bounds__(256) rowsum_warp_kernel(const float* __restrict__ in,
                                                          float* __restrict__ out, int64_t vrows,
                                                          int64_t len, int64_t chunk, int64_t nchunks,
                                                          float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < vrows; v += warps) {
    const int64_t row = v / nchunks, ck = v - row * nchunks;
    const int64_t b = ck * chunk;
    const int64_t e = (b + chunk < len) ? b + chunk : len;
    const float* p = in + row * len;
    float a0 = 0.f, a1 = 0.f;
    if constexpr (VEC == 4) {
      const float4* p4 = reinterpret_cast<const float4*>(p);
      const int64_t b4 = b >> 2, e4 = e >> 2;  // b % 4 == 0 by construction
      int64_t i = b4 + lane;
      for (; i + 32 < e4; i += 64) {
        float4 x = ldg_stream(p4 + i), y = ldg_stream(p4 + i + 32);
        a0 += (x.x + x.y) + (x.z + x.w);
        a1 += (y.x + y.y) + (y.z + y.w);
      }
      for (; i < e4; i += 32) {
        float4 x = ldg_stream(p4 + i);
        a0 += (x.x + x.y) + (x.z + x.w);
      }
      for (int64_t t = (e4 << 2) + lane; t < e; t += 32) a1 += __ldg(p + t);
    } else {
      for (int64_t i = b + lane; i < e; i += 32) a0 += __ldg(p + i);
    }
    float s = warp_sum(a0 + a1);
    if (lane == 0) out[v] = s * scale;
  }
}

// ---- 1 < inner < 32: flat, coalesced; each thread keeps a fixed column ---------------------------
__global__ void __launch_bounds__(256) reduce_narrow_kernel(const float* __restrict__ in,
                                                            float* __restrict__ out, int64_t red,
                                                            int inner, int64_t rows_per_slab,
                                                            float scale) {
  const int tid = threadIdx.x;
  const int k = 256 / inner;  // rows per pass
  const int L = k * inner;    // active threads
  const int64_t slab = blockIdx.y, o = blockIdx.z, nslabs = gridDim.y;
  const int64_t r0 = slab * rows_per_slab;
  const int64_t r1 = (r0 + rows_per_slab < red) ? r0 + rows_per_slab : red;
  float acc = 0.f;
  if (tid < L) {
    const int c = tid % inner, rr = tid / inner;
    const float* p = in + (o * red) * inner + c;
    for (int64_t r = r0 + rr; r < r1; r += k) acc += __ldg(p + r * inner);
  }
  __shared__ float sm[256];
  sm[tid] = acc;
  __syncthreads();
  if (tid < inner) {
    float s = 0.f;
    for (int j = 0; j < k; ++j) s += sm[j * inner + tid];
    out[(o * nslabs + slab) * inner + tid] = s * scale;
  }
}

__global__ void __launch_bounds__(256) repeat_axis_kernel(const float* __restrict__ in,
                                                          float* __restrict__ out, int64_t total,
                                                          int64_t rep, int64_t inner, float scale,
                                                          int vec_ok) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec_ok) {  // inner % 4 == 0
    const int64_t n4 = total >> 2, inner4 = inner >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      const int64_t o = i / (rep * inner4), c = i % inner4;
      float4 v = __ldg(reinterpret_cast<const float4*>(in) + o * inner4 + c);
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      stg_stream(reinterpret_cast<float4*>(out) + i, v);
    }
  } else {
    for (int64_t i = tid; i < total; i += stride) {
      const int64_t o = i / (rep * inner), c = i % inner;
      out[i] = __ldg(in + o * inner + c) * scale;
    }
  }
}

__global__ void __launch_bounds__(256) fill_kernel(float* __restrict__ out, int64_t n, float v) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = v;
}

__global__ void __launch_bounds__(256) axpby_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                    int64_t n, float a, float b, int vec_ok) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      float4 xv = ldg_stream(reinterpret_cast<const float4*>(x) + i);
      float4 yv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b != 0.f) yv = reinterpret_cast<const float4*>(y)[i];
      yv.x = a * xv.x + b * yv.x; yv.y = a * xv.y + b * yv.y;
      yv.z = a * xv.z + b * yv.z; yv.w = a * xv.w + b * yv.w;
      reinterpret_cast<float4*>(y)[i] = yv;
    }
    const int64_t t = (n4 << 2) + tid;
    if (t < n) y[t] = a * x[t] + (b != 0.f ? b * y[t] : 0.f);
  } else {
    for (int64_t i = tid; i < n; i += stride) y[i] = a * __ldg(x + i) + (b != 0.f ? b * y[i] : 0.f);
  }
}

// TimeDropout forward (reference layers.py:685-692): mask = (u > ratio) * scale, y = x * mask; the uniform draws u are
// the caller's (device RNG, or the reference's host stream for parity runs).
__global__ void __launch_bounds__(256) dropout_fwd_kernel(const float* __restrict__ x, const float* __restrict__ u,
                                                          float* __restrict__ mask, float* __restrict__ y, int64_t n,
                                                          float ratio, float scale, int vec_ok) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 xv = ldg_stream(reinterpret_cast<const float4*>(x) + i);
      const float4 uv = ldg_stream(reinterpret_cast<const float4*>(u) + i);
      float4 mv;
      mv.x = uv.x > ratio ? scale : 0.f; mv.y = uv.y > ratio ? scale : 0.f;
      mv.z = uv.z > ratio ? scale : 0.f; mv.w = uv.w > ratio ? scale : 0.f;
      reinterpret_cast<float4*>(mask)[i] = mv;
      reinterpret_cast<float4*>(y)[i] = make_float4(xv.x * mv.x, xv.y * mv.y, xv.z * mv.z, xv.w * mv.w);
    }
    const int64_t t = (n4 << 2) + tid;
    if (t < n) {
      const float m = u[t] > ratio ? scale : 0.f;
      mask[t] = m;
      y[t] = x[t] * m;
    }
  } else {
    for (int64_t i = tid; i < n; i += stride) {
      const float m = __ldg(u + i) > ratio ? scale : 0.f;
      mask[i] = m;
      y[i] = __ldg(x + i) * m;
    }
  }
}

// U[0,1) draws for TimeDropout (the reference draws np.random.rand, layers.py:687): counter-based Philox4x32-10, four
// floats per counter, so the stream depends only on (seed, offset) — not on the grid — and consecutive calls continue
// it by advancing `offset` by ceil(n / 4) counters.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__global__ void __launch_bounds__(256) uniform_kernel(float* __restrict__ out, int64_t n, uint64_t seed, uint64_t offset) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = (n + 3) >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint64_t c = offset + (uint64_t)i;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    // 24 random bits -> [0, 1) exactly representable in fp32
    const float f[4] = {(r.x >> 8) * 5.9604644775390625e-8f, (r.y >> 8) * 5.9604644775390625e-8f,
                        (r.z >> 8) * 5.9604644775390625e-8f, (r.w >> 8) * 5.9604644775390625e-8f};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * i + j < n) out[4 * i + j] = f[j];
  }
}

struct StridedDesc {
  int rank;
  int out_strided;
  int64_t shape[6];
  int64_t stride[6];
  int64_t ostride[6];
};

// Generic gather: one thread per output element (writes coalesced; reads coalesced whenever the
// innermost output axis has input stride 1, which is the case for the head split/merge permutes).
__global__ void __launch_bounds__(256) strided_copy_kernel(const float* __restrict__ in,
                                                           float* __restrict__ out, int64_t total,
                                                           StridedDesc d, int vec_ok) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec_ok) {  // innermost axis: input stride 1, extent % 4 == 0, bases and strides 16B aligned
    const int64_t n4 = total >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      int64_t rem = i << 2, off = 0;
#pragma unroll
      for (int a = 5; a >= 0; --a) {
        if (a < d.rank) {
          const int64_t q = rem / d.shape[a];
          off += (rem - q * d.shape[a]) * d.stride[a];
          rem = q;
        }
      }
      stg_stream(reinterpret_cast<float4*>(out) + i, __ldg(reinterpret_cast<const float4*>(in + off)));
    }
  } else {
    for (int64_t i = tid; i < total; i += stride) {
      int64_t rem = i, off = 0, ooff = 0;
#pragma unroll
      for (int a = 5; a >= 0; --a) {
        if (a < d.rank) {
          const int64_t q = rem / d.shape[a];
          const int64_t c = rem - q * d.shape[a];
          off += c * d.stride[a];
          ooff += c * d.ostride[a];
          rem = q;
        }
      }
      out[d.out_strided ? ooff : i] = __ldg(in + off);
    }
  }
}

// 2-D tiled transpose for the case where the innermost OUTPUT axis is strided in the input and some
// other axis is contiguous in the input: batch x (R x C) -> (C x R) through a padded smem tile.
__global__ void __launch_bounds__(256) transpose2d_kernel(const float* __restrict__ in,
                                                          float* __restrict__ out, int64_t rows,
                                                          int64_t cols, int64_t in_batch_stride,
                                                          int64_t in_row_stride) {
  // in[b][r][c] at in + b*in_batch_stride + r*in_row_stride + c ; out[b][c][r] contiguous
  __shared__ float tile[32][33];
  const int64_t b = blockIdx.z;
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* ib = in + b * in_batch_stride;
  float* ob = out + b * rows * cols;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int64_t r = r0 + ty + j, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + j][tx] = __ldg(ib + r * in_row_stride + c);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int64_t c = c0 + ty + j, r = r0 + tx;
    if (r < rows && c < cols) ob[c * rows + r] = tile[tx][ty + j];
  }
}

}  // namespace b200ad

using namespace b200ad;

static int64_t narrow_slabs(int64_t outer, int64_t red, int64_t inner) {
  const int64_t k = 256 / inner;
  int64_t target = (int64_t)sm_count() * 8 / (outer > 0 ? outer : 1);
  if (target < 1) target = 1;
  int64_t max_slabs = red / (8 * k);
  if (max_slabs < 1) max_slabs = 1;
  int64_t s = target < max_slabs ? target : max_slabs;
  if (s > 4096) s = 4096;
  return s;
}

static void rowsum_plan(int64_t rows, int64_t len, int64_t* chunk, int64_t* nchunks) {
  // want ~148*8*8 warps busy; only split rows that are long.
  const int64_t want = (int64_t)sm_count() * 64;
  int64_t nc = 1;
  if (rows < want && len >= 8192) {
    nc = want / (rows > 0 ? rows : 1);
    const int64_t maxc = len / 2048;
    if (nc > maxc) nc = maxc;
    if (nc < 1) nc = 1;
  }
  int64_t ck = ceil_div(len, nc);
  ck = (ck + 3) & ~(int64_t)3;
  if (ck < 4) ck = 4;
  *chunk = ck;
  *nchunks = ceil_div(len, ck);
}

extern "C" {

int64_t b200ad_colsum_ws_floats(int64_t rows, int64_t cols) {
  int64_t s4 = colsum_slabs(1, rows, cols, 4), s1 = colsum_slabs(1, rows, cols, 1);
  int64_t s = s4 > s1 ? s4 : s1;
  return s > 1 ? s * cols : 0;
}

int64_t b200ad_reduce_ws_floats(int64_t outer, int64_t red, int64_t inner) {
  if (outer <= 0 || red <= 1 || inner <= 0) return 0;
  if (inner == 1) {
    int64_t ck, nc;
    rowsum_plan(outer, red, &ck, &nc);
    return nc > 1 ? outer * nc : 0;
  }
  if (inner < 32) {
    int64_t s = narrow_slabs(outer, red, inner);
    return s > 1 ? outer * s * inner : 0;
  }
  const int64_t o = outer > 65535 ? 65535 : outer;
  int64_t s4 = colsum_slabs(o, red, inner, 4), s1 = colsum_slabs(o, red, inner, 1);
  int64_t s = s4 > s1 ? s4 : s1;
  return s > 1 ? o * s * inner : 0;
}

int b200ad_reduce(const float* in, float* out, int64_t outer, int64_t red, int64_t inner, float scale,
                  float* ws, void* stream) {
  B200AD_REQUIRE(outer >= 0 && red >= 0 && inner >= 0, "reduce: negative extent");
  if (outer == 0 || inner == 0) return B200AD_OK;
  B200AD_REQUIRE(out, "reduce: null output");
  cudaStream_t s = (cudaStream_t)stream;
  if (red == 0) {
    fill_kernel<<<stream_grid(outer * inner, 256), 256, 0, s>>>(out, outer * inner, 0.f);
    count_launch();
    return check_launch("reduce(fill)");
  }
  B200AD_REQUIRE(in, "reduce: null input");
  if (red == 1) {
    const int64_t n = outer * inner;
    const int vec_ok = aligned16(in) && aligned16(out);
    axpby_kernel<<<stream_grid(vec_ok ? (n + 3) / 4 : n, 256), 256, 0, s>>>(in, out, n, scale, 0.f, vec_ok);
    count_launch();
    return check_launch("reduce(copy)");
  }
  if (inner == 1) {
    int64_t ck, nc;
    rowsum_plan(outer, red, &ck, &nc);
    const bool vec = aligned16(in) && (red % 4 == 0);
    const int64_t vrows = outer * nc;
    float* o1 = nc > 1 ? ws : out;
    B200AD_REQUIRE(nc == 1 || ws, "reduce: workspace required");
    const int grid = stream_grid(vrows, 8);
    if (vec) rowsum_warp_kernel<4><<<grid, 256, 0, s>>>(in, o1, vrows, red, ck, nc, nc > 1 ? 1.f : scale);
    else rowsum_warp_kernel<1><<<grid, 256, 0, s>>>(in, o1, vrows, red, ck, nc, nc > 1 ? 1.f : scale);
    count_launch();
    if (nc > 1) {
      rowsum_warp_kernel<1><<<stream_grid(outer, 8), 256, 0, s>>>(ws, out, outer, nc, (nc + 3) & ~(int64_t)3, 1, scale);
      count_launch();
    }
    return check_launch("reduce(rowsum)");
  }
  if (inner < 32) {
    B200AD_REQUIRE(outer <= 65535, "reduce(narrow): outer %lld > 65535", (long long)outer);
    const int64_t slabs = narrow_slabs(outer, red, inner);
    const int64_t rps = ceil_div(red, slabs);
    B200AD_REQUIRE(slabs == 1 || ws, "reduce: workspace required");
    dim3 grid(1, (unsigned)slabs, (unsigned)outer);
    reduce_narrow_kernel<<<grid, 256, 0, s>>>(in, slabs > 1 ? ws : out, red, (int)inner, rps, slabs > 1 ? 1.f : scale);
    count_launch();
    if (slabs > 1) {
      dim3 grid2(1, 1, (unsigned)outer);
      reduce_narrow_kernel<<<grid2, 256, 0, s>>>(ws, out, slabs, (int)inner, slabs, scale);
      count_launch();
    }
    return check_launch("reduce(narrow)");
  }
  const bool vec = (inner % 4 == 0) && aligned16(in);
  for (int64_t o0 = 0; o0 < outer; o0 += 65535) {
    const int64_t oc = (outer - o0 < 65535) ? outer - o0 : 65535;
    LoadProducer lp{in + o0 * red * inner};
    int rc = colsum_launch(lp, out + o0 * inner, oc, red, inner, scale, ws, vec, s);
    if (rc) return rc;
  }
  return B200AD_OK;
}

int b200ad_repeat_axis(const float* in, float* out, int64_t outer, int64_t rep, int64_t inner,
                       float scale, void* stream) {
  B200AD_REQUIRE(outer >= 0 && rep >= 0 && inner >= 0, "repeat_axis: negative extent");
  const int64_t total = outer * rep * inner;
  if (total == 0) return B200AD_OK;
  B200AD_REQUIRE(in && out, "repeat_axis: null pointer");
  const int vec_ok = (inner % 4 == 0) && aligned16(in) && aligned16(out);
  repeat_axis_kernel<<<stream_grid(vec_ok ? total / 4 : total, 256), 256, 0, (cudaStream_t)stream>>>(
      in, out, total, rep, inner, scale, vec_ok);
  count_launch();
  return check_launch("repeat_axis");
}

int b200ad_fill(float* out, int64_t n, float value, void* stream) {
  B200AD_REQUIRE(n >= 0, "fill: negative n");
  if (n == 0) return B200AD_OK;
  B200AD_REQUIRE(out, "fill: null pointer");
  fill_kernel<<<stream_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(out, n, value);
  count_launch();
  return check_launch("fill");
}

int b200ad_axpby(const float* x, float* y, int64_t n, float a, float b, void* stream) {
  B200AD_REQUIRE(n >= 0, "axpby: negative n");
  if (n == 0) return B200AD_OK;
  B200AD_REQUIRE(x && y, "axpby: null pointer");
  const int vec_ok = aligned16(x) && aligned16(y);
  axpby_kernel<<<stream_grid(vec_ok ? (n + 3) / 4 : n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n, a, b, vec_ok);
  count_launch();
  return check_launch("axpby");
}

int b200ad_dropout_fwd(const float* x, const float* u, float* mask, float* y, int64_t n, float ratio,
                       float scale, void* stream) {
  B200AD_REQUIRE(n >= 0, "dropout_fwd: negative n");
  if (n == 0) return B200AD_OK;
  B200AD_REQUIRE(x && u && mask && y, "dropout_fwd: null pointer");
  const int vec_ok = aligned16(x) && aligned16(u) && aligned16(mask) && aligned16(y);
  dropout_fwd_kernel<<<stream_grid(vec_ok ? (n + 3) / 4 : n, 256), 256, 0, (cudaStream_t)stream>>>(x, u, mask, y, n, ratio,
                                                                                                  scale, vec_ok);
  count_launch();
  return check_launch("dropout_fwd");
}

int b200ad_uniform(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream) {
  B200AD_REQUIRE(n >= 0, "uniform: negative size");
  if (n == 0) return B200AD_OK;
  B200AD_REQUIRE(out, "uniform: null pointer");
  uniform_kernel<<<stream_grid((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(out, n, seed, offset);
  count_launch();
  return check_launch("uniform");
}

int b200ad_strided_copy(const float* in, float* out, int rank, const int64_t* shape,
                        const int64_t* in_strides, const int64_t* out_strides, void* stream) {
  B200AD_REQUIRE(rank >= 0 && rank <= 6, "strided_copy: rank %d not in [0,6]", rank);
  StridedDesc d;
  d.rank = rank;
  d.out_strided = out_strides != nullptr;
  int64_t total = 1;
  for (int a = 0; a < 6; ++a) {
    d.shape[a] = a < rank ? shape[a] : 1;
    d.stride[a] = a < rank ? in_strides[a] : 0;
    d.ostride[a] = (a < rank && out_strides) ? out_strides[a] : 0;
    B200AD_REQUIRE(d.shape[a] >= 0, "strided_copy: negative extent");
    total *= d.shape[a];
  }
  if (total == 0) return B200AD_OK;
  B200AD_REQUIRE(in && out, "strided_copy: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  if (rank == 0) {
    d.rank = 1; d.shape[0] = 1; d.stride[0] = 1; d.ostride[0] = 1;
  }
  // Fast path: pure (batch, R, C) -> (batch, C, R) transposes where the input rows are contiguous.
  if (!d.out_strided && d.rank >= 2 && d.stride[d.rank - 1] != 1 && d.stride[d.rank - 2] == 1) {
    // out axes (..., C, R): out inner axis R has input stride = in_row_stride, C is contiguous in input
    bool batch_ok = true;
    int64_t nb = 1, bstride = 0;
    if (d.rank == 3) { nb = d.shape[0]; bstride = d.stride[0]; }
    else if (d.rank > 3) {
      // leading axes must collapse into one batch axis with a uniform stride
      nb = 1;
      bstride = d.stride[d.rank - 3];
      int64_t expect = bstride;
      for (int a = d.rank - 3; a >= 0; --a) {
        if (d.shape[a] != 1 && d.stride[a] != expect) { batch_ok = false; break; }
        expect *= d.shape[a];
        nb *= d.shape[a];
      }
    }
    if (batch_ok && nb <= 65535) {
      const int64_t C = d.shape[d.rank - 2], R = d.shape[d.rank - 1];
      const int64_t in_row_stride = d.stride[d.rank - 1];
      if (ceil_div(R, 32) <= 65535) {
        dim3 grid((unsigned)ceil_div(C, 32), (unsigned)ceil_div(R, 32), (unsigned)nb);
        transpose2d_kernel<<<grid, 256, 0, s>>>(in, out, R, C, bstride, in_row_stride);
        count_launch();
        return check_launch("strided_copy(transpose2d)");
      }
    }
  }
  int vec_ok = !d.out_strided && aligned16(in) && aligned16(out) && d.stride[d.rank - 1] == 1 &&
               (d.shape[d.rank - 1] % 4 == 0);
  for (int a = 0; a < d.rank - 1 && vec_ok; ++a)
    if (d.stride[a] % 4 != 0) vec_ok = 0;
  strided_copy_kernel<<<stream_grid(vec_ok ? total / 4 : total, 256), 256, 0, s>>>(in, out, total, d, vec_ok);
  count_launch();
  return check_launch("strided_copy");
}
}
