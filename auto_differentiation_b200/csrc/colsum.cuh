// Deterministic two-stage "sum over the middle axis" for wide inner extents:
//   out[o, i] = scale * sum_r value(o, r, i),  value produced by a functor (a plain load, or a fused
//   elementwise backward that also writes its full-shape output).
// This is the broadcast-reduce behind check_shape (tgrad.py:2163-2187) and the db = sum(rows)
// reductions of layers.py:306,:674,:1264.  Stage 1: grid (col-blocks, slabs, outer), a warp covers
// 32*VEC consecutive columns (coalesced 512 B with VEC=4), 8 warps stride over the slab's rows with
// four row-loads in flight; cross-warp combine in shared memory.  Stage 2 folds the slab partials.
// No atomics: results are run-to-run reproducible.
#pragma once
#include "common.cuh"

namespace b200ad {

struct LoadProducer {
  const float* in;
  __device__ __forceinline__ float4 get4(int64_t off, int64_t) const {
    return ldg_stream(reinterpret_cast<const float4*>(in + off));
  }
  __device__ __forceinline__ float get1(int64_t off, int64_t) const { return __ldg(in + off); }
};

template <int VEC, class P>
__global__ void __launch_bounds__(256) colsum_wide_kernel(P prod, float* __restrict__ out, int64_t red,
                                                          int64_t inner, int64_t rows_per_slab,
                                                          float scale) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t col = ((int64_t)blockIdx.x * 32 + tx) * VEC;
  const int64_t slab = blockIdx.y, o = blockIdx.z, nslabs = gridDim.y;
  const int64_t r0 = slab * rows_per_slab;
  const int64_t r1 = (r0 + rows_per_slab < red) ? r0 + rows_per_slab : red;
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  if (col < inner) {
    const int64_t base = o * red * inner + col;
    int64_t r = r0 + ty;
    if constexpr (VEC == 4) {
      for (; r + 24 < r1; r += 32) {
        float4 v0 = prod.get4(base + r * inner, col);
        float4 v1 = prod.get4(base + (r + 8) * inner, col);
        float4 v2 = prod.get4(base + (r + 16) * inner, col);
        float4 v3 = prod.get4(base + (r + 24) * inner, col);
        acc[0] += (v0.x + v1.x) + (v2.x + v3.x);
        acc[1] += (v0.y + v1.y) + (v2.y + v3.y);
        acc[2] += (v0.z + v1.z) + (v2.z + v3.z);
        acc[3] += (v0.w + v1.w) + (v2.w + v3.w);
      }
      for (; r < r1; r += 8) {
        float4 v = prod.get4(base + r * inner, col);
        acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
      }
    } else {
      for (; r + 24 < r1; r += 32) {
        float v0 = prod.get1(base + r * inner, col);
        float v1 = prod.get1(base + (r + 8) * inner, col);
        float v2 = prod.get1(base + (r + 16) * inner, col);
        float v3 = prod.get1(base + (r + 24) * inner, col);
        acc[0] += (v0 + v1) + (v2 + v3);
      }
      for (; r < r1; r += 8) acc[0] += prod.get1(base + r * inner, col);
    }
  }
  __shared__ float sm[8][32 * VEC + 1];
#pragma unroll
  for (int j = 0; j < VEC; ++j) sm[ty][tx * VEC + j] = acc[j];
  __syncthreads();
  if (ty == 0 && col < inner) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += sm[w][tx * VEC + j];
      out[(o * nslabs + slab) * inner + col + j] = s * scale;
    }
  }
}

// Slab count for stage 1: fill ~148 x 8 CTAs but keep >= 64 rows per slab.
inline int64_t colsum_slabs(int64_t outer, int64_t red, int64_t inner, int vec) {
  const int64_t colblocks = ceil_div(inner, 32 * vec);
  int64_t target = (int64_t)sm_count() * 8 / (colblocks * outer > 0 ? colblocks * outer : 1);
  if (target < 1) target = 1;
  int64_t max_slabs = red / 64;
  if (max_slabs < 1) max_slabs = 1;
  int64_t s = target < max_slabs ? target : max_slabs;
  if (s > 4096) s = 4096;
  return s;
}

// Launches stage 1 (+ stage 2 when slabs > 1). `ws` needs outer*slabs*inner floats when slabs > 1.
template <class P>
int colsum_launch(P prod, float* out, int64_t outer, int64_t red, int64_t inner, float scale,
                  float* ws, bool vec, cudaStream_t s) {
  if (outer > 65535) return set_error(B200AD_ERR_UNSUPPORTED, "colsum: outer %lld > 65535", (long long)outer);
  const int V = vec ? 4 : 1;
  const int64_t colblocks = ceil_div(inner, 32 * V);
  const int64_t slabs = colsum_slabs(outer, red, inner, V);
  const int64_t rows_per_slab = ceil_div(red, slabs);
  dim3 grid((unsigned)colblocks, (unsigned)slabs, (unsigned)outer);
  float* stage1_out = slabs > 1 ? ws : out;
  const float stage1_scale = slabs > 1 ? 1.f : scale;
  if (slabs > 1 && !ws) return set_error(B200AD_ERR_INVALID, "colsum: workspace required");
  if (vec) colsum_wide_kernel<4, P><<<grid, 256, 0, s>>>(prod, stage1_out, red, inner, rows_per_slab, stage1_scale);
  else colsum_wide_kernel<1, P><<<grid, 256, 0, s>>>(prod, stage1_out, red, inner, rows_per_slab, stage1_scale);
  count_launch();
  if (slabs > 1) {
    LoadProducer lp{ws};
    dim3 grid2((unsigned)colblocks, 1, (unsigned)outer);
    const bool vec2 = vec && aligned16(ws);
    if (vec2) colsum_wide_kernel<4, LoadProducer><<<grid2, 256, 0, s>>>(lp, out, slabs, inner, slabs, scale);
    else {
      dim3 grid2s((unsigned)ceil_div(inner, 32), 1, (unsigned)outer);
      colsum_wide_kernel<1, LoadProducer><<<grid2s, 256, 0, s>>>(lp, out, slabs, inner, slabs, scale);
    }
    count_launch();
  }
  return check_launch("colsum");
}

}  // namespace b200ad
