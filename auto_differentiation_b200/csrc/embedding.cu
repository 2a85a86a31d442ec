// Embedding gather and deterministic scatter-add (integer index work is bit-exact).
// Reference bodies replaced: Embedding.forward `W[idx]` layers.py:448-452, Embedding.backward
// `dW[...] = 0; np.add.at(dW, idx, dout)` :454-458, TimeEmbedding :461-507.
#include "common.cuh"

namespace b200ad {

// NumPy index semantics: -vocab <= i < 0 wraps around.
__device__ __forceinline__ int64_t wrap_index(int64_t i, int64_t vocab) { return i < 0 ? i + vocab : i; }

// one warp per gathered row, four rows in flight per warp: the four index loads go out together and so do the row
// loads, so a warp exposes one round trip per four rows instead of one per row (ncu before: 55 % of the copy bandwidth)
__global__ void __launch_bounds__(256) embedding_gather_kernel(const float* __restrict__ w, const int64_t* __restrict__ idx,
                                                               float* __restrict__ out, int64_t n_idx, int64_t vocab,
                                                               int64_t dim, int vec_ok) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int U = 4;
  for (int64_t r0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * U; r0 < n_idx; r0 += warps * U) {
    int64_t v[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u;
      v[u] = r < n_idx ? wrap_index(__ldg(idx + r), vocab) : -1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) ok[u] = (v[u] >= 0 && v[u] < vocab);
    if (vec_ok) {
      const int64_t d4 = dim >> 2;
      for (int64_t c = lane; c < d4; c += 32) {
        float4 t[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
          t[u] = ok[u] ? __ldg(reinterpret_cast<const float4*>(w + v[u] * dim) + c) : make_float4(NAN, NAN, NAN, NAN);
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (r0 + u < n_idx) stg_stream(reinterpret_cast<float4*>(out + (r0 + u) * dim) + c, t[u]);
      }
    } else {
      for (int64_t c = lane; c < dim; c += 32) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (r0 + u < n_idx) out[(r0 + u) * dim + c] = ok[u] ? __ldg(w + v[u] * dim + c) : NAN;
      }
    }
  }
}

// One warp per vocabulary row: scans the index list in ascending position order (ballot over 32
// positions at a time) and accumulates matching rows of dout.  Fixed order → bit-reproducible;
// also performs the `dW[...] = 0` of the reference (rows without hits are written as zeros).
__global__ void __launch_bounds__(256) embedding_scatter_kernel(const float* __restrict__ dout,
                                                                const int64_t* __restrict__ idx,
                                                                float* __restrict__ dw, int64_t n_idx,
                                                                int64_t vocab, int64_t dim) {
  const int lane = threadIdx.x & 31;
  const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= vocab) return;
  // each lane owns columns lane, lane+32, ... ; process the row in column panels of 32*8
  for (int64_t c0 = 0; c0 < dim; c0 += 256) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int64_t p0 = 0; p0 < n_idx; p0 += 32) {
      const int64_t p = p0 + lane;
      const bool hit = (p < n_idx) && (wrap_index(idx[p], vocab) == v);
      unsigned mask = __ballot_sync(0xffffffffu, hit);
      while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const float* src = dout + (p0 + b) * dim + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int64_t c = c0 + lane + 32 * j;
          if (c < dim) acc[j] += __ldg(src + lane + 32 * j);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t c = c0 + lane + 32 * j;
      if (c < dim) dw[v * dim + c] = acc[j];
    }
  }
}

}  // namespace b200ad

using namespace b200ad;

extern "C" {

int b200ad_embedding_gather(const float* w, const int64_t* idx, float* out, int64_t n_idx, int64_t vocab, int64_t dim,
                            void* stream) {
  B200AD_REQUIRE(n_idx >= 0 && vocab > 0 && dim >= 0, "embedding_gather: bad shape");
  if (n_idx == 0 || dim == 0) return B200AD_OK;
  B200AD_REQUIRE(w && idx && out, "embedding_gather: null pointer");
  const int vec_ok = (dim % 4 == 0) && aligned16(w) && aligned16(out);
  embedding_gather_kernel<<<stream_grid(n_idx, 8), 256, 0, (cudaStream_t)stream>>>(w, idx, out, n_idx, vocab, dim, vec_ok);
  count_launch();
  return check_launch("embedding_gather");
}

int b200ad_embedding_scatter_add(const float* dout, const int64_t* idx, float* dw, int64_t n_idx, int64_t vocab,
                                 int64_t dim, void* stream) {
  B200AD_REQUIRE(n_idx >= 0 && vocab > 0 && dim >= 0, "embedding_scatter_add: bad shape");
  if (dim == 0) return B200AD_OK;
  B200AD_REQUIRE(dw && (n_idx == 0 || (dout && idx)), "embedding_scatter_add: null pointer");
  const int64_t blocks = ceil_div(vocab, 8);
  B200AD_REQUIRE(blocks < (int64_t)0x7fffffff, "embedding_scatter_add: vocab too large");
  embedding_scatter_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dout, idx, dw, n_idx, vocab, dim);
  count_launch();
  return check_launch("embedding_scatter_add");
}
}
