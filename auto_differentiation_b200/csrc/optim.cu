// Multi-tensor optimizer kernels: ONE launch walks a device-side table of (param, grad, m, v, numel)
// entries in chunks of B200AD_OPT_CHUNK elements (one CTA per chunk, float4 access).
// Reference bodies replaced: Adam.update utils/optimizer.py:14-29 (== layers/layers.py:59-73),
// SGD.update layers.py:829-831, clip_grads :132-141.  HBM-bound: Adam 28 B/param, SGD 12 B/param.
#include "common.cuh"

namespace b200ad {

constexpr int kChunk = B200AD_OPT_CHUNK;  // 4096 elements = 256 threads x 4 float4

struct ChunkRef {
  const b200ad_tensor_entry* e;
  int64_t off;
  int count;
};
__device__ __forceinline__ ChunkRef chunk_ref(const b200ad_tensor_entry* table, const int32_t* ct,
                                              const int32_t* ci) {
  ChunkRef r;
  r.e = table + ct[blockIdx.x];
  r.off = (int64_t)ci[blockIdx.x] * kChunk;
  const int64_t left = r.e->numel - r.off;
  r.count = left < kChunk ? (int)left : kChunk;
  return r;
}
__device__ __forceinline__ bool all_aligned16(const void* a, const void* b, const void* c, const void* d) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d)) & 15u) == 0;
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float lr_t, float b1c, float b2c,
                                         float eps, float gs) {
  g *= gs;
  m += b1c * (g - m);          // self.m[i] += (1 - beta1) * (grads[i] - self.m[i])      optimizer.py:26
  v += b2c * (g * g - v);      // self.v[i] += (1 - beta2) * (grads[i]**2 - self.v[i])   :27
  p -= lr_t * m / (sqrtf(v) + eps);  // params[i] -= lr_t * m / (sqrt(v) + 1e-7)         :29
}

// iter += 1; lr_t = lr*sqrt(1-b2^iter)/(1-b1^iter) in fp64 (the host expression of optimizer.py:22-23), so a
// captured CUDA graph of the update stays correct across replays without any host involvement.
__global__ void adam_step_scalar_kernel(int64_t* __restrict__ iter, double lr, double b1, double b2,
                                        float* __restrict__ lr_t) {
  const int64_t t = *iter + 1;
  *iter = t;
  *lr_t = (float)(lr * sqrt(1.0 - pow(b2, (double)t)) / (1.0 - pow(b1, (double)t)));
}

__global__ void __launch_bounds__(256) adam_multi_kernel(const b200ad_tensor_entry* __restrict__ table,
                                                         const int32_t* __restrict__ ct, const int32_t* __restrict__ ci,
                                                         float lr_t, float b1c, float b2c, float eps, float gs,
                                                         const float* __restrict__ lr_t_dev) {
  if (lr_t_dev) lr_t = *lr_t_dev;
  const ChunkRef r = chunk_ref(table, ct, ci);
  float* p = r.e->param + r.off;
  const float* g = r.e->grad + r.off;
  float* m = r.e->m + r.off;
  float* v = r.e->v + r.off;
  if (all_aligned16(p, g, m, v)) {
    const int n4 = r.count >> 2;
    for (int i = threadIdx.x; i < n4; i += 256) {
      float4 pp = reinterpret_cast<float4*>(p)[i];
      const float4 gg = ldg_stream(reinterpret_cast<const float4*>(g) + i);
      float4 mm = reinterpret_cast<float4*>(m)[i];
      float4 vv = reinterpret_cast<float4*>(v)[i];
      adam_one(pp.x, gg.x, mm.x, vv.x, lr_t, b1c, b2c, eps, gs);
      adam_one(pp.y, gg.y, mm.y, vv.y, lr_t, b1c, b2c, eps, gs);
      adam_one(pp.z, gg.z, mm.z, vv.z, lr_t, b1c, b2c, eps, gs);
      adam_one(pp.w, gg.w, mm.w, vv.w, lr_t, b1c, b2c, eps, gs);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < r.count; i += 256) adam_one(p[i], g[i], m[i], v[i], lr_t, b1c, b2c, eps, gs);
  } else {
    for (int i = threadIdx.x; i < r.count; i += 256) adam_one(p[i], g[i], m[i], v[i], lr_t, b1c, b2c, eps, gs);
  }
}

__global__ void __launch_bounds__(256) sgd_multi_kernel(const b200ad_tensor_entry* __restrict__ table,
                                                        const int32_t* __restrict__ ct, const int32_t* __restrict__ ci,
                                                        float lr, float gs) {
  const ChunkRef r = chunk_ref(table, ct, ci);
  float* p = r.e->param + r.off;
  const float* g = r.e->grad + r.off;
  const float k = lr * gs;
  if (all_aligned16(p, g, p, g)) {
    const int n4 = r.count >> 2;
    for (int i = threadIdx.x; i < n4; i += 256) {
      float4 pp = reinterpret_cast<float4*>(p)[i];
      const float4 gg = ldg_stream(reinterpret_cast<const float4*>(g) + i);
      pp.x -= k * gg.x; pp.y -= k * gg.y; pp.z -= k * gg.z; pp.w -= k * gg.w;  // params[i] -= lr*grads[i]  layers.py:831
      reinterpret_cast<float4*>(p)[i] = pp;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < r.count; i += 256) p[i] -= k * g[i];
  } else {
    for (int i = threadIdx.x; i < r.count; i += 256) p[i] -= k * g[i];
  }
}

__global__ void __launch_bounds__(256) sumsq_chunk_kernel(const b200ad_tensor_entry* __restrict__ table,
                                                          const int32_t* __restrict__ ct, const int32_t* __restrict__ ci,
                                                          float* __restrict__ ws) {
  __shared__ float red[32];
  const ChunkRef r = chunk_ref(table, ct, ci);
  const float* g = r.e->grad + r.off;
  float s = 0.f;
  for (int i = threadIdx.x; i < r.count; i += 256) {
    const float x = __ldg(g + i);
    s += x * x;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) ws[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024) sum_final_kernel(const float* __restrict__ ws, int64_t n, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += ws[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s;
}

__global__ void __launch_bounds__(256) clip_scale_kernel(const b200ad_tensor_entry* __restrict__ table,
                                                         const int32_t* __restrict__ ct, const int32_t* __restrict__ ci,
                                                         const float* __restrict__ norm_sq, float max_norm) {
  // total_norm = sqrt(sum g^2); rate = max_norm/(total_norm + 1e-6); if rate < 1: grad *= rate   layers.py:132-141
  const float rate = max_norm / (sqrtf(norm_sq[0]) + 1e-6f);
  if (!(rate < 1.f)) return;
  const ChunkRef r = chunk_ref(table, ct, ci);
  float* g = r.e->grad + r.off;
  for (int i = threadIdx.x; i < r.count; i += 256) g[i] *= rate;
}

}  // namespace b200ad

using namespace b200ad;

static int check_table(const char* who, const void* table, int n_tensors, int64_t chunks, const void* ct, const void* ci) {
  B200AD_REQUIRE(n_tensors >= 0 && chunks >= 0, "%s: negative count", who);
  if (chunks == 0) return 1;
  B200AD_REQUIRE(table && ct && ci, "%s: null table", who);
  B200AD_REQUIRE(chunks < (int64_t)0x7fffffff, "%s: too many chunks", who);
  return B200AD_OK;
}

extern "C" {

int b200ad_adam_step_scalar(int64_t* iter_dev, double lr, double beta1, double beta2, float* lr_t_dev, void* stream) {
  B200AD_REQUIRE(iter_dev && lr_t_dev, "adam_step_scalar: null pointer");
  adam_step_scalar_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(iter_dev, lr, beta1, beta2, lr_t_dev);
  count_launch();
  return check_launch("adam_step_scalar");
}

int b200ad_adam_multi(const b200ad_tensor_entry* table_dev, int n_tensors, int64_t total_chunks,
                      const int32_t* chunk_tensor_dev, const int32_t* chunk_index_dev, float lr_t, float beta1,
                      float beta2, float eps, float grad_scale, const float* lr_t_dev, void* stream) {
  int rc = check_table("adam_multi", table_dev, n_tensors, total_chunks, chunk_tensor_dev, chunk_index_dev);
  if (rc) return rc < 0 ? rc : B200AD_OK;
  adam_multi_kernel<<<(unsigned)total_chunks, 256, 0, (cudaStream_t)stream>>>(table_dev, chunk_tensor_dev, chunk_index_dev,
                                                                              lr_t, 1.f - beta1, 1.f - beta2, eps, grad_scale, lr_t_dev);
  count_launch();
  return check_launch("adam_multi");
}

int b200ad_sgd_multi(const b200ad_tensor_entry* table_dev, int n_tensors, int64_t total_chunks,
                     const int32_t* chunk_tensor_dev, const int32_t* chunk_index_dev, float lr, float grad_scale,
                     void* stream) {
  int rc = check_table("sgd_multi", table_dev, n_tensors, total_chunks, chunk_tensor_dev, chunk_index_dev);
  if (rc) return rc < 0 ? rc : B200AD_OK;
  sgd_multi_kernel<<<(unsigned)total_chunks, 256, 0, (cudaStream_t)stream>>>(table_dev, chunk_tensor_dev, chunk_index_dev, lr,
                                                                             grad_scale);
  count_launch();
  return check_launch("sgd_multi");
}

int b200ad_sumsq_multi(const b200ad_tensor_entry* table_dev, int n_tensors, int64_t total_chunks,
                       const int32_t* chunk_tensor_dev, const int32_t* chunk_index_dev, float* ws, float* norm_out,
                       void* stream) {
  B200AD_REQUIRE(norm_out, "sumsq_multi: null output");
  int rc = check_table("sumsq_multi", table_dev, n_tensors, total_chunks, chunk_tensor_dev, chunk_index_dev);
  if (rc < 0) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (rc == 1) {
    sum_final_kernel<<<1, 1024, 0, s>>>(nullptr, 0, norm_out);
    count_launch();
    return check_launch("sumsq_multi");
  }
  B200AD_REQUIRE(ws, "sumsq_multi: null workspace");
  sumsq_chunk_kernel<<<(unsigned)total_chunks, 256, 0, s>>>(table_dev, chunk_tensor_dev, chunk_index_dev, ws);
  sum_final_kernel<<<1, 1024, 0, s>>>(ws, total_chunks, norm_out);
  count_launch(2);
  return check_launch("sumsq_multi");
}

int b200ad_clip_scale_multi(const b200ad_tensor_entry* table_dev, int n_tensors, int64_t total_chunks,
                            const int32_t* chunk_tensor_dev, const int32_t* chunk_index_dev, const float* norm_sq,
                            float max_norm, void* stream) {
  B200AD_REQUIRE(norm_sq, "clip_scale_multi: null norm");
  int rc = check_table("clip_scale_multi", table_dev, n_tensors, total_chunks, chunk_tensor_dev, chunk_index_dev);
  if (rc) return rc < 0 ? rc : B200AD_OK;
  clip_scale_kernel<<<(unsigned)total_chunks, 256, 0, (cudaStream_t)stream>>>(table_dev, chunk_tensor_dev, chunk_index_dev,
                                                                              norm_sq, max_norm);
  count_launch();
  return check_launch("clip_scale_multi");
}
}
