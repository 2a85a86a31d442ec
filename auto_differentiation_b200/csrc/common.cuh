// Shared device/host helpers for libb200ad.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/b200ad.h"

#define B200AD_NUM_SMS 148   /* B200; only the fallback when the device query fails: launch sizing uses sm_count() */

namespace b200ad {

// Programmatic dependent launch: let the next kernel of the stream start its prologue / wait until the previous kernel of
// the stream has completed and its writes are visible.  Both are no-ops in a kernel launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }


int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
// multiprocessor count of the current device (queried once per process; 148 on B200)
int sm_count();

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B200AD_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return B200AD_OK;
}

#define B200AD_REQUIRE(cond, ...)                                            \
  do {                                                                       \
    if (!(cond)) return ::b200ad::set_error(B200AD_ERR_INVALID, __VA_ARGS__); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Grid for a grid-stride streaming kernel: enough CTAs for the work, capped at a whole number of
// waves (148 SMs x 8 resident 256-thread CTAs = full occupancy).
inline int stream_grid(int64_t work_items, int per_block, int ctas_per_sm = 8) {
  int64_t need = ceil_div(work_items, per_block);
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---------------------------------------------------------------------------------------------
// 128-bit streaming loads/stores. Streaming data is read once: bypass L1 allocation.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// Coherent variant for kernels that may run IN PLACE (output aliasing an input): ld.global.nc on memory the same
// kernel writes is undefined, so the aliasing-tolerant backward kernels use this one.
__device__ __forceinline__ float4 ld_stream_coherent(const float4* p) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum (all threads get the result). blockDim.x must be a multiple of 32, <= 1024.
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect smem32 reuse across consecutive calls
  if (lane == 0) smem32[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? smem32[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* smem32) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) smem32[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? smem32[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

}  // namespace b200ad
