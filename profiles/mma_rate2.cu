// Microbenchmark: cycles per tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, A from TMEM, kind::f16) as a function of N,
// against the cta_group::1 rate at the same N — the attention kernels issue M = 256, N = 64, K = 16 instructions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate2 mma_rate2.cu && ./mma_rate2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../auto_differentiation_b200/csrc/tcgen05.cuh"

using namespace b200ad;
namespace b200ad {
int set_error(int code, const char* fmt, ...) { return code; }
void count_launch(int) {}
}  // namespace b200ad

__device__ __forceinline__ void mma1(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma2(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}

template <bool PAIR>
__global__ void __launch_bounds__(128, 1) rate2_kernel(int n, int iters, int nd, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < 65536 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) { if (PAIR) tmem_alloc2(&tmem_slot, 512); else tmem_alloc(&tmem_slot, 512); }
  fence_proxy_async();
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
    const uint32_t sb = smem_u32(smem);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      uint64_t bd[4];
      uint32_t at[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        bd[k] = make_smem_desc(sb + k * 32, 16, 1024, 2);
        at[k] = tmem_base + 448 + k * 8;
      }
      t0 = clock64();
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t d = tmem_base + (uint32_t)((u % nd) * n);
          if (PAIR) mma2(d, at[u & 3], bd[u & 3], idesc); else mma1(d, at[u & 3], bd[u & 3], idesc);
        }
      }
      t1 = clock64();
      if (PAIR) umma2_commit_mc(&bar); else umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  if (warp == 0) { if (PAIR) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  const int iters = 4096;
  cudaFuncSetAttribute(rate2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
  cudaFuncSetAttribute(rate2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
  printf("cycles per kind::f16 tcgen05.mma, A from TMEM, %d back-to-back, 148 CTAs\n", iters);
  printf("%-14s %5s %5s %12s %12s\n", "group", "N", "ndest", "issue cyc", "total cyc");
  for (int pair = 0; pair < 2; ++pair)
    for (int n : {64, 128, 256})
      for (int nd : {1, 2}) {
        if (nd * n > 256) continue;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = 66560;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = pair ? 2 : 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = pair ? cudaLaunchKernelEx(&cfg, rate2_kernel<true>, n, iters, nd, d_out)
                             : cudaLaunchKernelEx(&cfg, rate2_kernel<false>, n, iters, nd, d_out);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("pair %d N=%d: %s\n", pair, n, cudaGetErrorString(e)); return 1; }
        long long h[2];
        cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
        printf("%-14s %5d %5d %12.1f %12.1f\n", pair ? "cta_group::2" : "cta_group::1", n, nd, (double)h[0] / iters, (double)h[1] / iters);
      }
  return 0;
}
