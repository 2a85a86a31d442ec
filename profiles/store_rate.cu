// Microbenchmark: how fast can 8 warps per SM (one 256-thread CTA per SM, like a GEMM epilogue) write 128 x 256 fp32
// output tiles of a (131072 x 512) row-major matrix, as a function of the store pattern?
//   V0  row per thread: lane = row, 32 consecutive float4 stores per thread (the tcgen05.ld-natural epilogue)
//   V1  row-contiguous: each warp instruction writes 512 contiguous bytes of one row
//   V2  fully contiguous: each CTA writes one contiguous 128 KB chunk (what an elementwise kernel does)
//   V3  like V1 but with st.global.cs (streaming / evict-first)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_rate store_rate.cu && ./store_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 1) store_kernel(float* c, int variant, int tiles_total) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ld = 512;
  for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
    const int n_tile = tile & 1, m_tile = tile >> 1;
    float* base = c + (int64_t)m_tile * 128 * ld + n_tile * 256;
    const float4 v = make_float4((float)tile, 1.f, 2.f, 3.f);
    if (variant == 0) {
      float* row = base + (int64_t)((warp & 3) * 32 + lane) * ld + (warp >> 2) * 128;
#pragma unroll 8
      for (int i = 0; i < 32; ++i) *reinterpret_cast<float4*>(row + 4 * i) = v;
    } else if (variant == 1 || variant == 3) {
      // warp w writes rows w*16 .. w*16+15; per row two instructions of 512 B
#pragma unroll 4
      for (int r = 0; r < 16; ++r) {
        float* row = base + (int64_t)(warp * 16 + r) * ld;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float* dst = row + h * 128 + lane * 4;
          if (variant == 1) *reinterpret_cast<float4*>(dst) = v;
          else asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
      }
    } else {
      float* chunk = c + (int64_t)tile * 128 * 256;
#pragma unroll 8
      for (int i = 0; i < 32; ++i) *reinterpret_cast<float4*>(chunk + (i * 256 + threadIdx.x) * 4) = v;
    }
  }
}

int main() {
  const int64_t M = 131072, N = 512;
  float* c;
  cudaMalloc(&c, M * N * 4);
  const int tiles = (int)(M / 128 * 2);
  const char* names[4] = {"V0 row per thread (16 B per lane, 2 KB pitch)", "V1 row-contiguous 512 B per instruction",
                          "V2 contiguous 128 KB per CTA", "V3 row-contiguous, st.global.cs"};
  for (int v = 0; v < 4; ++v) {
    for (int rep = 0; rep < 2; ++rep) store_kernel<<<148, 256>>>(c, v, tiles);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int rep = 0; rep < 5; ++rep) store_kernel<<<148, 256>>>(c, v, tiles);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    printf("%-50s %.3f ms  %.0f GB/s  (%.1f B/clk/SM at 1.9 GHz)\n", names[v], ms, M * N * 4 / ms / 1e6,
           M * N * 4 / (ms * 1e-3) / 148 / 1.9e9);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
  return 0;
}
