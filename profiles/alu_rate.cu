// Microbenchmark: issue cost of the instructions the attention row owners / converters are made of, per SM sub-partition:
// ex2.approx (MUFU), cvt.rn.f16x2.f32 (F2FP pack), FFMA, LOP3, and the split2 sequence (2 LOP + 2 FADD + 2 F2FP per pair).
// One CTA, `warps` warps (4 = one per sub-partition, 8 = two, ...); cycles per warp-instruction per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_rate alu_rate.cu && ./alu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int OP>
__global__ void rate(int iters, float seed, long long* out, float* sink) {
  float a[8];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f; u[i] = threadIdx.x + i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      else if (OP == 1) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "+r"(u[i]) : "f"(__uint_as_float(u[i])), "f"(a[i]));
      else if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(seed), "f"(1.0f));
      else if (OP == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(0x12345u));
      else if (OP == 4) {   // split2 of a pair (inputs depend on the previous trip through u[i])
        const float v0 = __uint_as_float(u[i]), v1 = a[i];
        float h0, h1, l0, l1;
        asm volatile("and.b32 %0, %1, 0xffffe000;" : "=f"(h0) : "f"(v0));
        asm volatile("and.b32 %0, %1, 0xffffe000;" : "=f"(h1) : "f"(v1));
        asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(l0) : "f"(v0), "f"(h0));
        asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(l1) : "f"(v1), "f"(h1));
        uint32_t hh;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hh) : "f"(h1), "f"(h0));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(l1), "f"(__uint_as_float(hh)));
      } else if (OP == 5) {  // one owner element: FFMA + EX2 + FFMA + FMUL
        float pv;
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(pv) : "f"(a[i]), "f"(seed), "f"(-1.0f));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(pv));
        a[i] = pv * fmaf(a[(i + 1) & 7], seed, -0.5f);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
  sink[threadIdx.x] = s;
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  long long* d_out;
  float* d_sink;
  cudaMalloc(&d_out, 8);
  cudaMalloc(&d_sink, 4096 * 4);
  const int iters = 4096;
  const char* names[6] = {"ex2.approx (MUFU)", "cvt.rn.f16x2.f32 (F2FP)", "fma.rn.f32", "lop3", "split2 of a pair (6 instr)", "owner element (4 instr)"};
  printf("%-30s %6s %22s\n", "instruction", "warps", "cycles / warp-op / SMSP");
  for (int op = 0; op < 6; ++op)
    for (int warps : {4, 8, 16}) {
      switch (op) {
        case 0: rate<0><<<1, warps * 32>>>(iters, 0.5f, d_out, d_sink); break;
        case 1: rate<1><<<1, warps * 32>>>(iters, 0.5f, d_out, d_sink); break;
        case 2: rate<2><<<1, warps * 32>>>(iters, 0.5f, d_out, d_sink); break;
        case 3: rate<3><<<1, warps * 32>>>(iters, 0.5f, d_out, d_sink); break;
        case 4: rate<4><<<1, warps * 32>>>(iters, 0.5f, d_out, d_sink); break;
        default: rate<5><<<1, warps * 32>>>(iters, 0.5f, d_out, d_sink); break;
      }
      cudaDeviceSynchronize();
      long long h;
      cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
      // per sub-partition: warps / 4 warps each issuing iters * 8 ops
      printf("%-30s %6d %22.2f\n", names[op], warps, (double)h / ((double)iters * 8 * (warps / 4)));
    }
  return 0;
}
