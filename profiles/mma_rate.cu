// Microbenchmark: cycles per tcgen05.mma (cta_group::1, M = 128) as a function of N, operand source (A from shared
// memory = SS, A from TMEM = TS) and kind (tf32: K = 8 per instruction, f16/bf16: K = 16), issued back to back by
// one thread on fixed operands, one commit at the end.  One CTA per SM on every SM (so the numbers include any
// chip-level power/clock effects), clock64 around the issue loop + completion wait.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../auto_differentiation_b200/csrc/tcgen05.cuh"

using namespace b200ad;

namespace b200ad {
int set_error(int code, const char* fmt, ...) { return code; }
void count_launch(int) {}
}  // namespace b200ad

__device__ __forceinline__ void umma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}

// mode: 0 = tf32 SS, 1 = tf32 TS, 2 = bf16 SS, 3 = bf16 TS
template <int mode, bool BMN>
__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int iters, int nacc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // zero the operands (64 KB): A tile 128 x 32 fp32 (16 KB) at 0, B tile 256 x 32 fp32 (32 KB) at 16 KB
  for (int i = threadIdx.x; i < 49152 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0) {
    const bool bf16 = mode >= 2, ts = mode & 1;
    // idesc: D = f32 (1<<4); A/B format: tf32 = 2, bf16 = 1 at bits 7 / 10; K-major both; N>>3 at 17; M>>4 at 24
    const uint32_t fmt = bf16 ? 1u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((BMN ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      uint64_t bd[4], ad[4];
      uint32_t at[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        bd[k] = BMN ? make_smem_desc(sb + k * 1024, 4096, 512, 1) : make_smem_desc(sb + k * 32, 16, 1024, 2);
        ad[k] = make_smem_desc(sa + k * 32, 16, 1024, 2);
        at[k] = tmem_base + 256 + k * 8;
      }
      const uint32_t d0 = tmem_base, d1 = tmem_base + (nacc > 1 ? (uint32_t)n : 0u);
      t0 = clock64();
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t d = (u & 1) ? d1 : d0;
          if (mode == 0) umma_tf32(d, ad[u & 3], bd[u & 3], idesc, 1u);
          else if (mode == 1) umma_tf32_ts(d, at[u & 3], bd[u & 3], idesc, 1u);
          else if (mode == 2) umma_f16_ss(d, ad[u & 3], bd[u & 3], idesc, 1u);
          else umma_f16_ts(d, at[u & 3], bd[u & 3], idesc, 1u);
        }
      }
      t1 = clock64();
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) {
      out[0] = t1 - t0;   // issue time
      out[1] = t2 - t0;   // until all MMAs retired
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// Attention-like MMA stream (TS, N = 64): groups of 24 MMAs (3 terms x 8 k-steps) alternating between a "score" group
// (A = columns 0..127, D = X_b, K-major B) and an "output" group (A = the X_b / L_b columns, D = OUT, MN-major B), with the
// commits the attention kernel issues (2 per group).  variant: 0 = as described, 1 = no commits inside the loop,
// 2 = all groups use the same A columns / D (no operand switching), 3 = B operand rotates over 6 smem slots of 32 KB,
// 4 / 5 / 6 = variant 0 while four other warps stream tcgen05.ld / tcgen05.st / both on TMEM columns the MMAs do not touch
__global__ void __launch_bounds__(160, 1) attn_stream_kernel(int variant, int groups, long long* out) {
  __shared__ volatile int done_flag;
  if (threadIdx.x == 0) done_flag = 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[4];
  __shared__ uint64_t dummy[3];
  __shared__ uint64_t hs[2];        // variant 17: "the other issuing warp has issued its group"
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 196608 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    for (int i = 0; i < 3; ++i) mbar_init(&dummy[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&hs[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (variant == 17 && warp <= 1) {
    // TWO issuing warps alternate the groups (warp w issues groups g with (g & 1) == w): per group three waits on completed
    // barriers + tcgen05.fence + 24 MMAs + 2 commits as in variant 8, plus a handshake so that group g is issued after
    // group g - 1.  Question: does one warp's wait / fence round trip still drain the pipe when the OTHER warp's MMAs are
    // queued behind it?
    constexpr uint32_t IDESC_K = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t IDESC_MN = IDESC_K | (1u << 16);
    if (warp == 0 && lane == 0) { mbar_arrive(&bars[1]); mbar_arrive(&bars[2]); mbar_arrive(&bars[3]); }
    __syncwarp();
    asm volatile("bar.sync 2, 64;" ::: "memory");
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int g = warp; g < groups; g += 2) {
        const bool outg = (g & 1);
        if (g >= 1) mbar_wait(&hs[g & 1], (warp == 1 ? ((g - 1) >> 1) : ((g >> 1) - 1)) & 1);
        mbar_wait(&bars[1], 0); mbar_wait(&bars[2], 0); mbar_wait(&bars[3], 0);
        tc_fence_after();
        const uint32_t slot = smem_u32(smem) + (uint32_t)(g % 6) * 32768u;
        const uint32_t xb = 128u + (uint32_t)((g >> 1) & 1) * 128u;
        const uint32_t d = tmem_base + (outg ? 384u : xb);
        const uint32_t a_hi = tmem_base + (outg ? xb : 0u), a_lo = a_hi + 64u;
        const uint64_t d0 = outg ? make_smem_desc(slot, 4096, 512, 1) : make_smem_desc(slot, 16, 1024, 2);
        const uint32_t kstep = outg ? 1024u : 32u;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_tf32_ts(d, (t == 0 ? a_lo : a_hi) + kk * 8,
                         d0 + (uint64_t)(((t == 1 ? 16384u : 0u) + (kk >> 2) * 8192u + (kk & 3) * kstep) >> 4),
                         outg ? IDESC_MN : IDESC_K, (t | kk) ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[g & 1])) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[2])) : "memory");
        mbar_arrive(&hs[(g + 1) & 1]);                       // the other warp may issue group g + 1
        if (g == groups - 1) umma_commit(&bars[0]);
      }
      t1 = clock64();
    }
    __syncwarp();
    if (warp == 0) {
      mbar_wait(&bars[0], 0);
      const long long t2 = clock64();
      done_flag = 1;
      if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  } else
  if (warp == 0) {
    constexpr uint32_t IDESC_K = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t IDESC_MN = IDESC_K | (1u << 16);
    long long t0 = 0, t1 = 0;
    if (variant == 7) {
      // the kernel's MMA-role structure: the whole warp loops; per group three waits on (already completed) barriers,
      // a fence, an elected issue block and a warp sync
      if (lane == 0) { mbar_arrive(&bars[1]); mbar_arrive(&bars[2]); mbar_arrive(&bars[3]); }   // complete phase 0
      __syncwarp();
      t0 = clock64();
      for (int g = 0; g < groups; ++g) {
        const bool outg = (g & 1);
        mbar_wait(&bars[1], 0);
        mbar_wait(&bars[2], 0);
        mbar_wait(&bars[3], 0);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t slot = smem_u32(smem) + (uint32_t)(g % 6) * 32768u;
          const uint32_t xb = 128u + (uint32_t)((g >> 1) & 1) * 128u;
          const uint32_t d = tmem_base + (outg ? 384u : xb);
          const uint32_t a_hi = tmem_base + (outg ? xb : 0u), a_lo = a_hi + 64u;
          const uint64_t d0 = outg ? make_smem_desc(slot, 4096, 512, 1) : make_smem_desc(slot, 16, 1024, 2);
          const uint32_t kstep = outg ? 1024u : 32u;
#pragma unroll
          for (int t = 0; t < 3; ++t) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_tf32_ts(d, (t == 0 ? a_lo : a_hi) + kk * 8,
                           d0 + (uint64_t)(((t == 1 ? 16384u : 0u) + (kk >> 2) * 8192u + (kk & 3) * kstep) >> 4),
                           outg ? IDESC_MN : IDESC_K, (t | kk) ? 1u : 0u);
          }
          // commits to barriers nobody waits on (separate from the ones polled above)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[g & 1])) : "memory");
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[2])) : "memory");
        }
        __syncwarp();
      }
      t1 = clock64();
      if (elect_one()) umma_commit(&bars[0]);
    } else if (variant == 8 || variant == 9) {
      // one elected thread runs the whole role, waits included (9: only one wait per group)
      if (lane == 0) { mbar_arrive(&bars[1]); mbar_arrive(&bars[2]); mbar_arrive(&bars[3]); }
      __syncwarp();
      if (elect_one()) {
        t0 = clock64();
        for (int g = 0; g < groups; ++g) {
          const bool outg = (g & 1);
          mbar_wait(&bars[1], 0);
          if (variant == 8) { mbar_wait(&bars[2], 0); mbar_wait(&bars[3], 0); }
          tc_fence_after();
          const uint32_t slot = smem_u32(smem) + (uint32_t)(g % 6) * 32768u;
          const uint32_t xb = 128u + (uint32_t)((g >> 1) & 1) * 128u;
          const uint32_t d = tmem_base + (outg ? 384u : xb);
          const uint32_t a_hi = tmem_base + (outg ? xb : 0u), a_lo = a_hi + 64u;
          const uint64_t d0 = outg ? make_smem_desc(slot, 4096, 512, 1) : make_smem_desc(slot, 16, 1024, 2);
          const uint32_t kstep = outg ? 1024u : 32u;
#pragma unroll
          for (int t = 0; t < 3; ++t) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_tf32_ts(d, (t == 0 ? a_lo : a_hi) + kk * 8,
                           d0 + (uint64_t)(((t == 1 ? 16384u : 0u) + (kk >> 2) * 8192u + (kk & 3) * kstep) >> 4),
                           outg ? IDESC_MN : IDESC_K, (t | kk) ? 1u : 0u);
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[g & 1])) : "memory");
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[2])) : "memory");
        }
        t1 = clock64();
        umma_commit(&bars[0]);
      }
    } else if (variant >= 13 && variant <= 16) {
      // one elected thread; 13: 3 x test_wait + fence, 14: 3 x try_wait, no fence, 15: fence only, 16: 3 x test_wait, no fence
      if (lane == 0) { mbar_arrive(&bars[1]); mbar_arrive(&bars[2]); mbar_arrive(&bars[3]); }
      __syncwarp();
      if (elect_one()) {
        t0 = clock64();
        for (int g = 0; g < groups; ++g) {
          const bool outg = (g & 1);
          if (variant == 13 || variant == 16) {
#pragma unroll
            for (int b = 1; b <= 3; ++b) {
              uint32_t ok;
              do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&bars[b])), "r"(0) : "memory");
              } while (!ok);
            }
          }
          if (variant == 14) { mbar_wait(&bars[1], 0); mbar_wait(&bars[2], 0); mbar_wait(&bars[3], 0); }
          if (variant == 13 || variant == 15) tc_fence_after();
          const uint32_t slot = smem_u32(smem) + (uint32_t)(g % 6) * 32768u;
          const uint32_t xb = 128u + (uint32_t)((g >> 1) & 1) * 128u;
          const uint32_t d = tmem_base + (outg ? 384u : xb);
          const uint32_t a_hi = tmem_base + (outg ? xb : 0u), a_lo = a_hi + 64u;
          const uint64_t d0 = outg ? make_smem_desc(slot, 4096, 512, 1) : make_smem_desc(slot, 16, 1024, 2);
          const uint32_t kstep = outg ? 1024u : 32u;
#pragma unroll
          for (int t = 0; t < 3; ++t) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_tf32_ts(d, (t == 0 ? a_lo : a_hi) + kk * 8,
                           d0 + (uint64_t)(((t == 1 ? 16384u : 0u) + (kk >> 2) * 8192u + (kk & 3) * kstep) >> 4),
                           outg ? IDESC_MN : IDESC_K, (t | kk) ? 1u : 0u);
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[g & 1])) : "memory");
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[2])) : "memory");
        }
        t1 = clock64();
        umma_commit(&bars[0]);
      }
    } else if (variant >= 10 && variant <= 12) {
      // warp-wide loop; 10: waits only (no tcgen05.fence), 11: fence only (no waits),
      // 12: the next group's waits + fence are done BEFORE the last 8 MMAs of the current group are issued
      if (lane == 0) { mbar_arrive(&bars[1]); mbar_arrive(&bars[2]); mbar_arrive(&bars[3]); }
      __syncwarp();
      t0 = clock64();
      for (int g = 0; g < groups; ++g) {
        const bool outg = (g & 1);
        if (variant == 10 || (variant == 12 && g == 0)) { mbar_wait(&bars[1], 0); mbar_wait(&bars[2], 0); mbar_wait(&bars[3], 0); }
        if (variant == 11 || (variant == 12 && g == 0)) tc_fence_after();
        const uint32_t slot = smem_u32(smem) + (uint32_t)(g % 6) * 32768u;
        const uint32_t xb = 128u + (uint32_t)((g >> 1) & 1) * 128u;
        const uint32_t d = tmem_base + (outg ? 384u : xb);
        const uint32_t a_hi = tmem_base + (outg ? xb : 0u), a_lo = a_hi + 64u;
        const uint64_t d0 = outg ? make_smem_desc(slot, 4096, 512, 1) : make_smem_desc(slot, 16, 1024, 2);
        const uint32_t kstep = outg ? 1024u : 32u;
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_tf32_ts(d, (t == 0 ? a_lo : a_hi) + kk * 8,
                           d0 + (uint64_t)(((t == 1 ? 16384u : 0u) + (kk >> 2) * 8192u + (kk & 3) * kstep) >> 4),
                           outg ? IDESC_MN : IDESC_K, (t | kk) ? 1u : 0u);
          }
        }
        __syncwarp();
        if (variant == 12) {      // prefetch the next group's synchronisation while 16 MMAs are queued
          mbar_wait(&bars[1], 0); mbar_wait(&bars[2], 0); mbar_wait(&bars[3], 0);
          tc_fence_after();
        }
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_tf32_ts(d, a_hi + kk * 8, d0 + (uint64_t)(((kk >> 2) * 8192u + (kk & 3) * kstep) >> 4), outg ? IDESC_MN : IDESC_K, 1u);
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[g & 1])) : "memory");
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[2])) : "memory");
        }
        __syncwarp();
      }
      t1 = clock64();
      if (elect_one()) umma_commit(&bars[0]);
    } else
    if (elect_one()) {
      t0 = clock64();
      for (int g = 0; g < groups; ++g) {
        const bool outg = (variant != 2) && (g & 1);
        const uint32_t slot = smem_u32(smem) + (variant == 3 ? (uint32_t)(g % 6) * 32768u : 0u);
        const uint32_t xb = 128u + (uint32_t)((g >> 1) & 1) * 128u;
        const uint32_t d = tmem_base + (outg ? 384u : xb);
        const uint32_t a_hi = tmem_base + (outg ? xb : 0u), a_lo = a_hi + 64u;
        const uint64_t d0 = outg ? make_smem_desc(slot, 4096, 512, 1) : make_smem_desc(slot, 16, 1024, 2);
        const uint32_t kstep = outg ? 1024u : 32u;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_tf32_ts(d, (t == 0 ? a_lo : a_hi) + kk * 8,
                         d0 + (uint64_t)(((t == 1 ? 16384u : 0u) + (kk >> 2) * 8192u + (kk & 3) * kstep) >> 4),
                         outg ? IDESC_MN : IDESC_K, (t | kk) ? 1u : 0u);
        }
        if (variant != 1) {
          umma_commit(&bars[1 + (g & 1)]);
          umma_commit(&bars[3]);
        }
      }
      t1 = clock64();
      umma_commit(&bars[0]);
    }
    __syncwarp();
    mbar_wait(&bars[0], 0);
    const long long t2 = clock64();
    done_flag = 1;
    if (lane == 0 && blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  } else if (variant >= 4) {
    // thread-side TMEM traffic on columns [448, 512) (not used by the MMA stream)
    const uint32_t addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 448u;
    uint32_t x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = i;
    while (!done_flag) {
      if (variant == 4 || variant == 6) { tmem_ld_32x32b_x32(addr, x); tmem_ld_wait(); }
      if (variant == 5 || variant == 6) { tmem_st_32x32b_x32(addr + 32, x); tmem_st_wait(); }
    }
    if (x[3] == 0x12345678u) out[2] = 1;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  cudaFuncSetAttribute(rate_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
  cudaFuncSetAttribute(rate_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
  cudaFuncSetAttribute(rate_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
  cudaFuncSetAttribute(rate_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
  cudaFuncSetAttribute(rate_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
  cudaFuncSetAttribute(rate_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
  const char* names[6] = {"tf32 SS", "tf32 TS", "bf16 SS", "bf16 TS", "tf32 SS Bmn", "tf32 TS Bmn"};
  const int iters = 4096;
  printf("cycles per tcgen05.mma (M=128, cta_group::1), %d back-to-back MMAs, 148 CTAs\n", iters);
  printf("%-12s %5s %5s %12s %12s\n", "mode", "N", "nacc", "issue cyc", "total cyc");
  for (int mode = 0; mode < 6; ++mode)
    for (int n : {16, 32, 64, 128, 256})
      for (int nacc : {1, 2}) {
        if (nacc * n > 256) continue;
        if (mode == 0) rate_kernel<0, false><<<148, 128, 66560>>>(n, iters, nacc, d_out);
        else if (mode == 1) rate_kernel<1, false><<<148, 128, 66560>>>(n, iters, nacc, d_out);
        else if (mode == 2) rate_kernel<2, false><<<148, 128, 66560>>>(n, iters, nacc, d_out);
        else if (mode == 3) rate_kernel<3, false><<<148, 128, 66560>>>(n, iters, nacc, d_out);
        else if (mode == 4) rate_kernel<0, true><<<148, 128, 66560>>>(n, iters, nacc, d_out);
        else rate_kernel<1, true><<<148, 128, 66560>>>(n, iters, nacc, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("%s N=%d: %s\n", names[mode], n, cudaGetErrorString(e));
          return 1;
        }
        long long h[2];
        cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
        printf("%-12s %5d %5d %12.1f %12.1f\n", names[mode], n, nacc, (double)h[0] / iters, (double)h[1] / iters);
      }
  {
    cudaFuncSetAttribute(attn_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 197632);
    const int groups = 512;
    const char* vn[18] = {"attention-like stream", "... without commits", "... no operand / accumulator switching", "... B rotating over 6 slots",
                         "... + 4 warps streaming tcgen05.ld", "... + 4 warps streaming tcgen05.st", "... + 4 warps streaming ld + st", "kernel-like role: warp loop, 3 waits + elect + syncwarp per group",
                         "one elected thread runs the role (3 waits per group)", "one elected thread, 1 wait per group",
                         "warp loop: 3 waits, no tcgen05.fence", "warp loop: fence only", "warp loop: next group's waits + fence before the last 8 MMAs",
                         "one thread: 3 x test_wait + fence", "one thread: 3 x try_wait, no fence", "one thread: fence only", "one thread: 3 x test_wait, no fence",
                         "TWO issuing warps alternate groups (3 waits + fence + handshake per group)"};
    for (int v = 0; v < 18; ++v) {
      attn_stream_kernel<<<148, 160, 197632>>>(v, groups, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("attn stream: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2];
      cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
      printf("%-44s issue %.1f  total %.1f cycles per MMA (24 MMAs per group)\n", vn[v], (double)h[0] / (groups * 24), (double)h[1] / (groups * 24));
    }
  }
  return 0;
}
