// Persistent recurrent kernel: the WHOLE time loop of the tanh RNN in one launch (forward) / one launch (BPTT).
//
// Reference loops replaced: TimeRNN.forward layers/layers.py:203-207 (h_t = tanh(x_t Wx + h_{t-1} Wh + b), the x_t Wx + b
// part hoisted into one big GEMM by the caller) and TimeRNN.backward :210-225 in its intended form:
//   forward  phase t      : hs[t+1] = tanh(hs[t+1] + hs[t] @ Wh)                                   t = 0 .. T-1
//   backward phase p      : dA[t]   = (dhs[:, t, :] + dA[t+1] @ Wh^T) * (1 - hs[t+1]^2)            t = T-1-p, p = 0 .. T-1
//            phase T      : dh0     = dA[0] @ Wh^T
// Every phase is one dependent (N <= 64) x H x H product.  Launched step by step it costs ~10 us per step at H = 1024 (512
// dependent launches per training step, VERDICT r1 "missing 1"); here the recurrent weight never leaves the chip:
//
//   * decomposition: the H output columns are cut into 128-wide slices (TMEM lanes), the reduction dimension into 8
//     K-slices = the 8 CTAs of a thread-block cluster, the batch rows optionally into 2 halves -> (H/128) x 8 x n_split
//     CTAs (64 or 128 at H = 1024), all co-resident (checked with cudaOccupancyMaxActiveClusters before launch);
//   * each CTA keeps ITS 128 x (H/8) block of Wh (fwd) / Wh^T (bwd), split once into tf32 hi / lo, RESIDENT IN TENSOR
//     MEMORY for the whole launch as the TS-mode A operand (lane = output column, TMEM column = k);
//   * per phase it loads its (rows x H/8) slice of the previous state from L2 (ld.global.cg), splits it into hi / lo
//     SWIZZLE_128B K-major tiles in shared memory (the B operand) and issues 3 x (H/8)/8 tcgen05.mma kind::tf32
//     (A_lo.B_hi + A_hi.B_lo + A_hi.B_hi: the same fp32-faithful 3xTF32 scheme as gemm_tc.cu, one <= 128-deep chain);
//   * the 8 K-partials of a cluster are summed through distributed shared memory in a fixed order (bit-reproducible):
//     CTA j of the cluster finalises rows [j*rows/8, (j+1)*rows/8) of the 128-column slice, applies the fused epilogue
//     (+ pre-activation, tanh / BPTT product) and writes the new state to global memory;
//   * dependent phases are ordered by per-slice arrival counters in global memory (release: __syncthreads +
//     __threadfence + red.release.gpu by one thread; acquire: one thread spins on ld.acquire.gpu, then __syncthreads):
//     a consumer only waits for the ONE cluster that produces its K-slice, never for the whole grid.
// All waits are bounded (trap after ~2 s) so that a protocol bug cannot hang the GPU.
#include "common.cuh"
#include "tcgen05.cuh"

namespace b200ad {

constexpr int RS_THREADS = 256;
constexpr int RS_CL = 8;            // CTAs per cluster = K-slices
constexpr int RS_BM = 128;          // output columns per cluster (= TMEM lanes)
constexpr uint32_t RS_ACC_COLS = 64;   // TMEM columns reserved for the accumulator (<= 64 batch rows)

struct RnnSeqParams {
  const float* w;        // Wh, (H, H) row-major
  int64_t ldw;
  float* hs;             // (T+1, N, H): fwd reads hs[t] and completes hs[t+1] in place; bwd reads hs[t+1]
  const float* dhs;      // bwd: upstream gradient of the outputs, element (n, t, c) at dhs[n*dhs_sn + t*dhs_st + c]
  int64_t dhs_sn, dhs_st;
  float* dA;             // bwd: (T, N, H) pre-activation gradients (output, also the next phase's operand)
  float* dh0;            // bwd: (N, H) gradient of the initial state (may be null)
  unsigned int* flags;   // [H/128 * n_split] arrival counters, zero at launch
  int32_t T, N, H;
  int32_t ks;            // K elements per CTA = H / 8  (32, 64 or 128)
  int32_t n_split;       // batch halves (1 or 2)
  int32_t bwd;
};

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// (the memory clobber keeps the load behind the acquire / barrier that precedes it; callers issue all loads of a batch
// before the first use so that they are in flight together)
__device__ __forceinline__ float4 ld_cg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_addr, uint32_t cta) {
  uint32_t ra;
  float4 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(cta));
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t tf32_hi(uint32_t x) { return x & 0xffffe000u; }
__device__ __forceinline__ uint32_t tf32_lo(uint32_t x) {
  return __float_as_uint(__uint_as_float(x) - __uint_as_float(x & 0xffffe000u)) & 0xffffe000u;
}

// NB: batch rows per cluster (32 or 64) = the MMA N dimension.
template <int NB>
__global__ void __launch_bounds__(RS_THREADS, 1) rnn_seq_kernel(const RnnSeqParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int KS = p.ks;
  const uint32_t KB_BYTES = NB * 128;                    // one 32-deep k-block of the B tile: NB rows x 128 B
  const uint32_t B_BYTES = (uint32_t)(KS / 32) * KB_BYTES;
  uint8_t* b_hi = smem;                                  // [KS/32][NB][128 B] SWIZZLE_128B, K-major
  uint8_t* b_lo = smem + B_BYTES;
  float* part = reinterpret_cast<float*>(smem + 2 * B_BYTES);          // [2][NB][128] this CTA's K-partials (double buffered)
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(smem + 2 * B_BYTES + 2 * NB * RS_BM * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();               // K-slice of this CTA
  const int cluster_id = blockIdx.x / RS_CL;
  const int cs = cluster_id / p.n_split, ns = cluster_id % p.n_split;   // output-column slice, batch half
  const int c0 = cs * RS_BM, n0 = ns * NB, k0 = (int)rank * KS;
  const uint32_t tmem_cols = (RS_ACC_COLS + 2u * KS) <= 128u ? 128u : ((RS_ACC_COLS + 2u * KS) <= 256u ? 256u : 512u);

  if (tid == 0) {
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t a_hi = tmem_base + RS_ACC_COLS, a_lo = a_hi + (uint32_t)KS;

  // ---- one-time: this CTA's 128 x KS block of Wop^T -> TMEM (hi: warps 0-3, lo: warps 4-7; a warp owns lane quarter warp & 3)
  {
    const int q = warp & 3, half = warp >> 2;
    const int c = c0 + q * 32 + lane;                    // output column = TMEM lane
    for (int j = 0; j < KS / 32; ++j) {
      uint32_t v[32];
      if (!p.bwd) {          // Wop = Wh: A[c][r] = Wh[k0 + r][c]  (coalesced across lanes for each r)
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__ldg(p.w + (int64_t)(k0 + j * 32 + i) * p.ldw + c));
      } else {               // Wop = Wh^T: A[c][r] = Wh[c][k0 + r]  (each lane reads 128 contiguous bytes)
        const float4* row = reinterpret_cast<const float4*>(p.w + (int64_t)c * p.ldw + k0 + j * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t4 = __ldg(row + i);
          v[4 * i] = __float_as_uint(t4.x); v[4 * i + 1] = __float_as_uint(t4.y);
          v[4 * i + 2] = __float_as_uint(t4.z); v[4 * i + 3] = __float_as_uint(t4.w);
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = half ? tf32_lo(v[i]) : tf32_hi(v[i]);
      tmem_st_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + RS_ACC_COLS + (half ? (uint32_t)KS : 0u) + (uint32_t)j * 32u, v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  // instruction descriptor: D = f32, A = B = tf32, both K-major, N = NB, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(RS_BM >> 4) << 24);
  const uint32_t sb_hi = smem_u32(b_hi), sb_lo = smem_u32(b_lo);
  const uint32_t part_addr = smem_u32(part);

  // the rows this thread finalises after the cluster reduction: NB/8 rows x 32 float4 per CTA
  const bool fin = tid < NB * 4;
  const int fin_row = (int)rank * (NB / 8) + (tid >> 5);        // row inside the cluster's NB rows
  const int fin_n = n0 + fin_row;                                // batch row
  const int fin_c = c0 + (tid & 31) * 4;                         // first of 4 output columns
  const bool fin_live = fin && fin_n < p.N;
  const int64_t NH = (int64_t)p.N * p.H;

  // which cluster produces this CTA's K-slice of the state, and its counter
  const unsigned int* wait_flag = p.flags + ((k0 / RS_BM) * p.n_split + ns);
  unsigned int* my_flag = p.flags + (cs * p.n_split + ns);

  const int phases = p.bwd ? p.T + 1 : p.T;
  uint32_t mma_phase = 0;
  for (int ph = 0; ph < phases; ++ph) {
    // ---- operand / epilogue pointers of this phase
    const float* src;            // (N, H) state consumed by the product (nullptr: no product, acc = 0)
    const int t = p.bwd ? p.T - 1 - ph : ph;
    if (!p.bwd) src = p.hs + (int64_t)ph * NH;
    else src = ph == 0 ? nullptr : p.dA + (int64_t)(p.T - ph) * NH;

    // ---- epilogue operands do not depend on the recurrence: fetch them before waiting
    float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f), e1 = e0;
    if (fin_live) {
      if (!p.bwd) {
        e0 = ld_cg_f4(p.hs + (int64_t)(ph + 1) * NH + (int64_t)fin_n * p.H + fin_c);          // x_t Wx + b (pre-activation)
      } else if (ph < p.T) {
        e0 = ld_cg_f4(p.dhs + (int64_t)fin_n * p.dhs_sn + (int64_t)t * p.dhs_st + fin_c);     // upstream dh_t
        e1 = ld_cg_f4(p.hs + (int64_t)(t + 1) * NH + (int64_t)fin_n * p.H + fin_c);           // h_t
      }
    }

    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src != nullptr) {
      // ---- wait until the producing cluster has published the previous phase (8 CTAs x ph arrivals)
      if (ph > 0) {
        if (tid == 0) {
          const unsigned int need = (unsigned int)RS_CL * (unsigned int)ph;
          if (ld_acquire_gpu_u32(wait_flag) < need) {
            const long long t0 = clock64();
            uint32_t spins = 0;
            while (ld_acquire_gpu_u32(wait_flag) < need) {
              if ((++spins & 255u) == 0 && clock64() - t0 > 4000000000ll) __trap();
            }
          }
        }
        __syncthreads();
      }
      // ---- B operand: rows [n0, n0+NB) x k [k0, k0+KS) of the state -> hi / lo SWIZZLE_128B tiles
      {
        // all loads of the slice are issued before the first one is consumed: ONE L2 round trip per phase
        const int vec_per_row = KS / 4, total = NB * vec_per_row;
        constexpr int MAX_IT = NB / 8;                       // NB * (128 / 4) / 256 at the largest K-slice
        float4 xs[MAX_IT];
#pragma unroll
        for (int i = 0; i < MAX_IT; ++i) {
          const int idx = tid + i * RS_THREADS;
          const int row = idx / vec_per_row, c16 = idx - row * vec_per_row;
          xs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (idx < total && n0 + row < p.N) xs[i] = ld_cg_f4(src + (int64_t)(n0 + row) * p.H + k0 + c16 * 4);
        }
#pragma unroll
        for (int i = 0; i < MAX_IT; ++i) {
          const int idx = tid + i * RS_THREADS;
          if (idx >= total) break;
          const int row = idx / vec_per_row, c16 = idx - row * vec_per_row;
          const float4 x = xs[i];
          const uint32_t off = (uint32_t)(c16 >> 3) * KB_BYTES + (uint32_t)row * 128u + (uint32_t)(((c16 & 7) ^ (row & 7)) << 4);
          uint4 h, l;
          h.x = tf32_hi(__float_as_uint(x.x)); h.y = tf32_hi(__float_as_uint(x.y));
          h.z = tf32_hi(__float_as_uint(x.z)); h.w = tf32_hi(__float_as_uint(x.w));
          l.x = tf32_lo(__float_as_uint(x.x)); l.y = tf32_lo(__float_as_uint(x.y));
          l.z = tf32_lo(__float_as_uint(x.z)); l.w = tf32_lo(__float_as_uint(x.w));
          *reinterpret_cast<uint4*>(b_hi + off) = h;
          *reinterpret_cast<uint4*>(b_lo + off) = l;
        }
        fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
      }
      __syncthreads();
      // ---- 3 x KS/8 MMAs (one elected thread), accumulator = TMEM columns [0, NB)
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
          for (int kb = 0; kb < KS / 32; ++kb) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t koff = (uint32_t)kb * KB_BYTES + (uint32_t)kk * 32u;
              const uint32_t acol = (uint32_t)kb * 32u + (uint32_t)kk * 8u;
              const uint64_t dh = make_smem_desc(sb_hi + koff, 16u, 1024u, 2u);
              const uint64_t dl = make_smem_desc(sb_lo + koff, 16u, 1024u, 2u);
              umma_tf32_ts(tmem_base, a_lo + acol, dh, idesc, (kb == 0 && kk == 0) ? 0u : 1u);
              umma_tf32_ts(tmem_base, a_hi + acol, dl, idesc, 1u);
              umma_tf32_ts(tmem_base, a_hi + acol, dh, idesc, 1u);
            }
          }
          umma_commit(mma_bar);
        }
        __syncwarp();
      }
      mbar_wait(mma_bar, mma_phase & 1);
      ++mma_phase;
      tc_fence_after();
      // ---- accumulator (lane = column c, TMEM column = row n) -> part[ph & 1][n][c]
      {
        const int q = warp & 3, half = warp >> 2;
        float* dst = part + (size_t)(ph & 1) * NB * RS_BM + (q * 32 + lane);
        if constexpr (NB == 32) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)half * 16u, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) dst[(size_t)(half * 16 + j) * RS_BM] = __uint_as_float(r[j]);
        } else {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)half * 32u, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[(size_t)(half * 32 + j) * RS_BM] = __uint_as_float(r[j]);
        }
        tc_fence_before();
      }
      // ---- all 8 K-partials of the cluster are in shared memory after this barrier
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
      if (fin) {
        const uint32_t a = part_addr + (uint32_t)((ph & 1) * NB * RS_BM + fin_row * RS_BM + (tid & 31) * 4) * 4u;
        float4 pv[RS_CL];
#pragma unroll
        for (uint32_t peer = 0; peer < (uint32_t)RS_CL; ++peer) pv[peer] = ld_dsmem_f4(a, peer);   // 8 loads in flight
#pragma unroll
        for (int peer = 0; peer < RS_CL; ++peer) {                     // fixed order: bit-reproducible
          acc.x += pv[peer].x; acc.y += pv[peer].y; acc.z += pv[peer].z; acc.w += pv[peer].w;
        }
      }
    }

    // ---- fused epilogue + store of the new state
    if (fin_live) {
      if (!p.bwd) {
        float4 o;
        o.x = tanhf(e0.x + acc.x); o.y = tanhf(e0.y + acc.y); o.z = tanhf(e0.z + acc.z); o.w = tanhf(e0.w + acc.w);
        *reinterpret_cast<float4*>(p.hs + (int64_t)(ph + 1) * NH + (int64_t)fin_n * p.H + fin_c) = o;
      } else if (ph < p.T) {
        float4 o;
        o.x = (e0.x + acc.x) * (1.f - e1.x * e1.x); o.y = (e0.y + acc.y) * (1.f - e1.y * e1.y);
        o.z = (e0.z + acc.z) * (1.f - e1.z * e1.z); o.w = (e0.w + acc.w) * (1.f - e1.w * e1.w);
        *reinterpret_cast<float4*>(p.dA + (int64_t)t * NH + (int64_t)fin_n * p.H + fin_c) = o;
      } else if (p.dh0 != nullptr) {
        *reinterpret_cast<float4*>(p.dh0 + (int64_t)fin_n * p.H + fin_c) = acc;
      }
    }
    // ---- publish: every store of this CTA happens-before the counter increment
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      red_release_gpu_add(my_flag, 1u);
    }
  }

  // peers may still be reading this CTA's last partial through DSMEM: leave together
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static size_t rs_smem_bytes(int nb, int ks) {
  return 1024 + 2 * (size_t)(ks / 32) * nb * 128 + 2 * (size_t)nb * RS_BM * 4 + 64;
}

template <int NB>
static cudaError_t rs_prepare(int ks) {
  static bool done[3] = {false, false, false};
  const int slot = ks == 32 ? 0 : (ks == 64 ? 1 : 2);
  if (done[slot]) return cudaSuccess;
  // the attribute is per function: set it to the largest footprint once
  cudaError_t e = cudaFuncSetAttribute(rnn_seq_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)rs_smem_bytes(NB, 128));
  if (e == cudaSuccess) done[slot] = true;
  return e;
}

template <int NB>
static int rs_max_clusters(int ks) {
  if (rs_prepare<NB>(ks) != cudaSuccess) { cudaGetLastError(); return 0; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(RS_CL * 16);
  cfg.blockDim = dim3(RS_THREADS);
  cfg.dynamicSmemBytes = rs_smem_bytes(NB, ks);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = RS_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, rnn_seq_kernel<NB>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

struct RsPlan { int nb, n_split, ks, clusters; bool ok; };

static RsPlan rs_plan(int64_t N, int64_t H) {
  RsPlan pl{0, 0, 0, 0, false};
  if (N < 1 || N > 64) return pl;
  if (H != 256 && H != 512 && H != 1024) return pl;            // K-slice H/8 in {32, 64, 128}: divides the 128-column slices
  pl.ks = (int)(H / RS_CL);
  const int slices = (int)(H / RS_BM);
  static int max32[3] = {-1, -1, -1}, max64[3] = {-1, -1, -1};
  const int slot = pl.ks == 32 ? 0 : (pl.ks == 64 ? 1 : 2);
  if (max32[slot] < 0) max32[slot] = rs_max_clusters<32>(pl.ks);
  if (max64[slot] < 0) max64[slot] = rs_max_clusters<64>(pl.ks);
  if (N <= 32) {
    pl.nb = 32; pl.n_split = 1;
    pl.ok = max32[slot] >= slices;
  } else if (max32[slot] >= 2 * slices) {
    pl.nb = 32; pl.n_split = 2; pl.ok = true;
  } else {
    pl.nb = 64; pl.n_split = 1;
    pl.ok = max64[slot] >= slices;
  }
  pl.clusters = slices * pl.n_split;
  return pl;
}

static int rs_launch(const RnnSeqParams& prm, const RsPlan& pl, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(prm.flags, 0, sizeof(unsigned int) * 64, s);
  if (e != cudaSuccess) return set_error(B200AD_ERR_CUDA, "rnn_seq: memset: %s", cudaGetErrorString(e));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(RS_CL * pl.clusters);
  cfg.blockDim = dim3(RS_THREADS);
  cfg.dynamicSmemBytes = rs_smem_bytes(pl.nb, pl.ks);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = RS_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  e = pl.nb == 32 ? cudaLaunchKernelEx(&cfg, rnn_seq_kernel<32>, prm) : cudaLaunchKernelEx(&cfg, rnn_seq_kernel<64>, prm);
  if (e != cudaSuccess) return set_error(B200AD_ERR_CUDA, "rnn_seq: launch: %s", cudaGetErrorString(e));
  count_launch();
  return check_launch("rnn_seq");
}

}  // namespace b200ad

using namespace b200ad;

extern "C" {

int b200ad_rnn_seq_supported(int64_t batch, int64_t hidden) { return rs_plan(batch, hidden).ok ? 1 : 0; }

int64_t b200ad_rnn_seq_ws_bytes(void) { return 64 * (int64_t)sizeof(unsigned int); }

int b200ad_rnn_tanh_seq_fwd(float* hs_ext, const float* wh, int64_t T, int64_t N, int64_t H, void* ws, void* stream) {
  B200AD_REQUIRE(T >= 0 && N >= 1 && H >= 1, "rnn_tanh_seq_fwd: bad shape");
  if (T == 0) return B200AD_OK;
  B200AD_REQUIRE(hs_ext && wh && ws, "rnn_tanh_seq_fwd: null pointer");
  B200AD_REQUIRE(aligned16(hs_ext) && aligned16(wh), "rnn_tanh_seq_fwd: operands must be 16-byte aligned");
  const RsPlan pl = rs_plan(N, H);
  if (!pl.ok) return set_error(B200AD_ERR_UNSUPPORTED, "rnn_tanh_seq_fwd: N=%lld H=%lld not supported by the persistent kernel "
                               "(N <= 64, H in {256, 512, 1024}, all clusters co-resident)", (long long)N, (long long)H);
  RnnSeqParams prm{};
  prm.w = wh; prm.ldw = H; prm.hs = hs_ext; prm.flags = static_cast<unsigned int*>(ws);
  prm.T = (int32_t)T; prm.N = (int32_t)N; prm.H = (int32_t)H; prm.ks = pl.ks; prm.n_split = pl.n_split; prm.bwd = 0;
  return rs_launch(prm, pl, (cudaStream_t)stream);
}

int b200ad_rnn_tanh_seq_bwd(const float* dhs, int64_t dhs_stride_n, int64_t dhs_stride_t, const float* hs_ext,
                            const float* wh, float* dA, float* dh0, int64_t T, int64_t N, int64_t H, void* ws,
                            void* stream) {
  B200AD_REQUIRE(T >= 0 && N >= 1 && H >= 1, "rnn_tanh_seq_bwd: bad shape");
  if (T == 0) return B200AD_OK;
  B200AD_REQUIRE(dhs && hs_ext && wh && dA && ws, "rnn_tanh_seq_bwd: null pointer");
  B200AD_REQUIRE(aligned16(dhs) && aligned16(hs_ext) && aligned16(wh) && aligned16(dA) && (!dh0 || aligned16(dh0)) &&
                 (dhs_stride_n & 3) == 0 && (dhs_stride_t & 3) == 0,
                 "rnn_tanh_seq_bwd: operands must be 16-byte aligned with strides that are multiples of 4 floats");
  const RsPlan pl = rs_plan(N, H);
  if (!pl.ok) return set_error(B200AD_ERR_UNSUPPORTED, "rnn_tanh_seq_bwd: N=%lld H=%lld not supported by the persistent kernel",
                               (long long)N, (long long)H);
  RnnSeqParams prm{};
  prm.w = wh; prm.ldw = H; prm.hs = const_cast<float*>(hs_ext); prm.dhs = dhs; prm.dhs_sn = dhs_stride_n;
  prm.dhs_st = dhs_stride_t; prm.dA = dA; prm.dh0 = dh0; prm.flags = static_cast<unsigned int*>(ws);
  prm.T = (int32_t)T; prm.N = (int32_t)N; prm.H = (int32_t)H; prm.ks = pl.ks; prm.n_split = pl.n_split; prm.bwd = 1;
  return rs_launch(prm, pl, (cudaStream_t)stream);
}
}
