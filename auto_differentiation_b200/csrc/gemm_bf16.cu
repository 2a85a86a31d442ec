// bf16 tensor-core GEMM variant (B200AD_GEMM_BF16): fp32 tensors in HBM, operands rounded to bf16 (RN-even) by a
// conversion pre-pass, tcgen05.mma kind::f16 with fp32 accumulation in TMEM, fp32 output.  Looser tolerance than the
// default 3xTF32 path (bf16 inputs: 8-bit mantissa; the stated bound is |err| <= 2^-8 * sum_k |a_mk b_kn|, tests use
// rtol 2e-2 of that row-norm bound) — the north_star's "bf16 tensor-core GEMM variants to a stated looser tolerance".
//
// Kernel: CTA pair (cta_group::2), 256 x 256 tiles, BK = 64 bf16 (128 B rows, SWIZZLE_128B, both operands K-major: the
// conversion pass also transposes whatever is stored the other way), 6-stage TMA ring, two TMEM accumulator buffers
// (the epilogue of tile i overlaps the main loop of tile i+1), 8 epilogue warps, compact epilogue through per-warp
// staging tiles + cp.async.bulk.tensor stores (see gemm_tc3_kernel for why the epilogue code must stay small).
//   warp 0 TMA (both CTAs)   warp 1 MMA (leader CTA)   warps 2-3 idle   warps 4-11 epilogue
// Split-K (deterministic two-pass, same reduce kernel as the 3xTF32 path) covers the skinny weight-gradient shapes.
#include "common.cuh"
#include "gemm_internal.cuh"
#include "tcgen05.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

namespace b200ad {

constexpr int BF_THREADS = 384;
constexpr int BF_STAGES = 6;
constexpr int BF_BK = 64;                                    // bf16 elements per k-block (128 B)
constexpr uint32_t BF_STAGE_BYTES = 2 * 128 * BF_BK * 2;     // A 128 rows + B half 128 rows
constexpr uint32_t BF_SMEM = BF_STAGES * BF_STAGE_BYTES + 8 * 4096 + 1024;

struct BfParams {
  float* c;
  const float* bias;
  float* partial;
  int64_t m, n, k;
  int64_t ldc, c_stride0, c_stride1, batch1;
  int32_t batches, m_tiles, n_tiles, splits, kblocks_total, kblocks_per_split;
  int32_t act, tma_store;
  float alpha, beta;
};

__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// TMA load whose completion bytes are signalled on the LEADER CTA's mbarrier (cta_group::2): the peer's tiles land in the
// peer's shared memory but count towards the barrier the leader's MMA warp waits on — no relay hop.
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* map, uint32_t leader_bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ uint4 bf_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void bf_sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <int ACT>
__device__ __forceinline__ float bf_act(float x) {
  if constexpr (ACT == 1) return tanhf(x);
  else if constexpr (ACT == 2) return 1.f / (1.f + expf(-x));
  else if constexpr (ACT == 3) return (x > 0.f || x != x) ? x : 0.f;
  else return x;
}

// Tile order: groups of 8 m-tiles, n fastest inside a group, so that the ~74 tiles in flight share 8 A row-panels and
// ~9 B column-panels (L2-resident) instead of 2 A panels and every B panel (streamed from HBM).
__device__ __forceinline__ void bf_decode_tile(int tile, const BfParams& p, int& m_tile, int& n_tile, int& split, int& batch) {
  const int per_mat = p.m_tiles * p.n_tiles;
  int t = tile / per_mat;
  const int r = tile - t * per_mat;
  split = t % p.splits;
  batch = t / p.splits;
  constexpr int GM = 8;
  const int group = r / (GM * p.n_tiles);
  const int gm = (p.m_tiles - group * GM) < GM ? (p.m_tiles - group * GM) : GM;    // m-tiles in this group
  const int rr = r - group * GM * p.n_tiles;
  m_tile = group * GM + rr % gm;
  n_tile = rr / gm;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BF_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_c, const BfParams p) {
  constexpr int BN = 256, R = BF_STAGES;
  constexpr uint32_t A_BYTES = 128 * BF_BK * 2, B_BYTES = 128 * BF_BK * 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[2 * R + 4];
  __shared__ uint32_t tmem_base_smem;
  uint64_t* full = bars;                 // [R] (leader's copy is the live one) TMA of BOTH CTAs -> leader's MMA warp
  uint64_t* empty = bars + R;            // [R] MMA commit (multicast) -> TMA of both CTAs
  uint64_t* tmem_full = bars + 2 * R;    // [2] MMA commit (multicast) -> epilogues
  uint64_t* tmem_empty = tmem_full + 2;  // [2] epilogues of both CTAs (512 arrivals) -> leader's MMA warp

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_c);
    for (int s = 0; s < R; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * 256);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(&tmem_base_smem, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int total_tiles = p.batches * p.splits * p.m_tiles * p.n_tiles;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) =====================================
    // Each CTA loads its own A rows and B-tile half into its own shared memory; the completion bytes of BOTH CTAs are
    // signalled on the leader's full barrier (cta_group::2 TMA), which the leader arms with the total byte count.
    // (The peer's complete_tx may reach the barrier before the leader's expect_tx: the transaction count goes
    // transiently negative, the phase cannot complete before the leader's arrival.)
    if (elect_one()) {
      uint32_t it = 0;
      uint32_t leader_full0;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(leader_full0) : "r"(smem_u32(&full[0])), "r"(0));
      for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
        int m_tile, n_tile, split, batch;
        bf_decode_tile(tile, p, m_tile, n_tile, split, batch);
        const int kb0 = split * p.kblocks_per_split;
        int kb1 = kb0 + p.kblocks_per_split;
        if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
        const int m0 = m_tile * 256 + (int)rank * 128, n0 = n_tile * BN + (int)rank * 128;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % R;
          mbar_wait(&empty[s], ((it / R) & 1) ^ 1);
          uint8_t* sa = smem + (size_t)s * BF_STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(&full[s], 2 * (A_BYTES + B_BYTES));
          const uint32_t bar = leader_full0 + (uint32_t)s * 8u;
          tma_load_4d_pair(&map_a, bar, sa, kb * BF_BK, m0, batch, 0);      // operand batches are flattened
          tma_load_4d_pair(&map_b, bar, sa + A_BYTES, kb * BF_BK, n0, batch, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===================================== MMA issuer (leader CTA) =====================================
      // idesc: D = f32, A = B = bf16 (format 1), both K-major, N = 256, M = 256
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      uint32_t it = 0, tl = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += nclusters, ++tl) {
        int m_tile, n_tile, split, batch;
        bf_decode_tile(tile, p, m_tile, n_tile, split, batch);
        const int kb0 = split * p.kblocks_per_split;
        int kb1 = kb0 + p.kblocks_per_split;
        if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
        const uint32_t buf = tl & 1;
        mbar_wait(&tmem_empty[buf], ((tl >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % R;
          const uint32_t ph = (it / R) & 1;
          mbar_wait(&full[s], ph);                           // both CTAs' tiles of this stage have landed
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + (size_t)s * BF_STAGE_BYTES);
            const uint64_t da = make_smem_desc(sa, 16, 1024, 2), db = make_smem_desc(sa + A_BYTES, 16, 1024, 2);
            const uint32_t d = tmem_base + buf * BN;
#pragma unroll
            for (int kk = 0; kk < BF_BK / 16; ++kk)
              umma2_bf16(d, da + (uint64_t)((kk * 32) >> 4), db + (uint64_t)((kk * 32) >> 4), idesc, (kb == kb0 && kk == 0) ? 0u : 1u);
            umma2_commit_mc(&empty[s]);
            if (kb == kb1 - 1) umma2_commit_mc(&tmem_full[buf]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue (both CTAs, own 128 rows) =====================================
    const int q = warp & 3, hf = (warp - 4) >> 2;
    const uint32_t stg = smem_u32(smem + (size_t)R * BF_STAGE_BYTES) + (uint32_t)(warp - 4) * 4096u;
    const int sub_row = lane >> 3, chunk = lane & 7;
    uint32_t tl = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += nclusters, ++tl) {
      int m_tile, n_tile, split, batch;
      bf_decode_tile(tile, p, m_tile, n_tile, split, batch);
      const int64_t b0 = batch / p.batch1, b1 = batch % p.batch1;
      const uint32_t buf = tl & 1;
      mbar_wait(&tmem_full[buf], (tl >> 1) & 1);
      tc_fence_after();
      const uint32_t x_addr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + (uint32_t)hf * 128u;
      const int m_row0 = m_tile * 256 + (int)rank * 128 + q * 32;
      const int c2 = p.partial ? batch : (int)b1, c3 = p.partial ? split : (int)b0;
      const bool identity = p.partial || (p.alpha == 1.f && p.beta == 0.f && !p.bias && p.act == 0);
      float* cwarp;
      if (p.partial) cwarp = p.partial + (((int64_t)split * p.batches + batch) * p.m + m_row0) * p.n;
      else cwarp = p.c + b0 * p.c_stride0 + b1 * p.c_stride1 + (int64_t)m_row0 * p.ldc;
      const int64_t row_pitch = p.partial ? p.n : p.ldc;
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        uint32_t x[32];
        tmem_ld_32x32b_x32(x_addr + g * 32, x);
        tmem_ld_wait();
        if (g == 3) {                                        // this thread has read its part of the accumulator
          tc_fence_before();
          mbar_arrive_cluster(&tmem_empty[buf], 0);
        }
        const int n_col0 = n_tile * BN + hf * 128 + g * 32;
        if (n_col0 < p.n && m_row0 < p.m) {                  // warp-uniform
          if (p.tma_store && elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 v4;
            v4.x = x[4 * c]; v4.y = x[4 * c + 1]; v4.z = x[4 * c + 2]; v4.w = x[4 * c + 3];
            bf_sts128(stg + lane * 128 + ((c ^ (lane & 7)) << 4), v4);
          }
          __syncwarp();
          if (!identity || !p.tma_store) {
            const int gn0 = n_col0 + chunk * 4;
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
              const int rl = 4 * i + sub_row;
              const uint32_t sa = stg + rl * 128 + ((chunk ^ (rl & 7)) << 4);
              const uint4 v4 = bf_lds128(sa);
              float v[4] = {__uint_as_float(v4.x), __uint_as_float(v4.y), __uint_as_float(v4.z), __uint_as_float(v4.w)};
              float* dst = cwarp + (int64_t)rl * row_pitch + gn0;
              const bool in_range = (m_row0 + rl < p.m) && (gn0 < p.n);
              if (!p.partial) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  v[j] *= p.alpha;
                  if (in_range && gn0 + j < p.n) {
                    if (p.beta != 0.f) v[j] += p.beta * dst[j];
                    if (p.bias) v[j] += __ldg(p.bias + gn0 + j);
                  }
                  if (p.act == 1) v[j] = bf_act<1>(v[j]);
                  else if (p.act == 2) v[j] = bf_act<2>(v[j]);
                  else if (p.act == 3) v[j] = bf_act<3>(v[j]);
                }
              }
              if (p.tma_store) {
                uint4 o4;
                o4.x = __float_as_uint(v[0]); o4.y = __float_as_uint(v[1]); o4.z = __float_as_uint(v[2]); o4.w = __float_as_uint(v[3]);
                bf_sts128(sa, o4);
              } else if (in_range) {
                if (gn0 + 4 <= p.n && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                  *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    if (gn0 + j < p.n) dst[j] = v[j];
                }
              }
            }
          }
          if (p.tma_store) {
            fence_proxy_async();
            __syncwarp();
            if (elect_one()) {          // the same (lowest) lane every time: bulk groups are per thread
              asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&map_c),
                           "r"(stg), "r"(n_col0), "r"(m_row0), "r"(c2), "r"(c3)
                           : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        }
      }
    }
    if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 1) tmem_dealloc2(tmem_base, 512);
}

// fp32 -> bf16 (round to nearest even) into a compact K-major [batch][rows][kpad] buffer.
//   src element (r, k) of batch (b0, b1) at src + b0*s0 + b1*s1 + (trans ? k*ld + r : r*ld + k)
__global__ void __launch_bounds__(256) cvt_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t rows,
                                                       int64_t k, int64_t kpad, int64_t ld, int trans, int64_t batch1, int64_t s0,
                                                       int64_t s1) {
  __shared__ float tile[32][33];
  const int64_t batch = blockIdx.z;
  const float* sb = src + (batch / batch1) * s0 + (batch % batch1) * s1;
  __nv_bfloat16* db = dst + batch * rows * kpad;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  const int64_t r0 = (int64_t)blockIdx.y * 32, k0 = (int64_t)blockIdx.x * 32;
  if (!trans) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t r = r0 + ty + 8 * i, kk = k0 + tx;
      if (r < rows && kk < kpad) db[r * kpad + kk] = __float2bfloat16_rn(kk < k ? sb[r * ld + kk] : 0.f);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {          // read (k, r) with r contiguous
      const int64_t kk = k0 + ty + 8 * i, r = r0 + tx;
      tile[ty + 8 * i][tx] = (kk < k && r < rows) ? sb[kk * ld + r] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {          // write (r, k) with k contiguous
      const int64_t r = r0 + ty + 8 * i, kk = k0 + tx;
      if (r < rows && kk < kpad) db[r * kpad + kk] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
    }
  }
}

__global__ void __launch_bounds__(256) bf16_splitk_reduce_kernel(const float* __restrict__ partial, int splits, BfParams p) {
  const int64_t per_batch = p.m * p.n, total = per_batch * p.batches;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += __ldg(partial + (int64_t)sp * total + i);
    const int64_t batch = i / per_batch, rem = i - batch * per_batch;
    const int64_t gm = rem / p.n, gn = rem - gm * p.n;
    float* dst = p.c + (batch / p.batch1) * p.c_stride0 + (batch % p.batch1) * p.c_stride1 + gm * p.ldc + gn;
    float v = p.alpha * s;
    if (p.beta != 0.f) v += p.beta * *dst;
    if (p.bias) v += __ldg(p.bias + gn);
    if (p.act == 1) v = bf_act<1>(v);
    else if (p.act == 2) v = bf_act<2>(v);
    else if (p.act == 3) v = bf_act<3>(v);
    *dst = v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
struct BfPlan {
  int64_t kpad, a_bytes, b_bytes, part_bytes;
  int m_tiles, n_tiles, batches, kblocks, splits, kblocks_per_split;
};

static BfPlan bf_plan(const GemmArgs& g) {
  BfPlan pl;
  pl.kpad = ceil_div(g.k, 8) * 8;                              // TMA row stride must be a multiple of 16 B
  pl.batches = (int)(g.batch0 * g.batch1);
  pl.a_bytes = ceil_div(pl.batches * g.m * pl.kpad * 2, 1024) * 1024;
  pl.b_bytes = ceil_div(pl.batches * g.n * pl.kpad * 2, 1024) * 1024;
  pl.m_tiles = (int)ceil_div(g.m, 256);
  pl.n_tiles = (int)ceil_div(g.n, 256);
  pl.kblocks = (int)ceil_div(g.k, BF_BK);
  const int64_t tiles = (int64_t)pl.m_tiles * pl.n_tiles * pl.batches;
  const int units = num_sms() / 2;
  int splits = 1;
  if (tiles < units && pl.kblocks >= 16) {
    splits = (int)(units / tiles);
    if (splits > pl.kblocks / 8) splits = pl.kblocks / 8;      // >= 512 K-elements per split
    if (splits < 1) splits = 1;
    if (splits > 64) splits = 64;
  }
  pl.kblocks_per_split = (int)ceil_div(pl.kblocks, splits);
  pl.splits = (int)ceil_div(pl.kblocks, pl.kblocks_per_split);
  pl.part_bytes = pl.splits > 1 ? (int64_t)pl.splits * pl.batches * g.m * g.n * 4 : 0;
  return pl;
}

bool gemm_bf16_supported(const GemmArgs& g, const char** why) {
  auto no = [&](const char* w) { if (why) *why = w; return false; };
  if (g.m < 1 || g.n < 1 || g.k < 1) return no("empty problem");
  if (g.m >= (1ll << 31) || g.n >= (1ll << 31) || g.k >= (1ll << 31)) return no("dimension too large");
  if (g.batch0 * g.batch1 >= 65536) return no("too many batches");
  if (g.batch0 > 1 && (g.a_stride0 < 0 || g.b_stride0 < 0)) return no("negative batch stride");
  if (get_encode_fn() == nullptr) return no("cuTensorMapEncodeTiled unavailable (no driver)");
  return true;
}

int64_t gemm_bf16_workspace_bytes(const GemmArgs& g) {
  if (g.m < 1 || g.n < 1 || g.k < 1) return 0;
  BfPlan pl = bf_plan(g);
  return pl.a_bytes + pl.b_bytes + pl.part_bytes;
}

static int encode_bf16(CUtensorMap* map, const void* base, int64_t rows, int64_t kpad, int64_t batches) {
  EncodeTiledFn enc = get_encode_fn();
  cuuint64_t dims[4] = {(cuuint64_t)kpad, (cuuint64_t)rows, (cuuint64_t)batches, 1};
  cuuint64_t strides[3] = {(cuuint64_t)kpad * 2, (cuuint64_t)rows * kpad * 2, (cuuint64_t)batches * rows * kpad * 2};
  cuuint32_t box[4] = {(cuuint32_t)BF_BK, 128u, 1u, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(B200AD_ERR_CUDA, "gemm(bf16): cuTensorMapEncodeTiled failed (%d)", (int)r);
  return B200AD_OK;
}

int gemm_bf16_launch(const GemmArgs& g, void* ws, int64_t ws_bytes, cudaStream_t s) {
  BfPlan pl = bf_plan(g);
  if (!ws || ws_bytes < pl.a_bytes + pl.b_bytes)
    return set_error(B200AD_ERR_INVALID, "gemm(bf16): workspace of %lld bytes required", (long long)(pl.a_bytes + pl.b_bytes + pl.part_bytes));
  if (ws_bytes < pl.a_bytes + pl.b_bytes + pl.part_bytes) {   // no room for split-K partials: run un-split
    pl.splits = 1;
    pl.kblocks_per_split = pl.kblocks;
    pl.part_bytes = 0;
  }
  __nv_bfloat16* abuf = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* bbuf = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(ws) + pl.a_bytes);
  float* part = pl.splits > 1 ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + pl.a_bytes + pl.b_bytes) : nullptr;
  // conversion passes: A -> [batch][M][kpad], B -> [batch][N][kpad] (both K-major).
  // B200AD_BF16_SKIP_CVT=1 (diagnostic) times the tensor-core kernel alone on whatever the workspace holds.
  if (!getenv("B200AD_BF16_SKIP_CVT")) {
    dim3 ga((unsigned)ceil_div(pl.kpad, 32), (unsigned)ceil_div(g.m, 32), (unsigned)pl.batches);
    cvt_bf16_kernel<<<ga, 256, 0, s>>>(g.a, abuf, g.m, g.k, pl.kpad, g.lda, g.trans_a ? 1 : 0, g.batch1, g.a_stride0, g.a_stride1);
    dim3 gb((unsigned)ceil_div(pl.kpad, 32), (unsigned)ceil_div(g.n, 32), (unsigned)pl.batches);
    cvt_bf16_kernel<<<gb, 256, 0, s>>>(g.b, bbuf, g.n, g.k, pl.kpad, g.ldb, g.trans_b ? 0 : 1, g.batch1, g.b_stride0, g.b_stride1);
    count_launch(2);
    int rc = check_launch("gemm(bf16) conversion");
    if (rc) return rc;
  }
  CUtensorMap map_a, map_b, map_c;
  int rc;
  if ((rc = encode_bf16(&map_a, abuf, g.m, pl.kpad, pl.batches))) return rc;
  if ((rc = encode_bf16(&map_b, bbuf, g.n, pl.kpad, pl.batches))) return rc;
  BfParams p{};
  p.c = g.c; p.bias = g.bias; p.partial = part;
  p.m = g.m; p.n = g.n; p.k = g.k;
  p.ldc = g.ldc; p.c_stride0 = g.c_stride0; p.c_stride1 = g.c_stride1; p.batch1 = 1;   // operand batches are flattened
  p.batches = pl.batches; p.m_tiles = pl.m_tiles; p.n_tiles = pl.n_tiles; p.splits = pl.splits;
  p.kblocks_total = pl.kblocks; p.kblocks_per_split = pl.kblocks_per_split;
  p.act = g.act; p.alpha = g.alpha; p.beta = g.beta;
  // C addressing keeps the caller's two batch levels: batch = b0 * batch1 + b1
  p.batch1 = g.batch1;
  map_c = map_a;
  p.tma_store = 0;
  {
    const float* cbase = part ? part : g.c;
    const int64_t ld = part ? g.n : g.ldc;
    const int64_t s1 = part ? g.m * g.n : g.c_stride1, s0 = part ? (int64_t)pl.batches * g.m * g.n : g.c_stride0;
    const int64_t d2 = part ? pl.batches : g.batch1, d3 = part ? pl.splits : g.batch0;
    const bool ok = (part || g.beta == 0.f) && aligned16(cbase) && (ld & 3) == 0 && (d2 <= 1 || (s1 > 0 && (s1 & 3) == 0)) &&
                    (d3 <= 1 || (s0 > 0 && (s0 & 3) == 0));
    if (ok) {
      EncodeTiledFn enc = get_encode_fn();
      const int64_t fb = ld * g.m;
      cuuint64_t dims[4] = {(cuuint64_t)g.n, (cuuint64_t)g.m, (cuuint64_t)d2, (cuuint64_t)d3};
      cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)(d2 > 1 ? s1 : fb) * 4, (cuuint64_t)(d3 > 1 ? s0 : fb) * 4};
      cuuint32_t box[4] = {32u, 32u, 1u, 1u};
      cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
      CUresult r = enc(&map_c, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(cbase), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r == CUDA_SUCCESS) p.tma_store = 1;
    }
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM);
    if (e != cudaSuccess) return set_error(B200AD_ERR_CUDA, "gemm(bf16): smem attribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  const int64_t total_tiles = (int64_t)pl.batches * pl.splits * pl.m_tiles * pl.n_tiles;
  const int pairs = num_sms() / 2;
  const int grid = 2 * (int)(total_tiles < pairs ? total_tiles : pairs);
  gemm_bf16_kernel<<<grid, BF_THREADS, BF_SMEM, s>>>(map_a, map_b, map_c, p);
  count_launch();
  rc = check_launch("gemm(bf16)");
  if (rc) return rc;
  if (pl.splits > 1) {
    const int64_t total = g.m * g.n * pl.batches;
    bf16_splitk_reduce_kernel<<<stream_grid(total, 256), 256, 0, s>>>(part, pl.splits, p);
    count_launch();
    rc = check_launch("gemm(bf16 split-K reduce)");
  }
  return rc;
}

}  // namespace b200ad
