// fp32-faithful tensor-core GEMM at the kind::f16 rate (B200AD_GEMM_F16X3): every operand element is carried as an fp16
// hi / lo pair (11 + 11 mantissa bits, the same 22 bits the 3xTF32 split carries) times one power of two per operand row,
// and D = Al.Bh + Ah.Bl + Ah.Bh runs as three kind::f16 tcgen05.mma per k-step — K = 16 per instruction in the cycles a
// kind::tf32 instruction covers K = 8 (profiles/r1_mma_rate.log), so the fp32-faithful ceiling doubles.  Numerics:
// profiles/fp16_split_probe.py (row-scaled fp16 split 2-3x more accurate than 3xTF32; scaling is mandatory, fp16 has 5
// exponent bits).  Reference calls replaced: the same np.matmul / np.dot sites as gemm_tc.cu (Matmul tgrad.py:523-551,
// TimeAffine layers.py:280-308, MultiHeadAttention projections :1347-1349,:1360).
//
// Two variants of one kernel (gemm_h16_kernel<A_RAW>), both a CTA pair (cta_group::2) on 256 x 256 tiles, BK = 64 (128 B rows,
// SWIZZLE_128B), a 3-stage ring of {A hi, A lo, B hi, B lo} tiles (64 KB per stage and CTA), 12 MMAs per k-block, chunked
// accumulation (a chain covers 128 K-elements in one of TWO TMEM accumulator buffers; eight fold warps add it into a
// register-resident running total with round-to-nearest adds — TMEM accumulation truncates, DESIGN 3.1 — while the next chain
// fills the other buffer), the finished tile scaled by invA[row] invB[col] (exact: powers of two) on its way through two 2 KB
// swizzled staging tiles per warp + cp.async.bulk.tensor stores:
//   * planes (A_RAW = false): a pre-pass writes K-major fp16 hi / lo planes + the inverse row scales of BOTH operands into the
//     caller's workspace (one fused pass for row-major operands: row maximum, scale, split; column maximum + transposing split
//     for the others); the kernel is pure TMA + MMA, the loads of both CTAs signal the leader's mbarrier.
//   * in-kernel conversion (A_RAW = true; op(A) one row-major matrix, N <= 2048): only A's row scales are computed ahead
//     (read-only pass); TMA delivers the raw fp32 tile and four converter warps rewrite it IN PLACE as the hi / lo tiles (see
//     the kernel).  B always comes as planes (weights: small).
//   warp 0 TMA (both CTAs)   warp 1 MMA (leader CTA)   warps 2-3 idle   warps 4-7 converter (A_RAW) / idle   warps 8-15 fold +
//   epilogue; setmaxnreg 56 / 24 / 208 (planes), 48 / 72 / 192 (A_RAW) of the 512 x 128 registers launched.
// Split-K (deterministic two-pass) covers the skinny shapes.  b200ad_gemm_shared_a prepares ONE A operand for several
// B / C pairs (the Q / K / V projections of self-attention share their input, layers.py:1347-1349).  Measurements: DESIGN 3.5.
#include "common.cuh"
#include "gemm_internal.cuh"
#include "tcgen05.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace b200ad {

constexpr int H_THREADS = 512;                                 // 4 warpgroups; setmaxnreg moves registers to the fold warps
constexpr int H_STAGES = 3;
constexpr int H_BK = 64;                                       // fp16 elements per k-block (128 B)
constexpr int H_CHUNK_KB = 2;                                  // accumulation chain: 128 K-elements
constexpr uint32_t H_TILE_BYTES = 128 * H_BK * 2;              // 16 KB: 128 rows of one plane
constexpr uint32_t H_STAGE_BYTES = 4 * H_TILE_BYTES;           // A hi, A lo, B-half hi, B-half lo
constexpr uint32_t H_SMEM = H_STAGES * H_STAGE_BYTES + 8 * 4096 + 1024;

struct H16Params {
  float* c;
  const float* bias;
  float* partial;
  const float* scale_a;        // A_RAW only: [m] forward row scales of A (the kernel converts the raw fp32 tile itself)
  const float* inv_a;          // [batches][m] inverse row scales of A (powers of two)
  const float* inv_b;          // [batches][n]
  int64_t m, n, k;
  int64_t ldc, c_stride0, c_stride1, batch1;
  int32_t batches, m_tiles, n_tiles, splits, kblocks_total, kblocks_per_split;
  int32_t act, tma_store, probe;
  float alpha, beta;
};

__device__ __forceinline__ void h16_umma2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// TMA load whose completion bytes are signalled on the LEADER CTA's mbarrier (see gemm_bf16.cu)
__device__ __forceinline__ void h16_tma_load_pair(const CUtensorMap* map, uint32_t leader_bar, uint32_t dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ uint4 h16_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void h16_sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float h16_act(float x, int act) {
  if (act == 1) return tanhf(x);
  if (act == 2) return 1.f / (1.f + expf(-x));
  if (act == 3) return (x > 0.f || x != x) ? x : 0.f;
  return x;
}

// scale = 2^(14 - floor(log2 amax)): the largest element of the row lands in [2^14, 2^15); hi = RN16(x s),
// lo = RN16(x s - hi) (the difference is exact in fp32).  Elements below 2^-18 of the row maximum lose relative precision
// to fp16's subnormal floor (absolute error <= 2^-40 of the row maximum).  amax_bits: bits of the non-negative maximum.
__device__ __forceinline__ void h16_scales(uint32_t amax_bits, float& scale, float& inv) {
  int se = 268 - (int)((amax_bits >> 23) & 0xffu);
  se = se < 1 ? 1 : (se > 253 ? 253 : se);
  scale = __uint_as_float((uint32_t)se << 23);
  inv = __uint_as_float((uint32_t)(254 - se) << 23);
}
__device__ __forceinline__ void h16_split(float x, float scale, __half& hi, __half& lo) {
  const float xs = x * scale;
  hi = __float2half_rn(xs);
  lo = __float2half_rn(xs - __half2float(hi));
}
// The in-kernel converter's split of a pair (same scheme as attn_h16.cu's ah_split2): hi = x s TRUNCATED to 11 significant bits
// (a mask; exactly an fp16 value, so the packed conversion is exact), lo = RN16(x s - hi).  One bit less in lo than the
// round-to-nearest split of the pre-pass, a third of its instruction cost (scalar F2F conversions issue at a quarter rate).
__device__ __forceinline__ void h16_split_pair_trunc(float x0, float x1, float scale, uint32_t& hi, uint32_t& lo) {
  const float v0 = x0 * scale, v1 = x1 * scale;
  const float h0 = __uint_as_float(__float_as_uint(v0) & 0xffffe000u), h1 = __uint_as_float(__float_as_uint(v1) & 0xffffe000u);
  const __half2 hh = __floats2half2_rn(h0, h1);
  const __half2 ll = __floats2half2_rn(v0 - h0, v1 - h1);
  hi = *reinterpret_cast<const uint32_t*>(&hh);
  lo = *reinterpret_cast<const uint32_t*>(&ll);
}
__device__ __forceinline__ uint32_t h16_absbits(float x) { return __float_as_uint(x) & 0x7fffffffu; }
__device__ __forceinline__ uint32_t h16_warp_umax(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint32_t w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}
__device__ __forceinline__ uint32_t h16_pack(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// Tile order: groups of 8 m-tiles, n fastest inside a group (the tiles in flight share L2-resident panels).
__device__ __forceinline__ void h16_decode_tile(int tile, const H16Params& p, int& m_tile, int& n_tile, int& split, int& batch) {
  const int per_mat = p.m_tiles * p.n_tiles;
  const int t = tile / per_mat;
  const int r = tile - t * per_mat;
  split = t % p.splits;
  batch = t / p.splits;
  constexpr int GM = 8;
  const int group = r / (GM * p.n_tiles);
  const int gm = (p.m_tiles - group * GM) < GM ? (p.m_tiles - group * GM) : GM;
  const int rr = r - group * GM * p.n_tiles;
  m_tile = group * GM + rr % gm;
  n_tile = rr / gm;
}

// A_RAW: op(A) is row-major fp32 in HBM and is converted INSIDE the kernel (no plane pre-pass, only the row scales): TMA
// delivers the raw 128 x 64 fp32 tile as two 128 B-wide halves (k 0..31 | k 32..63, SWIZZLE_128B) into the 32 KB the hi and lo
// tiles will occupy; converter thread r (warps 4-7) owns row r of BOTH halves and rewrites them IN PLACE as row r of the hi tile
// (where half 0 was) and of the lo tile (where half 1 was) — same swizzle phase (r & 7), no other thread touches those 256
// bytes, so no barrier inside the conversion.  Loads then signal the CTA's own barrier, the converter warps of both CTAs
// arrive on the leader's conv_full (8 arrivals) and the MMA warp waits for that instead.
template <bool A_RAW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(H_THREADS, 1)
gemm_h16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_c, const H16Params p) {
  constexpr int BN = 256, R = H_STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[3 * R + 4];
  __shared__ uint32_t tmem_base_smem;
  uint64_t* conv_full = bars + 2 * R + 4;  // [R] A_RAW: converter warps of BOTH CTAs (8 arrivals) -> leader's MMA warp
  uint64_t* full = bars;                 // [R] (leader's copy is the live one) TMA of BOTH CTAs -> leader's MMA warp
                                         //     A_RAW: every CTA's own copy, its TMA -> its converter warps
  uint64_t* empty = bars + R;            // [R] MMA commit (multicast) -> TMA of both CTAs
  uint64_t* tmem_full = bars + 2 * R;    // [2] MMA commit (multicast) -> fold warps of both CTAs: a chunk sum is complete
  uint64_t* tmem_empty = tmem_full + 2;  // [2] fold warps of both CTAs (16 arrivals) -> leader's MMA warp

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_c);
    for (int s = 0; s < R; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&conv_full[s], 2 * 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * 8);
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

  // register budget: 512 x 128 = 65536 at launch must cover the three setmaxnreg values; an .inc beyond what the .decs release
  // would block forever, so the arithmetic is part of the kernel
  if (warp < 4) {
  if constexpr (A_RAW) asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");    // A_RAW: 128 x 48 + 128 x 72 + 256 x 192 = 64512
  else asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");                    // else:  128 x 56 + 128 x 24 + 256 x 208 = 63488
  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) =====================================
    if (elect_one()) {
      uint32_t it = 0;
      uint32_t leader_full0;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(leader_full0) : "r"(smem_u32(&full[0])), "r"(0));
      for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
        int m_tile, n_tile, split, batch;
        h16_decode_tile(tile, p, m_tile, n_tile, split, batch);
        const int kb0 = split * p.kblocks_per_split;
        int kb1 = kb0 + p.kblocks_per_split;
        if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
        const int m0 = m_tile * 256 + (int)rank * 128, n0 = n_tile * BN + (int)rank * 128;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % R;
          mbar_wait(&empty[s], ((it / R) & 1) ^ 1);
          const uint32_t sa = smem_u32(smem + (size_t)s * H_STAGE_BYTES);
          const int k0 = kb * H_BK;
          if constexpr (A_RAW) {
            uint8_t* st = smem + (size_t)s * H_STAGE_BYTES;
            mbar_expect_tx(&full[s], H_STAGE_BYTES);
            tma_load_4d(&map_a, &full[s], st, k0, m0, 0, 0);                             // raw fp32, k0 .. k0+31
            tma_load_4d(&map_a, &full[s], st + H_TILE_BYTES, k0 + 32, m0, 0, 0);         // raw fp32, k0+32 .. k0+63
            tma_load_4d(&map_b, &full[s], st + 2 * H_TILE_BYTES, k0, n0, batch, 0);      // B hi (this CTA's 128 columns)
            tma_load_4d(&map_b, &full[s], st + 3 * H_TILE_BYTES, k0, n0, batch, 1);      // B lo
          } else {
            if (rank == 0) mbar_expect_tx(&full[s], 2 * H_STAGE_BYTES);
            const uint32_t bar = leader_full0 + (uint32_t)s * 8u;
            h16_tma_load_pair(&map_a, bar, sa, k0, m0, batch, 0);                        // A hi
            h16_tma_load_pair(&map_a, bar, sa + H_TILE_BYTES, k0, m0, batch, 1);         // A lo
            h16_tma_load_pair(&map_b, bar, sa + 2 * H_TILE_BYTES, k0, n0, batch, 0);     // B hi (this CTA's 128 columns)
            h16_tma_load_pair(&map_b, bar, sa + 3 * H_TILE_BYTES, k0, n0, batch, 1);     // B lo
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===================================== MMA issuer (leader CTA) =====================================
      // idesc: D = f32 (bit 4), A = B = f16 (format 0), both K-major, N = 256, M = 256
      constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      uint32_t it = 0, uit = 0;
      const bool prof = (p.probe & 8) != 0;                      // diagnostic: where the issuing warp waits
      long long w_full = 0, w_empty = 0;
      const long long t_start = clock64();
      for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
        int m_tile, n_tile, split, batch;
        h16_decode_tile(tile, p, m_tile, n_tile, split, batch);
        const int kb0 = split * p.kblocks_per_split;
        int kb1 = kb0 + p.kblocks_per_split;
        if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
        for (int cb0 = kb0; cb0 < kb1; cb0 += H_CHUNK_KB, ++uit) {
          const int cb1 = (cb0 + H_CHUNK_KB < kb1) ? cb0 + H_CHUNK_KB : kb1;
          const uint32_t buf = uit & 1;
          long long q0 = prof ? clock64() : 0;
          mbar_wait(&tmem_empty[buf], ((uit >> 1) & 1) ^ 1);     // the fold warps have read this buffer's previous chunk
          if (prof) w_empty += clock64() - q0;
          tc_fence_after();
          for (int kb = cb0; kb < cb1; ++kb, ++it) {
            const int s = it % R;
            q0 = prof ? clock64() : 0;
            mbar_wait(A_RAW ? &conv_full[s] : &full[s], (it / R) & 1);   // both CTAs' four tiles of this stage are ready
            if (prof) w_full += clock64() - q0;
            tc_fence_after();
            if (elect_one()) {
              const uint32_t sa = smem_u32(smem + (size_t)s * H_STAGE_BYTES);
              const uint64_t ah = make_smem_desc(sa, 16, 1024, 2), al = make_smem_desc(sa + H_TILE_BYTES, 16, 1024, 2);
              const uint64_t bh = make_smem_desc(sa + 2 * H_TILE_BYTES, 16, 1024, 2);
              const uint64_t bl = make_smem_desc(sa + 3 * H_TILE_BYTES, 16, 1024, 2);
              const uint32_t d = tmem_base + buf * BN;
              // cross terms first, hi.hi last (as in gemm_tc3_kernel); 32 B per K = 16 step inside the 128 B row
#pragma unroll
              for (int kk = 0; kk < H_BK / 16; ++kk)
                h16_umma2(d, al + (uint64_t)(kk * 2), bh + (uint64_t)(kk * 2), idesc, (kb == cb0 && kk == 0) ? 0u : 1u);
#pragma unroll
              for (int kk = 0; kk < H_BK / 16; ++kk) h16_umma2(d, ah + (uint64_t)(kk * 2), bl + (uint64_t)(kk * 2), idesc, 1u);
#pragma unroll
              for (int kk = 0; kk < H_BK / 16; ++kk) h16_umma2(d, ah + (uint64_t)(kk * 2), bh + (uint64_t)(kk * 2), idesc, 1u);
              umma2_commit_mc(&empty[s]);
              if (kb == cb1 - 1) umma2_commit_mc(&tmem_full[buf]);
            }
            __syncwarp();
          }
        }
      }
      if (prof && blockIdx.x == 0 && lane == 0)
        printf("[h16] MMA warp: total %lld cycles, %u k-blocks, %u chunks, wait full (TMA) %lld, wait tmem_empty (fold) %lld\n",
               clock64() - t_start, it, uit, w_full, w_empty);
    }
  }
  } else if (warp < 8) {
    if constexpr (!A_RAW) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    } else {
      // ===================================== converter (both CTAs): raw fp32 A tile -> hi / lo fp16 tiles, in place =====
      asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
      const int r = (int)threadIdx.x - 128;                    // row of this CTA's 128-row A tile
      const uint32_t sw = (uint32_t)(r & 7);
      uint32_t it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
        int m_tile, n_tile, split, batch;
        h16_decode_tile(tile, p, m_tile, n_tile, split, batch);
        const int kb0 = split * p.kblocks_per_split;
        int kb1 = kb0 + p.kblocks_per_split;
        if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
        const int64_t gm = (int64_t)m_tile * 256 + (int64_t)rank * 128 + r;
        const float sc = gm < p.m ? __ldg(p.scale_a + gm) : 1.f;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % R;
          mbar_wait(&full[s], (it / R) & 1);                   // this CTA's raw A halves (and its B planes) have landed
          const uint32_t row0 = smem_u32(smem + (size_t)s * H_STAGE_BYTES) + (uint32_t)r * 128u, row1 = row0 + H_TILE_BYTES;
          // half 1 (k 32..63) goes to registers first: its bytes receive the lo chunks of half 0
          uint4 h1v[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) h1v[c] = h16_lds128(row1 + (((uint32_t)c ^ sw) << 4));
          // half 0 progressively in place: output chunk j (8 elements) needs raw chunks 2j, 2j+1 and overwrites position j <= 2j
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 a = h16_lds128(row0 + (((uint32_t)(2 * j) ^ sw) << 4)), b = h16_lds128(row0 + (((uint32_t)(2 * j + 1) ^ sw) << 4));
            uint4 h, l;
            h16_split_pair_trunc(__uint_as_float(a.x), __uint_as_float(a.y), sc, h.x, l.x);
            h16_split_pair_trunc(__uint_as_float(a.z), __uint_as_float(a.w), sc, h.y, l.y);
            h16_split_pair_trunc(__uint_as_float(b.x), __uint_as_float(b.y), sc, h.z, l.z);
            h16_split_pair_trunc(__uint_as_float(b.z), __uint_as_float(b.w), sc, h.w, l.w);
            h16_sts128(row0 + (((uint32_t)j ^ sw) << 4), h);
            h16_sts128(row1 + (((uint32_t)j ^ sw) << 4), l);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {                          // half 1 from registers -> chunks 4..7 of both tiles
            const uint4 a = h1v[2 * j], b = h1v[2 * j + 1];
            uint4 h, l;
            h16_split_pair_trunc(__uint_as_float(a.x), __uint_as_float(a.y), sc, h.x, l.x);
            h16_split_pair_trunc(__uint_as_float(a.z), __uint_as_float(a.w), sc, h.y, l.y);
            h16_split_pair_trunc(__uint_as_float(b.x), __uint_as_float(b.y), sc, h.z, l.z);
            h16_split_pair_trunc(__uint_as_float(b.z), __uint_as_float(b.w), sc, h.w, l.w);
            h16_sts128(row0 + (((uint32_t)(4 + j) ^ sw) << 4), h);
            h16_sts128(row1 + (((uint32_t)(4 + j) ^ sw) << 4), l);
          }
          fence_proxy_async();                                 // generic-proxy writes -> visible to the MMA's shared-memory reads
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&conv_full[s], 0);  // one (cumulative) release per warp; the leader counts both CTAs
        }
      }
    }
  } else {
    // ===================================== fold + epilogue (both CTAs, own 128 rows) =====================================
    if constexpr (A_RAW) asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int q = warp & 3, hf = (warp - 8) >> 2;              // TMEM lane quarter, 128-column half of the tile
    const uint32_t stg = smem_u32(smem + (size_t)R * H_STAGE_BYTES) + (uint32_t)(warp - 8) * 4096u;
    const uint32_t x_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)hf * 128u;
    uint32_t uit = 0, stc = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
      int m_tile, n_tile, split, batch;
      h16_decode_tile(tile, p, m_tile, n_tile, split, batch);
      const int64_t b0 = batch / p.batch1, b1 = batch % p.batch1;
      const int kb0 = split * p.kblocks_per_split;
      int kb1 = kb0 + p.kblocks_per_split;
      if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
      const int nchunks = (kb1 - kb0 + H_CHUNK_KB - 1) / H_CHUNK_KB;
      // the scales this thread applies at the end are fetched now: their L2 latency hides behind the folds
      const int m_row0 = m_tile * 256 + (int)rank * 128 + q * 32;  // first row of this warp
      const float ra = (m_row0 + lane < p.m) ? __ldg(p.inv_a + (int64_t)batch * p.m + m_row0 + lane) : 0.f;
      float cbl[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = n_tile * BN + hf * 128 + g * 32 + lane;
        cbl[g] = (col < p.n) ? __ldg(p.inv_b + (int64_t)batch * p.n + col) : 0.f;
      }
      float y[128];                                            // running total: this thread's row, 128 columns
      for (int ch = 0; ch < nchunks; ++ch, ++uit) {
        const uint32_t buf = uit & 1;
        mbar_wait(&tmem_full[buf], (uit >> 1) & 1);
        tc_fence_after();
#pragma unroll
        uint32_t dep = 0;                                        // see below
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 16) {
          uint32_t x[16];
          tmem_ld_32x32b_x16(x_addr + buf * BN + (uint32_t)c0 + dep, x);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) y[c0 + j] = (ch == 0) ? __uint_as_float(x[j]) : y[c0 + j] + __uint_as_float(x[j]);
          // ptxas otherwise hoists all eight loads above the adds (128 live load registers next to the 128 totals: spills);
          // a zero it cannot prove (this lane's own value shuffled back) makes the next load's address depend on this sum
          if (c0 & 16) {
            const uint32_t yb = __float_as_uint(y[c0 + 15]);
            dep = __shfl_sync(0xffffffffu, yb, lane) ^ yb;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&tmem_empty[buf], 0);   // the buffer may take the chunk after next
      }
      // ---- tile epilogue (compact on purpose: see gemm_tc3_kernel) ----
      // The scales are applied on the way into the staging tile: this thread owns row m_row0 + lane (one inverse row scale),
      // the 32 inverse column scales of a group are loaded once per warp and broadcast by shuffles.  An identity epilogue
      // then leaves straight through the TMA store; alpha / beta / bias / activation take the rolled loop.
      if (p.probe & 1) continue;                                 // diagnostic: no output (times MMA + fold alone)
      const int c2 = p.partial ? batch : (int)b1, c3 = p.partial ? split : (int)b0;
      const bool identity = p.partial || (p.alpha == 1.f && p.beta == 0.f && !p.bias && p.act == 0);
      float* cwarp;
      if (p.partial) cwarp = p.partial + (((int64_t)split * p.batches + batch) * p.m + m_row0) * p.n;
      else cwarp = p.c + b0 * p.c_stride0 + b1 * p.c_stride1 + (int64_t)m_row0 * p.ldc;
      const int64_t row_pitch = p.partial ? p.n : p.ldc;
      // Output in 16-column groups through TWO 2 KB staging tiles per warp (32 rows x 64 B, SWIZZLE_64B: 16 B chunk c of row r
      // sits at chunk c ^ ((r >> 1) & 3)): while the TMA engine reads one, the warp fills the other (with one 4 KB tile per warp
      // the fill of every group waited for the previous group's store: ~11 k cycles per tile against 12.3 k of MMAs at K = 512).
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int n_col0 = n_tile * BN + hf * 128 + g * 16;
        if (n_col0 < p.n && m_row0 < p.m) {                    // warp-uniform
          // staging tiles alternate by a RUNNING count of issued stores, not by g: with partial column tiles (N % 32 != 0) a
          // tile may issue a single group, and two consecutive stores of this warp must never share a staging tile
          const uint32_t stgb = stg + (stc & 1u) * 2048u;
          ++stc;
          // the store issued two stores ago read this tile; the most recent one (other tile) may still be in flight
          if (p.tma_store && elect_one()) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int col = (g & 1) * 16 + 4 * c;                // column inside the 32-column scale group cbl[g >> 1]
            uint4 v4;                                            // exact: both factors are powers of two
            v4.x = __float_as_uint(y[g * 16 + 4 * c] * ra * __shfl_sync(0xffffffffu, cbl[g >> 1], col));
            v4.y = __float_as_uint(y[g * 16 + 4 * c + 1] * ra * __shfl_sync(0xffffffffu, cbl[g >> 1], col + 1));
            v4.z = __float_as_uint(y[g * 16 + 4 * c + 2] * ra * __shfl_sync(0xffffffffu, cbl[g >> 1], col + 2));
            v4.w = __float_as_uint(y[g * 16 + 4 * c + 3] * ra * __shfl_sync(0xffffffffu, cbl[g >> 1], col + 3));
            h16_sts128(stgb + lane * 64 + (((uint32_t)c ^ (((uint32_t)lane >> 1) & 3u)) << 4), v4);
          }
          __syncwarp();
          if (!identity || !p.tma_store) {
            const int sub_row = lane >> 2, chunk = lane & 3;
            const int gn0 = n_col0 + chunk * 4;
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
              const int rl = 8 * i + sub_row;
              const uint32_t sa = stgb + rl * 64 + (((uint32_t)chunk ^ (((uint32_t)rl >> 1) & 3u)) << 4);
              const uint4 v4 = h16_lds128(sa);
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
                  v[j] = h16_act(v[j], p.act);
                }
              }
              if (p.tma_store) {
                uint4 o4;
                o4.x = __float_as_uint(v[0]); o4.y = __float_as_uint(v[1]); o4.z = __float_as_uint(v[2]); o4.w = __float_as_uint(v[3]);
                h16_sts128(sa, o4);
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
                           "r"(stgb), "r"(n_col0), "r"(m_row0), "r"(c2), "r"(c3)
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

// ---------------------------------------------------------------------------------------------------------
// operand pre-pass: fp32 -> (hi, lo) fp16 planes, K-major [plane][batch][rows][kpad], one power-of-two scale per row
// ---------------------------------------------------------------------------------------------------------
// Row-major source (element (r, k) at r*ld + k): one warp per row does maximum, scale and split in one pass.
// VEC: k % 8 == 0 (kpad == k), 16 B aligned rows: 128-bit loads, 64-bit stores; up to K = 1024 the row stays in registers
// between the two steps, longer rows are read a second time (L1 / L2 hits).
template <bool VEC>
__global__ void __launch_bounds__(256) h16_split_rows_kernel(const float* __restrict__ src, __half* __restrict__ hi,
                                                             __half* __restrict__ lo, float* __restrict__ inv_scale, int64_t rows,
                                                             int64_t k, int64_t kpad, int64_t ld, int64_t batch1, int64_t s0, int64_t s1) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * 8 + warp;
  if (r >= rows) return;
  const int64_t batch = blockIdx.y;
  const float* row = src + (batch / batch1) * s0 + (batch % batch1) * s1 + r * ld;
  const int64_t orow = (batch * rows + r) * kpad;
  float scale, inv;
  if (VEC && k > 1024) {
    const float4* row4 = reinterpret_cast<const float4*>(row);
    const int n4 = (int)(k >> 2);
    uint32_t m = 0;
#pragma unroll 4
    for (int i = lane; i < n4; i += 32) {
      const float4 v = __ldg(row4 + i);
      uint32_t a = h16_absbits(v.x), b = h16_absbits(v.y), c = h16_absbits(v.z), d = h16_absbits(v.w);
      a = a > b ? a : b; c = c > d ? c : d; a = a > c ? a : c;
      m = m > a ? m : a;
    }
    m = h16_warp_umax(m);
    h16_scales(m, scale, inv);
#pragma unroll 4
    for (int i = lane; i < n4; i += 32) {
      const float4 v = __ldg(row4 + i);
      __half h0, h1, h2, h3, l0, l1, l2, l3;
      h16_split(v.x, scale, h0, l0); h16_split(v.y, scale, h1, l1);
      h16_split(v.z, scale, h2, l2); h16_split(v.w, scale, h3, l3);
      *reinterpret_cast<uint2*>(hi + orow + 4 * i) = make_uint2(h16_pack(h0, h1), h16_pack(h2, h3));
      *reinterpret_cast<uint2*>(lo + orow + 4 * i) = make_uint2(h16_pack(l0, l1), h16_pack(l2, l3));
    }
  } else if constexpr (VEC) {
    float4 v[8];
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = (i * 32 + lane) * 4;
      if (idx < k) {
        v[i] = ldg_stream(reinterpret_cast<const float4*>(row + idx));
        uint32_t a = h16_absbits(v[i].x), b = h16_absbits(v[i].y), c = h16_absbits(v[i].z), d = h16_absbits(v[i].w);
        a = a > b ? a : b; c = c > d ? c : d; a = a > c ? a : c;
        m = m > a ? m : a;
      }
    }
    m = h16_warp_umax(m);
    h16_scales(m, scale, inv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = (i * 32 + lane) * 4;
      if (idx < k) {
        __half h0, h1, h2, h3, l0, l1, l2, l3;
        h16_split(v[i].x, scale, h0, l0); h16_split(v[i].y, scale, h1, l1);
        h16_split(v[i].z, scale, h2, l2); h16_split(v[i].w, scale, h3, l3);
        *reinterpret_cast<uint2*>(hi + orow + idx) = make_uint2(h16_pack(h0, h1), h16_pack(h2, h3));
        *reinterpret_cast<uint2*>(lo + orow + idx) = make_uint2(h16_pack(l0, l1), h16_pack(l2, l3));
      }
    }
  } else {
    uint32_t m = 0;
    for (int64_t kk = lane; kk < k; kk += 32) {
      const uint32_t a = h16_absbits(row[kk]);
      m = m > a ? m : a;
    }
    m = h16_warp_umax(m);
    h16_scales(m, scale, inv);
    for (int64_t kk = lane; kk < kpad; kk += 32) {
      __half h, l;
      h16_split(kk < k ? row[kk] : 0.f, scale, h, l);
      hi[orow + kk] = h;
      lo[orow + kk] = l;
    }
  }
  if (lane == 0) inv_scale[batch * rows + r] = inv;
}

// A_RAW path: only the row scales are computed ahead of the GEMM (read-only pass; a warp per row)
__global__ void __launch_bounds__(256) h16_rowscale_kernel(const float* __restrict__ src, float* __restrict__ scale_out,
                                                           float* __restrict__ inv_out, int64_t rows, int64_t k, int64_t ld) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * 8 + warp;
  if (r >= rows) return;
  const float* row = src + r * ld;
  uint32_t m = 0;
  if ((k & 3) == 0) {                                         // callers guarantee 16 B aligned rows
    const float4* row4 = reinterpret_cast<const float4*>(row);
    const int n4 = (int)(k >> 2);
#pragma unroll 4
    for (int i = lane; i < n4; i += 32) {
      const float4 v = ldg_stream(row4 + i);
      uint32_t a = h16_absbits(v.x), b = h16_absbits(v.y), c = h16_absbits(v.z), d = h16_absbits(v.w);
      a = a > b ? a : b; c = c > d ? c : d; a = a > c ? a : c;
      m = m > a ? m : a;
    }
  } else {
    for (int64_t kk = lane; kk < k; kk += 32) {
      const uint32_t a = h16_absbits(row[kk]);
      m = m > a ? m : a;
    }
  }
  m = h16_warp_umax(m);
  float scale, inv;
  h16_scales(m, scale, inv);
  if (lane == 0) {
    scale_out[r] = scale;
    inv_out[r] = inv;
  }
}

// Long rows (1024 < K <= 8192, K % 8 == 0, 16 B aligned): one 256-thread block per row, the row in registers (a warp per row
// walking a 32 KB row twice measured 4x slower than the short-row kernel per element).
__global__ void __launch_bounds__(256) h16_split_rows_block_kernel(const float* __restrict__ src, __half* __restrict__ hi,
                                                                   __half* __restrict__ lo, float* __restrict__ inv_scale,
                                                                   int64_t rows, int64_t k, int64_t ld, int64_t batch1, int64_t s0,
                                                                   int64_t s1) {
  __shared__ uint32_t red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r = blockIdx.x, batch = blockIdx.y;
  const float* row = src + (batch / batch1) * s0 + (batch % batch1) * s1 + r * ld;
  const int64_t orow = (batch * rows + r) * k;
  float4 v[8];
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = (i * 256 + (int)threadIdx.x) * 4;
    if (idx < k) {
      v[i] = ldg_stream(reinterpret_cast<const float4*>(row + idx));
      uint32_t a = h16_absbits(v[i].x), b = h16_absbits(v[i].y), c = h16_absbits(v[i].z), d = h16_absbits(v[i].w);
      a = a > b ? a : b; c = c > d ? c : d; a = a > c ? a : c;
      m = m > a ? m : a;
    }
  }
  m = h16_warp_umax(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) m = red[i] > m ? red[i] : m;
  float scale, inv;
  h16_scales(m, scale, inv);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = (i * 256 + (int)threadIdx.x) * 4;
    if (idx < k) {
      __half h0, h1, h2, h3, l0, l1, l2, l3;
      h16_split(v[i].x, scale, h0, l0); h16_split(v[i].y, scale, h1, l1);
      h16_split(v[i].z, scale, h2, l2); h16_split(v[i].w, scale, h3, l3);
      *reinterpret_cast<uint2*>(hi + orow + idx) = make_uint2(h16_pack(h0, h1), h16_pack(h2, h3));
      *reinterpret_cast<uint2*>(lo + orow + idx) = make_uint2(h16_pack(l0, l1), h16_pack(l2, l3));
    }
  }
  if (threadIdx.x == 0) inv_scale[batch * rows + r] = inv;
}

// Column-major source (element (r, k) at k*ld + r): maximum per r over k (atomicMax on the bits of |x|: order-independent,
// hence deterministic), then a transposing split through a shared-memory tile.
__global__ void __launch_bounds__(256) h16_colmax_kernel(const float* __restrict__ src, uint32_t* __restrict__ amax, int64_t rows,
                                                         int64_t k, int64_t ld, int64_t batch1, int64_t s0, int64_t s1) {
  __shared__ uint32_t red[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t batch = blockIdx.z;
  const float* sb = src + (batch / batch1) * s0 + (batch % batch1) * s1;
  const int64_t r = (int64_t)blockIdx.x * 32 + tx;
  const int64_t k0 = (int64_t)blockIdx.y * 256;
  const int64_t k1 = k0 + 256 < k ? k0 + 256 : k;
  uint32_t m = 0;
  if (r < rows)
    for (int64_t kk = k0 + ty; kk < k1; kk += 8) {
      const uint32_t a = h16_absbits(sb[kk * ld + r]);
      m = m > a ? m : a;
    }
  red[ty][tx] = m;
  __syncthreads();
  if (ty == 0 && r < rows) {
#pragma unroll
    for (int i = 1; i < 8; ++i) m = red[i][tx] > m ? red[i][tx] : m;
    atomicMax(amax + batch * rows + r, m);
  }
}
__global__ void __launch_bounds__(256) h16_split_trans_kernel(const float* __restrict__ src, __half* __restrict__ hi,
                                                              __half* __restrict__ lo, const uint32_t* __restrict__ amax,
                                                              float* __restrict__ inv_scale, int64_t rows, int64_t k, int64_t kpad,
                                                              int64_t ld, int64_t batch1, int64_t s0, int64_t s1) {
  __shared__ float tile[32][33];
  const int64_t batch = blockIdx.z;
  const float* sb = src + (batch / batch1) * s0 + (batch % batch1) * s1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  const int64_t r0 = (int64_t)blockIdx.y * 32, k0 = (int64_t)blockIdx.x * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {          // read (k, r) with r contiguous
    const int64_t kk = k0 + ty + 8 * i, r = r0 + tx;
    tile[ty + 8 * i][tx] = (kk < k && r < rows) ? sb[kk * ld + r] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {          // write (r, k) with k contiguous
    const int64_t r = r0 + ty + 8 * i, kk = k0 + tx;
    if (r < rows && kk < kpad) {
      float scale, inv;
      h16_scales(amax[batch * rows + r], scale, inv);
      __half h, l;
      h16_split(tile[tx][ty + 8 * i], scale, h, l);
      const int64_t o = (batch * rows + r) * kpad + kk;
      hi[o] = h;
      lo[o] = l;
      if (blockIdx.x == 0 && tx == 0) inv_scale[batch * rows + r] = inv;
    }
  }
}

// Small column-major operands (the weights of a forward projection): maximum and transposing split in ONE launch, a block
// per 32 rows r; the second walk over the block's columns hits L1 / L2.
__global__ void __launch_bounds__(256) h16_split_trans_small_kernel(const float* __restrict__ src, __half* __restrict__ hi,
                                                                    __half* __restrict__ lo, float* __restrict__ inv_scale,
                                                                    int64_t rows, int64_t k, int64_t kpad, int64_t ld, int64_t batch1,
                                                                    int64_t s0, int64_t s1) {
  __shared__ float tile[32][33];
  __shared__ uint32_t red[8][32];
  const int64_t batch = blockIdx.y;
  const float* sb = src + (batch / batch1) * s0 + (batch % batch1) * s1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  uint32_t m = 0;
  if (r0 + tx < rows)
    for (int64_t kk = ty; kk < k; kk += 8) {
      const uint32_t a = h16_absbits(sb[kk * ld + r0 + tx]);
      m = m > a ? m : a;
    }
  red[ty][tx] = m;
  __syncthreads();
  if (ty == 0) {
#pragma unroll
    for (int i = 1; i < 8; ++i) m = red[i][tx] > m ? red[i][tx] : m;
    red[0][tx] = m;
    float scale, inv;
    h16_scales(m, scale, inv);
    if (r0 + tx < rows) inv_scale[batch * rows + r0 + tx] = inv;
  }
  __syncthreads();
  for (int64_t k0 = 0; k0 < kpad; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {          // read (k, r) with r contiguous
      const int64_t kk = k0 + ty + 8 * i, r = r0 + tx;
      tile[ty + 8 * i][tx] = (kk < k && r < rows) ? sb[kk * ld + r] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {          // write (r, k) with k contiguous
      const int64_t r = r0 + ty + 8 * i, kk = k0 + tx;
      if (r < rows && kk < kpad) {
        float scale, inv;
        h16_scales(red[0][ty + 8 * i], scale, inv);
        __half h, l;
        h16_split(tile[tx][ty + 8 * i], scale, h, l);
        const int64_t o = (batch * rows + r) * kpad + kk;
        hi[o] = h;
        lo[o] = l;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) h16_splitk_reduce_kernel(const float* __restrict__ partial, int splits, H16Params p) {
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
    *dst = h16_act(v, p.act);
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
// B200AD_H16_PROBE (diagnostic, read once): 1 = no output stores, 4 = skip the operand split (times the GEMM kernel alone on
// whatever the workspace holds), 8 = print the MMA warp's wait cycles, 16 = no in-kernel conversion of A (plane path)
static int h16_probe() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200AD_H16_PROBE");
    v = e ? atoi(e) : 0;
  }
  return v;
}

struct H16Plan {
  int64_t kpad, a_bytes, b_bytes, sa_bytes, sb_bytes, part_bytes;
  int m_tiles, n_tiles, batches, kblocks, splits, kblocks_per_split;
  int64_t total() const { return a_bytes + b_bytes + 2 * sa_bytes + 2 * sb_bytes + part_bytes; }
};
static inline int64_t round1k(int64_t x) { return ceil_div(x, 1024) * 1024; }

static H16Plan h16_plan(const GemmArgs& g) {
  H16Plan pl;
  pl.kpad = ceil_div(g.k, 8) * 8;                              // TMA row stride must be a multiple of 16 B
  pl.batches = (int)(g.batch0 * g.batch1);
  pl.a_bytes = round1k(2 * pl.batches * g.m * pl.kpad * 2);    // hi + lo planes
  pl.b_bytes = round1k(2 * pl.batches * g.n * pl.kpad * 2);
  pl.sa_bytes = round1k(pl.batches * g.m * 4);                 // inverse scales; the same again for the maxima
  pl.sb_bytes = round1k(pl.batches * g.n * 4);
  pl.m_tiles = (int)ceil_div(g.m, 256);
  pl.n_tiles = (int)ceil_div(g.n, 256);
  pl.kblocks = (int)ceil_div(g.k, H_BK);
  const int64_t tiles = (int64_t)pl.m_tiles * pl.n_tiles * pl.batches;
  const int units = num_sms() / 2;
  int splits = 1;
  if (tiles < units && pl.kblocks >= 16) {
    splits = (int)(units / tiles);
    if (splits > pl.kblocks / 8) splits = pl.kblocks / 8;      // >= 512 K-elements per split
    if (splits < 1) splits = 1;
    if (splits > 64) splits = 64;
  }
  pl.kblocks_per_split = (int)ceil_div(ceil_div(pl.kblocks, splits), H_CHUNK_KB) * H_CHUNK_KB;   // whole chains per split
  pl.splits = (int)ceil_div(pl.kblocks, pl.kblocks_per_split);
  pl.part_bytes = pl.splits > 1 ? (int64_t)pl.splits * pl.batches * g.m * g.n * 4 : 0;
  return pl;
}

bool gemm_h16_supported(const GemmArgs& g, const char** why) {
  auto no = [&](const char* w) { if (why) *why = w; return false; };
  if (g.m < 1 || g.n < 1 || g.k < 1) return no("empty problem");
  if (g.m >= (1ll << 31) || g.n >= (1ll << 31) || g.k >= (1ll << 31)) return no("dimension too large");
  if (g.m > 32 * 65535ll || g.n > 32 * 65535ll || g.k > 256 * 65535ll) return no("larger than the transposing pre-pass grids");
  if (g.batch0 * g.batch1 >= 65536) return no("too many batches");
  if (g.batch0 > 1 && (g.a_stride0 < 0 || g.b_stride0 < 0)) return no("negative batch stride");
  if (get_encode_fn() == nullptr) return no("cuTensorMapEncodeTiled unavailable (no driver)");
  return true;
}

int64_t gemm_h16_workspace_bytes(const GemmArgs& g) {
  if (g.m < 1 || g.n < 1 || g.k < 1) return 0;
  return h16_plan(g).total();
}

static int encode_h16_planes(CUtensorMap* map, const void* base, int64_t rows, int64_t kpad, int64_t batches) {
  EncodeTiledFn enc = get_encode_fn();
  cuuint64_t dims[4] = {(cuuint64_t)kpad, (cuuint64_t)rows, (cuuint64_t)batches, 2};
  cuuint64_t strides[3] = {(cuuint64_t)kpad * 2, (cuuint64_t)rows * kpad * 2, (cuuint64_t)batches * rows * kpad * 2};
  cuuint32_t box[4] = {(cuuint32_t)H_BK, 128u, 1u, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(B200AD_ERR_CUDA, "gemm(f16x3): cuTensorMapEncodeTiled failed (%d)", (int)r);
  return B200AD_OK;
}

// One operand -> planes + inverse scales.  `k_contig`: element (r, k) at r*ld + k (else at k*ld + r).
// `amax` is scratch of rows * batches words (only touched on the transposing path).
static int h16_prepare(const float* src, int64_t rows, int64_t k, int64_t kpad, int64_t ld, bool k_contig, int batches,
                       int64_t batch1, int64_t s0, int64_t s1, __half* planes, float* inv_scale, uint32_t* amax, cudaStream_t s) {
  __half* hi = planes;
  __half* lo = planes + (int64_t)batches * rows * kpad;
  if (h16_probe() & 4) return B200AD_OK;
  if (k_contig) {
    const bool vec = (k % 8 == 0) && aligned16(src) && (ld % 4 == 0) && (batches == 1 || (s0 % 4 == 0 && s1 % 4 == 0));
    dim3 grid((unsigned)ceil_div(rows, 8), (unsigned)batches);
    if (vec && k > 1024 && k <= 8192)
      h16_split_rows_block_kernel<<<dim3((unsigned)rows, (unsigned)batches), 256, 0, s>>>(src, hi, lo, inv_scale, rows, k, ld, batch1, s0, s1);
    else if (vec) h16_split_rows_kernel<true><<<grid, 256, 0, s>>>(src, hi, lo, inv_scale, rows, k, kpad, ld, batch1, s0, s1);
    else h16_split_rows_kernel<false><<<grid, 256, 0, s>>>(src, hi, lo, inv_scale, rows, k, kpad, ld, batch1, s0, s1);
    count_launch();
  } else if (rows * k <= (1 << 20) && k <= 4096) {
    dim3 grid((unsigned)ceil_div(rows, 32), (unsigned)batches);
    h16_split_trans_small_kernel<<<grid, 256, 0, s>>>(src, hi, lo, inv_scale, rows, k, kpad, ld, batch1, s0, s1);
    count_launch();
  } else {
    cudaError_t e = cudaMemsetAsync(amax, 0, (size_t)batches * rows * 4, s);
    if (e != cudaSuccess) return set_error(B200AD_ERR_CUDA, "gemm(f16x3): memset: %s", cudaGetErrorString(e));
    dim3 g1((unsigned)ceil_div(rows, 32), (unsigned)ceil_div(k, 256), (unsigned)batches);
    h16_colmax_kernel<<<g1, 256, 0, s>>>(src, amax, rows, k, ld, batch1, s0, s1);
    dim3 g2((unsigned)ceil_div(kpad, 32), (unsigned)ceil_div(rows, 32), (unsigned)batches);
    h16_split_trans_kernel<<<g2, 256, 0, s>>>(src, hi, lo, amax, inv_scale, rows, k, kpad, ld, batch1, s0, s1);
    count_launch(2);
  }
  return check_launch("gemm(f16x3) operand split");
}

// op(A) row-major, one matrix, rows 16 B aligned: the kernel converts the raw tiles itself (A_RAW), the pre-pass only finds the
// row scales.  B200AD_H16_PROBE bit 16 forces the plane path (A/B measurements).
static bool h16_a_raw_ok(const GemmArgs& g) {
  // (each of the N / 256 column tiles of a row panel converts the panel again: beyond N = 2048 the one-time plane split is cheaper)
  return !g.trans_a && g.batch0 * g.batch1 == 1 && (g.lda % 4) == 0 && aligned16(g.a) && g.n <= 2048 && !(h16_probe() & 16);
}
static int h16_prepare_raw_a(const GemmArgs& g, float* scale_a, float* inv_a, cudaStream_t s) {
  if (h16_probe() & 4) return B200AD_OK;
  h16_rowscale_kernel<<<(unsigned)ceil_div(g.m, 8), 256, 0, s>>>(g.a, scale_a, inv_a, g.m, g.k, g.lda);
  count_launch();
  return check_launch("gemm(f16x3) row scales");
}

// GEMM proper on prepared planes (a_planes == nullptr: A_RAW, the raw fp32 A of `g` + scale_a)
static int h16_run(const GemmArgs& g, H16Plan pl, const __half* a_planes, const float* scale_a, const float* inv_a,
                   const __half* b_planes, const float* inv_b, float* part, cudaStream_t s) {
  CUtensorMap map_a, map_b, map_c;
  int rc;
  const bool a_raw = a_planes == nullptr;
  if (a_raw) {
    EncodeTiledFn enc = get_encode_fn();
    cuuint64_t dims[4] = {(cuuint64_t)g.k, (cuuint64_t)g.m, 1, 1};
    cuuint64_t strides[3] = {(cuuint64_t)g.lda * 4, (cuuint64_t)g.lda * 4 * (cuuint64_t)g.m, (cuuint64_t)g.lda * 4 * (cuuint64_t)g.m};
    cuuint32_t box[4] = {32u, 128u, 1u, 1u};
    cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    CUresult r = enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(g.a), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(B200AD_ERR_CUDA, "gemm(f16x3): cuTensorMapEncodeTiled (raw A) failed (%d)", (int)r);
  } else if ((rc = encode_h16_planes(&map_a, a_planes, g.m, pl.kpad, pl.batches))) return rc;
  if ((rc = encode_h16_planes(&map_b, b_planes, g.n, pl.kpad, pl.batches))) return rc;
  if (!part) {
    pl.splits = 1;
    pl.kblocks_per_split = pl.kblocks;
  }
  H16Params p{};
  p.c = g.c; p.bias = g.bias; p.partial = part; p.scale_a = scale_a; p.inv_a = inv_a; p.inv_b = inv_b;
  p.m = g.m; p.n = g.n; p.k = g.k;
  p.ldc = g.ldc; p.c_stride0 = g.c_stride0; p.c_stride1 = g.c_stride1;
  p.batch1 = g.batch1;                   // C addressing keeps the caller's two batch levels: batch = b0 * batch1 + b1
  p.batches = pl.batches; p.m_tiles = pl.m_tiles; p.n_tiles = pl.n_tiles; p.splits = pl.splits;
  p.kblocks_total = pl.kblocks; p.kblocks_per_split = pl.kblocks_per_split;
  p.act = g.act; p.alpha = g.alpha; p.beta = g.beta; p.probe = h16_probe();
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
      cuuint32_t box[4] = {16u, 32u, 1u, 1u};
      cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
      CUresult r = enc(&map_c, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(cbase), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r == CUDA_SUCCESS) p.tma_store = 1;
    }
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_h16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_h16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM);
    if (e != cudaSuccess) return set_error(B200AD_ERR_CUDA, "gemm(f16x3): smem attribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  const int64_t total_tiles = (int64_t)pl.batches * pl.splits * pl.m_tiles * pl.n_tiles;
  const int pairs = num_sms() / 2;
  const int grid = 2 * (int)(total_tiles < pairs ? total_tiles : pairs);
  if (a_raw) gemm_h16_kernel<true><<<grid, H_THREADS, H_SMEM, s>>>(map_a, map_b, map_c, p);
  else gemm_h16_kernel<false><<<grid, H_THREADS, H_SMEM, s>>>(map_a, map_b, map_c, p);
  count_launch();
  rc = check_launch("gemm(f16x3)");
  if (rc) return rc;
  if (pl.splits > 1) {
    const int64_t total = g.m * g.n * pl.batches;
    h16_splitk_reduce_kernel<<<stream_grid(total, 256), 256, 0, s>>>(part, pl.splits, p);
    count_launch();
    rc = check_launch("gemm(f16x3 split-K reduce)");
  }
  return rc;
}

struct H16Workspace {
  __half *a_planes, *b_planes;
  float *inv_a, *inv_b;
  uint32_t *amax_a, *amax_b;
  float* part;
};
static H16Workspace h16_carve(const H16Plan& pl, void* ws, int64_t ws_bytes) {
  uint8_t* p = reinterpret_cast<uint8_t*>(ws);
  H16Workspace w;
  w.a_planes = reinterpret_cast<__half*>(p); p += pl.a_bytes;
  w.b_planes = reinterpret_cast<__half*>(p); p += pl.b_bytes;
  w.inv_a = reinterpret_cast<float*>(p); p += pl.sa_bytes;
  w.inv_b = reinterpret_cast<float*>(p); p += pl.sb_bytes;
  w.amax_a = reinterpret_cast<uint32_t*>(p); p += pl.sa_bytes;
  w.amax_b = reinterpret_cast<uint32_t*>(p); p += pl.sb_bytes;
  w.part = (pl.part_bytes > 0 && ws_bytes >= pl.total()) ? reinterpret_cast<float*>(p) : nullptr;   // no room: run un-split
  return w;
}

int gemm_h16_launch(const GemmArgs& g, void* ws, int64_t ws_bytes, cudaStream_t s) {
  const H16Plan pl = h16_plan(g);
  if (!ws || ws_bytes < pl.total() - pl.part_bytes)
    return set_error(B200AD_ERR_INVALID, "gemm(f16x3): workspace of %lld bytes required", (long long)pl.total());
  const H16Workspace w = h16_carve(pl, ws, ws_bytes);
  int rc;
  // A: (m x k), row-major unless trans_a; B: (k x n) row-major, i.e. K-contiguous only when trans_b
  const bool a_raw = h16_a_raw_ok(g);
  float* scale_a = reinterpret_cast<float*>(w.amax_a);         // the maxima scratch doubles as the forward scales of a raw A
  if (a_raw) rc = h16_prepare_raw_a(g, scale_a, w.inv_a, s);
  else rc = h16_prepare(g.a, g.m, g.k, pl.kpad, g.lda, !g.trans_a, pl.batches, g.batch1, g.a_stride0, g.a_stride1, w.a_planes,
                        w.inv_a, w.amax_a, s);
  if (rc) return rc;
  if ((rc = h16_prepare(g.b, g.n, g.k, pl.kpad, g.ldb, g.trans_b != 0, pl.batches, g.batch1, g.b_stride0, g.b_stride1, w.b_planes,
                        w.inv_b, w.amax_b, s)))
    return rc;
  return h16_run(g, pl, a_raw ? nullptr : w.a_planes, scale_a, w.inv_a, w.b_planes, w.inv_b, w.part, s);
}

// One A operand, `count` (B, C, bias) triples: A is split once.  Workspace: gemm_h16_workspace_bytes of the descriptor.
int gemm_h16_shared_a_launch(const GemmArgs& g0, int count, const float* const* bs, float* const* cs, const float* const* biases,
                             void* ws, int64_t ws_bytes, cudaStream_t s) {
  const H16Plan pl = h16_plan(g0);
  if (!ws || ws_bytes < pl.total() - pl.part_bytes)
    return set_error(B200AD_ERR_INVALID, "gemm_shared_a(f16x3): workspace of %lld bytes required", (long long)pl.total());
  const H16Workspace w = h16_carve(pl, ws, ws_bytes);
  int rc;
  const bool a_raw = h16_a_raw_ok(g0);
  float* scale_a = reinterpret_cast<float*>(w.amax_a);
  if (a_raw) rc = h16_prepare_raw_a(g0, scale_a, w.inv_a, s);
  else rc = h16_prepare(g0.a, g0.m, g0.k, pl.kpad, g0.lda, !g0.trans_a, pl.batches, g0.batch1, g0.a_stride0, g0.a_stride1,
                        w.a_planes, w.inv_a, w.amax_a, s);
  if (rc) return rc;
  for (int i = 0; i < count; ++i) {
    GemmArgs g = g0;
    g.b = bs[i]; g.c = cs[i]; g.bias = biases ? biases[i] : nullptr;
    if ((rc = h16_prepare(g.b, g.n, g.k, pl.kpad, g.ldb, g.trans_b != 0, pl.batches, g.batch1, g.b_stride0, g.b_stride1,
                          w.b_planes, w.inv_b, w.amax_b, s)))
      return rc;
    if ((rc = h16_run(g, pl, a_raw ? nullptr : w.a_planes, scale_a, w.inv_a, w.b_planes, w.inv_b, w.part, s))) return rc;
  }
  return B200AD_OK;
}

}  // namespace b200ad
