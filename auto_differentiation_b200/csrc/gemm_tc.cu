// tcgen05 / TMEM / TMA GEMM for sm_100a — the dense-contraction hot kernel.
//
//   C[b0,b1] = alpha * op(A) @ op(B) + beta * C (+ bias[n], activation), fp32 in HBM, fp32 accumulate in TMEM.
//
// Precisions
//   TF32    one tcgen05.mma.kind::tf32 pass (10-bit mantissa operands)        — diagnostic / roofline probe
//   TF32X3  error-compensated split: A = Ah + Al, B = Bh + Bl (each exactly representable in tf32),
//           D = Ah*Bh + Ah*Bl + Al*Bh  (the dropped Al*Bl term is ~2^-22 relative) — fp32-faithful, the default
// The hi/lo split is done IN SHARED MEMORY by four "splitter" warps on the tile TMA just delivered (an
// elementwise rewrite, so it is independent of the swizzled layout): no extra HBM traffic, no pre-pass.
//
// Structure (persistent, warp-specialised, one CTA per SM, static tile scheduler):
//   warp 0      TMA producer: cp.async.bulk.tensor (4-D maps: inner, rows, batch1, batch0; 128B swizzle)
//   warp 1      MMA issuer (one elected lane): tcgen05.mma 128 x BN x 8, accumulators in TMEM (double buffered)
//   warps 2-9   splitter (TF32X3 only): raw tile -> masked hi (in place) + lo tile, fence.proxy.async
//   warps 10-13 epilogue: tcgen05.ld TMEM -> registers -> alpha/beta/bias/activation -> global
// Pipelines: raw_full / conv_full / empty mbarriers over a ring of smem stages, tmem_full / tmem_empty over
// the two accumulator buffers, so the epilogue of tile i overlaps the main loop of tile i+1.
// Both operand majors are fed natively: a row-major (rows x K) operand is K-major, a (K x cols) row-major
// operand is MN-major (UMMA descriptor major bit), so NN / NT / TN / TT need no transposes.
// Skinny problems (weight gradients: M,N ~ 512, K ~ 1e5) use split-K with a deterministic two-pass reduce.
#include <atomic>
#include <mutex>
#include "common.cuh"
#include "gemm_internal.cuh"
#include "tcgen05.cuh"
#include <cuda.h>
#include <mutex>

namespace b200ad {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                 // 32 fp32 = 128 B rows (K-major: swizzle-128B).  BK=16 (swizzle-64B, 4-deep
                                          // ring) was measured slower: smaller TMA boxes, same smem-bandwidth bound
constexpr int TC_SPLIT_WARPS = 8;
constexpr int TC_SPLIT_THREADS = TC_SPLIT_WARPS * 32;
constexpr int TC_THREADS = 64 + TC_SPLIT_THREADS + 128;   // TMA, MMA, 8 splitter warps, 4 epilogue warps
constexpr uint32_t TC_SMEM_LIMIT = 232448;  // 227 KB


// ---------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------
struct TcParams {
  float* c;
  const float* c_in;       // beta source (== c)
  const float* bias;
  float* partial;          // split-K workspace [split][batch][M][N] or nullptr
  int64_t m, n, k;
  int64_t ldc, c_stride0, c_stride1;
  int64_t batch1;
  int32_t batches, m_tiles, n_tiles, splits, kblocks_total, kblocks_per_split;
  int32_t a_mn_major, b_mn_major, x3, mask_hi, act;
  int32_t a_tmem;              // BN = 64, TF32X3: the splitter stores the hi/lo A tiles into TMEM (TS-mode MMAs)
  int32_t probe;               // diagnostic bits: 1 = skip the TMA loads, 2 = skip the split (3 = pure MMA + epilogue), 4 = skip the stores
  int32_t tma_store;           // gemm_tc3_kernel: the output tile leaves through shared memory + cp.async.bulk.tensor stores
  int32_t chunk_kb, chunked;   // TMEM accumulation chunk (k-blocks); chunked: buffer 0 = chunk sum X, buffer 1 = running total Y
  float alpha, beta;
  int32_t pdl;                 // launched as a programmatic dependent: set up first, then wait for the predecessor's data
};

template <int ACT>
__device__ __forceinline__ float tc_act(float x) {
  if constexpr (ACT == 1) return tanhf(x);
  else if constexpr (ACT == 2) return 1.f / (1.f + expf(-x));
  else if constexpr (ACT == 3) return (x > 0.f || x != x) ? x : 0.f;
  else return x;
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const TcParams p, const int stages) {
  constexpr uint32_t A_BYTES = TC_BM * TC_BK * 4;          // 16 KB
  constexpr uint32_t B_BYTES = BN * TC_BK * 4;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem base is only guaranteed 16B aligned by the ABI: round up to 1024 (swizzle-128B atoms)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_bytes = (A_BYTES + B_BYTES) * (p.x3 ? 2 : 1);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* raw_full = bars;                  // [stages]  TMA -> splitter (or MMA)
  uint64_t* conv_full = bars + stages;        // [stages]  splitter -> MMA
  uint64_t* empty = bars + 2 * stages;        // [stages]  MMA -> TMA
  uint64_t* tmem_full = bars + 3 * stages;    // [2]       MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;       // [2]       epilogue -> MMA
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (p.pdl) pdl_launch_dependents();       // the successor may begin ITS set-up; it waits for this grid's completion itself
  // TMEM: two accumulator buffers; when the chunk accumulators can ping-pong (BN <= 128) a third region holds the
  // running total Y, so the fold of chunk c overlaps the MMAs of chunk c+1.  BN = 256 has room for X and Y only.
  constexpr bool PINGPONG = (BN <= 128);
  constexpr uint32_t Y_COL = PINGPONG ? 2 * BN : BN;
  // BN = 64 additionally keeps a ring of A tiles in TMEM (hi: 32 columns, lo: 32 columns per stage, <= 4 stages):
  // every tf32 tcgen05.mma pays ~128 cycles for fetching its 4 KB A operand from shared memory regardless of N
  // (measured), which caps N = 64 tiles at 25 % of the tensor rate; with A in TMEM the MMA takes N/2 cycles.
  constexpr uint32_t A_TMEM_COL = 3 * BN;                       // after X0, X1, Y
  constexpr uint32_t TMEM_COLS = PINGPONG ? (BN == 64 ? 512 : 512) : 2 * BN;
  const bool a_tmem = (BN == 64) && p.a_tmem && p.x3 && stages <= 4;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&conv_full[s], TC_SPLIT_THREADS);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  // barriers, tensor memory and descriptors are in place: from here on the kernel reads its operands and writes its output
  if (p.pdl) pdl_wait();

  const int total_tiles = p.batches * p.splits * p.m_tiles * p.n_tiles;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int t = tile;
        const int n_tile = t % p.n_tiles; t /= p.n_tiles;
        const int m_tile = t % p.m_tiles; t /= p.m_tiles;
        const int split = t % p.splits;
        const int batch = t / p.splits;
        const int b0 = (int)(batch / p.batch1), b1 = (int)(batch % p.batch1);
        const int kb0 = split * p.kblocks_per_split;
        int kb1 = kb0 + p.kblocks_per_split;
        if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % stages;
          const uint32_t ph = (it / stages) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          uint8_t* sb = sa + A_BYTES;
          if (p.probe & 1) { mbar_arrive(&raw_full[s]); continue; }
          mbar_expect_tx(&raw_full[s], A_BYTES + B_BYTES);
          const int k0 = kb * TC_BK;
          if (!p.a_mn_major) {
            tma_load_4d(&map_a, &raw_full[s], sa, k0, m_tile * TC_BM, b1, b0);
          } else {
#pragma unroll
            for (int c = 0; c < TC_BM / 32; ++c)
              tma_load_4d(&map_a, &raw_full[s], sa + c * (128 * TC_BK), m_tile * TC_BM + c * 32, k0, b1, b0);
          }
          if (!p.b_mn_major) {
            tma_load_4d(&map_b, &raw_full[s], sb, k0, n_tile * BN, b1, b0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 32; ++c)
              tma_load_4d(&map_b, &raw_full[s], sb + c * (128 * TC_BK), n_tile * BN + c * 32, k0, b1, b0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    // instruction descriptor: D=f32, A=B=tf32, majors, N>>3, M>>4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn_major << 15) |
                           ((uint32_t)p.b_mn_major << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    // K-major (SWIZZLE_128B for BK=32 / SWIZZLE_64B for BK=16): rows of 4*BK B, 8-row groups 32*BK B apart (SBO),
    // k-step = +32 B inside the swizzle atom.
    // MN-major (SWIZZLE_128B_BASE32B): 32-element chunks 128*BK B apart (LBO), 4-k-row groups 512 B apart
    // (SBO), k-step (8 rows) = +1024 B.
    constexpr uint32_t MN_LBO = 128u * TC_BK;
    const uint32_t a_lbo = p.a_mn_major ? MN_LBO : 16u, b_lbo = p.b_mn_major ? MN_LBO : 16u;
    constexpr uint32_t K_SBO = 32u * TC_BK, K_LAY = (TC_BK == 32) ? 2u : 4u;
    const uint32_t a_sbo = p.a_mn_major ? 512u : K_SBO, b_sbo = p.b_mn_major ? 512u : K_SBO;
    const uint32_t a_lay = p.a_mn_major ? 1u : K_LAY, b_lay = p.b_mn_major ? 1u : K_LAY;
    const uint32_t a_kstep = p.a_mn_major ? 1024u : 32u, b_kstep = p.b_mn_major ? 1024u : 32u;
    uint32_t it = 0, uit = 0;   // k-block counter (smem ring) and accumulation-unit counter (TMEM buffers)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int t = tile / (p.n_tiles * p.m_tiles);
      const int split = t % p.splits;
      const int kb0 = split * p.kblocks_per_split;
      int kb1 = kb0 + p.kblocks_per_split;
      if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
      for (int cb0 = kb0; cb0 < kb1; cb0 += p.chunk_kb, ++uit) {
        int cb1 = cb0 + p.chunk_kb;
        if (cb1 > kb1) cb1 = kb1;
        const bool single = p.chunked && !PINGPONG;        // BN = 256 chunked: X is buffer 0 only
        const uint32_t buf = single ? 0u : (uit & 1);
        const uint32_t buf_ph = single ? (uit & 1) : ((uit >> 1) & 1);
        mbar_wait(&tmem_empty[buf], buf_ph ^ 1);          // epilogue has drained / folded this buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = cb0; kb < cb1; ++kb, ++it) {
          const int s = it % stages;
          const uint32_t ph = (it / stages) & 1;
          mbar_wait(p.x3 ? &conv_full[s] : &raw_full[s], ph);
          tc_fence_after();
          if (elect_one()) {   // always the same thread: tcgen05.commit tracks the MMAs issued by the committing thread
            const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
            const uint32_t sb = sa + A_BYTES;
            const uint32_t sal = sb + B_BYTES, sbl = sal + A_BYTES;     // lo tiles (TF32X3)
            const bool first_kb = (kb == cb0);
            if (a_tmem) {
              const uint32_t idesc_ts = idesc & ~(1u << 15);          // A from TMEM is always K-major
              const uint32_t ah = tmem_base + A_TMEM_COL + (uint32_t)s * 64u, al = ah + 32u;
#pragma unroll
              for (int kk = 0; kk < TC_BK / 8; ++kk)
                umma_tf32_ts(d_tmem, al + kk * 8, make_smem_desc(sb + kk * b_kstep, b_lbo, b_sbo, b_lay), idesc_ts,
                             (first_kb && kk == 0) ? 0u : 1u);
#pragma unroll
              for (int kk = 0; kk < TC_BK / 8; ++kk)
                umma_tf32_ts(d_tmem, ah + kk * 8, make_smem_desc(sbl + kk * b_kstep, b_lbo, b_sbo, b_lay), idesc_ts, 1u);
#pragma unroll
              for (int kk = 0; kk < TC_BK / 8; ++kk)
                umma_tf32_ts(d_tmem, ah + kk * 8, make_smem_desc(sb + kk * b_kstep, b_lbo, b_sbo, b_lay), idesc_ts, 1u);
            } else if (p.x3) {
              // small cross terms first, the dominant hi*hi term last
#pragma unroll
              for (int kk = 0; kk < TC_BK / 8; ++kk)
                umma_tf32(d_tmem, make_smem_desc(sal + kk * a_kstep, a_lbo, a_sbo, a_lay),
                          make_smem_desc(sb + kk * b_kstep, b_lbo, b_sbo, b_lay), idesc, (first_kb && kk == 0) ? 0u : 1u);
#pragma unroll
              for (int kk = 0; kk < TC_BK / 8; ++kk)
                umma_tf32(d_tmem, make_smem_desc(sa + kk * a_kstep, a_lbo, a_sbo, a_lay),
                          make_smem_desc(sbl + kk * b_kstep, b_lbo, b_sbo, b_lay), idesc, 1u);
#pragma unroll
              for (int kk = 0; kk < TC_BK / 8; ++kk)
                umma_tf32(d_tmem, make_smem_desc(sa + kk * a_kstep, a_lbo, a_sbo, a_lay),
                          make_smem_desc(sb + kk * b_kstep, b_lbo, b_sbo, b_lay), idesc, 1u);
            } else {
#pragma unroll
              for (int kk = 0; kk < TC_BK / 8; ++kk)
                umma_tf32(d_tmem, make_smem_desc(sa + kk * a_kstep, a_lbo, a_sbo, a_lay),
                          make_smem_desc(sb + kk * b_kstep, b_lbo, b_sbo, b_lay), idesc, (first_kb && kk == 0) ? 0u : 1u);
            }
            umma_commit(&empty[s]);                          // smem stage free once these MMAs retire
            if (kb == cb1 - 1) umma_commit(&tmem_full[buf]);  // chunk / tile accumulator complete -> epilogue
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 2 + TC_SPLIT_WARPS) {
    // ===================================== splitter (TF32X3) =====================================
    if (p.x3) {
      const int tid = threadIdx.x - 64;   // 0..TC_SPLIT_THREADS-1
      uint32_t it = 0;
      constexpr uint32_t VECS = (A_BYTES + B_BYTES) / 16;   // float4 count of the raw [A|B] block
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int t = tile / (p.n_tiles * p.m_tiles);
        const int split = t % p.splits;
        const int kb0 = split * p.kblocks_per_split;
        int kb1 = kb0 + p.kblocks_per_split;
        if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % stages;
          const uint32_t ph = (it / stages) & 1;
          mbar_wait(&raw_full[s], ph);
          uint4* hi = reinterpret_cast<uint4*>(smem + (size_t)s * stage_bytes);
          uint4* lo = hi + VECS;
          if (p.probe & 2) { mbar_arrive(&conv_full[s]); continue; }
          if (a_tmem) {
            // A: each thread owns one row (TMEM lane) and stores its 32 k-values as hi (warps 2-5) or lo (6-9)
            const int q = warp & 3, half = (warp - 2) >> 2;
            const int r = q * 32 + lane;
            const uint8_t* sa = smem + (size_t)s * stage_bytes;
            uint32_t v[32];
            if (!p.a_mn_major) {          // K-major, SWIZZLE_128B: 16 B chunk c of row r sits at chunk c ^ (r & 7)
              const uint4* row = reinterpret_cast<const uint4*>(sa + r * 128);
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint4 t4 = row[c ^ (r & 7)];
                v[4 * c] = t4.x; v[4 * c + 1] = t4.y; v[4 * c + 2] = t4.z; v[4 * c + 3] = t4.w;
              }
            } else {                      // MN-major, SWIZZLE_128B_BASE32B: 32 B granule g of k-row kk sits at g ^ (kk & 3)
              const uint8_t* chunk = sa + (r >> 5) * (128 * TC_BK);
              const int ml = r & 31;
#pragma unroll
              for (int kk = 0; kk < 32; ++kk)
                v[kk] = *reinterpret_cast<const uint32_t*>(chunk + kk * 128 + ((((ml >> 3) ^ (kk & 3)) << 5) | ((ml & 7) << 2)));
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const uint32_t h = v[j] & 0xffffe000u;
              v[j] = half ? (__float_as_uint(__uint_as_float(v[j]) - __uint_as_float(h)) & 0xffffe000u) : h;
            }
            tmem_st_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + A_TMEM_COL + (uint32_t)s * 64u + (half ? 32u : 0u), v);
            // B: lo tile in shared memory, as before
            constexpr uint32_t A_VECS = A_BYTES / 16;
            for (uint32_t i = A_VECS + tid; i < VECS; i += TC_SPLIT_THREADS) {
              const uint4 bv = hi[i];
              uint4 l;
              l.x = __float_as_uint(__uint_as_float(bv.x) - __uint_as_float(bv.x & 0xffffe000u)) & 0xffffe000u;
              l.y = __float_as_uint(__uint_as_float(bv.y) - __uint_as_float(bv.y & 0xffffe000u)) & 0xffffe000u;
              l.z = __float_as_uint(__uint_as_float(bv.z) - __uint_as_float(bv.z & 0xffffe000u)) & 0xffffe000u;
              l.w = __float_as_uint(__uint_as_float(bv.w) - __uint_as_float(bv.w & 0xffffe000u)) & 0xffffe000u;
              lo[i] = l;
            }
            tmem_st_wait();
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(&conv_full[s]);
            continue;
          }
#pragma unroll 6
          for (uint32_t i = tid; i < VECS; i += TC_SPLIT_THREADS) {
            const uint4 v = hi[i];
            uint4 h, l;
            h.x = v.x & 0xffffe000u; h.y = v.y & 0xffffe000u; h.z = v.z & 0xffffe000u; h.w = v.w & 0xffffe000u;
            l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x)) & 0xffffe000u;
            l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y)) & 0xffffe000u;
            l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z)) & 0xffffe000u;
            l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w)) & 0xffffe000u;
            if (p.mask_hi) hi[i] = h;
            lo[i] = l;
          }
          fence_proxy_async();            // generic-proxy writes -> visible to the tensor core (async proxy)
          mbar_arrive(&conv_full[s]);
        }
      }
    }
  } else {
    // ===================================== epilogue =====================================
    const int ew = warp & 3;                       // TMEM lane quarter this warp may access
    const int row_in_tile = ew * 32 + lane;
    const uint32_t lane_base = (uint32_t)(ew * 32) << 16;
    uint32_t uit = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int t = tile;
      const int n_tile = t % p.n_tiles; t /= p.n_tiles;
      const int m_tile = t % p.m_tiles; t /= p.m_tiles;
      const int split = t % p.splits;
      const int batch = t / p.splits;
      const int64_t b0 = batch / p.batch1, b1 = batch % p.batch1;
      const int kb0 = split * p.kblocks_per_split;
      int kb1 = kb0 + p.kblocks_per_split;
      if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
      const int nchunks = (kb1 - kb0 + p.chunk_kb - 1) / p.chunk_kb;
      const int64_t gm = (int64_t)m_tile * TC_BM + row_in_tile;
      const bool row_ok = gm < p.m;
      float* crow;
      if (p.partial) crow = p.partial + (((int64_t)split * p.batches + batch) * p.m + gm) * p.n;
      else crow = p.c + b0 * p.c_stride0 + b1 * p.c_stride1 + gm * p.ldc;
      for (int ch = 0; ch < nchunks; ++ch, ++uit) {
        const bool single = p.chunked && !PINGPONG;
        const uint32_t buf = single ? 0u : (uit & 1);
        const uint32_t buf_ph = single ? (uit & 1) : ((uit >> 1) & 1);
        mbar_wait(&tmem_full[buf], buf_ph);
        tc_fence_after();
        const uint32_t x_addr = tmem_base + lane_base + buf * BN;
        const uint32_t y_addr = tmem_base + lane_base + Y_COL;       // running total (chunked mode only)
        const bool last = (ch == nchunks - 1);
        if (!last) {
          // fold the chunk sum into the running total with round-to-nearest adds (TMEM accumulation inside the
          // tensor core truncates; bounding the chain length keeps the result fp32-faithful for any K)
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t x[32];
            tmem_ld_32x32b_x32(x_addr + c0, x);
            if (ch > 0) {
              uint32_t y[32];
              tmem_ld_32x32b_x32(y_addr + c0, y);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] = __float_as_uint(__uint_as_float(x[j]) + __uint_as_float(y[j]));
            } else {
              tmem_ld_wait();
            }
            tmem_st_32x32b_x32(y_addr + c0, x);
          }
          tmem_st_wait();
        } else {
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(x_addr + c0, r);
            if (nchunks > 1) {
              uint32_t y[32];
              tmem_ld_32x32b_x32(y_addr + c0, y);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(y[j]));
            } else {
              tmem_ld_wait();
            }
            const int64_t gn0 = (int64_t)n_tile * BN + c0;
            if (row_ok && gn0 < p.n && !(p.probe & 4)) {
              float* dst = crow + gn0;
              const bool full = (gn0 + 32 <= p.n);
              if (p.partial) {
                if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                  for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                       __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                } else {
                  for (int j = 0; j < 32 && gn0 + j < p.n; ++j) dst[j] = __uint_as_float(r[j]);
                }
              } else {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = p.alpha * __uint_as_float(r[j]);
                if (p.beta != 0.f) {
                  for (int j = 0; j < 32 && gn0 + j < p.n; ++j) v[j] += p.beta * dst[j];
                }
                if (p.bias) {
                  for (int j = 0; j < 32 && gn0 + j < p.n; ++j) v[j] += __ldg(p.bias + gn0 + j);
                }
                if (p.act == 1) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = tc_act<1>(v[j]);
                } else if (p.act == 2) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = tc_act<2>(v[j]);
                } else if (p.act == 3) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = tc_act<3>(v[j]);
                }
                if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                  for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
                  for (int j = 0; j < 32 && gn0 + j < p.n; ++j) dst[j] = v[j];
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty[buf]);               // 128 arrivals release the buffer to the MMA warp
      }
    }
  }

  // teardown: every role has left its loop; the MMA warp's last commits were consumed by the epilogue
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}


// =========================================================================================================
// CTA-pair kernel (TF32X3 only): one 256 x 256 output tile per cluster of two CTAs (tcgen05 cta_group::2,
// UMMA M=256, N=256).  CTA r of the pair owns rows [128r, 128r+128) of the tile and stages columns
// [128r, 128r+128) of the B tile, so each SM loads / splits / serves half as much B as the 1-CTA kernel:
// 64 KB stages -> a 3-deep ring (the 1-CTA kernel only fits 2 stages and exposes TMA latency + split time),
// half the L2->SM operand traffic, and half the shared-memory operand reads per MMA.
// Roles per CTA are the same as above; only the leader CTA (rank 0) issues MMAs.  Cross-CTA signalling:
//   conv_full[s]   (leader)  <- 512 splitter threads of both CTAs (remote mbarrier.arrive via mapa)
//   empty[s]       (both)    <- tcgen05.commit ... multicast::cluster
//   tmem_full[b]   (both)    <- tcgen05.commit ... multicast::cluster
//   tmem_empty[b]  (leader)  <- 256 epilogue threads of both CTAs
// =========================================================================================================
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const TcParams p, const int stages) {
  constexpr int BN = 256;
  constexpr uint32_t A_BYTES = TC_BM * TC_BK * 4;            // this CTA's 128 rows of A
  constexpr uint32_t B_BYTES = (BN / 2) * TC_BK * 4;         // this CTA's half of the B tile
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // smem: a 4-deep ring of raw (= hi) tiles filled by TMA and a 2-deep ring of lo tiles filled by the splitter:
  // TMA gets three k-blocks of lead time while the lo tiles only need to stay one k-block ahead of the MMA.
  constexpr uint32_t raw_bytes = A_BYTES + B_BYTES;           // 32 KB
  constexpr int R = 4, LS = 2;
  (void)stages;
  uint8_t* lo_base = smem + (size_t)R * raw_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(lo_base + (size_t)LS * raw_bytes);
  uint64_t* raw_full = bars;                 // [R]  TMA -> splitter (local)
  uint64_t* empty = bars + R;                // [R]  MMA -> TMA (both CTAs, multicast commit)
  uint64_t* conv_full = bars + 2 * R;        // [LS] splitter -> MMA (leader, 512 arrivals)
  uint64_t* lo_empty = conv_full + LS;       // [LS] MMA -> splitter (both CTAs, multicast commit)
  uint64_t* tmem_full = lo_empty + LS;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  constexpr uint32_t TMEM_COLS = 2 * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < R; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < LS; ++s) {
      mbar_init(&conv_full[s], 2 * TC_SPLIT_THREADS);
      mbar_init(&lo_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_base_slot, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();                      // both CTAs' barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int total_tiles = p.batches * p.splits * p.m_tiles * p.n_tiles;   // m_tiles counts 256-row tiles
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) =====================================
    if (elect_one()) {
      uint32_t it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
        int t = tile;
        const int n_tile = t % p.n_tiles; t /= p.n_tiles;
        const int m_tile = t % p.m_tiles; t /= p.m_tiles;
        const int split = t % p.splits;
        const int batch = t / p.splits;
        const int b0 = (int)(batch / p.batch1), b1 = (int)(batch % p.batch1);
        const int kb0 = split * p.kblocks_per_split;
        int kb1 = kb0 + p.kblocks_per_split;
        if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
        const int m0 = m_tile * 256 + (int)rank * TC_BM;
        const int n0 = n_tile * BN + (int)rank * (BN / 2);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % R;
          const uint32_t ph = (it / R) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + (size_t)s * raw_bytes;
          uint8_t* sb = sa + A_BYTES;
          if (p.probe & 1) { mbar_arrive(&raw_full[s]); continue; }
          mbar_expect_tx(&raw_full[s], A_BYTES + B_BYTES);
          const int k0 = kb * TC_BK;
          if (!p.a_mn_major) {
            tma_load_4d(&map_a, &raw_full[s], sa, k0, m0, b1, b0);
          } else {
#pragma unroll
            for (int c = 0; c < TC_BM / 32; ++c)
              tma_load_4d(&map_a, &raw_full[s], sa + c * (128 * TC_BK), m0 + c * 32, k0, b1, b0);
          }
          if (!p.b_mn_major) {
            tma_load_4d(&map_b, &raw_full[s], sb, k0, n0, b1, b0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_4d(&map_b, &raw_full[s], sb + c * (128 * TC_BK), n0 + c * 32, k0, b1, b0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) =====================================
    if (rank == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn_major << 15) |
                             ((uint32_t)p.b_mn_major << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      constexpr uint32_t MN_LBO = 128u * TC_BK;
      const uint32_t a_lbo = p.a_mn_major ? MN_LBO : 16u, b_lbo = p.b_mn_major ? MN_LBO : 16u;
      constexpr uint32_t K_SBO = 32u * TC_BK, K_LAY = (TC_BK == 32) ? 2u : 4u;
      const uint32_t a_sbo = p.a_mn_major ? 512u : K_SBO, b_sbo = p.b_mn_major ? 512u : K_SBO;
      const uint32_t a_lay = p.a_mn_major ? 1u : K_LAY, b_lay = p.b_mn_major ? 1u : K_LAY;
      const uint32_t a_kstep = p.a_mn_major ? 1024u : 32u, b_kstep = p.b_mn_major ? 1024u : 32u;
      uint32_t it = 0, uit = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
        int t = tile / (p.n_tiles * p.m_tiles);
        const int split = t % p.splits;
        const int kb0 = split * p.kblocks_per_split;
        int kb1 = kb0 + p.kblocks_per_split;
        if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
        for (int cb0 = kb0; cb0 < kb1; cb0 += p.chunk_kb, ++uit) {
          int cb1 = cb0 + p.chunk_kb;
          if (cb1 > kb1) cb1 = kb1;
          const uint32_t buf = p.chunked ? 0u : (uit & 1);
          const uint32_t buf_ph = p.chunked ? (uit & 1) : ((uit >> 1) & 1);
          mbar_wait(&tmem_empty[buf], buf_ph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * BN;
          for (int kb = cb0; kb < cb1; ++kb, ++it) {
            const int s = it % R, ls = it % LS;
            const uint32_t lph = (it / LS) & 1;
            mbar_wait(&conv_full[ls], lph);        // both CTAs' hi/lo tiles of this k-block are in place
            tc_fence_after();
            if (elect_one()) {
              const uint32_t sa = smem_u32(smem + (size_t)s * raw_bytes);
              const uint32_t sb = sa + A_BYTES;
              const uint32_t sal = smem_u32(lo_base + (size_t)ls * raw_bytes), sbl = sal + A_BYTES;
              const bool first_kb = (kb == cb0);
#pragma unroll
              for (int kk = 0; kk < TC_BK / 8; ++kk)
                umma2_tf32(d_tmem, make_smem_desc(sal + kk * a_kstep, a_lbo, a_sbo, a_lay),
                           make_smem_desc(sb + kk * b_kstep, b_lbo, b_sbo, b_lay), idesc, (first_kb && kk == 0) ? 0u : 1u);
#pragma unroll
              for (int kk = 0; kk < TC_BK / 8; ++kk)
                umma2_tf32(d_tmem, make_smem_desc(sa + kk * a_kstep, a_lbo, a_sbo, a_lay),
                           make_smem_desc(sbl + kk * b_kstep, b_lbo, b_sbo, b_lay), idesc, 1u);
#pragma unroll
              for (int kk = 0; kk < TC_BK / 8; ++kk)
                umma2_tf32(d_tmem, make_smem_desc(sa + kk * a_kstep, a_lbo, a_sbo, a_lay),
                           make_smem_desc(sb + kk * b_kstep, b_lbo, b_sbo, b_lay), idesc, 1u);
              umma2_commit_mc(&empty[s]);                          // frees the raw stage in BOTH CTAs
              umma2_commit_mc(&lo_empty[ls]);                      // ... and the lo stage
              if (kb == cb1 - 1) umma2_commit_mc(&tmem_full[buf]);  // accumulator ready in BOTH CTAs
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp < 2 + TC_SPLIT_WARPS) {
    // ===================================== splitter (both CTAs) =====================================
    const int tid = threadIdx.x - 64;
    uint32_t it = 0;
    constexpr uint32_t VECS = (A_BYTES + B_BYTES) / 16;
    for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
      int t = tile / (p.n_tiles * p.m_tiles);
      const int split = t % p.splits;
      const int kb0 = split * p.kblocks_per_split;
      int kb1 = kb0 + p.kblocks_per_split;
      if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % R, ls = it % LS;
        mbar_wait(&raw_full[s], (it / R) & 1);
        mbar_wait(&lo_empty[ls], ((it / LS) & 1) ^ 1);
        uint4* hi = reinterpret_cast<uint4*>(smem + (size_t)s * raw_bytes);
        uint4* lo = reinterpret_cast<uint4*>(lo_base + (size_t)ls * raw_bytes);
        if (!(p.probe & 2)) {
#pragma unroll 4
          for (uint32_t i = tid; i < VECS; i += TC_SPLIT_THREADS) {
            const uint4 v = hi[i];
            uint4 h, l;
            h.x = v.x & 0xffffe000u; h.y = v.y & 0xffffe000u; h.z = v.z & 0xffffe000u; h.w = v.w & 0xffffe000u;
            l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x)) & 0xffffe000u;
            l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y)) & 0xffffe000u;
            l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z)) & 0xffffe000u;
            l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w)) & 0xffffe000u;
            if (p.mask_hi) hi[i] = h;
            lo[i] = l;
          }
          fence_proxy_async();
        }
        mbar_arrive_cluster(&conv_full[ls], 0);     // the leader's barrier counts both CTAs' splitter threads
      }
    }
  } else {
    // ===================================== epilogue (both CTAs, own 128 rows) =====================================
    const int ew = warp & 3;
    const int row_in_tile = ew * 32 + lane;
    const uint32_t lane_base = (uint32_t)(ew * 32) << 16;
    uint32_t uit = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
      int t = tile;
      const int n_tile = t % p.n_tiles; t /= p.n_tiles;
      const int m_tile = t % p.m_tiles; t /= p.m_tiles;
      const int split = t % p.splits;
      const int batch = t / p.splits;
      const int64_t b0 = batch / p.batch1, b1 = batch % p.batch1;
      const int kb0 = split * p.kblocks_per_split;
      int kb1 = kb0 + p.kblocks_per_split;
      if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
      const int nchunks = (kb1 - kb0 + p.chunk_kb - 1) / p.chunk_kb;
      const int64_t gm = (int64_t)m_tile * 256 + rank * TC_BM + row_in_tile;
      const bool row_ok = gm < p.m;
      float* crow;
      if (p.partial) crow = p.partial + (((int64_t)split * p.batches + batch) * p.m + gm) * p.n;
      else crow = p.c + b0 * p.c_stride0 + b1 * p.c_stride1 + gm * p.ldc;
      for (int ch = 0; ch < nchunks; ++ch, ++uit) {
        const uint32_t buf = p.chunked ? 0u : (uit & 1);
        const uint32_t buf_ph = p.chunked ? (uit & 1) : ((uit >> 1) & 1);
        mbar_wait(&tmem_full[buf], buf_ph);
        tc_fence_after();
        const uint32_t x_addr = tmem_base + lane_base + buf * BN;
        const uint32_t y_addr = tmem_base + lane_base + BN;
        const bool last = (ch == nchunks - 1);
        if (!last) {
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t x[32];
            tmem_ld_32x32b_x32(x_addr + c0, x);
            if (ch > 0) {
              uint32_t y[32];
              tmem_ld_32x32b_x32(y_addr + c0, y);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] = __float_as_uint(__uint_as_float(x[j]) + __uint_as_float(y[j]));
            } else {
              tmem_ld_wait();
            }
            tmem_st_32x32b_x32(y_addr + c0, x);
          }
          tmem_st_wait();
        } else {
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(x_addr + c0, r);
            if (nchunks > 1) {
              uint32_t y[32];
              tmem_ld_32x32b_x32(y_addr + c0, y);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(y[j]));
            } else {
              tmem_ld_wait();
            }
            const int64_t gn0 = (int64_t)n_tile * BN + c0;
            if (row_ok && gn0 < p.n) {
              float* dst = crow + gn0;
              const bool full = (gn0 + 32 <= p.n);
              float v[32];
              if (p.partial) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = p.alpha * __uint_as_float(r[j]);
                if (p.beta != 0.f) {
                  for (int j = 0; j < 32 && gn0 + j < p.n; ++j) v[j] += p.beta * dst[j];
                }
                if (p.bias) {
                  for (int j = 0; j < 32 && gn0 + j < p.n; ++j) v[j] += __ldg(p.bias + gn0 + j);
                }
                if (p.act == 1) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = tc_act<1>(v[j]);
                } else if (p.act == 2) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = tc_act<2>(v[j]);
                } else if (p.act == 3) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = tc_act<3>(v[j]);
                }
              }
              if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
                for (int j = 0; j < 32 && gn0 + j < p.n; ++j) dst[j] = v[j];
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive_cluster(&tmem_empty[buf], 0);    // leader's MMA warp waits for both CTAs' epilogues
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();          // the peer's smem / barriers / TMEM stay valid until both CTAs are done
  tc_fence_after();
  if (warp == 1) tmem_dealloc2(tmem_base, TMEM_COLS);
}

// =========================================================================================================
// CTA-pair kernel, generation 3 (TF32X3): 256 x 256 tiles on cta_group::2 with
//   * the A operand in TENSOR MEMORY (TS-mode MMAs): the splitter writes hi/lo A rows straight into a 4-stage
//     TMEM ring, so A costs no shared-memory write and no MMA-side shared-memory read;
//   * the running total of the chunked accumulation in REGISTERS (8 epilogue warps x 128 columns per thread),
//     which frees the TMEM columns the A ring needs (X 256 + A ring 4 x 64 = 512).
// Shared-memory traffic per k-block and CTA drops from ~192 KB (gemm_tc2_kernel: at the 128 B/clk limit) to
// ~128 KB (TMA 32 + splitter 48 + MMA B-half reads 48), which moves the kernel from shared-memory-bound to
// MMA-bound.  One ring of 4 stages; a stage = {raw A 16 KB, raw B half 16 KB, lo B half 16 KB, A hi/lo 64 TMEM
// columns}, filled by TMA, converted by the splitter, released by one multicast tcgen05.commit.
// Warps (512 threads = 4 warpgroups; setmaxnreg moves registers to the epilogue):
//   0 TMA   1 MMA (leader CTA)   2-3 idle   4-7 splitter (A rows -> TMEM, then the B lo tile)
//   8-15 epilogue (lane quarter = warp & 3, column half = (warp - 8) >> 2)
// =========================================================================================================
constexpr int TC3_THREADS = 512;
constexpr int TC3_STAGES = 4;
constexpr uint32_t TC3_STAGE_BYTES = 3 * 16384;

__device__ __forceinline__ uint4 tc_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t tc_lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tc_sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC3_THREADS, 1)
gemm_tc3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_c, const TcParams p) {
  constexpr int BN = 256;
  constexpr int R = TC3_STAGES;
  constexpr uint32_t A_BYTES = TC_BM * TC_BK * 4;            // 16 KB: this CTA's 128 rows of A (raw)
  constexpr uint32_t B_BYTES = (BN / 2) * TC_BK * 4;         // 16 KB: this CTA's half of the B tile
  constexpr uint32_t A_RING_COL = 256;                       // TMEM: X [0,256), A ring stage s: hi 256+64s, lo +32
  constexpr uint32_t TMEM_COLS = 512;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[3 * R + 4];
  __shared__ uint32_t tmem_base_smem;
  uint64_t* raw_full = bars;             // [R] TMA -> splitter (local)
  uint64_t* conv_full = bars + R;        // [R] splitter warps of BOTH CTAs -> leader's MMA warp (8 arrivals)
  uint64_t* empty = bars + 2 * R;        // [R] MMA commit (multicast) -> TMA + splitter of both CTAs
  uint64_t* tmem_full = bars + 3 * R;    // [2: column half] MMA commit (multicast) -> epilogue of both CTAs
  uint64_t* tmem_empty = tmem_full + 2;  // [2] all epilogue warps of both CTAs (16 arrivals) -> leader's MMA warp

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < R; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&conv_full[s], 2 * 4);
      mbar_init(&empty[s], 1);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(&tmem_full[h], 1);
      mbar_init(&tmem_empty[h], 2 * 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(&tmem_base_smem, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int total_tiles = p.batches * p.splits * p.m_tiles * p.n_tiles;   // m_tiles counts 256-row tiles
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ===================================== TMA producer (both CTAs) =====================================
      if (elect_one()) {
        uint32_t it = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
          int t = tile;
          const int n_tile = t % p.n_tiles; t /= p.n_tiles;
          const int m_tile = t % p.m_tiles; t /= p.m_tiles;
          const int split = t % p.splits;
          const int batch = t / p.splits;
          const int b0 = (int)(batch / p.batch1), b1 = (int)(batch % p.batch1);
          const int kb0 = split * p.kblocks_per_split;
          int kb1 = kb0 + p.kblocks_per_split;
          if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
          const int m0 = m_tile * 256 + (int)rank * TC_BM;
          // column half h of the tile = columns [128h, 128h + 128); CTA r stages columns 128h + 64r .. + 64 of each half
          const int n0 = n_tile * BN + (int)rank * 64;
          for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const int s = it % R;
            mbar_wait(&empty[s], ((it / R) & 1) ^ 1);
            uint8_t* sa = smem + (size_t)s * TC3_STAGE_BYTES;
            uint8_t* sb = sa + A_BYTES;
            mbar_expect_tx(&raw_full[s], A_BYTES + B_BYTES);
            const int k0 = kb * TC_BK;
            if (!p.a_mn_major) {
              tma_load_4d(&map_a, &raw_full[s], sa, k0, m0, b1, b0);
            } else {
#pragma unroll
              for (int c = 0; c < TC_BM / 32; ++c)
                tma_load_4d(&map_a, &raw_full[s], sa + c * (128 * TC_BK), m0 + c * 32, k0, b1, b0);
            }
            if (!p.b_mn_major) {
              tma_load_4d(&map_b, &raw_full[s], sb, k0, n0, b1, b0);                 // 64-row boxes
              tma_load_4d(&map_b, &raw_full[s], sb + 8192, k0, n0 + 128, b1, b0);
            } else {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                tma_load_4d(&map_b, &raw_full[s], sb + c * (128 * TC_BK), n0 + (c >> 1) * 128 + (c & 1) * 32, k0, b1, b0);
            }
          }
        }
      }
    } else if (warp == 1 && rank == 0) {
      // ===================================== MMA issuer (leader CTA) =====================================
      // A from TMEM is always K-major; idesc: D=f32, A=B=tf32, N=128 (one column half per MMA), M=256.
      // The two column halves are separate accumulation chains advanced in sub-steps of two k-blocks
      // (a a b b | a a b b per 128-K chunk), so the epilogue folds half a while the MMAs of half b run:
      // the fold latency is hidden without a second accumulator buffer.
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.b_mn_major << 16) |
                             ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t b_lbo = p.b_mn_major ? 128u * TC_BK : 16u, b_sbo = p.b_mn_major ? 512u : 32u * TC_BK;
      const uint32_t b_lay = p.b_mn_major ? 1u : 2u, b_kstep = p.b_mn_major ? 1024u : 32u;
      uint32_t it = 0, uit = 0;
#ifdef TC3_PROFILE
      long long w_conv = 0, w_empty = 0; const long long t_start = clock64();
      long long w_conv_kb[16] = {0}, w_empty_ch[8] = {0};     // by k-block / chunk index inside the tile (first 16 / 4)
#endif
      for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
        int t = tile / (p.n_tiles * p.m_tiles);
        const int split = t % p.splits;
        const int kb0 = split * p.kblocks_per_split;
        int kb1 = kb0 + p.kblocks_per_split;
        if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
        for (int cb0 = kb0; cb0 < kb1; cb0 += p.chunk_kb, ++uit) {
          int cb1 = cb0 + p.chunk_kb;
          if (cb1 > kb1) cb1 = kb1;
          for (int s0 = cb0; s0 < cb1; s0 += 2) {
            const int s1 = (s0 + 2 < cb1) ? s0 + 2 : cb1;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (s0 == cb0) {
#ifdef TC3_PROFILE
                const long long q0 = clock64();
#endif
                mbar_wait(&tmem_empty[h], (uit & 1) ^ 1);     // that half's epilogues have read the previous chunk sum
#ifdef TC3_PROFILE
                { const long long dq = clock64() - q0; w_empty += dq; const int ci = (cb0 - kb0) / p.chunk_kb; if (ci < 4) w_empty_ch[ci * 2 + h] += dq; }
#endif
                tc_fence_after();
              }
              for (int kb = s0; kb < s1; ++kb) {
                const uint32_t its = it + (uint32_t)(kb - s0);
                const int s = its % R;
                if (h == 0) {
#ifdef TC3_PROFILE
                  const long long q0 = clock64();
#endif
                  mbar_wait(&conv_full[s], (its / R) & 1);    // both CTAs: A hi/lo in TMEM, B lo tile in smem
#ifdef TC3_PROFILE
                  { const long long dq = clock64() - q0; w_conv += dq; if (kb - kb0 < 16) w_conv_kb[kb - kb0] += dq; }
#endif
                  tc_fence_after();
                }
                if (elect_one()) {
                  const uint32_t sb = smem_u32(smem + (size_t)s * TC3_STAGE_BYTES) + A_BYTES + (uint32_t)h * 8192u;
                  const uint64_t bh = make_smem_desc(sb, b_lbo, b_sbo, b_lay);
                  const uint64_t bl = make_smem_desc(sb + B_BYTES, b_lbo, b_sbo, b_lay);
                  const uint32_t ah = tmem_base + A_RING_COL + (uint32_t)s * 64u, al = ah + 32u;
                  const uint32_t d = tmem_base + (uint32_t)h * 128u;
                  const bool first_kb = (kb == cb0);
#pragma unroll
                  for (int kk = 0; kk < TC_BK / 8; ++kk)
                    umma2_tf32_ts(d, al + kk * 8, bh + (uint64_t)((kk * b_kstep) >> 4), idesc, (first_kb && kk == 0) ? 0u : 1u);
#pragma unroll
                  for (int kk = 0; kk < TC_BK / 8; ++kk)
                    umma2_tf32_ts(d, ah + kk * 8, bl + (uint64_t)((kk * b_kstep) >> 4), idesc, 1u);
#pragma unroll
                  for (int kk = 0; kk < TC_BK / 8; ++kk)
                    umma2_tf32_ts(d, ah + kk * 8, bh + (uint64_t)((kk * b_kstep) >> 4), idesc, 1u);
                  if (h == 1) umma2_commit_mc(&empty[s]);      // frees the stage (smem + A ring columns) in BOTH CTAs
                  if (kb == cb1 - 1) umma2_commit_mc(&tmem_full[h]);
                }
                __syncwarp();
              }
            }
            it += (uint32_t)(s1 - s0);
          }
        }
      }
#ifdef TC3_PROFILE
      if (blockIdx.x == 0 && lane == 0)
      {
        printf("[tc3] MMA warp: total %lld cycles, k-blocks %u, wait conv %lld, wait tmem_empty %lld\n", clock64() - t_start, it, w_conv, w_empty);
        printf("[tc3]   wait conv by k-block in tile:");
        for (int i = 0; i < 16; ++i) printf(" %lld", w_conv_kb[i]);
        printf("\n[tc3]   wait tmem_empty by (chunk, half):");
        for (int i = 0; i < 8; ++i) printf(" %lld", w_empty_ch[i]);
        printf("\n");
      }
#endif
    }
  } else if (warp < 8) {
    // ===================================== splitter (both CTAs) =====================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    const int q = warp & 3;
    const int r = q * 32 + lane;                            // A part: the row (TMEM lane) this thread owns
    const int btid = threadIdx.x - 128;                     // B part: 0..127
    uint32_t it = 0;
#ifdef TC3_PROFILE
    long long w_raw = 0, t_a = 0, t_b = 0, t_w = 0, t_f = 0, t_f1 = 0, t_f2 = 0, t_f3 = 0; const long long t_start = clock64();
#endif
    for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
      int t = tile / (p.n_tiles * p.m_tiles);
      const int split = t % p.splits;
      const int kb0 = split * p.kblocks_per_split;
      int kb1 = kb0 + p.kblocks_per_split;
      if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % R;
#ifdef TC3_PROFILE
        const long long q0 = clock64();
#endif
        mbar_wait(&raw_full[s], (it / R) & 1);
#ifdef TC3_PROFILE
        w_raw += clock64() - q0;
        const long long p0 = clock64();
#endif

        const uint32_t sa = smem_u32(smem + (size_t)s * TC3_STAGE_BYTES);
        {
          uint32_t v[32], l[32];
          if (!p.a_mn_major) {          // K-major, SWIZZLE_128B: 16 B chunk c of row r sits at chunk c ^ (r & 7)
            const uint32_t row = sa + r * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint4 t4 = tc_lds128(row + ((c ^ (r & 7)) << 4));
              v[4 * c] = t4.x; v[4 * c + 1] = t4.y; v[4 * c + 2] = t4.z; v[4 * c + 3] = t4.w;
            }
          } else {                      // MN-major, SWIZZLE_128B_BASE32B: 32 B granule g of k-row kk sits at g ^ (kk & 3)
            const uint32_t chunk = sa + (r >> 5) * (128 * TC_BK);
            const int ml = r & 31;
#pragma unroll
            for (int kk = 0; kk < 32; ++kk)
              v[kk] = tc_lds32(chunk + kk * 128 + ((((ml >> 3) ^ (kk & 3)) << 5) | ((ml & 7) << 2)));
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const uint32_t h = v[j] & 0xffffe000u;
            l[j] = __float_as_uint(__uint_as_float(v[j]) - __uint_as_float(h)) & 0xffffe000u;
            v[j] = h;
          }
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + A_RING_COL + (uint32_t)s * 64u;
          tmem_st_32x32b_x32(ta, v);
          tmem_st_32x32b_x32(ta + 32, l);
        }
#ifdef TC3_PROFILE
        const long long p1 = clock64(); t_a += p1 - p0;
#endif
        {
          const uint32_t sb = sa + A_BYTES + btid * 16;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 bv = tc_lds128(sb + i * 2048);
            uint4 lo;
            lo.x = __float_as_uint(__uint_as_float(bv.x) - __uint_as_float(bv.x & 0xffffe000u)) & 0xffffe000u;
            lo.y = __float_as_uint(__uint_as_float(bv.y) - __uint_as_float(bv.y & 0xffffe000u)) & 0xffffe000u;
            lo.z = __float_as_uint(__uint_as_float(bv.z) - __uint_as_float(bv.z & 0xffffe000u)) & 0xffffe000u;
            lo.w = __float_as_uint(__uint_as_float(bv.w) - __uint_as_float(bv.w & 0xffffe000u)) & 0xffffe000u;
            tc_sts128(sb + B_BYTES + i * 2048, lo);
          }
        }
#ifdef TC3_PROFILE
        const long long p2 = clock64(); t_b += p2 - p1;
#endif
        tmem_st_wait();
#ifdef TC3_PROFILE
        const long long p3 = clock64(); t_w += p3 - p2;
#endif
        tc_fence_before();
#ifdef TC3_PROFILE
        const long long p4 = clock64(); t_f1 += p4 - p3;
#endif
        fence_proxy_async();
#ifdef TC3_PROFILE
        const long long p5 = clock64(); t_f2 += p5 - p4;
#endif
        __syncwarp();
#ifdef TC3_PROFILE
        const long long p6 = clock64(); t_f3 += p6 - p5;
#endif
        // one (cumulative) release per warp
        if (lane == 0) mbar_arrive_cluster(&conv_full[s], 0);     // the leader's barrier counts both CTAs' splitter warps
#ifdef TC3_PROFILE
        t_f += clock64() - p3;
#endif
      }
    }
#ifdef TC3_PROFILE
    if (blockIdx.x == 0 && threadIdx.x == 128)
      printf("[tc3] splitter: total %lld cycles, wait raw_full (TMA) %lld, A part %lld, B part %lld, st wait %lld, fences+arrive %lld (tcgen05 fence %lld, proxy fence %lld, syncwarp %lld)\n",
             clock64() - t_start, w_raw, t_a, t_b, t_w, t_f, t_f1, t_f2, t_f3);
#endif
  } else {
    // ===================================== epilogue (both CTAs, own 128 rows) =====================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
    const int q = warp & 3;
    const int hf = (warp - 8) >> 2;                         // which 64 columns of EACH 128-column half this warp folds
    const int row_in_tile = q * 32 + lane;
    const uint32_t x_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)hf * 64u;
    uint32_t uit = 0;
#ifdef TC3_PROFILE
    long long w_full = 0, t_fold = 0, t_store = 0; const long long t_start = clock64();
#endif
    for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
      int t = tile;
      const int n_tile = t % p.n_tiles; t /= p.n_tiles;
      const int m_tile = t % p.m_tiles; t /= p.m_tiles;
      const int split = t % p.splits;
      const int batch = t / p.splits;
      const int64_t b0 = batch / p.batch1, b1 = batch % p.batch1;
      const int kb0 = split * p.kblocks_per_split;
      int kb1 = kb0 + p.kblocks_per_split;
      if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
      const int nchunks = (kb1 - kb0 + p.chunk_kb - 1) / p.chunk_kb;
      const int64_t gm = (int64_t)m_tile * 256 + rank * TC_BM + row_in_tile;
      const bool row_ok = gm < p.m;
      float* crow;
      if (p.partial) crow = p.partial + (((int64_t)split * p.batches + batch) * p.m + gm) * p.n;
      else crow = p.c + b0 * p.c_stride0 + b1 * p.c_stride1 + gm * p.ldc;
      // Tile epilogue.  Code size matters here: a fully unrolled alpha/beta/bias/activation epilogue over 128 register
      // columns made this kernel 24 k instructions (390 KB) and the store phase instruction-fetch bound (~66 k cycles
      // per tile, measured).  So the unrolled part only moves the registers into a per-warp 4 KB staging tile
      // (32 rows x 32 columns, SWIZZLE_128B; every index into y[] is a compile-time constant), a ROLLED loop applies
      // the epilogue math on row-contiguous segments, and the tile leaves either through cp.async.bulk.tensor stores
      // (asynchronous; rows / columns beyond M / N are clipped by the tensor map) or through coalesced stores.
      // running totals: y[0..63] = columns hf*64.. of half a, y[64..127] = the same columns of half b.  All eight warps
      // fold every half (64 columns each), so a chunk sum leaves TMEM in half the time of a 4-warp fold and the
      // tmem_empty round trip fits inside the two k-blocks the other half still has queued.
      float y[128];
      for (int ch = 0; ch < nchunks; ++ch, ++uit) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#ifdef TC3_PROFILE
          const long long q0 = clock64();
#endif
          mbar_wait(&tmem_full[h], uit & 1);
#ifdef TC3_PROFILE
          const long long q1 = clock64(); w_full += q1 - q0;
#endif
          tc_fence_after();
          // fold the chunk sum into the running total with round-to-nearest adds (TMEM accumulation truncates)
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t x[16];
            tmem_ld_32x32b_x16(x_addr + (uint32_t)(h * 128 + c0), x);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
              y[h * 64 + c0 + j] = (ch == 0) ? __uint_as_float(x[j]) : y[h * 64 + c0 + j] + __uint_as_float(x[j]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tmem_empty[h], 0);   // this half of X may be overwritten by the next chunk's MMAs
#ifdef TC3_PROFILE
          const long long q2 = clock64(); t_fold += q2 - q1;
#endif
          // Last chunk: this half is final, so its 64 columns leave NOW: half a's stores run under half b's last
          // two k-blocks of MMAs, half b's under the next tile's first chunk (the SM accepts ~21-32 B/clk of stores, so
          // the 128 KB tile costs ~5 k cycles of the epilogue warps' time; what cannot be hidden between the folds shows
          // up as a ~1.1 k-cycle tmem_empty stall per tile, whichever way the four segments are ordered - measured).
          if (ch == nchunks - 1) {
            const uint32_t stg = smem_u32(smem + (size_t)R * TC3_STAGE_BYTES) + (uint32_t)(warp - 8) * 4096u;
            const int m_row0 = m_tile * 256 + (int)rank * TC_BM + q * 32;       // first row of this warp
            const int c2 = p.partial ? batch : (int)b1, c3 = p.partial ? split : (int)b0;
            const bool identity = p.partial || (p.alpha == 1.f && p.beta == 0.f && !p.bias && p.act == 0);
            float* cwarp;
            if (p.partial) cwarp = p.partial + (((int64_t)split * p.batches + batch) * p.m + m_row0) * p.n;
            else cwarp = p.c + b0 * p.c_stride0 + b1 * p.c_stride1 + (int64_t)m_row0 * p.ldc;
            const int64_t row_pitch = p.partial ? p.n : p.ldc;
            const int sub_row = lane >> 3, chunk = lane & 7;
    #pragma unroll
          for (int c0 = h * 64; c0 < h * 64 + 64; c0 += 32) {
            const int n_col0 = n_tile * BN + (c0 >> 6) * 128 + hf * 64 + (c0 & 63);
            if (n_col0 < p.n && m_row0 < p.m && !(p.probe & 4)) {            // warp-uniform; probe bit 4 skips the stores
              if (p.tma_store && elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging free
              __syncwarp();
    #pragma unroll
              for (int c = 0; c < 8; ++c) {
                uint4 v4;
                v4.x = __float_as_uint(y[c0 + 4 * c]); v4.y = __float_as_uint(y[c0 + 4 * c + 1]);
                v4.z = __float_as_uint(y[c0 + 4 * c + 2]); v4.w = __float_as_uint(y[c0 + 4 * c + 3]);
                tc_sts128(stg + lane * 128 + ((c ^ (lane & 7)) << 4), v4);
              }
              __syncwarp();
              if (!identity || !p.tma_store) {
                const int gn0 = n_col0 + chunk * 4;
    #pragma unroll 1
                for (int i = 0; i < 8; ++i) {
                  const int rl = 4 * i + sub_row;
                  const uint32_t sa = stg + rl * 128 + ((chunk ^ (rl & 7)) << 4);
                  const uint4 v4 = tc_lds128(sa);
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
                      if (p.act == 1) v[j] = tc_act<1>(v[j]);
                      else if (p.act == 2) v[j] = tc_act<2>(v[j]);
                      else if (p.act == 3) v[j] = tc_act<3>(v[j]);
                    }
                  }
                  if (p.tma_store) {
                    uint4 o4;
                    o4.x = __float_as_uint(v[0]); o4.y = __float_as_uint(v[1]); o4.z = __float_as_uint(v[2]); o4.w = __float_as_uint(v[3]);
                    tc_sts128(sa, o4);
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
#ifdef TC3_PROFILE
            t_store += clock64() - q2;
#endif
          }
        }
      }
    }
    if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every output tile has been written
#ifdef TC3_PROFILE
    if (blockIdx.x == 0 && threadIdx.x == 256)
      printf("[tc3] epilogue: total %lld cycles, chunks %u, wait tmem_full %lld, fold %lld, store %lld\n", clock64() - t_start, uit, w_full, t_fold, t_store);
#endif
  }

  tc_fence_before();
  cluster_sync_all();          // the peer's smem / barriers / TMEM stay valid until both CTAs are done
  tc_fence_after();
  if (warp == 1) tmem_dealloc2(tmem_base, TMEM_COLS);
}

// split-K second pass: C = alpha * sum_s partial[s] + beta*C + bias, activation.  One thread per 4 outputs.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, int splits, TcParams p) {
  const int64_t per_batch = p.m * p.n;
  const int64_t total = per_batch * p.batches;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += __ldg(partial + (int64_t)sp * total + i);
    const int64_t batch = i / per_batch, rem = i - batch * per_batch;
    const int64_t gm = rem / p.n, gn = rem - gm * p.n;
    const int64_t b0 = batch / p.batch1, b1 = batch % p.batch1;
    float* dst = p.c + b0 * p.c_stride0 + b1 * p.c_stride1 + gm * p.ldc + gn;
    float v = p.alpha * s;
    if (p.beta != 0.f) v += p.beta * *dst;
    if (p.bias) v += __ldg(p.bias + gn);
    if (p.act == 1) v = tc_act<1>(v);
    else if (p.act == 2) v = tc_act<2>(v);
    else if (p.act == 3) v = tc_act<3>(v);
    *dst = v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    // resolved through the runtime: no link-time dependency on libcuda (the library must load on CPU-only hosts)
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

struct OperandPlan {
  bool mn_major;
  int64_t inner, rows, ld;     // global matrix as stored: `rows` rows of `inner` contiguous elements
  int box_rows;
};

static int encode_operand(CUtensorMap* map, const float* base, const OperandPlan& o, int64_t batch0, int64_t batch1,
                          int64_t stride0, int64_t stride1) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(B200AD_ERR_CUDA, "gemm(tcgen05): cuTensorMapEncodeTiled is not available");
  // dims: inner, rows, batch1, batch0.  A size-1 batch axis may carry stride 0: give it a legal dummy stride.
  cuuint64_t dims[4] = {(cuuint64_t)o.inner, (cuuint64_t)o.rows, (cuuint64_t)batch1, (cuuint64_t)batch0};
  const int64_t fallback = o.ld * o.rows;
  int64_t s1 = (batch1 > 1) ? stride1 : fallback, s0 = (batch0 > 1) ? stride0 : fallback;
  if (s1 <= 0) s1 = fallback;
  if (s0 <= 0) s0 = fallback;
  cuuint64_t strides[3] = {(cuuint64_t)o.ld * 4, (cuuint64_t)s1 * 4, (cuuint64_t)s0 * 4};
  cuuint32_t box[4] = {o.mn_major ? 32u : (cuuint32_t)TC_BK, (cuuint32_t)o.box_rows, 1u, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   o.mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                              : (TC_BK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(B200AD_ERR_CUDA, "gemm(tcgen05): cuTensorMapEncodeTiled failed (%d)", (int)r);
  return B200AD_OK;
}

// The tuning / debugging knobs of this file, read from the environment ONCE (ten getenv calls per launch were 2-5 us of
// host time on paths that are host-bound: the recurrent steps, the B = 16 attention layer).  b200ad_gemm_reload_env()
// re-reads them (profiles/gemm_tc_probe.py flips them between launches).
struct TcEnv {      // parsed values (not the getenv pointers: the host language may replace the strings later)
  bool has_bn, has_pair, has_chunk_kb, no_tma_store;
  int bn, pair, stages, splits, mask_hi, atmem, probe, chunk_kb;
};
static std::atomic<int> g_tc_env_gen{1};
static const TcEnv& tc_env() {
  static TcEnv env;
  static int seen = 0;
  static std::mutex mu;
  const int gen = g_tc_env_gen.load(std::memory_order_acquire);
  if (seen != gen) {
    std::lock_guard<std::mutex> lock(mu);
    auto num = [](const char* name, int dflt, bool* has = nullptr) {
      const char* v = getenv(name);
      if (has) *has = v != nullptr;
      return v ? atoi(v) : dflt;
    };
    env.bn = num("B200AD_TC_BN", 0, &env.has_bn);
    env.pair = num("B200AD_TC_PAIR", 3, &env.has_pair);
    env.stages = num("B200AD_TC_STAGES", 0);
    env.splits = num("B200AD_TC_SPLITS", 0);
    env.mask_hi = num("B200AD_TC_MASK_HI", 0);
    env.atmem = num("B200AD_TC_ATMEM", 1);
    env.probe = num("B200AD_TC_PROBE", 0);
    env.chunk_kb = num("B200AD_TC_CHUNK_KB", 0, &env.has_chunk_kb);
    env.no_tma_store = getenv("B200AD_TC_NO_TMA_STORE") != nullptr;
    seen = gen;
  }
  return env;
}
void gemm_tc_reload_env() { g_tc_env_gen.fetch_add(1, std::memory_order_acq_rel); }

static int pick_bn(int64_t n) {
  if (tc_env().has_bn) {
    const int v = tc_env().bn;
    if (v == 64 || v == 128 || v == 256) return v;
  }
  if (n <= 64) return 64;
  if (n <= 128) return 128;
  return 256;
}

struct TcPlan {
  int bn, stages, splits, kblocks_total, kblocks_per_split, m_tiles, n_tiles, batches;
  uint32_t smem_bytes;
  bool x3;
  bool pair;     // CTA-pair kernel (cta_group::2, 256 x 256 tiles)
  bool gen3;     // ... generation 3: A operand in TMEM, running totals in registers (gemm_tc3_kernel)
};

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = B200AD_NUM_SMS;
  }
  return n;
}

static TcPlan make_plan(const GemmArgs& g, int precision) {
  TcPlan pl;
  pl.x3 = (precision == B200AD_GEMM_TF32X3);
  pl.bn = pick_bn(g.n);
  // One row tile only (the M <= 128 recurrent steps of config 4): 64-column tiles (A in TMEM, 32-cycle MMAs) spread over
  // the SMs by a deeper split-K.  Measured on the 64 x 1024 x 1024 step inside a CUDA graph: BN = 256 / 4 splits 19.3 us,
  // BN = 64 / 8 splits 9.7 us, 16 splits 11.6 us (profiles/r1_summary.md, config 4).  A fused "last CTA sums the
  // partials" epilogue was measured too: 38.8 us (one CTA's 128 threads are far slower than the grid-wide second pass).
  const bool single_row_tile = pl.x3 && g.m <= TC_BM && g.n >= 256 && g.batch0 * g.batch1 == 1 && !tc_env().has_bn;
  // ... and 128-column tiles when they still fill the machine with >= 128 K-elements per split: half the re-reads of the
  // activation rows, which every column tile loads (the LSTM's 4H-wide products: 64 x 4096 x 1024 15.2 vs 18.3 us,
  // 64 x 1024 x 4096 14.7 vs 17.0 us per product + reduce, profiles/lstm_step_gemm_probe.py; BN = 256: 20 us)
  if (single_row_tile) pl.bn = (ceil_div(g.n, 128) * (g.k / 128) >= num_sms()) ? 128 : 64;
  // CTA-pair kernel: opt-in.  Measured on B200 (profiles/r1_gemm_tc_probe.md): +3..7 % on 4096^3 / 8192^3, but
  // -5 % on the (131072 x 512 x 512) projection shapes of config 5 (the chunk fold waits for two epilogues).
  const int pair_mode = tc_env().pair;
  // Measured on B200 (profiles/r1_summary.md §4): the generation-3 pair kernel (A in TMEM, running totals in registers,
  // interleaved half-tile chains, compact TMA-store epilogue) beats the 1-CTA kernel on every 256-wide shape:
  // 131072 x 512 x 512 projections 150 -> 236, 8192^3 192 -> 260, split-K dW 179 -> 234, K = 256 gates 139 -> 225 TFLOP/s.
  // B200AD_TC_PAIR: 0 = never, 1 = gemm_tc2_kernel, 2 = gen3 for deep-K problems only, 3 = gen3 whenever eligible.
  const int64_t m_tiles256 = ceil_div(g.m, 2 * TC_BM), n_tiles256 = ceil_div(g.n, 256);
  const int64_t tiles256 = m_tiles256 * n_tiles256 * g.batch0 * g.batch1;
  const int64_t kblocks = ceil_div(g.k, TC_BK);
  const bool deep_k = (tiles256 >= num_sms() / 2) ? (kblocks >= 128) : (kblocks * tiles256 >= 64 * (num_sms() / 2));
  pl.pair = pl.x3 && pl.bn == 256 && g.m > TC_BM && (pair_mode == 1 || pair_mode == 3 || (pair_mode == 2 && deep_k));
  pl.gen3 = pl.pair && pair_mode != 1;
  const uint32_t stage = pl.pair ? (TC_BM * TC_BK * 4 + (pl.bn / 2) * TC_BK * 4) * 2
                                 : (TC_BM * TC_BK * 4 + pl.bn * TC_BK * 4) * (pl.x3 ? 2 : 1);
  int st = (int)((TC_SMEM_LIMIT - 1024 - 256) / stage);
  if (st > 8) st = 8;
  if (tc_env().stages >= 1 && tc_env().stages <= st) st = tc_env().stages;
  pl.stages = st;
  pl.smem_bytes = (uint32_t)st * stage + 1024 + 256;
  if (pl.pair) {   // 4 raw stages + 2 lo stages of 32 KB
    pl.stages = 4;
    pl.smem_bytes = 6u * (TC_BM * TC_BK * 4 + (pl.bn / 2) * TC_BK * 4) + 1024 + 256;
  }
  if (pl.gen3) pl.smem_bytes = TC3_STAGES * TC3_STAGE_BYTES + 8 * 4096 + 1024;   // ring + per-warp store staging
  pl.m_tiles = (int)ceil_div(g.m, pl.pair ? 2 * TC_BM : TC_BM);
  pl.n_tiles = (int)ceil_div(g.n, pl.bn);
  pl.batches = (int)(g.batch0 * g.batch1);
  pl.kblocks_total = (int)ceil_div(g.k, TC_BK);
  // split-K when the output tiles cannot fill the machine and K is deep
  const int64_t tiles = (int64_t)pl.m_tiles * pl.n_tiles * pl.batches;
  int splits = 1;
  const int sms = pl.pair ? num_sms() / 2 : num_sms();   // schedulable units: CTAs or CTA pairs
  if (tiles < sms && pl.kblocks_total * TC_BK >= 512) {
    splits = (int)(sms / tiles);
    // keep >= 256 K-elements per split (one 128-K chunk when a single row tile has to fill the machine)
    const int max_by_k = pl.kblocks_total * TC_BK / (single_row_tile ? 128 : 256);
    if (splits > max_by_k) splits = max_by_k;
    if (splits < 1) splits = 1;
    if (splits > 64) splits = 64;
  }
  if (tc_env().splits >= 1) splits = tc_env().splits;
  pl.kblocks_per_split = (int)ceil_div(pl.kblocks_total, splits);
  pl.splits = (int)ceil_div(pl.kblocks_total, pl.kblocks_per_split);   // no empty splits
  return pl;
}

bool gemm_tc_supported(const GemmArgs& g, int precision, const char** why) {
  auto no = [&](const char* w) { if (why) *why = w; return false; };
  if (precision != B200AD_GEMM_TF32X3 && precision != B200AD_GEMM_TF32)
    return no("only TF32 / TF32X3 are built on the tcgen05 path (bf16 variant pending)");
  if (g.m < 1 || g.n < 1 || g.k < 1) return no("empty problem");
  if (g.k < 16 || g.m * g.n * g.k < (int64_t)(1 << 18)) return no("problem too small for the TMA path");
  if ((reinterpret_cast<uintptr_t>(g.a) & 15) || (reinterpret_cast<uintptr_t>(g.b) & 15)) return no("operand base not 16B aligned");
  if ((g.lda & 3) || (g.ldb & 3)) return no("leading dimension not a multiple of 4 elements (TMA stride must be 16B)");
  if ((g.a_stride0 & 3) || (g.a_stride1 & 3) || (g.b_stride0 & 3) || (g.b_stride1 & 3)) return no("batch stride not 16B aligned");
  if (g.batch0 > 1 && (g.a_stride0 <= 0 || g.b_stride0 <= 0)) return no("broadcast (stride 0) batch operands use the SIMT path");
  if (g.batch1 > 1 && (g.a_stride1 <= 0 || g.b_stride1 <= 0)) return no("broadcast (stride 0) batch operands use the SIMT path");
  if (g.m >= (1ll << 31) || g.n >= (1ll << 31) || g.k >= (1ll << 31)) return no("dimension too large");
  if (g.batch0 * g.batch1 >= (1ll << 24)) return no("too many batches");
  if (get_encode_fn() == nullptr) return no("cuTensorMapEncodeTiled unavailable (no driver)");
  return true;
}

int64_t gemm_tc_workspace_bytes(const GemmArgs& g, int precision) {
  if (g.m < 1 || g.n < 1 || g.k < 1) return 0;
  TcPlan pl = make_plan(g, precision);
  if (pl.splits <= 1) return 0;
  return (int64_t)pl.splits * pl.batches * g.m * g.n * 4;
}

int gemm_tc_launch(const GemmArgs& g, int precision, void* ws, int64_t ws_bytes, cudaStream_t s, TcSplitInfo* split_out) {
  TcPlan pl = make_plan(g, precision);
  if (split_out) *split_out = TcSplitInfo{nullptr, 1, 0};
  const int64_t need = pl.splits > 1 ? (int64_t)pl.splits * pl.batches * g.m * g.n * 4 : 0;
  if (need > 0 && (!ws || ws_bytes < need)) {
    // not enough workspace for split-K: run un-split
    pl.splits = 1;
    pl.kblocks_per_split = pl.kblocks_total;
  }
  OperandPlan oa, ob;
  oa.mn_major = g.trans_a != 0;                       // stored (K x M): M contiguous
  oa.inner = oa.mn_major ? g.m : g.k;
  oa.rows = oa.mn_major ? g.k : g.m;
  oa.ld = g.lda;
  oa.box_rows = oa.mn_major ? TC_BK : TC_BM;
  ob.mn_major = g.trans_b == 0;                       // stored (K x N): N contiguous
  ob.inner = ob.mn_major ? g.n : g.k;
  ob.rows = ob.mn_major ? g.k : g.n;
  ob.ld = g.ldb;
  ob.box_rows = ob.mn_major ? TC_BK : (pl.gen3 ? 64 : (pl.pair ? pl.bn / 2 : pl.bn));
  CUtensorMap map_a, map_b;
  int rc = encode_operand(&map_a, g.a, oa, g.batch0, g.batch1, g.a_stride0, g.a_stride1);
  if (rc) return rc;
  rc = encode_operand(&map_b, g.b, ob, g.batch0, g.batch1, g.b_stride0, g.b_stride1);
  if (rc) return rc;

  TcParams p;
  p.c = g.c; p.c_in = g.c; p.bias = g.bias;
  p.partial = pl.splits > 1 ? reinterpret_cast<float*>(ws) : nullptr;
  p.m = g.m; p.n = g.n; p.k = g.k;
  p.ldc = g.ldc; p.c_stride0 = g.c_stride0; p.c_stride1 = g.c_stride1;
  p.batch1 = g.batch1;
  p.batches = pl.batches; p.m_tiles = pl.m_tiles; p.n_tiles = pl.n_tiles; p.splits = pl.splits;
  p.kblocks_total = pl.kblocks_total; p.kblocks_per_split = pl.kblocks_per_split;
  p.a_mn_major = oa.mn_major; p.b_mn_major = ob.mn_major;
  p.x3 = pl.x3 ? 1 : 0;
  // tcgen05 reads fp32 smem operands as tf32 by IGNORING the 13 low mantissa bits (measured: identical error with
  // and without rewriting the hi tile, profiles/gemm_tc_probe.py "x3 without hi masking"), so the raw tile already
  // is Ah = trunc(A); only the lo tile is materialised.  B200AD_TC_MASK_HI=1 restores the explicit rewrite.
  p.mask_hi = tc_env().mask_hi;
  p.act = g.act; p.alpha = g.alpha; p.beta = g.beta;
  p.pdl = 0;
  p.a_tmem = tc_env().atmem;
  p.tma_store = 0;
  p.probe = tc_env().probe;
  // TMEM accumulation chain length: 128 K-elements (see the epilogue's fold).  Measured on B200
  // (profiles/gemm_tc_probe.py): max err/sqrt(K) 6.5e-6 at K=4096 (fp32 SIMT: 1.6e-5) for a 3 % slowdown; 8
  // k-blocks give 1.2e-5, no chunking grows linearly with K.  Single-pass TF32 is a 10-bit-mantissa probe
  // anyway, so it accumulates the whole K range in place.
  int chunk_kb = tc_env().has_chunk_kb ? tc_env().chunk_kb : 128 / TC_BK;
  if (chunk_kb < 1 || !pl.x3) chunk_kb = 1 << 30;
  p.chunked = (pl.kblocks_per_split > chunk_kb) ? 1 : 0;
  p.chunk_kb = p.chunked ? chunk_kb : (1 << 30);

  const int64_t total_tiles = (int64_t)pl.batches * pl.splits * pl.m_tiles * pl.n_tiles;
  if (total_tiles >= (1ll << 31)) return set_error(B200AD_ERR_UNSUPPORTED, "gemm(tcgen05): too many tiles");
  const int grid = (int)(total_tiles < num_sms() ? total_tiles : num_sms());
  cudaError_t e;
  if (pl.gen3) {
    static bool attr3 = false;
    if (!attr3) {
      e = cudaFuncSetAttribute(gemm_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT - 1024);
      if (e != cudaSuccess) return set_error(B200AD_ERR_CUDA, "gemm(tcgen05 pair/TS): smem attribute: %s", cudaGetErrorString(e));
      attr3 = true;
    }
    const int pairs = num_sms() / 2;
    const int grid3 = 2 * (int)(total_tiles < pairs ? total_tiles : pairs);
    // output tensor map (TMA stores): C itself, or the split-K partial buffer [split][batch][M][N]
    CUtensorMap map_c = map_a;
    p.tma_store = 0;
    {
      const float* cbase = p.partial ? p.partial : g.c;
      const int64_t ld = p.partial ? g.n : g.ldc;
      const int64_t s1 = p.partial ? g.m * g.n : g.c_stride1, s0 = p.partial ? (int64_t)pl.batches * g.m * g.n : g.c_stride0;
      const int64_t d2 = p.partial ? pl.batches : g.batch1, d3 = p.partial ? pl.splits : g.batch0;
      const bool ok = (p.partial || g.beta == 0.f) && aligned16(cbase) && (ld & 3) == 0 && (d2 <= 1 || (s1 > 0 && (s1 & 3) == 0)) &&
                      (d3 <= 1 || (s0 > 0 && (s0 & 3) == 0)) && !tc_env().no_tma_store;
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
    gemm_tc3_kernel<<<grid3, TC3_THREADS, pl.smem_bytes, s>>>(map_a, map_b, map_c, p);   // __cluster_dims__(2,1,1)
    count_launch();
    rc = check_launch("gemm(tcgen05 pair/TS)");
    if (rc) return rc;
    if (pl.splits > 1 && split_out) {
      *split_out = TcSplitInfo{reinterpret_cast<const float*>(ws), pl.splits, g.m * g.n * pl.batches};
    } else if (pl.splits > 1) {
      const int64_t total = g.m * g.n * pl.batches;
      splitk_reduce_kernel<<<stream_grid(total, 256), 256, 0, s>>>(reinterpret_cast<const float*>(ws), pl.splits, p);
      count_launch();
      rc = check_launch("gemm(tcgen05 split-K reduce)");
    }
    return rc;
  }
  if (pl.pair) {
    static bool attr2 = false;
    if (!attr2) {
      e = cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
      if (e != cudaSuccess) return set_error(B200AD_ERR_CUDA, "gemm(tcgen05 pair): smem attribute: %s", cudaGetErrorString(e));
      attr2 = true;
    }
    const int pairs = num_sms() / 2;
    const int grid2 = 2 * (int)(total_tiles < pairs ? total_tiles : pairs);
    gemm_tc2_kernel<<<grid2, TC_THREADS, pl.smem_bytes, s>>>(map_a, map_b, p, pl.stages);   // __cluster_dims__(2,1,1)
    count_launch();
    rc = check_launch("gemm(tcgen05 pair)");
    if (rc) return rc;
    if (pl.splits > 1 && split_out) {
      *split_out = TcSplitInfo{reinterpret_cast<const float*>(ws), pl.splits, g.m * g.n * pl.batches};
    } else if (pl.splits > 1) {
      const int64_t total = g.m * g.n * pl.batches;
      splitk_reduce_kernel<<<stream_grid(total, 256), 256, 0, s>>>(reinterpret_cast<const float*>(ws), pl.splits, p);
      count_launch();
      rc = check_launch("gemm(tcgen05 split-K reduce)");
    }
    return rc;
  }
#define LAUNCH_BN(BNV)                                                                                          \
  do {                                                                                                          \
    static bool attr_set = false;                                                                               \
    if (!attr_set) {                                                                                            \
      e = cudaFuncSetAttribute(gemm_tc_kernel<BNV>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT); \
      if (e != cudaSuccess) return set_error(B200AD_ERR_CUDA, "gemm(tcgen05): smem attribute: %s", cudaGetErrorString(e)); \
      attr_set = true;                                                                                          \
    }                                                                                                           \
    if (g.pdl) {                                                                                                \
      p.pdl = 1;                                                                                                \
      cudaLaunchConfig_t cfg = {};                                                                              \
      cfg.gridDim = dim3((unsigned)grid);                                                                       \
      cfg.blockDim = dim3(TC_THREADS);                                                                          \
      cfg.dynamicSmemBytes = pl.smem_bytes;                                                                     \
      cfg.stream = s;                                                                                           \
      cudaLaunchAttribute attr[1];                                                                              \
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                          \
      attr[0].val.programmaticStreamSerializationAllowed = 1;                                                   \
      cfg.attrs = attr;                                                                                         \
      cfg.numAttrs = 1;                                                                                         \
      e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BNV>, map_a, map_b, (const TcParams)p, (const int)pl.stages); \
      if (e != cudaSuccess) return set_error(B200AD_ERR_CUDA, "gemm(tcgen05): launch: %s", cudaGetErrorString(e)); \
    } else {                                                                                                    \
      gemm_tc_kernel<BNV><<<grid, TC_THREADS, pl.smem_bytes, s>>>(map_a, map_b, p, pl.stages);                  \
    }                                                                                                           \
  } while (0)
  if (pl.bn == 64) LAUNCH_BN(64);
  else if (pl.bn == 128) LAUNCH_BN(128);
  else LAUNCH_BN(256);
#undef LAUNCH_BN
  count_launch();
  rc = check_launch("gemm(tcgen05)");
  if (rc) return rc;
  if (pl.splits > 1 && split_out) {
    *split_out = TcSplitInfo{reinterpret_cast<const float*>(ws), pl.splits, g.m * g.n * pl.batches};
  } else if (pl.splits > 1) {
    const int64_t total = g.m * g.n * pl.batches;
    splitk_reduce_kernel<<<stream_grid(total, 256), 256, 0, s>>>(reinterpret_cast<const float*>(ws), pl.splits, p);
    count_launch();
    rc = check_launch("gemm(tcgen05 split-K reduce)");
  }
  return rc;
}

}  // namespace b200ad
