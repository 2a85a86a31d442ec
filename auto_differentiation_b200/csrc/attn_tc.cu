// Fused multi-head attention core on tcgen05 / TMEM / TMA (sm_100a): the S x S score / probability matrices
// never leave the SM.
//
// Replaces, for MultiHeadAttention (reference layers.py:1351-1357 forward, and the tape backward of
// Test/MultiHeadAttentionTest.py:70-85): np.matmul(Q, K^T) / sqrt(d_k), exp/sum softmax, np.matmul(P, V), and in
// the backward the five products dV = P^T dO, dP = dO V^T, dS = P o (dP - rowsum(dO o O)) / sqrt(d_k),
// dQ = dS K, dK = dS^T Q.
//
// One generic "attention pass" kernel, four modes.  A CTA owns a 128-row tile (TMEM lanes = rows) of one
// (batch, head) and walks the other sequence dimension in blocks of 64 columns:
//     X = A1 . B1^T   (128 x 64, K = d_k)         [Y = A2 . B2^T   in the two-score modes]
//     Z = elementwise(X [,Y])                     (row owner threads, registers)
//     OUT_blk = Z . B3 (128 x d_k, K = 64)        folded into a register-resident running total (RN adds)
//   mode   lanes   A1  A2   B1  B2   B3       elementwise                                   output
//   FWD    q rows  Q   -    K   -    V        online softmax (running max / sum, rescale)   O, LSE (log2 domain)
//   DV     kv rows K   -    Q   -    dO       P^T = exp2(X c2 - lse[col])                   dV
//   DQ     q rows  Q   dO   K   V    K        Z = P (Y - D[row]) c                          dQ
//   DK     kv rows K   V    Q   dO   Q        Z = P^T (Y - D[col]) c                        dK
// All contractions are fp32-faithful 3xTF32 (A = Ah + Al, B = Bh + Bl; Al.Bh + Ah.Bl + Ah.Bh), every MMA is
// TS-mode: the A operand (the 128-row tile, and Z) lives in TMEM as hi/lo column groups written by threads, so
// a 128 x 64 x 8 tf32 MMA costs 32 cycles (an A operand in shared memory would cost 128, measured in gemm_tc.cu).
// The B operands stream through a 7-slot ring of 32 KB shared-memory slots (raw tile from TMA + lo tile from
// the converter warps), in exactly the order the MMA warp consumes them.
//
// Warp roles (512 threads = 4 warpgroups, one CTA per SM, persistent over work items; setmaxnreg 80 / 112 / 152 / 152):
//   warp 0      TMA producer            warp 1      MMA issuer (one elected lane)  warps 2-3 idle
//   warps 4-7   converter: A tiles smem -> TMEM (hi/lo), B tiles raw -> lo (in smem)
//   warps 8-15  row owners (two groups): a thread owns one row and half of the columns: TMEM -> registers ->
//               softmax / dS -> TMEM (Z hi/lo), fold of its half of OUT_blk, epilogue
// TMEM (512 columns):  A1 hi/lo [0,128)  |  one-score modes: X0/Zhi0 L0 X1/Zhi1 L1 [128,384) OUT [384,448)
//                      (one-score modes: a second OUT accumulator at [448,512))
//                      A2 hi/lo [128,256) | two-score modes: X0/Zlo0 X1/Zlo1 [256,384) Y/Zhi [384,448) OUT [448,512)
#include "common.cuh"
#include "tcgen05.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <atomic>

namespace b200ad {

constexpr int AT_DK = 64;                    // head dimension this kernel is built for
constexpr int AT_BM = 128;                   // rows (TMEM lanes) per work item
constexpr int AT_BC = 64;                    // columns per block
constexpr int AT_NSLOT = 6;                  // ring slots; a 7th 32 KB region stages the output tile for the TMA store
constexpr uint32_t AT_SLOT = 32768;          // [raw 16 KB | lo 16 KB] for a B block, or one raw A tile
constexpr int AT_THREADS = 512;              // 4 warpgroups: {TMA, MMA, 2 idle}, converter, 2 x row owners
constexpr uint32_t AT_SMEM = (AT_NSLOT + 1) * AT_SLOT + 1024 /*align*/;   // + ~1.3 KB static (barriers, column / row statistics)

enum { AT_FWD = 0, AT_DV = 1, AT_DQ = 2, AT_DKM = 3 };

struct AttnParams {
  float* out;                // FWD: O; DV: dV; DQ: dQ; DK: dK    element (b, h, row, d) at out + b*bs + h*hs + row*ld + d
  int64_t out_bs, out_hs, out_ld;
  float* lse_w;              // FWD: (B, H, R) log2-domain log-sum-exp, written
  const float* lse;          // other modes: the same array, read ((B, H, Sq))
  const float* dvec;         // DQ / DK: D[b, h, s] = sum_d dO * O
  int32_t H, R, C;           // heads; extent of the lane dimension; extent of the column dimension
  int32_t row_tiles, nb, total_items;
  float c2, c;               // scale * log2(e), scale
  float* amax_w;             // FWD, optional: (B*H, 4) table, slots 0..2 receive max |Q|, |K|, |V| of the (batch, head) slab
                             // (atomic max on the bit patterns; zeroed by the launcher) — handed to the fp16-split backward,
                             // which otherwise re-reads the three tensors for them
  long long* prof;           // -DAT_PROFILE: per-role cycle counters of block 0 (debug builds only)
  int32_t probe;             // diagnostic bits (results are garbage, only the rate matters): 1 = converter skips the B lo
                             // tiles, 2 = row owners skip the elementwise math and the Z stores, 4 = no TMA loads
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#ifdef AT_PROFILE
#define AT_T(var) const long long var = clock64()
#define AT_ACC(slot, a, b) prof_acc[slot] += (b) - (a)
#else
#define AT_T(var)
#define AT_ACC(slot, a, b)
#endif
__device__ __forceinline__ uint32_t tf32_hi(uint32_t v) { return v & 0xffffe000u; }
__device__ __forceinline__ uint32_t tf32_lo(uint32_t v, uint32_t h) {
  return __float_as_uint(__uint_as_float(v) - __uint_as_float(h)) & 0xffffe000u;
}

__device__ __forceinline__ float fast_exp2(float x) {      // one MUFU; results below 2^-126 flush to zero
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// CP = CTA pair (cluster of two CTAs, cta_group::2 MMAs with M = 256): CTA r owns row tile 2 * unit + r and stages HALF of
// every B block (32 of the 64 score columns / 32 of the 64 output columns), so the MMA-side shared-memory reads, the TMA
// fill and the converter work per CTA are halved: the 1-CTA passes are bound by exactly that traffic (~125 B/clk at full
// MMA rate).  All cross-role barriers that the MMA warp waits on live in the leader CTA (remote arrives), everything the
// MMA warp signals is a multicast commit.
template <int MODE, bool CP>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_a2,
               const __grid_constant__ CUtensorMap map_b1, const __grid_constant__ CUtensorMap map_b2,
               const __grid_constant__ CUtensorMap map_b3, const __grid_constant__ CUtensorMap map_out,
               const AttnParams p) {
  constexpr bool HAS2 = (MODE == AT_DQ || MODE == AT_DKM);
  constexpr bool ROWSTAT = (MODE == AT_FWD || MODE == AT_DQ);   // lanes are query rows
  constexpr bool PAIR = (MODE != AT_FWD);                       // two column blocks per OUT accumulator chain
  constexpr uint32_t A1_COL = 0, A2_COL = 128;
  constexpr uint32_t X_COL0 = HAS2 ? 256u : 128u;
  constexpr uint32_t X_STRIDE = HAS2 ? 64u : 128u;
  constexpr uint32_t Y_COL = 384;
  constexpr uint32_t OUT_COL = HAS2 ? 448u : 384u;
  constexpr uint32_t TMEM_COLS = 512;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* out_stage = smem + (size_t)AT_NSLOT * AT_SLOT;   // 32 KB: the output tile, two SWIZZLE_128B halves
  __shared__ uint64_t bars[3 * AT_NSLOT + 14];
  __shared__ uint32_t tmem_base_smem;
  // DV / DK: [buffer][lse (64) | D (64)] of the current column block;  FWD: [owner group][row] row-maximum / row-sum
  // exchange between the two owners of a row (the modes never need both, so they share the 1 KB)
  __shared__ __align__(16) float colstat[2][128];
  float (*rowx)[128] = colstat;
  uint64_t* full = bars;                       // [NSLOT] TMA -> converter
  uint64_t* conv = bars + AT_NSLOT;            // [NSLOT] converter -> MMA (128 arrivals): lo tile in smem / A tile in TMEM
  uint64_t* empty = bars + 2 * AT_NSLOT;       // [NSLOT] MMA commit (B slots) / converter (A slots) -> TMA
  uint64_t* a_empty = bars + 3 * AT_NSLOT;     // [2] MMA commit -> converter: last score MMA of the item retired
  uint64_t* x_full = a_empty + 2;              // [2] MMA commit -> row owners
  uint64_t* z_full = x_full + 2;               // [2] row owners -> MMA (128 arrivals)
  uint64_t* y_full = z_full + 2;               // [1]
  uint64_t* out_full = y_full + 1;             // [2] MMA commit -> row owners (one-score modes double-buffer the OUT accumulator)
  uint64_t* out_empty = out_full + 2;          // [2] row owners -> MMA
  uint32_t* tmem_base_slot = &tmem_base_smem;
  // One-score modes have 64 spare TMEM columns: a second OUT accumulator, so OUT(j) never waits for the fold of
  // OUT(j-1) (that wait put two mbarrier wake-ups and a TMEM read on the MMA warp's critical path every block).
  constexpr uint32_t OUT_BUFS = HAS2 ? 1u : 2u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CP ? cluster_ctarank() : 0u;
  const int unit_id = CP ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, units = CP ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // signalling helpers: barriers the MMA warp waits on live in the leader CTA; its own signals reach both CTAs
  // one arrival per WARP in pair mode (cluster-scope release arrives are expensive: ~500 cycles each when all 32 lanes
  // issue them); every lane has completed its own tcgen05.wait::st / fence before the __syncwarp
  auto arrive_leader = [&](uint64_t* bar) {
    if (CP) {
      __syncwarp();
      if (lane == 0) { if (rank == 0) mbar_arrive(bar); else mbar_arrive_cluster(bar, 0); }
    } else {
      mbar_arrive(bar);
    }
  };
  auto mma_commit = [&](uint64_t* bar) { if (CP) umma2_commit_mc(bar); else umma_commit(bar); };
  auto mma_ts = [&](uint32_t d, uint32_t a, uint64_t desc, uint32_t idesc, uint32_t acc) {
    if (CP) umma2_tf32_ts(d, a, desc, idesc, acc); else umma_tf32_ts(d, a, desc, idesc, acc);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a1);
    tma_prefetch_desc(&map_b1);
    tma_prefetch_desc(&map_b3);
    tma_prefetch_desc(&map_out);
    if (HAS2) {
      tma_prefetch_desc(&map_a2);
      tma_prefetch_desc(&map_b2);
    }
    for (int s = 0; s < AT_NSLOT; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&conv[s], CP ? 8 : 128);          // pair mode: one arrival per converter warp of both CTAs
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_empty[i], 1);
      mbar_init(&x_full[i], 1);
      mbar_init(&z_full[i], CP ? 16 : 256);       // pair mode: one arrival per owner warp of both CTAs
    }
    mbar_init(y_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&out_full[i], 1);
      mbar_init(&out_empty[i], CP ? 16 : 256);
    }
    fence_barrier_init();
  }
  if (warp == 1) { if (CP) tmem_alloc2(tmem_base_slot, TMEM_COLS); else tmem_alloc(tmem_base_slot, TMEM_COLS); }
  tc_fence_before();
  if (CP) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const int nb = p.nb;

  // register budgets are set INSIDE each warpgroup's branch so that ptxas allocates per role
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      uint32_t op = 0;
      for (int item = unit_id; item < p.total_items; item += units) {
        const int rt = (item % p.row_tiles) * (CP ? 2 : 1) + (int)rank, bh = item / p.row_tiles;
        const int h = bh % p.H, b = bh / p.H;
        const int row0 = rt * AT_BM;
        auto acquire = [&](uint32_t bytes, uint64_t*& bar) -> uint8_t* {
          const uint32_t s = op % AT_NSLOT;
          mbar_wait(&empty[s], ((op / AT_NSLOT) & 1) ^ 1);
          if (p.probe & 4) mbar_arrive(&full[s]); else mbar_expect_tx(&full[s], bytes);
          bar = &full[s];
          ++op;
          return smem + (size_t)s * AT_SLOT;
        };
        auto load_a = [&](const CUtensorMap* m) {          // 128 rows x 64: two 32-float halves, SWIZZLE_128B
          uint64_t* bar;
          uint8_t* dst = acquire(32768u, bar);
          if (p.probe & 4) return;
          tma_load_4d(m, bar, dst, 0, row0, h, b);
          tma_load_4d(m, bar, dst + 16384, 32, row0, h, b);
        };
        auto load_k = [&](const CUtensorMap* m, int col0) {   // K-major B block: 64 rows x 64 (pair: this CTA's 32 rows)
          uint64_t* bar;
          if (p.probe & 4) { acquire(0u, bar); return; }
          if (CP) {
            uint8_t* dst = acquire(8192u, bar);
            tma_load_4d(m, bar, dst, 0, col0 + 32 * (int)rank, h, b);
            tma_load_4d(m, bar, dst + 4096, 32, col0 + 32 * (int)rank, h, b);
          } else {
            uint8_t* dst = acquire(16384u, bar);
            tma_load_4d(m, bar, dst, 0, col0, h, b);
            tma_load_4d(m, bar, dst + 8192, 32, col0, h, b);
          }
        };
        auto load_mn = [&](const CUtensorMap* m, int col0) {  // MN-major B block: 64 k-rows x 64 (32 x 32 boxes)
          uint64_t* bar;
          if (p.probe & 4) { acquire(0u, bar); return; }
          if (CP) {                                           // pair: this CTA's 32 output columns (d), all 64 k-rows
            uint8_t* dst = acquire(8192u, bar);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) tma_load_4d(m, bar, dst + kb * 4096, 32 * (int)rank, col0 + kb * 32, h, b);
          } else {
            uint8_t* dst = acquire(16384u, bar);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
              for (int c = 0; c < 2; ++c) tma_load_4d(m, bar, dst + kb * 8192 + c * 4096, c * 32, col0 + kb * 32, h, b);
          }
        };
        load_a(&map_a1);
        if (HAS2) load_a(&map_a2);
        load_k(&map_b1, 0);
        if (HAS2) load_k(&map_b2, 0);
        for (int j = 1; j <= nb; ++j) {
          if (j < nb) load_k(&map_b1, j * AT_BC);
          load_mn(&map_b3, (j - 1) * AT_BC);
          if (HAS2 && j < nb) load_k(&map_b2, j * AT_BC);
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================================== MMA issuer (leader CTA of a pair) =====================================
    // instruction descriptor: D = f32, A = B = tf32, A K-major (TMEM), N = 64, M = 128; bit 16 = B MN-major
    constexpr uint32_t IDESC_K = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) |
                                 ((uint32_t)((CP ? 256 : AT_BM) >> 4) << 24);
    constexpr uint32_t IDESC_MN = IDESC_K | (1u << 16);
    constexpr uint32_t LO_OFF = CP ? 8192u : 16384u, KHALF = CP ? 4096u : 8192u;   // slot layout: [raw | lo], two k halves each
    uint32_t op = 0, it = 0, gb = 0, fo = 0;    // ring op, item, global block and fold (OUT accumulator use) counters
#ifdef AT_PROFILE
    long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long prof_t0 = clock64();
#endif
    // Measured (profiles/mma_rate.cu): every mbarrier wait / tcgen05.fence of the issuing thread travels through the same
    // in-order queue as the MMAs; with 32-cycle N = 64 MMAs the queue drains during each such round trip (~250-340
    // cycles).  So the role has ONE synchronisation point per block: all waits of the block's MMA groups first
    // (Z, the OUT accumulator, and the operand slots of OUT(j-1), Y(j) and X(j+1)), one fence, then every MMA of the
    // block back to back inside one elected region.  The ring order (= consumption order) is unchanged.
    auto wait_slot = [&]() -> uint32_t {                       // next ring slot: converted operand (or A tile) is in place
      const uint32_t s = op % AT_NSLOT;
      AT_T(w0);
      mbar_wait(&conv[s], (op / AT_NSLOT) & 1);
      AT_T(w1);
      AT_ACC(0, w0, w1);
      ++op;
      return s;
    };
    // score MMAs: D[128 x 64] = A(TMEM, hi at a_col, lo at a_col + 64) . B^T (K-major slot: two k halves); ELECTED thread
    auto issue_score = [&](uint32_t s, uint32_t a_col, uint32_t d_col, uint64_t* done, uint64_t* a_done) {
      const uint64_t d0 = make_smem_desc(smem_u32(smem + (size_t)s * AT_SLOT), 16, 1024, 2);
      const uint32_t dst = tmem_base + d_col, a_hi = tmem_base + a_col, a_lo = a_hi + 64u;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          mma_ts(dst, (t == 0 ? a_lo : a_hi) + kk * 8,
                 d0 + (uint64_t)(((t == 1 ? LO_OFF : 0u) + (kk >> 2) * KHALF + (kk & 3) * 32u) >> 4), IDESC_K, (t | kk) ? 1u : 0u);
      }
      mma_commit(&empty[s]);
      mma_commit(done);
      if (a_done) mma_commit(a_done);
    };
    for (int item = unit_id; item < p.total_items; item += units, ++it) {
      // prologue of the item: X(0) [, Y(0)] and X(1); the A tiles' ring slots are consumed by the converter, which
      // signals "hi/lo in TMEM" on the slot's conv barrier
      {
        const uint32_t s_a1 = wait_slot();
        (void)s_a1;
        uint32_t s_a2 = 0;
        if (HAS2) s_a2 = wait_slot();
        (void)s_a2;
        const uint32_t s_x0 = wait_slot();
        uint32_t s_y0 = 0, s_x1 = 0;
        if (HAS2) s_y0 = wait_slot();
        if (nb > 1) s_x1 = wait_slot();
        tc_fence_after();
        if (elect_one()) {
          issue_score(s_x0, A1_COL, X_COL0 + (gb & 1) * X_STRIDE, &x_full[gb & 1], nb == 1 ? &a_empty[0] : nullptr);
          if (HAS2) issue_score(s_y0, A2_COL, Y_COL, y_full, nb == 1 ? &a_empty[1] : nullptr);
          if (nb > 1) issue_score(s_x1, A1_COL, X_COL0 + ((gb + 1) & 1) * X_STRIDE, &x_full[(gb + 1) & 1], nb == 2 ? &a_empty[0] : nullptr);
        }
        __syncwarp();
      }
      for (int j = 1; j <= nb; ++j) {
        // block j-1 closes: OUT_blk (+)= Z(g) . B3(g) with Z hi/lo from TMEM and B3 MN-major from its slot; the backward
        // modes chain two blocks (128 K-elements) per accumulator before the row owners fold it; in the two-score
        // modes the lo term (which reads the Y columns) goes first so that Y(j) can overwrite them while the hi terms
        // run.  X(j+1) follows in the same issue burst.
        const uint32_t g = gb + j - 1;
        const uint32_t xb = X_COL0 + (g & 1) * X_STRIDE;
        const uint32_t zhi = xb, zlo = HAS2 ? Y_COL : xb + 64u;
        const bool first = !PAIR || ((j - 1) & 1) == 0;
        const bool last = !PAIR || ((j - 1) & 1) == 1 || j == nb;
        const uint32_t ob = fo % OUT_BUFS, oph = (fo / OUT_BUFS) & 1;   // OUT accumulator buffer and its phase
        AT_T(w0);
        mbar_wait(&z_full[g & 1], (g >> 1) & 1);
        AT_T(w1);
        if (first) mbar_wait(&out_empty[ob], oph ^ 1);
        AT_T(w2);
        AT_ACC(1, w0, w1);
        AT_ACC(2, w1, w2);
        const uint32_t s_o = wait_slot();                                // B3(j-1)
        uint32_t s_y = 0, s_x = 0;
        if (HAS2 && j < nb) s_y = wait_slot();                           // B2(j)
        if (j + 1 < nb) s_x = wait_slot();                               // B1(j+1)
        tc_fence_after();
        if (elect_one()) {
          const uint64_t d0 = make_smem_desc(smem_u32(smem + (size_t)s_o * AT_SLOT), 4096, 512, 1);
          const uint32_t dst = tmem_base + OUT_COL + ob * 64u, a_hi = tmem_base + zhi, a_lo = tmem_base + zlo;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            mma_ts(dst, a_lo + kk * 8, d0 + (uint64_t)(((kk >> 2) * KHALF + (kk & 3) * 1024u) >> 4), IDESC_MN,
                   (kk || !first) ? 1u : 0u);
          if (HAS2 && j < nb) issue_score(s_y, A2_COL, Y_COL, y_full, j == nb - 1 ? &a_empty[1] : nullptr);
#pragma unroll
          for (int t = 1; t < 3; ++t) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              mma_ts(dst, a_hi + kk * 8,
                     d0 + (uint64_t)(((t == 1 ? LO_OFF : 0u) + (kk >> 2) * KHALF + (kk & 3) * 1024u) >> 4), IDESC_MN, 1u);
          }
          mma_commit(&empty[s_o]);
          if (last) mma_commit(&out_full[ob]);
          if (j + 1 < nb) {
            const uint32_t gx = gb + j + 1;
            issue_score(s_x, A1_COL, X_COL0 + (gx & 1) * X_STRIDE, &x_full[gx & 1], j + 1 == nb - 1 ? &a_empty[0] : nullptr);
          }
        }
        __syncwarp();
        if (last) ++fo;
      }
      gb += nb;
    }
#ifdef AT_PROFILE
    if (p.prof && blockIdx.x == 0 && lane == 0) {
      p.prof[0] = clock64() - prof_t0;   // MMA warp total
      p.prof[1] = prof_acc[0];           // waiting for converted operand slots
      p.prof[2] = prof_acc[1];           // waiting for Z (row owners)
      p.prof[3] = prof_acc[2];           // waiting for the OUT buffer to be drained
    }
#endif
  }
  } else if (warp < 8) {
    // ===================================== converter =====================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
    const int tid = threadIdx.x - 128;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t op = 0, it = 0;
    const bool track = (MODE == AT_FWD) && p.amax_w != nullptr;
    uint32_t mx_a = 0, mx_k = 0, mx_v = 0;        // running max |x| bit patterns of this thread's Q, K, V elements
    auto conv_a = [&](int which) {
      const uint32_t s = op % AT_NSLOT;
      mbar_wait(&full[s], (op / AT_NSLOT) & 1);
      ++op;
      mbar_wait(&a_empty[which], (it & 1) ^ 1);
      tc_fence_after();
      const uint32_t slot = smem_u32(smem + (size_t)s * AT_SLOT);
      const uint32_t col = which ? A2_COL : A1_COL;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const uint32_t row = slot + hh * 16384 + r * 128;
        uint32_t v[32], l[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {     // SWIZZLE_128B: 16 B chunk c of row r sits at chunk c ^ (r & 7)
          const uint4 t4 = lds128(row + ((c ^ (r & 7)) << 4));
          v[4 * c] = t4.x; v[4 * c + 1] = t4.y; v[4 * c + 2] = t4.z; v[4 * c + 3] = t4.w;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (MODE == AT_FWD) mx_a = max(mx_a, v[j] & 0x7fffffffu);
          const uint32_t hi = tf32_hi(v[j]);
          l[j] = tf32_lo(v[j], hi);
          v[j] = hi;
        }
        tmem_st_32x32b_x32(lane_base + col + hh * 32, v);
        tmem_st_32x32b_x32(lane_base + col + 64 + hh * 32, l);
      }
      tmem_st_wait();
      tc_fence_before();
      arrive_leader(&conv[s]);            // A tile (hi/lo) is in TMEM; keeps the slot's conv phase in step with its uses
      named_bar_sync(1, 128);             // every converter thread has read its row: the slot can be refilled
      if (tid == 0) mbar_arrive(&empty[s]);
    };
    auto conv_b = [&](uint32_t& mx) {
      const uint32_t s = op % AT_NSLOT;
      mbar_wait(&full[s], (op / AT_NSLOT) & 1);
      ++op;
      const uint32_t raw = smem_u32(smem + (size_t)s * AT_SLOT) + tid * 16;
      const int n_vec = (p.probe & 1) ? 0 : (CP ? 4 : 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) {                           // pair: 8 KB raw -> 8 KB lo per CTA
        if (i >= n_vec) break;
        const uint4 v = lds128(raw + i * 2048);
        if (MODE == AT_FWD) mx = max(max(mx, max(v.x & 0x7fffffffu, v.y & 0x7fffffffu)), max(v.z & 0x7fffffffu, v.w & 0x7fffffffu));
        uint4 l;
        l.x = tf32_lo(v.x, tf32_hi(v.x));
        l.y = tf32_lo(v.y, tf32_hi(v.y));
        l.z = tf32_lo(v.z, tf32_hi(v.z));
        l.w = tf32_lo(v.w, tf32_hi(v.w));
        sts128(raw + (CP ? 8192 : 16384) + i * 2048, l);
      }
      fence_proxy_async();                // generic-proxy writes -> visible to the tensor core (async proxy)
      arrive_leader(&conv[s]);
    };
    const int n_b = (HAS2 ? 3 : 2) * nb;
    for (int item = unit_id; item < p.total_items; item += units, ++it) {
      conv_a(0);
      if (HAS2) conv_a(1);
      // FWD ring order of the B blocks: K(0), then { K(j), V(j-1) } for j = 1 .. nb-1, then V(nb-1)
      for (int i = 0; i < n_b; ++i) conv_b((MODE != AT_FWD || i == 0 || (i < n_b - 1 && (i & 1))) ? mx_k : mx_v);
      if (track) {      // one atomic per warp, slot and item ((batch, head) changes between items)
        const int bh = item / p.row_tiles;
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) {
          mx_a = max(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, ofs));
          mx_k = max(mx_k, __shfl_xor_sync(0xffffffffu, mx_k, ofs));
          mx_v = max(mx_v, __shfl_xor_sync(0xffffffffu, mx_v, ofs));
        }
        if (lane == 0) {
          unsigned int* t = reinterpret_cast<unsigned int*>(p.amax_w + (int64_t)bh * 4);
          if (mx_a) atomicMax(t + 0, mx_a);
          if (mx_k) atomicMax(t + 1, mx_k);
          if (mx_v) atomicMax(t + 2, mx_v);
        }
      }
      mx_a = mx_k = mx_v = 0;
    }
  } else {
    // ===================================== row owners =====================================
    // Two warpgroups; a thread owns one row (TMEM lane) and HALF of the columns: group gi handles columns
    // [32 gi, 32 gi + 32) of every 64-column score block and of the d_k = 64 output.  Halving the per-thread work
    // halves the elementwise latency on the MMA <-> owner critical path; only the forward's row maximum (and the
    // final row sum) need an exchange between the two owners of a row, through shared memory.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int tid = threadIdx.x - 256;                   // 0..255
    const int gi = (warp - 8) >> 2;                      // column half
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t cofs = (uint32_t)gi * 32u;            // this group's column offset inside a 64-column region
    uint32_t gb = 0, fo = 0;
#ifdef AT_PROFILE
    long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long prof_t0 = clock64();
#endif
    for (int item = unit_id; item < p.total_items; item += units) {
      const int rt = (item % p.row_tiles) * (CP ? 2 : 1) + (int)rank, bh = item / p.row_tiles;
      const int h = bh % p.H, b = bh / p.H;
      const int grow = rt * AT_BM + r;
      const bool row_ok = grow < p.R;
      float run[32];
#pragma unroll
      for (int d = 0; d < 32; ++d) run[d] = 0.f;
      float m_run = -INFINITY, l_run = 0.f, a_prev = 0.f;     // FWD (l_run: this group's share of the row sum)
      float ls_r = 0.f, d_r = 0.f;                             // DQ: per-row statistics
      if (MODE == AT_DQ && row_ok) {
        ls_r = __ldg(p.lse + (int64_t)bh * p.R + grow);
        d_r = __ldg(p.dvec + (int64_t)bh * p.R + grow) * p.c;      // pre-scaled: Z = P (Y c - D c)
      }
      // fold this group's half of OUT_blk into the running total: run = run * a + OUT
      auto fold = [&](float a) {
        AT_T(f0);
        const uint32_t ob = fo % OUT_BUFS, oph = (fo / OUT_BUFS) & 1;
        mbar_wait(&out_full[ob], oph);
        ++fo;
        AT_T(f1);
        AT_ACC(1, f0, f1);
        tc_fence_after();
        uint32_t o[32];
        tmem_ld_32x32b_x32(lane_base + OUT_COL + ob * 64u + cofs, o);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 32; ++d) {
          if (MODE == AT_FWD) run[d] = fmaf(run[d], a, __uint_as_float(o[d]));
          else run[d] += __uint_as_float(o[d]);
        }
        tc_fence_before();
        arrive_leader(&out_empty[ob]);
        AT_T(f2);
        AT_ACC(2, f1, f2);
      };
      auto load_colstat = [&](int c0) -> float {
        if (tid < 64) {
          const int col = c0 + tid;
          return (col < p.C) ? __ldg(p.lse + (int64_t)bh * p.C + col) : INFINITY;
        }
        if (MODE == AT_DKM && tid < 128) {
          const int col = c0 + tid - 64;
          return (col < p.C) ? __ldg(p.dvec + (int64_t)bh * p.C + col) * p.c : 0.f;   // pre-scaled: Z = P (Y c - D c)
        }
        return 0.f;
      };
      float cs_next = ROWSTAT ? 0.f : load_colstat(0);
      for (int j = 0; j < nb; ++j) {
        const uint32_t g = gb + j;
        const uint32_t bsel = g & 1;
        const uint32_t xaddr = lane_base + X_COL0 + bsel * X_STRIDE + cofs;
        const int col0 = j * AT_BC + (int)cofs;          // first global column of this group's half
        float* cs = colstat[bsel];
        if (!ROWSTAT) {
          // per-column statistics of this block (lse, and D in the DK mode) -> shared memory; the next block's
          // values are fetched now so that the global-load latency hides behind this block's work
          if (tid < 128) cs[tid] = cs_next;
          named_bar_sync(2, 256);
          if (j + 1 < nb) cs_next = load_colstat((j + 1) * AT_BC);
        }
        AT_T(x0);
        mbar_wait(&x_full[bsel], (g >> 1) & 1);
        AT_T(x1);
        AT_ACC(0, x0, x1);
        tc_fence_after();
        if (p.probe & 2) {
          if (HAS2) mbar_wait(y_full, g & 1);
          tc_fence_before();
          arrive_leader(&z_full[bsel]);
          if (MODE == AT_FWD ? (j >= 1) : (j >= 1 && ((j - 1) & 1))) fold(1.f);
        } else if (MODE == AT_FWD) {
          uint32_t x[32], lo[32];
          tmem_ld_32x32b_x32(xaddr, x);
          tmem_ld_wait();
          AT_T(x2);
          AT_ACC(3, x1, x2);
          if (col0 + 32 > p.C) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (col0 + c >= p.C) x[c] = __float_as_uint(-INFINITY);
          }
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int c = 0; c < 32; ++c) mx[c & 3] = fmaxf(mx[c & 3], __uint_as_float(x[c]));
          float mb = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          // the row maximum over all 64 columns: exchange with the other owner of this row
          named_bar_sync(5, 256);                    // the previous block's exchange has been read by both groups
          rowx[gi][r] = mb;
          named_bar_sync(4, 256);
          mb = fmaxf(mb, rowx[gi ^ 1][r]);
          const float m_new = fmaxf(m_run, mb);
          const float alpha = fast_exp2((m_run - m_new) * p.c2);
          const float mc = m_new * p.c2;
          float sm[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float pv = fast_exp2(fmaf(__uint_as_float(x[c]), p.c2, -mc));
            sm[c & 3] += pv;
            const uint32_t hi = tf32_hi(__float_as_uint(pv));
            lo[c] = tf32_lo(__float_as_uint(pv), hi);
            x[c] = hi;
          }
          tmem_st_32x32b_x32(xaddr, x);
          tmem_st_32x32b_x32(xaddr + 64, lo);
          l_run = fmaf(l_run, alpha, (sm[0] + sm[1]) + (sm[2] + sm[3]));
          m_run = m_new;
          AT_T(x3);
          AT_ACC(4, x2, x3);
          tmem_st_wait();
          tc_fence_before();
          arrive_leader(&z_full[bsel]);
          AT_T(x4);
          AT_ACC(5, x3, x4);
          if (j >= 1) fold(a_prev);
          a_prev = alpha;
        } else if (MODE == AT_DV) {
          uint32_t x[32], lo[32];
          tmem_ld_32x32b_x32(xaddr, x);
          tmem_ld_wait();
#pragma unroll
          for (int c4 = 0; c4 < 32; c4 += 4) {
            const float4 l4 = *reinterpret_cast<const float4*>(cs + cofs + c4);
            const float ls[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c = c4 + i;
              const float pv = fast_exp2(fmaf(__uint_as_float(x[c]), p.c2, -ls[i]));
              const uint32_t hi = tf32_hi(__float_as_uint(pv));
              lo[c] = tf32_lo(__float_as_uint(pv), hi);
              x[c] = hi;
            }
          }
          tmem_st_32x32b_x32(xaddr, x);
          tmem_st_32x32b_x32(xaddr + 64, lo);
          tmem_st_wait();
          tc_fence_before();
          arrive_leader(&z_full[bsel]);
          if (j >= 1 && ((j - 1) & 1)) fold(1.f);      // block j-1 closed a two-block chain
        } else {
          // DQ / DK: Z = P (Y - D) c;   Z hi -> X columns, Z lo -> Y columns
          mbar_wait(y_full, g & 1);
          tc_fence_after();
          uint32_t x[32], y[32];
          tmem_ld_32x32b_x32(xaddr, x);
          tmem_ld_32x32b_x32(lane_base + Y_COL + cofs, y);
          tmem_ld_wait();
#pragma unroll
          for (int c4 = 0; c4 < 32; c4 += 4) {
            float ls[4], dd[4];
            if (MODE == AT_DQ) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                ls[i] = (col0 + c4 + i < p.C) ? ls_r : INFINITY;
                dd[i] = d_r;
              }
            } else {                                  // per-column statistics: two 16-byte shared-memory reads per 4 columns
              const float4 l4 = *reinterpret_cast<const float4*>(cs + cofs + c4);
              const float4 d4 = *reinterpret_cast<const float4*>(cs + 64 + cofs + c4);
              ls[0] = l4.x; ls[1] = l4.y; ls[2] = l4.z; ls[3] = l4.w;
              dd[0] = d4.x; dd[1] = d4.y; dd[2] = d4.z; dd[3] = d4.w;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c = c4 + i;
              const float pv = fast_exp2(fmaf(__uint_as_float(x[c]), p.c2, -ls[i]));
              const float z = pv * fmaf(__uint_as_float(y[c]), p.c, -dd[i]);
              const uint32_t hi = tf32_hi(__float_as_uint(z));
              y[c] = tf32_lo(__float_as_uint(z), hi);
              x[c] = hi;
            }
          }
          tmem_st_32x32b_x32(xaddr, x);
          tmem_st_32x32b_x32(lane_base + Y_COL + cofs, y);
          tmem_st_wait();
          tc_fence_before();
          arrive_leader(&z_full[bsel]);
          if (j >= 1 && ((j - 1) & 1)) fold(1.f);      // block j-1 closed a two-block chain
        }
      }
      fold(a_prev);
      gb += nb;
      AT_T(e0);
      // epilogue: stage the tile in shared memory (two SWIZZLE_128B halves = the two groups' column halves,
      // conflict-free row writes), one thread issues the TMA stores (coalesced, asynchronous, rows beyond R are
      // clipped by the tensor map)
      if (tid == 0) bulk_wait_read0();          // the previous item's store has finished reading the staging tile
      float sc = 1.f;
      if (MODE == AT_FWD) {
        named_bar_sync(5, 256);                 // the last block's exchange has been read by both groups
        rowx[gi][r] = l_run;                    // row sum = both groups' shares
      }
      named_bar_sync(3, 256);
      if (MODE == AT_FWD) {
        const float l_tot = l_run + rowx[gi ^ 1][r];
        sc = 1.f / l_tot;
        if (gi == 0 && row_ok) p.lse_w[(int64_t)bh * p.R + grow] = fmaf(m_run, p.c2, log2f(l_tot));
      }
      {
        const uint32_t srow = smem_u32(out_stage) + (uint32_t)gi * 16384u + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 v;
          v.x = __float_as_uint(run[4 * c] * sc); v.y = __float_as_uint(run[4 * c + 1] * sc);
          v.z = __float_as_uint(run[4 * c + 2] * sc); v.w = __float_as_uint(run[4 * c + 3] * sc);
          sts128(srow + ((c ^ (r & 7)) << 4), v);
        }
      }
      fence_proxy_async();
      named_bar_sync(3, 256);
      if (tid == 0) {
        tma_store_4d(&map_out, out_stage, 0, rt * AT_BM, h, b);
        tma_store_4d(&map_out, out_stage + 16384, 32, rt * AT_BM, h, b);
        bulk_commit();
      }
      AT_T(e1);
      AT_ACC(6, e0, e1);
    }
    if (tid == 0) bulk_wait0();                 // all output tiles have been written
#ifdef AT_PROFILE
    if (p.prof && blockIdx.x == 0 && threadIdx.x == 256) {
      p.prof[8] = clock64() - prof_t0;   // row owner total
      p.prof[9] = prof_acc[0];           // waiting for X (score MMA)
      p.prof[10] = prof_acc[1];          // waiting for OUT (second MMA)
      p.prof[11] = prof_acc[2];          // fold (TMEM load + FMAs)
      p.prof[12] = prof_acc[3];          // FWD: TMEM load of X
      p.prof[13] = prof_acc[4];          // FWD: softmax arithmetic + TMEM store issue
      p.prof[14] = prof_acc[5];          // FWD: wait::st + arrive
      p.prof[15] = prof_acc[6];          // epilogue stores
    }
#endif
  }

  tc_fence_before();
  if (CP) cluster_sync_all(); else __syncthreads();     // the peer's smem / barriers / TMEM stay valid until both CTAs are done
  tc_fence_after();
  if (warp == 1) { if (CP) tmem_dealloc2(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS); }
}

// D[b, h, s] = sum_d dO[b, s, h, d] * O[b, s, h, d]: 16 lanes (one float4 each) per (b, h, s) row of 64
__global__ void __launch_bounds__(256) attn_dvec_kernel(const float* __restrict__ o, const float* __restrict__ dout,
                                                        float* __restrict__ dvec, int64_t rows_total, int H, int S,
                                                        int64_t bs, int64_t hs, int64_t ld, int dk) {
  const int sub = threadIdx.x & 15;
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 4);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4); row < rows_total; row += stride) {
    // row enumerates (b, h, s) with s fastest: matches the (B, H, S) layout of dvec / lse
    const int64_t s = row % S, bh = row / S;
    const int64_t h = bh % H, b = bh / H;
    const int64_t off = b * bs + h * hs + s * ld + sub * 4;
    float v = 0.f;
    if (sub * 4 < dk) {
      const float4 a = *reinterpret_cast<const float4*>(o + off);
      const float4 g = *reinterpret_cast<const float4*>(dout + off);
      v = a.x * g.x + a.y * g.y + a.z * g.z + a.w * g.w;
    }
#pragma unroll
    for (int ofs = 8; ofs > 0; ofs >>= 1) v += __shfl_xor_sync(0xffffffffu, v, ofs, 16);
    if (sub == 0) dvec[row] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
struct AttnTensor {
  const float* ptr;
  int64_t bs, hs, ld;   // element strides: batch, head, row
  int64_t rows;         // sequence length of this tensor
  int64_t dk = AT_DK;   // head dimension (<= 64, multiple of 4): the tensor map's inner extent; columns dk .. 63 of every
                        // 64-wide tile are TMA zero fill on the way in and clipped on the way out
};

static int encode_attn(CUtensorMap* map, const AttnTensor& t, int64_t B, int64_t H, uint32_t box_rows, bool mn_major) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(B200AD_ERR_CUDA, "attention(tcgen05): cuTensorMapEncodeTiled is not available");
  cuuint64_t dims[4] = {(cuuint64_t)t.dk, (cuuint64_t)t.rows, (cuuint64_t)H, (cuuint64_t)B};
  // a size-1 axis may carry any stride: give it a legal dummy (multiple of 16 B)
  const int64_t fallback = t.ld * t.rows;
  const int64_t hs = (H > 1) ? t.hs : fallback;
  const int64_t bs = (B > 1) ? t.bs : fallback;
  if (hs <= 0 || bs <= 0) return set_error(B200AD_ERR_UNSUPPORTED, "attention(tcgen05): broadcast (stride 0) batch / head axes");
  cuuint64_t strides[3] = {(cuuint64_t)t.ld * 4, (cuuint64_t)hs * 4, (cuuint64_t)bs * 4};
  cuuint32_t box[4] = {32u, box_rows, 1u, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(t.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(B200AD_ERR_CUDA, "attention(tcgen05): cuTensorMapEncodeTiled failed (%d)", (int)r);
  return B200AD_OK;
}

static bool tensor_ok(const float* ptr, const b200ad_attn_strides& s) {
  return ptr && aligned16(ptr) && s.row > 0 && (s.row & 3) == 0 && (s.head & 3) == 0 && (s.batch & 3) == 0 &&
         s.head >= 0 && s.batch >= 0;
}

template <int MODE, bool CP>
static int launch_pass_impl(const AttnTensor& a1, const AttnTensor& a2, const AttnTensor& b1, const AttnTensor& b2,
                            const AttnTensor& b3, AttnParams p, int64_t B, int64_t H, cudaStream_t s) {
  constexpr bool HAS2 = (MODE == AT_DQ || MODE == AT_DKM);
  constexpr uint32_t BROWS = CP ? 32u : (uint32_t)AT_BC;      // rows of a K-major B block staged per CTA
  CUtensorMap m_a1, m_a2, m_b1, m_b2, m_b3, m_out;
  int rc;
  AttnTensor to{p.out, p.out_bs, p.out_hs, p.out_ld, p.R, a1.dk};
  if ((rc = encode_attn(&m_out, to, B, H, AT_BM, false))) return rc;
  if ((rc = encode_attn(&m_a1, a1, B, H, AT_BM, false))) return rc;
  if ((rc = encode_attn(&m_b1, b1, B, H, BROWS, false))) return rc;
  if ((rc = encode_attn(&m_b3, b3, B, H, 32, true))) return rc;
  if (HAS2) {
    if ((rc = encode_attn(&m_a2, a2, B, H, AT_BM, false))) return rc;
    if ((rc = encode_attn(&m_b2, b2, B, H, BROWS, false))) return rc;
  } else {
    m_a2 = m_a1;
    m_b2 = m_b1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel<MODE, CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e == cudaSuccess && CP) e = cudaFuncSetAttribute(attn_tc_kernel<MODE, CP>, cudaFuncAttributeNonPortableClusterSizeAllowed, 0);
    if (e != cudaSuccess) return set_error(B200AD_ERR_CUDA, "attention(tcgen05): smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  { const char* pb = getenv("B200AD_ATTN_PROBE"); p.probe = pb ? atoi(pb) : 0; }
  p.row_tiles = (int32_t)ceil_div(ceil_div(p.R, AT_BM), CP ? 2 : 1);      // work units per (batch, head): tiles or tile pairs
  p.nb = (int32_t)ceil_div(p.C, AT_BC);
  const int64_t items = B * H * p.row_tiles;
  p.total_items = (int32_t)items;
  const int max_units = CP ? num_sms() / 2 : num_sms();
  const int grid = (CP ? 2 : 1) * (int)(items < max_units ? items : max_units);
#ifdef AT_PROFILE
  static long long* prof_dev = nullptr;
  if (!prof_dev) cudaMalloc(&prof_dev, 16 * sizeof(long long));
  cudaMemsetAsync(prof_dev, 0, 16 * sizeof(long long), s);
  p.prof = prof_dev;
#endif
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(AT_THREADS);
    cfg.dynamicSmemBytes = AT_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CP ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, attn_tc_kernel<MODE, CP>, m_a1, m_a2, m_b1, m_b2, m_b3, m_out, p);
    if (e != cudaSuccess) return set_error(B200AD_ERR_CUDA, "attention(tcgen05): launch: %s", cudaGetErrorString(e));
  }
  count_launch();
#ifdef AT_PROFILE
  {
    long long h[16];
    cudaMemcpyAsync(h, prof_dev, sizeof(h), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    const double blocks = (double)ceil_div(p.total_items - 0, grid) * p.nb;
    fprintf(stderr, "[at_profile] mode %d items/CTA~%.1f nb=%d | per block: MMA warp total %.0f wait_conv %.0f wait_Z %.0f wait_out_empty %.0f | "
            "owner total %.0f wait_X %.0f wait_OUT %.0f fold %.0f ldX %.0f math %.0f st+arrive %.0f epilogue %.0f\n", MODE,
            (double)p.total_items / grid, p.nb, h[0] / blocks, h[1] / blocks, h[2] / blocks, h[3] / blocks, h[8] / blocks, h[9] / blocks,
            h[10] / blocks, h[11] / blocks, h[12] / blocks, h[13] / blocks, h[14] / blocks, h[15] / blocks);
  }
#endif
  if (getenv("B200AD_ATTN_SYNC")) {     // debugging aid: attribute an asynchronous failure to its pass
    cudaError_t e = cudaStreamSynchronize(s);
    fprintf(stderr, "[b200ad] attention pass mode %d (R=%d C=%d nb=%d items=%d): %s\n", MODE, p.R, p.C, p.nb,
            p.total_items, cudaGetErrorString(e));
  }
  return check_launch("attention(tcgen05)");
}

// CTA pairs (cta_group::2: each CTA stages half of every B block, the pair shares one MMA stream) whenever the row tiles
// of a (batch, head) come in pairs; otherwise the 1-CTA passes.  Measured on B200 at the config-5 micro-batch: pair
// forward 0.586 ms vs 0.664 ms, backward 2.44 vs 2.59 ms.  (With `mbarrier.arrive.release.cluster` hand-offs the pair
// variant had measured 1.01 / 3.36 ms: every such arrive carries a MEMBAR.ALL.GPU.)  B200AD_ATTN_PAIR=0 / 1 forces one.
template <int MODE>
static int launch_pass(const AttnTensor& a1, const AttnTensor& a2, const AttnTensor& b1, const AttnTensor& b2,
                       const AttnTensor& b3, AttnParams p, int64_t B, int64_t H, cudaStream_t s) {
  const char* e = getenv("B200AD_ATTN_PAIR");
  const bool pair = e ? atoi(e) == 1 : (ceil_div(p.R, AT_BM) % 2 == 0);
  if (pair) return launch_pass_impl<MODE, true>(a1, a2, b1, b2, b3, p, B, H, s);
  return launch_pass_impl<MODE, false>(a1, a2, b1, b2, b3, p, B, H, s);
}

static int attn_check(const b200ad_attn_desc* d, const char** why) {
  auto no = [&](const char* w) { if (why) *why = w; return 0; };
  if (!d) return no("null descriptor");
  if (d->head_dim < 4 || d->head_dim > AT_DK || (d->head_dim & 3)) return no("head_dim must be a multiple of 4 in 4..64");
  if (d->batch < 1 || d->heads < 1 || d->seq_q < 1 || d->seq_kv < 1) return no("empty problem");
  if (d->batch * d->heads * ceil_div(d->seq_q > d->seq_kv ? d->seq_q : d->seq_kv, AT_BM) >= (1ll << 31)) return no("too many tiles");
  if (d->seq_q >= (1ll << 31) || d->seq_kv >= (1ll << 31)) return no("sequence too long");
  if (get_encode_fn() == nullptr) return no("cuTensorMapEncodeTiled unavailable (no driver)");
  return 1;
}

// generation 2 (attn_h16.cu): the same passes carried by fp16 hi / lo pairs at the kind::f16 rate
int attn_h16_fwd(const b200ad_attn_desc* d, const float* q, const float* k, const float* v, float* o, float* lse, float* ws,
                 bool single, cudaStream_t s);
int attn_h16_bwd(const b200ad_attn_desc* d, const float* q, const float* k, const float* v, const float* o,
                 const float* dout, const float* lse, const float* fwd_maxima, float* dq, float* dk, float* dv, float* ws,
                 bool split_dkv, bool single, cudaStream_t s);

// 0 = fp16 hi/lo split (generation 2), 1 = tf32 hi/lo split (generation 1).  Default: generation 1 for the forward pass
// and generation 2 for the backward passes — measured at the config-5 micro-batch (256 sequences): forward 0.59 ms (tf32)
// vs 0.77 ms (fp16, slab-maximum pre-pass included), backward 2.45 ms (tf32) vs 1.80 ms (fp16).  -1 = that mix.
// 2 = generation 2 with dV and dK as two separate passes (the default fuses them into one).
// 3 = generation 2 issuing only the hi . hi product of every contraction: the looser-tolerance variant (11-bit operands,
//     relative error ~1e-3 like a bf16 GEMM; north_star's "bf16 tensor-core variants to a stated looser tolerance").
static std::atomic<int> g_attn_variant{getenv("B200AD_ATTN_TF32") ? 1 : (getenv("B200AD_ATTN_H16") ? 0 : -1)};

}  // namespace b200ad

using namespace b200ad;

extern "C" {

int b200ad_attn_set_variant(int variant) {
  if (variant < -1 || variant > 3) return set_error(B200AD_ERR_INVALID, "attn_set_variant: %d not in -1..3", variant);
  return g_attn_variant.exchange(variant);
}

int b200ad_attn_supported(const b200ad_attn_desc* d) {
  const char* why = nullptr;
  if (!attn_check(d, &why)) return 0;
  const b200ad_attn_strides* ss[4] = {&d->q, &d->k, &d->v, &d->o};
  for (int i = 0; i < 4; ++i)
    if (ss[i]->row <= 0 || (ss[i]->row & 3) || (ss[i]->head & 3) || (ss[i]->batch & 3) || ss[i]->head < 0 || ss[i]->batch < 0) return 0;
  return 1;
}

int64_t b200ad_attn_fwd_ws_floats(const b200ad_attn_desc* d) {
  if (!d) return 0;
  return 4 * d->batch * d->heads;          // slab maxima of Q, K, V (fp16-split variant)
}

int b200ad_attn_fwd(const b200ad_attn_desc* d, const float* q, const float* k, const float* v, float* o, float* lse,
                    float* ws, void* stream) {
  const char* why = "unsupported";
  if (!attn_check(d, &why)) return set_error(B200AD_ERR_UNSUPPORTED, "attn_fwd: %s", why);
  B200AD_REQUIRE(tensor_ok(q, d->q) && tensor_ok(k, d->k) && tensor_ok(v, d->v) && tensor_ok(o, d->o) && lse,
                 "attn_fwd: null / misaligned tensor or stride not a multiple of 4 elements");
  if (const int var = g_attn_variant.load(std::memory_order_relaxed); var == 0 || var == 2 || var == 3) {      // (the default mix runs the forward on generation 1)
    B200AD_REQUIRE(ws, "attn_fwd: null workspace (b200ad_attn_fwd_ws_floats)");
    return attn_h16_fwd(d, q, k, v, o, lse, ws, var == 3, (cudaStream_t)stream);
  }
  const int64_t hd = d->head_dim;
  AttnTensor tq{q, d->q.batch, d->q.head, d->q.row, d->seq_q, hd}, tk{k, d->k.batch, d->k.head, d->k.row, d->seq_kv, hd},
      tv{v, d->v.batch, d->v.head, d->v.row, d->seq_kv, hd};
  AttnParams p{};
  p.out = o; p.out_bs = d->o.batch; p.out_hs = d->o.head; p.out_ld = d->o.row;
  p.lse_w = lse; p.lse = nullptr; p.dvec = nullptr;
  p.H = (int32_t)d->heads; p.R = (int32_t)d->seq_q; p.C = (int32_t)d->seq_kv;
  p.c = d->scale; p.c2 = d->scale * 1.4426950408889634f;
  p.amax_w = ws;               // optional: slab maxima of Q / K / V for the backward (b200ad_attn_bwd_maxima)
  if (ws) {
    cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(float) * 4 * d->batch * d->heads, (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error(B200AD_ERR_CUDA, "attn_fwd: memset: %s", cudaGetErrorString(e));
  }
  return launch_pass<AT_FWD>(tq, tq, tk, tk, tv, p, d->batch, d->heads, (cudaStream_t)stream);
}

int64_t b200ad_attn_bwd_ws_floats(const b200ad_attn_desc* d) {
  if (!d) return 0;
  return d->batch * d->heads * (d->seq_q + 4);      // D = rowsum(dO o O), + slab maxima of Q, K, V, dO
}

int b200ad_attn_bwd(const b200ad_attn_desc* d, const float* q, const float* k, const float* v, const float* o,
                    const float* dout, const float* lse, float* dq, float* dk, float* dv, float* ws, void* stream) {
  return b200ad_attn_bwd_maxima(d, q, k, v, o, dout, lse, nullptr, dq, dk, dv, ws, stream);
}

int b200ad_attn_bwd_maxima(const b200ad_attn_desc* d, const float* q, const float* k, const float* v, const float* o,
                           const float* dout, const float* lse, const float* fwd_ws, float* dq, float* dk, float* dv,
                           float* ws, void* stream) {
  const char* why = "unsupported";
  if (!attn_check(d, &why)) return set_error(B200AD_ERR_UNSUPPORTED, "attn_bwd: %s", why);
  B200AD_REQUIRE(tensor_ok(q, d->q) && tensor_ok(k, d->k) && tensor_ok(v, d->v) && tensor_ok(o, d->o) &&
                     tensor_ok(dout, d->o) && tensor_ok(dq, d->q) && tensor_ok(dk, d->k) && tensor_ok(dv, d->v) && lse && ws,
                 "attn_bwd: null / misaligned tensor or stride not a multiple of 4 elements");
  cudaStream_t s = (cudaStream_t)stream;
  if (const int var = g_attn_variant.load(std::memory_order_relaxed); var != 1)
    return attn_h16_bwd(d, q, k, v, o, dout, lse, fwd_ws, dq, dk, dv, ws, var == 2, var == 3, s);
  const int64_t B = d->batch, H = d->heads;
  const int64_t hd = d->head_dim;
  AttnTensor tq{q, d->q.batch, d->q.head, d->q.row, d->seq_q, hd}, tk{k, d->k.batch, d->k.head, d->k.row, d->seq_kv, hd},
      tv{v, d->v.batch, d->v.head, d->v.row, d->seq_kv, hd}, tdo{dout, d->o.batch, d->o.head, d->o.row, d->seq_q, hd};
  // D = rowsum(dO o O)
  const int64_t rows = B * H * d->seq_q;
  attn_dvec_kernel<<<stream_grid(rows, 16), 256, 0, s>>>(o, dout, ws, rows, (int)H, (int)d->seq_q, d->o.batch, d->o.head,
                                                         d->o.row, (int)d->head_dim);
  count_launch();
  int rc = check_launch("attention dvec");
  if (rc) return rc;
  AttnParams p{};
  p.lse_w = nullptr; p.lse = lse; p.dvec = ws;
  p.H = (int32_t)H;
  p.c = d->scale; p.c2 = d->scale * 1.4426950408889634f;
  // dV: lanes = kv rows, columns = q rows
  p.out = dv; p.out_bs = d->v.batch; p.out_hs = d->v.head; p.out_ld = d->v.row;
  p.R = (int32_t)d->seq_kv; p.C = (int32_t)d->seq_q;
  if ((rc = launch_pass<AT_DV>(tk, tk, tq, tq, tdo, p, B, H, s))) return rc;
  // dK
  p.out = dk; p.out_bs = d->k.batch; p.out_hs = d->k.head; p.out_ld = d->k.row;
  if ((rc = launch_pass<AT_DKM>(tk, tv, tq, tdo, tq, p, B, H, s))) return rc;
  // dQ: lanes = q rows, columns = kv rows
  p.out = dq; p.out_bs = d->q.batch; p.out_hs = d->q.head; p.out_ld = d->q.row;
  p.R = (int32_t)d->seq_q; p.C = (int32_t)d->seq_kv;
  return launch_pass<AT_DQ>(tq, tdo, tk, tv, tk, p, B, H, s);
}
}
