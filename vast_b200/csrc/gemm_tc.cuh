// Warp-specialised, persistent sm_100a "NT" GEMM mainloop with pluggable epilogues.
//
//   S[prob][m][n] = sum_k A[prob][m][k] * B[prob][n][k]        (both operands K-major, 16-bit)
//
// This is the one tensor-core mainloop behind every dense contraction of the hot path
// (reference call sites: model/vast.py:405-408 `torch.matmul(feat, feat_all.permute(1,0))`,
// its autograd backward, and evaluation/evaluation_mm.py:223).  The accumulator tile lives in
// TMEM and is consumed in place by the epilogue functor (online log-sum-exp, softmax
// probabilities + hard-negative race, running top-k, or a plain store), so the [M, N] logit
// matrix never goes to HBM.
//
// Roles (one CTA per SM, persistent over work items):
//   warp 0 / lane 0 : TMA producer    -- cp.async.bulk.tensor into a STAGES-deep smem ring
//   warp 1 / lane 0 : MMA issuer      -- tcgen05.mma (128 x BN x 16 per instruction), fp32 in TMEM
//   warps 2..2+NE   : epilogue        -- tcgen05.ld TMEM -> registers -> Epi functor
// TMEM holds two accumulator stages (2 x BN columns) so the epilogue of tile i overlaps the
// mainloop of tile i+1.
//
// Work item = (problem, 128-row block, contiguous range of N-tiles, contiguous range of K-blocks).
// Epilogue state (row statistics, top-k lists, ...) persists across the N-tiles of one item and is
// flushed as a partial ("slot") at the end of the item; a small finalize kernel merges slots.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace vast {
namespace tc {

// -DVAST_EPI_TRACE (developer builds only, scripts/epi_trace.py): per-CTA %globaltimer stamps of the epilogue's phases --
// [0] roles start, [1 + 2 t] accumulator of the CTA's t-th tile ready, [2 + 2 t] its epilogue done, [15] after finish().
#ifdef VAST_EPI_TRACE
__device__ unsigned long long g_epi_trace[2 * 1024 * 16];  // [epilogues with a tile_end hook (symmetric S GEMM) ? 1 : 0][CTA][slot]
__device__ __forceinline__ void epi_trace(int slot, int tag) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  if (blockIdx.x < 1024 && slot < 16) g_epi_trace[(tag * 1024 + blockIdx.x) * 16 + slot] = t;
}
#define VAST_TRACE(cond, slot) do { if (cond) ::vast::tc::epi_trace(slot, HasTileEnd<Epi>::value ? 1 : 0); } while (0)
#else
#define VAST_TRACE(cond, slot) do { } while (0)
#endif

constexpr int BM = 128;
constexpr int BK = 64;  // 64 x 16-bit = one 128-byte swizzle span
constexpr int MAX_PROBLEMS = 2;
constexpr int MAX_SCHED = 192;  // entries of an explicit work-item table (two rounds of <= 96 clusters)

struct GemmShape {
  int num_problems;
  int M, N, K;
  int cl;        // CTAs per cluster (stacked along M, sharing every B tile through TMA multicast)
  int m_blocks;  // 128-row blocks
  int m_groups;  // groups of `cl` row blocks = what one cluster works on
  int n_tiles, n_splits, tiles_per_split;
  int k_blocks, k_splits, kb_per_split;
  int num_items;
  // Optional two-class schedule (streaming top-k): the first m_groups - tail_groups row-block groups are one
  // work item each (all N-tiles: a list that is never restarted), the last tail_groups are cut into
  // tail_splits column ranges of tail_tps tiles so that the final wave still fills the machine.
  int tail_groups, tail_splits, tail_tps;
  uint32_t idesc;
  // Optional explicit schedule (epilogues with kCustomTiles, MN-major B): work item idx is the single tile
  // sched[idx] = width64 | col64 << 3 | group << 11  -- `width64` 64-column units wide (0 = empty slot, 1..BN/64),
  // starting at column 64 * col64 of row-block group `group` (= prob * m_groups + group inside the problem).
  // Tiles narrower than BN let a static schedule balance a grid the regular tiling leaves 1.73 waves deep.
  int n_sched;
  uint32_t sched[MAX_SCHED];
};

template <class EpiParams>
struct alignas(64) KernelParams {
  CUtensorMap tmA[MAX_PROBLEMS + 1];  // [MAX_PROBLEMS]: the MN-major ("transposed view") A operand of problem a_mn_prob
  CUtensorMap tmB[MAX_PROBLEMS];
  GemmShape g;
  const int* gate;  // optional device flag: the whole launch is a no-op while *gate == 0
  // Optional MN-major A (epilogues with kAmnCapable): problem `a_mn_prob` reads its A operand as the TRANSPOSE of a
  // row-major matrix -- A[m][k] = X[k][m], X row-major [K, M] -- through tmA[MAX_PROBLEMS] (boxes of 64 k-rows x 64
  // m-columns, the same smem layout as the MN-major B operand) unless *a_mn_off != 0 (device flag read once per
  // role).  a_mn_prob1 = 1 + that problem's index; 0 (what memset leaves): every problem K-major.
  int a_mn_prob1;
  const int* a_mn_off;
  EpiParams epi;
};

// What an epilogue thread knows about the work item it is processing.
struct ItemCtx {
  int prob, m_blk, n_split, k_split;
  int row;         // row inside the problem (m_blk * 128 + TMEM lane)
  bool row_valid;  // row < M
  int slot;        // partial-result slot = n_split * HALVES + half
  int M, N;
  int lane, half;
};

struct WorkItem {
  int prob, m_blk, n_split, k_split;
  int tile_begin, tile_end, kb_begin, kb_end;
  int col_base, width;  // tile t covers columns [col_base + t * BN, +width)   (regular items: 0, BN)
};

// idx enumerates (problem, row-block group, N-split, K-split); the CTA's own row block inside the group
// is its rank in the cluster.
template <int BN, bool CUSTOM = false>
__device__ __forceinline__ WorkItem decode_item(const GemmShape& g, int idx, int cta_rank) {
  WorkItem w;
  w.col_base = 0;
  w.width = BN;
  if constexpr (CUSTOM) {
    if (g.n_sched > 0) {
      const uint32_t e = g.sched[idx];
      const int wu = static_cast<int>(e & 7u), cu = static_cast<int>((e >> 3) & 0xFFu), gg = static_cast<int>(e >> 11);
      w.prob = gg / g.m_groups;
      w.m_blk = (gg - w.prob * g.m_groups) * g.cl + cta_rank;
      w.n_split = cu;  // 0 <=> the tile that starts at column 0 (the row's publisher in EpiGrad)
      w.k_split = 0;
      w.tile_begin = 0;
      w.tile_end = wu ? 1 : 0;
      w.kb_begin = 0;
      w.kb_end = g.k_blocks;
      w.col_base = cu * 64;
      w.width = wu * 64;
      return w;
    }
  }
  if (g.tail_groups > 0) {
    const int full = g.m_groups - g.tail_groups;
    int grp = idx, split = 0;
    w.tile_begin = 0;
    w.tile_end = g.n_tiles;
    if (idx >= full) {
      // range-major: every tail group's first column range runs before any group's second one, so later ranges start
      // from the bound their finished siblings proved (a list that starts warm admits far fewer candidates)
      const int j = idx - full;
      split = j / g.tail_groups;
      grp = full + j - split * g.tail_groups;
      w.tile_begin = split * g.tail_tps;
      w.tile_end = min(w.tile_begin + g.tail_tps, g.n_tiles);
    }
    w.prob = 0;
    w.m_blk = grp * g.cl + cta_rank;
    w.n_split = split;
    w.k_split = 0;
    w.kb_begin = 0;
    w.kb_end = g.k_blocks;
    return w;
  }
  w.k_split = idx % g.k_splits;
  idx /= g.k_splits;
  w.n_split = idx % g.n_splits;
  idx /= g.n_splits;
  w.m_blk = (idx % g.m_groups) * g.cl + cta_rank;
  w.prob = idx / g.m_groups;
  w.tile_begin = w.n_split * g.tiles_per_split;
  w.tile_end = min(w.tile_begin + g.tiles_per_split, g.n_tiles);
  w.kb_begin = w.k_split * g.kb_per_split;
  w.kb_end = min(w.kb_begin + g.kb_per_split, g.k_blocks);
  return w;
}

// cluster split-K: pair `ks_rank` of the cluster takes its half of the item's k-blocks
__device__ __forceinline__ void slice_k(WorkItem& w, int ks_rank) {
  const int half = (w.kb_end - w.kb_begin + 1) >> 1;
  if (ks_rank == 0)
    w.kb_end = min(w.kb_begin + half, w.kb_end);
  else
    w.kb_begin = min(w.kb_begin + half, w.kb_end);
}

// A_RES > 0 ("A-stationary"): the CTA's whole A row block (up to A_RES k-blocks) stays resident in shared memory for
// all N-tiles of a work item and the ring stages carry B only.  For short K (retrieval at D = 512: 8 k-blocks) the
// operand traffic L2 -> SM is what binds the mainloop -- every 256 x 256 x 512 tile re-fetches 256 KB of A next to
// its 256 KB of B -- and residency halves it.
template <int BN, int STAGES, int CL, int A_RES = 0>
struct SmemLayout {
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = BN * BK * 2 / CL;  // a CTA pair keeps half of every B tile per CTA
  static constexpr uint32_t A_RES_BYTES = A_RES * A_BYTES;
  static constexpr uint32_t STAGE_BYTES = A_RES > 0 ? B_BYTES : A_BYTES + B_BYTES;
  static constexpr uint32_t RING_OFFSET = A_RES_BYTES;
  static constexpr uint32_t BAR_OFFSET = RING_OFFSET + STAGES * STAGE_BYTES;
  static constexpr uint32_t BAR_BYTES = 256;  // 2*STAGES + 6 barriers + tmem slot
  static constexpr uint32_t EPI_OFFSET = BAR_OFFSET + BAR_BYTES;
  static constexpr uint32_t ALIGN_SLACK = 1024;
  static_assert(2 * STAGES * 8 + 8 * 8 + 8 <= BAR_BYTES, "barrier block too small");
};

// Epilogues that declare `static constexpr bool kHasFinish` get finish(epilogue warp, lane, NE) called by every epilogue
// thread after the CTA's last work item (EpiGrad: the step's final loss / d tau reduction rides on the GEMM's tail).
template <class E, class = void>
struct HasCustomTiles : std::false_type {};
template <class E>
struct HasCustomTiles<E, std::void_t<decltype(E::kCustomTiles)>> : std::true_type {};
template <class E, class = void>
struct HasTileEnd : std::false_type {};
template <class E>
struct HasTileEnd<E, std::void_t<decltype(E::kHasTileEnd)>> : std::integral_constant<bool, E::kHasTileEnd> {};
// Epilogues with kSecondPass read every accumulator tile TWICE from TMEM: chunk() over the tile, then between()
// (e.g. a cross-CTA exchange of row statistics), then chunk2() over the same tile -- the accumulator stage is
// handed back to the MMA issuer only after the second pass.  TMEM reads are cheap; a global re-read is not.
template <class E, class = void>
struct HasSecondPass : std::false_type {};
template <class E>
struct HasSecondPass<E, std::void_t<decltype(E::kSecondPass)>> : std::integral_constant<bool, E::kSecondPass> {};
template <class E, class = void>
struct HasAmn : std::false_type {};
template <class E>
struct HasAmn<E, std::void_t<decltype(E::kAmnCapable)>> : std::true_type {};
template <class E, class = void>
struct HasFinish : std::false_type {};
template <class E>
struct HasFinish<E, std::void_t<decltype(E::kHasFinish)>> : std::true_type {};

// B_MN = false: B is [N rows, K cols] K-major (an "NT" GEMM, S = A . B^T with B given row-wise).
// B_MN = true : B is [K rows, N cols] row-major, i.e. the MN-major UMMA operand: each pipeline stage holds
//               BN/64 TMA boxes of 64 K-rows x 64 N-columns (128-byte swizzle atoms stacked along K),
//               so a row-major matrix is consumed as the right-hand side WITHOUT a transposed copy.
// CL = 1: one CTA per tile, tcgen05.mma.cta_group::1 (M = 128).
// CL = 2: CTA PAIRS (clusters of two SMs of one TPC) with tcgen05.mma.cta_group::2: one instruction of the
//         leader CTA multiplies a 256 x BN tile -- each CTA contributes its own 128 A rows and HALF of the B
//         tile from its shared memory and receives its 128 accumulator rows in its own TMEM.  Per CTA and
//         k-block that is 16 + 16 KB of TMA writes and 4 + 4 KB of operand reads per MMA instead of
//         16 + 32 and 4 + 8: a lone CTA at 128 x 256 needs ~190 B/clk of shared-memory traffic (TMA fill +
//         UMMA operand reads) against the 128 B/clk an SM has, which is what caps cta_group::1 near 60 % of
//         the tensor peak; the pair needs exactly 128 B/clk.
//         Protocol: both CTAs run their own TMA producer (transaction bytes of both land on the LEADER's
//         `full` barrier); only the leader issues MMAs; its tcgen05.commit is multicast to both CTAs' `empty`
//         (stage free) and `tfull` (accumulator ready) barriers; both CTAs' epilogue warps arrive on the
//         leader's `tempty` barrier (accumulator drained).
// MC = 2 (CTA pairs only): clusters of FOUR CTAs = two pairs that work on consecutive work items -- the same A rows, two
//         neighbouring column ranges -- in lockstep and SHARE every A tile: each of the four CTAs fetches a 64-row
//         slice (8 KB) of its own 128 A rows and TMA-multicasts it to the CTA of the same rank in the other pair.
//         These mainloops are bound by the L2 -> SM request rate (ncu: lts__t_sectors at 93 % of the ~6300 B/clk the
//         L2 slices deliver, tensor pipe at 61 %), not by the tensor pipe: 96 KB instead of 128 KB per k-block for
//         the two tiles.  Protocol on top of the pair's: `empty` barriers count the commits of BOTH leaders (a stage
//         is rewritten by two producers), commits to `empty` are multicast to all four CTAs; `tfull` / `tempty` stay
//         inside a pair.  The host guarantees that items 2j, 2j+1 differ only in their column range (can_share_a()).
// KS = 2 (CTA pairs only): CLUSTER SPLIT-K.  Clusters of four CTAs = two pairs that compute the SAME output tile, each over
//         half of the k-blocks; when both mainloops have drained, the second pair ships its accumulator (128 x BN fp32
//         per CTA) through distributed shared memory into the first pair's -- now idle -- operand ring, and the first
//         pair's epilogue adds it chunk by chunk before the functor sees the values: a fixed-order sum, no partials in
//         HBM, no reduce kernel.  For GEMMs with too few output tiles to fill the machine and a long K (dQ at small
//         per-rank batches).  One single-tile work item per cluster (the host sizes the grid accordingly).
// PROBE (diagnostics, scripts/gemm_probe.py): 1 = no operand loads, the MMA issuer never waits for data (the tensor
//         pipe's own rate); 2 = loads only, stages are released without MMAs (the L2 -> SM ingest rate).  Results are garbage.
template <class Epi, int BN, int STAGES, int NE, bool B_MN, int CL, int A_RES = 0, int MC = 1, int PROBE = 0, int KS = 1>
__global__ void __launch_bounds__(64 + 32 * NE + 32 * Epi::kAuxWarps, 1)
gemm_tc_kernel(const __grid_constant__ KernelParams<typename Epi::Params> P) {
  static_assert(KS == 1 || (KS == 2 && CL == 2 && MC == 1 && A_RES == 0 && BN <= 256 && Epi::kAuxWarps == 0 &&
                            !HasSecondPass<Epi>::value && !HasTileEnd<Epi>::value),
                "cluster split-K: two CTA pairs per output tile");
  static_assert(KS == 1 || STAGES * (BM * BK * 2 + BN * BK * 2 / 2) >= BM * BN * 4, "the ring must hold one accumulator tile");
  static_assert(MC == 1 || (MC == 2 && CL == 2 && A_RES == 0 && BN != 512), "A-sharing clusters of two CTA pairs");
  static_assert(NE == 4 || NE == 8, "4 or 8 epilogue warps");
  static_assert(BN == 64 || BN == 128 || BN == 256 || BN == 512, "BN");
  static_assert(BN != 512 || (B_MN && CL == 2 && A_RES == 0), "the 512-wide tile is implemented for CTA pairs with the MN-major B operand");
  static_assert(CL == 1 || CL == 2, "single CTA or CTA pair");
  static_assert(A_RES == 0 || !B_MN, "resident A is implemented for the K-major B operand");
  pdl_trigger();
  if (P.gate != nullptr) {  // the flag is written by a predecessor
    pdl_wait();
    if (*reinterpret_cast<const volatile int*>(P.gate) == 0) return;
  }
  using L = SmemLayout<BN, STAGES, CL, A_RES>;
  constexpr int HALVES = NE / 4;
  constexpr int COLS_PER_WARP = BN / HALVES;
  // BN = 512: ONE accumulator stage of all 512 TMEM columns, filled by two N = 256 instructions per k-step that share
  // the A tile -- a quarter less L2 -> SM operand traffic per FLOP than two 256-wide tiles (the mainloops here run into
  // the L2 request rate, ~12 TB/s, before they run into the tensor pipe) at the price of an epilogue that no longer
  // overlaps the next tile's MMAs: for GEMMs with one or two tiles per CTA pair and a long K loop (dQ = Pt . K).
  constexpr int ACC = BN == 512 ? 1 : 2;
  constexpr int NSUB = BN == 512 ? 2 : 1;
  constexpr int BNI = BN / NSUB;  // columns of one tcgen05.mma
  constexpr uint32_t TMEM_COLS = ACC * BN;
  constexpr bool PAIR = CL == 2;
  constexpr bool CUSTOM = HasCustomTiles<Epi>::value && B_MN;  // explicit single-tile work items of any width <= BN
  constexpr bool AMN = HasAmn<Epi>::value && A_RES == 0;      // one problem may read its A operand MN-major

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte aligned bases (in the shared address space).
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* afull = tempty + 2;   // resident A loaded (A_RES)
  uint64_t* aempty = afull + 1;   // every MMA of the item that read the resident A has retired
  uint64_t* sfull = aempty + 1;   // KS: the other pair's accumulator has landed in this CTA's ring
  uint64_t* sfree = sfull + 1;    // KS: the first pair's ring may be overwritten (its MMAs have retired)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sfree + 1);
  uint8_t* epi_smem = smem + L::EPI_OFFSET;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const GemmShape& g = P.g;
  const int cluster_rank = PAIR ? static_cast<int>(ptx::cluster_ctarank()) : 0;
  const int cta_rank = cluster_rank & 1;   // rank inside the CTA pair
  const int pair_rank = cluster_rank >> 1;  // MC = 2: which of the cluster's two pairs (pair p of cluster c = pair 2c + p of the grid)
  const bool leader = cta_rank == 0;
  const uint16_t pair_mask = static_cast<uint16_t>(3u << (2 * pair_rank));
  const uint16_t all_mask = MC == 2 ? static_cast<uint16_t>(0xF) : pair_mask;
  (void)pair_rank;
  const int cluster_id = static_cast<int>(blockIdx.x) / (CL * KS);  // KS = 2: both pairs of a cluster work on the same items
  const int num_clusters = static_cast<int>(gridDim.x) / (CL * KS);
  const int ks_rank = KS == 2 ? pair_rank : 0;

  if (warp == 0 && lane == 0) {
    for (int p = 0; p < g.num_problems; ++p) {
      ptx::tma_prefetch_desc(&P.tmA[p]);
      ptx::tma_prefetch_desc(&P.tmB[p]);
    }
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], MC);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull[a], 1);
      ptx::mbar_init(&tempty[a], NE * CL);
    }
    ptx::mbar_init(afull, 1);
    ptx::mbar_init(aempty, 1);
    ptx::mbar_init(sfull, NE);
    ptx::mbar_init(sfree, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc<CL>(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish<CL>();
  }
  ptx::tc_fence_before_sync();
  if constexpr (PAIR)
    ptx::cluster_sync_all();  // the peer's barriers and TMEM must exist before anything remote touches them
  else
    __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the predecessor's tail; operands and epilogue inputs are read below

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (every CTA)
    if (lane == 0 && PROBE != 1) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t a_phase = 0;
      int amn_prob = -1;
      if constexpr (AMN) {
        amn_prob = P.a_mn_prob1 - 1;
        if (amn_prob >= 0 && P.a_mn_off != nullptr && *reinterpret_cast<const volatile int*>(P.a_mn_off) != 0) amn_prob = -1;
      }
      (void)amn_prob;
      for (int item = cluster_id; item < g.num_items; item += num_clusters) {
        WorkItem w = decode_item<BN, CUSTOM>(g, item, cta_rank);
        if constexpr (KS == 2) slice_k(w, ks_rank);
        if constexpr (A_RES > 0) {
          // the item's A row block, once: k-block kb at A_BYTES * kb (the previous item's MMAs must have retired)
          ptx::mbar_wait(aempty, a_phase ^ 1);
          const int nkb = w.kb_end - w.kb_begin;
          if (!PAIR || leader) ptx::mbar_arrive_expect_tx(afull, static_cast<uint32_t>(CL * nkb) * L::A_BYTES);
          for (int kb = w.kb_begin; kb < w.kb_end; ++kb) {
            uint8_t* sa = smem + (kb - w.kb_begin) * L::A_BYTES;
            if constexpr (PAIR)
              ptx::tma_load_2d_pair(sa, &P.tmA[w.prob], afull, kb * BK, w.m_blk * BM);
            else
              ptx::tma_load_2d(sa, &P.tmA[w.prob], afull, kb * BK, w.m_blk * BM);
          }
          a_phase ^= 1;
        }
        for (int t = w.tile_begin; t < w.tile_end; ++t) {
          for (int kb = w.kb_begin; kb < w.kb_end; ++kb) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + L::RING_OFFSET + stage * L::STAGE_BYTES;
            uint8_t* sb = A_RES > 0 ? sa : sa + L::A_BYTES;
            if constexpr (A_RES > 0) {
              constexpr int HALF_N = BN / CL;
              if (!PAIR || leader) ptx::mbar_arrive_expect_tx(&full[stage], static_cast<uint32_t>(CL) * L::B_BYTES);
              if constexpr (PAIR)
                ptx::tma_load_2d_pair(sb, &P.tmB[w.prob], &full[stage], kb * BK, t * BN + cta_rank * HALF_N);
              else
                ptx::tma_load_2d(sb, &P.tmB[w.prob], &full[stage], kb * BK, t * BN);
            } else if constexpr (!PAIR) {
              if constexpr (B_MN) {
                const int nboxes = CUSTOM ? (w.width + 63) >> 6 : BN / 64;  // 64 K-rows x 64 columns each
                ptx::mbar_arrive_expect_tx(&full[stage], L::A_BYTES + static_cast<uint32_t>(nboxes) * (BK * 128));
                if (AMN && w.prob == amn_prob) {
                  ptx::tma_load_2d(sa, &P.tmA[MAX_PROBLEMS], &full[stage], w.m_blk * BM, kb * BK);
                  ptx::tma_load_2d(sa + BK * 128, &P.tmA[MAX_PROBLEMS], &full[stage], w.m_blk * BM + 64, kb * BK);
                } else {
                  ptx::tma_load_2d(sa, &P.tmA[w.prob], &full[stage], kb * BK, w.m_blk * BM);
                }
#pragma unroll
                for (int nb = 0; nb < BN / 64; ++nb)
                  if (nb < nboxes)
                    ptx::tma_load_2d(sb + nb * (BK * 128), &P.tmB[w.prob], &full[stage], w.col_base + t * BN + nb * 64, kb * BK);
              } else {
                ptx::mbar_arrive_expect_tx(&full[stage], L::STAGE_BYTES);
                ptx::tma_load_2d(sa, &P.tmA[w.prob], &full[stage], kb * BK, w.m_blk * BM);
                ptx::tma_load_2d(sb, &P.tmB[w.prob], &full[stage], kb * BK, t * BN);
              }
            } else {
              // the leader's barrier collects the bytes of BOTH CTAs' loads
              constexpr int HALF_N = BN / 2;  // this CTA's half of the B tile
              if constexpr (B_MN) {
                // a tile of `width` columns: this CTA holds columns [half_n * rank, +half_n) as 64-column boxes (a last,
                // partly used box is loaded whole: the tensor core reads only the first half_n columns)
                const int half_n = (CUSTOM && BN != 512) ? w.width >> 1 : HALF_N;
                const int nboxes = (CUSTOM && BN != 512) ? (half_n + 63) >> 6 : HALF_N / 64;
                if (leader) ptx::mbar_arrive_expect_tx(&full[stage], 2 * (L::A_BYTES + static_cast<uint32_t>(nboxes) * (BK * 128)));
                if constexpr (MC == 2) {
                  // this CTA's 64-row slice of the A rows it shares with CTA (rank ^ 2): lands in both
                  const uint16_t share = static_cast<uint16_t>(0x5u << cta_rank);
                  uint8_t* sl = sa + pair_rank * (BK * 128);
                  if (AMN && w.prob == amn_prob)
                    ptx::tma_load_2d_pair_mc(sl, &P.tmA[MAX_PROBLEMS], &full[stage], w.m_blk * BM + pair_rank * 64, kb * BK, share);
                  else  // (tmA: 64-row boxes in this mode)
                    ptx::tma_load_2d_pair_mc(sl, &P.tmA[w.prob], &full[stage], kb * BK, w.m_blk * BM + pair_rank * 64, share);
                } else if (AMN && w.prob == amn_prob) {  // A[m][k] = X[k][m]: two boxes of 64 k-rows x 64 m-columns
                  ptx::tma_load_2d_pair(sa, &P.tmA[MAX_PROBLEMS], &full[stage], w.m_blk * BM, kb * BK);
                  ptx::tma_load_2d_pair(sa + BK * 128, &P.tmA[MAX_PROBLEMS], &full[stage], w.m_blk * BM + 64, kb * BK);
                } else {
                  ptx::tma_load_2d_pair(sa, &P.tmA[w.prob], &full[stage], kb * BK, w.m_blk * BM);
                }
                if constexpr (BN == 512) {
                  // per N = 256 instruction this CTA supplies 128 of its 256 columns: boxes (sub, box) at (2 sub + box) * 8 KB
#pragma unroll
                  for (int sub = 0; sub < 2; ++sub)
#pragma unroll
                    for (int nb = 0; nb < 2; ++nb)
                      ptx::tma_load_2d_pair(sb + (sub * 2 + nb) * (BK * 128), &P.tmB[w.prob], &full[stage],
                                            t * BN + sub * 256 + cta_rank * 128 + nb * 64, kb * BK);
                } else {
#pragma unroll
                for (int nb = 0; nb < HALF_N / 64; ++nb)
                  if (nb < nboxes)
                    ptx::tma_load_2d_pair(sb + nb * (BK * 128), &P.tmB[w.prob], &full[stage],
                                          w.col_base + t * BN + cta_rank * half_n + nb * 64, kb * BK);
                }
              } else {
                if (leader) ptx::mbar_arrive_expect_tx(&full[stage], 2 * L::STAGE_BYTES);
                if constexpr (MC == 2)
                  ptx::tma_load_2d_pair_mc(sa + pair_rank * (BK * 128), &P.tmA[w.prob], &full[stage], kb * BK,
                                           w.m_blk * BM + pair_rank * 64, static_cast<uint16_t>(0x5u << cta_rank));
                else
                  ptx::tma_load_2d_pair(sa, &P.tmA[w.prob], &full[stage], kb * BK, w.m_blk * BM);
                ptx::tma_load_2d_pair(sb, &P.tmB[w.prob], &full[stage], kb * BK, t * BN + cta_rank * HALF_N);
              }
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA of a pair)
    if (lane == 0 && leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      uint32_t a_phase = 0;
      int amn_prob = -1;
      if constexpr (AMN) {
        amn_prob = P.a_mn_prob1 - 1;
        if (amn_prob >= 0 && P.a_mn_off != nullptr && *reinterpret_cast<const volatile int*>(P.a_mn_off) != 0) amn_prob = -1;
      }
      (void)amn_prob;
      for (int item = cluster_id; item < g.num_items; item += num_clusters) {
        WorkItem w = decode_item<BN, CUSTOM>(g, item, cta_rank);
        if constexpr (KS == 2) slice_k(w, ks_rank);
        if constexpr (A_RES > 0) {
          ptx::mbar_wait(afull, a_phase);
          ptx::tc_fence_after_sync();
          a_phase ^= 1;
        }
        for (int t = w.tile_begin; t < w.tile_end; ++t) {
          ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
          ptx::tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
          for (int kb = w.kb_begin; kb < w.kb_end; ++kb) {
            if constexpr (PROBE != 1) ptx::mbar_wait(&full[stage], phase);
            ptx::tc_fence_after_sync();
            const uint32_t sst = ptx::smem_u32(smem + L::RING_OFFSET + stage * L::STAGE_BYTES);
            const uint32_t sa = A_RES > 0 ? ptx::smem_u32(smem + (kb - w.kb_begin) * L::A_BYTES) : sst;
            const uint32_t sbb = A_RES > 0 ? sst : sst + L::A_BYTES;
            const bool a_mn = AMN && w.prob == amn_prob;
            const uint64_t da = a_mn ? ptx::umma_desc_sw128_mnmajor(sa, BK * 128) : ptx::umma_desc_sw128_kmajor(sa);
            const uint64_t db = B_MN ? ptx::umma_desc_sw128_mnmajor(sbb, BK * 128)
                                     : ptx::umma_desc_sw128_kmajor(sbb);
            // K-major: advance 16 elements (32 bytes) along K inside the swizzle span: +2 in 16-byte units.
            // MN-major: 16 K-rows of 128 bytes = 2048 bytes: +128.
            constexpr uint64_t B_KSTEP = B_MN ? 128 : 2;
            const uint64_t a_kstep = a_mn ? 128 : 2;
            // a narrower tile of an explicit schedule: the same descriptors, N taken from the item
            uint32_t idesc = (CUSTOM && BN != 512) ? ((g.idesc & ~(0x3Fu << 17)) | (static_cast<uint32_t>(w.width >> 3) << 17)) : g.idesc;
            if (a_mn) idesc |= 1u << 15;  // A operand MN-major
#pragma unroll
            for (int k = 0; k < (PROBE == 2 ? 0 : BK / 16); ++k) {
#pragma unroll
              for (int sub = 0; sub < NSUB; ++sub)  // BN = 512: two instructions share the A descriptor
                ptx::umma_f16<CL>(d_tmem + static_cast<uint32_t>(sub * BNI), da + a_kstep * k,
                                  db + static_cast<uint64_t>(sub) * ((BNI / CL) * BK * 2 >> 4) + B_KSTEP * k, idesc,
                                  (kb > w.kb_begin || k > 0) ? 1u : 0u);
            }
            ptx::umma_commit<CL>(&empty[stage], all_mask);  // frees the smem stage (in every CTA that fills it) once these MMAs retire
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          ptx::umma_commit<CL>(&tfull[acc], pair_mask);  // accumulator tile complete (in both CTAs' TMEM)
          if (++acc == ACC) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
        if constexpr (A_RES > 0) ptx::umma_commit<CL>(aempty, pair_mask);  // the resident A may be replaced (in both CTAs)
      }
    }
  } else if (warp >= 2 + NE) {
    // ------------------------------------------------------------ auxiliary warps of the epilogue (every CTA)
    // (EpiTopK: list-maintenance warps fed by the filter warps through shared-memory queues)
    if constexpr (Epi::kAuxWarps > 0) Epi::aux_main(P.epi, epi_smem, g, cluster_id, num_clusters, cta_rank, warp - 2 - NE, lane);
  } else {
    // ------------------------------------------------------------ epilogue warps (every CTA)
    const int ew = warp - 2;
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int half = ew >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    Epi epi(P.epi, epi_smem);
    int trace_tile = 0;
    (void)trace_tile;
    VAST_TRACE(ew == 0 && lane == 0, 0);
    if constexpr (KS == 2) {
      if (ks_rank == 1) {
        // ---- second pair of a split-K cluster: no functor -- the accumulator goes to the first pair's ring
        //      (layout: 16-byte vector j of chunk c of row r at ((c * 8 + j) * 128 + r) * 16: conflict-free on both sides)
        const uint32_t stage_remote = ptx::mapa_u32(smem + L::RING_OFFSET, static_cast<uint32_t>(cta_rank));
        const uint32_t sfull_remote = ptx::mapa_u32(sfull, static_cast<uint32_t>(cta_rank));
        const int row_in_cta = q * 32 + lane;
        for (int item = cluster_id; item < g.num_items; item += num_clusters) {
          ptx::mbar_wait(&tfull[acc], acc_phase);
          ptx::tc_fence_after_sync();
          ptx::mbar_wait_cluster(sfree, 0);  // (one item per cluster: phase 0)
#pragma unroll 1
          for (int c = 0; c < COLS_PER_WARP; c += 32) {
            const int col_in_tile = half * COLS_PER_WARP + c;
            uint32_t v[32];
            ptx::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN + col_in_tile), v);
            ptx::tmem_ld_wait();
            const uint32_t dst = stage_remote + static_cast<uint32_t>(((col_in_tile >> 5) * 8 * 128 + row_in_cta) * 16);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              ptx::st_cluster_v4(dst + static_cast<uint32_t>(j * 128 * 16), __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                 __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          }
          // every lane's stores are ordered before the warp's ONE remote arrival (fence, warp barrier, release-arrive)
          ptx::fence_acq_rel_cluster();
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            ptx::mbar_arrive_cluster(sfull_remote);
            ptx::mbar_arrive_leader(&tempty[acc]);
          }
          if (++acc == ACC) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      }  // (finish() below: contributes zeros, takes its ticket)
    }
    if (KS == 2 && ks_rank == 1) {
      // handled above
    } else
    for (int item = cluster_id; item < g.num_items; item += num_clusters) {
      const WorkItem w = decode_item<BN, CUSTOM>(g, item, cta_rank);
      if (CUSTOM && w.tile_end <= w.tile_begin) continue;  // empty slot of an explicit schedule
      const int cpw = CUSTOM ? w.width / HALVES : COLS_PER_WARP;  // columns of a tile handled by this warp (multiple of 32)
      ItemCtx ctx;
      ctx.prob = w.prob;
      ctx.m_blk = w.m_blk;
      ctx.n_split = w.n_split;
      ctx.k_split = w.k_split;
      ctx.row = w.m_blk * BM + q * 32 + lane;
      ctx.row_valid = ctx.row < g.M;
      ctx.slot = w.n_split * HALVES + half;
      ctx.M = g.M;
      ctx.N = g.N;
      ctx.lane = lane;
      ctx.half = half;
      epi.item_begin(ctx);
      for (int t = w.tile_begin; t < w.tile_end; ++t) {
        // Epilogues with global operands (EpiGrad) start loading the first chunk's rows while the MMAs finish.
        const int col_tile = w.col_base + t * BN;
        epi.prefetch(ctx, col_tile + half * cpw);
        ptx::mbar_wait(&tfull[acc], acc_phase);
        ptx::tc_fence_after_sync();
        VAST_TRACE(ew == 0 && lane == 0, 1 + 2 * trace_tile);
        if constexpr (KS == 2) {
          // this pair's MMAs have retired and its producer has nothing left to load (one item per cluster): the ring
          // is free for the other pair's accumulator
          if (ew == 0 && lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_u32(sfree, static_cast<uint32_t>(cta_rank + 2)));
          ptx::mbar_wait_cluster(sfull, 0);
        }
#pragma unroll(Epi::kUnrollChunks ? COLS_PER_WARP / 32 : 1)
        for (int c = 0; c < COLS_PER_WARP; c += 32) {
          if (CUSTOM && c >= cpw) break;
          const int col_in_tile = half * cpw + c;
          const uint32_t taddr =
              tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN + col_in_tile);
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(taddr, v);
          // software pipeline: the next chunk's global operands are requested before this chunk's wait
          epi.advance(ctx, col_tile + col_in_tile + 32, c + 32 < cpw);
          ptx::tmem_ld_wait();
          VAST_TRACE(!HasTileEnd<Epi>::value && ew == 0 && lane == 0 && trace_tile == 1 && c < 128, 5 + 2 * (c >> 5));
          if constexpr (KS == 2) {  // + the other half of K, in a fixed order
            const float4* part = reinterpret_cast<const float4*>(smem + L::RING_OFFSET) + ((col_in_tile >> 5) * 8 * 128 + q * 32 + lane);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 x = part[j * 128];
              v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + x.x);
              v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + x.y);
              v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + x.z);
              v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + x.w);
            }
          }
          epi.chunk(ctx, v, col_tile + col_in_tile);
          VAST_TRACE(!HasTileEnd<Epi>::value && ew == 0 && lane == 0 && trace_tile == 1 && c < 128, 6 + 2 * (c >> 5));
          __syncwarp();
        }
        if constexpr (HasSecondPass<Epi>::value) {
          epi.between(ctx);
          for (int c = 0; c < COLS_PER_WARP; c += 32) {
            const int col_in_tile = half * cpw + c;
            uint32_t v[32];
            ptx::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN + col_in_tile), v);
            ptx::tmem_ld_wait();
            epi.chunk2(ctx, v, col_tile + col_in_tile);
            __syncwarp();
          }
        }
        ptx::tc_fence_before_sync();
        if (lane == 0) {
          if constexpr (PAIR)
            ptx::mbar_arrive_leader(&tempty[acc]);  // the MMA issuer lives in the leader CTA
          else
            ptx::mbar_arrive(&tempty[acc]);
        }
        if (++acc == ACC) {
          acc = 0;
          acc_phase ^= 1;
        }
        if constexpr (HasTileEnd<Epi>::value) epi.tile_end(ctx, col_tile, ew, NE);
        VAST_TRACE(ew == 0 && lane == 0, 2 + 2 * trace_tile);
        ++trace_tile;
      }
      epi.item_end(ctx);
    }
    if constexpr (HasFinish<Epi>::value) epi.finish(ew, lane, NE);
    VAST_TRACE(ew == 0 && lane == 0, 15);
  }

  ptx::tc_fence_before_sync();
  if constexpr (PAIR)
    ptx::cluster_sync_all();  // no CTA may leave while its peer can still reach its barriers / TMEM
  else
    __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<CL>(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------ host side
int make_tmap_2d(CUtensorMap* map, const void* ptr, int dtype, int64_t rows, int64_t cols, int64_t ld_elems,
                 int box_rows);

// Pick how many N-splits (and K-splits) to cut each 128-row block into so the persistent grid
// is evenly loaded.  row_blocks = problems * m_blocks.
// item_overhead: fixed cost of one work item in units of one k-block's MMA time (pipeline fill, partial flush,
// and for the top-k epilogue the burst of list compactions after every restart).
void choose_splits(GemmShape* g, int sm_count, int max_n_splits, int max_k_splits, double item_overhead = 2.0);

// a_fmt / b_fmt: 0 = fp16, 1 = bf16 (may differ); b_mn: B operand is MN-major (see gemm_tc_kernel).
inline void fill_shape(GemmShape* g, int problems, int M, int N, int K, int BN, int a_fmt, int b_fmt = -1,
                       bool b_mn = false, int cl = 1) {
  if (b_fmt < 0) b_fmt = a_fmt;
  g->num_problems = problems;
  g->M = M;
  g->N = N;
  g->K = K;
  g->cl = cl;
  g->m_blocks = ceil_div(M, BM);
  g->m_groups = ceil_div(g->m_blocks, cl);
  g->n_tiles = ceil_div(N, BN);
  g->k_blocks = ceil_div(K, BK);
  g->n_splits = 1;
  g->tiles_per_split = g->n_tiles;
  g->k_splits = 1;
  g->kb_per_split = g->k_blocks;
  g->num_items = problems * g->m_groups;
  g->tail_groups = g->tail_splits = g->tail_tps = 0;
  g->n_sched = 0;
  g->idesc = ptx::umma_idesc_f16(static_cast<uint32_t>(a_fmt), static_cast<uint32_t>(b_fmt), b_mn ? 1u : 0u,
                                 static_cast<uint32_t>(BM * cl), static_cast<uint32_t>(BN > 256 ? 256 : BN));
}

inline int ceil_div_i(int a, int b) { return (a + b - 1) / b; }

// CTA pairs (cta_group::2) as soon as there are two 128-row blocks to pair.
inline int pick_cluster(int M) { return ceil_div(M, BM) >= 2 ? 2 : 1; }

// MC = 2 needs: whole-K items, an even number of equally long column ranges per row-block group (so that items 2j and
// 2j + 1 walk the same A tiles the same number of times), a regular schedule, A tensor maps with 64-row boxes.
inline bool can_share_a(const GemmShape& g) {
  return g.cl == 2 && g.k_splits == 1 && g.tail_groups == 0 && g.n_sched == 0 && g.n_splits % 2 == 0 &&
         g.n_tiles % g.n_splits == 0 && g.num_items % 2 == 0;
}

// How many clusters of `cluster` CTAs of `kern` the device holds at once (0 = none).
template <class K>
int resident_clusters(K kern, int cluster, int block, size_t smem) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(device_sm_count() / cluster * cluster), 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

template <class Epi, int BN, int STAGES, int NE, bool B_MN, int CL, int A_RES = 0, int MC = 1, int PROBE = 0, int KS = 1>
int launch_gemm_cl(const KernelParams<typename Epi::Params>& P, cudaStream_t stream, const char* name, size_t epi_smem_bytes) {
  using L = SmemLayout<BN, STAGES, CL, A_RES>;
  const size_t smem = L::EPI_OFFSET + L::ALIGN_SLACK + epi_smem_bytes;
  VAST_REQUIRE(smem <= 232448, VAST_ERR_UNSUPPORTED, "%s: %zu bytes of shared memory exceed the 227 KB limit", name, smem);
  VAST_REQUIRE(P.g.cl == CL, VAST_ERR_INVALID, "%s: shape planned for clusters of %d, launched with %d", name, P.g.cl, CL);
  VAST_REQUIRE(A_RES == 0 || (P.g.k_blocks <= A_RES && P.g.k_splits == 1), VAST_ERR_INVALID,
               "%s: %d k-blocks do not fit the %d resident ones", name, P.g.k_blocks, A_RES);
  VAST_REQUIRE(MC == 1 || can_share_a(P.g), VAST_ERR_INVALID, "%s: the work items do not pair up for A sharing", name);
  VAST_REQUIRE(KS == 1 || (P.g.k_splits == 1 && P.g.tiles_per_split == 1 && P.g.tail_groups == 0 && P.g.n_sched == 0),
               VAST_ERR_INVALID, "%s: cluster split-K takes single-tile whole-K work items", name);
  auto kern = gemm_tc_kernel<Epi, BN, STAGES, NE, B_MN, CL, A_RES, MC, PROBE, KS>;
  constexpr int block = 64 + 32 * NE + 32 * Epi::kAuxWarps;
  static size_t attr_smem = 0;  // per instantiation; grows monotonically
  static int resident = 0;      // MC = 2: clusters of four that fit the device at once
  if (smem > attr_smem) {
    VAST_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_smem = smem;
    resident = 0;
  }
  int clusters_max = device_sm_count() / CL;  // in CTA pairs (or lone CTAs)
  if constexpr (MC == 2 || KS == 2) {
    if (resident == 0) {
      resident = resident_clusters(kern, CL * MC * KS, block, smem);
      VAST_REQUIRE(resident > 0, VAST_ERR_UNSUPPORTED, "%s: no cluster of %d CTAs fits this device", name, CL * MC * KS);
    }
    clusters_max = resident * MC;
  }
  if constexpr (KS == 2) {
    // every work item is one cluster (two pairs), and all of them run at once
    VAST_REQUIRE(P.g.num_items <= resident, VAST_ERR_UNSUPPORTED, "%s: %d split-K items exceed the %d resident clusters", name,
                 P.g.num_items, resident);
    if (P.g.num_items <= 0) return VAST_OK;
    VAST_TIMED(stream, name, (launch_ex(kern, static_cast<unsigned>(P.g.num_items * CL * KS), block, smem, stream, CL * KS, P)));
    VAST_LAUNCH_OK(name);
    return VAST_OK;
  }
  int clusters = P.g.num_items < clusters_max ? P.g.num_items : clusters_max;
  if (MC == 2) clusters &= ~1;
  if (clusters <= 0) return VAST_OK;
  VAST_TIMED(stream, name, (launch_ex(kern, static_cast<unsigned>(clusters * CL), block, smem, stream, CL * MC, P)));
  VAST_LAUNCH_OK(name);
  return VAST_OK;
}

// Resident clusters of four CTAs of the A-sharing (MC = 2) or cluster split-K (KS = 2) form of a kernel: a decision-time
// query, so that a device that cannot hold such clusters simply keeps the lone-pair form (-1 until first asked).
template <class Epi, int BN, int STAGES, int NE, bool B_MN, int MC, int KS>
int quad_resident(size_t epi_smem_bytes) {
  static_assert(MC * KS == 2, "clusters of two CTA pairs");
  static int r = -1;
  if (r < 0) {
    using L = SmemLayout<BN, STAGES, 2, 0>;
    const size_t smem = L::EPI_OFFSET + L::ALIGN_SLACK + epi_smem_bytes;
    auto kern = gemm_tc_kernel<Epi, BN, STAGES, NE, B_MN, 2, 0, MC, 0, KS>;
    if (smem > 232448 || cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
      cudaGetLastError();
      r = 0;
    } else {
      r = resident_clusters(kern, 4, 64 + 32 * NE + 32 * Epi::kAuxWarps, smem);
    }
  }
  return r;
}
template <class Epi, int BN, int STAGES, int NE, bool B_MN>
int ks_resident(size_t epi_smem_bytes) {
  return quad_resident<Epi, BN, STAGES, NE, B_MN, 1, 2>(epi_smem_bytes);
}

// STAGES is the ring depth of a lone CTA (48 KB stages at BN = 256); a CTA pair has 32 KB stages and takes
// STAGES2 of them (default: the same shared-memory footprint -> 1.5x the depth).
template <class Epi, int BN, int STAGES, int NE, bool B_MN = false, int STAGES2 = STAGES + STAGES / 2>
int launch_gemm(const KernelParams<typename Epi::Params>& P, cudaStream_t stream, const char* name,
                size_t epi_smem_bytes = 0) {
  if (P.g.cl == 2) return launch_gemm_cl<Epi, BN, STAGES2, NE, B_MN, 2>(P, stream, name, epi_smem_bytes);
  return launch_gemm_cl<Epi, BN, STAGES, NE, B_MN, 1>(P, stream, name, epi_smem_bytes);
}

// ------------------------------------------------------------------ plain store epilogue
// C[k_split][prob][row][col] = alpha * S   (fp32).  Used by the dQ GEMM (split-K partials are
// summed in a fixed order by the gradient finalize kernel) and by the GEMM unit test.
struct EpiStore {
  struct Params {
    float* C;
    int64_t ldc;
    int64_t prob_stride;
    int64_t ksplit_stride;
    float alpha;
  };
  static constexpr bool kUnrollChunks = false;
  static constexpr int kAuxWarps = 0;
  const Params& p;
  __device__ EpiStore(const Params& p_, uint8_t*) : p(p_) {}
  __device__ __forceinline__ void item_begin(const ItemCtx&) {}
  __device__ __forceinline__ void prefetch(const ItemCtx&, int) {}
  __device__ __forceinline__ void advance(const ItemCtx&, int, bool) {}
  __device__ __forceinline__ void chunk(const ItemCtx& c, const uint32_t (&v)[32], int col0) {
    if (!c.row_valid || col0 >= c.N) return;
    float* dst = p.C + c.k_split * p.ksplit_stride + c.prob * p.prob_stride + static_cast<int64_t>(c.row) * p.ldc + col0;
    const bool vec = (col0 + 32 <= c.N) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    if (vec) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 o;
        o.x = __uint_as_float(v[i]) * p.alpha;
        o.y = __uint_as_float(v[i + 1]) * p.alpha;
        o.z = __uint_as_float(v[i + 2]) * p.alpha;
        o.w = __uint_as_float(v[i + 3]) * p.alpha;
        *reinterpret_cast<float4*>(dst + i) = o;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < c.N) dst[i] = __uint_as_float(v[i]) * p.alpha;
    }
  }
  __device__ __forceinline__ void item_end(const ItemCtx&) {}
};

}  // namespace tc
}  // namespace vast
