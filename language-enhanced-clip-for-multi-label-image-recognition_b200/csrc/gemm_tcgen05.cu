// Persistent warp-specialised bf16 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   out[M,N] = epi( A[M,K] · W[N,K]^T + bias (+ residual) )
//
// A tiles (128 x BK) and W tiles (BN x BK) are staged in shared memory by TMA in the canonical
// K-major swizzled layout (BK*2 bytes per row: 128B or 64B swizzle), multiplied by tcgen05.mma
// (one elected thread, M=128, N=BN, K=16 per instruction) into a double-buffered fp32 accumulator in
// TMEM, and drained by four epilogue warps with tcgen05.ld (thread == output row).  In CONV mode the A
// producer issues TMA *im2col* loads over the NHWC activation tensor — one (tap, channel-block) per K
// step, padding and row/image wrap handled by the TMA unit — so a 3x3 convolution is the same GEMM
// with K = 9*Cin.
//
// Epilogue (bf16 output): the tile leaves through shared memory in 64-column blocks.  A dedicated DMA
// warp TMA-loads the bf16 residual block into a swizzled staging buffer ahead of time, the epilogue
// warps add bias / residual, apply ReLU / QuickGELU, overwrite the block in place, and the DMA warp
// TMA-stores it — every HBM access of the epilogue is a full-line bulk transfer, overlapped with the
// next tile's MMAs.  fp32 outputs (small / final GEMMs) use direct vector stores instead.
//
// Roles (32 * (kEpiWarps + 3) = 352 threads): warps 0-7 epilogue in groups of four (TMEM lane quarter = warp id % 4;
// staged blocks are dealt round-robin to the groups, so two blocks are in flight per SM), warp 8 TMA producer, warp 9
// MMA issuer + TMEM owner, warp 10 epilogue DMA.  Single-thread roles run on one elected lane (elect.sync) so their
// code stays on the uniform datapath.  mbarrier rings: smem full/empty (operands), TMEM full/empty (accumulators),
// staging free/full (epilogue blocks).
#include <stdio.h>
#include <stdlib.h>

#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

// CTA-pair (cta_group::2) path for the wide layers: gemm_pair.cu
bool pair_gemm_eligible(int64_t M, int N, int K, unsigned flags, bool has_sumsq_f32_out);
bool pair_conv_eligible(int B, int H, int Wd, int Cin, int Cout, unsigned flags);
int launch_pair_gemm(const void* A, const void* Wt, const float* bias, const void* residual, void* out, float* row_sumsq,
                     int64_t M, int N, int K, unsigned flags, cudaStream_t stream, const void* A2 = nullptr, int K1 = 0);
int launch_pair_conv3x3(const void* x, const void* w, const float* bias, void* out, int B, int H, int Wd, int Cin, int Cout,
                        unsigned flags, cudaStream_t stream);

// experiment / kill switches, read from the environment ONCE at first use (round 1 called getenv on every launch)
static bool env_flag_no_resident() { static const bool v = getenv("LECB_NO_RESIDENT") != nullptr; return v; }
static bool env_flag_no_halo() { static const bool v = getenv("LECB_NO_HALO") != nullptr; return v; }
static bool env_flag_halo_single() { static const bool v = getenv("LECB_HALO_SINGLE") != nullptr; return v; }
static bool env_flag_no_mt2() { static const bool v = getenv("LECB_NO_MT2") != nullptr; return v; }

constexpr int kTileM = 128;
constexpr int kEpiWarps = 8;                       // groups of four (a warp reads TMEM lane quarter warp % 4).  Measured with
                                                   // 12: no faster on the short-K residual GEMMs and 3-20 % slower on the convs
                                                   // (the 128-register cap of 480 threads costs more than the third group hides)
constexpr int kGroups = kEpiWarps / 4;
constexpr int kTFull = (kGroups % 2 == 0) ? kGroups : 2 * kGroups;   // single-tile kernels: lcm(2 accumulators, kGroups)
constexpr int kTFullMax = 4;                       // barrier slots reserved in the layout (paired-tile kernels use four)
constexpr int kNumThreads = 32 * (kEpiWarps + 3);
constexpr int kWarpTma = kEpiWarps, kWarpMma = kEpiWarps + 1, kWarpDma = kEpiWarps + 2;
constexpr int kBiasBytes = 2048;                   // [2][256] floats: a tile's bias slice, double-buffered by tile parity
constexpr int kBarBytes = 512;
constexpr int kMaxStages = 8;

template <int BN, int BK, int NB>
struct GemmCfg {
  static constexpr int kABytes = kTileM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kCCols = BN >= 64 ? 64 : BN;               // columns per staged epilogue block
  static constexpr int kCBlocks = BN / kCCols;
  static constexpr int kCBytes = kTileM * kCCols * 2;             // 16 KB (8 KB for BN = 32)
  static constexpr int kStagesRaw = (227 * 1024 - NB * kCBytes - kBarBytes - kBiasBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = (2 * BN) < 32 ? 32 : (2 * BN);
  static constexpr int kSmemBytes = kStages * kStageBytes + NB * kCBytes + kBarBytes + kBiasBytes;
  static_assert(kStages >= 2, "need at least a double-buffered operand ring");
};

struct GemmParams {
  const float* bias;
  const void* residual;
  void* out;
  float* row_sumsq;
  int64_t M;
  int N;
  int num_kb;        // K / BK
  int num_m_tiles;
  int num_n_tiles;
  unsigned flags;
  int staged;        // output through swizzled smem blocks + TMA store (bf16: 64 columns per block, fp32: 32)
  int out_f32;       // staged fp32 output (+ optional fp32 residual)
  int cblocks;       // staged blocks per tile = BN / columns per block
  int b_resident;    // all K blocks of this CTA's W tile stay in smem for the CTA's lifetime (the CTA owns ONE n tile
                     // and strides over m tiles); the A ring gets the rest of the operand region
  int res_stages;    // A-ring depth in b_resident mode
  int operand_bytes; // size of the operand region (ring, or resident W + A ring); staging buffers follow it
  // conv mode
  int H, W, kb_per_tap;
  // halo-tile conv mode (Cin == 64, W resident): an m tile is a th x tw patch of one image; the stage holds three
  // (th+2) x tw halo copies (one per horizontal tap, so every tap's A operand is a dense, 1024-byte aligned
  // [128][64] K-major tile at copy[kx] + ky * tw rows): each input pixel crosses L2 -> SM ~3.4x instead of 9x
  int halo, tw, th, tiles_x, tiles_y, copy_bytes, batch;
  int mt;             // 1, or 2 = paired m tiles (BN <= 128, im2col convs): a stage holds TWO A tiles and one W tile, the W
                      // tile feeds two accumulators (four 128-column accumulators, still double-buffered), so W crosses
                      // L2 -> SM once per 256 pixels; the epilogue / DMA still see a sequence of 128-row tiles
  int pool;           // halo mode only: 2x2 average pool (M:147, M:27) fused into the epilogue — the four pixels of a window are
                      // lanes l, l^1, l^tw, l^tw^1 of one epilogue warp; the staged block is the (th/2 x tw/2) pooled patch
  int hilo;           // A is [M, 2K] = [hi | lo] (an exact fp16 split of an fp32 query, T:445): k block kb multiplies columns
                      // (kb & 1) * K + (kb >> 1) * BK of A with k block kb >> 1 of W, so both halves accumulate against the
                      // same W tile (its second fetch is an L2 hit) and the fp32 product needs no second pass over W
  float* topk_val;    // fused per-row top-10 epilogue (caption retrieval, T:446): instead of storing the tile, every epilogue
  int* topk_idx;      // thread keeps the 10 largest values of its row over the n tiles this CTA processes and writes them
  int topk_slots;     // to slot blockIdx.x of [rows][topk_slots][10]; a merge kernel finishes (retrieval.cu)
  int kb_split;       // > 0: TWO A operands (lecb_gemm_bf16_dual): k blocks [0, kb_split) come from A1 (tmA), the rest from A2
                      // (which travels in the residual tensor-map slot: the mode has no residual) — the K-concatenated GEMM
                      // [A1 | A2] . [W1 | W2]^T without ever materialising the concatenation
  int halo_single;    // tw == 8: ONE (th+2) x (tw+2) halo copy per stage.  The swizzle of a K-major operand is a function of
                      // the absolute shared-memory address bits (measured: a descriptor may start on any 128-byte row
                      // with base_offset 0), so tap (ky,kx) is just start = copy + (ky*(tw+2) + kx) * 128 with an
                      // 8-row-group stride of (tw+2) * 128 bytes: every input pixel crosses L2 -> SM 1.4x instead of 9x
};

// MT = m tiles per stage (1, or 2 = paired tiles: see GemmParams::mt); a template parameter so the single-tile kernels
// carry none of the paired-tile code (as a runtime switch it cost the short-tile kernels 5-30 %)
template <int BN, int BK, int NB, bool kConv, int MT>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const GemmParams p) {
  using Cfg = GemmCfg<BN, BK, NB>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kCCols = Cfg::kCCols;
  constexpr int kCBlocks = Cfg::kCBlocks;
  extern __shared__ __align__(1024) uint8_t smem[];     // swizzled tiles need 1024-byte alignment (checked below)
  const int nstages = (p.b_resident || MT == 2) ? p.res_stages : kStages;
  const int a_stride = MT * Cfg::kABytes;            // bytes of A per stage
  uint8_t* sA = smem;
  uint8_t* sB = smem + nstages * a_stride;
  uint8_t* sC = smem + p.operand_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sC + NB * Cfg::kCBytes);
  uint64_t* full = bars;                       // [kMaxStages] (the resident-W mode may run a deeper A ring)
  uint64_t* empty = bars + kMaxStages;
  // tfull: one barrier per (tile mod kTF), kTF = lcm(accumulators in flight, kGroups) — 2 for the single-tile kernels, FOUR
  // for the paired-tile kernels (four accumulators).  The barrier count must equal the accumulator count: the next completion
  // of tfull[i] then needs the accumulator it guards to be drained first, so a waiter can never be lapped.  Round 1 indexed
  // the paired-tile kernels' four accumulators with two barriers: the MMA thread, allowed to run four tiles ahead, could
  // complete tfull[1] twice more (tiles t+2, t+4... of the other accumulator pair) before a delayed epilogue warp had looked
  // at tile t — the waiter then aliased on the phase parity and the CTA deadlocked (seen once in ~10 long bench runs, caught by
  // the mbarrier time-out: MMA on tempty[3], epilogue on tfull[1], both parity 1).  Successive phases of one barrier
  // then belong to the same accumulator AND the same owner groups, so every waiter sees every phase of the barriers it
  // waits on (a group that skipped phases, or lagged two behind, would alias on the phase parity)
  uint64_t* tfull = bars + 2 * kMaxStages;
  constexpr int kTF = MT == 2 ? 4 : kTFull;
  static_assert(kTF <= kTFullMax, "tfull slots");
  uint64_t* tempty = tfull + kTFullMax;
  // staging barriers: one (free, full) pair per buffer — but never fewer than two pairs: with a single buffer the two
  // epilogue groups would share one barrier and a group could be two phases ahead of it (parity aliasing), so the
  // barrier index (block % kNBar) is decoupled from the buffer index (block % NB)
  constexpr int kNBar = NB < kGroups ? kGroups : NB;
  uint64_t* cfree = tempty + 4;                // tempty[4]: one per accumulator (two are used unless mt == 2)
  uint64_t* cfull = cfree + kNBar;
  uint64_t* bres = cfull + kNBar;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres + 1);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kBarBytes);      // [2][256]: bias slice of a tile
  // b_resident: [ W k-blocks (num_kb * kBBytes) | A ring (res_stages * kABytes) ] inside the operand region
  uint8_t* sA_ring = p.b_resident ? smem + p.num_kb * Cfg::kBBytes : sA;

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  // mt == 2: the schedule below runs over PAIRS of m tiles (num_m_tiles is even); sub-tile i belongs to pair i >> 1
  const int num_tiles = (p.num_m_tiles / MT) * p.num_n_tiles;
  // Tile schedule.  Default: tile t = blockIdx.x + i * gridDim.x, n fastest.  b_resident: the CTA keeps n tile
  // blockIdx.x % num_n_tiles and walks m tiles blockIdx.x / num_n_tiles + i * (gridDim.x / num_n_tiles).
  // Fused top-10 epilogue: the CTA owns ONE m block (blockIdx.x % num_m_tiles: its running lists are flushed once) and walks
  // n tiles blockIdx.x / num_m_tiles + i * (gridDim.x / num_m_tiles); the CTAs of different m blocks read the same bank
  // tile at about the same time, so the bank crosses HBM -> L2 once however many query blocks there are.
  const bool topk_sched = p.topk_val != nullptr;
  const int tk_stride = topk_sched ? static_cast<int>(gridDim.x) / p.num_m_tiles : 1;
  const int tk_n0 = topk_sched ? static_cast<int>(blockIdx.x) / p.num_m_tiles : 0;
  const int res_n = p.b_resident ? static_cast<int>(blockIdx.x) % p.num_n_tiles : 0;
  const int res_m0 = p.b_resident ? static_cast<int>(blockIdx.x) / p.num_n_tiles : 0;
  const int res_ms = p.b_resident ? static_cast<int>(gridDim.x) / p.num_n_tiles : 1;
  const int my_tiles = topk_sched ? (tk_n0 < p.num_n_tiles ? (p.num_n_tiles - tk_n0 + tk_stride - 1) / tk_stride : 0) :
                       MT * (p.b_resident
                           ? (res_m0 < p.num_m_tiles ? (p.num_m_tiles - res_m0 + res_ms - 1) / res_ms : 0)
                           : (static_cast<int>(blockIdx.x) < num_tiles ? (num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0));
  auto tile_coords = [&](int i, int& m_blk, int& n_blk) {
    if (topk_sched) {
      m_blk = static_cast<int>(blockIdx.x) % p.num_m_tiles;
      n_blk = tk_n0 + i * tk_stride;
    } else if (p.b_resident) {
      m_blk = res_m0 + i * res_ms;
      n_blk = res_n;
    } else if (MT == 2) {
      const int tile = static_cast<int>(blockIdx.x) + (i >> 1) * static_cast<int>(gridDim.x);
      const int mp = tile / p.num_n_tiles;
      n_blk = tile - mp * p.num_n_tiles;
      m_blk = 2 * mp + (i & 1);
    } else {
      const int tile = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
      m_blk = tile / p.num_n_tiles;
      n_blk = tile - m_blk * p.num_n_tiles;
    }
  };

  if (warp == kWarpTma && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.kb_split > 0) tma_prefetch_desc(&tmR);
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kTFullMax; ++i) mbar_init(&tfull[i], 1);
    for (int i = 0; i < 4; ++i) {
      // staged path: both epilogue groups read every accumulator (8 warps) unless a tile is a single block,
      // in which case the groups alternate tiles (4 warps); direct fp32 path: group 0 only
      mbar_init(&tempty[i], p.staged ? 4 * p.cblocks : 4);
    }
    for (int i = 0; i < kNBar; ++i) {
      mbar_init(&cfree[i], 1);
      mbar_init(&cfull[i], 4);
    }
    mbar_init(bres, 1);
    fence_barrier_init();
  }
  if (warp == kWarpMma) {
    tmem_alloc(tmem_slot, static_cast<uint32_t>(MT * Cfg::kTmemCols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();            // the prologue above overlapped the previous kernel's tail; no global access before this point
  if (my_tiles <= MT) pdl_trigger();       // at most one work item: it is the last one (else: the producer, below)

  if (warp == kWarpTma) {
    // ------------------------------- TMA producer -------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      if (p.b_resident && my_tiles > 0) {      // this CTA's whole W tile, once
        mbar_arrive_expect_tx(bres, static_cast<uint32_t>(p.num_kb) * Cfg::kBBytes);
        for (int kb = 0; kb < p.num_kb; ++kb) tma_load_2d(&tmB, bres, smem + kb * Cfg::kBBytes, kb * BK, res_n * BN);
      }
      for (int ti = 0; ti < my_tiles; ti += MT) {
        int m_blk, n_blk;
        tile_coords(ti, m_blk, n_blk);
        if (ti + MT >= my_tiles && my_tiles > MT) pdl_trigger();       // last work item of a longer run (lecb_common.cuh)
        if (kConv && p.halo) {
          const int tx = m_blk % p.tiles_x;
          const int ty = (m_blk / p.tiles_x) % p.tiles_y;
          const int img = m_blk / (p.tiles_x * p.tiles_y);
          mbar_wait(&empty[stage], phase ^ 1);
          if (p.halo_single) {
            mbar_arrive_expect_tx(&full[stage], static_cast<uint32_t>((p.th + 2) * (p.tw + 2) * 128));
            tma_load_4d(&tmA, &full[stage], sA_ring + stage * p.copy_bytes, 0, tx * p.tw - 1, ty * p.th - 1, img);
            if (++stage == nstages) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          mbar_arrive_expect_tx(&full[stage], 3u * static_cast<uint32_t>(p.copy_bytes));
          uint8_t* dst = sA_ring + stage * 3 * p.copy_bytes;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            tma_load_4d(&tmA, &full[stage], dst + kx * p.copy_bytes, 0, tx * p.tw - 1 + kx, ty * p.th - 1, img);
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
          continue;
        }
        // base pixel of the tile (and of its pair partner): scalars, not arrays — a runtime-indexed local array lives
        // in local memory, and with the whole carve-out given to shared memory every such load is an L2 round trip
        // on the producer's critical path (measured: +13 % on the 256-channel convs)
        int pw0 = 0, ph0 = 0, pn0 = 0, pw1 = 0, ph1 = 0, pn1 = 0;
        if (kConv) {
          const int hw = p.H * p.W;
          const int64_t m0 = static_cast<int64_t>(m_blk) * kTileM;
          pn0 = static_cast<int>(m0 / hw);
          const int rem0 = static_cast<int>(m0 - static_cast<int64_t>(pn0) * hw);
          ph0 = rem0 / p.W;
          pw0 = rem0 - ph0 * p.W;
          if (MT == 2) {
            const int64_t m1 = m0 + kTileM;
            pn1 = static_cast<int>(m1 / hw);
            const int rem1 = static_cast<int>(m1 - static_cast<int64_t>(pn1) * hw);
            ph1 = rem1 / p.W;
            pw1 = rem1 - ph1 * p.W;
          }
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], p.b_resident ? Cfg::kABytes : MT * Cfg::kABytes + Cfg::kBBytes);
          if (kConv) {
            const int tap = kb / p.kb_per_tap;
            const int cb = kb - tap * p.kb_per_tap;
            const int ky = tap / 3, kx = tap - ky * 3;
            tma_load_im2col_4d(&tmA, &full[stage], sA_ring + stage * a_stride, cb * BK, pw0 - 1, ph0 - 1, pn0,
                               static_cast<uint16_t>(kx), static_cast<uint16_t>(ky));
            if (MT == 2)
              tma_load_im2col_4d(&tmA, &full[stage], sA_ring + stage * a_stride + Cfg::kABytes, cb * BK, pw1 - 1, ph1 - 1,
                                 pn1, static_cast<uint16_t>(kx), static_cast<uint16_t>(ky));
          } else {
            const bool second = p.kb_split > 0 && kb >= p.kb_split;                 // dual-A mode: the tail of K is A2's
            const int a_col = p.hilo ? (kb & 1) * (p.num_kb >> 1) * BK + (kb >> 1) * BK : (second ? kb - p.kb_split : kb) * BK;
            tma_load_2d(second ? &tmR : &tmA, &full[stage], sA_ring + stage * a_stride, a_col, m_blk * kTileM);
          }
          if (!p.b_resident)
            tma_load_2d(&tmB, &full[stage], sB + stage * Cfg::kBBytes, (p.hilo ? (kb >> 1) : kb) * BK, n_blk * BN);
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ------------------------------- MMA issuer ---------------------------------
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(BN, (p.flags & LECB_GEMM_F16_OPERANDS) == 0);
      int stage = 0;
      uint32_t phase = 0;
      const int nacc = 2 * MT;               // accumulators: sub-tile t lives in accumulator t % nacc
      if (p.b_resident && my_tiles > 0) mbar_wait(bres, 0);
      // halo-tile conv: descriptor pieces that do not change from tile to tile
      uint32_t halo_off[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      uint64_t halo_a_hi = 0, halo_a0 = 0, halo_b0 = 0;
      if (kConv && p.halo) {
        const uint32_t pitch = static_cast<uint32_t>(p.tw + 2) * 128u;       // single-copy variant: halo row pitch
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t ky = tap / 3, kx = tap % 3;
          halo_off[tap] = (p.halo_single ? ky * pitch + kx * 128u
                                         : kx * static_cast<uint32_t>(p.copy_bytes) + ky * static_cast<uint32_t>(p.tw) * (BK * 2)) >> 4;
        }
        uint64_t d0 = make_kmajor_desc(0, BK * 2);
        if (p.halo_single) d0 = (d0 & ~(static_cast<uint64_t>(0x3FFF) << 32)) | (static_cast<uint64_t>(pitch >> 4) << 32);
        halo_a_hi = d0 & 0xFFFFFFFF00000000ull;
        halo_a0 = d0 & 0xFFFFFFFFull;                 // low word without the address field (LBO bits)
        halo_b0 = make_kmajor_desc(smem_u32(smem), BK * 2);
      }
      for (int ti = 0; ti < my_tiles; ti += MT) {
        // sub-tile t uses accumulator t % nacc in round t / nacc (nacc = 2 or 4)
        const uint32_t round_par = (static_cast<uint32_t>(ti) >> (MT == 2 ? 2 : 1)) & 1u;
        mbar_wait(&tempty[ti & (nacc - 1)], round_par ^ 1u);
        if (MT == 2) mbar_wait(&tempty[(ti + 1) & (nacc - 1)], round_par ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>((ti & (nacc - 1)) * BN);
        if (kConv && p.halo) {
          // The nine taps are issued from a fully unrolled sequence: every operand descriptor is the stage's base
          // descriptor plus a per-tap offset (16-byte units) computed once per CTA.  With N <= 128 an MMA lasts only
          // 45-64 cycles (smem-read bound, tools/micro/mma_chain.cu), so per-tap address arithmetic on the issuing
          // thread (ncu: ~250 cycles per tap in a rolled loop) was what bound these layers, not the tensor pipe.
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint8_t* src = sA_ring + stage * (p.halo_single ? 1 : 3) * p.copy_bytes;
          const uint32_t a_lo = static_cast<uint32_t>(halo_a0) + (smem_u32(src) >> 4);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t adesc = halo_a_hi | static_cast<uint64_t>(a_lo + halo_off[tap] + 2u * k);
              const uint64_t bdesc = halo_b0 + static_cast<uint64_t>(tap * (Cfg::kBBytes >> 4) + 2 * k);
              umma_f16(d_tmem, adesc, bdesc, idesc, (tap | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty[stage]);
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
          umma_commit(&tfull[ti % kTF]);
          continue;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = make_kmajor_desc(smem_u32(sA_ring + stage * a_stride), BK * 2);
          const uint64_t bdesc = make_kmajor_desc(
              smem_u32(p.b_resident ? smem + kb * Cfg::kBBytes : sB + stage * Cfg::kBBytes), BK * 2);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the (addr >> 4) field
            umma_f16(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                     (kb | k) != 0 ? 1u : 0u);
          }
          if (MT == 2) {                     // second m tile of the pair: next A tile of the stage, next accumulator
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_f16(d_tmem + BN, adesc + static_cast<uint64_t>((Cfg::kABytes >> 4) + 2 * k),
                       bdesc + static_cast<uint64_t>(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[ti % kTF]);
        if (MT == 2) umma_commit(&tfull[(ti + 1) % kTF]);
      }
    }
  } else if (warp == kWarpDma) {
    // ------------------------------- epilogue DMA (staged bf16 output) ----------
    if (p.staged && elect_one()) {
      const uint32_t cblocks = static_cast<uint32_t>(p.cblocks);
      const int ccols = BN / p.cblocks;
      const uint32_t total = static_cast<uint32_t>(my_tiles) * cblocks;
      const bool has_res = p.residual != nullptr;
      auto coords = [&](uint32_t g, int& m0, int& n0) {
        const int i = static_cast<int>(g / cblocks), cb = static_cast<int>(g % cblocks);
        int m_blk, n_blk;
        tile_coords(i, m_blk, n_blk);
        m0 = m_blk * kTileM;
        n0 = n_blk * BN + cb * ccols;
      };
      auto make_free = [&](uint32_t g) {
        const int buf = g % NB, bi = g % kNBar;
        if (has_res) {
          int m0, n0;
          coords(g, m0, n0);
          mbar_arrive_expect_tx(&cfree[bi], Cfg::kCBytes);
          tma_load_2d(&tmR, &cfree[bi], sC + buf * Cfg::kCBytes, n0, m0);
        } else {
          mbar_arrive(&cfree[bi]);
        }
      };
      for (uint32_t g = 0; g < NB && g < total; ++g) make_free(g);
      for (uint32_t g = 0; g < total; ++g) {
        const int buf = g % NB;
        mbar_wait(&cfull[g % kNBar], (g / kNBar) & 1);
        int m0, n0;
        coords(g, m0, n0);
        if (kConv && p.halo) {
          const int m_blk = m0 / kTileM;
          const int tx = m_blk % p.tiles_x;
          const int ty = (m_blk / p.tiles_x) % p.tiles_y;
          const int img = m_blk / (p.tiles_x * p.tiles_y);
          if (p.pool) tma_store_4d(&tmC, sC + buf * Cfg::kCBytes, n0, tx * (p.tw >> 1), ty * (p.th >> 1), img);
          else tma_store_4d(&tmC, sC + buf * Cfg::kCBytes, n0, tx * p.tw, ty * p.th, img);     // clips ragged tiles
        } else {
          tma_store_2d(&tmC, sC + buf * Cfg::kCBytes, n0, m0);
        }
        tma_store_commit();
        // recycle the buffer of the PREVIOUS block (its store had a whole block time to drain) so the lane
        // never stalls on the store it has just issued
        if (NB <= 2) {
          // one or two staging buffers: release THIS block's buffer as soon as its store has read it.  Recycling the
          // previous block's buffer (below) would hand block g+1 its buffer only after block g is finished, i.e.
          // serialise the two epilogue groups (ncu on the 64-channel halo convs: 30 % of all samples on this wait)
          if (g + NB < total) {
            tma_store_wait_read0();
            make_free(g + NB);
          }
        } else if (g >= 1 && g - 1 + NB < total) {
          tma_store_wait_read1();
          make_free(g - 1 + NB);
        }
      }
      tma_store_wait_all();
    }
  } else {
    // ------------------------------- epilogue (warps 0 .. kEpiWarps-1) ----------
    const int group = warp >> 2;                 // 0 .. kGroups-1
    const int quarter = warp & 3;                // TMEM lane quarter this warp may read
    const bool relu = p.flags & LECB_EPI_RELU;
    const bool gelu = p.flags & LECB_EPI_QUICKGELU;
    const bool mul_gelu_grad = p.flags & LECB_EPI_MUL_QGELU_GRAD;
    const bool res_f32 = p.flags & LECB_EPI_RES_F32;
    const uint32_t erow = static_cast<uint32_t>(quarter * 32 + lane);   // row inside the tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    int bias_n_blk = -1;
    // fused top-10 state (p.topk_val != nullptr): the thread's row keeps its ten best (value, column) pairs, sorted
    float tk_val[10];
    int tk_idx[10];
    int tk_m_blk = -1;
    auto topk_flush = [&](int mb) {
      const int64_t row = static_cast<int64_t>(mb) * kTileM + erow;
      if (row < p.M) {
        float* pv = p.topk_val + (row * p.topk_slots + blockIdx.x) * 10;
        int* pi = p.topk_idx + (row * p.topk_slots + blockIdx.x) * 10;
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          pv[j] = tk_val[j];
          pi[j] = tk_idx[j];
        }
      }
    };
    for (int tile_seq = 0; tile_seq < my_tiles; ++tile_seq) {
      int m_blk, n_blk;
      tile_coords(tile_seq, m_blk, n_blk);
      const int acc = tile_seq & (2 * MT - 1);        // accumulator of this (sub-)tile
      const int par = tile_seq & 1;                     // bias slice parity
      const uint32_t tf_phase = static_cast<uint32_t>(tile_seq / kTF) & 1u;
      const int64_t row = static_cast<int64_t>(m_blk) * kTileM + erow;
      const bool row_ok = row < p.M;
      float ssq = 0.f;
      // The tile's bias slice goes through shared memory: with the whole carve-out given to operand tiles there is
      // no L1 left, so per-block __ldg of the bias cost an exposed L2 round trip per 64 columns (measured: -20 % on
      // K = 768 GEMMs).  One element per epilogue thread, requested before the accumulator wait; double-buffered by
      // tile parity (the named barrier of tile i+1 orders every reader of tile i-1's slice before its overwrite).
      // Narrow tiles (BN <= 64) keep ONE slice that is reloaded only when the n tile changes (every epilogue warp walks
      // the same tile sequence, so the reload and its two barriers are warp-uniform; with a single n tile — the halo
      // convs, N <= 64 GEMMs — it happens once per CTA).
      const float* sb = BN <= 64 ? sbias : sbias + par * 256;
      if (p.bias != nullptr) {
        if (BN <= 64) {
          if (n_blk != bias_n_blk) {
            bias_n_blk = n_blk;
            const int tid = static_cast<int>(threadIdx.x);
            const int col = n_blk * BN + tid;
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");      // everyone is done with the old slice
            if (tid < BN) sbias[tid] = col < p.N ? __ldg(p.bias + col) : 0.f;
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          }
        } else {
          const int tid = static_cast<int>(threadIdx.x);      // 0..383: the epilogue warps
          const int col = n_blk * BN + tid;
          if (tid < BN) sbias[par * 256 + tid] = col < p.N ? __ldg(p.bias + col) : 0.f;
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        }
      }
      if (p.staged && p.out_f32) {
        // fp32 output (+ fp32 residual): 32-column blocks (128-byte rows), same buffer ring and DMA protocol
        const int cblocks = p.cblocks;
        // block g = tile_seq * cblocks + cb belongs to group g % kGroups
        const uint32_t g0 = static_cast<uint32_t>(tile_seq) * cblocks;
        const int cb_first = (group + kGroups - static_cast<int>(g0 % kGroups)) % kGroups;
        if (cb_first >= cblocks) continue;
        mbar_wait(&tfull[tile_seq % kTF], tf_phase);
        tc_fence_after();
#pragma unroll 1
        for (int cb = cb_first; cb < cblocks; cb += kGroups) {
          const uint32_t gblk = g0 + cb;
          const int buf = gblk % NB;
          uint8_t* cbuf = sC + buf * Cfg::kCBytes;
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + lane_base + static_cast<uint32_t>(acc * BN + cb * 32), r);
          tmem_ld_wait();
          tc_fence_before();                     // tempty counts one arrival per (block, warp)
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
          mbar_wait(&cfree[gblk % kNBar], (gblk / kNBar) & 1);
          const int n0 = n_blk * BN + cb * 32;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (p.bias != nullptr) {              // packed FADD2: see the bf16 path below
            const float4* bp = reinterpret_cast<const float4*>(sb + (n0 - n_blk * BN));
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = bp[j / 4];
              const float2 lo = fadd2(make_float2(v[j], v[j + 1]), make_float2(b.x, b.y));
              const float2 hi = fadd2(make_float2(v[j + 2], v[j + 3]), make_float2(b.z, b.w));
              v[j] = lo.x; v[j + 1] = lo.y; v[j + 2] = hi.x; v[j + 3] = hi.y;
            }
          }
          if (p.residual != nullptr) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 f = *reinterpret_cast<const float4*>(cbuf + swizzled_chunk_offset(erow, q, 128));
              const float2 lo = fadd2(make_float2(v[q * 4], v[q * 4 + 1]), make_float2(f.x, f.y));
              const float2 hi = fadd2(make_float2(v[q * 4 + 2], v[q * 4 + 3]), make_float2(f.z, f.w));
              v[q * 4] = lo.x; v[q * 4 + 1] = lo.y; v[q * 4 + 2] = hi.x; v[q * 4 + 3] = hi.y;
            }
          }
          if (relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (gelu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            *reinterpret_cast<float4*>(cbuf + swizzled_chunk_offset(erow, q, 128)) =
                make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
            if (p.row_sumsq != nullptr && n0 + q * 4 < p.N)
              ssq += v[q * 4] * v[q * 4] + v[q * 4 + 1] * v[q * 4 + 1] + v[q * 4 + 2] * v[q * 4 + 2] +
                     v[q * 4 + 3] * v[q * 4 + 3];
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&cfull[gblk % kNBar]);
        }
      } else if (p.staged) {
        const uint32_t g0 = static_cast<uint32_t>(tile_seq) * kCBlocks;
        const int cb_first = (group + kGroups - static_cast<int>(g0 % kGroups)) % kGroups;
        if (cb_first >= kCBlocks) continue;       // no block of this tile belongs to this group
        mbar_wait(&tfull[tile_seq % kTF], tf_phase);
        tc_fence_after();
#pragma unroll 1
        for (int cb = cb_first; cb < kCBlocks; cb += kGroups) {
          const uint32_t gblk = g0 + cb;
          const int buf = gblk % NB;
          uint8_t* cbuf = sC + buf * Cfg::kCBytes;
          // both 32-column halves of the block are fetched from TMEM back to back, then one wait
          uint32_t r[kCCols / 32][32];
#pragma unroll
          for (int half = 0; half < kCCols / 32; ++half)
            tmem_ld_32x32(tmem_base + lane_base + static_cast<uint32_t>(acc * BN + cb * kCCols + half * 32), r[half]);
          tmem_ld_wait();
          tc_fence_before();             // the block is in registers: hand its share of the accumulator back
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
          mbar_wait(&cfree[gblk % kNBar], (gblk / kNBar) & 1);
#pragma unroll
          for (int half = 0; half < kCCols / 32; ++half) {
            const int n0 = n_blk * BN + cb * kCCols + half * 32;
            // column pairs: bias and residual go in with packed FADD2, ReLU is applied to the packed bf16 pair after
            // the rounding (identical result: rounding is monotone and keeps the sign) — 6 instead of 9 ALU
            // instructions per pair on the warps that bound the short-K residual GEMMs
            float2 v2[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v2[j] = make_float2(__uint_as_float(r[half][2 * j]), __uint_as_float(r[half][2 * j + 1]));
            if (p.bias != nullptr) {
              const float4* bp = reinterpret_cast<const float4*>(sb + (n0 - n_blk * BN));
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = bp[j];
                v2[2 * j] = fadd2(v2[2 * j], make_float2(b.x, b.y));
                v2[2 * j + 1] = fadd2(v2[2 * j + 1], make_float2(b.z, b.w));
              }
            }
            if (p.residual != nullptr) {
              if (mul_gelu_grad) {        // `residual` holds the fc pre-activation v: y = (A W^T) * QuickGELU'(v)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint4 u = *reinterpret_cast<const uint4*>(cbuf + swizzled_chunk_offset(erow, half * 4 + q, kCCols * 2));
                  v2[q * 4 + 0] = mul_quick_gelu_grad(v2[q * 4 + 0], unpack_bf16(u.x));
                  v2[q * 4 + 1] = mul_quick_gelu_grad(v2[q * 4 + 1], unpack_bf16(u.y));
                  v2[q * 4 + 2] = mul_quick_gelu_grad(v2[q * 4 + 2], unpack_bf16(u.z));
                  v2[q * 4 + 3] = mul_quick_gelu_grad(v2[q * 4 + 3], unpack_bf16(u.w));
                }
              } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint4 u = *reinterpret_cast<const uint4*>(cbuf + swizzled_chunk_offset(erow, half * 4 + q, kCCols * 2));
                  v2[q * 4 + 0] = fadd2(v2[q * 4 + 0], unpack_bf16(u.x));
                  v2[q * 4 + 1] = fadd2(v2[q * 4 + 1], unpack_bf16(u.y));
                  v2[q * 4 + 2] = fadd2(v2[q * 4 + 2], unpack_bf16(u.z));
                  v2[q * 4 + 3] = fadd2(v2[q * 4 + 3], unpack_bf16(u.w));
                }
              }
            }
            if (gelu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v2[j] = make_float2(quick_gelu(v2[j].x), quick_gelu(v2[j].y));
            }
            if (kConv && p.pool) {
              float v[32];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                v[2 * j] = relu ? fmaxf(v2[j].x, 0.f) : v2[j].x;
                v[2 * j + 1] = relu ? fmaxf(v2[j].y, 0.f) : v2[j].y;
              }
              // 2x2 average in fp32 before the bf16 rounding.  Patch row = y * tw + x (tw = 8 or 16), so the window is
              // lanes {l, l^1, l^tw, l^tw^1}.  Each exchange step sends the half of the columns the lane gives up and
              // keeps the other half (24 shuffles per 32 columns instead of 64); the lane ends up with 8 of the 32
              // pooled columns = one 16-byte chunk of the pooled row, so all 32 lanes store.
              const bool bx = (erow & 1u) != 0, by = (erow & static_cast<uint32_t>(p.tw)) != 0;
              float s1[16], s2[8];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float send = bx ? v[j] : v[16 + j];
                const float keep = bx ? v[16 + j] : v[j];
                s1[j] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float send = by ? s1[j] : s1[8 + j];
                const float keep = by ? s1[8 + j] : s1[j];
                s2[j] = 0.25f * (keep + __shfl_xor_sync(0xffffffffu, send, p.tw));
              }
              const uint32_t px = erow & static_cast<uint32_t>(p.tw - 1), py = erow / static_cast<uint32_t>(p.tw);
              const uint32_t wrow = (py >> 1) * static_cast<uint32_t>(p.tw >> 1) + (px >> 1);
              uint4 u;
              u.x = pack_bf16(s2[0], s2[1]);
              u.y = pack_bf16(s2[2], s2[3]);
              u.z = pack_bf16(s2[4], s2[5]);
              u.w = pack_bf16(s2[6], s2[7]);
              *reinterpret_cast<uint4*>(cbuf + swizzled_chunk_offset(wrow, half * 4 + (bx ? 2 : 0) + (by ? 1 : 0), kCCols * 2)) = u;
              continue;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 u;
              u.x = pack_bf16(v2[q * 4 + 0].x, v2[q * 4 + 0].y);
              u.y = pack_bf16(v2[q * 4 + 1].x, v2[q * 4 + 1].y);
              u.z = pack_bf16(v2[q * 4 + 2].x, v2[q * 4 + 2].y);
              u.w = pack_bf16(v2[q * 4 + 3].x, v2[q * 4 + 3].y);
              if (relu) {
                u.x = relu_bf16x2(u.x);
                u.y = relu_bf16x2(u.y);
                u.z = relu_bf16x2(u.z);
                u.w = relu_bf16x2(u.w);
              }
              *reinterpret_cast<uint4*>(cbuf + swizzled_chunk_offset(erow, half * 4 + q, kCCols * 2)) = u;
              if (p.row_sumsq != nullptr && n0 + q * 8 < p.N) {
                float2 f;
                f = unpack_bf16(u.x); ssq += f.x * f.x + f.y * f.y;
                f = unpack_bf16(u.y); ssq += f.x * f.x + f.y * f.y;
                f = unpack_bf16(u.z); ssq += f.x * f.x + f.y * f.y;
                f = unpack_bf16(u.w); ssq += f.x * f.x + f.y * f.y;
              }
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&cfull[gblk % kNBar]);
        }
      } else if (p.topk_val != nullptr) {
        // ---- fused per-row top-10 (caption retrieval): no output tile at all ----
        if (group != 0) continue;
        if (m_blk != tk_m_blk) {                  // tiles arrive in increasing order: the m block changes at most once or twice
          if (tk_m_blk >= 0) topk_flush(tk_m_blk);
          tk_m_blk = m_blk;
#pragma unroll
          for (int j = 0; j < 10; ++j) {
            tk_val[j] = -INFINITY;
            tk_idx[j] = -1;
          }
        }
        mbar_wait(&tfull[tile_seq % kTF], tf_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + lane_base + static_cast<uint32_t>(acc * BN + c * 32), r);
          tmem_ld_wait();
          const int n0 = n_blk * BN + c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = __uint_as_float(r[j]);
            if (v > tk_val[9] && n0 + j < p.N) {           // strict: of equal values the lower index (seen first) stays
              // branch-free sorted insertion with STATIC indices only: the list must stay in registers (a runtime-indexed
              // array lives in local memory, and with the whole carve-out given to shared memory there is no L1: every
              // access would be an L2 round trip on the epilogue's critical path)
              const int col = n0 + j;
#pragma unroll
              for (int q = 9; q >= 1; --q) {
                const bool above = v > tk_val[q - 1];        // the element above moves down into slot q
                const bool here = v > tk_val[q];             // ... else v lands here if it beats the current occupant
                tk_idx[q] = above ? tk_idx[q - 1] : (here ? col : tk_idx[q]);
                tk_val[q] = above ? tk_val[q - 1] : (here ? v : tk_val[q]);
              }
              if (v > tk_val[0]) {
                tk_val[0] = v;
                tk_idx[0] = col;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
      } else {
        if (group != 0) continue;                 // direct fp32 stores: one group is plenty (small GEMMs)
        mbar_wait(&tfull[tile_seq % kTF], tf_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + lane_base + static_cast<uint32_t>(acc * BN + c * 32), r);
          tmem_ld_wait();
          const int n0 = n_blk * BN + c * 32;
          if (n0 >= p.N) continue;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (p.bias != nullptr) {
            const float4* bp = reinterpret_cast<const float4*>(sb + (n0 - n_blk * BN));
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = bp[j / 4];
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
          if (p.residual != nullptr && row_ok && res_f32) {
            const float4* rp = reinterpret_cast<const float4*>(static_cast<const float*>(p.residual) + row * p.N + n0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (n0 + q * 4 < p.N) {
                const float4 f = __ldg(rp + q);
                v[q * 4 + 0] += f.x; v[q * 4 + 1] += f.y; v[q * 4 + 2] += f.z; v[q * 4 + 3] += f.w;
              }
            }
          } else if (p.residual != nullptr && row_ok) {
            const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.residual) + row * p.N + n0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (n0 + q * 8 < p.N) {
                const uint4 u = __ldg(rp + q);
                float2 f;
                f = unpack_bf16(u.x); v[q * 8 + 0] += f.x; v[q * 8 + 1] += f.y;
                f = unpack_bf16(u.y); v[q * 8 + 2] += f.x; v[q * 8 + 3] += f.y;
                f = unpack_bf16(u.z); v[q * 8 + 4] += f.x; v[q * 8 + 5] += f.y;
                f = unpack_bf16(u.w); v[q * 8 + 6] += f.x; v[q * 8 + 7] += f.y;
              }
            }
          }
          if (relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (gelu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
          }
          if (row_ok) {
            float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + row * p.N + n0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (n0 + q * 4 < p.N) {
                op[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                if (p.row_sumsq != nullptr)
                  ssq += v[q * 4] * v[q * 4] + v[q * 4 + 1] * v[q * 4 + 1] + v[q * 4 + 2] * v[q * 4 + 2] +
                         v[q * 4 + 3] * v[q * 4 + 3];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
      }
      if (p.row_sumsq != nullptr && row_ok) atomicAdd(p.row_sumsq + row, ssq);
    }
    if (p.topk_val != nullptr && group == 0 && tk_m_blk >= 0) topk_flush(tk_m_blk);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(MT * Cfg::kTmemCols));
  }
}

template <int BN, int BK, int NB, bool kConv>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams& p, cudaStream_t stream,
                       const CUtensorMap* tmA2 = nullptr) {
  using Cfg = GemmCfg<BN, BK, NB>;
  constexpr bool kPairable = kConv && BN == 128 && BK == 64 && NB == 2;      // the one instantiation of MT = 2
  static DeviceOnce once;                    // the attribute is per device: one flag per device ordinal
  bool& configured = once.flag();
  auto kern = gemm_kernel<BN, BK, NB, kConv, 1>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess && kPairable)
      e = cudaFuncSetAttribute(gemm_kernel<BN, BK, NB, kConv, kPairable ? 2 : 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(LECB_ERR_CUDA, "cudaFuncSetAttribute(smem=227K): %s", cudaGetErrorString(e));
    configured = true;
  }
  if (p.mt == 2 && !kPairable) p.mt = 1;
  CUtensorMap tmC = tmA, tmR = tmA;      // placeholders when the staged path is off
  const bool want_f32 = (p.flags & LECB_EPI_OUT_F32) != 0;
  if (!want_f32 && (p.flags & LECB_EPI_RES_F32)) return fail(LECB_ERR_ARG, "fp32 residual requires LECB_EPI_OUT_F32");
  // fp32 output is staged too (32-column blocks) when the tile is wide enough and the residual, if any, is fp32
  p.out_f32 = (want_f32 && BN >= 64 && (p.residual == nullptr || (p.flags & LECB_EPI_RES_F32))) ? 1 : 0;
  p.staged = (!want_f32 || p.out_f32) ? 1 : 0;
  if (p.topk_val != nullptr) p.staged = p.out_f32 = 0;      // nothing is stored: the direct-path barrier counts apply
  p.cblocks = p.out_f32 ? BN / 32 : Cfg::kCBlocks;
  if (p.staged) {
    const uint32_t ccols = p.out_f32 ? 32 : Cfg::kCCols;
    const uint32_t esz = p.out_f32 ? 4 : 2;
    int st = p.halo ? (p.pool ? encode_tiled_4d_nhwc(&tmC, p.out, p.batch, p.H / 2, p.W / 2, p.N, ccols, p.tw / 2, p.th / 2)
                              : encode_tiled_4d_nhwc(&tmC, p.out, p.batch, p.H, p.W, p.N, ccols, p.tw, p.th))
                    : encode_tiled_2d_ex(&tmC, p.out, static_cast<uint64_t>(p.M), static_cast<uint64_t>(p.N), kTileM, ccols, esz);
    if (st) return st;
    if (p.residual != nullptr) {
      st = encode_tiled_2d_ex(&tmR, p.residual, static_cast<uint64_t>(p.M), static_cast<uint64_t>(p.N), kTileM, ccols, esz);
      if (st) return st;
    }
  }
  if (p.kb_split > 0) {                  // dual-A mode: the second A operand rides in the (otherwise unused) residual slot
    if (tmA2 == nullptr || p.residual != nullptr) return fail(LECB_ERR_ARG, "dual-A GEMM: second operand missing or residual given");
    tmR = *tmA2;
  }
  // Weight tiles that fit stay resident in shared memory for the CTA's lifetime (64-channel 3x3 convs: 9 x 8 KB;
  // the K <= 256 expand convs: 128 KB): the mainloop then streams only A tiles, which removes the W re-fetch from
  // the L2 -> SM traffic those layers are bound by.  The CTA owns one n tile and strides over m tiles.
  const int sms = sm_count();
  if (sms <= 0) return fail(LECB_ERR_CUDA, "no CUDA device");
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  p.operand_bytes = Cfg::kStages * Cfg::kStageBytes;
  int smem_bytes = Cfg::kSmemBytes;
  int grid = tiles < sms ? tiles : sms;
  if (p.mt == 2) {                        // paired m tiles: stages of two A tiles + one W tile
    const int stage_bytes = 2 * Cfg::kABytes + Cfg::kBBytes;
    p.res_stages = (227 * 1024 - NB * Cfg::kCBytes - kBarBytes - kBiasBytes) / stage_bytes;
    if (p.res_stages > kMaxStages) p.res_stages = kMaxStages;
    if (p.res_stages < 2 || 4 * BN > 512) return fail(LECB_ERR_UNSUPPORTED, "paired-tile mode does not fit (BN=%d BK=%d)", BN, BK);
    p.operand_bytes = p.res_stages * stage_bytes;
    smem_bytes = p.operand_bytes + NB * Cfg::kCBytes + kBarBytes + kBiasBytes;
    const int pairs = (p.num_m_tiles / 2) * p.num_n_tiles;
    grid = pairs < sms ? pairs : sms;
  } else if (p.b_resident) {              // planned by dispatch(): W tile + A ring + NB staging buffers fit
    p.operand_bytes = p.num_kb * Cfg::kBBytes + p.res_stages * (p.halo ? (p.halo_single ? 1 : 3) * p.copy_bytes : Cfg::kABytes);
    smem_bytes = p.operand_bytes + NB * Cfg::kCBytes + kBarBytes + kBiasBytes;
    grid = (sms / p.num_n_tiles) * p.num_n_tiles;
  }
  if (p.topk_val != nullptr) {             // one m block per CTA (see the kernel's tile schedule)
    int per_m = sms / p.num_m_tiles;
    if (per_m > p.num_n_tiles) per_m = p.num_n_tiles;
    if (per_m < 1) return fail(LECB_ERR_ARG, "top-10 epilogue: too many row blocks (%d) for %d SMs", p.num_m_tiles, sms);
    grid = per_m * p.num_m_tiles;
  }
  if (p.mt == 2) launch_k(gemm_kernel<BN, BK, NB, kConv, kPairable ? 2 : 1>, dim3(grid), dim3(kNumThreads), smem_bytes, stream, tmA, tmB, tmC, tmR, p);
  else launch_k(kern, dim3(grid), dim3(kNumThreads), smem_bytes, stream, tmA, tmB, tmC, tmR, p);
  count_launch();
  return check_launch("gemm_kernel");
}

static int pick_bn(int N) {
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  return 256;
}

// NB = number of 16 KB epilogue staging buffers: 4 when the K loop is short (the tile is epilogue/HBM-bound and
// deeper residual prefetch + store overlap matters more than operand stages), else 2.
template <bool kConv, int NB>
static int dispatch_nb(int BN, int BK, const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams& p, cudaStream_t s,
                       const CUtensorMap* tmA2 = nullptr) {
  if (BK == 64) {
    switch (BN) {
      case 32: return launch_gemm<32, 64, NB, kConv>(tmA, tmB, p, s, tmA2);
      case 64: return launch_gemm<64, 64, NB, kConv>(tmA, tmB, p, s, tmA2);
      case 128: return launch_gemm<128, 64, NB, kConv>(tmA, tmB, p, s, tmA2);
      default: return launch_gemm<256, 64, NB, kConv>(tmA, tmB, p, s, tmA2);
    }
  }
  return BN == 32 ? launch_gemm<32, 32, NB, kConv>(tmA, tmB, p, s) : launch_gemm<64, 32, NB, kConv>(tmA, tmB, p, s);
}

// The staging-buffer count trades operand stages for epilogue depth: with a short K loop the tile is bound by
// the residual read + output write, and the number of 16 KB residual blocks that can be in flight per SM
// (each buffer cycles free -> TMA load -> add -> TMA store -> drain) sets the achievable HBM bandwidth.
// A ring depth the resident-W mode would get with `nb` staging buffers (0 = does not fit).
static int resident_ring(int BN, int BK, int num_kb, int nb) {
  const int ccols = BN >= 64 ? 64 : BN;
  const int budget = 227 * 1024 - nb * (kTileM * ccols * 2) - kBarBytes - kBiasBytes;
  const int wbytes = num_kb * BN * BK * 2;
  if (wbytes > 128 * 1024) return 0;
  int ring = (budget - wbytes) / (kTileM * BK * 2);
  return ring > kMaxStages ? kMaxStages : (ring < 0 ? 0 : ring);
}

template <bool kConv>
static int dispatch(int BN, int BK, const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams& p, cudaStream_t s,
                    const CUtensorMap* tmA2 = nullptr) {
  int nb = 2;
  if (!kConv) {
    if (p.num_kb <= 2) nb = 8;               // measured 6 / 8: equal
    else if (p.num_kb <= 4) nb = 5;          // measured 4 / 5 / 8 on layer3's expand conv: 4.91 / 4.56 / 4.52 ms per 23 launches
    else if (p.num_kb <= 8) nb = 4;
    // fp32 output / residual moves twice the bytes per element: keep four blocks in flight up to K = 1024
    else if ((p.flags & LECB_EPI_OUT_F32) && p.num_kb <= 16) nb = 4;
    // QuickGELU costs two MUFU ops per element: the epilogue of a 128x256 tile is then almost as long as a K = 768
    // mainloop, and with two staging buffers it stalls on store drain (measured 971 -> 1164 TF/s with four)
    else if ((p.flags & LECB_EPI_QUICKGELU) && p.num_kb <= 16) nb = 4;
  }
  // Resident-W mode (see launch_gemm): worth it when every CTA walks many m tiles of one n tile AND the staging
  // depth does not have to shrink for it — measured on layer3's expand conv (K 256 -> N 1024, 128 KB W tile):
  // giving up two of the five residual/staging buffers costs more HBM overlap than the W re-fetch saves.
  p.b_resident = 0;
  const int sms = sm_count();
  if (kConv && p.halo) {                  // halo-tile conv: W resident (9 x BN x 64), stages of three halo copies
    auto stages_with = [&](int nbuf) {
      const int budget = 227 * 1024 - nbuf * (kTileM * (BN >= 64 ? 64 : BN) * 2) - kBarBytes - kBiasBytes - 9 * BN * BK * 2;
      const int st = budget / ((p.halo_single ? 1 : 3) * p.copy_bytes);
      return st > kMaxStages ? kMaxStages : st;
    };
    // Cin = 32 (stem convs, 64-byte rows): a tile is 24 KB of HBM traffic against ~0.5 us of MMAs, so the output
    // blocks need the deeper store pipeline more than the halo ring needs a fifth stage
    if (BK == 32) nb = 4;                     // measured 2 / 3 / 4 on the stem convs: 4 is 10-25 % faster
    int stages = stages_with(nb);
    if (stages < 2 && stages_with(1) >= 1) {       // Cout = 128: 144 KB of weights leave room for ONE halo stage and
      nb = 1;                                      // one staging buffer; still ~3x faster than re-fetching per tap
      stages = stages_with(1);
    }
    if (stages < 1) return fail(LECB_ERR_UNSUPPORTED, "halo conv does not fit shared memory (BN=%d copy=%d)", BN, p.copy_bytes);
    p.b_resident = 1;
    p.res_stages = stages;
  } else if (p.mt == 1 && p.num_kb >= 2 && p.num_n_tiles <= 8 && sms > 0 && p.num_m_tiles >= 4 * sms && !env_flag_no_resident()) {
    const int ring = resident_ring(BN, BK, p.num_kb, nb);
    if (ring >= 3) {                          // a two-stage A ring next to a resident 128 KB W tile measured slower than streaming
      p.b_resident = 1;
      p.res_stages = ring;
    }
  }
  if (kConv && nb == 1) return dispatch_nb<kConv, 1>(BN, BK, tmA, tmB, p, s);
  switch (nb) {
    case 8: return dispatch_nb<kConv, 8>(BN, BK, tmA, tmB, p, s, tmA2);
    case 5: return dispatch_nb<kConv, 5>(BN, BK, tmA, tmB, p, s, tmA2);
    case 4: return dispatch_nb<kConv, 4>(BN, BK, tmA, tmB, p, s, tmA2);
    case 3: return dispatch_nb<kConv, 3>(BN, BK, tmA, tmB, p, s, tmA2);
    default: return dispatch_nb<kConv, 2>(BN, BK, tmA, tmB, p, s, tmA2);
  }
}

}  // namespace lecb

using namespace lecb;

extern "C" int lecb_gemm_bf16(const void* A, const void* W, const float* bias, const void* residual, void* out,
                              float* row_sumsq, int64_t M, int N, int K, unsigned flags, void* stream) {
  LECB_CHECK_ARG(A && W && out, "lecb_gemm_bf16: null pointer");
  LECB_CHECK_ARG(M > 0 && N > 0 && K > 0, "lecb_gemm_bf16: empty problem M=%lld N=%d K=%d", (long long)M, N, K);
  LECB_CHECK_ARG(K % 32 == 0, "lecb_gemm_bf16: K=%d must be a multiple of 32", K);
  LECB_CHECK_ARG(N % 8 == 0, "lecb_gemm_bf16: N=%d must be a multiple of 8", N);
  LECB_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
                 "lecb_gemm_bf16: operands must be 16-byte aligned");
  if (flags & LECB_EPI_MUL_QGELU_GRAD)
    LECB_CHECK_ARG(residual != nullptr && !(flags & (LECB_EPI_RELU | LECB_EPI_QUICKGELU | LECB_EPI_OUT_F32 | LECB_EPI_RES_F32)),
                   "lecb_gemm_bf16: LECB_EPI_MUL_QGELU_GRAD needs the bf16 pre-activation in `residual`, a bf16 output and no "
                   "other activation");
  // (a bf16 residual with an fp32 output stays on the single-CTA kernel's direct path)
  if (pair_gemm_eligible(M, N, K, flags, (flags & LECB_EPI_OUT_F32) && residual != nullptr && !(flags & LECB_EPI_RES_F32)))
    return launch_pair_gemm(A, W, bias, residual, out, row_sumsq, M, N, K, flags, static_cast<cudaStream_t>(stream));
  const int BK = (K % 64 == 0) ? 64 : 32;
  int BN = pick_bn(N);
  if (BK == 32 && BN > 64) BN = 64;
  if (BN == 256) {
    // wave quantisation: with few m tiles (text tower, small batches) 128-wide tiles can fill the last wave much
    // better than 256-wide ones; switch when that buys more than 15 %
    const int sms = sm_count();
    if (sms > 0) {
      const int64_t mt = (M + kTileM - 1) / kTileM;
      const int64_t t256 = mt * ((N + 255) / 256), t128 = mt * ((N + 127) / 128);
      const double e256 = static_cast<double>(t256) / (((t256 + sms - 1) / sms) * sms);
      const double e128 = static_cast<double>(t128) / (((t128 + sms - 1) / sms) * sms);
      if (e128 > 1.15 * e256) BN = 128;
    }
  }
  GemmParams p{};
  p.bias = bias;
  p.residual = residual;
  p.out = out;
  p.row_sumsq = row_sumsq;
  p.mt = 1;
  p.M = M;
  p.N = N;
  p.num_kb = K / BK;
  p.num_m_tiles = static_cast<int>((M + kTileM - 1) / kTileM);
  p.num_n_tiles = (N + BN - 1) / BN;
  p.flags = flags;
  CUtensorMap tmA, tmB;
  int st = encode_tiled_2d(&tmA, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), kTileM, BK);
  if (st) return st;
  st = encode_tiled_2d(&tmB, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K), BN, BK);
  if (st) return st;
  return dispatch<false>(BN, BK, tmA, tmB, p, static_cast<cudaStream_t>(stream));
}

// K-concatenated GEMM over two A operands (the projection shortcut of a bottleneck, M:44-52):
//   out[M,N] = epi( A1[M,K1] . W[:, :K1]^T + A2[M,K2] . W[:, K1:]^T + bias ),   W = [W1 | W2] bf16 [N, K1+K2]
// `out = relu(bn3(conv3(y)) + downsample(x))` is one GEMM with [y | x] against [W3 | Wd] and the summed folded biases: the
// shortcut tensor (M x N bf16) is neither written nor read back as a residual.  bf16 output, K1 % 64 == K2 % 64 == 0.
extern "C" int lecb_gemm_bf16_dual(const void* A1, int K1, const void* A2, int K2, const void* W, const float* bias, void* out,
                                   int64_t M, int N, unsigned flags, void* stream) {
  LECB_CHECK_ARG(A1 && A2 && W && out, "lecb_gemm_bf16_dual: null pointer");
  LECB_CHECK_ARG(M > 0 && N > 0 && K1 > 0 && K2 > 0, "lecb_gemm_bf16_dual: empty problem M=%lld N=%d K1=%d K2=%d", (long long)M, N, K1, K2);
  LECB_CHECK_ARG(K1 % 64 == 0 && K2 % 64 == 0, "lecb_gemm_bf16_dual: K1=%d and K2=%d must be multiples of 64", K1, K2);
  LECB_CHECK_ARG(N % 8 == 0, "lecb_gemm_bf16_dual: N=%d must be a multiple of 8", N);
  LECB_CHECK_ARG((flags & ~static_cast<unsigned>(LECB_EPI_RELU)) == 0, "lecb_gemm_bf16_dual: only LECB_EPI_RELU is supported (flags=0x%x)", flags);
  LECB_CHECK_ARG(((reinterpret_cast<uintptr_t>(A1) | reinterpret_cast<uintptr_t>(A2) | reinterpret_cast<uintptr_t>(W) |
                   reinterpret_cast<uintptr_t>(out)) & 15) == 0, "lecb_gemm_bf16_dual: operands must be 16-byte aligned");
  const int K = K1 + K2;
  if (pair_gemm_eligible(M, N, K, flags, false))
    return launch_pair_gemm(A1, W, bias, nullptr, out, nullptr, M, N, K, flags, static_cast<cudaStream_t>(stream), A2, K1);
  const int BK = 64;
  const int BN = pick_bn(N);
  GemmParams p{};
  p.bias = bias;
  p.out = out;
  p.mt = 1;
  p.M = M;
  p.N = N;
  p.num_kb = K / BK;
  p.kb_split = K1 / BK;
  p.num_m_tiles = static_cast<int>((M + kTileM - 1) / kTileM);
  p.num_n_tiles = (N + BN - 1) / BN;
  p.flags = flags;
  CUtensorMap tmA, tmA2, tmB;
  int st = encode_tiled_2d(&tmA, A1, static_cast<uint64_t>(M), static_cast<uint64_t>(K1), kTileM, BK);
  if (st) return st;
  st = encode_tiled_2d(&tmA2, A2, static_cast<uint64_t>(M), static_cast<uint64_t>(K2), kTileM, BK);
  if (st) return st;
  st = encode_tiled_2d(&tmB, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K), BN, BK);
  if (st) return st;
  return dispatch<false>(BN, BK, tmA, tmB, p, static_cast<cudaStream_t>(stream), &tmA2);
}

// Similarity GEMM with the per-row top-10 fused into the epilogue (caption retrieval, T:444-446):
//   sim[m, n] = <q_m, bank_n> with q given as an exact fp16 pair q = hi + lo (A = [hi | lo], fp16 [M, 2K]) against the
//   fp16 bank [N, K]; instead of the [M, N] matrix (225 MB at 256 x 220 000) each CTA writes the ten best (value, index)
//   pairs of every row over the n tiles it processed to slot blockIdx.x of part_val / part_idx [M][slots][10]
//   (slots >= the grid size returned through *slots_used; every slot is first reset to -inf / -1 by a fill kernel here,
//   so slots a CTA never touches for a row block drop out of the merge).
namespace lecb {
__global__ void __launch_bounds__(256) topk_partial_fill_kernel(float* __restrict__ v, int* __restrict__ i, int64_t n) {
  pdl_grid_sync();
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    v[t] = -INFINITY;
    i[t] = -1;
  }
}
}  // namespace lecb

extern "C" int lecb_gemm_topk10(const void* A_hilo, const void* bank, int64_t M, int N, int K, float* part_val, int* part_idx,
                                int slots, int* slots_used, void* stream) {
  LECB_CHECK_ARG(A_hilo && bank && part_val && part_idx, "lecb_gemm_topk10: null pointer");
  LECB_CHECK_ARG(M > 0 && N >= 10 && K > 0 && K % 64 == 0, "lecb_gemm_topk10: need M > 0, N >= 10, K %% 64 == 0 (M=%lld N=%d K=%d)",
                 (long long)M, N, K);
  LECB_CHECK_ARG((reinterpret_cast<uintptr_t>(A_hilo) & 15) == 0 && (reinterpret_cast<uintptr_t>(bank) & 15) == 0,
                 "lecb_gemm_topk10: operands must be 16-byte aligned");
  const int sms = sm_count();
  if (sms <= 0) return fail(LECB_ERR_CUDA, "no CUDA device");
  const int BN = 256, BK = 64;
  GemmParams p{};
  p.mt = 1;
  p.M = M;
  p.N = N;
  p.hilo = 1;
  p.num_kb = 2 * (K / BK);
  p.num_m_tiles = static_cast<int>((M + kTileM - 1) / kTileM);
  p.num_n_tiles = (N + BN - 1) / BN;
  p.flags = LECB_GEMM_F16_OPERANDS;
  p.topk_val = part_val;
  p.topk_idx = part_idx;
  p.topk_slots = slots;
  LECB_CHECK_ARG(p.num_m_tiles <= sms, "lecb_gemm_topk10: at most %d query rows per call (M=%lld)", sms * kTileM, (long long)M);
  int per_m = sms / p.num_m_tiles;                 // CTAs per m block
  if (per_m > p.num_n_tiles) per_m = p.num_n_tiles;
  const int grid = per_m * p.num_m_tiles;
  if (slots_used) *slots_used = grid;
  LECB_CHECK_ARG(slots >= grid, "lecb_gemm_topk10: %d partial slots, the launch needs %d", slots, grid);
  CUtensorMap tmA, tmB;
  int st = encode_tiled_2d(&tmA, A_hilo, static_cast<uint64_t>(M), static_cast<uint64_t>(2) * K, kTileM, BK);
  if (st) return st;
  st = encode_tiled_2d(&tmB, bank, static_cast<uint64_t>(N), static_cast<uint64_t>(K), BN, BK);
  if (st) return st;
  const int64_t nfill = M * slots * 10;
  launch_k(topk_partial_fill_kernel, dim3(static_cast<unsigned>((nfill + 255) / 256 < 1024 ? (nfill + 255) / 256 : 1024)), dim3(256), 0, static_cast<cudaStream_t>(stream), part_val, part_idx, nfill);
  count_launch();
  st = check_launch("topk_partial_fill_kernel");
  if (st) return st;
  return dispatch<false>(BN, BK, tmA, tmB, p, static_cast<cudaStream_t>(stream));
}

// Halo-tile eligibility and patch shape of a 3x3 conv (see lecb_conv3x3_bf16): Cin 32 / 64 (one K block per tap), one n
// tile, at least two patches per SM, and a patch shape (16x8 or 8x16) that tiles the image with <= 15 % waste.
static bool halo_plan(int B, int H, int Wd, int Cin, int BN, int& th, int& tw) {
  const int sms = sm_count();
  if (!((Cin == 64 || Cin == 32) && BN <= 128 && sms > 0) || env_flag_no_halo()) return false;
  auto waste = [&](int a, int b) {       // padded / real pixels for an a x b patch
    return static_cast<double>(((H + a - 1) / a) * a) * (((Wd + b - 1) / b) * b) / (static_cast<double>(H) * Wd);
  };
  th = 16;
  tw = 8;
  if (waste(8, 16) < waste(16, 8)) {
    th = 8;
    tw = 16;
  }
  const int64_t tiles = static_cast<int64_t>(B) * ((Wd + tw - 1) / tw) * ((H + th - 1) / th);
  return tiles >= 2 * sms && tiles < 0x7fffffff && waste(th, tw) <= 1.15;
}

static void conv_tile_shape(int Cin, int Cout, int& BN, int& BK) {
  BK = (Cin % 64 == 0) ? 64 : 32;
  BN = pick_bn(Cout);
  if (BK == 32 && BN > 64) BN = 64;
}

extern "C" int lecb_conv3x3_pool_fusable(int B, int H, int Wd, int Cin, int Cout) {
  if (B <= 0 || H <= 0 || Wd <= 0 || Cin % 32 != 0 || Cout % 8 != 0 || (H & 1) || (Wd & 1)) return 0;
  int BN, BK, th, tw;
  conv_tile_shape(Cin, Cout, BN, BK);
  return ((Cout + BN - 1) / BN == 1 && halo_plan(B, H, Wd, Cin, BN, th, tw)) ? 1 : 0;
}

extern "C" int lecb_conv3x3_bf16(const void* x, const void* w, const float* bias, void* out, int B, int H, int Wd,
                                 int Cin, int Cout, unsigned flags, void* stream) {
  LECB_CHECK_ARG(x && w && out, "lecb_conv3x3_bf16: null pointer");
  LECB_CHECK_ARG(B > 0 && H > 0 && Wd > 0, "lecb_conv3x3_bf16: empty problem");
  LECB_CHECK_ARG(Cin % 32 == 0, "lecb_conv3x3_bf16: Cin=%d must be a multiple of 32", Cin);
  LECB_CHECK_ARG(Cout % 8 == 0, "lecb_conv3x3_bf16: Cout=%d must be a multiple of 8", Cout);
  LECB_CHECK_ARG((flags & (LECB_EPI_OUT_F32 | LECB_EPI_RES_F32 | LECB_GEMM_F16_OPERANDS)) == 0,
                 "lecb_conv3x3_bf16: only LECB_EPI_RELU / LECB_EPI_QUICKGELU / LECB_EPI_AVGPOOL2 are supported");
  const bool want_pool = (flags & LECB_EPI_AVGPOOL2) != 0;
  LECB_CHECK_ARG(!want_pool || (H % 2 == 0 && Wd % 2 == 0), "lecb_conv3x3_bf16: LECB_EPI_AVGPOOL2 needs even H and W (H=%d W=%d)", H, Wd);
  if (pair_conv_eligible(B, H, Wd, Cin, Cout, flags))
    return launch_pair_conv3x3(x, w, bias, out, B, H, Wd, Cin, Cout, flags, static_cast<cudaStream_t>(stream));
  int BN, BK;
  conv_tile_shape(Cin, Cout, BN, BK);
  const int64_t M = static_cast<int64_t>(B) * H * Wd;
  GemmParams p{};
  p.bias = bias;
  p.residual = nullptr;
  p.out = out;
  p.row_sumsq = nullptr;
  p.mt = 1;
  p.N = Cout;
  p.kb_per_tap = Cin / BK;
  p.num_kb = 9 * p.kb_per_tap;
  p.num_n_tiles = (Cout + BN - 1) / BN;
  p.flags = flags;
  p.H = H;
  p.W = Wd;
  p.batch = B;
  CUtensorMap tmA, tmB;
  int st = encode_tiled_2d(&tmB, w, static_cast<uint64_t>(Cout), static_cast<uint64_t>(9) * Cin, BN, BK);
  if (st) return st;
  // Halo-tile mode for the 64-channel layers (the im2col path re-fetches every input pixel once per tap and those
  // layers are bound by that L2 -> SM fill): pick the patch shape that tiles the image with the least waste.
  int th = 0, tw = 0;
  if (p.num_n_tiles == 1 && halo_plan(B, H, Wd, Cin, BN, th, tw)) {
    const int tiles_x = (Wd + tw - 1) / tw, tiles_y = (H + th - 1) / th;
    const int64_t tiles = static_cast<int64_t>(B) * tiles_x * tiles_y;
    {
      p.halo = 1;
      p.pool = want_pool ? 1 : 0;
      p.th = th;
      p.tw = tw;
      p.tiles_x = tiles_x;
      p.tiles_y = tiles_y;
      p.copy_bytes = (th + 2) * tw * Cin * 2;        // rows of Cin * 2 = 128 (64) bytes, a whole number of swizzle atoms
      p.num_m_tiles = static_cast<int>(tiles);
      p.M = tiles * kTileM;
      // Measured (RN101, layer1 conv2): the single-copy layout moves 2.4x fewer bytes and fits five stages, yet runs
      // 18 % SLOWER than three aligned copies — operand groups that straddle 1024-byte swizzle atoms cost the tensor
      // pipe more than the L2 traffic saved.  Kept behind a switch as the record of that experiment.
      if (tw == 8 && Cin == 64 && env_flag_halo_single()) {
        p.halo_single = 1;
        p.copy_bytes = ((th + 2) * (tw + 2) * 128 + 1023) / 1024 * 1024;
        st = encode_tiled_4d_nhwc(&tmA, x, B, H, Wd, Cin, 64, tw + 2, th + 2);
      } else {
        st = encode_tiled_4d_nhwc(&tmA, x, B, H, Wd, Cin, Cin, tw, th + 2);
      }
      if (st) return st;
      return dispatch<true>(BN, BK, tmA, tmB, p, static_cast<cudaStream_t>(stream));
    }
  }
  if (want_pool)
    return fail(LECB_ERR_UNSUPPORTED, "lecb_conv3x3_bf16: LECB_EPI_AVGPOOL2 is fused only in halo-tile mode "
                "(Cin 32/64, one n tile, >= 2 patches per SM); see lecb_conv3x3_pool_fusable()");
  p.M = M;
  p.num_m_tiles = static_cast<int>((M + kTileM - 1) / kTileM);
  // 128-wide convs (Cin >= 128: the halo mode does not apply, W is too large to stay resident) are bound by the
  // L2 -> SM fill of A (once per tap) plus W (once per m tile): pairing m tiles halves the W share
  {
    const int sms = sm_count();
    if (BN == 128 && BK == 64 && p.num_n_tiles == 1 && p.num_m_tiles % 2 == 0 && sms > 0 && p.num_m_tiles >= 4 * sms &&
        !env_flag_no_mt2())
      p.mt = 2;
  }
  st = encode_im2col_3x3(&tmA, x, B, H, Wd, Cin, BK, kTileM);
  if (st) return st;
  return dispatch<true>(BN, BK, tmA, tmB, p, static_cast<cudaStream_t>(stream));
}
