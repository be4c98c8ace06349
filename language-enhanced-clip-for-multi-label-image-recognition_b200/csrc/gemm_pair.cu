// CTA-pair (cta_group::2) variant of the bf16 GEMM / implicit-GEMM 3x3 convolution for the wide layers.
//
//   out[M,N] = epi( A[M,K] · W[N,K]^T + bias (+ residual) ),   N >= 256, K % 64 == 0, bf16 in / bf16 out
//
// Why: ncu on round 1's single-CTA kernel showed every layer3 / layer4 launch of the RN101 trunk sitting at the L2 -> SM
// fill limit (~11 TB/s of lts sectors): with one CTA per 128 x 256 tile each SM re-fetches the whole 256-row W tile for
// every 128 rows of A.  Here two CTAs of a cluster (one TPC) own a 256 x 256 tile: each stages ITS 128 rows of A and HALF
// of the W tile (128 of the 256 n rows) per k block, and one `tcgen05.mma.cta_group::2` (M = 256, issued by the leader
// CTA only) multiplies both A halves against both W halves — the hardware reads the peer's shared memory — leaving each
// CTA's 128 x 256 accumulator in its own TMEM.  W crosses L2 -> SM once per 256 output rows: operand fill per flop drops by
// a third for N = 256 (A 16 KB + W 16 KB per CTA per k block instead of 16 + 32), and a stage is 32 KB instead of 48 KB.
// (tools/micro/mma_pair.cu is the probe that pinned the conventions: M = 256 instruction descriptor, cta_group::2 TMEM
// allocation, multicast commit; its M=256,N=256,K=16 MMA issues in 128 cycles, the same as the 1-CTA M=128 one.)
//
// Protocol (per CTA: 8 epilogue warps, TMA producer warp, MMA warp, epilogue DMA warp — as in gemm_tcgen05.cu):
//   full[s]   lives in the LEADER: both CTAs' producers arrive on it (count 2, the peer remotely) with their own byte
//             counts, and both CTAs' TMA loads (cta_group::2 form) complete their bytes on it
//   empty[s]  per CTA: the leader's tcgen05.commit multicasts the arrival to both CTAs when the MMAs that read stage s retire
//   tfull[a]  per CTA: multicast commit after a tile's last MMA — each CTA's epilogue drains its own 128 rows
//   tempty[a] lives in the leader: every epilogue warp of BOTH CTAs arrives (the peer's remotely) once its block is in registers
// Epilogue: identical to the staged bf16 path of gemm_tcgen05.cu (bias slice through smem, TMA-prefetched residual blocks,
// packed adds, ReLU / QuickGELU, swizzled 64-column staging blocks, TMA stores by the DMA warp), per CTA on its own rows.
#include <stdio.h>
#include <stdlib.h>

#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

constexpr int kPM = 128;                       // rows per CTA (256 per pair)
constexpr int kPN = 256;                       // columns per pair tile
constexpr int kPK = 64;
constexpr int kPEpiWarps = 8;
constexpr int kPGroups = 2;
constexpr int kPThreads = 32 * (kPEpiWarps + 3);
constexpr int kPWarpTma = kPEpiWarps, kPWarpMma = kPEpiWarps + 1, kPWarpDma = kPEpiWarps + 2;
constexpr int kPABytes = kPM * kPK * 2;        // 16 KB: this CTA's rows of A
constexpr int kPBBytes = (kPN / 2) * kPK * 2;  // 16 KB: this CTA's half of the W tile
constexpr int kPStageBytes = kPABytes + kPBBytes;
constexpr int kPCBytes = kPM * 64 * 2;         // one staged 64-column output block
constexpr int kPCBlocks = kPN / 64;
constexpr int kPBarBytes = 512;
constexpr int kPBiasBytes = 2048;
constexpr int kPMaxStages = 8;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;    // clears the CTA-rank bit of a shared::cluster address: the leader's copy

template <int NB>
struct PairCfg {
  static constexpr int kStagesRaw = (227 * 1024 - NB * kPCBytes - kPBarBytes - kPBiasBytes) / kPStageBytes;
  static constexpr int kStages = kStagesRaw > kPMaxStages ? kPMaxStages : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kPStageBytes + NB * kPCBytes + kPBarBytes + kPBiasBytes;
  static_assert(kStages >= 3, "operand ring too shallow");
};

struct PairParams {
  const float* bias;
  const void* residual;
  float* row_sumsq;
  int64_t M;
  int N;
  int num_kb;
  int num_m_pairs;     // ceil(M / 256)
  int num_n_tiles;     // ceil(N / 256)
  unsigned flags;
  int H, W, kb_per_tap;   // conv mode
  int kb_split;           // > 0: two A operands (lecb_gemm_bf16_dual): k blocks >= kb_split come from A2, whose tensor map travels
                          // in the residual slot (the mode has no residual)
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (+ expect_tx) on the LEADER CTA's copy of a barrier; a no-op mask on the leader itself
__device__ __forceinline__ void mbar_arrive_expect_tx_leader(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(smem_u32(bar) & kPeerMask), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}
// TMA loads of a CTA pair: bytes complete on the leader's barrier
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* m, uint64_t* bar, void* smem, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_im2col_4d(const CUtensorMap* m, uint64_t* bar, void* smem, int c, int w, int h, int n,
                                                    uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerMask), "r"(c), "r"(w), "r"(h),
      "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` of BOTH CTAs when complete
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
// kind::f16 instruction descriptor, M = 256 across the pair, bf16 operands, fp32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_pair(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}

// kF32: fp32 output (+ optional fp32 residual: the transformers' residual stream, M:226-227) in 32-column staged blocks
// (128-byte rows) instead of bf16 output (+ bf16 residual) in 64-column blocks
template <int NB, bool kConv, bool kF32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const PairParams p) {
  using Cfg = PairCfg<NB>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kCCols = kF32 ? 32 : 64;                // columns per staged block (128-byte rows either way)
  constexpr int kCBlocks = kPN / kCCols;                // 4 or 8 blocks per CTA tile: even, so block cb belongs to group cb % 2
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                                   // [kStages][16 KB]
  uint8_t* sB = smem + kStages * kPABytes;              // [kStages][16 KB]
  uint8_t* sC = smem + kStages * kPStageBytes;          // [NB][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sC + NB * kPCBytes);
  uint64_t* full = bars;                                // leader's copies are the live ones
  uint64_t* empty = bars + kPMaxStages;
  uint64_t* tfull = bars + 2 * kPMaxStages;             // [2]
  uint64_t* tempty = tfull + 2;                         // [2], leader's copies are the live ones
  constexpr int kNBar = NB < kPGroups ? kPGroups : NB;
  uint64_t* cfree = tempty + 2;
  uint64_t* cfull = cfree + kNBar;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cfull + kNBar);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kPBarBytes);      // [2][256]

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = static_cast<int>(blockIdx.x) >> 1;
  const int num_clusters = static_cast<int>(gridDim.x) >> 1;
  const int num_tiles = p.num_m_pairs * p.num_n_tiles;
  const int my_tiles = cluster_id < num_tiles ? (num_tiles - cluster_id + num_clusters - 1) / num_clusters : 0;
  // tile i of this cluster: n fastest (neighbouring clusters share the A rows through L2)
  auto tile_coords = [&](int i, int& m_pair, int& n_blk) {
    const int tile = cluster_id + i * num_clusters;
    m_pair = tile / p.num_n_tiles;
    n_blk = tile - m_pair * p.num_n_tiles;
  };

  if (warp == kPWarpTma && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kPMaxStages; ++i) {
      mbar_init(&full[i], 2);                  // one arrival per CTA of the pair
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 2 * 4 * kCBlocks);         // every (block, epilogue warp) of both CTAs
    }
    for (int i = 0; i < kNBar; ++i) {
      mbar_init(&cfree[i], 1);
      mbar_init(&cfull[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == kPWarpMma) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();            // prologue done while the previous kernel drained; no global access before this point
  if (my_tiles <= 1) pdl_trigger();        // at most one work item: it is the last one (else: the producer, below)

  if (warp == kPWarpTma) {
    // ------------------------------- TMA producer (both CTAs) -------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        int m_pair, n_blk;
        tile_coords(ti, m_pair, n_blk);
        if (ti + 1 == my_tiles && my_tiles > 1) pdl_trigger();         // last work item of a longer run (lecb_common.cuh)
        const int64_t m0 = (static_cast<int64_t>(m_pair) * 2 + rank) * kPM;         // this CTA's first row
        int pw0 = 0, ph0 = 0, pn0 = 0;
        if (kConv) {
          const int hw = p.H * p.W;
          pn0 = static_cast<int>(m0 / hw);
          const int rem0 = static_cast<int>(m0 - static_cast<int64_t>(pn0) * hw);
          ph0 = rem0 / p.W;
          pw0 = rem0 - ph0 * p.W;
        }
        const int n0 = n_blk * kPN + static_cast<int>(rank) * (kPN / 2);             // this CTA's half of the W rows
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx_leader(&full[stage], kPStageBytes);
          if (kConv) {
            const int tap = kb / p.kb_per_tap;
            const int cb = kb - tap * p.kb_per_tap;
            const int ky = tap / 3, kx = tap - ky * 3;
            // rows past the end of the tensor (ragged last pair): the image index is out of range -> zero fill
            tma2_load_im2col_4d(&tmA, &full[stage], sA + stage * kPABytes, cb * kPK, pw0 - 1, ph0 - 1, pn0,
                                static_cast<uint16_t>(kx), static_cast<uint16_t>(ky));
          } else {
            const bool second = p.kb_split > 0 && kb >= p.kb_split;
            tma2_load_2d(second ? &tmR : &tmA, &full[stage], sA + stage * kPABytes, (second ? kb - p.kb_split : kb) * kPK,
                         static_cast<int>(m0));
          }
          tma2_load_2d(&tmB, &full[stage], sB + stage * kPBBytes, kb * kPK, n0);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == kPWarpMma) {
    // ------------------------------- MMA issuer (leader CTA only) ---------------------------
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = make_idesc_pair(kPN);
      int stage = 0;
      uint32_t phase = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int acc = ti & 1;
        mbar_wait(&tempty[acc], ((static_cast<uint32_t>(ti) >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kPN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = make_kmajor_desc(smem_u32(sA + stage * kPABytes), kPK * 2);
          const uint64_t bdesc = make_kmajor_desc(smem_u32(sB + stage * kPBBytes), kPK * 2);
#pragma unroll
          for (int k = 0; k < kPK / 16; ++k)
            umma2_f16(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma2_commit(&empty[stage]);           // both CTAs' producers may refill the stage
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma2_commit(&tfull[acc]);               // both CTAs' epilogues
      }
    }
  } else if (warp == kPWarpDma) {
    // ------------------------------- epilogue DMA (per CTA, own rows) -----------------------
    if (elect_one()) {
      const uint32_t total = static_cast<uint32_t>(my_tiles) * kCBlocks;
      const bool has_res = p.residual != nullptr;
      auto coords = [&](uint32_t g, int& m0, int& n0) {
        const int i = static_cast<int>(g / kCBlocks), cb = static_cast<int>(g % kCBlocks);
        int m_pair, n_blk;
        tile_coords(i, m_pair, n_blk);
        m0 = (m_pair * 2 + static_cast<int>(rank)) * kPM;
        n0 = n_blk * kPN + cb * kCCols;
      };
      auto make_free = [&](uint32_t g) {
        const int buf = g % NB, bi = g % kNBar;
        if (has_res) {
          int m0, n0;
          coords(g, m0, n0);
          mbar_arrive_expect_tx(&cfree[bi], kPCBytes);
          tma_load_2d(&tmR, &cfree[bi], sC + buf * kPCBytes, n0, m0);
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
        tma_store_2d(&tmC, sC + buf * kPCBytes, n0, m0);      // rows / columns past the tensor are clipped
        tma_store_commit();
        if (NB <= 2) {
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
    // ------------------------------- epilogue (warps 0..7, per CTA, own rows) ---------------
    const int group = warp >> 2;
    const int quarter = warp & 3;
    const bool relu = p.flags & LECB_EPI_RELU;
    const bool gelu = p.flags & LECB_EPI_QUICKGELU;
    const bool mul_gelu_grad = p.flags & LECB_EPI_MUL_QGELU_GRAD;
    const uint32_t erow = static_cast<uint32_t>(quarter * 32 + lane);
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    for (int tile_seq = 0; tile_seq < my_tiles; ++tile_seq) {
      int m_pair, n_blk;
      tile_coords(tile_seq, m_pair, n_blk);
      const int acc = tile_seq & 1;
      const int par = tile_seq & 1;
      const int64_t row = (static_cast<int64_t>(m_pair) * 2 + rank) * kPM + erow;
      const bool row_ok = row < p.M;
      float ssq = 0.f;
      const float* sb = sbias + par * 256;
      if (p.bias != nullptr) {
        const int tid = static_cast<int>(threadIdx.x);          // 0..255: the epilogue warps
        const int col = n_blk * kPN + tid;
        sbias[par * 256 + tid] = col < p.N ? __ldg(p.bias + col) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(kPEpiWarps * 32) : "memory");
      }
      const uint32_t g0 = static_cast<uint32_t>(tile_seq) * kCBlocks;
      mbar_wait(&tfull[acc], (static_cast<uint32_t>(tile_seq) >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int cb = group; cb < kCBlocks; cb += kPGroups) {
        const uint32_t gblk = g0 + cb;
        const int buf = gblk % NB;
        uint8_t* cbuf = sC + buf * kPCBytes;
        if constexpr (kF32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + lane_base + static_cast<uint32_t>(acc * kPN + cb * 32), r);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tempty[acc]);
          mbar_wait(&cfree[gblk % kNBar], (gblk / kNBar) & 1);
          const int n0 = n_blk * kPN + cb * 32;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (p.bias != nullptr) {
            const float4* bp = reinterpret_cast<const float4*>(sb + (n0 - n_blk * kPN));
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
              ssq += v[q * 4] * v[q * 4] + v[q * 4 + 1] * v[q * 4 + 1] + v[q * 4 + 2] * v[q * 4 + 2] + v[q * 4 + 3] * v[q * 4 + 3];
          }
        } else {
        uint32_t r[2][32];
#pragma unroll
        for (int half = 0; half < 2; ++half)
          tmem_ld_32x32(tmem_base + lane_base + static_cast<uint32_t>(acc * kPN + cb * 64 + half * 32), r[half]);
        tmem_ld_wait();
        tc_fence_before();             // the block is in registers: hand its share of the accumulator back (to the leader)
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tempty[acc]);
        mbar_wait(&cfree[gblk % kNBar], (gblk / kNBar) & 1);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int n0 = n_blk * kPN + cb * 64 + half * 32;
          float2 v2[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v2[j] = make_float2(__uint_as_float(r[half][2 * j]), __uint_as_float(r[half][2 * j + 1]));
          if (p.bias != nullptr) {
            const float4* bp = reinterpret_cast<const float4*>(sb + (n0 - n_blk * kPN));
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
                const uint4 u = *reinterpret_cast<const uint4*>(cbuf + swizzled_chunk_offset(erow, half * 4 + q, 128));
                v2[q * 4 + 0] = mul_quick_gelu_grad(v2[q * 4 + 0], unpack_bf16(u.x));
                v2[q * 4 + 1] = mul_quick_gelu_grad(v2[q * 4 + 1], unpack_bf16(u.y));
                v2[q * 4 + 2] = mul_quick_gelu_grad(v2[q * 4 + 2], unpack_bf16(u.z));
                v2[q * 4 + 3] = mul_quick_gelu_grad(v2[q * 4 + 3], unpack_bf16(u.w));
              }
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 u = *reinterpret_cast<const uint4*>(cbuf + swizzled_chunk_offset(erow, half * 4 + q, 128));
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
            *reinterpret_cast<uint4*>(cbuf + swizzled_chunk_offset(erow, half * 4 + q, 128)) = u;
            if (p.row_sumsq != nullptr && n0 + q * 8 < p.N) {
              float2 f;
              f = unpack_bf16(u.x); ssq += f.x * f.x + f.y * f.y;
              f = unpack_bf16(u.y); ssq += f.x * f.x + f.y * f.y;
              f = unpack_bf16(u.z); ssq += f.x * f.x + f.y * f.y;
              f = unpack_bf16(u.w); ssq += f.x * f.x + f.y * f.y;
            }
          }
        }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&cfull[gblk % kNBar]);
      }
      if (p.row_sumsq != nullptr && row_ok) atomicAdd(p.row_sumsq + row, ssq);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // neither CTA frees TMEM / exits while the pair's MMAs or remote arrivals are in flight
  if (warp == kPWarpMma) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int NB, bool kConv, bool kF32>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                       const PairParams& p, cudaStream_t stream) {
  using Cfg = PairCfg<NB>;
  auto kern = gemm_pair_kernel<NB, kConv, kF32>;
  static DeviceOnce once;
  bool& configured = once.flag();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(LECB_ERR_CUDA, "cudaFuncSetAttribute(pair smem=227K): %s", cudaGetErrorString(e));
    configured = true;
  }
  const int sms = sm_count();
  if (sms <= 0) return fail(LECB_ERR_CUDA, "no CUDA device");
  const int tiles = p.num_m_pairs * p.num_n_tiles;
  int clusters = sms / 2;
  if (clusters > tiles) clusters = tiles;
  launch_k(kern, dim3(2 * clusters), dim3(kPThreads), Cfg::kSmemBytes, stream, tmA, tmB, tmC, tmR, p);
  count_launch();
  return check_launch("gemm_pair_kernel");
}

template <bool kConv, bool kF32>
static int dispatch_pair(int nb, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                         const PairParams& p, cudaStream_t s) {
  switch (nb) {
    case 5: return launch_pair<5, kConv, kF32>(tmA, tmB, tmC, tmR, p, s);
    case 4: return launch_pair<4, kConv, kF32>(tmA, tmB, tmC, tmR, p, s);
    default: return launch_pair<2, kConv, kF32>(tmA, tmB, tmC, tmR, p, s);
  }
}

// Eligibility: wide bf16-out layers with enough 256 x 256 tiles to fill the machine.  LECB_NO_PAIR=1 switches the path
// off (A/B runs of the same build).
static int g_pair_mode = -1;           // -1: not initialised (LECB_NO_PAIR decides), 0: off, 1: on

bool pair_gemm_eligible(int64_t M, int N, int K, unsigned flags, bool has_sumsq_f32_out) {
  if (g_pair_mode < 0) g_pair_mode = getenv("LECB_NO_PAIR") != nullptr ? 0 : 1;
  if (g_pair_mode == 0 || has_sumsq_f32_out) return false;
  if (flags & (LECB_GEMM_F16_OPERANDS | LECB_EPI_AVGPOOL2)) return false;
  if ((flags & LECB_EPI_RES_F32) && !(flags & LECB_EPI_OUT_F32)) return false;
  if (N < kPN || N % 8 != 0 || K % kPK != 0 || K < 4 * kPK) return false;      // K <= 128 layers already run at their HBM bound
  const int sms = sm_count();
  if (sms <= 1) return false;
  const int64_t tiles = ((M + 2 * kPM - 1) / (2 * kPM)) * ((N + kPN - 1) / kPN);
  return tiles >= sms;                 // at least two tiles per cluster
}

static int staging_buffers(int num_kb, bool has_res) {
  if (num_kb <= 4) return 5;           // short K: the tile is bound by its residual read + output write
  if (num_kb <= 8 || has_res) return 4;
  return 2;
}

int launch_pair_gemm(const void* A, const void* Wt, const float* bias, const void* residual, void* out, float* row_sumsq,
                     int64_t M, int N, int K, unsigned flags, cudaStream_t stream, const void* A2, int K1) {
  if (A2 != nullptr && (residual != nullptr || K1 <= 0 || K1 >= K || K1 % kPK != 0 || (flags & LECB_EPI_OUT_F32)))
    return fail(LECB_ERR_ARG, "pair GEMM, dual-A mode: bad split K1=%d of K=%d, residual or fp32 output given", K1, K);
  PairParams p{};
  p.kb_split = A2 != nullptr ? K1 / kPK : 0;
  p.bias = bias;
  p.residual = residual;
  p.row_sumsq = row_sumsq;
  p.M = M;
  p.N = N;
  p.num_kb = K / kPK;
  p.num_m_pairs = static_cast<int>((M + 2 * kPM - 1) / (2 * kPM));
  p.num_n_tiles = (N + kPN - 1) / kPN;
  p.flags = flags;
  CUtensorMap tmA, tmB, tmC, tmR;
  int st = encode_tiled_2d(&tmA, A, static_cast<uint64_t>(M), static_cast<uint64_t>(A2 != nullptr ? K1 : K), kPM, kPK);
  if (st) return st;
  st = encode_tiled_2d(&tmB, Wt, static_cast<uint64_t>(N), static_cast<uint64_t>(K), kPN / 2, kPK);
  if (st) return st;
  const bool f32 = (flags & LECB_EPI_OUT_F32) != 0;
  const uint32_t ccols = f32 ? 32 : 64, esz = f32 ? 4 : 2;
  st = encode_tiled_2d_ex(&tmC, out, static_cast<uint64_t>(M), static_cast<uint64_t>(N), kPM, ccols, esz);
  if (st) return st;
  tmR = tmC;
  if (residual != nullptr) {
    st = encode_tiled_2d_ex(&tmR, residual, static_cast<uint64_t>(M), static_cast<uint64_t>(N), kPM, ccols, esz);
    if (st) return st;
  }
  if (A2 != nullptr) {                   // the second A operand rides in the residual slot
    st = encode_tiled_2d(&tmR, A2, static_cast<uint64_t>(M), static_cast<uint64_t>(K - K1), kPM, kPK);
    if (st) return st;
  }
  // fp32 blocks carry half the columns: twice the blocks in flight for the same bytes
  const int nb = f32 ? (p.num_kb <= 16 ? 4 : 2) : staging_buffers(p.num_kb, residual != nullptr);
  if (f32) return dispatch_pair<false, true>(nb, tmA, tmB, tmC, tmR, p, stream);
  return dispatch_pair<false, false>(nb, tmA, tmB, tmC, tmR, p, stream);
}

bool pair_conv_eligible(int B, int H, int Wd, int Cin, int Cout, unsigned flags) {
  const int64_t M = static_cast<int64_t>(B) * H * Wd;
  // whole pairs only: a CTA whose 128 pixels lie entirely past the tensor would start its im2col walk out of range
  if (Cin % kPK != 0 || M % (2 * kPM) != 0) return false;
  return pair_gemm_eligible(M, Cout, 9 * Cin, flags, false);
}

int launch_pair_conv3x3(const void* x, const void* w, const float* bias, void* out, int B, int H, int Wd, int Cin, int Cout,
                        unsigned flags, cudaStream_t stream) {
  PairParams p{};
  p.bias = bias;
  p.M = static_cast<int64_t>(B) * H * Wd;
  p.N = Cout;
  p.kb_per_tap = Cin / kPK;
  p.num_kb = 9 * p.kb_per_tap;
  p.num_m_pairs = static_cast<int>((p.M + 2 * kPM - 1) / (2 * kPM));
  p.num_n_tiles = (Cout + kPN - 1) / kPN;
  p.flags = flags;
  p.H = H;
  p.W = Wd;
  CUtensorMap tmA, tmB, tmC;
  int st = encode_im2col_3x3(&tmA, x, B, H, Wd, Cin, kPK, kPM);
  if (st) return st;
  st = encode_tiled_2d(&tmB, w, static_cast<uint64_t>(Cout), static_cast<uint64_t>(9) * Cin, kPN / 2, kPK);
  if (st) return st;
  st = encode_tiled_2d(&tmC, out, static_cast<uint64_t>(p.M), static_cast<uint64_t>(Cout), kPM, 64);
  if (st) return st;
  return dispatch_pair<true, false>(2, tmA, tmB, tmC, tmC, p, stream);
}

}  // namespace lecb

// Switch the CTA-pair (cta_group::2) path of lecb_gemm_bf16 / lecb_conv3x3_bf16 on or off at run time (A/B measurements and
// parity tests of both kernels on the same shapes); returns the previous setting.
extern "C" int lecb_set_pair_gemm(int enable) {
  if (lecb::g_pair_mode < 0) lecb::g_pair_mode = getenv("LECB_NO_PAIR") != nullptr ? 0 : 1;
  const int prev = lecb::g_pair_mode;
  lecb::g_pair_mode = enable ? 1 : 0;
  return prev;
}
