// Multi-head self-attention forward on tcgen05 tensor cores (head dim 64): the ViT visual encoder's
// full attention (M:207-228 inside M:240-276) and, with `causal`, the text transformer's masked
// attention (M:364-370).
//
// Work item = one (image, head, 128-query tile); persistent CTAs, two co-resident per SM, loop over items.  K / V arrive as 128-key TMA
// tiles (128B swizzle) and are consumed in 64-key sub-blocks with double-buffered S (TMEM) and P (smem), so
// the tensor pipe computes Q K^T of sub-block j+1 while the softmax warps work on sub-block j:
//   S = Q K^T      tcgen05.mma  M=128 N=64 K=64, fp32 S in TMEM
//   P = exp2(S*c - m*c)   four softmax warps, thread == query row (tcgen05.ld 32x32b), row sums in registers,
//                         P packed to bf16 pairs and written back to TMEM (tcgen05.st): it never touches smem
//   O += P V       tcgen05.mma  M=128 N=64 K=64, A = P from TMEM, V tile used in place as an MN-major B operand
// O accumulates in TMEM across key blocks.  The softmax reference point m is only moved when a block's
// probabilities would grow past 2^15 relative to it (lazy rescaling): block 0 takes an exact two-pass max,
// later sub-blocks are a single pass over S with the running m, and only a warp that sees a row sum above the
// threshold re-does the block exactly and rescales its 32 rows of O in TMEM (tcgen05.ld / tcgen05.st) before
// the block's PV MMA is released.  The result o / l is independent of where m sits.
// TMEM: S0 [0,64) | S1 [64,128) | O [128,192) | P0 [192,224) | P1 [224,256).
//
// Warps: 0-3 softmax (TMEM lane quarter == warp id), 4 TMA producer, 5 MMA issuer + TMEM owner.
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

constexpr int kAtTile = 128;        // queries per CTA and keys per TMA tile
constexpr int kAtSub = 64;          // keys per softmax / MMA sub-block
constexpr int kAtDh = 64;
constexpr int kAtThreads = 192;
constexpr int kAtTileBytes = kAtTile * kAtDh * 2;    // 16 KB
constexpr int kAtSmemQ = 0;                          // 2 buffers (next item's Q prefetched)
constexpr int kAtSmemK = 2 * kAtTileBytes;           // 2 stages
constexpr int kAtSmemV = 4 * kAtTileBytes;           // 2 stages
constexpr int kAtSmemBars = 6 * kAtTileBytes;
constexpr int kAtSmemBytes = kAtSmemBars + 192;
constexpr int kAtTmemCols = 256;

struct AttnParams {
  __nv_bfloat16* out;
  int T, W, q_rows, causal;
  int q_tiles, heads, total_items;
  float sc;            // log2(e) / sqrt(dh)
};

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* smem, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA pipe: round-to-nearest split x = n + f (magic-number add: no F2I / FRND, which share the MUFU's quarter-rate
// datapath), cubic in f on [-0.5, 0.5] (relative error 7.7e-5 = 4 % of half a bf16 ulp, below the rounding the probabilities
// get anyway; tools/micro/exp2_poly.py), n added into the exponent field.  The kernel is bound by the 16 ex2 per clock and SM of
// the MUFU (one per score); evaluating every kPoly-th score here moves that share of the work to the 128-lane FMA pipe.
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -125.0f);                              // keeps 2^n normal; such probabilities are zero after the bf16 rounding
  const float t = x + 12582912.0f;                    // 1.5 * 2^23: the low mantissa bits now hold round(x)
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.0550886838f, f, 0.242604051f);
  p = fmaf(p, f, 0.693276242f);
  p = fmaf(p, f, 0.99992894f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// kind::f16 instruction descriptor with selectable B major-ness (bit 16: 1 = MN-major)
__host__ __device__ constexpr uint32_t attn_idesc(uint32_t n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// kPoly: 0 = every exponential on the MUFU; n > 0 = scores whose column index is n-1 mod n use poly_exp2 (FMA pipe)
template <int kPoly>
__global__ void __launch_bounds__(kAtThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem + kAtSmemQ;     // 2 buffers: the next work item's Q tile is prefetched
  uint8_t* sK = smem + kAtSmemK;
  uint8_t* sV = smem + kAtSmemV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAtSmemBars);
  uint64_t* q_full = bars;         // [2]
  uint64_t* q_empty = bars + 2;    // [2] last Q K^T of the item issued and complete
  uint64_t* k_full = bars + 4;     // [2] 128-key K tiles
  uint64_t* k_empty = bars + 6;    // [2]
  uint64_t* v_full = bars + 8;     // [2]
  uint64_t* v_empty = bars + 10;   // [2]
  uint64_t* s_full = bars + 12;    // [2] S buffer holds Q K^T of a 64-key sub-block
  uint64_t* s_free = bars + 14;    // [2] softmax has read it
  uint64_t* p_full = bars + 16;    // [2] P block written
  uint64_t* p_free = bars + 18;    // [2] PV MMA of the sub-block complete: P reusable, O includes the sub-block
  uint64_t* o_free = bars + 20;    // softmax warps have read the finished item's O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;

  // Persistent CTA: work item w = (image b, head h, query tile qt), w = blockIdx.x + i * gridDim.x.  Every
  // barrier parity below is derived from RUNNING counters (items, K/V tiles, sub-blocks) that all three roles
  // advance identically, so the pipelines never drain between items: the producer prefetches the next item's
  // Q / K / V and the tensor pipe starts its Q K^T while the softmax warps are still storing the previous O.
  auto decode = [&](int w, int& q0, int& h, int& b, int& n_sub, int& n_kv) {
    const int qt = w % p.q_tiles;
    h = (w / p.q_tiles) % p.heads;
    b = w / (p.q_tiles * p.heads);
    q0 = qt * kAtTile;
    const int kv_end = p.causal ? min(p.T, q0 + kAtTile) : p.T;
    n_sub = (kv_end + kAtSub - 1) / kAtSub;      // 64-key sub-blocks
    n_kv = (n_sub + 1) / 2;                      // 128-key K / V tiles
  };

  if (warp == 4 && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();     // swizzled tiles need 1024-byte alignment
    tma_prefetch_desc(&tmQKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&p_free[i], 1);
    }
    mbar_init(o_free, 4);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, kAtTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();            // prologue done while the previous kernel drained; no global access before this point
  const bool one_item = static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) >= p.total_items;
  if (one_item) pdl_trigger();             // at most one work item: it is the last one (else: the producer, below)

  if (warp == 4) {
    // ------------------------------- TMA producer -------------------------------
    if (elect_one()) {
      int it = 0, kvc = 0;
      for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++it) {
        int q0, h, b, n_sub, n_kv;
        decode(w, q0, h, b, n_sub, n_kv);
        if (!one_item && w + static_cast<int>(gridDim.x) >= p.total_items) pdl_trigger();      // last item of a longer run
        const int qb = it & 1;
        mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[qb], kAtTileBytes);
        tma_load_3d(&tmQKV, &q_full[qb], sQ + qb * kAtTileBytes, h * kAtDh, q0, b);
        for (int j = 0; j < n_kv; ++j, ++kvc) {
          const int s = kvc & 1;
          const uint32_t ph = (kvc >> 1) & 1;
          mbar_wait(&k_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[s], kAtTileBytes);
          tma_load_3d(&tmQKV, &k_full[s], sK + s * kAtTileBytes, p.W + h * kAtDh, j * kAtTile, b);
          mbar_wait(&v_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[s], kAtTileBytes);
          tma_load_3d(&tmQKV, &v_full[s], sV + s * kAtTileBytes, 2 * p.W + h * kAtDh, j * kAtTile, b);
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------- MMA issuer ---------------------------------
    // Sub-block g (running index) uses S/P buffer g & 1; sub-block jj of an item uses the (jj & 1) half of K/V tile
    // kvc + (jj >> 1).  Q K^T of sub-block jj+1 is issued before P V of sub-block jj, so the tensor pipe always runs
    // one S ahead of the softmax warps.
    if (elect_one()) {
      constexpr uint32_t idesc_qk = attn_idesc(kAtSub, false);
      constexpr uint32_t idesc_pv = attn_idesc(kAtDh, true);
      const uint32_t tO = tmem_base + 128u;
      int it = 0, kvc = 0, sc = 0;
      for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++it) {
        int q0, h, b, n_sub, n_kv;
        decode(w, q0, h, b, n_sub, n_kv);
        const int qb = it & 1;
        auto issue_pv = [&](int i) {
          const int g = sc + i, bf = g & 1, half = i & 1, tile = kvc + (i >> 1), s = tile & 1;
          if (half == 0) mbar_wait(&v_full[s], (tile >> 1) & 1);
          mbar_wait(&p_full[bf], (g >> 1) & 1);
          if (i == 0 && it >= 1) mbar_wait(o_free, (it - 1) & 1);      // the previous item's O has been read out
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < kAtSub / 16; ++kk) {
            // A = P block in TMEM: bf16 pairs packed per column, 16 keys = 8 columns per step
            const uint32_t tP = tmem_base + 192u + static_cast<uint32_t>(bf * (kAtSub / 2) + kk * 8);
            // B = V tile [keys][dh] used in place as an MN-major operand: 16 keys = 2048 bytes per step
            const uint64_t bdesc = make_kmajor_desc(smem_u32(sV + s * kAtTileBytes + half * (kAtSub * 128) + kk * 2048), 128);
            umma_f16_ts(tO, tP, bdesc, idesc_pv, (i | kk) != 0 ? 1u : 0u);
          }
          if (half == 1 || i == n_sub - 1) umma_commit(&v_empty[s]);
          umma_commit(&p_free[bf]);
        };
        mbar_wait(&q_full[qb], (it >> 1) & 1);
        for (int jj = 0; jj < n_sub; ++jj) {
          const int g = sc + jj, bf = g & 1, half = jj & 1, tile = kvc + (jj >> 1), s = tile & 1;
          if (half == 0) mbar_wait(&k_full[s], (tile >> 1) & 1);
          if (g >= 2) mbar_wait(&s_free[bf], ((g >> 1) - 1) & 1);
          tc_fence_after();
          const uint64_t adesc = make_kmajor_desc(smem_u32(sQ + qb * kAtTileBytes), 128);
          const uint64_t bdesc = make_kmajor_desc(smem_u32(sK + s * kAtTileBytes + half * (kAtSub * 128)), 128);
#pragma unroll
          for (int k = 0; k < kAtDh / 16; ++k)
            umma_f16(tmem_base + static_cast<uint32_t>(bf * kAtSub), adesc + static_cast<uint64_t>(2 * k),
                     bdesc + static_cast<uint64_t>(2 * k), idesc_qk, k != 0 ? 1u : 0u);
          if (half == 1 || jj == n_sub - 1) umma_commit(&k_empty[s]);
          if (jj == n_sub - 1) umma_commit(&q_empty[qb]);
          umma_commit(&s_full[bf]);
          if (jj >= 1) issue_pv(jj - 1);
        }
        issue_pv(n_sub - 1);
        sc += n_sub;
        kvc += n_kv;
      }
    }
  } else {
    // ------------------------------- softmax warps (thread == query row) --------
    const int row = warp * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t tO = tmem_base + lane_base + 128u;

    // exact row maximum of the (masked) 64-key sub-block in S buffer `tS`
    auto block_max = [&](uint32_t tS, bool need_mask, int lim) {
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(tS, ra);
      tmem_ld_32x32(tS + 32u, rb);
      tmem_ld_wait();
      float mx = -INFINITY;
      if (need_mask) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          mx = fmaxf(mx, (i <= lim) ? __uint_as_float(ra[i]) : -INFINITY);
          mx = fmaxf(mx, (32 + i <= lim) ? __uint_as_float(rb[i]) : -INFINITY);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(ra[i]), __uint_as_float(rb[i])));
      }
      return mx;
    };
    // P = exp2(S*sc - msc) -> bf16 K-major operand block in shared memory; returns the row sum.  The second
    // 32-column TMEM load is in flight while the first half is exponentiated; `wait_bar` (P block free) is
    // taken just before the first store so that wait hides behind the first half's math.
    // kHalf (ragged blocks only): no row of this warp sees a key of the second 32-column half — those 32 exponentials per
    // row are skipped and the half of P is written as zeros
    auto write_p_impl = [&](auto mask_tag, auto half_tag, uint32_t tS, uint32_t tP, float msc, int lim, uint64_t* wait_bar,
                            uint32_t wait_parity) {
      constexpr bool kMask = decltype(mask_tag)::value;
      constexpr bool kHalf = decltype(half_tag)::value;
      float sum = 0.f;
      uint32_t ra[32], rb[32], pk[32];
      tmem_ld_32x32(tS, ra);
      tmem_ld_wait();
      if constexpr (!kHalf) tmem_ld_32x32(tS + 32u, rb);
      if (kHalf) {
#pragma unroll
        for (int i = 16; i < 32; ++i) pk[i] = 0u;
      }
#pragma unroll
      for (int c2 = 0; c2 < (kHalf ? 1 : 2); ++c2) {
        uint32_t(&r)[32] = c2 ? rb : ra;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float x0 = fmaf(__uint_as_float(r[i]), p.sc, -msc), x1 = fmaf(__uint_as_float(r[i + 1]), p.sc, -msc);
          float e0 = (kPoly > 0 && i % kPoly == kPoly - 1) ? poly_exp2(x0) : fast_exp2(x0);
          float e1 = (kPoly > 0 && (i + 1) % kPoly == kPoly - 1) ? poly_exp2(x1) : fast_exp2(x1);
          if (kMask && c2 * 32 + i > lim) e0 = 0.f;
          if (kMask && c2 * 32 + i + 1 > lim) e1 = 0.f;
          sum += e0 + e1;
          pk[c2 * 16 + i / 2] = pack_bf16(e0, e1);
        }
        if (c2 == 0 && !kHalf) tmem_ld_wait();
      }
      if (wait_bar != nullptr) mbar_wait(wait_bar, wait_parity);      // PV of sub-block jj-2 has consumed this P buffer
      tmem_st_32x32(tP, pk);
      tmem_st_wait();
      return sum;
    };
    // the mask test costs two extra instructions per score: only the ragged last block / causal diagonal pays it
    auto write_p = [&](uint32_t tS, uint32_t tP, float msc, bool need_mask, int lim, uint64_t* wait_bar,
                       uint32_t wait_parity) {
      // ragged last block (T = 1 + patches is never a multiple of 64: 17 live keys of 64 at ViT-B/16 @448, ONE at ViT-L/14) and the
      // causal diagonal: warp-uniform choice of the half-block variant
      if (need_mask && __reduce_max_sync(0xffffffffu, lim) < 32)
        return write_p_impl(std::true_type{}, std::true_type{}, tS, tP, msc, lim, wait_bar, wait_parity);
      return need_mask ? write_p_impl(std::true_type{}, std::false_type{}, tS, tP, msc, lim, wait_bar, wait_parity)
                       : write_p_impl(std::false_type{}, std::false_type{}, tS, tP, msc, lim, wait_bar, wait_parity);
    };

    int sc = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x) {
      int q0, h, b, n_sub, n_kv;
      decode(w, q0, h, b, n_sub, n_kv);
      const int qi = q0 + row;
      const int limit = p.causal ? min(p.T - 1, qi) : p.T - 1;       // last key index this row may see
      float m = -INFINITY, l = 0.f;
      // ragged last query tile (17 live rows of 128 at T = 785, one at T = 1025; rows 77.. of the text tower): a warp whose 32 rows
      // are all past q_rows keeps the barrier protocol — it is paced by s_full, so it can never run a phase ahead of the live
      // warps — but does none of the softmax work; its rows of P / O hold garbage that is never stored (MMA rows are independent)
      const bool warp_live = q0 + warp * 32 < p.q_rows;
      for (int jj = 0; jj < n_sub; ++jj) {
        const int g = sc + jj, bf = g & 1;
        const uint32_t par = (g >> 1) & 1;
        if (!warp_live) {
          mbar_wait(&s_full[bf], par);
          if (lane == 0) {
            mbar_arrive(&p_full[bf]);
            mbar_arrive(&s_free[bf]);
          }
          continue;
        }
        const int kv0 = jj * kAtSub;
        const bool need_mask = (kv0 + kAtSub > p.T) || (p.causal && kv0 + kAtSub - 1 > q0);
        const int lim = limit - kv0;             // columns c <= lim are visible
        const uint32_t tS = tmem_base + lane_base + static_cast<uint32_t>(bf * kAtSub);
        const uint32_t pblk = tmem_base + lane_base + 192u + static_cast<uint32_t>(bf * (kAtSub / 2));
        uint64_t* pf = g >= 2 ? &p_free[bf] : nullptr;          // PV of sub-block g-2 read this P buffer
        mbar_wait(&s_full[bf], par);
        tc_fence_after();
        if (jj == 0) {
          m = block_max(tS, need_mask, lim);
          l = write_p(tS, pblk, m * p.sc, need_mask, lim, pf, par ^ 1);
        } else {
          float sum = write_p(tS, pblk, m * p.sc, need_mask, lim, pf, par ^ 1);
          // lazy rescaling: keep m unless some probability of this sub-block is enormous relative to it
          if (__any_sync(0xffffffffu, !(sum <= 32768.f))) {
            const float m_new = fmaxf(m, block_max(tS, need_mask, lim));
            const float alpha = fast_exp2((m - m_new) * p.sc);
            // every PV up to sub-block g-1 must have landed in O; PV of sub-block g is not released yet, so
            // this warp's 32 rows of O can be rescaled in place
            mbar_wait(&p_free[bf ^ 1], ((g - 1) >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int hlf = 0; hlf < 2; ++hlf) {
              uint32_t r[32];
              tmem_ld_32x32(tO + static_cast<uint32_t>(hlf * 32), r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
              tmem_st_32x32(tO + static_cast<uint32_t>(hlf * 32), r);
            }
            tmem_st_wait();
            sum = write_p(tS, pblk, m_new * p.sc, need_mask, lim, nullptr, 0);
            l *= alpha;
            m = m_new;
          }
          l += sum;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&p_full[bf]);        // before s_free: once s_free completes (and the next S of this buffer can be issued, which
          mbar_arrive(&s_free[bf]);        // is what paces a warp without live rows) every p_full arrival of the sub-block is in
        }
      }
      // all PV MMAs of the item done -> O complete (the last two sub-blocks' commits cover every earlier MMA)
      const int gl = sc + n_sub - 1;
      if (gl >= 1) mbar_wait(&p_free[(gl - 1) & 1], ((gl - 1) >> 1) & 1);
      mbar_wait(&p_free[gl & 1], (gl >> 1) & 1);
      tc_fence_after();
      {      // the tcgen05.ld is warp-collective: every lane loads, only valid query rows store
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(tO, r0);
        tmem_ld_32x32(tO + 32u, r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);          // O is in registers: the next item's first PV may overwrite it
        if (qi < p.q_rows) {
          const float inv = 1.0f / l;
          uint4* op = reinterpret_cast<uint4*>(p.out + (static_cast<int64_t>(b) * p.T + qi) * p.W + h * kAtDh);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 u;
            u.x = pack_bf16(__uint_as_float(r0[q * 8 + 0]) * inv, __uint_as_float(r0[q * 8 + 1]) * inv);
            u.y = pack_bf16(__uint_as_float(r0[q * 8 + 2]) * inv, __uint_as_float(r0[q * 8 + 3]) * inv);
            u.z = pack_bf16(__uint_as_float(r0[q * 8 + 4]) * inv, __uint_as_float(r0[q * 8 + 5]) * inv);
            u.w = pack_bf16(__uint_as_float(r0[q * 8 + 6]) * inv, __uint_as_float(r0[q * 8 + 7]) * inv);
            op[q] = u;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 u;
            u.x = pack_bf16(__uint_as_float(r1[q * 8 + 0]) * inv, __uint_as_float(r1[q * 8 + 1]) * inv);
            u.y = pack_bf16(__uint_as_float(r1[q * 8 + 2]) * inv, __uint_as_float(r1[q * 8 + 3]) * inv);
            u.z = pack_bf16(__uint_as_float(r1[q * 8 + 4]) * inv, __uint_as_float(r1[q * 8 + 5]) * inv);
            u.w = pack_bf16(__uint_as_float(r1[q * 8 + 6]) * inv, __uint_as_float(r1[q * 8 + 7]) * inv);
            op[4 + q] = u;
          }
        }
      }
      sc += n_sub;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAtTmemCols);
  }
}

}  // namespace lecb

using namespace lecb;

namespace lecb {
static int g_attn_poly = -1;        // -1: not initialised (LECB_ATTN_POLY or the default), else 0 / 2 / 3 / 4
// Default 0 since the ragged-edge skipping (dead query warps, half key blocks): the MUFU is no longer the co-bottleneck it was, and
// with every exponential on it the kernel is 3-4 % faster than with a quarter on the FMA pipe (456-462 vs 476-492 us per
// ViT-B/16 layer, 499-508 vs 525-528 at ViT-L/14; before the skipping: 473 vs 449 — profiles/r02_attn_poly_ab.jsonl)
constexpr int kAttnPolyDefault = 0;

template <int kPoly>
static int launch_attn(const CUtensorMap& tm, const AttnParams& p, int grid, cudaStream_t stream) {
  static DeviceOnce once;                    // the attribute is per device: one flag per device ordinal
  bool& configured = once.flag();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<kPoly>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemBytes);
    if (e != cudaSuccess) return fail(LECB_ERR_CUDA, "cudaFuncSetAttribute(attn smem=%d): %s", kAtSmemBytes, cudaGetErrorString(e));
    // two CTAs per SM need the full 228 KB shared-memory carve-out
    cudaFuncSetAttribute(attn_fwd_kernel<kPoly>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  launch_k(attn_fwd_kernel<kPoly>, dim3(grid), dim3(kAtThreads), kAtSmemBytes, stream, tm, p);
  count_launch();
  return check_launch("attn_fwd_kernel");
}
}  // namespace lecb

// Share of the attention softmax's exponentials evaluated by the polynomial on the FMA pipe: 0 = none, n in {2, 3, 4} = every
// n-th score.  Returns the previous setting (A/B measurements; environment LECB_ATTN_POLY sets the initial value).
extern "C" int lecb_set_attn_poly(int n) {
  if (lecb::g_attn_poly < 0) {
    const char* e = getenv("LECB_ATTN_POLY");
    lecb::g_attn_poly = e ? atoi(e) : lecb::kAttnPolyDefault;
  }
  const int prev = lecb::g_attn_poly;
  if (n == 0 || n == 2 || n == 3 || n == 4) lecb::g_attn_poly = n;
  return prev;
}

extern "C" int lecb_attn_fwd(const void* qkv, void* out, int B, int T, int W, int heads, int q_rows, int causal,
                             void* stream) {
  LECB_CHECK_ARG(qkv && out, "lecb_attn_fwd: null pointer");
  LECB_CHECK_ARG(B > 0 && T > 0 && heads > 0 && W == heads * kAtDh, "lecb_attn_fwd: W=%d must equal heads*64 (heads=%d)", W, heads);
  LECB_CHECK_ARG(q_rows > 0 && q_rows <= T, "lecb_attn_fwd: q_rows=%d out of range (T=%d)", q_rows, T);
  LECB_CHECK_ARG(B <= 65535 && heads <= 65535, "lecb_attn_fwd: grid too large");
  LECB_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "lecb_attn_fwd: operands must be 16-byte aligned");
  CUtensorMap tm;
  int st = encode_tiled_3d(&tm, qkv, static_cast<uint64_t>(3) * W, static_cast<uint64_t>(T), static_cast<uint64_t>(B),
                           kAtDh, kAtTile);
  if (st) return st;
  AttnParams p;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.T = T;
  p.W = W;
  p.q_rows = q_rows;
  p.causal = causal;
  p.sc = 1.4426950408889634f / 8.0f;
  p.q_tiles = (q_rows + kAtTile - 1) / kAtTile;
  p.heads = heads;
  const int64_t total = static_cast<int64_t>(p.q_tiles) * heads * B;
  LECB_CHECK_ARG(total < 0x7fffffff, "lecb_attn_fwd: too many work items");
  p.total_items = static_cast<int>(total);
  const int sms = sm_count();
  if (sms <= 0) return fail(LECB_ERR_CUDA, "no CUDA device");
  const int grid = p.total_items < 2 * sms ? p.total_items : 2 * sms;        // persistent: two CTAs per SM
  if (g_attn_poly < 0) lecb_set_attn_poly(-1);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // short sequences (the text tower's 77 tokens: a single key block per row) are latency-bound, not MUFU-bound: measured
  // 27.7 us with every exponential on the MUFU against 28.3 with a quarter on the FMA pipe — they keep the plain kernel
  switch (T >= 256 ? g_attn_poly : 0) {
    case 2: return launch_attn<2>(tm, p, grid, s);
    case 3: return launch_attn<3>(tm, p, grid, s);
    case 4: return launch_attn<4>(tm, p, grid, s);
    default: return launch_attn<0>(tm, p, grid, s);
  }
}
