// Multi-head self-attention forward on tcgen05 tensor cores (head dim 64): the ViT visual encoder's
// full attention (M:207-228 inside M:240-276) and, with `causal`, the text transformer's masked
// attention (M:364-370).
//
// One CTA = one (image, head, 128-query tile); two CTAs are co-resident per SM so one tile's softmax
// overlaps the other's MMAs.  Per 128-key block:
//   S = Q K^T      tcgen05.mma  M=128 N=128 K=64, Q/K tiles TMA-loaded (128B swizzle), fp32 S in TMEM
//   P = exp2(S*c - m*c)   four softmax warps, thread == query row (tcgen05.ld 32x32b), online max / sum,
//                         P written to shared memory as the bf16 K-major A operand of the next MMA
//   O_blk = P V    tcgen05.mma  M=128 N=64 K=128, V tile used in place as an MN-major B operand
// O accumulates in TMEM across key blocks.  The softmax reference point m is only moved when a block's
// probabilities would grow past 2^15 relative to it (lazy rescaling): block 0 takes an exact two-pass max,
// later blocks are a single pass over S with the running m, and only a warp that sees a row sum above the
// threshold re-does the block exactly and rescales its 32 rows of O in TMEM (tcgen05.ld / tcgen05.st) before
// the block's PV MMA is released.  The result o / l is independent of where m sits.
// TMEM: S [0,128) | O [128,192).
//
// Warps: 0-3 softmax (TMEM lane quarter == warp id), 4 TMA producer, 5 MMA issuer + TMEM owner.
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

constexpr int kAtTile = 128;        // queries per CTA and keys per block
constexpr int kAtDh = 64;
constexpr int kAtThreads = 192;
constexpr int kAtTileBytes = kAtTile * kAtDh * 2;    // 16 KB
constexpr int kAtSmemQ = 0;
constexpr int kAtSmemK = kAtTileBytes;               // 2 stages
constexpr int kAtSmemV = 3 * kAtTileBytes;           // 2 stages
constexpr int kAtSmemP = 5 * kAtTileBytes;           // 2 blocks of 64 keys (128 x 64 bf16 each)
constexpr int kAtSmemBars = 7 * kAtTileBytes;
constexpr int kAtSmemBytes = kAtSmemBars + 128;
constexpr int kAtTmemCols = 256;

struct AttnParams {
  __nv_bfloat16* out;
  int T, W, q_rows, causal;
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

// kind::f16 instruction descriptor with selectable B major-ness (bit 16: 1 = MN-major)
__host__ __device__ constexpr uint32_t attn_idesc(uint32_t n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__global__ void __launch_bounds__(kAtThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem + kAtSmemQ;
  uint8_t* sK = smem + kAtSmemK;
  uint8_t* sV = smem + kAtSmemV;
  uint8_t* sP = smem + kAtSmemP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAtSmemBars);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* k_empty = bars + 3;    // [2]
  uint64_t* v_full = bars + 5;     // [2]
  uint64_t* v_empty = bars + 7;    // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* s_free = bars + 10;
  uint64_t* p_full = bars + 11;
  uint64_t* p_free = bars + 12;    // PV MMA of the block complete: P reusable, O includes the block
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAtTile;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int kv_end = p.causal ? min(p.T, q0 + kAtTile) : p.T;
  const int n_kv = (kv_end + kAtTile - 1) / kAtTile;

  if (warp == 4 && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();     // swizzled tiles need 1024-byte alignment
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(p_full, 4);
    mbar_init(p_free, 1);
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

  if (warp == 4) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kAtTileBytes);
      tma_load_3d(&tmQKV, q_full, sQ, h * kAtDh, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], kAtTileBytes);
        tma_load_3d(&tmQKV, &k_full[s], sK + s * kAtTileBytes, p.W + h * kAtDh, j * kAtTile, b);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], kAtTileBytes);
        tma_load_3d(&tmQKV, &v_full[s], sV + s * kAtTileBytes, 2 * p.W + h * kAtDh, j * kAtTile, b);
      }
    }
  } else if (warp == 5) {
    // ------------------------------- MMA issuer ---------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc_qk = attn_idesc(kAtTile, false);
      constexpr uint32_t idesc_pv = attn_idesc(kAtDh, true);
      const uint32_t tS = tmem_base;
      auto issue_pv = [&](int i) {
        const int s = i & 1;
        mbar_wait(&v_full[s], (i >> 1) & 1);
        mbar_wait(p_full, i & 1);
        tc_fence_after();
        const uint32_t tO = tmem_base + 128u;
#pragma unroll
        for (int kk = 0; kk < kAtTile / 16; ++kk) {
          // A = P: two 64-key K-major blocks, 16 keys (32 bytes) per step inside a block
          const uint64_t adesc = make_kmajor_desc(smem_u32(sP + (kk >> 2) * kAtTileBytes), 128) + static_cast<uint64_t>(2 * (kk & 3));
          // B = V tile [keys][dh] used as an MN-major operand: 16 keys = 2048 bytes per step
          const uint64_t bdesc = make_kmajor_desc(smem_u32(sV + s * kAtTileBytes + kk * 2048), 128);
          umma_f16(tO, adesc, bdesc, idesc_pv, (i | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&v_empty[s]);
        umma_commit(p_free);
      };
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        mbar_wait(&k_full[s], (j >> 1) & 1);
        if (j >= 1) mbar_wait(s_free, (j - 1) & 1);
        tc_fence_after();
        const uint64_t adesc = make_kmajor_desc(smem_u32(sQ), 128);
        const uint64_t bdesc = make_kmajor_desc(smem_u32(sK + s * kAtTileBytes), 128);
#pragma unroll
        for (int k = 0; k < kAtDh / 16; ++k)
          umma_f16(tS, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(&k_empty[s]);
        umma_commit(s_full);
        if (j >= 1) issue_pv(j - 1);
      }
      issue_pv(n_kv - 1);
    }
  } else {
    // ------------------------------- softmax warps (thread == query row) --------
    const int row = warp * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const int qi = q0 + row;
    const int limit = p.causal ? min(p.T - 1, qi) : p.T - 1;       // last key index this row may see
    float m = -INFINITY, l = 0.f;
    const uint32_t tS = tmem_base + lane_base;
    const uint32_t tO = tmem_base + lane_base + 128u;

    // exact row maximum of the (masked) block
    auto block_max = [&](bool need_mask, int lim) {
      float mx = -INFINITY;
#pragma unroll 1
      for (int c4 = 0; c4 < 4; ++c4) {
        uint32_t r[32];
        tmem_ld_32x32(tS + static_cast<uint32_t>(c4 * 32), r);
        tmem_ld_wait();
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (c4 * 32 + i <= lim) ? __uint_as_float(r[i]) : -INFINITY);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      return mx;
    };
    // P = exp2(S*sc - msc) -> bf16 K-major operand in shared memory; returns the row sum.  The TMEM load of
    // chunk c+1 is in flight while chunk c is exponentiated.  `wait_bar` (P buffer free) is taken just before
    // the first store so the wait hides behind the first chunk's math.
    auto write_p_impl = [&](auto mask_tag, float msc, int lim, uint64_t* wait_bar, uint32_t wait_parity) {
      constexpr bool kMask = decltype(mask_tag)::value;
      float sum = 0.f;
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(tS, ra);
      tmem_ld_wait();
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        uint32_t(&r)[32] = (c4 & 1) ? rb : ra;
        uint32_t(&rn)[32] = (c4 & 1) ? ra : rb;
        if (c4 < 3) tmem_ld_32x32(tS + static_cast<uint32_t>((c4 + 1) * 32), rn);
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float e = fast_exp2(fmaf(__uint_as_float(r[i]), p.sc, -msc));
          if (kMask && c4 * 32 + i > lim) e = 0.f;
          pv[i] = e;
          sum += e;
        }
        if (c4 == 0 && wait_bar != nullptr) mbar_wait(wait_bar, wait_parity);
        uint8_t* pblk = sP + (c4 >> 1) * kAtTileBytes;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack_bf16(pv[q * 8 + 0], pv[q * 8 + 1]);
          u.y = pack_bf16(pv[q * 8 + 2], pv[q * 8 + 3]);
          u.z = pack_bf16(pv[q * 8 + 4], pv[q * 8 + 5]);
          u.w = pack_bf16(pv[q * 8 + 6], pv[q * 8 + 7]);
          *reinterpret_cast<uint4*>(pblk + swizzled_chunk_offset(row, (c4 & 1) * 4 + q, 128)) = u;
        }
        if (c4 < 3) tmem_ld_wait();
      }
      return sum;
    };
    // the mask test costs two extra instructions per score: only the ragged last block / causal diagonal pays it
    auto write_p = [&](float msc, bool need_mask, int lim, uint64_t* wait_bar, uint32_t wait_parity) {
      return need_mask ? write_p_impl(std::true_type{}, msc, lim, wait_bar, wait_parity)
                       : write_p_impl(std::false_type{}, msc, lim, wait_bar, wait_parity);
    };

    for (int j = 0; j < n_kv; ++j) {
      const int kv0 = j * kAtTile;
      const bool need_mask = (kv0 + kAtTile > p.T) || (p.causal && kv0 + kAtTile - 1 > q0);
      const int lim = limit - kv0;             // columns c <= lim are visible
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      if (j == 0) {
        m = block_max(need_mask, lim);
        l = write_p(m * p.sc, need_mask, lim, nullptr, 0);
      } else {
        float sum = write_p(m * p.sc, need_mask, lim, p_free, (j - 1) & 1);
        // lazy rescaling: keep m unless some probability of this block is enormous relative to it
        if (__any_sync(0xffffffffu, !(sum <= 32768.f))) {
          const float m_new = fmaxf(m, block_max(need_mask, lim));
          const float alpha = fast_exp2((m - m_new) * p.sc);
          // PV of block j-1 has completed (p_free waited in write_p) and PV of block j is not released yet:
          // this warp's 32 rows of O can be rescaled in place
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
          sum = write_p(m_new * p.sc, need_mask, lim, nullptr, 0);
          l *= alpha;
          m = m_new;
        }
        l += sum;
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(s_free);
        mbar_arrive(p_full);
      }
    }
    // all PV MMAs done -> O complete
    mbar_wait(p_free, (n_kv - 1) & 1);
    tc_fence_after();
    {      // the tcgen05.ld is warp-collective: every lane loads, only valid query rows store
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(tO, r0);
      tmem_ld_32x32(tO + 32u, r1);
      tmem_ld_wait();
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

extern "C" int lecb_attn_fwd(const void* qkv, void* out, int B, int T, int W, int heads, int q_rows, int causal,
                             void* stream) {
  LECB_CHECK_ARG(qkv && out, "lecb_attn_fwd: null pointer");
  LECB_CHECK_ARG(B > 0 && T > 0 && heads > 0 && W == heads * kAtDh, "lecb_attn_fwd: W=%d must equal heads*64 (heads=%d)", W, heads);
  LECB_CHECK_ARG(q_rows > 0 && q_rows <= T, "lecb_attn_fwd: q_rows=%d out of range (T=%d)", q_rows, T);
  LECB_CHECK_ARG(B <= 65535 && heads <= 65535, "lecb_attn_fwd: grid too large");
  LECB_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "lecb_attn_fwd: operands must be 16-byte aligned");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemBytes);
    if (e != cudaSuccess) return fail(LECB_ERR_CUDA, "cudaFuncSetAttribute(attn smem=%d): %s", kAtSmemBytes, cudaGetErrorString(e));
    // two CTAs per SM need the full 228 KB shared-memory carve-out
    cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (getenv("LECB_DEBUG")) {
      int occ = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, attn_fwd_kernel, kAtThreads, kAtSmemBytes);
      fprintf(stderr, "[lecb] attn_fwd_kernel: %d CTAs/SM, %d B dynamic smem\n", occ, kAtSmemBytes);
    }
    configured = true;
  }
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
  dim3 grid((q_rows + kAtTile - 1) / kAtTile, heads, B);
  attn_fwd_kernel<<<grid, kAtThreads, kAtSmemBytes, static_cast<cudaStream_t>(stream)>>>(tm, p);
  count_launch();
  return check_launch("attn_fwd_kernel");
}
