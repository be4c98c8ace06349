// Causal self-attention backward on tcgen05 tensor cores (text transformer, M:221-223 with the mask of M:364-370;
// the data-gradient half of loss.backward() through nn.MultiheadAttention — all CLIP weights are frozen, T:763-765).
//
// One CTA = one (sequence, head), L <= 128 tokens, head dim 64.  Q, K, V and dO tiles arrive by TMA (128B swizzle,
// rows past L zero-filled).  Five MMAs, all with M = 128 and fp32 accumulators in TMEM:
//   S  = Q K^T          A = Q  (K-major)    B = K  (K-major)     N = 128
//   dP = dO V^T         A = dO (K-major)    B = V  (K-major)     N = 128
//   -- four warps, thread == query row: P = softmax(S / 8) under the causal mask, D = rowsum(P . dP),
//      dS = P . (dP - D) / 8; P and dS are written to shared memory as bf16 [i][j] tiles --
//   dV = P^T dO         A = P  (MN-major)   B = dO (MN-major)    N = 64, K = i
//   dK = dS^T Q         A = dS (MN-major)   B = Q  (MN-major)    N = 64, K = i
//   dQ = dS K           A = dS (K-major)    B = K  (MN-major)    N = 64, K = j
// The transposed operands are never materialised: the same shared-memory tiles are re-read through MN-major
// descriptors.  K loops stop at ceil(L / 16) steps.
// TMEM (256 columns, so two CTAs share an SM): S [0,128) | dP [128,256) for the first two MMAs; once the row math has
// read them, dV [0,64) | dK [64,128) | dQ [128,192) reuse the same columns.  Shared memory (112 KB, two CTAs per SM):
// Q | K | dO | P(keys 0-63) | V | dS, and the second 64-key block of P is written over V, which only the dP MMA reads.
// One CTA per SM (round 1: 512 columns, 128 KB) left every phase — TMA latency, MMA, row math, epilogue — exposed.
#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

constexpr int kAbTile = 128;
constexpr int kAbDh = 64;
constexpr int kAbThreads = 160;                       // warps 0-3 row math + epilogue, warp 4 TMA + MMA issue
constexpr int kAbTileBytes = kAbTile * kAbDh * 2;     // 16 KB
constexpr int kAbSmemQ = 0, kAbSmemK = kAbTileBytes, kAbSmemO = 2 * kAbTileBytes;
constexpr int kAbSmemP = 3 * kAbTileBytes;            // [128 x 128] bf16 as two 64-column blocks; the second one IS the V tile
constexpr int kAbSmemV = 4 * kAbTileBytes;
constexpr int kAbSmemS = 5 * kAbTileBytes;            // dS, same shape
constexpr int kAbSmemBars = 7 * kAbTileBytes;
constexpr int kAbSmemBytes = kAbSmemBars + 64;
constexpr int kAbTmemCols = 256;
constexpr uint32_t kAbColDV = 0, kAbColDK = 64, kAbColDQ = 128;      // reuse of the S / dP columns

struct AttnBwdParams {
  __nv_bfloat16* dqkv;
  int L, W;
  float scale;       // 1 / sqrt(dh)
};

__device__ __forceinline__ void ab_tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* smem, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 128B-swizzled operand descriptor with explicit leading / stride byte offsets (MN-major tiles wider than one
// 64-element swizzle atom need the leading offset = distance between the atoms)
__device__ __forceinline__ uint64_t ab_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// kind::f16, bf16 operands, fp32 accumulate, M = 128; bits 15 / 16 select MN-major A / B
__host__ __device__ constexpr uint32_t ab_idesc(uint32_t n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((n >> 3) << 17) |
         ((128u >> 4) << 24);
}

__device__ __forceinline__ float ab_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAbThreads, 2)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem + kAbSmemQ;
  uint8_t* sK = smem + kAbSmemK;
  uint8_t* sV = smem + kAbSmemV;
  uint8_t* sO = smem + kAbSmemO;
  uint8_t* sP = smem + kAbSmemP;
  uint8_t* sS = smem + kAbSmemS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAbSmemBars);
  uint64_t* ld_full = bars;        // Q, K, V, dO landed
  uint64_t* mma1 = bars + 1;       // S and dP complete
  uint64_t* rows_done = bars + 2;  // P and dS written (4 warps)
  uint64_t* mma2 = bars + 3;       // dV, dK, dQ complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.x;
  const int n = blockIdx.y;
  const int L = p.L;
  const int nk = (L + 15) / 16;      // 16-row K steps that contain live tokens

  if (warp == 4) {
    if (lane == 0) {
      if ((smem_u32(smem) & 1023u) != 0) __trap();
      tma_prefetch_desc(&tmQKV);
      tma_prefetch_desc(&tmDO);
      mbar_init(ld_full, 1);
      mbar_init(mma1, 1);
      mbar_init(rows_done, 4);
      mbar_init(mma2, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kAbTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync();       // prologue done while the previous kernel drained; no global access before this point

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(ld_full, 4 * kAbTileBytes);
      ab_tma_load_3d(&tmQKV, ld_full, sQ, h * kAbDh, 0, n);
      ab_tma_load_3d(&tmQKV, ld_full, sK, p.W + h * kAbDh, 0, n);
      ab_tma_load_3d(&tmQKV, ld_full, sV, 2 * p.W + h * kAbDh, 0, n);
      ab_tma_load_3d(&tmDO, ld_full, sO, h * kAbDh, 0, n);
      mbar_wait(ld_full, 0);
      tc_fence_after();
      constexpr uint32_t id_kk128 = ab_idesc(128, false, false);
      constexpr uint32_t id_mm64 = ab_idesc(64, true, true);
      constexpr uint32_t id_km64 = ab_idesc(64, false, true);
#pragma unroll
      for (int k = 0; k < kAbDh / 16; ++k) {      // S = Q K^T, dP = dO V^T (K = head dim)
        umma_f16(tmem_base, make_kmajor_desc(smem_u32(sQ), 128) + 2u * k, make_kmajor_desc(smem_u32(sK), 128) + 2u * k,
                 id_kk128, k != 0 ? 1u : 0u);
      }
#pragma unroll
      for (int k = 0; k < kAbDh / 16; ++k) {
        umma_f16(tmem_base + 128u, make_kmajor_desc(smem_u32(sO), 128) + 2u * k,
                 make_kmajor_desc(smem_u32(sV), 128) + 2u * k, id_kk128, k != 0 ? 1u : 0u);
      }
      umma_commit(mma1);
      mbar_wait(rows_done, 0);
      tc_fence_after();
      for (int ks = 0; ks < nk; ++ks) {           // dV = P^T dO, dK = dS^T Q (K = query index i, 16 rows = 2048 B)
        const uint32_t off = static_cast<uint32_t>(ks) * 2048u;
        umma_f16(tmem_base + kAbColDV, ab_desc(smem_u32(sP) + off, kAbTileBytes, 1024), ab_desc(smem_u32(sO) + off, 16, 1024),
                 id_mm64, ks != 0 ? 1u : 0u);
        umma_f16(tmem_base + kAbColDK, ab_desc(smem_u32(sS) + off, kAbTileBytes, 1024), ab_desc(smem_u32(sQ) + off, 16, 1024),
                 id_mm64, ks != 0 ? 1u : 0u);
      }
      for (int ks = 0; ks < nk; ++ks) {           // dQ = dS K (K = key index j: 64-key blocks, 32 B per step inside)
        const uint32_t a_off = static_cast<uint32_t>(ks >> 2) * kAbTileBytes + static_cast<uint32_t>(ks & 3) * 32u;
        umma_f16(tmem_base + kAbColDQ, ab_desc(smem_u32(sS) + a_off, 16, 1024),
                 ab_desc(smem_u32(sK) + static_cast<uint32_t>(ks) * 2048u, 16, 1024), id_km64, ks != 0 ? 1u : 0u);
      }
      umma_commit(mma2);
    }
  } else {
    // ------------------------------- row math: thread == query row i ------------
    const int i = warp * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t tS = tmem_base + lane_base, tdP = tmem_base + lane_base + 128u;
    const int lim = i < L ? i : -1;                 // visible keys: j <= lim (causal); padding rows see none
    const float c = p.scale * 1.4426950408889634f;
    const int nchunk = (L + 31) / 32;               // 32-column chunks that contain live keys
    mbar_wait(mma1, 0);
    tc_fence_after();
    float mx = -INFINITY;
    for (int c4 = 0; c4 < nchunk; ++c4) {
      uint32_t r[32];
      tmem_ld_32x32(tS + static_cast<uint32_t>(c4 * 32), r);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) mx = fmaxf(mx, (c4 * 32 + q <= lim) ? __uint_as_float(r[q]) : -INFINITY);
    }
    const float mc = mx * c;
    float l = 0.f, dsum = 0.f;
    for (int c4 = 0; c4 < nchunk; ++c4) {
      uint32_t r[32], g[32];
      tmem_ld_32x32(tS + static_cast<uint32_t>(c4 * 32), r);
      tmem_ld_32x32(tdP + static_cast<uint32_t>(c4 * 32), g);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const float e = (c4 * 32 + q <= lim) ? ab_exp2(fmaf(__uint_as_float(r[q]), c, -mc)) : 0.f;
        l += e;
        dsum = fmaf(e, __uint_as_float(g[q]), dsum);
      }
    }
    const float inv = lim >= 0 ? 1.0f / l : 0.f;
    const float D = dsum * inv;
    for (int c4 = 0; c4 < 4; ++c4) {
      uint32_t pk[16], dk[16];
      if (c4 < nchunk) {
        uint32_t r[32], g[32];
        tmem_ld_32x32(tS + static_cast<uint32_t>(c4 * 32), r);
        tmem_ld_32x32(tdP + static_cast<uint32_t>(c4 * 32), g);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; q += 2) {
          float p0 = (c4 * 32 + q <= lim) ? ab_exp2(fmaf(__uint_as_float(r[q]), c, -mc)) * inv : 0.f;
          float p1 = (c4 * 32 + q + 1 <= lim) ? ab_exp2(fmaf(__uint_as_float(r[q + 1]), c, -mc)) * inv : 0.f;
          const float d0 = p0 * (__uint_as_float(g[q]) - D) * p.scale;
          const float d1 = p1 * (__uint_as_float(g[q + 1]) - D) * p.scale;
          pk[q / 2] = pack_bf16(p0, p1);
          dk[q / 2] = pack_bf16(d0, d1);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) pk[q] = dk[q] = 0u;
      }
      const uint32_t blk = static_cast<uint32_t>(c4 >> 1) * kAbTileBytes;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t off = blk + swizzled_chunk_offset(i, (c4 & 1) * 4 + q, 128);
        *reinterpret_cast<uint4*>(sP + off) = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
        *reinterpret_cast<uint4*>(sS + off) = make_uint4(dk[q * 4], dk[q * 4 + 1], dk[q * 4 + 2], dk[q * 4 + 3]);
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(rows_done);
    // ------------------------------- epilogue: thread == token row ---------------
    mbar_wait(mma2, 0);
    tc_fence_after();
    __nv_bfloat16* orow = p.dqkv + (static_cast<int64_t>(n) * L + i) * 3 * p.W + h * kAbDh;
#pragma unroll 1
    for (int which = 0; which < 3; ++which) {       // 0: dQ, 1: dK, 2: dV
      const uint32_t col = which == 0 ? kAbColDQ : (which == 1 ? kAbColDK : kAbColDV);
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(tmem_base + lane_base + col, r0);
      tmem_ld_32x32(tmem_base + lane_base + col + 32u, r1);
      tmem_ld_wait();
      if (i < L) {
        uint4* op = reinterpret_cast<uint4*>(orow + which * p.W);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          op[q] = make_uint4(pack_bf16(__uint_as_float(r0[q * 8 + 0]), __uint_as_float(r0[q * 8 + 1])),
                             pack_bf16(__uint_as_float(r0[q * 8 + 2]), __uint_as_float(r0[q * 8 + 3])),
                             pack_bf16(__uint_as_float(r0[q * 8 + 4]), __uint_as_float(r0[q * 8 + 5])),
                             pack_bf16(__uint_as_float(r0[q * 8 + 6]), __uint_as_float(r0[q * 8 + 7])));
          op[4 + q] = make_uint4(pack_bf16(__uint_as_float(r1[q * 8 + 0]), __uint_as_float(r1[q * 8 + 1])),
                                 pack_bf16(__uint_as_float(r1[q * 8 + 2]), __uint_as_float(r1[q * 8 + 3])),
                                 pack_bf16(__uint_as_float(r1[q * 8 + 4]), __uint_as_float(r1[q * 8 + 5])),
                                 pack_bf16(__uint_as_float(r1[q * 8 + 6]), __uint_as_float(r1[q * 8 + 7])));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAbTmemCols);
  }
}

}  // namespace lecb

using namespace lecb;

extern "C" int lecb_attn_causal_bwd(const void* qkv, const void* dout, void* dqkv, int N, int L, int W, int heads,
                                    void* stream) {
  LECB_CHECK_ARG(qkv && dout && dqkv, "lecb_attn_causal_bwd: null pointer");
  LECB_CHECK_ARG(N > 0 && L > 0 && L <= kAbTile, "lecb_attn_causal_bwd: need 0 < L <= 128 (L=%d)", L);
  LECB_CHECK_ARG(heads > 0 && W == heads * kAbDh, "lecb_attn_causal_bwd: W=%d must equal heads*64 (heads=%d)", W, heads);
  LECB_CHECK_ARG(N <= 65535, "lecb_attn_causal_bwd: grid too large");
  LECB_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(dout) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0,
                 "lecb_attn_causal_bwd: operands must be 16-byte aligned");
  static DeviceOnce once;                    // the attribute is per device: one flag per device ordinal
  bool& configured = once.flag();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAbSmemBytes);
    if (e != cudaSuccess) return fail(LECB_ERR_CUDA, "cudaFuncSetAttribute(attn bwd smem=%d): %s", kAbSmemBytes, cudaGetErrorString(e));
    e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return fail(LECB_ERR_CUDA, "cudaFuncSetAttribute(attn bwd carveout): %s", cudaGetErrorString(e));
    configured = true;
  }
  CUtensorMap tmQKV, tmDO;
  int st = encode_tiled_3d(&tmQKV, qkv, static_cast<uint64_t>(3) * W, static_cast<uint64_t>(L), static_cast<uint64_t>(N),
                           kAbDh, kAbTile);
  if (st) return st;
  st = encode_tiled_3d(&tmDO, dout, static_cast<uint64_t>(W), static_cast<uint64_t>(L), static_cast<uint64_t>(N), kAbDh, kAbTile);
  if (st) return st;
  AttnBwdParams p;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.L = L;
  p.W = W;
  p.scale = 0.125f;
  launch_k(attn_bwd_kernel, dim3(dim3(heads, N)), dim3(kAbThreads), kAbSmemBytes, static_cast<cudaStream_t>(stream), tmQKV, tmDO, p);
  count_launch();
  return check_launch("attn_bwd_kernel");
}
