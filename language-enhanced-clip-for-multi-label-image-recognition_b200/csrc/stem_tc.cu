// Stem conv1 (3x3, stride 2, pad 1, 3 -> 32 channels, folded BN + ReLU; M:144-145,174-175) on tcgen05.
//
// The CUDA-core version is bound by fp32 FMA issue (27 x 32 FMAs per output pixel); as an implicit GEMM the
// arithmetic is trivial (M = pixels, N = 32, K = 27 padded to 32) and the kernel becomes a streaming transform:
// each thread gathers the 27 input values of one output pixel from the NCHW fp32 image, packs them to bf16 and
// writes one 64-byte im2col row of a 128 x 32 K-major (64B-swizzled) A tile in shared memory; two tcgen05.mma
// (M=128, N=32, K=16) against the 32 x 32 weight tile produce the 128 x 32 outputs in TMEM; the same thread reads
// its pixel's 32 channels back (tcgen05.ld), adds the bias, applies ReLU and stores 64 contiguous bytes of the
// NHWC bf16 output.  Persistent CTAs (weights, TMEM and barrier set up once), several CTAs per SM for overlap.
#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

constexpr int kStemCo = 32;
constexpr int kStemK = 32;          // 27 taps + 5 zero columns

__global__ void __launch_bounds__(128)
stem_conv1_tc_kernel(const float* __restrict__ x, const float* __restrict__ w27, const float* __restrict__ bias,
                     __nv_bfloat16* __restrict__ out, int B, int H, int W, int num_tiles) {
  __shared__ __align__(1024) uint8_t sA[128 * kStemK * 2];        // 8 KB, rows of 64 bytes
  __shared__ __align__(1024) uint8_t sW[kStemCo * kStemK * 2];    // 2 KB
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float sbias[kStemCo];
  const int t = threadIdx.x;
  const int warp = t >> 5;
  const int HO = H / 2, WO = W / 2;
  const int64_t total = static_cast<int64_t>(B) * HO * WO;

  // weights: W[co][k] = w27[k][co] (k = ci*9 + ky*3 + kx), K-major rows of 64 bytes, 64B swizzle
  for (int i = t; i < kStemCo * kStemK; i += 128) {
    const int co = i / kStemK, k = i % kStemK;
    const float v = k < 27 ? w27[k * kStemCo + co] : 0.f;
    const uint32_t off = swizzled_chunk_offset(co, k / 8, 64) + (k % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sW + off) = __float2bfloat16(v);
  }
  if (t < kStemCo) sbias[t] = bias[t];
  if (t == 0) {
    mbar_init(&mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 32);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t idesc = make_idesc_f16(kStemCo, true);
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t pix = static_cast<int64_t>(tile) * 128 + t;
    const bool live = pix < total;
    // ---- gather the 3x3x3 patch of this output pixel -> one bf16 im2col row ----
    float v[kStemK];
#pragma unroll
    for (int k = 27; k < kStemK; ++k) v[k] = 0.f;
    if (live) {
      const int wo = static_cast<int>(pix % WO);
      const int ho = static_cast<int>((pix / WO) % HO);
      const int b = static_cast<int>(pix / (static_cast<int64_t>(WO) * HO));
      const int wi0 = 2 * wo - 1, hi0 = 2 * ho - 1;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float* xp = x + (static_cast<int64_t>(b) * 3 + ci) * H * W;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int hi = hi0 + ky;
          const bool hok = hi >= 0 && hi < H;
          const float* rp = xp + static_cast<int64_t>(hi) * W + wi0;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int wi = wi0 + kx;
            v[ci * 9 + ky * 3 + kx] = (hok && wi >= 0 && wi < W) ? __ldg(rp + kx) : 0.f;
          }
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 27; ++k) v[k] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < kStemK / 8; ++c) {
      uint4 u;
      u.x = pack_bf16(v[c * 8 + 0], v[c * 8 + 1]);
      u.y = pack_bf16(v[c * 8 + 2], v[c * 8 + 3]);
      u.z = pack_bf16(v[c * 8 + 4], v[c * 8 + 5]);
      u.w = pack_bf16(v[c * 8 + 6], v[c * 8 + 7]);
      *reinterpret_cast<uint4*>(sA + swizzled_chunk_offset(t, c, 64)) = u;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (t == 0) {
      tc_fence_after();
      const uint64_t adesc = make_kmajor_desc(smem_u32(sA), 64);
      const uint64_t bdesc = make_kmajor_desc(smem_u32(sW), 64);
      umma_f16(tmem_base, adesc, bdesc, idesc, 0u);
      umma_f16(tmem_base, adesc + 2u, bdesc + 2u, idesc, 1u);
      umma_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: thread == output pixel (TMEM lane) ----
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16), r);
    tmem_ld_wait();
    tc_fence_before();
    if (live) {
      uint4* op = reinterpret_cast<uint4*>(out + pix * kStemCo);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 u;
        u.x = pack_bf16(fmaxf(__uint_as_float(r[c * 8 + 0]) + sbias[c * 8 + 0], 0.f), fmaxf(__uint_as_float(r[c * 8 + 1]) + sbias[c * 8 + 1], 0.f));
        u.y = pack_bf16(fmaxf(__uint_as_float(r[c * 8 + 2]) + sbias[c * 8 + 2], 0.f), fmaxf(__uint_as_float(r[c * 8 + 3]) + sbias[c * 8 + 3], 0.f));
        u.z = pack_bf16(fmaxf(__uint_as_float(r[c * 8 + 4]) + sbias[c * 8 + 4], 0.f), fmaxf(__uint_as_float(r[c * 8 + 5]) + sbias[c * 8 + 5], 0.f));
        u.w = pack_bf16(fmaxf(__uint_as_float(r[c * 8 + 6]) + sbias[c * 8 + 6], 0.f), fmaxf(__uint_as_float(r[c * 8 + 7]) + sbias[c * 8 + 7], 0.f));
        op[c] = u;
      }
    }
    // the next iteration's __syncthreads (after the A tile is rebuilt) orders these TMEM reads before the next MMA;
    // the A tile itself may be overwritten right away: the MMA that read it has completed (mma_bar)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

int launch_stem_conv1_tc(const float* x, const float* w, const float* bias, void* out, int B, int H, int W,
                         cudaStream_t s) {
  const int64_t total = static_cast<int64_t>(B) * (H / 2) * (W / 2);
  const int64_t tiles = (total + 127) / 128;
  if (tiles > 0x7fffffff) return fail(LECB_ERR_ARG, "lecb_stem_conv1: problem too large");
  const int sms = sm_count();
  if (sms <= 0) return fail(LECB_ERR_CUDA, "no CUDA device");
  const int64_t cap = static_cast<int64_t>(sms) * 12;
  const int grid = static_cast<int>(tiles < cap ? tiles : cap);
  stem_conv1_tc_kernel<<<grid, 128, 0, s>>>(x, w, bias, static_cast<__nv_bfloat16*>(out), B, H, W, static_cast<int>(tiles));
  count_launch();
  return check_launch("stem_conv1_tc_kernel");
}

}  // namespace lecb
