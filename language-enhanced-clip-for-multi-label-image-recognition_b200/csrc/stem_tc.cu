// Stem conv1 (3x3, stride 2, pad 1, 3 -> 32 channels, folded BN + ReLU; M:144-145,174-175) on tcgen05.
//
// As an implicit GEMM the arithmetic is trivial (M = pixels, N = 32, K = 27 padded to 32): the kernel is a streaming
// transform bound by how it touches memory.  Round 1 gathered the 27 taps of every output pixel with scalar loads
// straight from the NCHW fp32 image (0.37 of HBM peak: every input value was requested 2.25 times, as 4-byte sectors).
// Now a CTA owns an 8 x 16 patch of output pixels, stages the 17 x 33 x 3 input patch it needs ONCE through shared
// memory (coalesced row segments, converted to bf16 on the way, conv padding = zeros), each thread then builds the
// 64-byte im2col row of its pixel from shared memory, two tcgen05.mma (M=128, N=32, K=16) against the 32 x 32 weight tile
// produce the outputs in TMEM and the same thread stores its pixel's 32 channels (64 contiguous bytes, NHWC bf16).
//
// Two input formats:
//   * fp32 NCHW, already normalised — the tensor the reference's DataLoader hands to DenseCLIP.forward (T:401-403);
//   * uint8 NHWC raw pixels — what the test-time window kernel (window_resize.cu) produces and what a host that skips
//     the float conversion uploads (4x fewer bytes over PCIe and HBM).  ToTensor + Normalize (transforms.py:396-402:
//     v / 255, then (x - mean) / std in fp32) is applied on the fly through a 3 x 256 entry table built with exactly
//     those fp32 operations, so the bf16 operand is bit-identical to converting the reference's float tensor.
// K order inside a row: k = ky*9 + kx*3 + ci (the patch is pixel-interleaved), weights permuted to match.
#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

constexpr int kStemCo = 32;
constexpr int kStemK = 32;          // 27 taps + 5 zero columns
constexpr int kTileH = 8, kTileW = 16;                    // output pixels per CTA tile (128 = TMEM lanes)
constexpr int kPatchH = 2 * kTileH + 1, kPatchW = 2 * kTileW + 1;      // 17 x 33 input pixels
constexpr int kPatchPitch = 100;                          // bf16 elements per patch row (33 * 3 = 99, even pitch: 4-byte rows)

struct StemNorm {
  float mean[3], std[3];
};

template <bool kU8>
__global__ void __launch_bounds__(128)
stem_conv1_tc_kernel(const void* __restrict__ xin, const float* __restrict__ w27, const float* __restrict__ bias,
                     __nv_bfloat16* __restrict__ out, int B, int H, int W, int tiles_x, int tiles_y, int num_tiles, StemNorm nrm) {
  __shared__ __align__(1024) uint8_t sA[128 * kStemK * 2];        // 8 KB, rows of 64 bytes
  __shared__ __align__(1024) uint8_t sW[kStemCo * kStemK * 2];    // 2 KB
  __shared__ __align__(16) __nv_bfloat16 patch[kPatchH * kPatchPitch];
  __shared__ __nv_bfloat16 lut[kU8 ? 3 * 256 : 1];
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float sbias[kStemCo];
  const int t = threadIdx.x;
  const int warp = t >> 5;
  const int HO = H / 2, WO = W / 2;

  // weights: sW[co][k], k = ky*9 + kx*3 + ci  <-  w27[(ci*9 + ky*3 + kx)][co]; K-major rows of 64 bytes, 64B swizzle
  for (int i = t; i < kStemCo * kStemK; i += 128) {
    const int co = i / kStemK, k = i % kStemK;
    float v = 0.f;
    if (k < 27) {
      const int ky = k / 9, kx = (k % 9) / 3, ci = k % 3;
      v = w27[(ci * 9 + ky * 3 + kx) * kStemCo + co];
    }
    const uint32_t off = swizzled_chunk_offset(co, k / 8, 64) + (k % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sW + off) = __float2bfloat16(v);
  }
  if (kU8) {
    // ToTensor: byte / 255 (fp32 division), Normalize: (x - mean) / std, both correctly rounded like ATen's CPU ops
    for (int i = t; i < 3 * 256; i += 128) {
      const int c = i >> 8;
      const float v = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(i & 255), 255.0f), nrm.mean[c]), nrm.std[c]);
      lut[i] = __float2bfloat16(v);
    }
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
  const int py = t / kTileW, px = t % kTileW;

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int tx = tile % tiles_x;
    const int ty = (tile / tiles_x) % tiles_y;
    const int b = tile / (tiles_x * tiles_y);
    const int ho0 = ty * kTileH, wo0 = tx * kTileW;
    const int hi0 = 2 * ho0 - 1, wi0 = 2 * wo0 - 1;
    // ---- stage the input patch: [17][33][3] bf16, zeros outside the image (conv padding, ragged tiles) ----
    // One warp per patch row segment, lanes along the contiguous axis (coalesced), no per-element divisions.
    const int lane = t & 31;
    if (kU8) {
      // a patch row is 99 contiguous bytes of the NHWC image: lanes 0..25 fetch the aligned 32-bit words covering it
      const uint8_t* xb = static_cast<const uint8_t*>(xin);
      const int64_t total = static_cast<int64_t>(B) * H * W * 3;
      constexpr int kRowsPerWarp = (kPatchH + 3) / 4;                    // 5: every word load is issued before the first use
      uint32_t words[kRowsPerWarp];
#pragma unroll
      for (int it = 0; it < kRowsPerWarp; ++it) {
        const int r = warp + 4 * it;
        const int hi = hi0 + r;
        const bool row_ok = r < kPatchH && hi >= 0 && hi < H;
        const int64_t g0 = ((static_cast<int64_t>(b) * H + (row_ok ? hi : 0)) * W + wi0) * 3;      // first byte of the row (may be < 0)
        const int64_t o = (g0 >= 0 ? g0 : g0 - 3) / 4 * 4 + 4 * lane;                             // aligned down (floor)
        uint32_t word = 0;
        if (row_ok && lane < 26) {
          if (o >= 0 && o + 4 <= total) {
            word = __ldg(reinterpret_cast<const uint32_t*>(xb + o));
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (o + k >= 0 && o + k < total) word |= static_cast<uint32_t>(__ldg(xb + o + k)) << (8 * k);
          }
        }
        words[it] = word;
      }
#pragma unroll
      for (int it = 0; it < kRowsPerWarp; ++it) {
        const int r = warp + 4 * it;
        const int hi = hi0 + r;
        const bool row_ok = hi >= 0 && hi < H;
        if (r < kPatchH && lane < 26) {
          const int64_t g0 = ((static_cast<int64_t>(b) * H + (row_ok ? hi : 0)) * W + wi0) * 3;
          const int shift = static_cast<int>(g0 - (g0 >= 0 ? g0 : g0 - 3) / 4 * 4);                // 0..3
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int cc = 4 * lane + k - shift;                         // byte position inside the row: col * 3 + ci
            if (cc >= 0 && cc < kPatchW * 3) {
              const int c = (cc * 171) >> 9;                             // cc / 3 for cc < 512
              const int ci = cc - 3 * c;
              const int wi = wi0 + c;
              const bool ok = row_ok && wi >= 0 && wi < W;
              patch[r * kPatchPitch + cc] = ok ? lut[ci * 256 + ((words[it] >> (8 * k)) & 0xffu)] : __float2bfloat16(0.f);
            }
          }
        }
      }
    } else {
      // 3 planes x 17 rows = 51 segments of 33 floats; all of a warp's loads are issued before the first store
      const float* x = static_cast<const float*>(xin) + static_cast<int64_t>(b) * 3 * H * W;
      constexpr int kSeg = 3 * kPatchH;                                  // 51
      constexpr int kPerWarp = (kSeg + 3) / 4;                           // 13
      float v0[kPerWarp], v1[kPerWarp];
#pragma unroll
      for (int it = 0; it < kPerWarp; ++it) {
        const int sgm = warp + 4 * it;
        const int ci = sgm / kPatchH, r = sgm - ci * kPatchH;
        const int hi = hi0 + r;
        const bool row_ok = sgm < kSeg && hi >= 0 && hi < H;
        const float* rp = x + (static_cast<int64_t>(ci) * H + (row_ok ? hi : 0)) * W;
        const int wa = wi0 + lane, wb = wi0 + 32;
        v0[it] = (row_ok && wa >= 0 && wa < W) ? __ldg(rp + wa) : 0.f;
        v1[it] = (row_ok && lane == 0 && wb < W) ? __ldg(rp + wb) : 0.f;       // wb >= 31 > 0 always
      }
#pragma unroll
      for (int it = 0; it < kPerWarp; ++it) {
        const int sgm = warp + 4 * it;
        if (sgm < kSeg) {
          const int ci = sgm / kPatchH, r = sgm - ci * kPatchH;
          patch[r * kPatchPitch + lane * 3 + ci] = __float2bfloat16(v0[it]);
          if (lane == 0) patch[r * kPatchPitch + 32 * 3 + ci] = __float2bfloat16(v1[it]);
        }
      }
    }
    __syncthreads();
    // ---- this thread's im2col row: three runs of 9 consecutive patch elements (ky = 0, 1, 2) ----
    {
      uint16_t v[kStemK];
#pragma unroll
      for (int k = 27; k < kStemK; ++k) v[k] = 0;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        // element offset (2py+ky)*100 + 6px is even: five aligned 32-bit words cover the nine values
        const uint32_t* src = reinterpret_cast<const uint32_t*>(patch + (2 * py + ky) * kPatchPitch + 6 * px);
#pragma unroll
        for (int wd = 0; wd < 5; ++wd) {
          const uint32_t u = src[wd];
          v[ky * 9 + 2 * wd] = static_cast<uint16_t>(u & 0xffffu);
          if (2 * wd + 1 < 9) v[ky * 9 + 2 * wd + 1] = static_cast<uint16_t>(u >> 16);
        }
      }
#pragma unroll
      for (int c = 0; c < kStemK / 8; ++c) {
        uint4 u;
        u.x = v[c * 8 + 0] | (static_cast<uint32_t>(v[c * 8 + 1]) << 16);
        u.y = v[c * 8 + 2] | (static_cast<uint32_t>(v[c * 8 + 3]) << 16);
        u.z = v[c * 8 + 4] | (static_cast<uint32_t>(v[c * 8 + 5]) << 16);
        u.w = v[c * 8 + 6] | (static_cast<uint32_t>(v[c * 8 + 7]) << 16);
        *reinterpret_cast<uint4*>(sA + swizzled_chunk_offset(t, c, 64)) = u;
      }
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
    const int ho = ho0 + py, wo = wo0 + px;
    if (ho < HO && wo < WO) {
      uint4* op = reinterpret_cast<uint4*>(out + ((static_cast<int64_t>(b) * HO + ho) * WO + wo) * kStemCo);
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
    // The next iteration's first __syncthreads (after the patch is rebuilt) orders these TMEM reads before the next MMA
    // and every thread's reads of `patch` (done before the MMA above) before its overwrite; the A tile may be
    // overwritten once the MMA that read it has completed (mma_bar).
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

int launch_stem_conv1_tc(const void* x, int x_is_u8, const float* w, const float* bias, const float* mean, const float* stdv,
                         void* out, int B, int H, int W, cudaStream_t s) {
  const int HO = H / 2, WO = W / 2;
  const int tiles_x = (WO + kTileW - 1) / kTileW, tiles_y = (HO + kTileH - 1) / kTileH;
  const int64_t tiles = static_cast<int64_t>(B) * tiles_x * tiles_y;
  if (tiles > 0x7fffffff) return fail(LECB_ERR_ARG, "lecb_stem_conv1: problem too large");
  const int sms = sm_count();
  if (sms <= 0) return fail(LECB_ERR_CUDA, "no CUDA device");
  const int64_t cap = static_cast<int64_t>(sms) * 12;
  const int grid = static_cast<int>(tiles < cap ? tiles : cap);
  StemNorm nrm{};
  for (int c = 0; c < 3; ++c) {
    nrm.mean[c] = mean ? mean[c] : 0.f;
    nrm.std[c] = stdv ? stdv[c] : 1.f;
  }
  if (x_is_u8)
    stem_conv1_tc_kernel<true><<<grid, 128, 0, s>>>(x, w, bias, static_cast<__nv_bfloat16*>(out), B, H, W, tiles_x, tiles_y,
                                                    static_cast<int>(tiles), nrm);
  else
    stem_conv1_tc_kernel<false><<<grid, 128, 0, s>>>(x, w, bias, static_cast<__nv_bfloat16*>(out), B, H, W, tiles_x, tiles_y,
                                                     static_cast<int>(tiles), nrm);
  count_launch();
  return check_launch("stem_conv1_tc_kernel");
}

}  // namespace lecb
