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
constexpr int kStemK = 32;          // 27 taps + 2 bias columns (A = 1.0, W = bias as a bf16 hi / lo pair) + 3 zero columns
constexpr int kTileH = 8, kTileW = 16;                    // output pixels per CTA tile (128 = TMEM lanes)
constexpr int kPatchH = 2 * kTileH + 1, kPatchW = 2 * kTileW + 1;      // 17 x 33 input pixels
constexpr int kPatchPitch = 100;                          // bf16 elements per patch row (33 * 3 = 99, even pitch: 4-byte rows)

// max(x, 0) and round to bf16, two values per instruction (lo -> bits 0-15)
__device__ __forceinline__ uint32_t pack_relu_bf16(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

struct StemNorm {
  float mean[3], std[3];
};

template <bool kU8>
__global__ void __launch_bounds__(128, 6)
stem_conv1_tc_kernel(const void* __restrict__ xin, const float* __restrict__ w27, const float* __restrict__ bias,
                     __nv_bfloat16* __restrict__ out, int B, int H, int W, int tiles_x, int tiles_y, int num_tiles, StemNorm nrm) {
  __shared__ __align__(1024) uint8_t sA[128 * kStemK * 2];        // 8 KB, rows of 64 bytes
  __shared__ __align__(1024) uint8_t sW[kStemCo * kStemK * 2];    // 2 KB
  __shared__ __align__(16) __nv_bfloat16 patch[kPatchH * kPatchPitch];
  __shared__ __nv_bfloat16 lut[kU8 ? 3 * 256 : 1];
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
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
    } else if (k == 27) {
      v = bias[co];                                                        // rounded to bf16 below: the "hi" half
    } else if (k == 28) {
      v = bias[co] - __bfloat162float(__float2bfloat16(bias[co]));        // the "lo" half: hi + lo = bias to 2^-17 relative
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
  // the prologue read only the layer's own constants (packed weights, bias: written at engine construction, many kernels
  // ago); the pixels may come from the previous kernel (crop + resize) and `out` may still be read by it
  pdl_wait();
  const uint32_t idesc = make_idesc_f16(kStemCo, true);
  uint32_t phase = 0;
  const int py = t / kTileW, px = t % kTileW;

  // ---- input staging, split in two so that the NEXT tile's global loads are in flight while this tile is multiplied and
  // stored: fetch() issues the loads into registers, stash() converts them into the [17][33][3] bf16 patch (zeros outside
  // the image: conv padding, ragged tiles).  ncu on the first version of this kernel: 62 % issue-slot utilisation with 40 % of
  // all instructions in per-segment 64-bit address arithmetic and 16 % in the tile -> (image, row, column) divisions — so
  // the loops below run over compile-time (plane, row group) indices with 32-bit offsets from one per-tile base pointer, the
  // tile origin is computed once per tile, and the 33rd column of the 51 segments is fetched by 51 threads in one step.
  const int lane = t & 31;
  constexpr int kRowsPerWarp = (kPatchH + 3) / 4;                      // 5 patch rows per warp (row = warp + 4 j)
  uint32_t words[kU8 ? kRowsPerWarp : 1];                              // u8: one aligned word of each of the warp's rows
  float v0[kU8 ? 1 : 3 * kRowsPerWarp], v1 = 0.f;                      // fp32: columns 0..31 of (plane, row); column 32 of segment t
  const int64_t total_u8 = static_cast<int64_t>(B) * H * W * 3;
  const int tiles_per_image = tiles_x * tiles_y;
  struct Origin {
    int b, hi0, wi0;
  };
  auto tile_origin = [&](int tile) {
    Origin o;
    o.b = tile / tiles_per_image;
    const int rem = tile - o.b * tiles_per_image;
    const int ty = rem / tiles_x;
    o.hi0 = 2 * (ty * kTileH) - 1;
    o.wi0 = 2 * ((rem - ty * tiles_x) * kTileW) - 1;
    return o;
  };
  auto fetch = [&](const Origin& o) {
    if (kU8) {
      // a patch row is 99 contiguous bytes of the NHWC image: lanes 0..25 fetch the aligned 32-bit words covering it
      const uint8_t* xb = static_cast<const uint8_t*>(xin);
      const int64_t img0 = static_cast<int64_t>(o.b) * H * W * 3;
#pragma unroll
      for (int it = 0; it < kRowsPerWarp; ++it) {
        const int r = warp + 4 * it;
        const int hi = o.hi0 + r;
        const bool row_ok = r < kPatchH && hi >= 0 && hi < H;
        const int64_t g0 = img0 + ((row_ok ? hi : 0) * W + o.wi0) * 3;                            // first byte of the row (may be < 0)
        const int64_t off = (g0 & ~static_cast<int64_t>(3)) + 4 * lane;                           // aligned down (floor, two's complement)
        uint32_t word = 0;
        if (row_ok && lane < 26) {
          if (off >= 0 && off + 4 <= total_u8) {
            word = __ldg(reinterpret_cast<const uint32_t*>(xb + off));
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (off + k >= 0 && off + k < total_u8) word |= static_cast<uint32_t>(__ldg(xb + off + k)) << (8 * k);
          }
        }
        words[kU8 ? it : 0] = word;
      }
    } else {
      const float* xi = static_cast<const float*>(xin) + static_cast<int64_t>(o.b) * 3 * H * W;       // this image
      const int plane = H * W;
      const int wa = o.wi0 + lane;
      const bool col_ok = wa >= 0 && wa < W;
#pragma unroll
      for (int j = 0; j < kRowsPerWarp; ++j) {
        const int r = warp + 4 * j;
        const int hi = o.hi0 + r;
        const bool ok = r < kPatchH && hi >= 0 && hi < H && col_ok;
        const int off = hi * W + wa;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) v0[kU8 ? 0 : ci * kRowsPerWarp + j] = ok ? __ldg(xi + ci * plane + off) : 0.f;
      }
      // column 32 of segment t = (plane t / 17, row t % 17): 51 threads, one load each
      {
        const int ci = t / kPatchH, r = t - ci * kPatchH;
        const int hi = o.hi0 + r, wb = o.wi0 + 32;
        v1 = (t < 3 * kPatchH && hi >= 0 && hi < H && wb < W) ? __ldg(xi + ci * plane + hi * W + wb) : 0.f;      // wb >= 31 > 0
      }
    }
  };
  auto stash = [&](const Origin& o) {
    if (kU8) {
      // lane l holds bytes [4l, 4l+4) of the aligned window; byte cc = 4l + k - shift of the row is (column cc / 3, channel cc % 3).
      // Only tiles on the left / right image border have invalid columns: the common tiles skip the per-byte column test.
      const int img_lo2 = static_cast<int>((static_cast<int64_t>(o.b) * H * W * 3) & 3);
      const bool edge_cols = o.wi0 < 0 || o.wi0 + kPatchW > W;
#pragma unroll
      for (int it = 0; it < kRowsPerWarp; ++it) {
        const int r = warp + 4 * it;
        const int hi = o.hi0 + r;
        const bool row_ok = hi >= 0 && hi < H;
        if (r < kPatchH && lane < 26) {
          // low two bits of the row's first byte address (floor-mod also for the negative offset of the first image row)
          const int shift = (img_lo2 + (((row_ok ? hi : 0) * W + o.wi0) * 3)) & 3;
          const int cc0 = 4 * lane - shift;                              // row byte of this lane's first byte (may be < 0)
          const uint32_t w4 = words[kU8 ? it : 0];
          __nv_bfloat16* prow = patch + r * kPatchPitch;
          if (row_ok && !edge_cols) {
            // interior row of an interior tile (block-uniform x warp-uniform condition): every column is a pixel of the image,
            // so only the two ends of the 99-byte window need a range test; the channel of byte k is (ci0 + k) mod 3 with ONE
            // division per lane and row (4 = 1 mod 3)
            const int q = ((cc0 + 3) * 171) >> 9;                          // (cc0 + 3) / 3, cc0 >= -3
            const int ci0 = cc0 + 3 - 3 * q;                               // cc0 mod 3 (floor)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              int ci = ci0 + k;
              ci -= ci >= 3 ? 3 : 0;                                       // ci0 + k <= 5
              const __nv_bfloat16 val = lut[ci * 256 + ((w4 >> (8 * k)) & 0xffu)];
              if (static_cast<unsigned>(cc0 + k) < static_cast<unsigned>(kPatchW * 3)) prow[cc0 + k] = val;
            }
            continue;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int cc = cc0 + k;
            if (cc >= 0 && cc < kPatchW * 3) {
              const int c = (cc * 171) >> 9;                             // cc / 3 for 0 <= cc < 512
              const int ci = cc - 3 * c;
              bool ok = row_ok;
              if (edge_cols) {
                const int wi = o.wi0 + c;
                ok = ok && wi >= 0 && wi < W;
              }
              prow[cc] = ok ? lut[ci * 256 + ((w4 >> (8 * k)) & 0xffu)] : __float2bfloat16(0.f);
            }
          }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < kRowsPerWarp; ++j) {
        const int r = warp + 4 * j;
        if (r < kPatchH) {
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) patch[r * kPatchPitch + lane * 3 + ci] = __float2bfloat16(v0[kU8 ? 0 : ci * kRowsPerWarp + j]);
        }
      }
      if (t < 3 * kPatchH) {
        const int ci = t / kPatchH, r = t - ci * kPatchH;
        patch[r * kPatchPitch + 32 * 3 + ci] = __float2bfloat16(v1);
      }
    }
  };

  Origin cur = tile_origin(blockIdx.x < static_cast<unsigned>(num_tiles) ? blockIdx.x : 0);
  if (static_cast<int>(blockIdx.x) < num_tiles) fetch(cur);
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    if (tile + static_cast<int>(gridDim.x) >= num_tiles) pdl_trigger();      // last tile of this CTA (late trigger: lecb_common.cuh)
    const int b = cur.b, ho0 = (cur.hi0 + 1) / 2, wo0 = (cur.wi0 + 1) / 2;
    // the previous iteration's readers of `patch` finished before its second __syncthreads
    stash(cur);
    __syncthreads();
    if (tile + static_cast<int>(gridDim.x) < num_tiles) {      // the next tile's loads are in flight during the MMA and the stores
      cur = tile_origin(tile + gridDim.x);
      fetch(cur);
    }
    // ---- this thread's im2col row: three runs of 9 consecutive patch elements (ky = 0, 1, 2) ----
    {
      uint16_t v[kStemK];
      // the folded-BN bias rides through the MMA: columns 27 / 28 of every A row are 1.0 against the bias's bf16 (hi, lo) pair in
      // W — the accumulator leaves TMEM with the bias already added, and the epilogue is one cvt.relu per channel pair
      v[27] = v[28] = 0x3F80;
#pragma unroll
      for (int k = 29; k < kStemK; ++k) v[k] = 0;
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
        u.x = pack_relu_bf16(__uint_as_float(r[c * 8 + 0]), __uint_as_float(r[c * 8 + 1]));
        u.y = pack_relu_bf16(__uint_as_float(r[c * 8 + 2]), __uint_as_float(r[c * 8 + 3]));
        u.z = pack_relu_bf16(__uint_as_float(r[c * 8 + 4]), __uint_as_float(r[c * 8 + 5]));
        u.w = pack_relu_bf16(__uint_as_float(r[c * 8 + 6]), __uint_as_float(r[c * 8 + 7]));
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
    launch_k(stem_conv1_tc_kernel<true>, dim3(grid), dim3(128), 0, s, x, w, bias, static_cast<__nv_bfloat16*>(out), B, H, W, tiles_x, tiles_y,
                                                    static_cast<int>(tiles), nrm);
  else
    launch_k(stem_conv1_tc_kernel<false>, dim3(grid), dim3(128), 0, s, x, w, bias, static_cast<__nv_bfloat16*>(out), B, H, W, tiles_x, tiles_y,
                                                     static_cast<int>(tiles), nrm);
  count_launch();
  return check_launch("stem_conv1_tc_kernel");
}

}  // namespace lecb
