// HBM-bound row / pixel kernels: stem conv1 (3x3 stride 2, 3->C), 2x2 average pool, token mean,
// row L2 normalisation, LayerNorm.  128-bit vectorised, coalesced, warp-shuffle reductions.
#include <stdlib.h>

#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

int launch_stem_conv1_tc(const void* x, int x_is_u8, const float* w, const float* bias, const float* mean, const float* stdv,
                         void* out, int B, int H, int W, cudaStream_t s);      // stem_tc.cu

// ------------------------------------------------------------------------------------------------
// Stem conv1: NCHW fp32 [B,3,H,W] -> NHWC bf16 [B,H/2,W/2,CO], 3x3 stride 2 pad 1, folded BN + ReLU.
// One thread = one output pixel x CO channels (weights broadcast from shared memory).
// ------------------------------------------------------------------------------------------------
template <int CO>
__global__ void __launch_bounds__(128) stem_conv1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias,
                                                         __nv_bfloat16* __restrict__ out, int B, int H, int W) {
  pdl_grid_sync();
  // One thread = TWO horizontally adjacent output pixels x CO channels: the 5x3x3 input patch is shared and
  // every 128-bit weight broadcast from shared memory feeds 8 FMAs (the kernel is FMA-issue bound).
  __shared__ __align__(16) float sw[27 * CO];   // [tap(ci,ky,kx)][co]
  __shared__ float sb[CO];
  for (int i = threadIdx.x; i < 27 * CO; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < CO; i += blockDim.x) sb[i] = bias[i];
  __syncthreads();
  const int HO = H / 2, WO = W / 2, WP = (WO + 1) / 2;
  const int64_t total = static_cast<int64_t>(B) * HO * WP;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wp = static_cast<int>(idx % WP);
  const int ho = static_cast<int>((idx / WP) % HO);
  const int b = static_cast<int>(idx / (static_cast<int64_t>(WP) * HO));
  const int wo = 2 * wp;
  const bool second = wo + 1 < WO;
  float in[3][3][5];
#pragma unroll
  for (int ci = 0; ci < 3; ++ci) {
    const float* xp = x + (static_cast<int64_t>(b) * 3 + ci) * H * W;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hi = 2 * ho - 1 + ky;
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        const int wi = 2 * wo - 1 + c;
        in[ci][ky][c] = (hi >= 0 && hi < H && wi >= 0 && wi < W) ? __ldg(xp + static_cast<int64_t>(hi) * W + wi) : 0.f;
      }
    }
  }
  __nv_bfloat16* op = out + ((static_cast<int64_t>(b) * HO + ho) * WO + wo) * CO;
#pragma unroll
  for (int c0 = 0; c0 < CO; c0 += 8) {
    float a0[8], a1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a0[j] = a1[j] = sb[c0 + j];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int t = ci * 9 + ky * 3 + kx;
          const float4 w0 = *reinterpret_cast<const float4*>(&sw[t * CO + c0]);
          const float4 w1 = *reinterpret_cast<const float4*>(&sw[t * CO + c0 + 4]);
          const float p0 = in[ci][ky][kx], p1 = in[ci][ky][kx + 2];
          a0[0] = fmaf(p0, w0.x, a0[0]); a0[1] = fmaf(p0, w0.y, a0[1]); a0[2] = fmaf(p0, w0.z, a0[2]); a0[3] = fmaf(p0, w0.w, a0[3]);
          a0[4] = fmaf(p0, w1.x, a0[4]); a0[5] = fmaf(p0, w1.y, a0[5]); a0[6] = fmaf(p0, w1.z, a0[6]); a0[7] = fmaf(p0, w1.w, a0[7]);
          a1[0] = fmaf(p1, w0.x, a1[0]); a1[1] = fmaf(p1, w0.y, a1[1]); a1[2] = fmaf(p1, w0.z, a1[2]); a1[3] = fmaf(p1, w0.w, a1[3]);
          a1[4] = fmaf(p1, w1.x, a1[4]); a1[5] = fmaf(p1, w1.y, a1[5]); a1[6] = fmaf(p1, w1.z, a1[6]); a1[7] = fmaf(p1, w1.w, a1[7]);
        }
    uint4 u;
    u.x = pack_bf16(fmaxf(a0[0], 0.f), fmaxf(a0[1], 0.f));
    u.y = pack_bf16(fmaxf(a0[2], 0.f), fmaxf(a0[3], 0.f));
    u.z = pack_bf16(fmaxf(a0[4], 0.f), fmaxf(a0[5], 0.f));
    u.w = pack_bf16(fmaxf(a0[6], 0.f), fmaxf(a0[7], 0.f));
    *reinterpret_cast<uint4*>(op + c0) = u;
    if (second) {
      u.x = pack_bf16(fmaxf(a1[0], 0.f), fmaxf(a1[1], 0.f));
      u.y = pack_bf16(fmaxf(a1[2], 0.f), fmaxf(a1[3], 0.f));
      u.z = pack_bf16(fmaxf(a1[4], 0.f), fmaxf(a1[5], 0.f));
      u.w = pack_bf16(fmaxf(a1[6], 0.f), fmaxf(a1[7], 0.f));
      *reinterpret_cast<uint4*>(op + CO + c0) = u;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 2x2 average pool, NHWC bf16.  One thread = 8 channels of one output pixel (4 x 16-byte loads).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) avgpool2_kernel(const __nv_bfloat16* __restrict__ x,
                                                       __nv_bfloat16* __restrict__ out, int B, int H, int W, int C) {
  pdl_grid_sync();
  const int HO = H / 2, WO = W / 2, CV = C / 8;
  const int64_t total = static_cast<int64_t>(B) * HO * WO * CV;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % CV);
    int64_t t = idx / CV;
    const int wo = static_cast<int>(t % WO);
    t /= WO;
    const int ho = static_cast<int>(t % HO);
    const int b = static_cast<int>(t / HO);
    const __nv_bfloat16* p = x + ((static_cast<int64_t>(b) * H + 2 * ho) * W + 2 * wo) * C + cv * 8;
    const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(p + C));
    const uint4 a2 = __ldg(reinterpret_cast<const uint4*>(p + static_cast<int64_t>(W) * C));
    const uint4 a3 = __ldg(reinterpret_cast<const uint4*>(p + static_cast<int64_t>(W) * C + C));
    const uint32_t* u0 = reinterpret_cast<const uint32_t*>(&a0);
    const uint32_t* u1 = reinterpret_cast<const uint32_t*>(&a1);
    const uint32_t* u2 = reinterpret_cast<const uint32_t*>(&a2);
    const uint32_t* u3 = reinterpret_cast<const uint32_t*>(&a3);
    uint4 r;
    uint32_t* ur = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f0 = unpack_bf16(u0[j]), f1 = unpack_bf16(u1[j]), f2 = unpack_bf16(u2[j]), f3 = unpack_bf16(u3[j]);
      ur[j] = pack_bf16(0.25f * (f0.x + f1.x + f2.x + f3.x), 0.25f * (f0.y + f1.y + f2.y + f3.y));
    }
    *reinterpret_cast<uint4*>(out + idx * 8) = r;
  }
}

// ------------------------------------------------------------------------------------------------
// Token mean over P: x bf16 [B,P,C] -> out bf16 [B,C] (and optional fp32 copy).  grid (C/256... , B)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) token_mean_kernel(const __nv_bfloat16* __restrict__ x,
                                                         __nv_bfloat16* __restrict__ out_bf16,
                                                         float* __restrict__ out_f32, int P, int C) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  const int c2 = blockIdx.x * blockDim.x + threadIdx.x;   // pair of channels
  if (c2 * 2 >= C) return;
  const uint32_t* p = reinterpret_cast<const uint32_t*>(x + static_cast<int64_t>(b) * P * C) + c2;
  float s0 = 0.f, s1 = 0.f;
  for (int t = 0; t < P; ++t) {
    const float2 f = unpack_bf16(__ldg(p + static_cast<int64_t>(t) * (C / 2)));
    s0 += f.x;
    s1 += f.y;
  }
  const float inv = 1.0f / static_cast<float>(P);
  s0 *= inv;
  s1 *= inv;
  if (out_bf16) reinterpret_cast<uint32_t*>(out_bf16 + static_cast<int64_t>(b) * C)[c2] = pack_bf16(s0, s1);
  if (out_f32) {
    out_f32[static_cast<int64_t>(b) * C + 2 * c2] = s0;
    out_f32[static_cast<int64_t>(b) * C + 2 * c2 + 1] = s1;
  }
}

// ------------------------------------------------------------------------------------------------
// Row L2 normalisation  y = x / ||x||_2  (no epsilon: T:441-442, T:485-488).  One warp per row,
// 128-bit loads, row kept in registers between the reduction and the scaled store.
// ------------------------------------------------------------------------------------------------
template <typename TIn, typename TOut, int kVecPerLane>
__global__ void __launch_bounds__(256) l2norm_kernel(const TIn* __restrict__ x, TOut* __restrict__ y, int64_t rows,
                                                     int D) {
  pdl_grid_sync();
  constexpr int kElemsPerVec = 16 / sizeof(TIn);
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int nvec = D / kElemsPerVec;
  for (int64_t row = warp_global; row < rows; row += nwarps) {
    const uint4* xp = reinterpret_cast<const uint4*>(x + row * D);
    float v[kVecPerLane][kElemsPerVec];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        const uint4 u = __ldg(xp + vi);
        if constexpr (sizeof(TIn) == 4) {
          v[i][0] = __uint_as_float(u.x); v[i][1] = __uint_as_float(u.y);
          v[i][2] = __uint_as_float(u.z); v[i][3] = __uint_as_float(u.w);
        } else {
          float2 f;
          f = unpack_bf16(u.x); v[i][0] = f.x; v[i][1] = f.y;
          f = unpack_bf16(u.y); v[i][2] = f.x; v[i][3] = f.y;
          f = unpack_bf16(u.z); v[i][4] = f.x; v[i][5] = f.y;
          f = unpack_bf16(u.w); v[i][6] = f.x; v[i][7] = f.y;
        }
#pragma unroll
        for (int j = 0; j < kElemsPerVec; ++j) ss = fmaf(v[i][j], v[i][j], ss);
      }
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / sqrtf(ss);
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        if constexpr (sizeof(TOut) == 4) {
          float4* yp = reinterpret_cast<float4*>(y + row * D) + vi * (kElemsPerVec / 4);
#pragma unroll
          for (int q = 0; q < kElemsPerVec / 4; ++q)
            yp[q] = make_float4(v[i][4 * q] * inv, v[i][4 * q + 1] * inv, v[i][4 * q + 2] * inv, v[i][4 * q + 3] * inv);
        } else {
          uint32_t* yp = reinterpret_cast<uint32_t*>(y + row * D) + vi * (kElemsPerVec / 2);
#pragma unroll
          for (int q = 0; q < kElemsPerVec / 2; ++q) yp[q] = pack_bf16(v[i][2 * q] * inv, v[i][2 * q + 1] * inv);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm forward over the last dim (fp32 statistics, eps inside sqrt: M:193-199).  x fp32 [rows,D]
// -> y bf16 (GEMM operand) and optional fp32 mean / rstd for the backward.  One warp per row.
// ------------------------------------------------------------------------------------------------
template <int kVecPerLane>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                            const float* __restrict__ bta,
                                                            __nv_bfloat16* __restrict__ y_bf16,
                                                            float* __restrict__ y_f32, float* __restrict__ mean_out,
                                                            float* __restrict__ rstd_out, int64_t rows, int D,
                                                            float eps) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int nvec = D / 4;
  for (int64_t row = warp_global; row < rows; row += nwarps) {
    const float4* xp = reinterpret_cast<const float4*>(x + row * D);
    float4 v[kVecPerLane];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        v[i] = __ldg(xp + vi);
        s += v[i].x + v[i].y + v[i].z + v[i].w;
      }
    }
    const float mean = warp_sum(s) / static_cast<float>(D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += a * a + b * b + c * c + d * d;
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(D) + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + vi);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bta) + vi);
        const float o0 = (v[i].x - mean) * rstd * gg.x + bb.x, o1 = (v[i].y - mean) * rstd * gg.y + bb.y;
        const float o2 = (v[i].z - mean) * rstd * gg.z + bb.z, o3 = (v[i].w - mean) * rstd * gg.w + bb.w;
        if (y_bf16) {
          uint2 u;
          u.x = pack_bf16(o0, o1);
          u.y = pack_bf16(o2, o3);
          reinterpret_cast<uint2*>(y_bf16 + row * D)[vi] = u;
        }
        if (y_f32) reinterpret_cast<float4*>(y_f32 + row * D)[vi] = make_float4(o0, o1, o2, o3);
      }
    }
  }
}

static int grid_for(int64_t work_items, int per_block, int max_waves = 8) {
  const int64_t blocks = (work_items + per_block - 1) / per_block;
  const int64_t cap = static_cast<int64_t>(sm_count()) * max_waves;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace lecb

using namespace lecb;

extern "C" int lecb_stem_conv1(const float* x, const float* w, const float* bias, void* out, int B, int H, int W,
                               int Cout, void* stream) {
  LECB_CHECK_ARG(x && w && bias && out, "lecb_stem_conv1: null pointer");
  LECB_CHECK_ARG(B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "lecb_stem_conv1: H, W must be even and positive");
  const int64_t total = static_cast<int64_t>(B) * (H / 2) * ((W / 2 + 1) / 2);     // two output pixels per thread
  const unsigned grid = static_cast<unsigned>((total + 127) / 128);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static const bool cuda_core_stem = getenv("LECB_STEM_CUDA_CORES") != nullptr;     // read once, not per launch
  if (Cout == 32 && !cuda_core_stem)      // the real CLIP ResNets (width 64): tensor-core implicit GEMM
    return launch_stem_conv1_tc(x, 0, w, bias, nullptr, nullptr, out, B, H, W, s);
  if (Cout == 32)
    launch_k(stem_conv1_kernel<32>, dim3(grid), dim3(128), 0, s, x, w, bias, static_cast<__nv_bfloat16*>(out), B, H, W);
  else if (Cout == 48)
    launch_k(stem_conv1_kernel<48>, dim3(grid), dim3(128), 0, s, x, w, bias, static_cast<__nv_bfloat16*>(out), B, H, W);
  else if (Cout == 8)
    launch_k(stem_conv1_kernel<8>, dim3(grid), dim3(128), 0, s, x, w, bias, static_cast<__nv_bfloat16*>(out), B, H, W);
  else
    return fail(LECB_ERR_UNSUPPORTED, "lecb_stem_conv1: Cout=%d (supported: 8, 32, 48)", Cout);
  count_launch();
  return check_launch("stem_conv1_kernel");
}

extern "C" int lecb_stem_conv1_u8(const uint8_t* x, const float* w, const float* bias, const float* mean, const float* stdv,
                                  void* out, int B, int H, int W, int Cout, void* stream) {
  LECB_CHECK_ARG(x && w && bias && out && mean && stdv, "lecb_stem_conv1_u8: null pointer");
  LECB_CHECK_ARG(B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "lecb_stem_conv1_u8: H, W must be even and positive");
  LECB_CHECK_ARG(stdv[0] != 0.f && stdv[1] != 0.f && stdv[2] != 0.f, "lecb_stem_conv1_u8: zero std");
  LECB_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 3) == 0, "lecb_stem_conv1_u8: the image batch must be 4-byte aligned");
  if (Cout != 32) return fail(LECB_ERR_UNSUPPORTED, "lecb_stem_conv1_u8: Cout=%d (only the CLIP ResNet stem width 32)", Cout);
  return launch_stem_conv1_tc(x, 1, w, bias, mean, stdv, out, B, H, W, static_cast<cudaStream_t>(stream));
}

extern "C" int lecb_avgpool2x2(const void* x, void* out, int B, int H, int W, int C, void* stream) {
  LECB_CHECK_ARG(x && out, "lecb_avgpool2x2: null pointer");
  LECB_CHECK_ARG(B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0,
                 "lecb_avgpool2x2: need even H, W and C %% 8 == 0 (H=%d W=%d C=%d)", H, W, C);
  const int64_t total = static_cast<int64_t>(B) * (H / 2) * (W / 2) * (C / 8);
  launch_k(avgpool2_kernel, dim3(grid_for(total, 256, 16)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), B, H, W, C);
  count_launch();
  return check_launch("avgpool2_kernel");
}

extern "C" int lecb_token_mean(const void* x, void* out_bf16, float* out_f32, int B, int P, int C, void* stream) {
  LECB_CHECK_ARG(x && (out_bf16 || out_f32), "lecb_token_mean: null pointer");
  LECB_CHECK_ARG(B > 0 && P > 0 && C > 0 && C % 2 == 0, "lecb_token_mean: bad shape");
  dim3 grid((C / 2 + 127) / 128, B);
  launch_k(token_mean_kernel, dim3(grid), dim3(128), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out_bf16), out_f32, P, C);
  count_launch();
  return check_launch("token_mean_kernel");
}

template <typename TIn, typename TOut>
static int launch_l2norm(const void* x, void* y, int64_t rows, int D, cudaStream_t s) {
  constexpr int kElems = 16 / sizeof(TIn);
  const int nvec = D / kElems;
  const int per_lane = (nvec + 31) / 32;
  const int grid = grid_for(rows, 8, 16);
  const TIn* xi = static_cast<const TIn*>(x);
  TOut* yo = static_cast<TOut*>(y);
  if (per_lane <= 1) launch_k(l2norm_kernel<TIn, TOut, 1>, dim3(grid), dim3(256), 0, s, xi, yo, rows, D);
  else if (per_lane <= 2) launch_k(l2norm_kernel<TIn, TOut, 2>, dim3(grid), dim3(256), 0, s, xi, yo, rows, D);
  else if (per_lane <= 4) launch_k(l2norm_kernel<TIn, TOut, 4>, dim3(grid), dim3(256), 0, s, xi, yo, rows, D);
  else if (per_lane <= 8) launch_k(l2norm_kernel<TIn, TOut, 8>, dim3(grid), dim3(256), 0, s, xi, yo, rows, D);
  else return fail(LECB_ERR_UNSUPPORTED, "lecb_l2norm_rows: D=%d too wide", D);
  count_launch();
  return check_launch("l2norm_kernel");
}

extern "C" int lecb_l2norm_rows(const void* x, void* y, int64_t rows, int D, int in_is_bf16, int out_is_bf16,
                                void* stream) {
  LECB_CHECK_ARG(x && y, "lecb_l2norm_rows: null pointer");
  LECB_CHECK_ARG(rows > 0 && D > 0 && D % 8 == 0, "lecb_l2norm_rows: need D %% 8 == 0 (D=%d)", D);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (in_is_bf16 && out_is_bf16) return launch_l2norm<__nv_bfloat16, __nv_bfloat16>(x, y, rows, D, s);
  if (in_is_bf16) return launch_l2norm<__nv_bfloat16, float>(x, y, rows, D, s);
  if (out_is_bf16) return launch_l2norm<float, __nv_bfloat16>(x, y, rows, D, s);
  return launch_l2norm<float, float>(x, y, rows, D, s);
}

extern "C" int lecb_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32,
                                  float* mean, float* rstd, int64_t rows, int D, float eps, void* stream) {
  LECB_CHECK_ARG(x && gamma && beta && (y_bf16 || y_f32), "lecb_layernorm_fwd: null pointer");
  LECB_CHECK_ARG(rows > 0 && D > 0 && D % 4 == 0, "lecb_layernorm_fwd: need D %% 4 == 0 (D=%d)", D);
  const int per_lane = (D / 4 + 31) / 32;
  const int grid = grid_for(rows, 8, 16);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(y_bf16);
  if (per_lane <= 2) launch_k(layernorm_fwd_kernel<2>, dim3(grid), dim3(256), 0, s, x, gamma, beta, yb, y_f32, mean, rstd, rows, D, eps);
  else if (per_lane <= 4) launch_k(layernorm_fwd_kernel<4>, dim3(grid), dim3(256), 0, s, x, gamma, beta, yb, y_f32, mean, rstd, rows, D, eps);
  else if (per_lane <= 8) launch_k(layernorm_fwd_kernel<8>, dim3(grid), dim3(256), 0, s, x, gamma, beta, yb, y_f32, mean, rstd, rows, D, eps);
  else return fail(LECB_ERR_UNSUPPORTED, "lecb_layernorm_fwd: D=%d too wide", D);
  count_launch();
  return check_launch("layernorm_fwd_kernel");
}
