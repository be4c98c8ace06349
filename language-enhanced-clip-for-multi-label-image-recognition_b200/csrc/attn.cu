// Attention kernels (head dim 64):
//  * attnpool_query0: CLIP AttentionPool2d with if_pos=False (M:89-127 via T:413) reduced to what reaches
//    token 0: one query (the mean token) per image and head against P patch keys + the mean-token key.
//    By linearity of k_proj / v_proj the mean token's key/value are the means of the patch keys/values,
//    so the kernel only needs q [B,C], K = k_proj(x) [B*P,C] and V = v_proj(x) [B*P,C] (V is shared
//    with the per-patch local-feature path, T:409).
//  * causal_attn_fwd / bwd: the text transformer's masked self-attention (M:221-223, mask M:364-370),
//    L <= 128 tokens, whole (sequence, head) problem resident in one CTA's shared memory.
#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

constexpr int kDh = 64;

// grid (heads/8, B), 256 threads: warp = head.  Dynamic smem: 8 * (P+1) floats (scores / probs).
__global__ void __launch_bounds__(256)
attnpool_query0_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ kmat,
                       const __nv_bfloat16* __restrict__ vmat, __nv_bfloat16* __restrict__ out, int P, int C,
                       float qscale) {
  pdl_grid_sync();
  extern __shared__ float s_scores[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int h = blockIdx.x * 8 + warp;
  float* sc = s_scores + warp * (P + 1);
  const float* qp = q + static_cast<int64_t>(b) * C + h * kDh;
  float qr[kDh];
#pragma unroll
  for (int d = 0; d < kDh; ++d) qr[d] = __ldg(qp + d) * qscale;
  const __nv_bfloat16* kb = kmat + static_cast<int64_t>(b) * P * C + h * kDh;
  // pass 1: lane-parallel over patches, each lane does a full 64-dim dot (8 x 16-byte loads)
  float ssum = 0.f, smax = -INFINITY;
  for (int p = lane; p < P; p += 32) {
    const uint4* kp = reinterpret_cast<const uint4*>(kb + static_cast<int64_t>(p) * C);
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < kDh / 8; ++v) {
      const uint4 u = __ldg(kp + v);
      float2 f;
      f = unpack_bf16(u.x); s = fmaf(qr[8 * v + 0], f.x, s); s = fmaf(qr[8 * v + 1], f.y, s);
      f = unpack_bf16(u.y); s = fmaf(qr[8 * v + 2], f.x, s); s = fmaf(qr[8 * v + 3], f.y, s);
      f = unpack_bf16(u.z); s = fmaf(qr[8 * v + 4], f.x, s); s = fmaf(qr[8 * v + 5], f.y, s);
      f = unpack_bf16(u.w); s = fmaf(qr[8 * v + 6], f.x, s); s = fmaf(qr[8 * v + 7], f.y, s);
    }
    sc[p] = s;
    ssum += s;
    smax = fmaxf(smax, s);
  }
  ssum = warp_sum(ssum);
  smax = warp_max(smax);
  const float s_mean = ssum / static_cast<float>(P);      // score of the mean token (key = mean of keys)
  smax = fmaxf(smax, s_mean);
  float den = 0.f;
  for (int p = lane; p < P; p += 32) {
    const float e = __expf(sc[p] - smax);
    sc[p] = e;
    den += e;
  }
  den = warp_sum(den);
  const float e_mean = __expf(s_mean - smax);
  den += e_mean;
  __syncwarp();
  // pass 2: lanes over the 64 dims (2 each), loop over patches
  const __nv_bfloat16* vb = vmat + static_cast<int64_t>(b) * P * C + h * kDh + 2 * lane;
  float o0 = 0.f, o1 = 0.f, m0 = 0.f, m1 = 0.f;
  // eight row loads in flight per lane: the loop is a chain of dependent FMAs behind one 128-byte load per patch, and with
  // 16 warps per SM the load latency, not HBM, set the pace (1.8 TB/s before)
  int p = 0;
  for (; p + 8 <= P; p += 8) {
    uint32_t raw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) raw[j] = __ldg(reinterpret_cast<const uint32_t*>(vb + static_cast<int64_t>(p + j) * C));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 f = unpack_bf16(raw[j]);
      const float a = sc[p + j];
      o0 = fmaf(a, f.x, o0);
      o1 = fmaf(a, f.y, o1);
      m0 += f.x;
      m1 += f.y;
    }
  }
  for (; p < P; ++p) {
    const float2 f = unpack_bf16(__ldg(reinterpret_cast<const uint32_t*>(vb + static_cast<int64_t>(p) * C)));
    const float a = sc[p];
    o0 = fmaf(a, f.x, o0);
    o1 = fmaf(a, f.y, o1);
    m0 += f.x;
    m1 += f.y;
  }
  const float invP = 1.0f / static_cast<float>(P), invd = 1.0f / den;
  o0 = (o0 + e_mean * m0 * invP) * invd;
  o1 = (o1 + e_mean * m1 * invP) * invd;
  reinterpret_cast<uint32_t*>(out + static_cast<int64_t>(b) * C + h * kDh)[lane] = pack_bf16(o0, o1);
}

// ------------------------------------------------------------------------------------------------
// Causal self-attention forward.  qkv bf16 [N*L, 3W] (q | k | v per token), out bf16 [N*L, W].
// grid (heads, N); 128 threads; K/V of the (sequence, head) in smem (rows padded to 66 bf16).
// ------------------------------------------------------------------------------------------------
constexpr int kLMax = 128;
constexpr int kKvPitch = kDh + 2;

__global__ void __launch_bounds__(128)
causal_attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int L, int W,
                       float scale) {
  pdl_grid_sync();
  __shared__ __nv_bfloat16 sK[kLMax * kKvPitch];
  __shared__ __nv_bfloat16 sV[kLMax * kKvPitch];
  __shared__ float sP[4][kLMax];
  __shared__ float sQ[4][kDh];
  const int h = blockIdx.x, n = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* base = qkv + static_cast<int64_t>(n) * L * 3 * W + h * kDh;
  for (int i = threadIdx.x; i < L * (kDh / 2); i += blockDim.x) {
    const int r = i / (kDh / 2), c2 = i % (kDh / 2);
    const uint32_t kk = __ldg(reinterpret_cast<const uint32_t*>(base + static_cast<int64_t>(r) * 3 * W + W) + c2);
    const uint32_t vv = __ldg(reinterpret_cast<const uint32_t*>(base + static_cast<int64_t>(r) * 3 * W + 2 * W) + c2);
    reinterpret_cast<uint32_t*>(sK + r * kKvPitch)[c2] = kk;
    reinterpret_cast<uint32_t*>(sV + r * kKvPitch)[c2] = vv;
  }
  __syncthreads();
  for (int i = warp; i < L; i += 4) {
    const float2 qf = unpack_bf16(__ldg(reinterpret_cast<const uint32_t*>(base + static_cast<int64_t>(i) * 3 * W) + lane));
    sQ[warp][2 * lane] = qf.x * scale;
    sQ[warp][2 * lane + 1] = qf.y * scale;
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j <= i; j += 32) {
      const uint32_t* kr = reinterpret_cast<const uint32_t*>(sK + j * kKvPitch);
      float s = 0.f;
#pragma unroll
      for (int d2 = 0; d2 < kDh / 2; ++d2) {
        const float2 kf = unpack_bf16(kr[d2]);
        s = fmaf(sQ[warp][2 * d2], kf.x, s);
        s = fmaf(sQ[warp][2 * d2 + 1], kf.y, s);
      }
      sP[warp][j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float den = 0.f;
    for (int j = lane; j <= i; j += 32) {
      const float e = __expf(sP[warp][j] - mx);
      sP[warp][j] = e;
      den += e;
    }
    den = warp_sum(den);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j <= i; ++j) {
      const float2 vf = unpack_bf16(reinterpret_cast<const uint32_t*>(sV + j * kKvPitch)[lane]);
      const float a = sP[warp][j];
      o0 = fmaf(a, vf.x, o0);
      o1 = fmaf(a, vf.y, o1);
    }
    const float inv = 1.0f / den;
    reinterpret_cast<uint32_t*>(out + (static_cast<int64_t>(n) * L + i) * W + h * kDh)[lane] = pack_bf16(o0 * inv, o1 * inv);
    __syncwarp();
  }
}

}  // namespace lecb

using namespace lecb;

extern "C" int lecb_attnpool_query0(const float* q, const void* kmat, const void* vmat, void* out, int B, int P, int C,
                                    int heads, void* stream) {
  LECB_CHECK_ARG(q && kmat && vmat && out, "lecb_attnpool_query0: null pointer");
  LECB_CHECK_ARG(B > 0 && P > 0 && heads > 0 && C == heads * kDh, "lecb_attnpool_query0: head dim must be 64 (C=%d heads=%d)", C, heads);
  LECB_CHECK_ARG(heads % 8 == 0, "lecb_attnpool_query0: heads=%d must be a multiple of 8", heads);
  const size_t smem = static_cast<size_t>(8) * (P + 1) * sizeof(float);
  LECB_CHECK_ARG(smem <= 48 * 1024, "lecb_attnpool_query0: P=%d too large", P);
  dim3 grid(heads / 8, B);
  launch_k(attnpool_query0_kernel, dim3(grid), dim3(256), smem, static_cast<cudaStream_t>(stream), 
      q, static_cast<const __nv_bfloat16*>(kmat), static_cast<const __nv_bfloat16*>(vmat),
      static_cast<__nv_bfloat16*>(out), P, C, 1.0f / sqrtf(static_cast<float>(kDh)));
  count_launch();
  return check_launch("attnpool_query0_kernel");
}

extern "C" int lecb_causal_attn_fwd(const void* qkv, void* out, int N, int L, int W, int heads, void* stream) {
  LECB_CHECK_ARG(qkv && out, "lecb_causal_attn_fwd: null pointer");
  LECB_CHECK_ARG(N > 0 && L > 0 && L <= kLMax, "lecb_causal_attn_fwd: need 0 < L <= %d (L=%d)", kLMax, L);
  LECB_CHECK_ARG(W == heads * kDh, "lecb_causal_attn_fwd: head dim must be 64 (W=%d heads=%d)", W, heads);
  dim3 grid(heads, N);
  launch_k(causal_attn_fwd_kernel, dim3(grid), dim3(128), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), L, W, 1.0f / sqrtf(static_cast<float>(kDh)));
  count_launch();
  return check_launch("causal_attn_fwd_kernel");
}
