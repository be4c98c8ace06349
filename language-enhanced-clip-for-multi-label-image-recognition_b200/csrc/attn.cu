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
// Eight lanes share one patch row (lane & 7 = 16-byte chunk of the head's 128-byte row, lane >> 3 = patch slot), so a warp-wide
// load fetches four complete rows, every lane keeps only its 8 query / 8 output values (round 1 / 2: 64 query registers per lane
// held the kernel at 16 warps per SM, and with one full row per lane it ran at the load latency: 0.37 of the HBM peak), and
// sixteen rows per warp are in flight in both passes.
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 t;
  t = unpack_bf16(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16(u.w); f[6] = t.x; f[7] = t.y;
}

__global__ void __launch_bounds__(256)
attnpool_query0_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ kmat,
                       const __nv_bfloat16* __restrict__ vmat, __nv_bfloat16* __restrict__ out, int P, int C,
                       float qscale) {
  pdl_grid_sync();
  extern __shared__ float s_scores[];
  constexpr int kU = 4;                          // 4 x 4 patch rows per warp per step
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = lane >> 3, chunk = lane & 7;
  const int b = blockIdx.y;
  const int h = blockIdx.x * 8 + warp;
  float* sc = s_scores + warp * (P + 1);
  float q8[8];
  {
    const float4* qp = reinterpret_cast<const float4*>(q + static_cast<int64_t>(b) * C + h * kDh + chunk * 8);
    const float4 a = __ldg(qp), c = __ldg(qp + 1);
    q8[0] = a.x * qscale; q8[1] = a.y * qscale; q8[2] = a.z * qscale; q8[3] = a.w * qscale;
    q8[4] = c.x * qscale; q8[5] = c.y * qscale; q8[6] = c.z * qscale; q8[7] = c.w * qscale;
  }
  // pass 1: scores.  Row p of this warp's head: 128 bytes at kb + p*C; lane reads its 16-byte chunk of rows slot, slot+4, ...
  const __nv_bfloat16* kb = kmat + static_cast<int64_t>(b) * P * C + h * kDh + chunk * 8;
  float ssum = 0.f, smax = -INFINITY;
  for (int p0 = 0; p0 < P; p0 += 4 * kU) {
    uint4 raw[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int p = p0 + 4 * u + slot;
      raw[u] = p < P ? __ldg(reinterpret_cast<const uint4*>(kb + static_cast<int64_t>(p) * C)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int p = p0 + 4 * u + slot;
      float f[8];
      unpack8(raw[u], f);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(q8[i], f[i], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (p < P && chunk == 0) {                  // one lane per row keeps the statistics
        sc[p] = s;
        ssum += s;
        smax = fmaxf(smax, s);
      }
    }
  }
  ssum = warp_sum(ssum);
  smax = warp_max(smax);
  const float s_mean = ssum / static_cast<float>(P);      // score of the mean token (key = mean of keys)
  smax = fmaxf(smax, s_mean);
  __syncwarp();
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
  // pass 2: probability-weighted sum and plain sum (the mean token's value) of the value rows, same row -> lane mapping
  const __nv_bfloat16* vb = vmat + static_cast<int64_t>(b) * P * C + h * kDh + chunk * 8;
  float o[8], m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    o[i] = 0.f;
    m[i] = 0.f;
  }
  for (int p0 = 0; p0 < P; p0 += 4 * kU) {
    uint4 raw[kU];
    float a[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int p = p0 + 4 * u + slot;
      const bool in = p < P;
      raw[u] = in ? __ldg(reinterpret_cast<const uint4*>(vb + static_cast<int64_t>(p) * C)) : make_uint4(0u, 0u, 0u, 0u);
      a[u] = in ? sc[p] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      float f[8];
      unpack8(raw[u], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        o[i] = fmaf(a[u], f[i], o[i]);
        m[i] += f[i];
      }
    }
  }
  const float invP = 1.0f / static_cast<float>(P), invd = 1.0f / den;
#pragma unroll
  for (int i = 0; i < 8; ++i) {                   // sum over the four patch slots
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
    m[i] += __shfl_xor_sync(0xffffffffu, m[i], 8);
    m[i] += __shfl_xor_sync(0xffffffffu, m[i], 16);
    o[i] = (o[i] + e_mean * m[i] * invP) * invd;
  }
  if (slot == 0) {
    uint4 u;
    u.x = pack_bf16(o[0], o[1]);
    u.y = pack_bf16(o[2], o[3]);
    u.z = pack_bf16(o[4], o[5]);
    u.w = pack_bf16(o[6], o[7]);
    *reinterpret_cast<uint4*>(out + static_cast<int64_t>(b) * C + h * kDh + chunk * 8) = u;
  }
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
