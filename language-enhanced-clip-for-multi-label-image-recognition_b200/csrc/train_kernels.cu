// Backward-pass kernels of the text-only prompt-tuning step (T:473-545 + loss.backward()).
// All CLIP weights are frozen (T:763-765), so the backward is data-gradient only: no weight gradients,
// the only trainable tensors are the prompt contexts reached through the prompt embeddings.
//   quick_gelu fwd / bwd (M:202-204)           elementwise, bf16, 128-bit vectorised
//   layernorm bwd (M:193-199)                   one warp per row, dgamma/dbeta not needed
//   causal attention bwd (M:221-223)            one CTA per (sequence, head), probabilities recomputed
//   l2norm bwd (T:487-488,503)                  one warp per row
//   head aggregate bwd (T:496-514)              one CTA per caption, forward statistics recomputed
//   tn_gemm_small                               out[J,D] = sum_r a[r,J] * b[r,D]  (prompt-feature grads)
#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

// ---------------------------------------------------------------------------------------------
// QuickGELU  u = v * sigmoid(1.702 v)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float qgelu_grad(float v) {
  const float s = 1.0f / (1.0f + __expf(-1.702f * v));
  return s * (1.0f + 1.702f * v * (1.0f - s));
}

__global__ void __launch_bounds__(256)
quick_gelu_fwd_kernel(const uint4* __restrict__ v, uint4* __restrict__ u, int64_t nvec) {
  pdl_grid_sync();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 a = __ldg(v + i);
    const uint32_t* ai = reinterpret_cast<const uint32_t*>(&a);
    uint4 r;
    uint32_t* ri = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16(ai[j]);
      ri[j] = pack_bf16(quick_gelu(f.x), quick_gelu(f.y));
    }
    u[i] = r;
  }
}

// dv = du * f'(v), all bf16
__global__ void __launch_bounds__(256)
quick_gelu_bwd_kernel(const uint4* __restrict__ du, const uint4* __restrict__ v, uint4* __restrict__ dv, int64_t nvec) {
  pdl_grid_sync();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 g = __ldg(du + i), a = __ldg(v + i);
    const uint32_t* gi = reinterpret_cast<const uint32_t*>(&g);
    const uint32_t* ai = reinterpret_cast<const uint32_t*>(&a);
    uint4 r;
    uint32_t* ri = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 gf = unpack_bf16(gi[j]), vf = unpack_bf16(ai[j]);
      ri[j] = pack_bf16(gf.x * qgelu_grad(vf.x), gf.y * qgelu_grad(vf.y));
    }
    dv[i] = r;
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward (data gradient only):  dx_out = dx_in + rstd * (g - mean(g) - xhat * mean(g*xhat)),
// g = dy * gamma.  dy fp32 [rows,D]; x fp32 (forward input); writes fp32 and/or bf16.
// ---------------------------------------------------------------------------------------------
template <int kVecPerLane>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dx_in,
                     float* __restrict__ dx_f32, __nv_bfloat16* __restrict__ dx_bf16, int64_t rows, int D) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int nvec = D / 4;
  for (int64_t row = warp_global; row < rows; row += nwarps) {
    const float mu = mean[row], rs = rstd[row];
    float4 g[kVecPerLane], xh[kVecPerLane];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        const float4 d = __ldg(reinterpret_cast<const float4*>(dy + row * D) + vi);
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + row * D) + vi);
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + vi);
        g[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        s1 += g[i].x + g[i].y + g[i].z + g[i].w;
        s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
      }
    }
    s1 = warp_sum(s1) / static_cast<float>(D);
    s2 = warp_sum(s2) / static_cast<float>(D);
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        float4 o = make_float4(rs * (g[i].x - s1 - xh[i].x * s2), rs * (g[i].y - s1 - xh[i].y * s2),
                               rs * (g[i].z - s1 - xh[i].z * s2), rs * (g[i].w - s1 - xh[i].w * s2));
        if (dx_in != nullptr) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(dx_in + row * D) + vi);
          o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        }
        if (dx_f32 != nullptr) reinterpret_cast<float4*>(dx_f32 + row * D)[vi] = o;
        if (dx_bf16 != nullptr) {
          uint2 u;
          u.x = pack_bf16(o.x, o.y);
          u.y = pack_bf16(o.z, o.w);
          reinterpret_cast<uint2*>(dx_bf16 + row * D)[vi] = u;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Causal attention backward.  qkv bf16 [N*L,3W]; dout bf16 [N*L,W]; dqkv bf16 [N*L,3W].  grid (heads, N).
// ---------------------------------------------------------------------------------------------
constexpr int kDh = 64;
constexpr int kLMaxB = 96;
constexpr int kPitch = kDh + 2;

__global__ void __launch_bounds__(256)
causal_attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                       __nv_bfloat16* __restrict__ dqkv, int L, int W, float scale) {
  pdl_grid_sync();
  extern __shared__ uint8_t smem_raw[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sK = sQ + kLMaxB * kPitch;
  __nv_bfloat16* sV = sK + kLMaxB * kPitch;
  __nv_bfloat16* sO = sV + kLMaxB * kPitch;                   // dO
  float* sP = reinterpret_cast<float*>(sO + kLMaxB * kPitch);  // [L][L+1] probabilities, then dS
  float* sD = sP + kLMaxB * (kLMaxB + 1);                       // [L] row dot(P, dP)
  const int h = blockIdx.x, n = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int PP = L + 1;
  const __nv_bfloat16* base = qkv + static_cast<int64_t>(n) * L * 3 * W + h * kDh;
  const __nv_bfloat16* dob = dout + static_cast<int64_t>(n) * L * W + h * kDh;
  for (int i = tid; i < L * (kDh / 2); i += blockDim.x) {
    const int r = i / (kDh / 2), c2 = i % (kDh / 2);
    const uint32_t* rowp = reinterpret_cast<const uint32_t*>(base + static_cast<int64_t>(r) * 3 * W);
    reinterpret_cast<uint32_t*>(sQ + r * kPitch)[c2] = __ldg(rowp + c2);
    reinterpret_cast<uint32_t*>(sK + r * kPitch)[c2] = __ldg(rowp + W / 2 + c2);
    reinterpret_cast<uint32_t*>(sV + r * kPitch)[c2] = __ldg(rowp + W + c2);
    reinterpret_cast<uint32_t*>(sO + r * kPitch)[c2] = __ldg(reinterpret_cast<const uint32_t*>(dob + static_cast<int64_t>(r) * W) + c2);
  }
  __syncthreads();
  // S = scale * Q K^T (causal) and dP = dO V^T, one (i,j) pair per thread iteration
  for (int idx = tid; idx < L * L; idx += blockDim.x) {
    const int i = idx / L, j = idx % L;
    if (j > i) continue;
    const uint32_t* qi = reinterpret_cast<const uint32_t*>(sQ + i * kPitch);
    const uint32_t* kj = reinterpret_cast<const uint32_t*>(sK + j * kPitch);
    float s = 0.f;
#pragma unroll 8
    for (int d2 = 0; d2 < kDh / 2; ++d2) {
      const float2 a = unpack_bf16(qi[d2]), b = unpack_bf16(kj[d2]);
      s = fmaf(a.x, b.x, s);
      s = fmaf(a.y, b.y, s);
    }
    sP[i * PP + j] = s * scale;
  }
  __syncthreads();
  for (int i = warp; i < L; i += blockDim.x / 32) {     // row softmax
    float mx = -INFINITY;
    for (int j = lane; j <= i; j += 32) mx = fmaxf(mx, sP[i * PP + j]);
    mx = warp_max(mx);
    float den = 0.f;
    for (int j = lane; j <= i; j += 32) {
      const float e = __expf(sP[i * PP + j] - mx);
      sP[i * PP + j] = e;
      den += e;
    }
    den = warp_sum(den);
    const float inv = 1.0f / den;
    for (int j = lane; j <= i; j += 32) sP[i * PP + j] *= inv;
  }
  __syncthreads();
  // dV[j][d] = sum_{i>=j} P[i][j] dO[i][d]   (before P is overwritten by dS)
  __nv_bfloat16* dq_out = dqkv + static_cast<int64_t>(n) * L * 3 * W + h * kDh;
  for (int idx = tid; idx < L * (kDh / 2); idx += blockDim.x) {
    const int j = idx / (kDh / 2), d2 = idx % (kDh / 2);
    float a0 = 0.f, a1 = 0.f;
    for (int i = j; i < L; ++i) {
      const float pij = sP[i * PP + j];
      const float2 o = unpack_bf16(reinterpret_cast<const uint32_t*>(sO + i * kPitch)[d2]);
      a0 = fmaf(pij, o.x, a0);
      a1 = fmaf(pij, o.y, a1);
    }
    reinterpret_cast<uint32_t*>(dq_out + static_cast<int64_t>(j) * 3 * W + 2 * W)[d2] = pack_bf16(a0, a1);
  }
  __syncthreads();
  // dS = P * (dP - rowsum(P * dP)), dP[i][j] = dO[i] . V[j]
  for (int i = warp; i < L; i += blockDim.x / 32) {
    float part = 0.f;
    for (int j = lane; j <= i; j += 32) {
      const uint32_t* oi = reinterpret_cast<const uint32_t*>(sO + i * kPitch);
      const uint32_t* vj = reinterpret_cast<const uint32_t*>(sV + j * kPitch);
      float dp = 0.f;
#pragma unroll 8
      for (int d2 = 0; d2 < kDh / 2; ++d2) {
        const float2 a = unpack_bf16(oi[d2]), b = unpack_bf16(vj[d2]);
        dp = fmaf(a.x, b.x, dp);
        dp = fmaf(a.y, b.y, dp);
      }
      const float pij = sP[i * PP + j];
      part += pij * dp;
      sP[i * PP + j] = pij * dp;            // temporarily P*dP; fixed up below
      // keep P for the fix-up: dS = P*dP - P*D  => store P in the (unused) upper triangle mirror
      sP[j * PP + i + 1] = pij;             // (j, i+1) with i+1 > j: strictly upper part incl. column L
    }
    part = warp_sum(part);
    if (lane == 0) sD[i] = part;
  }
  __syncthreads();
  for (int idx = tid; idx < L * L; idx += blockDim.x) {
    const int i = idx / L, j = idx % L;
    if (j > i) continue;
    const float pij = sP[j * PP + i + 1];
    sP[i * PP + j] = (sP[i * PP + j] - pij * sD[i]) * scale;      // dS * scale (chain through S = scale*QK^T)
  }
  __syncthreads();
  // dQ[i][d] = sum_{j<=i} dS[i][j] K[j][d];   dK[j][d] = sum_{i>=j} dS[i][j] Q[i][d]
  for (int idx = tid; idx < L * (kDh / 2); idx += blockDim.x) {
    const int r = idx / (kDh / 2), d2 = idx % (kDh / 2);
    float q0 = 0.f, q1 = 0.f, k0 = 0.f, k1 = 0.f;
    for (int j = 0; j <= r; ++j) {
      const float ds = sP[r * PP + j];
      const float2 kk = unpack_bf16(reinterpret_cast<const uint32_t*>(sK + j * kPitch)[d2]);
      q0 = fmaf(ds, kk.x, q0);
      q1 = fmaf(ds, kk.y, q1);
    }
    for (int i = r; i < L; ++i) {
      const float ds = sP[i * PP + r];
      const float2 qq = unpack_bf16(reinterpret_cast<const uint32_t*>(sQ + i * kPitch)[d2]);
      k0 = fmaf(ds, qq.x, k0);
      k1 = fmaf(ds, qq.y, k1);
    }
    reinterpret_cast<uint32_t*>(dq_out + static_cast<int64_t>(r) * 3 * W)[d2] = pack_bf16(q0, q1);
    reinterpret_cast<uint32_t*>(dq_out + static_cast<int64_t>(r) * 3 * W + W)[d2] = pack_bf16(k0, k1);
  }
}

// ---------------------------------------------------------------------------------------------
// y = x/||x||  backward:  dx = (dy - y (y.dy)) / ||x||.   fp32 rows, one warp per row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int64_t rows, int D) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float* xp = x + row * D;
  const float* gp = dy + row * D;
  float ss = 0.f, dot = 0.f;
  for (int d = lane; d < D; d += 32) {
    ss = fmaf(xp[d], xp[d], ss);
    dot = fmaf(xp[d], gp[d], dot);
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float inv = rsqrtf(ss);
  const float c = dot * inv * inv * inv;          // (y.dy)/||x|| * (1/||x||) with y = x*inv
  for (int d = lane; d < D; d += 32) dx[row * D + d] = gp[d] * inv - xp[d] * c;
}

// ---------------------------------------------------------------------------------------------
// Head aggregation backward (T:496-514): gradient of logits_local w.r.t. the raw dot products.
// One CTA (8 warps) per caption; pass 1 recomputes the per-class spatial-softmax statistics (max, sum,
// weighted sum), pass 2 emits d_dots[row, n_txt*K] (positive-prompt block = 0).  Masked rows get zeros.
// ---------------------------------------------------------------------------------------------
constexpr int kAggWarps = 8;

template <int kJ>
__global__ void __launch_bounds__(kAggWarps * 32)
head_aggregate_bwd_kernel(const float* __restrict__ dots, int ldn, const float* __restrict__ row_sumsq,
                          const uint8_t* __restrict__ row_mask, const float* __restrict__ grad_local,
                          float* __restrict__ d_dots, int B, int P, int K, int n_txt, float logit_scale,
                          float spatial_scale) {
  pdl_grid_sync();
  __shared__ float s_m[kAggWarps][kJ * 32];
  __shared__ float s_s[kAggWarps][kJ * 32];
  __shared__ float s_a[kAggWarps][kJ * 32];
  __shared__ float s_M[kJ * 32], s_S[kJ * 32], s_O[kJ * 32];   // final max, sum, sum(pi*neg)
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool evidence = n_txt >= 3;

  auto row_forward = [&](int64_t row, float (&neg)[kJ], float (&negw)[kJ], float (&w)[kJ], float (&t)[kJ], float& mx,
                         float& rn) {
    rn = row_sumsq != nullptr ? rsqrtf(__ldg(row_sumsq + row)) : 1.0f;
    const float* dp = dots + row * ldn;
    float evi[kJ];
    mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      const int k = lane + 32 * j;
      if (k < K) {
        neg[j] = __ldg(dp + K + k) * rn;
        evi[j] = evidence ? __ldg(dp + 2 * K + k) * rn : 0.f;
        mx = fmaxf(mx, neg[j]);
      } else {
        neg[j] = evi[j] = 0.f;
      }
    }
    if (evidence) {
      mx = warp_max(mx);
      const float gain = spatial_scale * (mx + 1.0f);
      float zmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        w[j] = (lane + 32 * j < K) ? gain * neg[j] : -INFINITY;
        zmax = fmaxf(zmax, w[j]);
      }
      zmax = warp_max(zmax);
      float den = 0.f;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        w[j] = (lane + 32 * j < K) ? __expf(w[j] - zmax) : 0.f;
        den += w[j];
      }
      den = warp_sum(den);
      const float inv = 1.0f / den;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        w[j] *= inv;
        negw[j] = neg[j] * w[j];
        t[j] = spatial_scale * evi[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        w[j] = 1.f;
        negw[j] = neg[j];
        t[j] = spatial_scale * neg[j];
      }
    }
  };

  // ---- pass 1: statistics ----
  float m[kJ], ssum[kJ], acc[kJ];
#pragma unroll
  for (int j = 0; j < kJ; ++j) {
    m[j] = -INFINITY;
    ssum[j] = 0.f;
    acc[j] = 0.f;
  }
  for (int p = warp; p < P; p += kAggWarps) {
    const int64_t row = static_cast<int64_t>(b) * P + p;
    if (row_mask != nullptr && row_mask[row]) continue;
    float neg[kJ], negw[kJ], w[kJ], t[kJ], mx, rn;
    row_forward(row, neg, negw, w, t, mx, rn);
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      if (t[j] > m[j]) {
        const float sc = __expf(m[j] - t[j]);
        ssum[j] = ssum[j] * sc + 1.0f;
        acc[j] = acc[j] * sc + negw[j];
        m[j] = t[j];
      } else {
        const float e = __expf(t[j] - m[j]);
        ssum[j] += e;
        acc[j] += e * negw[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kJ; ++j) {
    s_m[warp][lane + 32 * j] = m[j];
    s_s[warp][lane + 32 * j] = ssum[j];
    s_a[warp][lane + 32 * j] = acc[j];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < kJ * 32; k += blockDim.x) {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kAggWarps; ++w) M = fmaxf(M, s_m[w][k]);
    float S = 0.f, A = 0.f;
#pragma unroll
    for (int w = 0; w < kAggWarps; ++w) {
      const float sc = (s_m[w][k] == -INFINITY) ? 0.f : __expf(s_m[w][k] - M);
      S += s_s[w][k] * sc;
      A += s_a[w][k] * sc;
    }
    s_M[k] = M;
    s_S[k] = S;
    s_O[k] = (S > 0.f) ? A / S : 0.f;
  }
  __syncthreads();

  // ---- pass 2: gradients ----
  float G[kJ];
#pragma unroll
  for (int j = 0; j < kJ; ++j) {
    const int k = lane + 32 * j;
    G[j] = (k < K) ? logit_scale * grad_local[static_cast<int64_t>(b) * K + k] : 0.f;
  }
  for (int p = warp; p < P; p += kAggWarps) {
    const int64_t row = static_cast<int64_t>(b) * P + p;
    float* op = d_dots + row * ldn;
    if (row_mask != nullptr && row_mask[row]) {
      for (int c = lane; c < n_txt * K; c += 32) op[c] = 0.f;
      continue;
    }
    float neg[kJ], negw[kJ], w[kJ], t[kJ], mx, rn;
    row_forward(row, neg, negw, w, t, mx, rn);
    float dn[kJ], de[kJ];
    if (evidence) {
      float dw[kJ], wsum = 0.f;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        const int k = lane + 32 * j;
        const float pi = (k < K) ? __expf(t[j] - s_M[k]) / s_S[k] : 0.f;
        const float dnw = G[j] * pi;                                  // d/d(neg*w)
        de[j] = spatial_scale * G[j] * pi * (negw[j] - s_O[k]);
        dn[j] = dnw * w[j];
        dw[j] = dnw * neg[j];
        wsum += w[j] * dw[j];
      }
      wsum = warp_sum(wsum);
      const float gain = spatial_scale * (mx + 1.0f);
      float dmx = 0.f;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        const float dz = w[j] * (dw[j] - wsum);
        dn[j] += dz * gain;
        dmx += dz * spatial_scale * neg[j];
      }
      dmx = warp_sum(dmx);
      // the max's gradient goes to the (first) arg-max class
      int arg = 1 << 30;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        const int k = lane + 32 * j;
        if (k < K && neg[j] == mx) arg = min(arg, k);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) arg = min(arg, __shfl_xor_sync(0xffffffffu, arg, o));
#pragma unroll
      for (int j = 0; j < kJ; ++j)
        if (lane + 32 * j == arg) dn[j] += dmx;
    } else {
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        const int k = lane + 32 * j;
        const float pi = (k < K) ? __expf(t[j] - s_M[k]) / s_S[k] : 0.f;
        dn[j] = G[j] * pi * (1.0f + spatial_scale * (neg[j] - s_O[k]));
        de[j] = 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      const int k = lane + 32 * j;
      if (k < K) {
        op[k] = 0.f;
        op[K + k] = dn[j] * rn;
        if (evidence) op[2 * K + k] = de[j] * rn;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// out[J,D] (+)= sum_r a[r,J] * b[r,D];  a fp32 [R,lda], b bf16 or fp32 [R,D].  Each CTA owns a 16 x 64 output
// tile and walks R in 64-row chunks: the next chunk's global loads (16-byte vectors when the operands allow) are in
// flight in registers while the current one is consumed from shared memory; a thread owns 1 x 4 outputs (one
// broadcast LDS + one LDS.128 per four FMAs); fp32 accumulate in a fixed order, so the result is reproducible.
// (J = n_txt*K = 160..240, D <= 1024, R = B*L: ~0.6-2.4 GFLOP.  The first version — scalar loads, 32-row chunks,
// nothing in flight across the barrier — took 253 us for R = 1792, J = 160, D = 1024.)
// ---------------------------------------------------------------------------------------------
constexpr int kTnRows = 64;
template <typename TB>
__global__ void __launch_bounds__(256)
tn_gemm_small_kernel(const float* __restrict__ a, int lda, const TB* __restrict__ b, float* __restrict__ out, int R,
                     int J, int D, float alpha, int accumulate, int a_vec, int b_vec) {
  pdl_grid_sync();
  constexpr bool kBf16 = sizeof(TB) == 2;
  constexpr int kBPer = kBf16 ? 2 : 4;           // 16-byte vectors of b per thread and chunk (64 x 64 elements)
  constexpr int kBElems = kBf16 ? 8 : 4;         // elements per vector
  constexpr int kBVecRow = 64 / kBElems;         // vectors per 64-element row
  __shared__ __align__(16) float sA[kTnRows][16];
  __shared__ __align__(16) float sB[kTnRows][64];
  const int j0 = blockIdx.y * 16, d0 = blockIdx.x * 64;
  const int tid = threadIdx.x;
  const int tj = tid >> 4;           // 0..15
  const int td = (tid & 15) * 4;     // 0..60
  const int ar = tid >> 2, aj = (tid & 3) * 4;   // this thread's float4 of the a chunk
  float4 ra;
  float rb[kBPer][kBElems];
  auto fetch = [&](int r0) {
    const int r = r0 + ar;
    const float* src = a + static_cast<int64_t>(r) * lda + j0 + aj;
    if (r < R && a_vec && j0 + aj + 3 < J) {
      ra = __ldg(reinterpret_cast<const float4*>(src));
    } else {
      ra.x = (r < R && j0 + aj + 0 < J) ? __ldg(src + 0) : 0.f;
      ra.y = (r < R && j0 + aj + 1 < J) ? __ldg(src + 1) : 0.f;
      ra.z = (r < R && j0 + aj + 2 < J) ? __ldg(src + 2) : 0.f;
      ra.w = (r < R && j0 + aj + 3 < J) ? __ldg(src + 3) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < kBPer; ++q) {
      const int idx = tid + 256 * q;
      const int rr = r0 + idx / kBVecRow, dd = d0 + (idx % kBVecRow) * kBElems;
      const TB* bs = b + static_cast<int64_t>(rr) * D + dd;
      if (rr < R && b_vec && dd + kBElems - 1 < D) {
        if constexpr (kBf16) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(bs));
          const uint32_t* vi = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_bf16(vi[e]);
            rb[q][2 * e] = f.x;
            rb[q][2 * e + 1] = f.y;
          }
        } else {
          const float4 v = __ldg(reinterpret_cast<const float4*>(bs));
          rb[q][0] = v.x;
          rb[q][1] = v.y;
          rb[q][2] = v.z;
          rb[q][3] = v.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < kBElems; ++e) {
          float v = 0.f;
          if (rr < R && dd + e < D) {
            if constexpr (kBf16) v = __bfloat162float(bs[e]);
            else v = bs[e];
          }
          rb[q][e] = v;
        }
      }
    }
  };
  auto stash = [&]() {
    *reinterpret_cast<float4*>(&sA[ar][aj]) = ra;
#pragma unroll
    for (int q = 0; q < kBPer; ++q) {
      const int idx = tid + 256 * q;
      float* dst = &sB[idx / kBVecRow][(idx % kBVecRow) * kBElems];
#pragma unroll
      for (int e = 0; e < kBElems; e += 4)
        *reinterpret_cast<float4*>(dst + e) = make_float4(rb[q][e], rb[q][e + 1], rb[q][e + 2], rb[q][e + 3]);
    }
  };
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  fetch(0);
  for (int r0 = 0; r0 < R; r0 += kTnRows) {
    __syncthreads();                 // the previous chunk has been consumed
    stash();
    __syncthreads();
    if (r0 + kTnRows < R) fetch(r0 + kTnRows);
#pragma unroll 16
    for (int r = 0; r < kTnRows; ++r) {
      const float av = sA[r][tj];
      const float4 bv = *reinterpret_cast<const float4*>(&sB[r][td]);
      acc[0] = fmaf(av, bv.x, acc[0]);
      acc[1] = fmaf(av, bv.y, acc[1]);
      acc[2] = fmaf(av, bv.z, acc[2]);
      acc[3] = fmaf(av, bv.w, acc[3]);
    }
  }
  if (j0 + tj < J) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (d0 + td + q < D) {
        float* o = out + static_cast<int64_t>(j0 + tj) * D + d0 + td + q;
        *o = accumulate ? (*o + alpha * acc[q]) : alpha * acc[q];
      }
    }
  }
}

static int ew_grid(int64_t nvec) {
  int64_t blocks = (nvec + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace lecb

using namespace lecb;

// ---- residual ReLU adapter (`x + Adapter(x)`, trainers/Caption_distill_double_adapter.py:304-317, :109): elementwise pieces ----
// out = x + max(z, 0) (fp32) and its bf16 copy for the next GEMM; z = a1 @ W2^T raw.
namespace lecb {
__global__ void __launch_bounds__(256)
residual_relu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ z, float* __restrict__ out, int64_t n) {
  pdl_grid_sync();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = x[i] + fmaxf(z[i], 0.f);
}
// dz = dy * 1[z > 0], rounded to bf16 (the A operand of the data-gradient GEMM); z fp32 or bf16 (post-activation works too)
template <typename TZ>
__global__ void __launch_bounds__(256)
relu_bwd_kernel(const float* __restrict__ dy, const TZ* __restrict__ z, __nv_bfloat16* __restrict__ dz, int64_t n) {
  pdl_grid_sync();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float zi;
    if constexpr (sizeof(TZ) == 2) zi = __bfloat162float(z[i]);
    else zi = z[i];
    dz[i] = __float2bfloat16(zi > 0.f ? dy[i] : 0.f);
  }
}
}  // namespace lecb

extern "C" int lecb_residual_relu_fwd(const float* x, const float* z, float* out, int64_t n, void* stream) {
  LECB_CHECK_ARG(x && z && out && n > 0, "lecb_residual_relu_fwd: bad argument");
  launch_k(lecb::residual_relu_fwd_kernel, dim3(ew_grid((n + 7) / 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, z, out, n);
  count_launch();
  return check_launch("residual_relu_fwd_kernel");
}

extern "C" int lecb_relu_bwd(const float* dy, const void* z, int z_is_bf16, void* dz_bf16, int64_t n, void* stream) {
  LECB_CHECK_ARG(dy && z && dz_bf16 && n > 0, "lecb_relu_bwd: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (z_is_bf16)
    launch_k(lecb::relu_bwd_kernel<__nv_bfloat16>, dim3(ew_grid((n + 7) / 8)), dim3(256), 0, s, dy, static_cast<const __nv_bfloat16*>(z),
                                                                             static_cast<__nv_bfloat16*>(dz_bf16), n);
  else
    launch_k(lecb::relu_bwd_kernel<float>, dim3(ew_grid((n + 7) / 8)), dim3(256), 0, s, dy, static_cast<const float*>(z), static_cast<__nv_bfloat16*>(dz_bf16), n);
  count_launch();
  return check_launch("relu_bwd_kernel");
}

extern "C" int lecb_quick_gelu_fwd(const void* v, void* u, int64_t n, void* stream) {
  LECB_CHECK_ARG(v && u && n > 0 && n % 8 == 0, "lecb_quick_gelu_fwd: need n %% 8 == 0");
  launch_k(quick_gelu_fwd_kernel, dim3(ew_grid(n / 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const uint4*>(v), static_cast<uint4*>(u), n / 8);
  count_launch();
  return check_launch("quick_gelu_fwd_kernel");
}

extern "C" int lecb_quick_gelu_bwd(const void* du, const void* v, void* dv, int64_t n, void* stream) {
  LECB_CHECK_ARG(du && v && dv && n > 0 && n % 8 == 0, "lecb_quick_gelu_bwd: need n %% 8 == 0");
  launch_k(quick_gelu_bwd_kernel, dim3(ew_grid(n / 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const uint4*>(du), static_cast<const uint4*>(v), static_cast<uint4*>(dv), n / 8);
  count_launch();
  return check_launch("quick_gelu_bwd_kernel");
}

extern "C" int lecb_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                                  const float* rstd, const float* dx_in, float* dx_f32, void* dx_bf16, int64_t rows,
                                  int D, void* stream) {
  LECB_CHECK_ARG(dy && x && gamma && mean && rstd && (dx_f32 || dx_bf16), "lecb_layernorm_bwd: null pointer");
  LECB_CHECK_ARG(rows > 0 && D > 0 && D % 4 == 0, "lecb_layernorm_bwd: need D %% 4 == 0");
  const int per_lane = (D / 4 + 31) / 32;
  int64_t blocks = (rows + 7) / 8;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  const int grid = static_cast<int>(blocks > cap ? cap : blocks);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* db = static_cast<__nv_bfloat16*>(dx_bf16);
  if (per_lane <= 2) launch_k(layernorm_bwd_kernel<2>, dim3(grid), dim3(256), 0, s, dy, x, gamma, mean, rstd, dx_in, dx_f32, db, rows, D);
  else if (per_lane <= 4) launch_k(layernorm_bwd_kernel<4>, dim3(grid), dim3(256), 0, s, dy, x, gamma, mean, rstd, dx_in, dx_f32, db, rows, D);
  else if (per_lane <= 8) launch_k(layernorm_bwd_kernel<8>, dim3(grid), dim3(256), 0, s, dy, x, gamma, mean, rstd, dx_in, dx_f32, db, rows, D);
  else return fail(LECB_ERR_UNSUPPORTED, "lecb_layernorm_bwd: D=%d too wide", D);
  count_launch();
  return check_launch("layernorm_bwd_kernel");
}

extern "C" int lecb_causal_attn_bwd(const void* qkv, const void* dout, void* dqkv, int N, int L, int W, int heads,
                                    void* stream) {
  LECB_CHECK_ARG(qkv && dout && dqkv, "lecb_causal_attn_bwd: null pointer");
  LECB_CHECK_ARG(N > 0 && L > 0 && L <= kLMaxB, "lecb_causal_attn_bwd: need 0 < L <= %d (L=%d)", kLMaxB, L);
  LECB_CHECK_ARG(W == heads * kDh, "lecb_causal_attn_bwd: head dim must be 64");
  const size_t smem = 4 * kLMaxB * kPitch * sizeof(__nv_bfloat16) + (kLMaxB * (kLMaxB + 1) + kLMaxB) * sizeof(float);
  static DeviceOnce once;                    // the attribute is per device: one flag per device ordinal
  bool& configured = once.flag();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(causal_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(LECB_ERR_CUDA, "lecb_causal_attn_bwd: smem attr: %s", cudaGetErrorString(e));
    configured = true;
  }
  dim3 grid(heads, N);
  launch_k(causal_attn_bwd_kernel, dim3(grid), dim3(256), smem, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(dout), static_cast<__nv_bfloat16*>(dqkv),
      L, W, 1.0f / sqrtf(static_cast<float>(kDh)));
  count_launch();
  return check_launch("causal_attn_bwd_kernel");
}

extern "C" int lecb_l2norm_bwd(const float* x, const float* dy, float* dx, int64_t rows, int D, void* stream) {
  LECB_CHECK_ARG(x && dy && dx && rows > 0 && D > 0, "lecb_l2norm_bwd: bad argument");
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  launch_k(l2norm_bwd_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), x, dy, dx, rows, D);
  count_launch();
  return check_launch("l2norm_bwd_kernel");
}

extern "C" int lecb_head_aggregate_bwd(const float* dots, int ldn, const float* row_sumsq, const uint8_t* row_mask,
                                       const float* grad_local, float* d_dots, int B, int P, int K, int n_txt,
                                       float logit_scale, float spatial_scale, void* stream) {
  LECB_CHECK_ARG(dots && grad_local && d_dots, "lecb_head_aggregate_bwd: null pointer");
  LECB_CHECK_ARG(B > 0 && P > 0 && K > 0 && K <= 128, "lecb_head_aggregate_bwd: need 0 < K <= 128");
  LECB_CHECK_ARG(n_txt == 2 || n_txt == 3, "lecb_head_aggregate_bwd: n_txt must be 2 or 3");
  LECB_CHECK_ARG(ldn >= n_txt * K, "lecb_head_aggregate_bwd: ldn too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int kj = (K + 31) / 32;
#define LECB_AGGB(J)                                                                                             \
  launch_k(head_aggregate_bwd_kernel<J>, dim3(B), dim3(kAggWarps * 32), 0, s, dots, ldn, row_sumsq, row_mask, grad_local, d_dots, B, P, \
                                                            K, n_txt, logit_scale, spatial_scale)
  if (kj == 1) LECB_AGGB(1);
  else if (kj == 2) LECB_AGGB(2);
  else if (kj == 3) LECB_AGGB(3);
  else LECB_AGGB(4);
#undef LECB_AGGB
  count_launch();
  return check_launch("head_aggregate_bwd_kernel");
}

extern "C" int lecb_tn_gemm_small(const float* a, int lda, const void* b, int b_is_bf16, float* out, int R, int J,
                                  int D, float alpha, int accumulate, void* stream) {
  LECB_CHECK_ARG(a && b && out && R > 0 && J > 0 && D > 0 && lda >= J, "lecb_tn_gemm_small: bad argument");
  dim3 grid((D + 63) / 64, (J + 15) / 16);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int a_vec = (lda % 4 == 0) && (reinterpret_cast<uintptr_t>(a) & 15) == 0;
  const int b_vec = (D % (b_is_bf16 ? 8 : 4) == 0) && (reinterpret_cast<uintptr_t>(b) & 15) == 0;
  if (b_is_bf16)
    launch_k(tn_gemm_small_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, a, lda, static_cast<const __nv_bfloat16*>(b), out, R, J, D, alpha,
                                                             accumulate, a_vec, b_vec);
  else
    launch_k(tn_gemm_small_kernel<float>, dim3(grid), dim3(256), 0, s, a, lda, static_cast<const float*>(b), out, R, J, D, alpha, accumulate, a_vec, b_vec);
  count_launch();
  return check_launch("tn_gemm_small_kernel");
}
