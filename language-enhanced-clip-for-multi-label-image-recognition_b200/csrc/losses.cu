// Fused loss forward+backward kernels: logits -> (scalar loss, dlogits) in one launch.
//  * asymmetric loss (U:126-173; ASL_loss U:184-190, dualcoop_loss U:175-181): elementwise, HBM-bound,
//    128-bit vectorised with a warp-shuffle + one-atomic-per-CTA reduction.  The focal weight
//    (1-p_t)^gamma is a constant in the backward (computed under set_grad_enabled(False), U:162-170).
//  * pairwise ranking hinge (U:85-93) and its co-occurrence weighted form (U:95-110): one warp per row over the
//    compacted list of non-zero targets, no [B,K,K] temporary.
//  * the EMA consistency KL terms of T:809-813 (log_softmax / softmax / KLDivLoss batchmean) fwd + bwd.
#include <stdlib.h>

#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

struct AslParams {
  float gamma_neg, gamma_pos, clip, eps, thresh_pos, thresh_neg, inv_denom;
};

__device__ __forceinline__ float pow_gamma(float base, float g) {
  // torch.pow semantics incl. pow(x, 0) == 1; gamma is a small non-negative number (0, 1, 2, 4 typical)
  if (g == 0.f) return 1.f;
  if (g == 1.f) return base;
  if (g == 2.f) return base * base;
  return powf(base, g);
}

// MUFU approximations with flush-to-zero: the plain intrinsics (__expf / __logf / __fdividef without -ftz) wrap every MUFU in a
// denormal fix-up (compare, predicated scale, predicated undo: 3-4 extra instructions each), and none of the arguments below can
// be denormal where it matters (1 + e >= 1, probabilities are clamped to eps first)
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// kLean (host-checked: gamma_pos = 1, gamma_neg = 2, thresh_neg <= thresh_pos so that a target is never positive AND negative,
// eps a normal number): branch-free, one logarithm on the selected probability, four MUFU operations per element.  Same
// expressions as the general form below, which stays for every other parameter set.
template <bool kFastGamma, bool kLean>
__device__ __forceinline__ void asl_elem(float x, float y, const AslParams& p, float& loss, float& grad) {
  if (kLean) {
    const float e = ex2_ftz(x * -1.4426950408889634f);
    const float s = rcp_ftz(1.0f + e);                                  // e = +inf -> 0
    const float oms = 1.0f - s;
    const float sneg_raw = oms + p.clip;
    const bool has_clip = p.clip > 0.f;
    const bool clipped = has_clip && sneg_raw > 1.0f;
    const float sneg = has_clip ? fminf(sneg_raw, 1.0f) : oms;
    const bool pos = y > p.thresh_pos, neg = y < p.thresh_neg;          // mutually exclusive here
    const float sel = pos ? s : sneg;                                   // p_t of a live element
    const float lg = 0.6931471805599453f * lg2_ftz(fmaxf(sel, p.eps));
    const float base = 1.0f - sel;
    const float w = pos ? base : base * base;                           // (1 - p_t)^gamma, gamma = 1 / 2
    const float dneg = (!clipped && sneg >= p.eps) ? -s * oms * rcp_ftz(sneg) : 0.f;     // d log(1 - s + clip) / dx
    const float dpos = s >= p.eps ? oms : 0.f;                                            // d log(s) / dx
    const float scale = (pos || neg) ? -w * p.inv_denom : 0.f;
    loss = lg * scale;
    grad = (pos ? dpos : dneg) * scale;
    return;
  }
  const float e = __expf(-x);
  const float s = __fdividef(1.0f, 1.0f + e);
  const float sneg_raw = 1.0f - s + p.clip;
  const bool clipped = (p.clip > 0.f) && (sneg_raw > 1.0f);
  const float sneg = (p.clip > 0.f) ? fminf(sneg_raw, 1.0f) : (1.0f - s);
  const bool pos = y > p.thresh_pos, neg = y < p.thresh_neg;
  // only one of the two log terms is live for binary targets: evaluate a single log on the selected probability
  const float lp = pos ? __logf(fmaxf(s, p.eps)) : 0.f;
  const float ln = neg ? __logf(fmaxf(sneg, p.eps)) : 0.f;
  const float ypos = pos ? 1.f : 0.f, yneg = neg ? 1.f : 0.f;
  const float pt = s * ypos + sneg * yneg;
  const float base = 1.0f - pt;
  float w;
  if (kFastGamma) {                       // gamma_pos = 1, gamma_neg = 2 (ASL_loss / dualcoop_loss, U:179,188)
    w = pos ? base : 1.0f;
    w = neg ? (pos ? w * base * base : base * base) : w;
  } else {
    w = (p.gamma_neg > 0.f || p.gamma_pos > 0.f) ? pow_gamma(base, p.gamma_pos * ypos + p.gamma_neg * yneg) : 1.0f;
  }
  loss = -(lp + ln) * w * p.inv_denom;
  const float dpos = (pos && s >= p.eps) ? (1.0f - s) : 0.f;                              // d log(sigmoid)/dx
  const float dneg = (neg && !clipped && sneg >= p.eps) ? __fdividef(-s * (1.0f - s), sneg) : 0.f;   // d log(1-s+clip)/dx
  grad = -(dpos + dneg) * w * p.inv_denom;
}

template <bool kFastGamma, bool kLean>
__global__ void __launch_bounds__(256)
asl_fwd_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ grad,
                   float* __restrict__ loss_out, int64_t n, AslParams p) {
  pdl_grid_sync();
  float acc = 0.f;
  const int64_t nvec = n / 4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const float4* y4 = reinterpret_cast<const float4*>(y);
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  // two independent 128-bit load pairs in flight per thread
  for (; i + stride < nvec; i += 2 * stride) {
    const float4 xa = __ldcs(x4 + i), ya = __ldcs(y4 + i);
    const float4 xb = __ldcs(x4 + i + stride), yb = __ldcs(y4 + i + stride);
    float4 ga, gb;
    float l;
    asl_elem<kFastGamma, kLean>(xa.x, ya.x, p, l, ga.x); acc += l;
    asl_elem<kFastGamma, kLean>(xa.y, ya.y, p, l, ga.y); acc += l;
    asl_elem<kFastGamma, kLean>(xa.z, ya.z, p, l, ga.z); acc += l;
    asl_elem<kFastGamma, kLean>(xa.w, ya.w, p, l, ga.w); acc += l;
    asl_elem<kFastGamma, kLean>(xb.x, yb.x, p, l, gb.x); acc += l;
    asl_elem<kFastGamma, kLean>(xb.y, yb.y, p, l, gb.y); acc += l;
    asl_elem<kFastGamma, kLean>(xb.z, yb.z, p, l, gb.z); acc += l;
    asl_elem<kFastGamma, kLean>(xb.w, yb.w, p, l, gb.w); acc += l;
    if (grad) {
      __stcs(reinterpret_cast<float4*>(grad) + i, ga);
      __stcs(reinterpret_cast<float4*>(grad) + i + stride, gb);
    }
  }
  for (; i < nvec; i += stride) {
    const float4 xv = __ldcs(x4 + i), yv = __ldcs(y4 + i);
    float4 g;
    float l;
    asl_elem<kFastGamma, kLean>(xv.x, yv.x, p, l, g.x); acc += l;
    asl_elem<kFastGamma, kLean>(xv.y, yv.y, p, l, g.y); acc += l;
    asl_elem<kFastGamma, kLean>(xv.z, yv.z, p, l, g.z); acc += l;
    asl_elem<kFastGamma, kLean>(xv.w, yv.w, p, l, g.w); acc += l;
    if (grad) __stcs(reinterpret_cast<float4*>(grad) + i, g);
  }
  for (int64_t t = nvec * 4 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += stride) {
    float l, g;
    asl_elem<kFastGamma, kLean>(x[t], y[t], p, l, g);
    acc += l;
    if (grad) grad[t] = g;
  }
  __shared__ float s_part[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? s_part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(loss_out, v);
  }
}

// Pairwise ranking hinge (U:85-93) and its co-occurrence weighted form (U:95-110):
//   loss_b = sum_{i,j} relu(margin - s*y_j + s*y_i) * t_j * (1 - t_i) * Wt[i,j]          (Wt = 1 without co-occurrence)
// ONE WARP PER ROW.  Only pairs whose weight t_j (1 - t_i) is non-zero matter, and multi-label targets are sparse (1-5
// positives of 80), so the warp first compacts the columns with t_j != 0 into shared memory (ballot + popcount) and
// then every lane walks that short list for its own columns i: O(K * n_pos) pair evaluations per row instead of the
// O(K^2) loop of round 1 (2.06 ms at [2^19, 80] = 0.037 of HBM).  The gradient of a "positive" column j is the negated
// sum of the pair indicators over all i: one warp reduction per list entry.  Targets may be any floats (the reference
// multiplies by them), i == j pairs included, exactly like the reference's dense [B,K,K] expression.
constexpr int kRankWarps = 8;
constexpr int kRankMaxK = 512;       // columns per row a warp keeps in registers (kRankChunks * 32); 3 lists of K floats per warp in smem
constexpr int kRankChunks = kRankMaxK / 32;
constexpr int kRankSmemWtK = 96;     // co-occurrence: the [K,K] pair weights are staged (transposed) in shared memory up to this K

template <int CH, bool kCooc>
__global__ void __launch_bounds__(kRankWarps * 32)
ranking_fwd_bwd_kernel(const float* __restrict__ ypred, const float* __restrict__ ytrue, const float* __restrict__ wt,
                       float* __restrict__ grad, float* __restrict__ loss_out, int64_t B, int K, float scale, float margin,
                       float inv_batch) {
  pdl_grid_sync();
  extern __shared__ float sm[];                  // per warp: [K] scaled score of list entry, [K] its target, [K] its column;
                                                 // then (co-occurrence, K <= kRankSmemWtK) the pair weights, transposed
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* l_y = sm + warp * 3 * K;
  float* l_t = l_y + K;
  int* l_j = reinterpret_cast<int*>(l_t + K);
  const float* wt_t = nullptr;                   // wt_t[j * K + i] = Wt[i, j]: lanes walk i, so the transposed copy is conflict-free
  if (kCooc && K <= kRankSmemWtK) {
    float* dst = sm + kRankWarps * 3 * K;
    for (int e = threadIdx.x; e < K * K; e += blockDim.x) {
      const int i = e / K, j = e - i * K;
      dst[j * K + i] = __ldg(wt + e);
    }
    wt_t = dst;
    __syncthreads();
  }
  float acc = 0.f;
  // The row loop is software-pipelined: the next row's scores and targets are requested before this row is processed, so a
  // warp always has two rows of loads in flight (one row per warp left the kernel latency-bound at 0.27 of HBM peak).
  const int64_t row_step = static_cast<int64_t>(gridDim.x) * kRankWarps;
  float yn[CH], tn[CH];
  {
    const int64_t b0 = static_cast<int64_t>(blockIdx.x) * kRankWarps + warp;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int k = c * 32 + lane;
      const bool in = k < K && b0 < B;
      yn[c] = in ? __ldcs(ypred + b0 * K + k) : 0.f;
      tn[c] = in ? __ldcs(ytrue + b0 * K + k) : 0.f;
    }
  }
  for (int64_t b = static_cast<int64_t>(blockIdx.x) * kRankWarps + warp; b < B; b += row_step) {
    float y[CH], t[CH], g[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      y[c] = yn[c] * scale;
      t[c] = tn[c];
      g[c] = 0.f;
    }
    {
      const int64_t bn = b + row_step;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int k = c * 32 + lane;
        const bool in = k < K && bn < B;
        yn[c] = in ? __ldcs(ypred + bn * K + k) : 0.f;
        tn[c] = in ? __ldcs(ytrue + bn * K + k) : 0.f;
      }
    }
    bool binary = true;
#pragma unroll
    for (int c = 0; c < CH; ++c) binary = binary && (t[c] == 0.f || t[c] == 1.f);
    binary = __all_sync(0xffffffffu, binary);
    if (!kCooc && binary) {
      // Multi-hot {0,1} targets (every shipped configuration) without pair weights: the pair weight is the indicator
      // "j positive, i negative".  No list in shared memory: the positives of chunk cj are the set bits of one ballot,
      // each one's score is broadcast with a shuffle, every lane tests its own negative columns (positives and the
      // columns past K carry -inf, so their hinge is never active), and the positive's gradient is the COUNT of active
      // hinges — one integer warp reduction (REDUX) per positive instead of a float shuffle tree.
      float yneg[CH];
      int gi[CH];                                 // active-hinge counts, kept as integers until the row is stored
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        yneg[c] = (c * 32 + lane < K && t[c] == 0.f) ? y[c] : -INFINITY;
        gi[c] = 0;
      }
#pragma unroll
      for (int cj = 0; cj < CH; ++cj) {
        unsigned m = __ballot_sync(0xffffffffu, t[cj] != 0.f);     // columns past K were loaded as 0
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const float base = margin - __shfl_sync(0xffffffffu, y[cj], src);
          int n = 0;
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const float h = base + yneg[c];
            const int hit = h > 0.f ? 1 : 0;
            acc += fmaxf(h, 0.f);                 // -inf (not a negative column) contributes 0
            gi[c] += hit;
            n += hit;
          }
          const int cnt = __reduce_add_sync(0xffffffffu, n);
          gi[cj] -= lane == src ? cnt : 0;
        }
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) g[c] = static_cast<float>(gi[c]);
    } else {
      int n_pos = 0;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int k = c * 32 + lane;
        const bool in = k < K;
        const unsigned m = __ballot_sync(0xffffffffu, in && t[c] != 0.f);
        if (in && t[c] != 0.f) {
          const int slot = n_pos + __popc(m & ((1u << lane) - 1u));
          l_y[slot] = y[c];
          l_t[slot] = t[c];
          l_j[slot] = k;
        }
        n_pos += __popc(m);
      }
      __syncwarp();
      for (int q = 0; q < n_pos; ++q) {
        const float yj = l_y[q], tj = l_t[q];
        const int j = l_j[q];
        float gj = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int i = c * 32 + lane;
          if (i < K) {
            const float h = margin - yj + y[c];
            float w = tj * (1.0f - t[c]);
            if (kCooc) w *= wt_t ? wt_t[j * K + i] : __ldg(wt + static_cast<int64_t>(i) * K + j);
            if (h > 0.f) {
              acc += h * w;
              g[c] += w;
              gj += w;
            }
          }
        }
        gj = warp_sum(gj);
        // the column that owns j subtracts the summed indicator (d relu(m - y_j + y_i) / d y_j = -1)
#pragma unroll
        for (int c = 0; c < CH; ++c) g[c] -= (c * 32 + lane == j) ? gj : 0.f;
      }
      __syncwarp();                             // the list is rebuilt for the next row
    }
    if (grad) {
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int k = c * 32 + lane;
        if (k < K) __stcs(grad + b * K + k, scale * inv_batch * g[c]);
      }
    }
  }
  __shared__ float s_part[kRankWarps];
  acc = warp_sum(acc);
  if (lane == 0) s_part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kRankWarps; ++w) v += s_part[w];
    atomicAdd(loss_out, v * inv_batch);
  }
}

// KL(softmax(xm) || softmax(x)) with reduction "batchmean" — `nn.KLDivLoss(reduction="batchmean")(log_softmax(x), softmax(xm))`,
// the EMA consistency terms of T:809-813 — and its gradient w.r.t. x: weight * (softmax(x) - softmax(xm)) / B.
// One warp per row, both softmaxes from registers; the target branch carries no gradient (it is computed under no_grad).
template <int CH>
__global__ void __launch_bounds__(kRankWarps * 32)
kl_softmax_fwd_bwd_kernel(const float* __restrict__ x, const float* __restrict__ xm, float* __restrict__ grad,
                          float* __restrict__ loss_out, int64_t B, int K, float weight, float inv_batch) {
  pdl_grid_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  const int64_t row_step = static_cast<int64_t>(gridDim.x) * kRankWarps;
  float an[CH], mn[CH];                         // next row, requested one iteration ahead (two rows of loads in flight)
  {
    const int64_t b0 = static_cast<int64_t>(blockIdx.x) * kRankWarps + warp;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int k = c * 32 + lane;
      const bool in = k < K && b0 < B;
      an[c] = in ? __ldcs(x + b0 * K + k) : -INFINITY;
      mn[c] = in ? __ldcs(xm + b0 * K + k) : -INFINITY;
    }
  }
  for (int64_t b = static_cast<int64_t>(blockIdx.x) * kRankWarps + warp; b < B; b += row_step) {
    float a[CH], m[CH];
    float amax = -INFINITY, mmax = -INFINITY;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      a[c] = an[c];
      m[c] = mn[c];
      amax = fmaxf(amax, a[c]);
      mmax = fmaxf(mmax, m[c]);
    }
    {
      const int64_t bn = b + row_step;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int k = c * 32 + lane;
        const bool in = k < K && bn < B;
        an[c] = in ? __ldcs(x + bn * K + k) : -INFINITY;
        mn[c] = in ? __ldcs(xm + bn * K + k) : -INFINITY;
      }
    }
    amax = warp_max(amax);
    mmax = warp_max(mmax);
    float ea[CH], em[CH];
    float asum = 0.f, msum = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int k = c * 32 + lane;
      // (flush-to-zero MUFU forms, see asl_elem: the arguments are <= 0 and a term below 2^-126 of the row maximum is nothing)
      ea[c] = k < K ? ex2_ftz((a[c] - amax) * 1.4426950408889634f) : 0.f;          // kept: p and q below are these over the row sums
      em[c] = k < K ? ex2_ftz((m[c] - mmax) * 1.4426950408889634f) : 0.f;
      asum += ea[c];
      msum += em[c];
    }
    asum = warp_sum(asum);
    msum = warp_sum(msum);
    const float la = 0.6931471805599453f * lg2_ftz(asum), lm = 0.6931471805599453f * lg2_ftz(msum);      // sums in [1, K]
    const float ra = rcp_ftz(asum), rm = rcp_ftz(msum);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int k = c * 32 + lane;
      if (k < K) {
        const float logp = a[c] - amax - la, logq = m[c] - mmax - lm;
        const float q = em[c] * rm;
        if (q > 0.f) acc += q * (logq - logp);                 // xlogy: a zero target contributes nothing
        if (grad) __stcs(grad + b * K + k, weight * inv_batch * (ea[c] * ra - q));
      }
    }
  }
  __shared__ float s_part[kRankWarps];
  acc = warp_sum(acc);
  if (lane == 0) s_part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kRankWarps; ++w) v += s_part[w];
    atomicAdd(loss_out, v * weight * inv_batch);
  }
}


// Distribution-balanced loss (`ResampleLoss`, trainers/dbl.py:263-445, built at T:818-830 for LOSSFUNC 'dbl') with
// use_sigmoid=True, partial=False: re-balanced weights (dbl.py:411-416), logit regularisation (dbl.py:401-409), weighted
// binary cross-entropy with logits averaged over all B*K elements (dbl.py:49-65) and the optional focal factor — which in
// the reference multiplies two SCALARS, because its `binary_cross_entropy` always reduces with 'mean' (dbl.py:61-63, 373-383).
// One warp per row (the repeat rate sum_k y_k / freq_k is a row reduction).
//   kMode 0: no focal term: loss and dloss/dlogits in one pass.
//   kMode 1: focal, pass 1: sums[0] += unweighted BCE, sums[1] += weighted BCE (no gradient).
//   kMode 2: focal, pass 2: loss = bp (1 - e^-L0)^g Lw with L0 = sums[0] / n, Lw = sums[1] / n;
//            grad = dLw * bp (1 - e^-L0)^g + dL0 * bp g (1 - e^-L0)^(g-1) e^-L0 Lw.
struct ResampleParams {
  float map_alpha, map_beta, map_gamma, neg_scale, focal_gamma, balance_param, loss_weight, inv_n;
  int reweight, has_neg_scale;
};

template <int CH, int kMode>
__global__ void __launch_bounds__(kRankWarps * 32)
resample_bce_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ freq_inv,
                    const float* __restrict__ init_bias, float* __restrict__ grad, float* __restrict__ loss_out,
                    float* __restrict__ sums, int64_t B, int K, ResampleParams p) {
  pdl_grid_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc0 = 0.f, accw = 0.f;
  float c_w = p.loss_weight, c_0 = 0.f;          // gradient coefficients of the weighted / unweighted mean
  if (kMode == 2) {
    const float l0 = sums[0] * p.inv_n, lw = sums[1] * p.inv_n;
    const float q = 1.0f - expf(-l0);
    const float f = p.balance_param * powf(q, p.focal_gamma);
    const float df = p.balance_param * p.focal_gamma * powf(q, p.focal_gamma - 1.0f) * expf(-l0);
    c_w = p.loss_weight * f;
    c_0 = p.loss_weight * df * lw;
    if (blockIdx.x == 0 && threadIdx.x == 0) *loss_out = p.loss_weight * f * lw;
  }
  for (int64_t b = static_cast<int64_t>(blockIdx.x) * kRankWarps + warp; b < B; b += static_cast<int64_t>(gridDim.x) * kRankWarps) {
    float xv[CH], yv[CH], fi[CH];
    float rr = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int k = c * 32 + lane;
      const bool in = k < K;
      xv[c] = in ? __ldcs(x + b * K + k) : 0.f;
      yv[c] = in ? __ldcs(y + b * K + k) : 0.f;
      fi[c] = (in && p.reweight) ? __ldg(freq_inv + k) : 0.f;
      rr += yv[c] * fi[c];
    }
    if (p.reweight) rr = warp_sum(rr);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int k = c * 32 + lane;
      if (k < K) {
        float w = 1.0f;
        if (p.reweight) {      // sigmoid(beta * (freq_inv / repeat_rate - gamma)) + alpha; a row without positives divides by zero -> 1 + alpha
          const float pw = fi[c] / rr;
          w = 1.0f / (1.0f + expf(-p.map_beta * (pw - p.map_gamma))) + p.map_alpha;
        }
        float z = xv[c] + (init_bias ? __ldg(init_bias + k) : 0.f);
        float dz = 1.0f;
        if (p.has_neg_scale) {
          dz = (1.0f - yv[c]) * p.neg_scale + yv[c];
          z = z * (1.0f - yv[c]) * p.neg_scale + z * yv[c];
          w = w / p.neg_scale * (1.0f - yv[c]) + w * yv[c];
        }
        // binary_cross_entropy_with_logits: max(z, 0) - z y + log(1 + exp(-|z|));  d/dz = sigmoid(z) - y
        const float e = expf(-fabsf(z));
        const float bce = fmaxf(z, 0.f) - z * yv[c] + log1pf(e);
        const float sg = z >= 0.f ? 1.0f / (1.0f + e) : e / (1.0f + e);
        acc0 += bce;
        accw += w * bce;
        if (kMode != 1 && grad) __stcs(grad + b * K + k, (c_w * w + c_0) * (sg - yv[c]) * dz * p.inv_n);
      }
    }
  }
  if (kMode == 2) return;
  __shared__ float s0[kRankWarps], sw[kRankWarps];
  acc0 = warp_sum(acc0);
  accw = warp_sum(accw);
  if (lane == 0) {
    s0[warp] = acc0;
    sw[warp] = accw;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, bsum = 0.f;
#pragma unroll
    for (int w = 0; w < kRankWarps; ++w) {
      a += s0[w];
      bsum += sw[w];
    }
    if (kMode == 0) {
      atomicAdd(loss_out, p.loss_weight * bsum * p.inv_n);
    } else {
      atomicAdd(sums, a);
      atomicAdd(sums + 1, bsum);
    }
  }
}

template <int kMode>
static int launch_resample(const float* x, const float* y, const float* freq_inv, const float* init_bias, float* grad, float* loss,
                           float* sums, int64_t B, int K, const ResampleParams& p, cudaStream_t s) {
  int64_t blocks = (B + kRankWarps - 1) / kRankWarps;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  const int chunks = (K + 31) / 32;
  if (chunks <= 3)
    launch_k(resample_bce_kernel<3, kMode>, dim3(static_cast<unsigned>(blocks)), dim3(kRankWarps * 32), 0, s, x, y, freq_inv, init_bias, grad, loss, sums, B, K, p);
  else
    launch_k(resample_bce_kernel<8, kMode>, dim3(static_cast<unsigned>(blocks)), dim3(kRankWarps * 32), 0, s, x, y, freq_inv, init_bias, grad, loss, sums, B, K, p);
  count_launch();
  return check_launch("resample_bce_kernel");
}

template <bool kCooc>
static int launch_ranking(const float* logits, const float* targets, const float* wt, float* grad, float* loss, int64_t B,
                          int K, float scale, float margin, cudaStream_t s) {
  const int chunks = (K + 31) / 32;
  size_t smem = static_cast<size_t>(kRankWarps) * 3 * K * sizeof(float);
  if (kCooc && K <= kRankSmemWtK) smem += static_cast<size_t>(K) * K * sizeof(float);
  int64_t blocks = (B + kRankWarps - 1) / kRankWarps;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  const float inv_b = 1.0f / static_cast<float>(B);
#define LECB_RANK_CASE(CH)                                                                                        \
  launch_k(ranking_fwd_bwd_kernel<CH, kCooc>, dim3(static_cast<unsigned>(blocks)), dim3(kRankWarps * 32), smem, s, logits, targets, wt, grad, \
                                                                                                loss, B, K, scale, margin, inv_b)
  if (chunks <= 1) LECB_RANK_CASE(1);
  else if (chunks <= 2) LECB_RANK_CASE(2);
  else if (chunks <= 3) LECB_RANK_CASE(3);
  else if (chunks <= 4) LECB_RANK_CASE(4);
  else if (chunks <= 8) LECB_RANK_CASE(8);
  else LECB_RANK_CASE(kRankChunks);
#undef LECB_RANK_CASE
  count_launch();
  return check_launch("ranking_fwd_bwd_kernel");
}

}  // namespace lecb

using namespace lecb;

extern "C" int lecb_asl_fwd_bwd(const float* logits, const float* targets, float* grad, float* loss, int64_t B, int K,
                                float gamma_neg, float gamma_pos, float clip, float eps, float thresh_pos,
                                float thresh_neg, int partial, void* stream) {
  LECB_CHECK_ARG(logits && targets && loss, "lecb_asl_fwd_bwd: null pointer");
  LECB_CHECK_ARG(B > 0 && K > 0, "lecb_asl_fwd_bwd: empty problem");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t n = B * K;
  AslParams p{gamma_neg, gamma_pos, clip, eps, thresh_pos, thresh_neg,
              partial ? 1.0f / static_cast<float>(B) : 1.0f / static_cast<float>(n)};
  cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), s);
  if (e != cudaSuccess) return fail(LECB_ERR_CUDA, "lecb_asl_fwd_bwd: memset: %s", cudaGetErrorString(e));
  int64_t blocks = (n / 4 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks < 1) blocks = 1;
  if (blocks > cap) blocks = cap;
  static const bool no_lean = getenv("LECB_ASL_GENERAL") != nullptr;      // A/B and cross-check switch
  const bool fast_gamma = gamma_pos == 1.0f && gamma_neg == 2.0f;
  if (fast_gamma && thresh_neg <= thresh_pos && eps >= 1e-30f && !no_lean)
    launch_k(asl_fwd_bwd_kernel<true, true>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, s, logits, targets, grad, loss, n, p);
  else if (fast_gamma)
    launch_k(asl_fwd_bwd_kernel<true, false>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, s, logits, targets, grad, loss, n, p);
  else
    launch_k(asl_fwd_bwd_kernel<false, false>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, s, logits, targets, grad, loss, n, p);
  count_launch();
  return check_launch("asl_fwd_bwd_kernel");
}

extern "C" int lecb_ranking_fwd_bwd(const float* logits, const float* targets, float* grad, float* loss, int B, int K,
                                    float scale, float margin, void* stream) {
  LECB_CHECK_ARG(logits && targets && loss, "lecb_ranking_fwd_bwd: null pointer");
  LECB_CHECK_ARG(B > 0 && K > 0 && K <= kRankMaxK, "lecb_ranking_fwd_bwd: need 0 < K <= %d", kRankMaxK);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), s);
  if (e != cudaSuccess) return fail(LECB_ERR_CUDA, "lecb_ranking_fwd_bwd: memset: %s", cudaGetErrorString(e));
  return launch_ranking<false>(logits, targets, nullptr, grad, loss, B, K, scale, margin, s);
}

extern "C" int lecb_ranking_cooc_fwd_bwd(const float* logits, const float* targets, const float* pair_weights, float* grad,
                                         float* loss, int B, int K, float scale, float margin, void* stream) {
  LECB_CHECK_ARG(logits && targets && pair_weights && loss, "lecb_ranking_cooc_fwd_bwd: null pointer");
  LECB_CHECK_ARG(B > 0 && K > 0 && K <= kRankMaxK, "lecb_ranking_cooc_fwd_bwd: need 0 < K <= %d", kRankMaxK);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), s);
  if (e != cudaSuccess) return fail(LECB_ERR_CUDA, "lecb_ranking_cooc_fwd_bwd: memset: %s", cudaGetErrorString(e));
  return launch_ranking<true>(logits, targets, pair_weights, grad, loss, B, K, scale, margin, s);
}

extern "C" int lecb_kl_softmax_fwd_bwd(const float* logits, const float* logits_target, float* grad, float* loss, int64_t B,
                                       int K, float weight, void* stream) {
  LECB_CHECK_ARG(logits && logits_target && loss, "lecb_kl_softmax_fwd_bwd: null pointer");
  LECB_CHECK_ARG(B > 0 && K > 0 && K <= 256, "lecb_kl_softmax_fwd_bwd: need 0 < K <= 256");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), s);
  if (e != cudaSuccess) return fail(LECB_ERR_CUDA, "lecb_kl_softmax_fwd_bwd: memset: %s", cudaGetErrorString(e));
  int64_t blocks = (B + kRankWarps - 1) / kRankWarps;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  const float inv_b = 1.0f / static_cast<float>(B);
  const int chunks = (K + 31) / 32;
  if (chunks <= 3)
    launch_k(kl_softmax_fwd_bwd_kernel<3>, dim3(static_cast<unsigned>(blocks)), dim3(kRankWarps * 32), 0, s, logits, logits_target, grad, loss, B, K, weight, inv_b);
  else
    launch_k(kl_softmax_fwd_bwd_kernel<8>, dim3(static_cast<unsigned>(blocks)), dim3(kRankWarps * 32), 0, s, logits, logits_target, grad, loss, B, K, weight, inv_b);
  count_launch();
  return check_launch("kl_softmax_fwd_bwd_kernel");
}

extern "C" int lecb_resample_bce_fwd_bwd(const float* logits, const float* labels, const float* freq_inv, const float* init_bias,
                                         float* grad, float* loss, float* scratch2, int64_t B, int K, float map_alpha,
                                         float map_beta, float map_gamma, float neg_scale, int focal, float focal_gamma,
                                         float balance_param, float loss_weight, void* stream) {
  LECB_CHECK_ARG(logits && labels && loss, "lecb_resample_bce_fwd_bwd: null pointer");
  LECB_CHECK_ARG(B > 0 && K > 0 && K <= 256, "lecb_resample_bce_fwd_bwd: need 0 < K <= 256");
  LECB_CHECK_ARG(!focal || scratch2, "lecb_resample_bce_fwd_bwd: the focal variant needs two floats of scratch");
  LECB_CHECK_ARG(neg_scale == 0.f || freq_inv, "lecb_resample_bce_fwd_bwd: neg_scale rescales the weights: it needs reweighting (freq_inv)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ResampleParams p{map_alpha, map_beta, map_gamma, neg_scale, focal_gamma, balance_param, loss_weight,
                   1.0f / static_cast<float>(B * K), freq_inv ? 1 : 0, neg_scale != 0.f ? 1 : 0};
  cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), s);
  if (e == cudaSuccess && focal) e = cudaMemsetAsync(scratch2, 0, 2 * sizeof(float), s);
  if (e != cudaSuccess) return fail(LECB_ERR_CUDA, "lecb_resample_bce_fwd_bwd: memset: %s", cudaGetErrorString(e));
  if (!focal) return launch_resample<0>(logits, labels, freq_inv, init_bias, grad, loss, nullptr, B, K, p, s);
  int st = launch_resample<1>(logits, labels, freq_inv, init_bias, nullptr, loss, scratch2, B, K, p, s);
  if (st) return st;
  return launch_resample<2>(logits, labels, freq_inv, init_bias, grad, loss, scratch2, B, K, p, s);
}
