// Fused loss forward+backward kernels: logits -> (scalar loss, dlogits) in one launch.
//  * asymmetric loss (U:126-173; ASL_loss U:184-190, dualcoop_loss U:175-181): elementwise, HBM-bound,
//    128-bit vectorised with a warp-shuffle + one-atomic-per-CTA reduction.  The focal weight
//    (1-p_t)^gamma is a constant in the backward (computed under set_grad_enabled(False), U:162-170).
//  * pairwise ranking hinge (U:85-93): one CTA per row, no [B,K,K] temporary.
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

template <bool kFastGamma>
__device__ __forceinline__ void asl_elem(float x, float y, const AslParams& p, float& loss, float& grad) {
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

template <bool kFastGamma>
__global__ void __launch_bounds__(256)
asl_fwd_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ grad,
                   float* __restrict__ loss_out, int64_t n, AslParams p) {
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
    asl_elem<kFastGamma>(xa.x, ya.x, p, l, ga.x); acc += l;
    asl_elem<kFastGamma>(xa.y, ya.y, p, l, ga.y); acc += l;
    asl_elem<kFastGamma>(xa.z, ya.z, p, l, ga.z); acc += l;
    asl_elem<kFastGamma>(xa.w, ya.w, p, l, ga.w); acc += l;
    asl_elem<kFastGamma>(xb.x, yb.x, p, l, gb.x); acc += l;
    asl_elem<kFastGamma>(xb.y, yb.y, p, l, gb.y); acc += l;
    asl_elem<kFastGamma>(xb.z, yb.z, p, l, gb.z); acc += l;
    asl_elem<kFastGamma>(xb.w, yb.w, p, l, gb.w); acc += l;
    if (grad) {
      __stcs(reinterpret_cast<float4*>(grad) + i, ga);
      __stcs(reinterpret_cast<float4*>(grad) + i + stride, gb);
    }
  }
  for (; i < nvec; i += stride) {
    const float4 xv = __ldcs(x4 + i), yv = __ldcs(y4 + i);
    float4 g;
    float l;
    asl_elem<kFastGamma>(xv.x, yv.x, p, l, g.x); acc += l;
    asl_elem<kFastGamma>(xv.y, yv.y, p, l, g.y); acc += l;
    asl_elem<kFastGamma>(xv.z, yv.z, p, l, g.z); acc += l;
    asl_elem<kFastGamma>(xv.w, yv.w, p, l, g.w); acc += l;
    if (grad) __stcs(reinterpret_cast<float4*>(grad) + i, g);
  }
  for (int64_t t = nvec * 4 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += stride) {
    float l, g;
    asl_elem<kFastGamma>(x[t], y[t], p, l, g);
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

// One CTA per batch row.  tmp[i,j] = margin - s*y_j + s*y_i;  loss += relu(tmp) * t_j * (1 - t_i).
__global__ void __launch_bounds__(128)
ranking_fwd_bwd_kernel(const float* __restrict__ ypred, const float* __restrict__ ytrue, float* __restrict__ grad,
                       float* __restrict__ loss_out, int K, float scale, float margin, float inv_batch) {
  extern __shared__ float sm[];
  float* sy = sm;
  float* st = sm + K;
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    sy[k] = ypred[static_cast<int64_t>(b) * K + k] * scale;
    st[k] = ytrue[static_cast<int64_t>(b) * K + k];
  }
  __syncthreads();
  float acc = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float yk = sy[k], tk = st[k];
    float as_i = 0.f, cnt_i = 0.f, cnt_j = 0.f;
    for (int o = 0; o < K; ++o) {
      const float h_ko = margin - sy[o] + yk;      // k plays i (the "negative" side)
      if (h_ko > 0.f) {
        as_i += h_ko * st[o];
        cnt_i += st[o];
      }
      const float h_ok = margin - yk + sy[o];      // k plays j (the "positive" side)
      if (h_ok > 0.f) cnt_j += 1.0f - st[o];
    }
    acc += as_i * (1.0f - tk);
    if (grad) grad[static_cast<int64_t>(b) * K + k] = scale * inv_batch * ((1.0f - tk) * cnt_i - tk * cnt_j);
  }
  __shared__ float s_part[4];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(loss_out, (s_part[0] + s_part[1] + s_part[2] + s_part[3]) * inv_batch);
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
  if (gamma_pos == 1.0f && gamma_neg == 2.0f)
    asl_fwd_bwd_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, s>>>(logits, targets, grad, loss, n, p);
  else
    asl_fwd_bwd_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, s>>>(logits, targets, grad, loss, n, p);
  count_launch();
  return check_launch("asl_fwd_bwd_kernel");
}

extern "C" int lecb_ranking_fwd_bwd(const float* logits, const float* targets, float* grad, float* loss, int B, int K,
                                    float scale, float margin, void* stream) {
  LECB_CHECK_ARG(logits && targets && loss, "lecb_ranking_fwd_bwd: null pointer");
  LECB_CHECK_ARG(B > 0 && K > 0 && K <= 4096, "lecb_ranking_fwd_bwd: need 0 < K <= 4096");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), s);
  if (e != cudaSuccess) return fail(LECB_ERR_CUDA, "lecb_ranking_fwd_bwd: memset: %s", cudaGetErrorString(e));
  ranking_fwd_bwd_kernel<<<B, 128, 2 * K * sizeof(float), s>>>(logits, targets, grad, loss, K, scale, margin,
                                                               1.0f / static_cast<float>(B));
  count_launch();
  return check_launch("ranking_fwd_bwd_kernel");
}
