// Dual-prompt head: winner-take-all + spatial-softmax aggregation of the patch-by-class similarity
// maps (T:456-470 test / T:496-514 train) and the global logits (T:453-455).  The patch-by-class
// contraction itself runs on tcgen05 (lecb_gemm_bf16 against the concatenated [pos;neg;evi] prompt
// matrix, fp32 output); these kernels consume its raw dot products with the per-row inverse norms, so
// the L2 normalisation of the local features (T:442 / T:486) never touches HBM as a separate pass.
#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

constexpr int kAggWarps = 8;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One CTA per image / caption.  Warp w streams rows p = w, w+8, ...; lane l owns classes l, l+32, ...
// The kernel is instruction-issue bound long before it is HBM bound (ncu: 91 % issue slots busy in the first
// version), so everything per row is kept to the minimum: compile-time variants for evidence / maps / mask,
// running pointers instead of per-row index arithmetic, one REDUX for the class max (gain >= 0 for cosines, so
// max_k gain*neg_k = gain*max_k neg_k), one shuffle tree for the class sum, exp2 with pre-folded log2(e), and a
// single-MUFU online spatial-softmax update per (row, class): of the two factors exp(m-m') and exp(t-m') one is
// always 1.  The next row's operands are requested before the current row's reduction chain starts.
template <int kJ, bool kEvi, bool kMaps, bool kMask>
__global__ void __launch_bounds__(kAggWarps * 32)
head_aggregate_kernel(const float* __restrict__ dots, int ldn, const float* __restrict__ row_sumsq,
                      const uint8_t* __restrict__ row_mask, float* __restrict__ logits_local,
                      float* __restrict__ neg_map, float* __restrict__ pos_map, int B, int P, int K,
                      float logit_scale, float spatial_scale) {
  __shared__ float s_m[kAggWarps][kJ * 32];
  __shared__ float s_s[kAggWarps][kJ * 32];
  __shared__ float s_a[kAggWarps][kJ * 32];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool last_ok = lane + 32 * (kJ - 1) < K;        // only the last class slot of a lane can be out of range
  const float s2 = spatial_scale * kLog2e;
  float m[kJ], ssum[kJ], acc[kJ];
#pragma unroll
  for (int j = 0; j < kJ; ++j) {
    m[j] = -INFINITY;
    ssum[j] = 0.f;
    acc[j] = 0.f;
  }
  // running pointers of this warp's current row
  const int64_t row0 = static_cast<int64_t>(b) * P + warp;
  const float* dp = dots + row0 * ldn + lane;
  const float* sq = row_sumsq != nullptr ? row_sumsq + row0 : nullptr;
  const uint8_t* mk = kMask ? row_mask + row0 : nullptr;
  float* np = kMaps ? neg_map + (static_cast<int64_t>(warp) * B + b) * K + lane : nullptr;
  float* pp = kMaps ? pos_map + (static_cast<int64_t>(warp) * B + b) * K + lane : nullptr;
  const int64_t dstep = static_cast<int64_t>(kAggWarps) * ldn;
  const int64_t mstep = static_cast<int64_t>(kAggWarps) * B * K;

  float npos[kJ], nneg[kJ], nevi[kJ], nsq = 1.f;
  bool nmasked = false;
  auto fetch = [&]() {       // operands of the row the running pointers designate
    nmasked = kMask ? (__ldg(mk) != 0) : false;
    if (nmasked) return;
    if (sq != nullptr) nsq = __ldg(sq);
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      if (j < kJ - 1 || last_ok) {
        if (kMaps) npos[j] = __ldcs(dp + 32 * j);
        nneg[j] = __ldcs(dp + K + 32 * j);
        if (kEvi) nevi[j] = __ldcs(dp + 2 * K + 32 * j);
      } else {
        npos[j] = 0.f;
        nneg[j] = 0.f;
        nevi[j] = 0.f;
      }
    }
  };
  if (warp < P) fetch();
  for (int p = warp; p < P; p += kAggWarps) {
    float pos[kJ], neg[kJ], evi[kJ];
    const bool masked = nmasked;
    const float rn = sq != nullptr ? rsqrtf(nsq) : 1.0f;
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      if (kMaps) pos[j] = npos[j] * rn;
      neg[j] = nneg[j] * rn;
      if (kEvi) evi[j] = nevi[j] * rn;
    }
    float* np_cur = np;
    float* pp_cur = pp;
    dp += dstep;
    if (sq != nullptr) sq += kAggWarps;
    if (kMask) mk += kAggWarps;
    if (kMaps) {
      np += mstep;
      pp += mstep;
    }
    if (p + kAggWarps < P) fetch();                 // prefetch the next row of this warp
    if (kMask && masked) continue;                  // padded token: weight underflows to exactly 0 (T:491-498)
    float val[kJ], t2[kJ];                          // summand and log2-domain spatial score per class
    if (kEvi) {
      // winner-take-all softmax over the classes of this row
      float mx = neg[0];
#pragma unroll
      for (int j = 1; j < kJ - 1; ++j) mx = fmaxf(mx, neg[j]);
      if (kJ > 1) mx = fmaxf(mx, last_ok ? neg[kJ - 1] : -INFINITY);
      else if (!last_ok) mx = -INFINITY;
      mx = warp_max_redux(mx);
      const float g2 = s2 * (mx + 1.0f);            // gain * log2(e); gain >= 0 because the scores are cosines
      float zmax2 = g2 * mx;
      if (g2 < 0.f) {                               // not cosines (max < -1): the largest gain*neg is at the MIN
        float mn = -neg[0];                         // (warp-uniform branch, never taken on unit features)
#pragma unroll
        for (int j = 1; j < kJ - 1; ++j) mn = fmaxf(mn, -neg[j]);
        if (kJ > 1) mn = fmaxf(mn, last_ok ? -neg[kJ - 1] : -INFINITY);
        else if (!last_ok) mn = -INFINITY;
        zmax2 = -g2 * warp_max_redux(mn);
      }
      float z[kJ], den = 0.f;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        z[j] = ex2f(fminf(fmaf(g2, neg[j], -zmax2), 0.f));       // clamp: exact for g2 >= 0, keeps g2 < 0 finite
        if (j == kJ - 1 && !last_ok) z[j] = 0.f;
        den += z[j];
      }
      den = warp_sum(den);
      const float inv = __fdividef(1.0f, den);
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        val[j] = neg[j] * (z[j] * inv);
        t2[j] = s2 * evi[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        val[j] = neg[j];
        t2[j] = s2 * neg[j];
      }
    }
    if (kMaps) {
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        if (j < kJ - 1 || last_ok) {
          __stcs(np_cur + 32 * j, val[j]);
          __stcs(pp_cur + 32 * j, pos[j]);
        }
      }
    }
    // online spatial softmax: d = t - m; one of exp(m - m'), exp(t - m') is 1, the other exp(-|d|)
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      const float d = t2[j] - m[j];
      const float x = ex2f(-fabsf(d));              // first row: m = -inf -> d = +inf -> x = 0
      const bool up = d > 0.f;
      const float sc = up ? x : 1.0f;
      const float e = up ? 1.0f : x;
      m[j] = fmaxf(m[j], t2[j]);
      ssum[j] = fmaf(ssum[j], sc, e);
      acc[j] = fmaf(acc[j], sc, e * val[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < kJ; ++j) {
    s_m[warp][lane + 32 * j] = m[j];
    s_s[warp][lane + 32 * j] = ssum[j];
    s_a[warp][lane + 32 * j] = acc[j];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kAggWarps; ++w) M = fmaxf(M, s_m[w][k]);
    float S = 0.f, A = 0.f;
#pragma unroll
    for (int w = 0; w < kAggWarps; ++w) {
      const float sc = (s_m[w][k] == -INFINITY) ? 0.f : ex2f(s_m[w][k] - M);
      S += s_s[w][k] * sc;
      A += s_a[w][k] * sc;
    }
    logits_local[static_cast<int64_t>(b) * K + k] = logit_scale * A / S;
  }
}

// logits_[b,k] = scale * sum_d x[b,d] * T[k,d],  x = g_unit (or 0.5*(g_unit + g_add): T:448).  One CTA
// per image; warp per class; fp32 throughout (tiny: B*K*D MACs).
__global__ void __launch_bounds__(256)
global_logits_kernel(const float* __restrict__ g_unit, const float* __restrict__ g_add, const float* __restrict__ tpos,
                     float* __restrict__ out, int D, int K, float scale) {
  extern __shared__ float sx[];
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float v = g_unit[static_cast<int64_t>(b) * D + d];
    if (g_add != nullptr) v = 0.5f * (v + g_add[static_cast<int64_t>(b) * D + d]);
    sx[d] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < K; k += blockDim.x / 32) {
    const float* tp = tpos + static_cast<int64_t>(k) * D;
    float s = 0.f;
    for (int d = lane * 4; d < D; d += 128) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(tp + d));
      s += sx[d] * t.x + sx[d + 1] * t.y + sx[d + 2] * t.z + sx[d + 3] * t.w;
    }
    s = warp_sum(s);
    if (lane == 0) out[static_cast<int64_t>(b) * K + k] = scale * s;
  }
}

}  // namespace lecb

using namespace lecb;

extern "C" int lecb_head_aggregate(const float* dots, int ldn, const float* row_sumsq, const uint8_t* row_mask,
                                   float* logits_local, float* neg_map, float* pos_map, int B, int P, int K,
                                   int n_txt, float logit_scale, float spatial_scale, void* stream) {
  LECB_CHECK_ARG(dots && logits_local, "lecb_head_aggregate: null pointer");
  LECB_CHECK_ARG((neg_map == nullptr) == (pos_map == nullptr), "lecb_head_aggregate: neg_map and pos_map go together");
  LECB_CHECK_ARG(B > 0 && P > 0 && K > 0 && K <= 128, "lecb_head_aggregate: need 0 < K <= 128 (K=%d)", K);
  LECB_CHECK_ARG(n_txt == 2 || n_txt == 3, "lecb_head_aggregate: n_txt must be 2 (pos,neg) or 3 (+evidence)");
  LECB_CHECK_ARG(ldn >= n_txt * K, "lecb_head_aggregate: ldn=%d < n_txt*K", ldn);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int kj = (K + 31) / 32;
  const bool evi = n_txt >= 3, maps = neg_map != nullptr, mask = row_mask != nullptr;
#define LECB_AGG4(J, E, M, K_)                                                                                     \
  head_aggregate_kernel<J, E, M, K_><<<B, kAggWarps * 32, 0, s>>>(dots, ldn, row_sumsq, row_mask, logits_local,    \
                                                                  neg_map, pos_map, B, P, K, logit_scale,         \
                                                                  spatial_scale)
#define LECB_AGG(J)                                                        \
  do {                                                                     \
    if (evi) {                                                             \
      if (maps) { if (mask) LECB_AGG4(J, true, true, true); else LECB_AGG4(J, true, true, false); }       \
      else      { if (mask) LECB_AGG4(J, true, false, true); else LECB_AGG4(J, true, false, false); }     \
    } else {                                                               \
      if (maps) { if (mask) LECB_AGG4(J, false, true, true); else LECB_AGG4(J, false, true, false); }     \
      else      { if (mask) LECB_AGG4(J, false, false, true); else LECB_AGG4(J, false, false, false); }   \
    }                                                                      \
  } while (0)
  if (kj == 1) LECB_AGG(1);
  else if (kj == 2) LECB_AGG(2);
  else if (kj == 3) LECB_AGG(3);
  else LECB_AGG(4);
#undef LECB_AGG
#undef LECB_AGG4
  count_launch();
  return check_launch("head_aggregate_kernel");
}

extern "C" int lecb_global_logits(const float* g_unit, const float* g_add, const float* tpos, float* out, int B, int D,
                                  int K, float scale, void* stream) {
  LECB_CHECK_ARG(g_unit && tpos && out, "lecb_global_logits: null pointer");
  LECB_CHECK_ARG(B > 0 && K > 0 && D > 0 && D % 4 == 0 && D <= 8192, "lecb_global_logits: need D %% 4 == 0, D <= 8192");
  global_logits_kernel<<<B, 256, D * sizeof(float), static_cast<cudaStream_t>(stream)>>>(g_unit, g_add, tpos, out, D,
                                                                                          K, scale);
  count_launch();
  return check_launch("global_logits_kernel");
}
