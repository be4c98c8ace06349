// Dual-prompt head: winner-take-all + spatial-softmax aggregation of the patch-by-class similarity
// maps (T:456-470 test / T:496-514 train) and the global logits (T:453-455).  The patch-by-class
// contraction itself runs on tcgen05 (lecb_gemm_bf16 against the concatenated [pos;neg;evi] prompt
// matrix, fp32 output); these kernels consume its raw dot products with the per-row inverse norms, so
// the L2 normalisation of the local features (T:442 / T:486) never touches HBM as a separate pass.
#include "lecb_common.cuh"
#include "lecb_host.h"

#include <type_traits>

namespace lecb {

constexpr int kAggWarps = 8;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float rsqrt_fast(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// warp max of floats in ONE instruction (CREDUX.MAX.F32, sm_100a)
__device__ __forceinline__ float warp_max_f32(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

template <int kJ, bool kEvi, bool kMaps>
struct AggRow {                // raw operands of one patch row, as loaded
  float pos[kJ], neg[kJ], evi[kJ];
  float sq;
  bool masked;
};

// One CTA per image / caption.  Warp w streams rows p = w, w+8, ...; lane l owns classes l, l+32, ...
// The kernel is instruction-issue bound long before it is HBM bound (ncu: 91 % issue slots busy in round 1's first version;
// round 2's still issued ~150 instructions per row, which is what held the variant without map outputs at 0.66 of the HBM peak),
// so everything per row is kept to the minimum:
//  * compile-time variants for evidence / maps / mask;
//  * the row's inverse norm is never multiplied into the operands: it is folded into the three scalars that use it
//    (class max, softmax gain, spatial scale) and into the softmax normaliser;
//  * class max = one CREDUX.MAX.F32 (gain >= 0 for cosines, so max_k gain*neg_k = gain*max_k neg_k); class sum = one integer
//    REDUX.ADD over 2^-24 fixed point (the largest term is exactly 1, so the sum is in [1, K]: relative error < 1e-6);
//  * spatial softmax with the FIXED reference 0 instead of a running maximum: the scores are spatial_scale * cosine, so
//    exp2(t) stays inside fp32 range (|t| <= 72 at the reference's scale 50) and the update is ex2 + add + fma.  Inputs
//    for which that would over- or underflow (not cosines, or a much larger scale) are detected on the RESULT (sum outside
//    [2^-100, 2^100], or non-finite) and the CTA reruns the rows with the online (running-maximum) update: same result for every
//    input, the common case pays nothing;
//  * three operand sets per warp in rotation (rows p, p+8, p+16): two rows of loads are always in flight behind the
//    row being reduced.
template <int kJ, bool kEvi, bool kMaps, bool kMask>
__global__ void __launch_bounds__(kAggWarps * 32)
head_aggregate_kernel(const float* __restrict__ dots, int ldn, const float* __restrict__ row_sumsq,
                      const uint8_t* __restrict__ row_mask, float* __restrict__ logits_local,
                      float* __restrict__ neg_map, float* __restrict__ pos_map, int B, int P, int K,
                      float logit_scale, float spatial_scale) {
  pdl_grid_sync();
  __shared__ float s_m[kAggWarps][kJ * 32];
  __shared__ float s_s[kAggWarps][kJ * 32];
  __shared__ float s_a[kAggWarps][kJ * 32];
  using Row = AggRow<kJ, kEvi, kMaps>;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool last_ok = lane + 32 * (kJ - 1) < K;        // only the last class slot of a lane can be out of range
  const float s2 = spatial_scale * kLog2e;
  const float* dots_b = dots + static_cast<int64_t>(b) * P * ldn + lane;
  const float* sq_b = row_sumsq != nullptr ? row_sumsq + static_cast<int64_t>(b) * P : nullptr;
  const uint8_t* mk_b = kMask ? row_mask + static_cast<int64_t>(b) * P : nullptr;
  const int64_t map_b = static_cast<int64_t>(b) * K + lane;
  const int64_t map_step = static_cast<int64_t>(B) * K;
  float m[kJ], ssum[kJ], acc[kJ];

  auto fetch = [&](Row& r, int p) {
    r.masked = kMask ? (__ldg(mk_b + p) != 0) : false;
    r.sq = 1.f;
    if (r.masked) return;
    if (sq_b != nullptr) r.sq = __ldg(sq_b + p);
    const float* dp = dots_b + static_cast<int64_t>(p) * ldn;
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      if (j < kJ - 1 || last_ok) {
        if (kMaps) r.pos[j] = __ldcs(dp + 32 * j);
        r.neg[j] = __ldcs(dp + K + 32 * j);
        if (kEvi) r.evi[j] = __ldcs(dp + 2 * K + 32 * j);
      } else {
        r.pos[j] = 0.f;
        r.neg[j] = 0.f;
        r.evi[j] = 0.f;
      }
    }
  };

  // kOnline = false: fixed reference 0 (m[] unused); true: running maximum (one of exp(m-m'), exp(t-m') is 1)
  auto row = [&](const Row& r, int p, auto online_tag) {
    constexpr bool kOnline = decltype(online_tag)::value;
    if (kMask && r.masked) return;                  // padded token: weight underflows to exactly 0 (T:491-498)
    const float rn = sq_b != nullptr ? rsqrt_fast(r.sq) : 1.0f;
    const float s2rn = s2 * rn;
    float val[kJ], t2[kJ];                          // summand and log2-domain spatial score per class
    if (kEvi) {
      // winner-take-all softmax over the classes of this row
      float mx = r.neg[0];
#pragma unroll
      for (int j = 1; j < kJ - 1; ++j) mx = fmaxf(mx, r.neg[j]);
      if (kJ > 1) mx = fmaxf(mx, last_ok ? r.neg[kJ - 1] : -INFINITY);
      else if (!last_ok) mx = -INFINITY;
      mx = warp_max_f32(mx) * rn;
      const float g2 = s2 * (mx + 1.0f);            // gain * log2(e); gain >= 0 because the scores are cosines
      float zmax2 = g2 * mx;
      if (g2 < 0.f) {                               // not cosines (max < -1): the largest gain*neg is at the MIN
        float mn = -r.neg[0];                       // (warp-uniform branch, never taken on unit features)
#pragma unroll
        for (int j = 1; j < kJ - 1; ++j) mn = fmaxf(mn, -r.neg[j]);
        if (kJ > 1) mn = fmaxf(mn, last_ok ? -r.neg[kJ - 1] : -INFINITY);
        else if (!last_ok) mn = -INFINITY;
        zmax2 = -g2 * (warp_max_f32(mn) * rn);
      }
      const float g2r = g2 * rn;
      float z[kJ], den = 0.f;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        z[j] = ex2f(fminf(fmaf(g2r, r.neg[j], -zmax2), 0.f));    // clamp: exact for g2 >= 0, keeps g2 < 0 finite
        if (j == kJ - 1 && !last_ok) z[j] = 0.f;
        den += z[j];
      }
      // sum over the warp in 2^-24 fixed point: every z is in [0, 1] and the largest is exactly 1
      const uint32_t den_i = __reduce_add_sync(0xffffffffu, __float2uint_rn(den * 16777216.0f));
      const float inv = rcp_fast(static_cast<float>(den_i)) * (16777216.0f * rn);
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        val[j] = r.neg[j] * (z[j] * inv);
        t2[j] = s2rn * r.evi[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        val[j] = r.neg[j] * rn;
        t2[j] = s2rn * r.neg[j];
      }
    }
    if (kMaps) {
      float* np_cur = neg_map + map_b + static_cast<int64_t>(p) * map_step;
      float* pp_cur = pos_map + map_b + static_cast<int64_t>(p) * map_step;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        if (j < kJ - 1 || last_ok) {
          __stcs(np_cur + 32 * j, val[j]);
          __stcs(pp_cur + 32 * j, r.pos[j] * rn);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      if (kOnline) {
        const float d = t2[j] - m[j];
        const float x = ex2f(-fabsf(d));            // first row: m = -inf -> d = +inf -> x = 0
        const bool up = d > 0.f;
        const float sc = up ? x : 1.0f;
        const float e = up ? 1.0f : x;
        m[j] = fmaxf(m[j], t2[j]);
        ssum[j] = fmaf(ssum[j], sc, e);
        acc[j] = fmaf(acc[j], sc, e * val[j]);
      } else {
        const float e = ex2f(t2[j]);
        ssum[j] += e;
        acc[j] = fmaf(e, val[j], acc[j]);
      }
    }
  };

  auto sweep = [&](auto online_tag) {
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      m[j] = -INFINITY;
      ssum[j] = 0.f;
      acc[j] = 0.f;
    }
    Row ra, rb, rc;
    if (warp < P) fetch(ra, warp);
    if (warp + kAggWarps < P) fetch(rb, warp + kAggWarps);
    if (warp + 2 * kAggWarps < P) fetch(rc, warp + 2 * kAggWarps);
    for (int p = warp; p < P; p += 3 * kAggWarps) {
      row(ra, p, online_tag);
      if (p + 3 * kAggWarps < P) fetch(ra, p + 3 * kAggWarps);
      if (p + kAggWarps < P) {
        row(rb, p + kAggWarps, online_tag);
        if (p + 4 * kAggWarps < P) fetch(rb, p + 4 * kAggWarps);
      }
      if (p + 2 * kAggWarps < P) {
        row(rc, p + 2 * kAggWarps, online_tag);
        if (p + 5 * kAggWarps < P) fetch(rc, p + 5 * kAggWarps);
      }
    }
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      s_m[warp][lane + 32 * j] = m[j];
      s_s[warp][lane + 32 * j] = ssum[j];
      s_a[warp][lane + 32 * j] = acc[j];
    }
  };

  // fast sweep: plain sums over the warps; the result tells whether the fixed reference was safe
  sweep(std::false_type{});
  __syncthreads();
  bool bad = false;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float S = 0.f, A = 0.f;
#pragma unroll
    for (int w = 0; w < kAggWarps; ++w) {
      S += s_s[w][k];
      A += s_a[w][k];
    }
    // 2^-100 < S < 2^100 and |A| finite: no term overflowed and the largest term kept full precision
    if (S > 7.888609e-31f && S < 1.2676506e30f && fabsf(A) < 1e37f) logits_local[static_cast<int64_t>(b) * K + k] = logit_scale * A / S;
    else bad = true;
  }
  if (!__syncthreads_or(bad ? 1 : 0)) return;

  sweep(std::true_type{});
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
  pdl_grid_sync();
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
  launch_k(head_aggregate_kernel<J, E, M, K_>, dim3(B), dim3(kAggWarps * 32), 0, s, dots, ldn, row_sumsq, row_mask, logits_local,    \
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
  launch_k(global_logits_kernel, dim3(B), dim3(256), D * sizeof(float), static_cast<cudaStream_t>(stream), g_unit, g_add, tpos, out, D,
                                                                                          K, scale);
  count_launch();
  return check_launch("global_logits_kernel");
}
