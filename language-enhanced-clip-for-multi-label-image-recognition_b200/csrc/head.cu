// Dual-prompt head: winner-take-all + spatial-softmax aggregation of the patch-by-class similarity
// maps (T:456-470 test / T:496-514 train) and the global logits (T:453-455).  The patch-by-class
// contraction itself runs on tcgen05 (lecb_gemm_bf16 against the concatenated [pos;neg;evi] prompt
// matrix, fp32 output); these kernels consume its raw dot products with the per-row inverse norms, so
// the L2 normalisation of the local features (T:442 / T:486) never touches HBM as a separate pass.
#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

constexpr int kAggWarps = 8;

// One CTA per image / caption.  Warp w streams rows p = w, w+8, ...; lane l owns classes l, l+32, ...
// Row-wise WTA softmax over classes via warp shuffles; column-wise spatial softmax over rows as an
// online (max, sum, weighted-sum) recurrence per (lane, class), merged across the 8 warps at the end.
template <int kJ>
__global__ void __launch_bounds__(kAggWarps * 32)
head_aggregate_kernel(const float* __restrict__ dots, int ldn, const float* __restrict__ row_sumsq,
                      const uint8_t* __restrict__ row_mask, float* __restrict__ logits_local,
                      float* __restrict__ neg_map, float* __restrict__ pos_map, int B, int P, int K, int n_txt,
                      float logit_scale, float spatial_scale) {
  __shared__ float s_m[kAggWarps][kJ * 32];
  __shared__ float s_s[kAggWarps][kJ * 32];
  __shared__ float s_a[kAggWarps][kJ * 32];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool evidence = n_txt >= 3;
  float m[kJ], ssum[kJ], acc[kJ];
#pragma unroll
  for (int j = 0; j < kJ; ++j) {
    m[j] = -INFINITY;
    ssum[j] = 0.f;
    acc[j] = 0.f;
  }
  // Software-pipelined row loop: the raw dot products of the NEXT row of this warp are requested before the
  // shuffle / exp chain of the current row starts, so the DRAM latency hides under the reductions.
  auto live_row = [&](int p) { return p < P && !(row_mask != nullptr && row_mask[static_cast<int64_t>(b) * P + p]); };
  auto fetch = [&](int p, float (&rp)[kJ], float (&rneg)[kJ], float (&re)[kJ], float& rn) {
    const int64_t row = static_cast<int64_t>(b) * P + p;
    const float* dp = dots + row * ldn;
    rn = row_sumsq != nullptr ? __ldg(row_sumsq + row) : 1.0f;
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      const int k = lane + 32 * j;
      if (k < K) {
        rp[j] = __ldcs(dp + k);
        rneg[j] = __ldcs(dp + K + k);
        re[j] = evidence ? __ldcs(dp + 2 * K + k) : 0.f;
      } else {
        rp[j] = rneg[j] = re[j] = 0.f;
      }
    }
  };
  int p = warp;
  while (p < P && !live_row(p)) p += kAggWarps;       // padded token: weight underflows to exactly 0 (T:491-498)
  float npos[kJ], nneg[kJ], nevi[kJ], nrn = 1.f;
  if (p < P) fetch(p, npos, nneg, nevi, nrn);
  while (p < P) {
    float pos[kJ], neg[kJ], evi[kJ];
    const float rn = row_sumsq != nullptr ? rsqrtf(nrn) : 1.0f;
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      pos[j] = npos[j] * rn;
      neg[j] = nneg[j] * rn;
      evi[j] = nevi[j] * rn;
      if (lane + 32 * j < K) mx = fmaxf(mx, neg[j]);
    }
    const int pcur = p;
    p += kAggWarps;
    while (p < P && !live_row(p)) p += kAggWarps;
    if (p < P) fetch(p, npos, nneg, nevi, nrn);          // prefetch the next live row
    float t[kJ];
    if (evidence) {
      mx = warp_max(mx);
      const float gain = spatial_scale * (mx + 1.0f);
      float z[kJ], zmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        z[j] = (lane + 32 * j < K) ? gain * neg[j] : -INFINITY;
        zmax = fmaxf(zmax, z[j]);
      }
      zmax = warp_max(zmax);
      float den = 0.f;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        z[j] = (lane + 32 * j < K) ? __expf(z[j] - zmax) : 0.f;
        den += z[j];
      }
      den = warp_sum(den);
      const float inv = __fdividef(1.0f, den);
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        neg[j] *= z[j] * inv;
        t[j] = spatial_scale * evi[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < kJ; ++j) t[j] = spatial_scale * neg[j];
    }
    if (neg_map != nullptr) {
      float* np = neg_map + (static_cast<int64_t>(pcur) * B + b) * K;
      float* pp = pos_map + (static_cast<int64_t>(pcur) * B + b) * K;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        const int k = lane + 32 * j;
        if (k < K) {
          __stcs(np + k, neg[j]);
          __stcs(pp + k, pos[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      if (t[j] > m[j]) {
        const float sc = __expf(m[j] - t[j]);     // exp(-inf) = 0 on the first row
        ssum[j] = ssum[j] * sc + 1.0f;
        acc[j] = acc[j] * sc + neg[j];
        m[j] = t[j];
      } else {
        const float e = __expf(t[j] - m[j]);
        ssum[j] += e;
        acc[j] += e * neg[j];
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
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
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
#define LECB_AGG(J)                                                                                                 \
  head_aggregate_kernel<J><<<B, kAggWarps * 32, 0, s>>>(dots, ldn, row_sumsq, row_mask, logits_local, neg_map,      \
                                                        pos_map, B, P, K, n_txt, logit_scale, spatial_scale)
  if (kj == 1) LECB_AGG(1);
  else if (kj == 2) LECB_AGG(2);
  else if (kj == 3) LECB_AGG(3);
  else LECB_AGG(4);
#undef LECB_AGG
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
