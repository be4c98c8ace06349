// Test-time score fusion on either side of the scoring path:
//  * block_fuse: aggregation of the per-window scores of one image — s_ag = max_n d if max_n d > thr else min_n d —
//    over optionally re-weighted scores d (Caption_distill_double.py:655-662 inline in `test`; gen_final_ans.py:18-71
//    `fuse` / `fuse6`: windows are re-weighted by 1 + mean retrieval similarity and by 1 + the unbiased variance of
//    their class scores), fused with the final `weight * s_ag + base` (T:662 / gen_final_ans.py);
//  * cooc_adjust: pred + w * pred @ P with the row-normalised class co-occurrence matrix (T:611-618).
// One CTA per image; a warp owns a window (row variance by shuffles), classes are reduced over windows in smem.
#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

constexpr int kFuseWarps = 8;
constexpr int kFuseMaxK = 128;

__global__ void __launch_bounds__(kFuseWarps * 32)
block_fuse_kernel(const float* __restrict__ data, const float* __restrict__ sims, int sims_ld, const float* __restrict__ base,
                  float* __restrict__ out, int NB, int K, int mode, float threshold, float weight) {
  pdl_grid_sync();
  __shared__ float s_max[kFuseWarps][kFuseMaxK];
  __shared__ float s_min[kFuseWarps][kFuseMaxK];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float vmax[kFuseMaxK / 32], vmin[kFuseMaxK / 32];
#pragma unroll
  for (int j = 0; j < kFuseMaxK / 32; ++j) {
    vmax[j] = -INFINITY;
    vmin[j] = INFINITY;
  }
  for (int n = warp; n < NB; n += kFuseWarps) {
    const float* row = data + (static_cast<int64_t>(b) * NB + n) * K;
    float d[kFuseMaxK / 32];
#pragma unroll
    for (int j = 0; j < kFuseMaxK / 32; ++j) d[j] = (lane + 32 * j < K) ? __ldg(row + lane + 32 * j) : 0.f;
    if (mode != 0) {
      // mean retrieval similarity of this window (gen_final_ans.py:20, 41: sims_scores.mean(-1))
      float sm = 0.f;
      for (int i = lane; i < sims_ld; i += 32) sm += __ldg(sims + (static_cast<int64_t>(b) * NB + n) * sims_ld + i);
      sm = warp_sum(sm) / static_cast<float>(sims_ld);
      auto row_var = [&](const float (&v)[kFuseMaxK / 32]) {       // torch.var(dim=2): unbiased
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kFuseMaxK / 32; ++j) s += (lane + 32 * j < K) ? v[j] : 0.f;
        const float mean = warp_sum(s) / static_cast<float>(K);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < kFuseMaxK / 32; ++j) {
          const float c = (lane + 32 * j < K) ? v[j] - mean : 0.f;
          q += c * c;
        }
        return warp_sum(q) / static_cast<float>(K - 1);
      };
      float scale;
      if (mode == 1) {                    // fuse: d1 = (1+sim) d; d2 = (1 + var(d1)) d1
        const float a = 1.0f + sm;
#pragma unroll
        for (int j = 0; j < kFuseMaxK / 32; ++j) d[j] *= a;
        scale = 1.0f + row_var(d);
      } else {                            // fuse6: (1 + var(d)) (1 + var((1+sim) d)) (1+sim) d
        const float v0 = row_var(d);
        const float a = 1.0f + sm;
#pragma unroll
        for (int j = 0; j < kFuseMaxK / 32; ++j) d[j] *= a;
        scale = (1.0f + v0) * (1.0f + row_var(d));
      }
#pragma unroll
      for (int j = 0; j < kFuseMaxK / 32; ++j) d[j] *= scale;
    }
#pragma unroll
    for (int j = 0; j < kFuseMaxK / 32; ++j) {
      vmax[j] = fmaxf(vmax[j], d[j]);
      vmin[j] = fminf(vmin[j], d[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < kFuseMaxK / 32; ++j) {
    s_max[warp][lane + 32 * j] = vmax[j];
    s_min[warp][lane + 32 * j] = vmin[j];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float a = -INFINITY, c = INFINITY;
#pragma unroll
    for (int w = 0; w < kFuseWarps; ++w) {
      a = fmaxf(a, s_max[w][k]);
      c = fminf(c, s_min[w][k]);
    }
    const float s_ag = a > threshold ? a : c;
    const float bs = base != nullptr ? base[static_cast<int64_t>(b) * K + k] : 0.f;
    out[static_cast<int64_t>(b) * K + k] = weight * s_ag + bs;
  }
}

// out[b,:] = pred[b,:] + w * pred[b,:] @ P   (P [K,K] row-major).  One CTA per image, pred row in smem.
__global__ void __launch_bounds__(128)
cooc_adjust_kernel(const float* __restrict__ pred, const float* __restrict__ P, float* __restrict__ out, int K, float w) {
  pdl_grid_sync();
  __shared__ float sp[kFuseMaxK];
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += blockDim.x) sp[k] = pred[static_cast<int64_t>(b) * K + k];
  __syncthreads();
  for (int j = threadIdx.x; j < K; j += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(sp[k], __ldg(P + static_cast<int64_t>(k) * K + j), acc);
    out[static_cast<int64_t>(b) * K + j] = sp[j] + w * acc;
  }
}

}  // namespace lecb

using namespace lecb;

extern "C" int lecb_block_fuse(const float* data, const float* sims, int sims_ld, const float* base, float* out, int B,
                               int NB, int K, int mode, float threshold, float weight, void* stream) {
  LECB_CHECK_ARG(data && out, "lecb_block_fuse: null pointer");
  LECB_CHECK_ARG(B > 0 && NB > 0 && K > 1 && K <= kFuseMaxK, "lecb_block_fuse: need B, NB > 0 and 1 < K <= 128 (K=%d)", K);
  LECB_CHECK_ARG(mode >= 0 && mode <= 2, "lecb_block_fuse: mode must be 0 (plain), 1 (fuse) or 2 (fuse6)");
  LECB_CHECK_ARG(mode == 0 || (sims != nullptr && sims_ld > 0), "lecb_block_fuse: modes 1 and 2 need the similarity scores");
  launch_k(block_fuse_kernel, dim3(B), dim3(kFuseWarps * 32), 0, static_cast<cudaStream_t>(stream), data, sims, sims_ld, base, out, NB, K, mode,
                                                                                threshold, weight);
  count_launch();
  return check_launch("block_fuse_kernel");
}

extern "C" int lecb_cooc_adjust(const float* pred, const float* P, float* out, int B, int K, float weight, void* stream) {
  LECB_CHECK_ARG(pred && P && out, "lecb_cooc_adjust: null pointer");
  LECB_CHECK_ARG(B > 0 && K > 0 && K <= kFuseMaxK, "lecb_cooc_adjust: need 0 < K <= 128 (K=%d)", K);
  launch_k(cooc_adjust_kernel, dim3(B), dim3(128), 0, static_cast<cudaStream_t>(stream), pred, P, out, K, weight);
  count_launch();
  return check_launch("cooc_adjust_kernel");
}
