// Row kernels of the ViT visual tower (VisionTransformer.forward, M:259-276): patch extraction for the
// patch-embedding GEMM, class-token / positional-embedding / ln_pre fusion, and a strided column copy used
// by the dense last block.  All HBM-bound: 128-bit accesses, one warp per token row where a row reduction
// is needed.
#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

// x NCHW fp32 [B,3,H,W] -> out bf16 [B*gh*gw, Kpad]; column k = c*p*p + py*p + px (the flattening of
// conv1.weight [width,3,p,p], M:247), columns >= 3*p*p are zero.  One thread = 8 consecutive columns.
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int H, int W, int p, int Kpad) {
  pdl_grid_sync();
  const int gw = W / p, gh = H / p;
  const int octs = Kpad / 8;
  const int64_t total = static_cast<int64_t>(B) * gh * gw * octs;
  const int K = 3 * p * p;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int oc = static_cast<int>(idx % octs);
    const int64_t row = idx / octs;
    const int gx = static_cast<int>(row % gw);
    const int gy = static_cast<int>((row / gw) % gh);
    const int b = static_cast<int>(row / (static_cast<int64_t>(gw) * gh));
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = oc * 8 + i;
      if (k < K) {
        const int c = k / (p * p);
        const int r = k - c * p * p;
        const int py = r / p, px = r - py * p;
        v[i] = __ldg(x + ((static_cast<int64_t>(b) * 3 + c) * H + gy * p + py) * W + gx * p + px);
      } else {
        v[i] = 0.f;
      }
    }
    uint4 u;
    u.x = pack_bf16(v[0], v[1]);
    u.y = pack_bf16(v[2], v[3]);
    u.z = pack_bf16(v[4], v[5]);
    u.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + row * Kpad + oc * 8) = u;
  }
}

// token row (b,t): v = (t == 0 ? class_embedding : emb[b*(T-1) + t-1]) + positional_embedding[t]; out = ln_pre(v)
// (M:262-266).  One warp per row; kVec float4 per lane.
template <int kVec>
__global__ void __launch_bounds__(256)
vit_embed_ln_kernel(const __nv_bfloat16* __restrict__ emb, const float* __restrict__ cls, const float* __restrict__ pos,
                    const float* __restrict__ g, const float* __restrict__ bta, float* __restrict__ out, int64_t rows,
                    int T, int D, float eps) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int nvec = D / 4;
  for (int64_t row = warp_global; row < rows; row += nwarps) {
    const int t = static_cast<int>(row % T);
    const int64_t b = row / T;
    const float4* pp = reinterpret_cast<const float4*>(pos + static_cast<int64_t>(t) * D);
    float4 v[kVec];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        float4 a;
        if (t == 0) {
          a = __ldg(reinterpret_cast<const float4*>(cls) + vi);
        } else {
          const uint2 u = __ldg(reinterpret_cast<const uint2*>(emb + (b * (T - 1) + t - 1) * D) + vi);
          const float2 lo = unpack_bf16(u.x), hi = unpack_bf16(u.y);
          a = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
        const float4 q = __ldg(pp + vi);
        v[i] = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
        s += v[i].x + v[i].y + v[i].z + v[i].w;
      }
    }
    const float mean = warp_sum(s) / static_cast<float>(D);
    float q2 = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        const float a = v[i].x - mean, b2 = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q2 += a * a + b2 * b2 + c * c + d * d;
      }
    }
    const float rstd = rsqrtf(warp_sum(q2) / static_cast<float>(D) + eps);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + vi);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bta) + vi);
        reinterpret_cast<float4*>(out + row * D)[vi] =
            make_float4((v[i].x - mean) * rstd * gg.x + bb.x, (v[i].y - mean) * rstd * gg.y + bb.y,
                        (v[i].z - mean) * rstd * gg.z + bb.z, (v[i].w - mean) * rstd * gg.w + bb.w);
      }
    }
  }
}

// dst[r, 0:cols] = src[r, col0:col0+cols]  (bf16, 16-byte vectors)
__global__ void __launch_bounds__(256)
copy_cols_kernel(const __nv_bfloat16* __restrict__ src, int64_t ld_src, int col0, __nv_bfloat16* __restrict__ dst,
                 int64_t ld_dst, int64_t rows, int cols) {
  pdl_grid_sync();
  const int vecs = cols / 8;
  const int64_t total = rows * vecs;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = idx / vecs;
    const int v = static_cast<int>(idx - r * vecs);
    *reinterpret_cast<uint4*>(dst + r * ld_dst + v * 8) = __ldg(reinterpret_cast<const uint4*>(src + r * ld_src + col0 + v * 8));
  }
}

static unsigned grid_for(int64_t total, int block, int per_sm) {
  const int64_t want = (total + block - 1) / block;
  const int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
  return static_cast<unsigned>(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace lecb

using namespace lecb;

extern "C" int lecb_patchify(const float* x, void* out, int B, int H, int W, int patch, int Kpad, void* stream) {
  LECB_CHECK_ARG(x && out, "lecb_patchify: null pointer");
  LECB_CHECK_ARG(B > 0 && patch > 0 && H % patch == 0 && W % patch == 0, "lecb_patchify: H=%d W=%d not multiples of patch=%d", H, W, patch);
  LECB_CHECK_ARG(Kpad % 8 == 0 && Kpad >= 3 * patch * patch, "lecb_patchify: Kpad=%d must be a multiple of 8 and >= 3*patch^2", Kpad);
  const int64_t total = static_cast<int64_t>(B) * (H / patch) * (W / patch) * (Kpad / 8);
  launch_k(patchify_kernel, dim3(grid_for(total, 256, 16)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      x, static_cast<__nv_bfloat16*>(out), B, H, W, patch, Kpad);
  count_launch();
  return check_launch("patchify_kernel");
}

extern "C" int lecb_vit_embed_ln(const void* emb, const float* cls, const float* pos, const float* gamma,
                                 const float* beta, float* out, int B, int T, int D, float eps, void* stream) {
  LECB_CHECK_ARG(emb && cls && pos && gamma && beta && out, "lecb_vit_embed_ln: null pointer");
  LECB_CHECK_ARG(B > 0 && T > 1 && D > 0 && D % 4 == 0, "lecb_vit_embed_ln: bad shape B=%d T=%d D=%d", B, T, D);
  const int64_t rows = static_cast<int64_t>(B) * T;
  const int per_lane = (D / 4 + 31) / 32;
  const unsigned grid = grid_for(rows * 32, 256, 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* e = static_cast<const __nv_bfloat16*>(emb);
  if (per_lane <= 2) launch_k(vit_embed_ln_kernel<2>, dim3(grid), dim3(256), 0, s, e, cls, pos, gamma, beta, out, rows, T, D, eps);
  else if (per_lane <= 4) launch_k(vit_embed_ln_kernel<4>, dim3(grid), dim3(256), 0, s, e, cls, pos, gamma, beta, out, rows, T, D, eps);
  else if (per_lane <= 8) launch_k(vit_embed_ln_kernel<8>, dim3(grid), dim3(256), 0, s, e, cls, pos, gamma, beta, out, rows, T, D, eps);
  else return fail(LECB_ERR_UNSUPPORTED, "lecb_vit_embed_ln: D=%d too wide", D);
  count_launch();
  return check_launch("vit_embed_ln_kernel");
}

extern "C" int lecb_copy_cols(const void* src, int64_t ld_src, int col0, void* dst, int64_t ld_dst, int64_t rows,
                              int cols, void* stream) {
  LECB_CHECK_ARG(src && dst, "lecb_copy_cols: null pointer");
  LECB_CHECK_ARG(rows > 0 && cols > 0 && cols % 8 == 0 && col0 % 8 == 0 && ld_src % 8 == 0 && ld_dst % 8 == 0,
                 "lecb_copy_cols: cols, col0 and leading dimensions must be multiples of 8");
  launch_k(copy_cols_kernel, dim3(grid_for(rows * (cols / 8), 256, 16)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(src), ld_src, col0, static_cast<__nv_bfloat16*>(dst), ld_dst, rows, cols);
  count_launch();
  return check_launch("copy_cols_kernel");
}
