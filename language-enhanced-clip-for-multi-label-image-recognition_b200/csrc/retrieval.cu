// Caption retrieval (T:444-448): sim = g_unit · bankᵀ, top-10 per image, mean of the selected bank rows.
// The similarity GEMM runs on tcgen05 through lecb_gemm_f16 with the fp32 query split into an exact
// fp16 (hi, lo) pair, so the ranking matches the reference's fp32 matmul against the fp16 bank; these
// kernels do the split, the per-row top-k selection and the gather-mean.
#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

__global__ void __launch_bounds__(256)
split_f16_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, int64_t n) {
  pdl_grid_sync();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  const __half h = __float2half_rn(v);
  hi[i] = h;
  lo[i] = __float2half_rn(v - __half2float(h));
}

constexpr int kTopK = 10;

__device__ __forceinline__ void topk_insert(float (&val)[kTopK], int (&idx)[kTopK], float v, int i) {
  if (v <= val[kTopK - 1]) return;
  int pos = kTopK - 1;
#pragma unroll
  for (int j = kTopK - 2; j >= 0; --j) {
    if (v > val[j]) {
      val[j + 1] = val[j];
      idx[j + 1] = idx[j];
      pos = j;
    }
  }
  val[pos] = v;
  idx[pos] = i;
}

// One CTA per query row: per-thread sorted top-10 over a strided slice, then 10 rounds of block arg-max.
__global__ void __launch_bounds__(256)
topk10_kernel(const float* __restrict__ sim, int64_t ld, int N, float* __restrict__ out_val,
              int* __restrict__ out_idx) {
  pdl_grid_sync();
  __shared__ float s_val[256 * kTopK];
  __shared__ int s_idx[256 * kTopK];
  __shared__ float r_val[8];
  __shared__ int r_slot[8];
  const int b = blockIdx.x, t = threadIdx.x;
  const float* row = sim + static_cast<int64_t>(b) * ld;
  float val[kTopK];
  int idx[kTopK];
#pragma unroll
  for (int j = 0; j < kTopK; ++j) {
    val[j] = -INFINITY;
    idx[j] = -1;
  }
  for (int i = t; i < N; i += 256) topk_insert(val, idx, __ldg(row + i), i);
#pragma unroll
  for (int j = 0; j < kTopK; ++j) {
    s_val[t * kTopK + j] = val[j];
    s_idx[t * kTopK + j] = idx[j];
  }
  __syncthreads();
  int head = 0;     // each thread's list is sorted: only its current head can be the global maximum
  for (int r = 0; r < kTopK; ++r) {
    float v = head < kTopK ? s_val[t * kTopK + head] : -INFINITY;
    int slot = t;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int os = __shfl_xor_sync(0xffffffffu, slot, o);
      if (ov > v || (ov == v && os < slot)) {
        v = ov;
        slot = os;
      }
    }
    if ((t & 31) == 0) {
      r_val[t >> 5] = v;
      r_slot[t >> 5] = slot;
    }
    __syncthreads();
    float bv = r_val[0];
    int bs = r_slot[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      if (r_val[w] > bv || (r_val[w] == bv && r_slot[w] < bs)) {
        bv = r_val[w];
        bs = r_slot[w];
      }
    }
    if (t == bs) {
      out_val[b * kTopK + r] = bv;
      out_idx[b * kTopK + r] = s_idx[t * kTopK + head];
      ++head;
    }
    __syncthreads();
  }
}

// Merge of the per-CTA partial top-10 lists written by lecb_gemm_topk10: one CTA per query row, candidates
// [slots][10] (each slot sorted descending, unused entries -inf / -1) -> the row's global top-10, descending; of equal
// values the lower bank index wins.
__global__ void __launch_bounds__(256)
topk10_merge_kernel(const float* __restrict__ part_val, const int* __restrict__ part_idx, int slots, float* __restrict__ out_val,
                    int* __restrict__ out_idx) {
  pdl_grid_sync();
  extern __shared__ uint8_t sm_raw[];
  const int b = blockIdx.x, t = threadIdx.x;
  const int n = slots * kTopK;
  float* c_val = reinterpret_cast<float*>(sm_raw);
  int* c_idx = reinterpret_cast<int*>(c_val + n);
  __shared__ float r_val[8];
  __shared__ int r_pos[8];
  for (int i = t; i < n; i += 256) {
    c_val[i] = part_val[static_cast<int64_t>(b) * n + i];
    c_idx[i] = part_idx[static_cast<int64_t>(b) * n + i];
  }
  __syncthreads();
  for (int r = 0; r < kTopK; ++r) {
    float v = -INFINITY;
    int pos = -1, id = 0x7fffffff;
    for (int i = t; i < n; i += 256) {
      const float cv = c_val[i];
      const int ci = c_idx[i];
      if (ci >= 0 && (cv > v || (cv == v && ci < id))) {
        v = cv;
        pos = i;
        id = ci;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int op = __shfl_xor_sync(0xffffffffu, pos, o);
      const int oi = __shfl_xor_sync(0xffffffffu, id, o);
      if (op >= 0 && (pos < 0 || ov > v || (ov == v && oi < id))) {
        v = ov;
        pos = op;
        id = oi;
      }
    }
    if ((t & 31) == 0) {
      r_val[t >> 5] = v;
      r_pos[t >> 5] = pos;
    }
    __syncthreads();
    if (t == 0) {
      float bv = -INFINITY;
      int bp = -1, bi = 0x7fffffff;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const int wp = r_pos[w];
        if (wp < 0) continue;
        const int wi = c_idx[wp];
        if (bp < 0 || r_val[w] > bv || (r_val[w] == bv && wi < bi)) {
          bv = r_val[w];
          bp = wp;
          bi = wi;
        }
      }
      out_val[b * kTopK + r] = bv;
      out_idx[b * kTopK + r] = bp >= 0 ? bi : -1;
      if (bp >= 0) c_idx[bp] = -1;                // taken
    }
    __syncthreads();
  }
}

// g_add[b,:] = mean of the 10 selected bank rows, rounded to the bank's dtype like the reference's
// `caption_text_feats[idx].view(-1, topk, D).mean(1)` on an fp16 tensor.
template <typename TBank>
__global__ void __launch_bounds__(256)
gather_mean_kernel(const TBank* __restrict__ bank, const int* __restrict__ idx, float* __restrict__ out, int D) {
  pdl_grid_sync();
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kTopK; ++j) {
      const int r = idx[b * kTopK + j];
      if constexpr (sizeof(TBank) == 2) s += __half2float(bank[static_cast<int64_t>(r) * D + d]);
      else s += bank[static_cast<int64_t>(r) * D + d];
    }
    s *= 1.0f / kTopK;
    if constexpr (sizeof(TBank) == 2) s = __half2float(__float2half_rn(s));
    out[static_cast<int64_t>(b) * D + d] = s;
  }
}

}  // namespace lecb

using namespace lecb;

// [rows, 2D] = [hi | lo]: the A operand of lecb_gemm_topk10
__global__ void __launch_bounds__(256)
split_f16_hilo_kernel(const float* __restrict__ x, __half* __restrict__ out, int64_t rows, int D) {
  pdl_grid_sync();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * D) return;
  const int64_t r = i / D;
  const int d = static_cast<int>(i - r * D);
  const float v = x[i];
  const __half h = __float2half_rn(v);
  out[r * 2 * D + d] = h;
  out[r * 2 * D + D + d] = __float2half_rn(v - __half2float(h));
}

extern "C" int lecb_split_f16_hilo(const float* x, void* out, int64_t rows, int D, void* stream) {
  LECB_CHECK_ARG(x && out && rows > 0 && D > 0, "lecb_split_f16_hilo: bad argument");
  const int64_t n = rows * D;
  launch_k(split_f16_hilo_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      x, static_cast<__half*>(out), rows, D);
  count_launch();
  return check_launch("split_f16_hilo_kernel");
}

extern "C" int lecb_split_f16(const float* x, void* hi, void* lo, int64_t n, void* stream) {
  LECB_CHECK_ARG(x && hi && lo && n > 0, "lecb_split_f16: bad argument");
  launch_k(split_f16_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      x, static_cast<__half*>(hi), static_cast<__half*>(lo), n);
  count_launch();
  return check_launch("split_f16_kernel");
}

extern "C" int lecb_topk10(const float* sim, int64_t ld, int B, int N, float* out_val, int* out_idx, void* stream) {
  LECB_CHECK_ARG(sim && out_val && out_idx, "lecb_topk10: null pointer");
  LECB_CHECK_ARG(B > 0 && N >= kTopK && ld >= N, "lecb_topk10: need N >= 10 and ld >= N (N=%d)", N);
  launch_k(topk10_kernel, dim3(B), dim3(256), 0, static_cast<cudaStream_t>(stream), sim, ld, N, out_val, out_idx);
  count_launch();
  return check_launch("topk10_kernel");
}

extern "C" int lecb_topk10_merge(const float* part_val, const int* part_idx, int slots, int B, float* out_val, int* out_idx,
                                 void* stream) {
  LECB_CHECK_ARG(part_val && part_idx && out_val && out_idx, "lecb_topk10_merge: null pointer");
  LECB_CHECK_ARG(B > 0 && slots > 0 && slots <= 1024, "lecb_topk10_merge: need 0 < slots <= 1024 (slots=%d)", slots);
  const size_t smem = static_cast<size_t>(slots) * kTopK * 8;
  static bool big_smem[64] = {};
  if (smem > 48 * 1024) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !big_smem[dev]) {
      cudaError_t e = cudaFuncSetAttribute(topk10_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * kTopK * 8);
      if (e != cudaSuccess) return fail(LECB_ERR_CUDA, "lecb_topk10_merge: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      big_smem[dev] = true;
    }
  }
  launch_k(topk10_merge_kernel, dim3(B), dim3(256), smem, static_cast<cudaStream_t>(stream), part_val, part_idx, slots, out_val, out_idx);
  count_launch();
  return check_launch("topk10_merge_kernel");
}

extern "C" int lecb_gather_mean10(const void* bank, int bank_is_f16, const int* idx, float* out, int B, int D,
                                  void* stream) {
  LECB_CHECK_ARG(bank && idx && out && B > 0 && D > 0, "lecb_gather_mean10: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (bank_is_f16) launch_k(gather_mean_kernel<__half>, dim3(B), dim3(256), 0, s, static_cast<const __half*>(bank), idx, out, D);
  else launch_k(gather_mean_kernel<float>, dim3(B), dim3(256), 0, s, static_cast<const float*>(bank), idx, out, D);
  count_launch();
  return check_launch("gather_mean_kernel");
}
