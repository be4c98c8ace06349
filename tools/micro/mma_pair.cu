// Probe for round 2: tcgen05.mma.cta_group::2 — one M = 256 x N = 256 tile computed by a CTA pair,
// each CTA staging its own 128 rows of A and HALF of the W tile (128 of the 256 n rows), so W crosses L2 -> SM once per 256
// output rows.  Checks the 2-CTA conventions this repo has not used yet (instruction descriptor with M = 256, cta_group::2
// TMEM allocation by one warp of each CTA, multicast commit onto both CTAs' mbarriers, cluster-scope ordering of generic-
// proxy shared-memory writes before the async-proxy MMA) against a CPU product, then times the MMA issue rate.
// NOT part of liblecb.so; compiled here only (no GPU minutes were left in round 1 to run it):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I<csrc> -o mma_pair mma_pair.cu && ./mma_pair
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "lecb_common.cuh"

using namespace lecb;

constexpr int kK = 256;            // reduction length of the test problem (4 k blocks of 64)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// kind::f16 instruction descriptor with M = 256 (cta_group::2): same fields as make_idesc_f16, M >> 4 = 16
__device__ __forceinline__ uint32_t idesc_m256(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// A [256, kK], W [256, kK] bf16 row-major (K-major); D [256, 256] fp32.  One cluster of two CTAs, 128 threads each.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ W, float* __restrict__ D, int timing_iters,
            long long* cycles) {
  __shared__ __align__(1024) uint8_t sA[128 * 64 * 2];      // this CTA's 128 rows of A, one k block, 128B swizzle
  __shared__ __align__(1024) uint8_t sB[128 * 64 * 2];      // this CTA's 128 of the 256 n rows of W
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t cta = cluster_ctarank();
  const int t = threadIdx.x, warp = uniform_warp_idx();
  if (t == 0) {
    mbar_init(&mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {      // one warp of EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = idesc_m256(256);
  uint32_t phase = 0;
  for (int kb = 0; kb < kK / 64; ++kb) {
    // thread t stages row t of this CTA's A half and W half (eight 16-byte chunks each)
    const uint4* ga = reinterpret_cast<const uint4*>(A + static_cast<size_t>(cta * 128 + t) * kK + kb * 64);
    const uint4* gb = reinterpret_cast<const uint4*>(W + static_cast<size_t>(cta * 128 + t) * kK + kb * 64);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      *reinterpret_cast<uint4*>(sA + swizzled_chunk_offset(t, c, 128)) = ga[c];
      *reinterpret_cast<uint4*>(sB + swizzled_chunk_offset(t, c, 128)) = gb[c];
    }
    fence_proxy_async_smem();
    __syncthreads();
    cluster_sync_all();            // the peer's tiles are staged too
    if (cta == 0 && warp == 0 && elect_one()) {
      tc_fence_after();
      const uint64_t adesc = make_kmajor_desc(smem_u32(sA), 128);
      const uint64_t bdesc = make_kmajor_desc(smem_u32(sB), 128);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma2_f16(tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
      umma2_commit(&mma_bar, 0b11);
    }
    mbar_wait(&mma_bar, phase);    // both CTAs: the MMAs that read this k block are complete
    phase ^= 1;
    tc_fence_after();
  }
  // epilogue: thread == TMEM lane == row (cta * 128 + t)
  for (int c = 0; c < 256 / 32; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(c * 32), r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[static_cast<size_t>(cta * 128 + t) * 256 + c * 32 + j] = __uint_as_float(r[j]);
  }
  // issue-rate timing: back-to-back M = 256, N = 256, K = 16 MMAs on the staged tiles (results discarded)
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (timing_iters > 0) {
    if (cta == 0 && warp == 0 && elect_one()) {
      tc_fence_after();
      const uint64_t adesc = make_kmajor_desc(smem_u32(sA), 128);
      const uint64_t bdesc = make_kmajor_desc(smem_u32(sB), 128);
      const long long t0 = clock64();
      for (int it = 0; it < timing_iters; it += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma2_f16(tmem, adesc + 2u * k, bdesc + 2u * k, idesc, 1u);
      }
      umma2_commit(&mma_bar, 0b11);
      mbar_wait(&mma_bar, phase);
      cycles[0] = clock64() - t0;
    } else {
      mbar_wait(&mma_bar, phase);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

int main() {
  std::vector<__nv_bfloat16> hA(256 * kK), hW(256 * kK);
  std::vector<float> fA(256 * kK), fW(256 * kK), hD(256 * 256), ref(256 * 256);
  srand(7);
  for (int i = 0; i < 256 * kK; ++i) {
    hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f);
    hW[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f);
    fA[i] = __bfloat162float(hA[i]);
    fW[i] = __bfloat162float(hW[i]);
  }
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < 256; ++n) {
      float s = 0.f;
      for (int k = 0; k < kK; ++k) s += fA[m * kK + k] * fW[n * kK + k];
      ref[m * 256 + n] = s;
    }
  __nv_bfloat16 *dA, *dW;
  float* dD;
  long long* dC;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dW, hW.size() * 2);
  cudaMalloc(&dD, hD.size() * 4);
  cudaMalloc(&dC, 8);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, hD.size() * 4);
  const int iters = 4096;
  pair_kernel<<<2, 128>>>(dA, dW, dD, iters, dC);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("kernel failed: %s\n", cudaGetErrorString(e));
    return 1;
  }
  long long cyc = 0;
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
  double err = 0.0, scale = 0.0;
  for (size_t i = 0; i < hD.size(); ++i) {
    err = fmax(err, fabs(hD[i] - ref[i]));
    scale = fmax(scale, fabs(ref[i]));
  }
  printf("cta_group::2 M=256 N=256 K=%d: max |D - ref| = %.3g (scale %.3g) -> %s\n", kK, err, scale, err <= 1e-3 * scale ? "OK" : "MISMATCH");
  printf("issue rate: %.1f cycles per M=256,N=256,K=16 MMA (1-CTA M=128,N=256 measured 128.0: equal cycles = 2x the flops per W byte)\n",
         static_cast<double>(cyc) / iters);
  return err <= 1e-3 * scale ? 0 : 2;
}
