// Microbenchmark: issue rate of tcgen05.mma (M=128, K=16, bf16, SS mode) as a function of N, of the number of
// independent accumulators the stream alternates between, and of the operand swizzle (128B / 64B rows).
// One CTA per SM-sized grid of 1 (timing with clock64 inside the issuing thread, completion via commit + mbarrier).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I<csrc> -o mma_chain mma_chain.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "lecb_common.cuh"

using namespace lecb;

template <int NACC>
__global__ void __launch_bounds__(128, 1) chain_kernel(int n, int swz, int iters, int same_ab, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16(n, true);
    const uint64_t a0 = make_kmajor_desc(smem_u32(smem), swz);
    const uint64_t b0 = make_kmajor_desc(smem_u32(smem + 131072), swz);
    const uint64_t astep = same_ab ? 0 : (16384 >> 4);      // next A tile (8 distinct 16 KB tiles)
    long long t0 = clock64();
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        umma_f16(tmem + (j % NACC) * n, a0 + j * astep, b0 + 2 * (j & 1), idesc, 1u);
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int NACC>
static void run(int n, int swz, int iters, int same, long long* d) {
  cudaFuncSetAttribute(chain_kernel<NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  chain_kernel<NACC><<<1, 128, 180 * 1024>>>(n, swz, iters, same, d);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const int iters = 2048;
  printf("%5s %5s %4s %6s | %10s %10s\n", "N", "accs", "swz", "sameAB", "issue cyc", "cyc/mma");
  for (int swz : {128, 64})
    for (int n : {32, 64, 128, 256})
      for (int naccs : {1, 2, 4})
        for (int same : {0, 1}) {
          if (naccs * n > 512) continue;
          if (naccs == 1) run<1>(n, swz, iters, same, d);
          else if (naccs == 2) run<2>(n, swz, iters, same, d);
          else run<4>(n, swz, iters, same, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[2];
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("%5d %5d %4d %6d | %10.1f %10.1f\n", n, naccs, swz, same, (double)h[0] / iters, (double)h[1] / iters);
        }
  return 0;
}
