// Host-side plumbing shared by the C-ABI translation units: error slot, launch counter, TMA encoders.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lecb.h"

namespace lecb {

int fail(int status, const char* fmt, ...);          // records the thread-local message, returns status
void count_launch(unsigned n = 1);
int sm_count();                                       // SMs of the current device (cached per device)
int check_launch(const char* what);                   // cudaGetLastError -> status

// Every kernel launch of the library: cudaLaunchKernelEx, with the programmatic-stream-serialization attribute (PDL) where it
// pays, so the kernel's prologue overlaps the tail of the previous kernel in the stream (each kernel executes
// griddepcontrol.wait before it touches global memory: lecb_common.cuh).  Measured policy (profiles/r02_pdl_policy.txt):
//  * the tensor-core kernels (more than 48 KB of dynamic shared memory: GEMM / conv, attention) always carry the attribute;
//  * a row kernel carries it only when its grid is small (<= 8 CTAs per SM: the prompt-tuning step, where launch latency is
//    a fifth of every kernel).  Big row kernels launched early behind a GEMM ran the ViT-B step 2.5 % SLOWER than no PDL at
//    all (24.7 vs 24.1 ms) and the RN101 step 2 % slower, wherever the primary's trigger sat; launched the ordinary way they
//    cost nothing (23.9 ms).
// LECB_PDL_MODE (read once): 1 = that policy (default), 0 = never (same as LECB_NO_PDL=1), 2 = row kernels only, 3 = tensor-core
// kernels only, 4 = every kernel.
int pdl_mode();
template <typename... P, typename... A>
inline void launch_k(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  const int mode = pdl_mode();
  const bool heavy = smem > 48 * 1024;
  bool on = false;
  switch (mode) {
    case 1: on = heavy || static_cast<unsigned long long>(grid.x) * grid.y * grid.z <= 8ull * static_cast<unsigned>(sm_count()); break;
    case 2: on = !heavy; break;
    case 3: on = heavy; break;
    case 4: on = true; break;
    default: break;
  }
  cfg.numAttrs = on ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kern, static_cast<A&&>(args)...);      // errors surface through check_launch()
}

// cudaFuncSetAttribute is a per-DEVICE setting: a `static DeviceOnce once; bool& done = once.flag();` at the call site keeps
// one "configured" flag per device ordinal (a process that drives several GPUs configures each of them once).  A race
// between host threads only repeats the idempotent attribute call.
struct DeviceOnce {
  bool done[64] = {};
  bool scratch = false;
  bool& flag() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
      scratch = false;
      return scratch;
    }
    return done[dev];
  }
};

// cuTensorMapEncode* resolved at run time through cudaGetDriverEntryPoint (no link-time libcuda).
// 2-D row-major bf16 matrix [rows, cols]; box = box_rows x box_cols elements; swizzle = box_cols*2 bytes.
int encode_tiled_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                    uint32_t box_cols);
// same for 2-byte (bf16 / fp16) or 4-byte (fp32) elements
int encode_tiled_2d_ex(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                       uint32_t box_cols, uint32_t elem_bytes);
// 3-D bf16 tensor [batches, rows, cols] (dense); box = 1 x box_rows x box_cols; rows past the end of a batch
// are zero-filled, so a tile never reads the next batch's rows.
int encode_tiled_3d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t batches,
                    uint32_t box_cols, uint32_t box_rows);
// NHWC bf16 tensor as a plain 4-D tiled map (dims {C,W,H,N}); box = 1 x box_h x box_w x box_c
int encode_tiled_4d_nhwc(CUtensorMap* out, const void* base, int B, int H, int W, int C, uint32_t box_c, uint32_t box_w,
                         uint32_t box_h);
// NHWC bf16 tensor, 3x3 / pad 1 / stride 1 im2col window; box = `pixels` x `channels`.
int encode_im2col_3x3(CUtensorMap* out, const void* base, int B, int H, int W, int C, uint32_t channels,
                      uint32_t pixels);

}  // namespace lecb

#define LECB_CHECK_ARG(cond, ...) \
  do {                            \
    if (!(cond)) return ::lecb::fail(LECB_ERR_ARG, __VA_ARGS__); \
  } while (0)
