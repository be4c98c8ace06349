/* lecb.h — C ABI of liblecb.so: the B200 (sm_100a) kernels behind the dual-prompt CLIP scoring path.
 *
 * The reference (JarvisUSTC/Language-Enhanced-CLIP-For-Multi-label-Image-Recognition) is pure
 * Python/PyTorch and has no FFI of its own: its "operator interface" for this path is
 * the ATen call issued at each line cited below.  Every entry point replaces the cited call site(s).
 * Paths are relative to project/my_code/:  T = trainers/Caption_distill_double.py,
 * M = clip/model.py, U = trainers/utils.py.
 *
 * Conventions (all entry points):
 *   - plain C: raw DEVICE pointers + explicit sizes; no torch / C++ types in any signature;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised;
 *   - never allocates, never frees, never owns memory; scratch comes in through `ws` arguments;
 *   - returns 0 on success, a negative lecb_status otherwise; `lecb_last_error()` gives the
 *     thread-local message of the last failure on this host thread;
 *   - activations are row-major bf16 "pixel-major" matrices: NHWC images == [B*H*W, C] rows;
 *   - there is NO CPU fallback: without a CUDA device every compute call returns LECB_ERR_CUDA.
 */
#ifndef LECB_H_
#define LECB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LECB_ABI_VERSION 1

enum lecb_status {
  LECB_OK = 0,
  LECB_ERR_ARG = -1,     /* bad shape / alignment / null pointer */
  LECB_ERR_CUDA = -2,    /* a CUDA runtime or driver call failed */
  LECB_ERR_UNSUPPORTED = -3
};

/* epilogue flags for lecb_gemm_bf16 / lecb_conv3x3_bf16 */
#define LECB_EPI_RELU 1u        /* y = max(y, 0) after bias (+ residual)                     */
#define LECB_EPI_QUICKGELU 2u   /* y = y * sigmoid(1.702 y)   (M:202-204)                    */
#define LECB_EPI_OUT_F32 4u     /* store fp32 instead of bf16                                 */

int lecb_abi_version(void);
const char* lecb_last_error(void);
/* Number of kernels this library has launched from this process (bench.py's gpu_launches). */
unsigned long long lecb_launch_count(void);

/* ---- tensor-core GEMM (tcgen05 + TMA + TMEM) --------------------------------------------------
 * out[M,N] = epi( A[M,K] · W[N,K]^T + bias[N] (+ residual[M,N]) ), bf16 operands, fp32 accumulate.
 * Replaces every F.linear / 1x1 nn.Conv2d (+ folded eval BatchNorm + ReLU + residual add) on the path:
 * T:409-410 (v_proj, c_proj per patch), M:43,46,49-52 (Bottleneck conv1/conv3/downsample),
 * M:225-227 (transformer in_proj/out_proj/c_fc/c_proj), T:95,100 (text_projection).
 * Requirements: K % 32 == 0, N % 8 == 0, A/W 16-byte aligned, lda == K, ldw == K.
 * `bias` may be NULL; `residual` (bf16 [M,N]) may be NULL; `row_sumsq` (fp32 [M], pre-zeroed) if
 * non-NULL receives sum_n out[m,n]^2 of the stored (rounded) values — feeds the L2 normalisation. */
int lecb_gemm_bf16(const void* A, const void* W, const float* bias, const void* residual, void* out,
                   float* row_sumsq, int64_t M, int N, int K, unsigned flags, void* stream);

/* ---- implicit-GEMM 3x3 convolution, stride 1, pad 1 (TMA im2col + tcgen05) ---------------------
 * x: NHWC bf16 [B,H,W,Cin]; w: bf16 [Cout][3][3][Cin] with eval-BatchNorm folded in; bias fp32 [Cout];
 * out: NHWC bf16 [B,H,W,Cout].  Replaces M:44 (Bottleneck conv2+bn2+relu) and M:174-175 (stem
 * conv2/conv3).  Requirements: Cin % 32 == 0, Cout % 8 == 0. */
int lecb_conv3x3_bf16(const void* x, const void* w, const float* bias, void* out, int B, int H, int Wd,
                      int Cin, int Cout, unsigned flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LECB_H_ */
