// Multi-tensor elementwise kernels over the prompt-learner parameter list (six small tensors: ctx, ctx_double,
// ctx_evidence and three scalars — 24 579 floats, or 1.32 M with class-specific contexts): ONE launch per operation
// instead of one ATen kernel per tensor per operation.
//   lecb_ema_update     twin <- momentum * twin + (1 - momentum) * live                    (T:554-559 `_momentum_update`)
//   lecb_pack_f32       flat <- concat(src_i)  (NULL source = zeros: a parameter without gradient, the case DDP's
//                       find_unused_parameters=True covers at T:787)                         (gradient bucket of T:786-787)
//   lecb_unpack_scale   dst_i <- scale * flat[off_i : off_i + n_i]                          (bucket -> .grad, averaged)
//   lecb_sgd_step       torch.optim.SGD semantics (momentum, dampening 0, weight decay, no nesterov) straight from the
//                       flat averaged gradient: the optimiser the reference builds at T:773 (dassl/optim/optimizer.py)
// The pointer lists are HOST arrays (at most kMaxTensors entries); they are copied into the kernel parameter block.
#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

constexpr int kMaxTensors = 16;

struct TensorList {
  const float* src[kMaxTensors];
  float* dst[kMaxTensors];
  float* aux[kMaxTensors];
  long long n[kMaxTensors];
  long long off[kMaxTensors];     // offset of tensor i inside the flat buffer
  int count;
};

// blockIdx.y = tensor, blockIdx.x strides over its elements
__global__ void __launch_bounds__(256) ema_update_kernel(const TensorList tl, float momentum, float one_minus) {
  pdl_grid_sync();
  const int t = blockIdx.y;
  const float* __restrict__ live = tl.src[t];
  float* __restrict__ twin = tl.dst[t];
  const long long n = tl.n[t];
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    // T:559 `param_m.data * momentum + param.data * (1. - momentum)`: two rounded products and a rounded sum (no FMA
    // contraction), `1. - momentum` evaluated in double by Python and rounded to fp32 once — bit-identical to ATen
    twin[i] = __fadd_rn(__fmul_rn(twin[i], momentum), __fmul_rn(live[i], one_minus));
}

__global__ void __launch_bounds__(256) pack_kernel(const TensorList tl, float* __restrict__ flat) {
  pdl_grid_sync();
  const int t = blockIdx.y;
  const float* __restrict__ src = tl.src[t];
  const long long n = tl.n[t], off = tl.off[t];
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    flat[off + i] = src ? src[i] : 0.f;
}

__global__ void __launch_bounds__(256) unpack_scale_kernel(const TensorList tl, const float* __restrict__ flat, float scale) {
  pdl_grid_sync();
  const int t = blockIdx.y;
  float* __restrict__ dst = tl.dst[t];
  const long long n = tl.n[t], off = tl.off[t];
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = scale * flat[off + i];
}

// p <- p - lr * (buf <- momentum * buf + g),  g = grad_scale * flat + weight_decay * p.  The momentum buffers start at
// zero, which reproduces torch's first step (buf <- g) without a step counter — so the launch is CUDA-graph replayable.
__global__ void __launch_bounds__(256)
sgd_step_kernel(const TensorList tl, const float* __restrict__ flat, float grad_scale, float lr, float momentum, float weight_decay) {
  pdl_grid_sync();
  const int t = blockIdx.y;
  float* __restrict__ p = tl.dst[t];
  float* __restrict__ buf = tl.aux[t];
  const long long n = tl.n[t], off = tl.off[t];
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float g = grad_scale * flat[off + i];
    const float w = p[i];
    if (weight_decay != 0.f) g += weight_decay * w;
    if (momentum != 0.f) {
      const float b = momentum * buf[i] + g;
      buf[i] = b;
      g = b;
    }
    p[i] = w - lr * g;
  }
}

static int fill_list(TensorList& tl, const float* const* src, float* const* dst, float* const* aux, const long long* n, int count,
                     const char* what, long long* max_n) {
  if (count <= 0 || count > kMaxTensors) return fail(LECB_ERR_ARG, "%s: need 1 <= count <= %d (count=%d)", what, kMaxTensors, count);
  if (!n) return fail(LECB_ERR_ARG, "%s: null size list", what);
  long long off = 0, mx = 0;
  tl.count = count;
  for (int i = 0; i < count; ++i) {
    if (n[i] <= 0) return fail(LECB_ERR_ARG, "%s: tensor %d is empty", what, i);
    tl.src[i] = src ? src[i] : nullptr;
    tl.dst[i] = dst ? dst[i] : nullptr;
    tl.aux[i] = aux ? aux[i] : nullptr;
    tl.n[i] = n[i];
    tl.off[i] = off;
    off += n[i];
    if (n[i] > mx) mx = n[i];
  }
  *max_n = mx;
  return LECB_OK;
}

static dim3 list_grid(long long max_n, int count) {
  long long bx = (max_n + 255) / 256;
  if (bx > 1024) bx = 1024;
  if (bx < 1) bx = 1;
  return dim3(static_cast<unsigned>(bx), static_cast<unsigned>(count));
}

}  // namespace lecb

using namespace lecb;

extern "C" int lecb_ema_update(const float* const* live, float* const* twin, const long long* n, int count, float momentum,
                               float one_minus_momentum, void* stream) {
  LECB_CHECK_ARG(live && twin, "lecb_ema_update: null pointer list");
  TensorList tl{};
  long long mx = 0;
  int st = fill_list(tl, live, twin, nullptr, n, count, "lecb_ema_update", &mx);
  if (st) return st;
  for (int i = 0; i < count; ++i) LECB_CHECK_ARG(tl.src[i] && tl.dst[i], "lecb_ema_update: tensor %d is null", i);
  launch_k(ema_update_kernel, dim3(list_grid(mx, count)), dim3(256), 0, static_cast<cudaStream_t>(stream), tl, momentum, one_minus_momentum);
  count_launch();
  return check_launch("ema_update_kernel");
}

extern "C" int lecb_pack_f32(const float* const* src, const long long* n, int count, float* flat, void* stream) {
  LECB_CHECK_ARG(src && flat, "lecb_pack_f32: null pointer");
  TensorList tl{};
  long long mx = 0;
  int st = fill_list(tl, src, nullptr, nullptr, n, count, "lecb_pack_f32", &mx);
  if (st) return st;
  launch_k(pack_kernel, dim3(list_grid(mx, count)), dim3(256), 0, static_cast<cudaStream_t>(stream), tl, flat);
  count_launch();
  return check_launch("pack_kernel");
}

extern "C" int lecb_unpack_scale_f32(const float* flat, float* const* dst, const long long* n, int count, float scale,
                                     void* stream) {
  LECB_CHECK_ARG(flat && dst, "lecb_unpack_scale_f32: null pointer");
  TensorList tl{};
  long long mx = 0;
  int st = fill_list(tl, nullptr, dst, nullptr, n, count, "lecb_unpack_scale_f32", &mx);
  if (st) return st;
  for (int i = 0; i < count; ++i) LECB_CHECK_ARG(tl.dst[i], "lecb_unpack_scale_f32: tensor %d is null", i);
  launch_k(unpack_scale_kernel, dim3(list_grid(mx, count)), dim3(256), 0, static_cast<cudaStream_t>(stream), tl, flat, scale);
  count_launch();
  return check_launch("unpack_scale_kernel");
}

extern "C" int lecb_sgd_step(const float* flat_grad, float* const* params, float* const* momentum_buf, const long long* n,
                             int count, float grad_scale, float lr, float momentum, float weight_decay, void* stream) {
  LECB_CHECK_ARG(flat_grad && params, "lecb_sgd_step: null pointer");
  LECB_CHECK_ARG(momentum == 0.f || momentum_buf, "lecb_sgd_step: momentum needs momentum buffers");
  TensorList tl{};
  long long mx = 0;
  int st = fill_list(tl, nullptr, params, momentum_buf, n, count, "lecb_sgd_step", &mx);
  if (st) return st;
  for (int i = 0; i < count; ++i)
    LECB_CHECK_ARG(tl.dst[i] && (momentum == 0.f || tl.aux[i]), "lecb_sgd_step: tensor %d is null", i);
  launch_k(sgd_step_kernel, dim3(list_grid(mx, count)), dim3(256), 0, static_cast<cudaStream_t>(stream), tl, flat_grad, grad_scale, lr, momentum,
                                                                                      weight_decay);
  count_launch();
  return check_launch("sgd_step_kernel");
}
