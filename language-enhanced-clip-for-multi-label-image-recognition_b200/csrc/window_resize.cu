// Test-time multi-window pipeline on the GPU: crop + Pillow-compatible resize (+ ToTensor + Normalize)
// of a batch of sliding windows of ONE image, straight into the trunk's input.
//
// Reference: `DatasetWrapperWithBlock._transform_image` (dassl/data/data_manager.py:348-492) cuts every window out of the
// decoded image (group 1 out of a reflect-padded / cropped copy) and runs the dataset transform on it:
// `Resize(INPUT.SIZE, bicubic)` on a PIL image, `ToTensor`, `Normalize(mean, std)` (dassl/data/transforms/transforms.py:379-411)
// — about a hundred PIL calls per image on the host, the real cost of the reference's 10-12 h inference (README.md:18).
// Pillow's 8-bit resampler is a two-pass integer convolution (Resample.c): per output pixel a short run of source
// pixels times 22-bit fixed-point taps, accumulated in int32 from 2^21, shifted down and clamped — horizontal pass first,
// ROUNDED TO uint8, then the vertical pass.  The taps are produced on the host by lecb_resize_plan (double precision, the
// same operation order as Pillow); the two integer passes run here and are therefore bit-exact:
//   resize_h_kernel   thread = (window row, output column): reads the window row in place from the source image
//                     (reflection / cropping of group-1 rows resolved per row), writes the [height, S, 3] intermediate
//   resize_v_kernel   thread = (output row, output column): writes uint8 NHWC [n,S,S,3] (input of lecb_stem_conv1_u8)
//                     and / or the reference's float tensor NCHW [n,3,S,S] = ((v / 255) - mean) / std in fp32
// The plan blob is built on the host by lecb_window_plan (below) and uploaded by the caller (the library never allocates).
#include "lecb_common.cuh"
#include "lecb_host.h"

namespace lecb {

constexpr int kWinRec = 16;      // ints per window record in the plan blob
// record: 0 top, 1 left, 2 height, 3 width, 4 pad_top, 5 pad_bottom, 6 hb, 7 hc, 8 hk, 9 vb, 10 vc, 11 vk, 12/13 tmp offset lo/hi

__host__ __device__ inline int padded_row_source(int p, int h, int pad_top, int pad_bottom) {
  // row p of the image after torchvision's F.pad with (top, bottom) = (pad_top, pad_bottom): negative paddings crop
  // first, positive ones reflect about the edges of the cropped image (no edge repeat) — data_manager.py:383-388
  const int crop_top = pad_top < 0 ? -pad_top : 0, crop_bottom = pad_bottom < 0 ? -pad_bottom : 0;
  const int he = h - crop_top - crop_bottom;
  int q = p - (pad_top > 0 ? pad_top : 0);
  if (q < 0) q = -q;
  else if (q >= he) q = 2 * (he - 1) - q;
  return crop_top + q;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= 22;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

struct NormParams {
  float mean[3], std[3];
};

__global__ void __launch_bounds__(256)
resize_h_kernel(const uint8_t* __restrict__ img, int H, int W, int S, const int* __restrict__ plan, uint8_t* __restrict__ tmp) {
  pdl_grid_sync();
  const int* rec = plan + blockIdx.y * kWinRec;
  const int top = rec[0], left = rec[1], height = rec[2], pad_top = rec[4], pad_bottom = rec[5];
  const int* bounds = plan + rec[6];
  const int* coeffs = plan + rec[7];
  const int hk = rec[8];
  uint8_t* wtmp = tmp + ((static_cast<int64_t>(static_cast<uint32_t>(rec[13])) << 32) | static_cast<uint32_t>(rec[12]));
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < height * S; i += gridDim.x * blockDim.x) {
    const int y = i / S, x = i - y * S;
    const int first = bounds[2 * x], cnt = bounds[2 * x + 1];
    const int* k = coeffs + x * hk;
    const uint8_t* row = img + (static_cast<int64_t>(padded_row_source(top + y, H, pad_top, pad_bottom)) * W + left + first) * 3;
    int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
    for (int t = 0; t < cnt; ++t) {
      const int c = __ldg(k + t);
      a0 += c * __ldg(row + 3 * t);
      a1 += c * __ldg(row + 3 * t + 1);
      a2 += c * __ldg(row + 3 * t + 2);
    }
    uint8_t* o = wtmp + (static_cast<int64_t>(y) * S + x) * 3;
    o[0] = clip8(a0);
    o[1] = clip8(a1);
    o[2] = clip8(a2);
  }
}

__global__ void __launch_bounds__(256)
resize_v_kernel(int S, const int* __restrict__ plan, const uint8_t* __restrict__ tmp, uint8_t* __restrict__ out_u8,
                float* __restrict__ out_f32, NormParams nrm) {
  pdl_grid_sync();
  const int* rec = plan + blockIdx.y * kWinRec;
  const int* bounds = plan + rec[9];
  const int* coeffs = plan + rec[10];
  const int vk = rec[11];
  const uint8_t* wtmp = tmp + ((static_cast<int64_t>(static_cast<uint32_t>(rec[13])) << 32) | static_cast<uint32_t>(rec[12]));
  const int64_t wbase = static_cast<int64_t>(blockIdx.y) * S * S * 3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S * S; i += gridDim.x * blockDim.x) {
    const int y = i / S, x = i - y * S;
    const int first = bounds[2 * y], cnt = bounds[2 * y + 1];
    const int* k = coeffs + y * vk;
    const uint8_t* col = wtmp + (static_cast<int64_t>(first) * S + x) * 3;
    int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
    for (int t = 0; t < cnt; ++t) {
      const int c = __ldg(k + t);
      const uint8_t* px = col + static_cast<int64_t>(t) * S * 3;
      a0 += c * px[0];
      a1 += c * px[1];
      a2 += c * px[2];
    }
    const uint8_t r = clip8(a0), g = clip8(a1), b = clip8(a2);
    if (out_u8) {
      uint8_t* o = out_u8 + wbase + (static_cast<int64_t>(y) * S + x) * 3;
      o[0] = r;
      o[1] = g;
      o[2] = b;
    }
    if (out_f32) {
      float* f = out_f32 + wbase + static_cast<int64_t>(y) * S + x;        // [3, S, S] planes of this window
      f[0] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(r), 255.0f), nrm.mean[0]), nrm.std[0]);
      f[static_cast<int64_t>(S) * S] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(g), 255.0f), nrm.mean[1]), nrm.std[1]);
      f[2 * static_cast<int64_t>(S) * S] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(b), 255.0f), nrm.mean[2]), nrm.std[2]);
    }
  }
}

static bool window_ok(const int* w, int H, int W) {
  const int top = w[0], left = w[1], height = w[2], width = w[3], pad_top = w[4], pad_bottom = w[5];
  if (height <= 0 || width <= 0 || left < 0 || left + width > W || top < 0) return false;
  const int crop_top = pad_top < 0 ? -pad_top : 0, crop_bottom = pad_bottom < 0 ? -pad_bottom : 0;
  const int he = H - crop_top - crop_bottom;
  if (he <= 0 || (pad_top > 0 ? pad_top : 0) >= he || (pad_bottom > 0 ? pad_bottom : 0) >= he) return false;
  return top + height <= H + pad_top + pad_bottom;
}

}  // namespace lecb

using namespace lecb;

// Sizes of the plan blob (ints) and of the intermediate buffer (bytes) for `n` windows given as HOST int32 [n][6]
// records (top, left, height, width, pad_top, pad_bottom) — lecb200.windows.Window.
extern "C" int lecb_window_plan_size(const int* wins, int n, int H, int W, int S, int filter, long long* plan_ints,
                                     long long* tmp_bytes) {
  LECB_CHECK_ARG(wins && plan_ints && tmp_bytes, "lecb_window_plan_size: null pointer");
  LECB_CHECK_ARG(n > 0 && H > 0 && W > 0 && S > 0, "lecb_window_plan_size: bad sizes (n=%d H=%d W=%d S=%d)", n, H, W, S);
  long long ints = static_cast<long long>(n) * kWinRec, tmp = 0;
  for (int i = 0; i < n; ++i) {
    const int* w = wins + 6 * i;
    LECB_CHECK_ARG(window_ok(w, H, W), "lecb_window_plan_size: window %d (top %d left %d %dx%d pad %d/%d) does not fit a %d x %d image",
                   i, w[0], w[1], w[2], w[3], w[4], w[5], H, W);
    const int hk = lecb_resize_ksize(w[3], S, filter), vk = lecb_resize_ksize(w[2], S, filter);
    if (hk < 0 || vk < 0) return hk < 0 ? hk : vk;
    ints += 2LL * S + static_cast<long long>(S) * hk + 2LL * S + static_cast<long long>(S) * vk;
    tmp += static_cast<long long>(w[2]) * S * 3;
  }
  LECB_CHECK_ARG(ints < 0x7fffffffLL, "lecb_window_plan_size: plan too large");
  *plan_ints = ints;
  *tmp_bytes = tmp;
  return LECB_OK;
}

// Fills the HOST blob `plan` (plan_ints int32, from lecb_window_plan_size): n window records followed by every window's
// horizontal (width -> S) and vertical (height -> S) bounds / coefficient arrays (lecb_resize_plan).
extern "C" int lecb_window_plan(const int* wins, int n, int H, int W, int S, int filter, int* plan, long long plan_ints) {
  LECB_CHECK_ARG(wins && plan, "lecb_window_plan: null pointer");
  long long need = 0, tmp_total = 0;
  int st = lecb_window_plan_size(wins, n, H, W, S, filter, &need, &tmp_total);
  if (st) return st;
  LECB_CHECK_ARG(plan_ints >= need, "lecb_window_plan: blob too small (%lld < %lld ints)", plan_ints, need);
  long long off = static_cast<long long>(n) * kWinRec, tmp = 0;
  for (int i = 0; i < n; ++i) {
    const int* w = wins + 6 * i;
    int* rec = plan + static_cast<long long>(i) * kWinRec;
    for (int j = 0; j < 6; ++j) rec[j] = w[j];
    for (int axis = 0; axis < 2; ++axis) {
      const int in_size = axis == 0 ? w[3] : w[2];
      const int ks = lecb_resize_ksize(in_size, S, filter);
      const long long bo = off, co = off + 2LL * S;
      st = lecb_resize_plan(in_size, S, filter, plan + bo, plan + co, ks);
      if (st) return st;
      rec[6 + 3 * axis] = static_cast<int>(bo);
      rec[7 + 3 * axis] = static_cast<int>(co);
      rec[8 + 3 * axis] = ks;
      off = co + static_cast<long long>(S) * ks;
    }
    rec[12] = static_cast<int>(static_cast<uint32_t>(tmp & 0xffffffffLL));
    rec[13] = static_cast<int>(static_cast<uint32_t>(tmp >> 32));
    rec[14] = rec[15] = 0;
    tmp += static_cast<long long>(w[2]) * S * 3;
  }
  return LECB_OK;
}

// img uint8 [H,W,3] (device); plan = the uploaded blob; tmp = tmp_bytes of device scratch; out_u8 [n,S,S,3] and / or
// out_f32 [n,3,S,S] (either may be NULL); mean / std are HOST float[3] (used for out_f32 only).
extern "C" int lecb_crop_resize_u8(const uint8_t* img, int H, int W, const int* plan, int n, int S, uint8_t* tmp,
                                   uint8_t* out_u8, float* out_f32, const float* mean, const float* stdv, void* stream) {
  LECB_CHECK_ARG(img && plan && tmp && (out_u8 || out_f32), "lecb_crop_resize_u8: null pointer");
  LECB_CHECK_ARG(n > 0 && n <= 65535 && H > 0 && W > 0 && S > 0, "lecb_crop_resize_u8: bad sizes (n=%d H=%d W=%d S=%d)", n, H, W, S);
  LECB_CHECK_ARG(!out_f32 || (mean && stdv && stdv[0] != 0.f && stdv[1] != 0.f && stdv[2] != 0.f),
                 "lecb_crop_resize_u8: the float output needs mean and a non-zero std");
  NormParams nrm{};
  for (int c = 0; c < 3; ++c) {
    nrm.mean[c] = mean ? mean[c] : 0.f;
    nrm.std[c] = stdv ? stdv[c] : 1.f;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int bx = (S * S + 255) / 256 < 64 ? (S * S + 255) / 256 : 64;
  const dim3 grid(bx > 0 ? bx : 1, n);
  launch_k(resize_h_kernel, dim3(grid), dim3(256), 0, s, img, H, W, S, plan, tmp);
  count_launch();
  int st = check_launch("resize_h_kernel");
  if (st) return st;
  launch_k(resize_v_kernel, dim3(grid), dim3(256), 0, s, S, plan, tmp, out_u8, out_f32, nrm);
  count_launch();
  return check_launch("resize_v_kernel");
}
