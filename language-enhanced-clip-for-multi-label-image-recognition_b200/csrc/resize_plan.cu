// Host side of the test-time window pipeline: Pillow-compatible resampling taps.
//
// The reference resizes the image and each of its sliding windows with PIL (`Resize(..., bicubic)`,
// dassl/data/transforms/transforms.py:384-394 via data_manager.py:348-492).  Pillow's 8-bit resampler
// (src/libImaging/Resample.c: precompute_coeffs + normalize_coeffs_8bpc) computes, per output pixel, the filter taps over
// the source pixels within support * max(scale, 1) of the output centre in DOUBLE precision, normalises them, and converts
// them to 22-bit fixed point; the two passes then are pure integer multiply-adds.  Bit-compatibility of a GPU resize
// therefore hinges on these taps: they are produced here, on the host, with the same operation order as Pillow (pinned
// against oracle/pil_resize.py, which is itself bit-exact against Pillow: tests/test_pil_resize.py).  The device kernels
// that consume a plan (crop + two integer passes + normalisation into the trunk's input) are the next step.
#include <cmath>
#include <cstdint>

#include "lecb_host.h"

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

inline double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

inline double bilinear_filter(double x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) return 1.0 - x;
  return 0.0;
}

inline int plan_ksize(int in_size, int out_size, int filter) {
  double filterscale = static_cast<double>(in_size) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = (filter == LECB_RESIZE_BICUBIC ? 2.0 : 1.0) * filterscale;
  return static_cast<int>(std::ceil(support)) * 2 + 1;
}

}  // namespace

using namespace lecb;

extern "C" int lecb_resize_ksize(int in_size, int out_size, int filter) {
  LECB_CHECK_ARG(in_size > 0 && out_size > 0, "lecb_resize_ksize: sizes must be positive (in=%d out=%d)", in_size, out_size);
  LECB_CHECK_ARG(filter == LECB_RESIZE_BILINEAR || filter == LECB_RESIZE_BICUBIC, "lecb_resize_ksize: unknown filter %d", filter);
  return plan_ksize(in_size, out_size, filter);
}

extern "C" int lecb_resize_plan(int in_size, int out_size, int filter, int* bounds, int* coeffs, int ksize) {
  LECB_CHECK_ARG(bounds && coeffs, "lecb_resize_plan: null pointer");
  LECB_CHECK_ARG(in_size > 0 && out_size > 0, "lecb_resize_plan: sizes must be positive (in=%d out=%d)", in_size, out_size);
  LECB_CHECK_ARG(filter == LECB_RESIZE_BILINEAR || filter == LECB_RESIZE_BICUBIC, "lecb_resize_plan: unknown filter %d", filter);
  LECB_CHECK_ARG(ksize >= plan_ksize(in_size, out_size, filter), "lecb_resize_plan: ksize=%d too small (need %d)", ksize,
                 plan_ksize(in_size, out_size, filter));
  const double in0 = 0.0, in1 = static_cast<double>(in_size);
  const double scale = (in1 - in0) / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = (filter == LECB_RESIZE_BICUBIC ? 2.0 : 1.0) * filterscale;
  const double ss = 1.0 / filterscale;
  double* k = new double[ksize];
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = in0 + (xx + 0.5) * scale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      const double arg = (x + xmin - center + 0.5) * ss;
      const double w = filter == LECB_RESIZE_BICUBIC ? bicubic_filter(arg) : bilinear_filter(arg);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    int* row = coeffs + static_cast<int64_t>(xx) * ksize;
    for (int x = 0; x < ksize; ++x) {
      if (x >= xmax) {
        row[x] = 0;
      } else if (k[x] < 0) {
        row[x] = static_cast<int>(-0.5 + k[x] * (1 << kPrecisionBits));
      } else {
        row[x] = static_cast<int>(0.5 + k[x] * (1 << kPrecisionBits));
      }
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  delete[] k;
  return LECB_OK;
}
