// Probe for the next §8f-1 step: crop + Pillow-compatible two-pass integer resize + normalisation of a batch
// of sliding windows on the GPU, straight into the trunk's NCHW fp32 input.  Stand-alone (not part of liblecb.so):
//   crop_resize_probe <in.bin> <out.bin>
// in.bin  : int32 header {H, W, S, n_win}, n_win x int32 {top, left, height, width, pad_top, pad_bottom} (lecb200.windows.Window),
//           then the uint8 HWC image.  out.bin: n_win x [S,S,3] uint8 resized windows, then n_win x [3,S,S] float32 normalised.
// tools/micro/crop_resize_check.py builds the input from lecb200.windows, runs this binary and compares the uint8 part
// bit for bit with oracle/pil_resize.py (itself bit-exact against Pillow) and the float part to 1e-6.
// The resampling taps come from liblecb.so's lecb_resize_plan (host), which the checker links in by path.
// Compiled only — no GPU minutes were left in round 1 to run it:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I<repo>/include -o crop_resize_probe crop_resize_probe.cu \
//        <repo>/language-enhanced-clip-for-multi-label-image-recognition_b200/csrc/liblecb.so
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#include "lecb.h"

struct Win {
  int top, left, height, width, pad_top, pad_bottom;
};
struct WinPlan {            // per window: offsets into the concatenated plan arrays
  int hb, hc, hk;           // horizontal: bounds offset (ints), coeff offset (ints), taps per output pixel
  int vb, vc, vk;           // vertical
  int tmp;                  // offset (bytes) of the window's [height, S, 3] intermediate
};

__host__ __device__ inline int padded_row_source(int p, int h, int pad_top, int pad_bottom) {
  const int crop_top = pad_top < 0 ? -pad_top : 0, crop_bottom = pad_bottom < 0 ? -pad_bottom : 0;
  const int he = h - crop_top - crop_bottom;
  int q = p - (pad_top > 0 ? pad_top : 0);
  if (q < 0) q = -q;
  else if (q >= he) q = 2 * (he - 1) - q;
  return crop_top + q;
}

__device__ inline uint8_t clip8(int v) {
  v >>= 22;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: thread = (window row y, output column x); three channels per thread
__global__ void resize_h_kernel(const uint8_t* __restrict__ img, int H, int W, int S, const Win* __restrict__ wins,
                                const WinPlan* __restrict__ plans, const int* __restrict__ bounds, const int* __restrict__ coeffs,
                                uint8_t* __restrict__ tmp) {
  const Win w = wins[blockIdx.y];
  const WinPlan p = plans[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w.height * S; i += gridDim.x * blockDim.x) {
    const int y = i / S, x = i - y * S;
    const int first = bounds[p.hb + 2 * x], cnt = bounds[p.hb + 2 * x + 1];
    const int* k = coeffs + p.hc + x * p.hk;
    const uint8_t* row = img + (static_cast<size_t>(padded_row_source(w.top + y, H, w.pad_top, w.pad_bottom)) * W + w.left + first) * 3;
    int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
    for (int t = 0; t < cnt; ++t) {
      const int c = k[t];
      a0 += c * row[3 * t];
      a1 += c * row[3 * t + 1];
      a2 += c * row[3 * t + 2];
    }
    uint8_t* o = tmp + p.tmp + (static_cast<size_t>(y) * S + x) * 3;
    o[0] = clip8(a0);
    o[1] = clip8(a1);
    o[2] = clip8(a2);
  }
}

// vertical pass + ToTensor + Normalize: thread = (output row y, output column x)
__global__ void resize_v_kernel(int S, const Win* __restrict__ wins, const WinPlan* __restrict__ plans, const int* __restrict__ bounds,
                                const int* __restrict__ coeffs, const uint8_t* __restrict__ tmp, uint8_t* __restrict__ out_u8,
                                float* __restrict__ out_f32, float m0, float m1, float m2, float s0, float s1, float s2) {
  const WinPlan p = plans[blockIdx.y];
  const size_t wbase = static_cast<size_t>(blockIdx.y) * S * S * 3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S * S; i += gridDim.x * blockDim.x) {
    const int y = i / S, x = i - y * S;
    const int first = bounds[p.vb + 2 * y], cnt = bounds[p.vb + 2 * y + 1];
    const int* k = coeffs + p.vc + y * p.vk;
    const uint8_t* col = tmp + p.tmp + (static_cast<size_t>(first) * S + x) * 3;
    int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
    for (int t = 0; t < cnt; ++t) {
      const int c = k[t];
      const uint8_t* px = col + static_cast<size_t>(t) * S * 3;
      a0 += c * px[0];
      a1 += c * px[1];
      a2 += c * px[2];
    }
    const uint8_t r = clip8(a0), g = clip8(a1), b = clip8(a2);
    uint8_t* o = out_u8 + wbase + (static_cast<size_t>(y) * S + x) * 3;
    o[0] = r; o[1] = g; o[2] = b;
    float* f = out_f32 + wbase + static_cast<size_t>(y) * S + x;        // [3, S, S] planes of this window
    f[0] = (r / 255.0f - m0) / s0;
    f[static_cast<size_t>(S) * S] = (g / 255.0f - m1) / s1;
    f[2 * static_cast<size_t>(S) * S] = (b / 255.0f - m2) / s2;
  }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  FILE* fi = fopen(argv[1], "rb");
  if (!fi) { perror("in.bin"); return 2; }
  int hdr[4];
  if (fread(hdr, 4, 4, fi) != 4) return 2;
  const int H = hdr[0], W = hdr[1], S = hdr[2], n = hdr[3];
  std::vector<Win> wins(n);
  if (fread(wins.data(), sizeof(Win), n, fi) != static_cast<size_t>(n)) return 2;
  std::vector<uint8_t> img(static_cast<size_t>(H) * W * 3);
  if (fread(img.data(), 1, img.size(), fi) != img.size()) return 2;
  fclose(fi);
  // plans: one horizontal (width -> S) and one vertical (height -> S) per window
  std::vector<int> bounds, coeffs;
  std::vector<WinPlan> plans(n);
  size_t tmp_bytes = 0;
  for (int i = 0; i < n; ++i) {
    WinPlan& p = plans[i];
    for (int axis = 0; axis < 2; ++axis) {
      const int in_size = axis == 0 ? wins[i].width : wins[i].height;
      const int ks = lecb_resize_ksize(in_size, S, LECB_RESIZE_BICUBIC);
      if (ks <= 0) { fprintf(stderr, "plan: %s\n", lecb_last_error()); return 1; }
      const int bo = static_cast<int>(bounds.size()), co = static_cast<int>(coeffs.size());
      bounds.resize(bo + 2 * S);
      coeffs.resize(co + static_cast<size_t>(S) * ks);
      if (lecb_resize_plan(in_size, S, LECB_RESIZE_BICUBIC, bounds.data() + bo, coeffs.data() + co, ks) != 0) {
        fprintf(stderr, "plan: %s\n", lecb_last_error());
        return 1;
      }
      if (axis == 0) { p.hb = bo; p.hc = co; p.hk = ks; } else { p.vb = bo; p.vc = co; p.vk = ks; }
    }
    p.tmp = static_cast<int>(tmp_bytes);
    tmp_bytes += static_cast<size_t>(wins[i].height) * S * 3;
  }
  uint8_t *d_img, *d_tmp, *d_u8;
  float* d_f32;
  Win* d_wins;
  WinPlan* d_plans;
  int *d_bounds, *d_coeffs;
  const size_t out_px = static_cast<size_t>(n) * S * S * 3;
  CK(cudaMalloc(&d_img, img.size()));
  CK(cudaMalloc(&d_tmp, tmp_bytes));
  CK(cudaMalloc(&d_u8, out_px));
  CK(cudaMalloc(&d_f32, out_px * 4));
  CK(cudaMalloc(&d_wins, n * sizeof(Win)));
  CK(cudaMalloc(&d_plans, n * sizeof(WinPlan)));
  CK(cudaMalloc(&d_bounds, bounds.size() * 4));
  CK(cudaMalloc(&d_coeffs, coeffs.size() * 4));
  CK(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_wins, wins.data(), n * sizeof(Win), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_plans, plans.data(), n * sizeof(WinPlan), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_bounds, bounds.data(), bounds.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_coeffs, coeffs.data(), coeffs.size() * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const dim3 grid(64, n);
  for (int rep = 0; rep < 2; ++rep) {          // second repetition is the timed one
    cudaEventRecord(e0);
    resize_h_kernel<<<grid, 256>>>(d_img, H, W, S, d_wins, d_plans, d_bounds, d_coeffs, d_tmp);
    resize_v_kernel<<<grid, 256>>>(S, d_wins, d_plans, d_bounds, d_coeffs, d_tmp, d_u8, d_f32, 0.48145466f, 0.4578275f, 0.40821073f,
                                   0.26862954f, 0.26130258f, 0.27577711f);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<uint8_t> h_u8(out_px);
  std::vector<float> h_f32(out_px);
  CK(cudaMemcpy(h_u8.data(), d_u8, out_px, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_f32.data(), d_f32, out_px * 4, cudaMemcpyDeviceToHost));
  FILE* fo = fopen(argv[2], "wb");
  fwrite(h_u8.data(), 1, out_px, fo);
  fwrite(h_f32.data(), 4, out_px, fo);
  fclose(fo);
  printf("%d windows of a %d x %d image -> %d x %d: %.3f ms on the GPU (two kernels)\n", n, H, W, S, S, ms);
  return 0;
}
