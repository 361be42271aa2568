// CLIPImageProcessor on the device (clip/clip.py:55-58 -> HF CLIPImageProcessor, torchvision backend): resize of the
// shortest edge with ANTIALIASED BICUBIC interpolation on uint8 (ATen's fixed-point kernel: horizontal pass, then
// vertical pass, int16 taps, a rounded uint8 after each pass), centre crop, fused (x - 255 mean) / (255 std).
// The taps come from the host (conzic_b200/imageproc.py restates ATen's weight computation); the kernels do the two
// integer passes for the output columns / rows inside the crop only.  HBM-bound byte work: n * rows * W * 3 bytes in.
#include "kernels.h"

namespace conzic {

namespace {

// tmp[b, r, x, c] = horizontal pass of input row row_lo + r at cropped output column x
__global__ void __launch_bounds__(256) image_resize_h_kernel(ImagePreArgs a) {
  PDL_ENTRY();
  const int r = blockIdx.x, b = blockIdx.y;
  const int R = a.row_hi - a.row_lo;
  const uint8_t* row = a.src + (static_cast<size_t>(b) * a.H + a.row_lo + r) * a.W * 3;
  uint8_t* o = a.tmp + (static_cast<size_t>(b) * R + r) * a.hz.n_out * 3;
  for (int x = threadIdx.x; x < a.hz.n_out; x += blockDim.x) {
    const int first = a.hz.first[x];
    if (a.hz.identity) {
      o[3 * x] = row[3 * first]; o[3 * x + 1] = row[3 * first + 1]; o[3 * x + 2] = row[3 * first + 2];
      continue;
    }
    const int cnt = a.hz.count[x];
    const int16_t* w = a.hz.w + static_cast<size_t>(x) * a.hz.taps;
    int acc0 = 1 << (a.hz.precision - 1), acc1 = acc0, acc2 = acc0;
    for (int j = 0; j < cnt; ++j) {
      const int wj = w[j];
      const uint8_t* px = row + 3 * (first + j);
      acc0 += wj * px[0]; acc1 += wj * px[1]; acc2 += wj * px[2];
    }
    o[3 * x] = static_cast<uint8_t>(min(max(acc0 >> a.hz.precision, 0), 255));
    o[3 * x + 1] = static_cast<uint8_t>(min(max(acc1 >> a.hz.precision, 0), 255));
    o[3 * x + 2] = static_cast<uint8_t>(min(max(acc2 >> a.hz.precision, 0), 255));
  }
}

// out[b, c, y, x] = ((vertical pass at cropped output row y) - mean[c]) / std[c]
__global__ void __launch_bounds__(256) image_resize_v_kernel(ImagePreArgs a) {
  PDL_ENTRY();
  const int y = blockIdx.x, b = blockIdx.y;
  const int R = a.row_hi - a.row_lo, NW = a.hz.n_out, S = a.vt.n_out;
  const int first = a.vt.first[y] - a.row_lo;
  const int cnt = a.vt.count[y];
  const int16_t* w = a.vt.w + static_cast<size_t>(y) * a.vt.taps;
  const uint8_t* t = a.tmp + (static_cast<size_t>(b) * R + first) * NW * 3;
  for (int x = threadIdx.x; x < NW; x += blockDim.x) {
    int v0, v1, v2;
    if (a.vt.identity) {
      v0 = t[3 * x]; v1 = t[3 * x + 1]; v2 = t[3 * x + 2];
    } else {
      int acc0 = 1 << (a.vt.precision - 1), acc1 = acc0, acc2 = acc0;
      for (int j = 0; j < cnt; ++j) {
        const int wj = w[j];
        const uint8_t* px = t + (static_cast<size_t>(j) * NW + x) * 3;
        acc0 += wj * px[0]; acc1 += wj * px[1]; acc2 += wj * px[2];
      }
      v0 = min(max(acc0 >> a.vt.precision, 0), 255);
      v1 = min(max(acc1 >> a.vt.precision, 0), 255);
      v2 = min(max(acc2 >> a.vt.precision, 0), 255);
    }
    float* o = a.out + (static_cast<size_t>(b) * 3 * S + y) * NW + x;
    const size_t plane = static_cast<size_t>(S) * NW;
    o[0] = (static_cast<float>(v0) - a.mean[0]) / a.std[0];
    o[plane] = (static_cast<float>(v1) - a.mean[1]) / a.std[1];
    o[2 * plane] = (static_cast<float>(v2) - a.mean[2]) / a.std[2];
  }
}

}  // namespace

void launch_image_preprocess(const ImagePreArgs& a, cudaStream_t st) {
  const int R = a.row_hi - a.row_lo;
  count_launch();
  {
    ProfScope prof_(CAT_EMBED, static_cast<double>(a.n) * R * a.W * 3, st);
    launch_k(image_resize_h_kernel, dim3(R, a.n), dim3(256), 0, st, a);
  }
  count_launch();
  ProfScope prof_(CAT_EMBED, static_cast<double>(a.n) * R * a.hz.n_out * 3, st);
  launch_k(image_resize_v_kernel, dim3(a.vt.n_out, a.n), dim3(256), 0, st, a);
}

}  // namespace conzic
