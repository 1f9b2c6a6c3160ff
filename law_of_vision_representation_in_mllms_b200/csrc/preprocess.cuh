// Image preprocessing in front of the tower, bit-exact with the reference's CPU path:
//   llava/mm_utils.py:63-95 (expand2square + HF CLIPImageProcessor.preprocess), the DINOv2 / SigLIP HF processors,
//   C_score/extract_feature.py:65-67 (img.resize((S,S)) + (x/255 - 0.5)*2).
// All of them resize with Pillow's BICUBIC filter on 8-bit RGB.  Pillow's algorithm (libImaging/Resample.c, an
// un-vendored dependency restated in oracle/preprocess_oracle.py) is a separable two-pass convolution, horizontal pass
// first, with the intermediate image rounded to 8 bits, coefficients computed in double precision, normalised and fixed
// to 22 fractional bits.  That makes it integer work and bit-exact on the GPU as long as the coefficient arithmetic
// uses IEEE round-to-nearest double operations without FMA contraction (explicit __d*_rn intrinsics below).
//
//   resample_coeffs_kernel   per image, axis and output index inside the centre crop: bounds + int32 taps
//   resample_h_kernel        RGB HWC uint8 (optionally the virtual expand2square canvas) -> planar temp [3][rows][R] uint8,
//                            only the input rows the vertical pass of the crop rows touches, only the crop columns
//   resample_v_kernel        temp -> 8-bit value -> float through a 3x256 table (HF rescale+normalise or the C-score map)
//                            -> [N,3,R,R] in the tower's pixel dtype
// HBM-bound byte work: per image it reads H*W*3 source bytes once and writes R*R*3 output elements; the temp image is
// written and read once (stays in the 126 MB L2 for typical batches).
#pragma once
#include <cuda_fp16.h>

#include "rowwise.cuh"

namespace lovr {

constexpr int kResamplePrecisionBits = 32 - 8 - 2;

struct PreImg {
  long long src_off;   // byte offset of this image (HWC, 3 bytes per pixel) in the packed source buffer
  long long tmp_off;   // byte offset of its temp image [3][rows][R]
  long long coef_off;  // int32 offset of its coefficient block: H axis R*(2+kx) then V axis R*(2+ky)
  int H, W;            // stored image
  int vH, vW;          // virtual canvas the resize sees (== H, W unless expand2square)
  int pad_top, pad_left;
  int newH, newW;      // resized size
  int crop_top, crop_left;
  int row0, rows;      // canvas rows [row0, row0+rows) feed the vertical pass of the crop rows
  int kx, ky;          // taps per output along x / y (Resample.c ksize)
};

__device__ __forceinline__ double bicubic_filter_rn(double x) {
  // Resample.c bicubic_filter with a = -0.5, every operation rounded separately (no FMA)
  if (x < 0.0) x = -x;
  if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.5, x), 2.5), x), x), 1.0);
  if (x < 2.0) return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), -0.5);
  return 0.0;
}

// One thread per (image, axis, i): output index xx = crop_offset + i of that axis.
// Block layout written: [xmin, count, k[0..ksize)] per output.
__global__ void resample_coeffs_kernel(const PreImg* __restrict__ imgs, int* __restrict__ coefs, int n_img, int R) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int axis = blockIdx.y;  // 0: x (horizontal pass), 1: y (vertical pass)
  const int n = blockIdx.z;
  if (i >= R || n >= n_img) return;
  const PreImg im = imgs[n];
  const int in_size = axis == 0 ? im.vW : im.vH;
  const int out_size = axis == 0 ? im.newW : im.newH;
  const int ksize = axis == 0 ? im.kx : im.ky;
  const int xx = (axis == 0 ? im.crop_left : im.crop_top) + i;
  int* dst = coefs + im.coef_off + (axis == 0 ? 0 : static_cast<long long>(R) * (2 + im.kx)) +
             static_cast<long long>(i) * (2 + ksize);
  const double scale = __ddiv_rn(static_cast<double>(in_size), static_cast<double>(out_size));
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, filterscale);
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dmul_rn(static_cast<double>(xx) + 0.5, scale);
  int xmin = static_cast<int>(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  if (xmax > ksize) xmax = ksize;  // cannot happen (ksize = 2*ceil(support)+1); keeps the writes in the block regardless
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x)
    ww = __dadd_rn(ww, bicubic_filter_rn(__dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss)));
  dst[0] = xmin;
  dst[1] = xmax;
  for (int x = 0; x < ksize; ++x) {
    int k = 0;
    if (x < xmax) {
      double w = bicubic_filter_rn(__dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss));
      if (ww != 0.0) w = __ddiv_rn(w, ww);
      const double f = __dmul_rn(w, static_cast<double>(1 << kResamplePrecisionBits));
      k = w < 0 ? static_cast<int>(__dadd_rn(-0.5, f)) : static_cast<int>(__dadd_rn(0.5, f));
    }
    dst[2 + x] = k;
  }
}

__device__ __forceinline__ int clip8(int v) {
  v >>= kResamplePrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// grid (row blocks, images); each block produces kRowsPerBlock temp rows; thread = output column (crop-relative).
constexpr int kPreRowsPerBlock = 4;
__global__ void __launch_bounds__(256)
resample_h_kernel(const uint8_t* __restrict__ src, const PreImg* __restrict__ imgs, const int* __restrict__ coefs,
                  uint8_t* __restrict__ tmp, int R, uint32_t pad_rgb) {
  const PreImg im = imgs[blockIdx.y];
  const int r_begin = blockIdx.x * kPreRowsPerBlock;
  if (r_begin >= im.rows) return;
  const int r_end = min(im.rows, r_begin + kPreRowsPerBlock);
  const int* cx = coefs + im.coef_off;
  const int stride = 2 + im.kx;
  uint8_t* t = tmp + im.tmp_off;
  const size_t plane = static_cast<size_t>(im.rows) * R;
  const int bg0 = pad_rgb & 255, bg1 = (pad_rgb >> 8) & 255, bg2 = (pad_rgb >> 16) & 255;
  for (int x = threadIdx.x; x < R; x += blockDim.x) {
    const int* c = cx + static_cast<size_t>(x) * stride;
    const int xmin = __ldg(c), cnt = __ldg(c + 1);
    for (int r = r_begin; r < r_end; ++r) {
      const int y = im.row0 + r - im.pad_top;  // stored-image row
      int s0 = 1 << (kResamplePrecisionBits - 1), s1 = s0, s2 = s0;
      if (y >= 0 && y < im.H) {
        const uint8_t* row = src + im.src_off + static_cast<size_t>(y) * im.W * 3;
        for (int j = 0; j < cnt; ++j) {
          const int k = __ldg(c + 2 + j);
          const int xs = xmin + j - im.pad_left;
          int p0 = bg0, p1 = bg1, p2 = bg2;
          if (xs >= 0 && xs < im.W) {
            const uint8_t* px = row + static_cast<size_t>(xs) * 3;
            p0 = __ldg(px);
            p1 = __ldg(px + 1);
            p2 = __ldg(px + 2);
          }
          s0 += p0 * k;
          s1 += p1 * k;
          s2 += p2 * k;
        }
      } else {
        for (int j = 0; j < cnt; ++j) {
          const int k = __ldg(c + 2 + j);
          s0 += bg0 * k;
          s1 += bg1 * k;
          s2 += bg2 * k;
        }
      }
      const size_t o = static_cast<size_t>(r) * R + x;
      t[o] = static_cast<uint8_t>(clip8(s0));
      t[plane + o] = static_cast<uint8_t>(clip8(s1));
      t[2 * plane + o] = static_cast<uint8_t>(clip8(s2));
    }
  }
}

enum : int { PIX_HF_RESCALE_NORMALIZE = 0, PIX_CSCORE = 1 };

struct PixelMap {
  int mode;
  float mean[3], stdv[3];
};

// grid (R output rows, images); thread = output column; all three channels per thread.
__global__ void __launch_bounds__(256)
resample_v_kernel(const uint8_t* __restrict__ tmp, const PreImg* __restrict__ imgs, const int* __restrict__ coefs,
                  void* __restrict__ out, int out_dt, int R, PixelMap pm) {
  __shared__ float lut[3][256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    const int c = i >> 8, v = i & 255;
    float f;
    if (pm.mode == PIX_HF_RESCALE_NORMALIZE) {
      // transformers rescale: float32(float64(v) * (1/255)); normalize: (x - mean) / std in float32
      const float x = __double2float_rn(__dmul_rn(static_cast<double>(v), 1.0 / 255.0));
      f = __fdiv_rn(__fsub_rn(x, pm.mean[c]), pm.stdv[c]);
    } else {
      // (PILToTensor()(img) / 255.0 - 0.5) * 2 in float32
      f = __fmul_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.0f), 0.5f), 2.0f);
    }
    lut[c][v] = f;
  }
  __syncthreads();
  const PreImg im = imgs[blockIdx.y];
  const int y = blockIdx.x;
  const int* c = coefs + im.coef_off + static_cast<size_t>(R) * (2 + im.kx) + static_cast<size_t>(y) * (2 + im.ky);
  const int ymin = __ldg(c) - im.row0, cnt = __ldg(c + 1);
  const uint8_t* t = tmp + im.tmp_off;
  const size_t plane = static_cast<size_t>(im.rows) * R;
  const size_t img_out = static_cast<size_t>(blockIdx.y) * 3 * R * R;
  for (int x = threadIdx.x; x < R; x += blockDim.x) {
    int s0 = 1 << (kResamplePrecisionBits - 1), s1 = s0, s2 = s0;
    for (int j = 0; j < cnt; ++j) {
      const int k = __ldg(c + 2 + j);
      const size_t o = static_cast<size_t>(ymin + j) * R + x;
      s0 += static_cast<int>(t[o]) * k;
      s1 += static_cast<int>(t[plane + o]) * k;
      s2 += static_cast<int>(t[2 * plane + o]) * k;
    }
    const float f[3] = {lut[0][clip8(s0)], lut[1][clip8(s1)], lut[2][clip8(s2)]};
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const size_t o = img_out + (static_cast<size_t>(ch) * R + y) * R + x;
      if (out_dt == DT_F32) reinterpret_cast<float*>(out)[o] = f[ch];
      else if (out_dt == DT_BF16) reinterpret_cast<__nv_bfloat16*>(out)[o] = __float2bfloat16(f[ch]);
      else reinterpret_cast<__half*>(out)[o] = __float2half_rn(f[ch]);
    }
  }
}

}  // namespace lovr
