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
//   pixel_lut_kernel         the 8-bit value -> float map as a 3x256 table (HF rescale+normalise or the C-score map)
//   resample_h_kernel        RGB HWC uint8 (optionally the virtual expand2square canvas) -> planar temp [3][rows][R] uint8,
//                            only the input rows the vertical pass of the crop rows touches, only the crop columns
//   resample_v4_kernel       temp -> 8-bit value -> table -> [N,3,R,R] in the tower's pixel dtype, 4 columns per thread
//                            (resample_v_kernel: any R, one column per thread)
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

// grid (row blocks, images); each block produces kPreRowsPerBlock temp rows; thread = output column (crop-relative).
// Taps outer, rows inner: one coefficient load feeds 4 rows x 3 channels.  PAD = the expand2square canvas (pixels outside
// the pasted image read the background colour); without it every tap is inside the image by construction of the bounds.
constexpr int kPreRowsPerBlock = 4;
template <bool PAD>
__global__ void __launch_bounds__(128)
resample_h_kernel(const uint8_t* __restrict__ src, const PreImg* __restrict__ imgs, const int* __restrict__ coefs,
                  uint8_t* __restrict__ tmp, int R, uint32_t pad_rgb) {
  const PreImg im = imgs[blockIdx.y];
  const int r_begin = blockIdx.x * kPreRowsPerBlock;
  if (r_begin >= im.rows) return;
  const int* cx = coefs + im.coef_off;
  const int stride = 2 + im.kx;
  uint8_t* t = tmp + im.tmp_off;
  const size_t plane = static_cast<size_t>(im.rows) * R;
  const int bg[3] = {static_cast<int>(pad_rgb & 255), static_cast<int>((pad_rgb >> 8) & 255),
                     static_cast<int>((pad_rgb >> 16) & 255)};
  const uint8_t* rowp[kPreRowsPerBlock];
  bool row_in[kPreRowsPerBlock], row_live[kPreRowsPerBlock];
#pragma unroll
  for (int q = 0; q < kPreRowsPerBlock; ++q) {
    const int r = r_begin + q;
    row_live[q] = r < im.rows;
    const int y = im.row0 + min(r, im.rows - 1) - im.pad_top;  // stored-image row
    row_in[q] = !PAD || (y >= 0 && y < im.H);
    rowp[q] = src + im.src_off + static_cast<size_t>(row_in[q] ? y : 0) * im.W * 3;
  }
  for (int x = threadIdx.x; x < R; x += blockDim.x) {
    const int* c = cx + static_cast<size_t>(x) * stride;
    const int xmin = __ldg(c), cnt = __ldg(c + 1);
    int acc[kPreRowsPerBlock][3];
#pragma unroll
    for (int q = 0; q < kPreRowsPerBlock; ++q) acc[q][0] = acc[q][1] = acc[q][2] = 1 << (kResamplePrecisionBits - 1);
    for (int j = 0; j < cnt; ++j) {
      const int k = __ldg(c + 2 + j);
      const int xs = xmin + j - (PAD ? im.pad_left : 0);
      const bool col_in = !PAD || (xs >= 0 && xs < im.W);
      const size_t off = static_cast<size_t>(col_in ? xs : 0) * 3;
#pragma unroll
      for (int q = 0; q < kPreRowsPerBlock; ++q) {
        int p0 = bg[0], p1 = bg[1], p2 = bg[2];
        if (col_in && row_in[q]) {
          const uint8_t* px = rowp[q] + off;
          p0 = __ldg(px);
          p1 = __ldg(px + 1);
          p2 = __ldg(px + 2);
        }
        acc[q][0] += p0 * k;
        acc[q][1] += p1 * k;
        acc[q][2] += p2 * k;
      }
    }
#pragma unroll
    for (int q = 0; q < kPreRowsPerBlock; ++q) {
      if (row_live[q]) {
        const size_t o = static_cast<size_t>(r_begin + q) * R + x;
        t[o] = static_cast<uint8_t>(clip8(acc[q][0]));
        t[plane + o] = static_cast<uint8_t>(clip8(acc[q][1]));
        t[2 * plane + o] = static_cast<uint8_t>(clip8(acc[q][2]));
      }
    }
  }
}

enum : int { PIX_HF_RESCALE_NORMALIZE = 0, PIX_CSCORE = 1 };

struct PixelMap {
  int mode;
  float mean[3], stdv[3];
};

// 8-bit value -> float table [3][256], built once per call (IEEE operations spelled out: results equal the CPU path's)
__global__ void pixel_lut_kernel(float* __restrict__ lut, PixelMap pm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 768) return;
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
  lut[i] = f;
}

template <int DT>
__device__ __forceinline__ void store_px4(void* out, size_t o, float a, float b, float c, float d) {
  if constexpr (DT == DT_F32) {
    *reinterpret_cast<float4*>(static_cast<float*>(out) + o) = make_float4(a, b, c, d);
  } else if constexpr (DT == DT_BF16) {
    *reinterpret_cast<uint2*>(static_cast<uint16_t*>(out) + o) = make_uint2(pack16x2<false>(a, b), pack16x2<false>(c, d));
  } else {
    *reinterpret_cast<uint2*>(static_cast<uint16_t*>(out) + o) = make_uint2(pack16x2<true>(a, b), pack16x2<true>(c, d));
  }
}

// Vertical pass + value map + NCHW store.  R % 4 == 0: thread = 4 consecutive output columns of one output row, all three
// channels (32-bit temp loads, 8- or 16-byte stores); work items flattened over (row, column group).  grid (items / 256, images).
template <int DT>
__global__ void __launch_bounds__(256)
resample_v4_kernel(const uint8_t* __restrict__ tmp, const PreImg* __restrict__ imgs, const int* __restrict__ coefs,
                   const float* __restrict__ lut_g, void* __restrict__ out, int R) {
  __shared__ float lut[768];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i] = __ldg(lut_g + i);
  __syncthreads();
  const int R4 = R >> 2;
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= R * R4) return;
  const PreImg im = imgs[blockIdx.y];
  const int y = item / R4, x = (item - y * R4) * 4;
  const int* c = coefs + im.coef_off + static_cast<size_t>(R) * (2 + im.kx) + static_cast<size_t>(y) * (2 + im.ky);
  const int ymin = __ldg(c) - im.row0, cnt = __ldg(c + 1);
  const uint8_t* t = tmp + im.tmp_off + static_cast<size_t>(ymin) * R + x;
  const size_t plane = static_cast<size_t>(im.rows) * R;
  int acc[3][4];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[ch][e] = 1 << (kResamplePrecisionBits - 1);
  for (int j = 0; j < cnt; ++j) {
    const int k = __ldg(c + 2 + j);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(t + ch * plane + static_cast<size_t>(j) * R);
      acc[ch][0] += static_cast<int>(w & 255u) * k;
      acc[ch][1] += static_cast<int>((w >> 8) & 255u) * k;
      acc[ch][2] += static_cast<int>((w >> 16) & 255u) * k;
      acc[ch][3] += static_cast<int>(w >> 24) * k;
    }
  }
  const size_t img_out = static_cast<size_t>(blockIdx.y) * 3 * R * R;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float* l = lut + ch * 256;
    store_px4<DT>(out, img_out + (static_cast<size_t>(ch) * R + y) * R + x, l[clip8(acc[ch][0])], l[clip8(acc[ch][1])],
                  l[clip8(acc[ch][2])], l[clip8(acc[ch][3])]);
  }
}

// Generic vertical pass (any R): grid (R output rows, images); thread = output column; all three channels per thread.
__global__ void __launch_bounds__(256)
resample_v_kernel(const uint8_t* __restrict__ tmp, const PreImg* __restrict__ imgs, const int* __restrict__ coefs,
                  const float* __restrict__ lut_g, void* __restrict__ out, int out_dt, int R) {
  __shared__ float lut[3][256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i >> 8][i & 255] = __ldg(lut_g + i);
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
