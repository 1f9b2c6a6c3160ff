// Plain CUDA-core attention (fp32 math on bf16 q/k/v).  NOT the product path: it exists so the
// tcgen05 attention kernel can be cross-checked on the device and bisected when it disagrees.
#pragma once
#include "ptx.cuh"

namespace lovr {

constexpr int kSimtAttnWarps = 8;
constexpr int kSimtKStride = 66;  // bf16 elements per smem row (132 B -> conflict-free column walks)

inline size_t attn_simt_smem_bytes(int T) {
  const int Tp = (T + 31) / 32 * 32;
  return static_cast<size_t>(2) * T * kSimtKStride * 2 + static_cast<size_t>(kSimtAttnWarps) * Tp * 4;
}

// qkv [N*T, 3*H*64] bf16 -> out [N*T, H*64] bf16; one block per (image, head).
template <bool F16>
__global__ void __launch_bounds__(kSimtAttnWarps * 32)
attention_simt_kernel(const uint16_t* __restrict__ qkv, uint16_t* __restrict__ out, int T, int H, float scale) {
  extern __shared__ uint8_t smem_raw[];
  const int Tp = (T + 31) / 32 * 32;
  uint16_t* sK = reinterpret_cast<uint16_t*>(smem_raw);
  uint16_t* sV = sK + static_cast<size_t>(T) * kSimtKStride;
  float* sP = reinterpret_cast<float*>(sV + static_cast<size_t>(T) * kSimtKStride);
  const int n = blockIdx.x / H, h = blockIdx.x % H;
  const int D = H * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint16_t* base = qkv + static_cast<size_t>(n) * T * 3 * D;
  for (int i = threadIdx.x; i < T * 32; i += blockDim.x) {
    const int j = i >> 5, w = i & 31;  // 32 x 4-byte words per 64-element row
    const uint32_t kw = *reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(j) * 3 * D + D + h * 64 + w * 2);
    const uint32_t vw = *reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(j) * 3 * D + 2 * D + h * 64 + w * 2);
    *reinterpret_cast<uint32_t*>(sK + static_cast<size_t>(j) * kSimtKStride + w * 2) = kw;
    *reinterpret_cast<uint32_t*>(sV + static_cast<size_t>(j) * kSimtKStride + w * 2) = vw;
  }
  __syncthreads();
  float* myP = sP + static_cast<size_t>(warp) * Tp;
  for (int t = warp; t < T; t += kSimtAttnWarps) {
    float q[64];
    const uint16_t* qp = base + static_cast<size_t>(t) * 3 * D + h * 64;
#pragma unroll
    for (int d = 0; d < 64; ++d) q[d] = load16<F16>(qp, d);
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) {
      const uint16_t* kr = sK + static_cast<size_t>(j) * kSimtKStride;
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < 64; ++d) s += q[d] * load16<F16>(kr, d);
      s *= scale;
      myP[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float e = expf(myP[j] - mx);
      myP[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < T; ++j) {
      const float pj = myP[j];
      const float2 vv = unpack16x2<F16>(*reinterpret_cast<const uint32_t*>(sV + static_cast<size_t>(j) * kSimtKStride + lane * 2));
      o0 += pj * vv.x;
      o1 += pj * vv.y;
    }
    const float inv = 1.0f / sum;
    *reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(n) * T + t) * D + h * 64 + lane * 2) =
        pack16x2<F16>(o0 * inv, o1 * inv);
    __syncwarp();
  }
}

}  // namespace lovr
