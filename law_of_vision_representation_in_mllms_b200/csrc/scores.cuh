// HBM-bound helper kernels of the A score: per-token inverse norms, fp32 -> bf16 hi/lo splitting,
// and the per-pair mean of the row maxima produced by the EPI_ROWMAX GEMM epilogue.
#pragma once
#include "ptx.cuh"

namespace lovr {

// Scale that turns a raw dot product into the reference's cosine (A_score/compute.py:12-15,64-65):
//   n(x) = x / (||x|| + 1e-10);  F.cosine_similarity divides by max(||n(x)||, 1e-8).
//   mode 1 (C score, C_score/pck_train.py:24-29): only n(x) = x / (||x|| + 1e-10).
__device__ __forceinline__ float inv_norm_scale(float sumsq, int mode = 0) {
  const float nrm = sqrtf(sumsq);
  const float denom1 = nrm + 1e-10f;
  if (mode == 1) return 1.0f / denom1;
  const float nn = nrm / denom1;  // norm of the normalised vector
  return 1.0f / (denom1 * fmaxf(nn, 1e-8f));
}

// bf16 rows [rows, D] -> inv[rows]; one warp per row, 16-byte loads.
// inv index = (row / rows_per_batch) * inv_ld + row % rows_per_batch (padded per-batch vectors for the GEMM epilogue).
__global__ void rownorm_inv_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ inv, size_t rows,
                                        int D, int mode = 0, int rows_per_batch = 0x7fffffff, int inv_ld = 0) {
  const int lane = threadIdx.x & 31;
  const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint4* p = reinterpret_cast<const uint4*>(x + row * D);
  float s = 0.f;
  for (int i = lane; i < D / 8; i += 32) {
    const uint4 v = __ldg(p + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
      s += f.x * f.x + f.y * f.y;
    }
  }
  s = warp_sum(s);
  if (lane == 0) inv[(row / rows_per_batch) * inv_ld + row % rows_per_batch] = inv_norm_scale(s, mode);
}

// fp32 rows [rows, D] -> bf16 [rows, 3D] and inv[rows].
//   x = hi + lo + O(2^-17 |x|), hi = bf16(x), lo = bf16(x - hi).
//   side 0 ("other"): [hi | lo | hi]      side 1 ("ref"): [hi | hi | lo]
// so one K=3D GEMM yields  hi·hi + lo·hi + hi·lo  (the dropped lo·lo term is ~2^-18 relative).
__global__ void split3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, float* __restrict__ inv,
                              size_t rows, int D, int side, int mode = 0, int rows_per_batch = 0x7fffffff,
                              int inv_ld = 0) {
  const int lane = threadIdx.x & 31;
  const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* p = reinterpret_cast<const float4*>(x + row * D);
  __nv_bfloat16* o = out + row * 3 * static_cast<size_t>(D);
  float s = 0.f;
  for (int i = lane; i < D / 4; i += 32) {
    const float4 v = __ldg(p + i);
    const float f[4] = {v.x, v.y, v.z, v.w};
    float hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s += f[k] * f[k];
      hi[k] = __bfloat162float(__float2bfloat16(f[k]));
      lo[k] = f[k] - hi[k];
    }
    uint2 whi, wlo;
    whi.x = pack_bf16x2(hi[0], hi[1]);
    whi.y = pack_bf16x2(hi[2], hi[3]);
    wlo.x = pack_bf16x2(lo[0], lo[1]);
    wlo.y = pack_bf16x2(lo[2], lo[3]);
    uint2* o2 = reinterpret_cast<uint2*>(o);
    const int seg = D / 4;
    o2[i] = whi;
    o2[seg + i] = (side == 0) ? wlo : whi;
    o2[2 * seg + i] = (side == 0) ? whi : wlo;
  }
  s = warp_sum(s);
  if (lane == 0) inv[(row / rows_per_batch) * inv_ld + row % rows_per_batch] = inv_norm_scale(s, mode);
}

// pair_score[b] = mean_i rowmax[b, i]; one block per pair, fixed-order tree (deterministic).
__global__ void pair_mean_kernel(const float* __restrict__ rowmax, float* __restrict__ pair_score, int x) {
  __shared__ float part[32];
  const size_t b = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < x; i += blockDim.x) s += rowmax[b * x + i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (threadIdx.x < (blockDim.x >> 5)) ? part[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) pair_score[b] = t / static_cast<float>(x);
  }
}

}  // namespace lovr
