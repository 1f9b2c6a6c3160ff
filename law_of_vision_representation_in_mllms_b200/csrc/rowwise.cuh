// HBM-bound row-wise kernels around the GEMMs: im2col for the patch-embed GEMM, embedding assembly
// (+cls, +pos, CLIP pre_layrnorm), LayerNorm fp32 -> bf16, final feature select/cast, row L2 norms.
// One warp per token row, 128-bit vectorised, fp32 statistics, warp-shuffle reductions.
#pragma once
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace lovr {

enum : int { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };

__device__ __forceinline__ float load_as_float(const void* base, size_t idx, int dt) {
  if (dt == DT_F32) return __ldg(reinterpret_cast<const float*>(base) + idx);
  if (dt == DT_BF16) return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(base) + idx));
  return __half2float(__ldg(reinterpret_cast<const __half*>(base) + idx));
}

// pixels [N,3,R,R] (dt) -> cols [N*G*G, Kpad] bf16, k = c*P*P + py*P + px (conv-weight flatten order),
// zero for k >= 3*P*P.  Reference arithmetic: the stride-P conv of HF *Embeddings (Appendix A).
__global__ void im2col_kernel(const void* __restrict__ pixels, int dt, uint16_t* __restrict__ cols, int n_img,
                              int R, int P, int G, int Kpad, int f16) {
  const size_t total = static_cast<size_t>(n_img) * G * G * Kpad;
  const int K = 3 * P * P;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % Kpad);
    const size_t row = i / Kpad;
    float v = 0.f;
    if (k < K) {
      const int gx = static_cast<int>(row % G);
      const int gy = static_cast<int>((row / G) % G);
      const size_t n = row / (static_cast<size_t>(G) * G);
      const int c = k / (P * P);
      const int py = (k / P) % P;
      const int px = k % P;
      v = load_as_float(pixels, ((n * 3 + c) * R + (gy * P + py)) * static_cast<size_t>(R) + gx * P + px, dt);
    }
    cols[i] = f16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16(v));
  }
}

// Same gather when the pixels already are in the 16-bit operand format (the extraction path: bf16 in, bf16 operands): pure
// 32-bit word copies, one warp per patch row, constant divisors.  P even, R even: every P-element source run and every
// destination word is 4-byte aligned.  The generic kernel above ran at 0.5 TB/s (a div/mod chain per element).
template <int P>
__global__ void im2col16_kernel(const uint32_t* __restrict__ pixels, uint32_t* __restrict__ cols, int n_img, int R, int G,
                                int Kpad) {
  constexpr int K = 3 * P * P;
  const int lane = threadIdx.x & 31;
  const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= static_cast<size_t>(n_img) * G * G) return;
  const int gx = static_cast<int>(row % G);
  const int gy = static_cast<int>((row / G) % G);
  const size_t n = row / (static_cast<size_t>(G) * G);
  uint32_t* dst = cols + row * (Kpad >> 1);
  for (int w = lane; w < (Kpad >> 1); w += 32) {
    const int k = 2 * w;
    uint32_t v = 0u;
    if (k < K) {
      const int seg = k / P, off = k - seg * P;       // seg = c * P + py
      const int c = seg / P, py = seg - c * P;
      v = __ldg(pixels + ((((n * 3 + c) * R + (gy * P + py)) * static_cast<size_t>(R) + gx * P + off) >> 1));
    }
    dst[w] = v;
  }
}

// V4 = float4 per lane (D = 128 * V4).  Two-pass statistics on register-resident values.
template <int V4>
__device__ __forceinline__ void warp_layernorm(float (&v)[V4 * 4], const float* __restrict__ gamma,
                                               const float* __restrict__ beta, float eps, int lane) {
  constexpr int D = V4 * 128;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V4 * 4; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V4 * 4; ++i) {
    const float d = v[i] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
  for (int j = 0; j < V4; ++j) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + j * 32 + lane);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + j * 32 + lane);
    v[j * 4 + 0] = (v[j * 4 + 0] - mean) * rstd * g.x + b.x;
    v[j * 4 + 1] = (v[j * 4 + 1] - mean) * rstd * g.y + b.y;
    v[j * 4 + 2] = (v[j * 4 + 2] - mean) * rstd * g.z + b.z;
    v[j * 4 + 3] = (v[j * 4 + 3] - mean) * rstd * g.w + b.w;
  }
}

// x[n*T + t, :] = (t is cls ? cls : patch[n*(T-has_cls) + t - has_cls, :]) + pos[t, :]; optional pre-LN.
// With LayerNorm folding (xhat != nullptr) the row also leaves as its 16-bit copy x^ and as one statistics slot per 128
// columns (shift, sum(x - shift), sum((x - shift)^2), 0) — the inputs of layer 0's folded LayerNorm (gemm.cuh).
template <int V4, bool F16>
__global__ void embed_finalize_kernel(const float* __restrict__ patch, const float* __restrict__ cls,
                                      const float* __restrict__ pos, const float* __restrict__ ln_w,
                                      const float* __restrict__ ln_b, float eps, float* __restrict__ x, int n_img,
                                      int T, int has_cls, uint16_t* __restrict__ xhat, float* __restrict__ stats) {
  constexpr int D = V4 * 128;
  const int lane = threadIdx.x & 31;
  const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= static_cast<size_t>(n_img) * T) return;
  const int t = static_cast<int>(row % T);
  const size_t n = row / T;
  float v[V4 * 4];
#pragma unroll
  for (int j = 0; j < V4; ++j) {
    float4 a;
    if (has_cls && t == 0) {
      a = __ldg(reinterpret_cast<const float4*>(cls) + j * 32 + lane);
    } else {
      a = __ldg(reinterpret_cast<const float4*>(patch + (n * (T - has_cls) + (t - has_cls)) * D) + j * 32 + lane);
    }
    const float4 pe = __ldg(reinterpret_cast<const float4*>(pos + static_cast<size_t>(t) * D) + j * 32 + lane);
    v[j * 4 + 0] = a.x + pe.x;
    v[j * 4 + 1] = a.y + pe.y;
    v[j * 4 + 2] = a.z + pe.z;
    v[j * 4 + 3] = a.w + pe.w;
  }
  if (ln_w != nullptr) warp_layernorm<V4>(v, ln_w, ln_b, eps, lane);
#pragma unroll
  for (int j = 0; j < V4; ++j)
    reinterpret_cast<float4*>(x + row * D)[j * 32 + lane] =
        make_float4(v[j * 4 + 0], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
  if (xhat != nullptr) {
#pragma unroll
    for (int j = 0; j < V4; ++j) {  // vector j of every lane = columns [128 j, 128 j + 128): one statistics slot
      uint2 o;
      o.x = pack16x2<F16>(v[j * 4 + 0], v[j * 4 + 1]);
      o.y = pack16x2<F16>(v[j * 4 + 2], v[j * 4 + 3]);
      reinterpret_cast<uint2*>(xhat + row * D)[j * 32 + lane] = o;
      const float shift = __shfl_sync(0xffffffffu, v[j * 4], 0);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d = v[j * 4 + e] - shift;
        s1 += d;
        s2 = fmaf(d, d, s2);
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) reinterpret_cast<float4*>(stats)[row * V4 + j] = make_float4(shift, s1, s2, 0.f);
    }
  }
}

// Weight ingest for the folded LayerNorm (once per tower): W [N, K] fp32, gamma / beta [K], bias [N] or nullptr ->
//   W16[j, k] = 16-bit(W[j, k] * gamma[k]),   cw[j] = sum_k W16[j, k] (of the ROUNDED values: the epilogue subtracts
//   mu * cw from x^ W16^T, which must cancel exactly),   bw[j] = sum_k W[j, k] * beta[k] + bias[j].   One warp per row j.
template <bool F16>
__global__ void ln_fold_weights_kernel(const float* __restrict__ W, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ bias, int N, int K,
                                       uint16_t* __restrict__ W16, float* __restrict__ cw, float* __restrict__ bw) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= N) return;
  float sc = 0.f, sb = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = W[static_cast<size_t>(j) * K + k];
    const float wg = w * __ldg(gamma + k);
    float r;
    uint16_t h;
    if constexpr (F16) {
      const __half hv = __float2half_rn(wg);
      h = __half_as_ushort(hv);
      r = __half2float(hv);
    } else {
      const __nv_bfloat16 hv = __float2bfloat16(wg);
      h = __bfloat16_as_ushort(hv);
      r = __bfloat162float(hv);
    }
    W16[static_cast<size_t>(j) * K + k] = h;
    sc += r;
    sb = fmaf(w, __ldg(beta + k), sb);
  }
  sc = warp_sum(sc);
  sb = warp_sum(sb);
  if (lane == 0) {
    cw[j] = sc;
    bw[j] = sb + (bias ? bias[j] : 0.f);
  }
}

// y[row,:] (bf16) = LayerNorm(x[row,:]) * gamma + beta, fp32 statistics.
template <int V4, bool F16>
__global__ void layernorm_bf16_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, float eps, uint16_t* __restrict__ y,
                                      size_t rows) {
  constexpr int D = V4 * 128;
  const int lane = threadIdx.x & 31;
  const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[V4 * 4];
#pragma unroll
  for (int j = 0; j < V4; ++j) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + row * D) + j * 32 + lane);
    v[j * 4 + 0] = a.x;
    v[j * 4 + 1] = a.y;
    v[j * 4 + 2] = a.z;
    v[j * 4 + 3] = a.w;
  }
  warp_layernorm<V4>(v, gamma, beta, eps, lane);
#pragma unroll
  for (int j = 0; j < V4; ++j) {
    uint2 o;
    o.x = pack16x2<F16>(v[j * 4 + 0], v[j * 4 + 1]);
    o.y = pack16x2<F16>(v[j * 4 + 2], v[j * 4 + 3]);
    reinterpret_cast<uint2*>(y + row * D)[j * 32 + lane] = o;
  }
}

// out[n, p, col_off : col_off + D] = x[n, p + drop, :]  cast to dt; out row stride = out_ld elements
// (dual-tower concat writes both towers into one [N,P,D1+D2] tensor; reference llava_arch.py:278-285).
template <int V4>
__global__ void select_cast_kernel(const float* __restrict__ x, void* __restrict__ out, int dt, int n_img, int T,
                                   int drop, int out_ld, int col_off) {
  constexpr int D = V4 * 128;
  const int lane = threadIdx.x & 31;
  const int Pout = T - drop;
  const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= static_cast<size_t>(n_img) * Pout) return;
  const int pidx = static_cast<int>(row % Pout);
  const size_t n = row / Pout;
  const float* src = x + (n * T + pidx + drop) * D;
  const size_t obase = row * out_ld + col_off;
#pragma unroll
  for (int j = 0; j < V4; ++j) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + j * 32 + lane);
    const size_t o = obase + (j * 32 + lane) * 4;
    if (dt == DT_F32) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = a;
    } else if (dt == DT_BF16) {
      uint2 w;
      w.x = pack_bf16x2(a.x, a.y);
      w.y = pack_bf16x2(a.z, a.w);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + o) = w;
    } else {
      __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
      uint2 w;
      w.x = *reinterpret_cast<uint32_t*>(&h0);
      w.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out) + o) = w;
    }
  }
}

}  // namespace lovr
