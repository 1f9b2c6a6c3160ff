// C-score kernels (reference: C_score/utils/utils_correspondence.py:226-277,297-337,345-382 and
// C_score/pck_train.py:142-163): windowed soft-argmax over similarity rows, the keypoint-rows similarity
// (only the K source-keypoint rows of the P^2 x P^2 matrix are consumed by the reference's callers), and PCK.
#pragma once
#include "ptx.cuh"

namespace lovr {

constexpr int kMaxGrid = 64;   // P <= 64
constexpr int kMaxKps = 32;    // keypoints per pair (SPair-71k: <= 30)

// One warp per similarity row.  sim row r: P*P fp32 values (targets, row-major ty*P+tx), row stride ld.
//   mode 0: hard  — flow = (argmax % P, argmax / P)                                   (:372-373)
//   mode 1: soft  — window > 0: entries outside the clamped (2W+1)^2 window around the row argmax are set to
//                   ZERO (not -inf) before softmax((s - max)/beta) over ALL P*P targets (:305-327, :226-232);
//                   the P*P - |window| zero entries are folded in analytically: each contributes exp(-M/beta)
//                   to the denominator and exp(-M/beta)*lin(.) to the numerators.
//                   window == 0: plain soft-argmax over the whole row.
// flow[r] = ((E[x]+1)(P-1)/2, (E[y]+1)(P-1)/2) in patch units (:239-255, :264-265), x,y on linspace(-1,1,P).
// If stride > 0 the row is a keypoint row and out = clamp(flow,0,P-1)*stride + floor(stride/2)   (:369,:376-377).
__global__ void flow_rows_kernel(const float* __restrict__ sim, size_t ld, size_t rows, int P, int window, float beta,
                                 int mode, float stride, float* __restrict__ flow, int* __restrict__ argmax_out) {
  __shared__ float lin[kMaxGrid];
  if (threadIdx.x < P) lin[threadIdx.x] = static_cast<float>(-1.0 + threadIdx.x * (2.0 / (P - 1)));
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* s = sim + row * ld;
  const int PP = P * P;
  // ---- row argmax, first index on ties (torch.argmax)
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int t = lane; t < PP; t += 32) {
    const float v = __ldg(s + t);
    if (v > bv) { bv = v; bi = t; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (argmax_out != nullptr && lane == 0) argmax_out[row] = bi;
  float fx, fy;
  if (mode == 0) {
    fx = static_cast<float>(bi % P);
    fy = static_cast<float>(bi / P);
  } else {
    int y0 = 0, y1 = P - 1, x0 = 0, x1 = P - 1;
    if (window > 0) {
      const int my = bi / P, mx = bi % P;
      y0 = max(my - window, 0); y1 = min(my + window, P - 1);
      x0 = max(mx - window, 0); x1 = min(mx + window, P - 1);
    }
    const int wx = x1 - x0 + 1;
    const int nwin = wx * (y1 - y0 + 1);
    const int tail = PP - nwin;
    float M = -INFINITY;
    for (int i = lane; i < nwin; i += 32) M = fmaxf(M, __ldg(s + (y0 + i / wx) * P + x0 + i % wx));
    M = warp_max(M);
    if (tail > 0) M = fmaxf(M, 0.f);
    float sw = 0.f, sx = 0.f, sy = 0.f, lx = 0.f, ly = 0.f;
    for (int i = lane; i < nwin; i += 32) {
      const int y = y0 + i / wx, x = x0 + i % wx;
      const float w = expf((__ldg(s + y * P + x) - M) / beta);
      sw += w;
      sx += w * lin[x];
      sy += w * lin[y];
      lx += lin[x];
      ly += lin[y];
    }
    sw = warp_sum(sw); sx = warp_sum(sx); sy = warp_sum(sy); lx = warp_sum(lx); ly = warp_sum(ly);
    if (tail > 0) {
      // sum of lin over ALL positions is 0 by symmetry, so the masked-out positions sum to -(window sum)
      const float wt = expf((0.f - M) / beta);
      sw += wt * static_cast<float>(tail);
      sx -= wt * lx;
      sy -= wt * ly;
    }
    fx = (sx / sw + 1.f) * (P - 1) / 2.0f;
    fy = (sy / sw + 1.f) * (P - 1) / 2.0f;
  }
  if (lane == 0) {
    if (stride > 0.f) {
      const float half = floorf(stride / 2.0f);  // python `anno_stride // 2` on floats
      fx = fminf(fmaxf(fx, 0.f), static_cast<float>(P - 1)) * stride + half;
      fy = fminf(fmaxf(fy, 0.f), static_cast<float>(P - 1)) * stride + half;
    }
    flow[row * 2 + 0] = fx;
    flow[row * 2 + 1] = fy;
  }
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// Keypoint rows of the similarity matrix in plain fp32 (the arithmetic the reference's fp32 matmul performs):
//   out[b, k, t] = n(d1[b, idx[b,k], :]) . n(d2[b, t, :]),   n(x) = x / (||x|| + 1e-10)   (pck_train.py:24-29)
// grid (ceil(PP / rows_per_block), n_pairs); smem: K normalised source rows (K*D floats).
template <typename T>
__global__ void kp_sim_rows_kernel(const T* __restrict__ d1, const T* __restrict__ d2, const int* __restrict__ idx,
                                   int Kall, int k0, int K, int PP, int D, int rows_per_block, float* __restrict__ out) {
  // keypoints [k0, k0 + K) of the pair's Kall: the host splits K so that K * D floats fit in shared memory
  extern __shared__ float s1[];  // [K][D]
  __shared__ float s_inv[kMaxKps];
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const T* d1b = d1 + static_cast<size_t>(b) * PP * D;
  const T* d2b = d2 + static_cast<size_t>(b) * PP * D;
  for (int k = warp; k < K; k += nwarps) {
    const int r = idx[static_cast<size_t>(b) * Kall + k0 + k];
    float ss = 0.f;
    for (int e = lane; e < D; e += 32) {
      const float v = to_f32<T>(d1b[static_cast<size_t>(r) * D + e]);
      s1[k * D + e] = v;
      ss += v * v;
    }
    ss = warp_sum(ss);
    if (lane == 0) s_inv[k] = 1.0f / (sqrtf(ss) + 1e-10f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * D; i += blockDim.x) s1[i] *= s_inv[i / D];
  __syncthreads();
  const int t0 = blockIdx.x * rows_per_block;
  const int t1 = min(PP, t0 + rows_per_block);
  for (int t = t0 + warp; t < t1; t += nwarps) {
    float dots[kMaxKps];
#pragma unroll
    for (int k = 0; k < kMaxKps; ++k) dots[k] = 0.f;
    float ss = 0.f;
    for (int c = 0; c < D; c += 1024) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int e = c + lane + 32 * i;
        v[i] = (e < D) ? to_f32<T>(d2b[static_cast<size_t>(t) * D + e]) : 0.f;
        ss += v[i] * v[i];
      }
#pragma unroll
      for (int k = 0; k < kMaxKps; ++k) {
        if (k < K) {
          float a = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int e = c + lane + 32 * i;
            if (e < D) a += v[i] * s1[k * D + e];
          }
          dots[k] += a;
        }
      }
    }
    ss = warp_sum(ss);
    const float inv2 = 1.0f / (sqrtf(ss) + 1e-10f);
#pragma unroll
    for (int k = 0; k < kMaxKps; ++k) {
      if (k < K) {
        const float d = warp_sum(dots[k]);
        if (lane == 0) out[(static_cast<size_t>(b) * Kall + k0 + k) * PP + t] = d * inv2;
      }
    }
  }
}

// Same arithmetic, D % 128 == 0: each lane owns 4 consecutive channels per 128-channel chunk, so the target row comes in as
// 8- or 16-byte loads and the K source rows are read from shared memory as float4 (one LDS.128 per 4 FMAs instead of one
// LDS.32 per FMA: the scalar version above was shared-memory-issue-bound at 28 us per P = 60 pair, 25x its HBM roofline).
template <typename T, int NC>  // NC = D / 128 chunks
__global__ void kp_sim_rows4_kernel(const T* __restrict__ d1, const T* __restrict__ d2, const int* __restrict__ idx, int Kall,
                                    int k0, int K, int PP, int rows_per_block, float* __restrict__ out) {
  constexpr int D = NC * 128;
  extern __shared__ float s1[];  // [K][D]
  __shared__ float s_inv[kMaxKps];
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const T* d1b = d1 + static_cast<size_t>(b) * PP * D;
  const T* d2b = d2 + static_cast<size_t>(b) * PP * D;
  auto load4 = [&](const T* p) -> float4 {
    if constexpr (sizeof(T) == 4) {
      return __ldg(reinterpret_cast<const float4*>(p));
    } else {
      const uint2 w = __ldg(reinterpret_cast<const uint2*>(p));
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.x));
      const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.y));
      return make_float4(a.x, a.y, c.x, c.y);
    }
  };
  for (int k = warp; k < K; k += nwarps) {
    const int r = idx[static_cast<size_t>(b) * Kall + k0 + k];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float4 v = load4(d1b + static_cast<size_t>(r) * D + c * 128 + 4 * lane);
      *reinterpret_cast<float4*>(&s1[k * D + c * 128 + 4 * lane]) = v;
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
    if (lane == 0) s_inv[k] = 1.0f / (sqrtf(ss) + 1e-10f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * D; i += blockDim.x) s1[i] *= s_inv[i / D];
  __syncthreads();
  const int t0 = blockIdx.x * rows_per_block;
  const int t1 = min(PP, t0 + rows_per_block);
  // two target rows per pass: every float4 of a source row read from shared memory feeds 8 FMAs
  for (int t = t0 + 2 * warp; t < t1; t += 2 * nwarps) {
    const bool two = t + 1 < t1;
    const int tb = two ? t + 1 : t;
    float4 va[NC], vb[NC];
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      va[c] = load4(d2b + static_cast<size_t>(t) * D + c * 128 + 4 * lane);
      vb[c] = load4(d2b + static_cast<size_t>(tb) * D + c * 128 + 4 * lane);
      sa += va[c].x * va[c].x + va[c].y * va[c].y + va[c].z * va[c].z + va[c].w * va[c].w;
      sb += vb[c].x * vb[c].x + vb[c].y * vb[c].y + vb[c].z * vb[c].z + vb[c].w * vb[c].w;
    }
    sa = warp_sum(sa);
    sb = warp_sum(sb);
    const float inva = 1.0f / (sqrtf(sa) + 1e-10f), invb = 1.0f / (sqrtf(sb) + 1e-10f);
    for (int k = 0; k < K; ++k) {
      float a = 0.f, bsum = 0.f;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float4 w = *reinterpret_cast<const float4*>(&s1[k * D + c * 128 + 4 * lane]);
        a = fmaf(va[c].x, w.x, a);
        a = fmaf(va[c].y, w.y, a);
        a = fmaf(va[c].z, w.z, a);
        a = fmaf(va[c].w, w.w, a);
        bsum = fmaf(vb[c].x, w.x, bsum);
        bsum = fmaf(vb[c].y, w.y, bsum);
        bsum = fmaf(vb[c].z, w.z, bsum);
        bsum = fmaf(vb[c].w, w.w, bsum);
      }
      a = warp_sum(a);
      bsum = warp_sum(bsum);
      if (lane == 0) {
        out[(static_cast<size_t>(b) * Kall + k0 + k) * PP + t] = a * inva;
        if (two) out[(static_cast<size_t>(b) * Kall + k0 + k) * PP + t + 1] = bsum * invb;
      }
    }
  }
}

// Two-model descriptors (C_score/pck_train_two.py:24-35, the DUMMY_NET branch): out[r] = [ a[r] / (|a[r]| + 1e-10) , b[r] / (|b[r]| + 1e-10) ];
// the final renormalisation of that file is the one the similarity kernels fuse.  One warp per row, fp32 arithmetic.
template <typename T>
__global__ void concat_normalize_kernel(const T* __restrict__ a, const T* __restrict__ b, size_t rows, int Da, int Db,
                                        T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* src[2] = {a + row * Da, b + row * Db};
  const int dim[2] = {Da, Db};
  T* dst = out + row * (static_cast<size_t>(Da) + Db);
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    float ss = 0.f;
    for (int e = lane; e < dim[m]; e += 32) {
      const float v = to_f32<T>(src[m][e]);
      ss += v * v;
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss) + 1e-10f;
    for (int e = lane; e < dim[m]; e += 32) {
      const float v = to_f32<T>(src[m][e]) / nrm;
      if constexpr (sizeof(T) == 4) dst[e] = v;
      else dst[e] = __float2bfloat16(v);
    }
    dst += dim[m];
  }
}

// Per-pair PCK (pck_train.py:150-163): fraction of visible keypoints with ||pred-gt|| < alpha*thr,
// alpha in {0.1, 0.05, 0.01}.  out[b] = {pck@.1, pck@.05, pck@.01, #visible}.
__global__ void pck_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                           const unsigned char* __restrict__ vis, const float* __restrict__ thr, int n, int K,
                           float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  int nv = 0, c1 = 0, c2 = 0, c3 = 0;
  const float t = thr[b];
  for (int k = 0; k < K; ++k) {
    if (!vis[static_cast<size_t>(b) * K + k]) continue;
    const size_t o = (static_cast<size_t>(b) * K + k) * 2;
    const float dx = gt[o] - pred[o], dy = gt[o + 1] - pred[o + 1];
    const float err = sqrtf(dx * dx + dy * dy);
    ++nv;
    c1 += err < 0.1f * t;
    c2 += err < 0.05f * t;
    c3 += err < 0.01f * t;
  }
  const float d = nv > 0 ? static_cast<float>(nv) : 1.f;
  out[b * 4 + 0] = c1 / d;
  out[b * 4 + 1] = c2 / d;
  out[b * 4 + 2] = c3 / d;
  out[b * 4 + 3] = static_cast<float>(nv);
}

// ---------------------------------------------------------------------------------------------------------------
// Mutual-nearest-neighbour distance of the adaptive-flip test (C_score/utils/utils_correspondence.py:55-73
// get_distance_mutual_nn; caller pck_train.py:111-126).  The reference builds torch.cdist(n(d1), n(d2)) [P^2,P^2]; with
// the cosine matrix sim = n(d1) n(d2)^T already produced by the tcgen05 GEMM, dist^2(i,j) = |n1_i|^2 + |n2_j|^2 - 2 sim(i,j)
// (cdist's own matmul form), |n|^2 = (1 - 1e-10 * inv)^2 from the stored inv = 1 / (|d| + 1e-10).
//   mnn_rows_kernel   warp per source row: min_j dist^2 and its first index
//   mnn_cols_kernel   thread per target column: first index of min_i dist^2 (coalesced across the warp)
//   mnn_reduce_kernel block per pair: mean over i with col_arg[row_arg[i]] == i of sqrt(max(dist^2, 0)); NaN when there is
//                     no mutual pair (torch: mean of an empty tensor)
__device__ __forceinline__ float unit_norm_sq(float inv) {
  const float q = 1.0f - 1e-10f * inv;
  return q * q;
}

__global__ void mnn_rows_kernel(const float* __restrict__ sim, size_t ld, const float* __restrict__ inv1,
                                const float* __restrict__ inv2, int scale_ld, int PP, size_t rows,
                                float* __restrict__ row_d2, int* __restrict__ row_arg) {
  const int lane = threadIdx.x & 31;
  const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const size_t b = row / PP;
  const int i = static_cast<int>(row % PP);
  const float* s = sim + row * ld;
  const float* i2 = inv2 + b * scale_ld;
  const float n1 = unit_norm_sq(__ldg(inv1 + b * scale_ld + i));
  float bv = INFINITY;
  int bi = 0x7fffffff;
  for (int t = lane; t < PP; t += 32) {
    const float v = n1 + unit_norm_sq(__ldg(i2 + t)) - 2.0f * __ldg(s + t);
    if (v < bv) { bv = v; bi = t; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) {
    row_d2[row] = bv;
    row_arg[row] = bi;
  }
}

// grid (ceil(PP / blockDim.x), row slabs, pairs): each thread scans one slab of rows of its column; slabs are merged
// (first index on ties) by mnn_reduce_kernel.
constexpr int kMnnSlabs = 8;
__global__ void mnn_cols_kernel(const float* __restrict__ sim, size_t ld, const float* __restrict__ inv1,
                                const float* __restrict__ inv2, int scale_ld, int PP, float* __restrict__ col_d2,
                                int* __restrict__ col_arg) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int slab = blockIdx.y;
  const size_t b = blockIdx.z;
  if (j >= PP) return;
  const int per = (PP + kMnnSlabs - 1) / kMnnSlabs;
  const int i0 = slab * per, i1 = min(PP, i0 + per);
  const float n2 = unit_norm_sq(__ldg(inv2 + b * scale_ld + j));
  const float* s = sim + (b * PP) * ld + j;
  const float* i1p = inv1 + b * scale_ld;
  float bv = INFINITY;
  int bi = 0x7fffffff;
  for (int i = i0; i < i1; ++i) {
    const float v = unit_norm_sq(__ldg(i1p + i)) + n2 - 2.0f * __ldg(s + static_cast<size_t>(i) * ld);
    if (v < bv) { bv = v; bi = i; }
  }
  const size_t o = (b * kMnnSlabs + slab) * PP + j;
  col_d2[o] = bv;
  col_arg[o] = bi;
}

__global__ void mnn_reduce_kernel(const float* __restrict__ row_d2, const int* __restrict__ row_arg,
                                  const float* __restrict__ col_d2, const int* __restrict__ col_arg, int PP,
                                  float* __restrict__ out) {
  const size_t b = blockIdx.x;
  __shared__ float s_sum[32];
  __shared__ int s_cnt[32];
  float sum = 0.f;
  int cnt = 0;
  for (int i = threadIdx.x; i < PP; i += blockDim.x) {
    const int j = row_arg[b * PP + i];
    float bv = INFINITY;
    int bi = 0x7fffffff;
    for (int sl = 0; sl < kMnnSlabs; ++sl) {  // slabs are in ascending row order: strict < keeps the first index
      const size_t o = (b * kMnnSlabs + sl) * PP + j;
      const float v = col_d2[o];
      if (v < bv) { bv = v; bi = col_arg[o]; }
    }
    if (bi == i) {
      sum += sqrtf(fmaxf(row_d2[b * PP + i], 0.f));
      ++cnt;
    }
  }
  sum = warp_sum(sum);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_sum[warp] = sum; s_cnt[warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    int c = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { t += s_sum[w]; c += s_cnt[w]; }
    out[b * 2 + 0] = t / static_cast<float>(c);  // 0/0 = NaN like torch's mean of an empty selection
    out[b * 2 + 1] = static_cast<float>(c);
  }
}

}  // namespace lovr
