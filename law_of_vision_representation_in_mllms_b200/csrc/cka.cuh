// Linear CKA support kernels (metric named by BASELINE.json north_star; builder-defined — the reference has no CKA,
// SURVEY.md §0 F1).  The Gram matrices X^T X, X^T Y, Y^T Y reduce over the ROW dimension, so the tensor-core GEMM
// (K-major operands) consumes a transposed bf16 copy Zt = [X | Y]^T whose rows are features and whose contiguous
// dimension is the sample index; the row range is split into chunks that become the GEMM batch dimension.
#pragma once
#include "ptx.cuh"

namespace lovr {

constexpr int kCkaRowsPerBlock = 1024;

// X [R, D] bf16 row-major -> Zt[d_off + d, r] (row stride Rpad, zero for r >= R), colsum[d_off + d] += sum_r X[r, d].
// grid (D/64, Rpad/1024), block 256.
__global__ void __launch_bounds__(256)
transpose_colsum_kernel(const uint16_t* __restrict__ X, size_t R, int D, uint16_t* __restrict__ Zt, size_t Rpad, int d_off,
                        double* __restrict__ colsum) {
  __shared__ __align__(16) uint16_t tile[64][72];
  const int tid = threadIdx.x;
  const int d0 = blockIdx.x * 64;
  const size_t rbase = static_cast<size_t>(blockIdx.y) * kCkaRowsPerBlock;
  double acc = 0.0;
  for (int sub = 0; sub < kCkaRowsPerBlock / 64; ++sub) {
    const size_t r0 = rbase + static_cast<size_t>(sub) * 64;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int row = pass * 32 + (tid >> 3), ch = tid & 7;
      const size_t r = r0 + row;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (r < R) v = __ldg(reinterpret_cast<const uint4*>(X + r * D + d0 + ch * 8));
      *reinterpret_cast<uint4*>(&tile[row][ch * 8]) = v;
    }
    __syncthreads();
    if (tid < 64) {
      float s = 0.f;
#pragma unroll 8
      for (int row = 0; row < 64; ++row) s += __bfloat162float(__ushort_as_bfloat16(tile[row][tid]));
      acc += static_cast<double>(s);
    }
    {
      const int d = tid >> 2, q = tid & 3;
      uint32_t w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        w[i] = static_cast<uint32_t>(tile[q * 16 + 2 * i][d]) | (static_cast<uint32_t>(tile[q * 16 + 2 * i + 1][d]) << 16);
      uint4* dst = reinterpret_cast<uint4*>(Zt + (static_cast<size_t>(d_off + d0 + d)) * Rpad + r0 + q * 16);
      dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
      dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
    __syncthreads();
  }
  if (tid < 64) atomicAdd(&colsum[d_off + d0 + tid], acc);
}

// out[i] (fp64) += sum_b partial[b * stride + i]; the accumulate form lets several calls (shards) add up.
// count (optional): *count += add_rows, by one thread — the shard's row count rides in the Gram buffer.
__global__ void reduce_partials_kernel(const float* __restrict__ partial, size_t stride, int splits, size_t n,
                                       double* __restrict__ out, double* __restrict__ count, double add_rows) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i == 0 && count != nullptr) *count += add_rows;
  if (i >= n) return;
  double s = 0.0;
  for (int b = 0; b < splits; ++b) s += static_cast<double>(partial[static_cast<size_t>(b) * stride + i]);
  out[i] += s;
}

// gram layout: Ga [Dx, Dx+Dy] (rows of X against [X|Y]), Gb [Dy, Dy], colsum [Dx+Dy]; all fp64, UNcentred sums over n rows.
// out[0] = ||Xc^T Yc||_F^2, out[1] = ||Xc^T Xc||_F^2, out[2] = ||Yc^T Yc||_F^2 with the centring applied as
// G - s s^T / n.
__global__ void cka_finalize_kernel(const double* __restrict__ Ga, const double* __restrict__ Gb,
                                    const double* __restrict__ s, double n_arg, const double* __restrict__ n_buf, int Dx,
                                    int Dy, double* __restrict__ out) {
  __shared__ double part[3][32];
  const double n = n_arg > 0.0 ? n_arg : *n_buf;
  const int Dz = Dx + Dy;
  const size_t na = static_cast<size_t>(Dx) * Dz, nb = static_cast<size_t>(Dy) * Dy;
  double a_xy = 0, a_xx = 0, a_yy = 0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < na + nb;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    if (i < na) {
      const int r = static_cast<int>(i / Dz), c = static_cast<int>(i % Dz);
      const double v = Ga[i] - s[r] * s[c] / n;
      if (c < Dx) a_xx += v * v; else a_xy += v * v;
    } else {
      const size_t j = i - na;
      const int r = static_cast<int>(j / Dy), c = static_cast<int>(j % Dy);
      const double v = Gb[j] - s[Dx + r] * s[Dx + c] / n;
      a_yy += v * v;
    }
  }
  a_xy = warp_sum(a_xy); a_xx = warp_sum(a_xx); a_yy = warp_sum(a_yy);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { part[0][w] = a_xy; part[1][w] = a_xx; part[2][w] = a_yy; }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    double x = l < nw ? part[0][l] : 0, y = l < nw ? part[1][l] : 0, z = l < nw ? part[2][l] : 0;
    x = warp_sum(x); y = warp_sum(y); z = warp_sum(z);
    if (l == 0) { atomicAdd(&out[0], x); atomicAdd(&out[1], y); atomicAdd(&out[2], z); }
  }
}

}  // namespace lovr
