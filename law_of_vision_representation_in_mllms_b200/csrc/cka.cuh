// Linear CKA finalisation (metric named by BASELINE.json north_star; builder-defined — the reference has no CKA,
// SURVEY.md §0 F1).  The Gram sums themselves come from gram.cuh (tcgen05, MN-major operands straight from the features).
#pragma once
#include "ptx.cuh"

namespace lovr {

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
