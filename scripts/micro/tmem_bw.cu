// Micro-benchmark: TMEM -> register bandwidth of tcgen05.ld.32x32b on sm_100a (one CTA per SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cuda_runtime.h>
#include "../../law_of_vision_representation_in_mllms_b200/csrc/ptx.cuh"
using namespace lovr;

template <int WAIT_EVERY>
__global__ void tmem_read_kernel(int iters, unsigned long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(base + ((i * 32) & 255) + ((warp >> 2) * 256), r);
    if ((i % WAIT_EVERY) == WAIT_EVERY - 1) tmem_ld_wait_dep(r);
    acc ^= r[i & 31];
  }
  tmem_ld_wait();
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  unsigned long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
  const int iters = 4096;
  for (int warps : {1, 2, 4, 8}) {
    for (int we : {1, 4}) {
      if (we == 1) tmem_read_kernel<1><<<148, warps * 32>>>(iters, cyc, sink);
      else tmem_read_kernel<4><<<148, warps * 32>>>(iters, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      unsigned long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      double c = h[0];
      double bytes = double(iters) * warps * 32 * 32 * 4;
      printf("warps=%d wait_every=%d: %s cycles=%.0f bytes/cycle/SM=%.1f cycles_per_x32_ld=%.1f\n", warps, we,
             cudaGetErrorString(e), c, bytes / c, c / iters);
    }
  }
  return 0;
}
