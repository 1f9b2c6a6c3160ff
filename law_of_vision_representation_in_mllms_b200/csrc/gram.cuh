// Gram sums for linear CKA on tcgen05, straight from the row-major features (no transposed copy):
//     G = Z^T Z,  Z = [X | Y]  ([R, Dx + Dy] bf16, R = all N*P tokens),   s = column sums of Z
// (metric named by BASELINE.json north_star; builder-defined — the reference has no CKA).
//
// Both MMA operands are MN-major views of Z itself: a TMA box of 64 rows x 64 features (128-byte swizzle) IS the canonical
// MN-major UMMA tile (64 MN-contiguous elements per 128-B row, 8-row groups SBO = 1024 B apart, 64-feature groups LBO = 8 KB
// apart, +2 KB per K = 16 step), so X^T never exists in memory (round 1 wrote a transposed bf16 copy first: one extra read
// and write of all the data) and the reduction dimension K is the row index.
//
// G is symmetric: only tiles (i: 128 features, j: 256 features) with 256 (j + 1) > 128 i — the upper triangle of the whole
// Z-Gram, i.e. X^T X (upper), X^T Y, Y^T Y (upper) — are computed: 72 of 128 tiles at Dx = Dy = 1024 (44 % fewer FLOPs).
// Unit = (row range, tile).  With n_ranges = floor(#SM / n_tiles) the whole job is ONE wave of persistent CTAs that all stream
// the same rows at the same time, so every row window is read from HBM once and served to the other tiles' CTAs by L2.
//
// Accuracy: the tensor core's fp32 accumulation truncates (error ~ run length; 1e-4 on CKA at 16384 rows), so a TMEM
// accumulation run is kChunk = 2048 rows; finished runs are added (fp32, round-to-nearest) into REGISTER accumulators of
// the eight epilogue warps (128 values per thread), which are written out once per unit as fp32 partials — 144 x 128 KB
// instead of round 1's one partial per 2048 rows (1.5 GB written and re-read per 262k rows).  Partials and column sums are
// reduced in fp64.
//   warp 0        TMA producer: per 64-row k-block 2 boxes of the A features and 4 boxes of the B features (48 KB stage)
//   warp 1        MMA issuer: 128 x 256 x 16, A and B MN-major, two TMEM accumulators (ping-pong between runs)
//   warps 4..7    column sums of the A features (only in the unit that owns them: j == i / 2), fp32 per run, fp64 atomics
//   warps 8..15   epilogue (warp % 4 = TMEM lane quarter, (warp - 8) / 4 = column half)
// setmaxnreg: 56 registers for warps 0..7, 200 for the epilogue warps.
#pragma once
#include "gemm.cuh"

namespace lovr {

constexpr int kGramChunk = 2048;
constexpr int kGramStages = 4;
constexpr int kGramThreads = 16 * 32;
constexpr int kGramABytes = 128 * 64 * 2;
constexpr int kGramStageBytes = (128 + 256) * 64 * 2;
constexpr int kGramSmem = kGramStages * kGramStageBytes + 256 + 1024;
constexpr int kGramTileElems = 128 * 256;

struct GramArgs {
  long long R;
  int Dx, Dy;
  int n_tiles, n_ranges, total_chunks;
  float* partial;   // [n_ranges * n_tiles][128][256]
  double* colsum;   // [Dx + Dy], accumulated with atomics
};

__host__ __device__ inline int gram_tile_count(int Dz) {
  const int ni = Dz / 128, nj = Dz / 256;
  int n = 0;
  for (int i = 0; i < ni; ++i) n += nj - (i >> 1);
  return n;
}
// tile index -> (i, j): row block of 128 features, column block of 256 features, j >= i / 2
__host__ __device__ inline void gram_tile_ij(int t, int Dz, int* i_out, int* j_out) {
  const int nj = Dz / 256;
  int i = 0;
  for (;;) {
    const int cnt = nj - (i >> 1);
    if (t < cnt) break;
    t -= cnt;
    ++i;
  }
  *i_out = i;
  *j_out = (i >> 1) + t;
}

__global__ void __launch_bounds__(kGramThreads, 1)
gram_tcgen05_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const GramArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kGramStages * kGramStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kGramStages;
  uint64_t* tmem_full = bars + 2 * kGramStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Dz = p.Dx + p.Dy;
  const int units = p.n_tiles * p.n_ranges;
  constexpr int kKb = kGramChunk / 64;  // k-blocks per accumulation run

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
    for (int i = 0; i < kGramStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1 + 4);  // tcgen05.commit + the four column-sum warps
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // chunk range of unit u
  auto unit_chunks = [&](int u, int* c0, int* c1, int* i, int* j) {
    const int range = u / p.n_tiles;
    gram_tile_ij(u % p.n_tiles, Dz, i, j);
    *c0 = static_cast<int>(static_cast<long long>(p.total_chunks) * range / p.n_ranges);
    *c1 = static_cast<int>(static_cast<long long>(p.total_chunks) * (range + 1) / p.n_ranges);
  };

  if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");      // 8 x 32 x 56 + 8 x 32 x 200 = 65536
    if (warp == 0) {
      // ===================================================================== TMA producer
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
          int c0, c1, ti, tj;
          unit_chunks(u, &c0, &c1, &ti, &tj);
          const int acol = ti * 128, bcol = tj * 256;
          const CUtensorMap* ma = acol < p.Dx ? &tmX : &tmY;
          const CUtensorMap* mb = bcol < p.Dx ? &tmX : &tmY;
          const int ac = acol < p.Dx ? acol : acol - p.Dx, bc = bcol < p.Dx ? bcol : bcol - p.Dx;
          for (int c = c0; c < c1; ++c)
            for (int kb = 0; kb < kKb; ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = stage_base + stage * kGramStageBytes;
              const int row = c * kGramChunk + kb * 64;
              mbar_arrive_expect_tx(&full_bar[stage], kGramStageBytes);
#pragma unroll
              for (int g = 0; g < 2; ++g) tma_load_3d(sa + g * 8192, ma, &full_bar[stage], ac + g * 64, row, 0);
#pragma unroll
              for (int g = 0; g < 4; ++g) tma_load_3d(sa + kGramABytes + g * 8192, mb, &full_bar[stage], bc + g * 64, row, 0);
              if (++stage == kGramStages) { stage = 0; phase ^= 1; }
            }
        }
      }
    } else if (warp == 1) {
      // ===================================================================== MMA issuer
      constexpr uint32_t kDescHi = smem_desc_hi_sw128(1024);               // SBO: 8-row (K) groups 1 KB apart
      constexpr uint32_t kIdesc = make_idesc_16(128, 256, 1, 1, false);    // A and B MN-major, bf16, fp32 accumulate
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(stage_base), 8192);     // LBO: 64-feature groups 8 KB apart
      const uint32_t b_lo0 = smem_desc_lo(smem_u32(stage_base) + kGramABytes, 8192);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int c0, c1, ti, tj;
        unit_chunks(u, &c0, &c1, &ti, &tj);
        for (int c = c0; c < c1; ++c) {
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * 256;
          for (int kb = 0; kb < kKb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (lane == 0) {
              const uint32_t a_lo = a_lo0 + stage * (kGramStageBytes >> 4);
              const uint32_t b_lo = b_lo0 + stage * (kGramStageBytes >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)   // K = 16 rows = 2 KB further into every 64-feature group
                umma_bf16_ss(d_tmem, smem_desc_join(kDescHi, a_lo + k * (2048 >> 4)), smem_desc_join(kDescHi, b_lo + k * (2048 >> 4)),
                             kIdesc, (kb | k) != 0 ? 1u : 0u);
              umma_commit(&empty_bar[stage]);
              if (kb == kKb - 1) umma_commit(&tmem_full[acc]);
            }
            __syncwarp();
            if (++stage == kGramStages) { stage = 0; phase ^= 1; }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    } else if (warp >= 4) {
      // ===================================================================== column sums of the A features
      const int f = threadIdx.x - 4 * 32;       // feature within the A tile
      const uint32_t col_off = (f >> 6) * 8192 + (f & 7) * 2;
      const uint32_t chunk16 = (f & 63) >> 3;   // 16-byte chunk of the 128-byte row before swizzling
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int c0, c1, ti, tj;
        unit_chunks(u, &c0, &c1, &ti, &tj);
        const bool owner = tj == (ti >> 1);
        for (int c = c0; c < c1; ++c) {
          float s = 0.f;
          for (int kb = 0; kb < kKb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            if (owner) {
              const uint32_t base = smem_u32(stage_base + stage * kGramStageBytes) + col_off;
#pragma unroll 8
              for (int r = 0; r < 64; ++r) {
                uint16_t h;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(base + r * 128 + ((chunk16 ^ (r & 7)) << 4)));
                s += __uint_as_float(static_cast<uint32_t>(h) << 16);
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == kGramStages) { stage = 0; phase ^= 1; }
          }
          if (owner) atomicAdd(&p.colsum[ti * 128 + f], static_cast<double>(s));
        }
      }
    }
  } else {
    // ===================================================================== epilogue: runs -> register accumulators -> partials
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    const int quarter = warp & 3, half = (warp - 8) >> 2;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + half * 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int c0, c1, ti, tj;
      unit_chunks(u, &c0, &c1, &ti, &tj);
      float sum[128];
#pragma unroll
      for (int e = 0; e < 128; ++e) sum[e] = 0.f;
      for (int c = c0; c < c1; ++c) {
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t r0[32];
          tmem_ld_32x32b_x32(t_lane + acc * 256 + q * 32, r0);
          tmem_ld_wait_dep(r0);
#pragma unroll
          for (int e = 0; e < 32; ++e) sum[q * 32 + e] += __uint_as_float(r0[e]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      float4* dst = reinterpret_cast<float4*>(p.partial + (static_cast<size_t>(u) * 128 + quarter * 32 + lane) * 256 + half * 128);
#pragma unroll
      for (int e = 0; e < 32; ++e) dst[e] = make_float4(sum[4 * e], sum[4 * e + 1], sum[4 * e + 2], sum[4 * e + 3]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// gram (fp64, layout of include/lovr_b200.h: Ga [Dx, Dx+Dy] | Gb [Dy, Dy]) += sum over ranges of the fp32 partials; the
// upper-triangle entries are mirrored inside X^T X and Y^T Y so the buffer stays a complete Gram.  Thread = tile entry.
__global__ void gram_reduce_kernel(const float* __restrict__ partial, int n_tiles, int n_ranges, int Dx, int Dy,
                                   double* __restrict__ Ga, double* __restrict__ Gb, double* __restrict__ count, double add_rows) {
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx == 0) *count += add_rows;  // the shard's row count rides in the Gram buffer
  if (idx >= static_cast<size_t>(n_tiles) * kGramTileElems) return;
  const int t = static_cast<int>(idx / kGramTileElems), e = static_cast<int>(idx % kGramTileElems);
  const int Dz = Dx + Dy;
  int ti, tj;
  gram_tile_ij(t, Dz, &ti, &tj);
  const int pz = ti * 128 + e / 256, qz = tj * 256 + e % 256;
  if (qz < pz) return;  // lower-triangle part of a tile that straddles the diagonal: its mirror is computed elsewhere
  double s = 0.0;
  for (int r = 0; r < n_ranges; ++r) s += static_cast<double>(partial[(static_cast<size_t>(r) * n_tiles + t) * kGramTileElems + e]);
  if (pz < Dx) {
    Ga[static_cast<size_t>(pz) * Dz + qz] += s;
    if (qz < Dx && qz != pz) Ga[static_cast<size_t>(qz) * Dz + pz] += s;
  } else {
    const int a = pz - Dx, b = qz - Dx;
    Gb[static_cast<size_t>(a) * Dy + b] += s;
    if (a != b) Gb[static_cast<size_t>(b) * Dy + a] += s;
  }
}

}  // namespace lovr
