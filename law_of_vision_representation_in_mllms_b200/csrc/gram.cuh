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
//   warp 0        TMA producer: per 64-row k-block one 3-D box {64 features, 64 rows, 2 groups} for A and one with 4 groups
//                 for B (the features are addressed as [group of 64][row][64]): 48 KB per stage
//   warp 1        MMA issuer: 128 x 256 x 16, A and B MN-major, two TMEM accumulators (ping-pong between runs)
//   warps 4..7    column sums of the A features (only in the unit that owns them: j == i / 2), fp32 per run, fp64 atomics
//   warps 8..15   epilogue (warp % 4 = TMEM lane quarter, (warp - 8) / 4 = column half)
// setmaxnreg: 56 registers for warps 0..7, 200 for the epilogue warps.
#pragma once
#include "gemm.cuh"

namespace lovr {

constexpr int kGramChunk = 2048;
constexpr int kGramStages = 4;
constexpr int kGramPrefetch = 0;     // k-blocks of L2 prefetch ahead of the TMA loads: measured slower (0: 5.07 ms, 4: 5.86, 12: 6.06 at 1.2M rows)
constexpr long long kGramPaceTimeoutCycles = 2000000;
constexpr int kGramPaceWindow = 2;   // runs (2048 rows) a CTA may lead the slowest CTA of its range by
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
  int debug;        // LOVR_GRAM_DEBUG: 1 = print per-unit cycle counts, 2 / 4 = timing experiments (see below)
  int prefetch;     // k-blocks of L2 prefetch ahead of the TMA loads
  int* pace;        // [n_ranges][chunks per range] arrival counters of the pacing window, zeroed per launch, or nullptr
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
gram_tcgen05_kernel(const __grid_constant__ CUtensorMap tmXa, const __grid_constant__ CUtensorMap tmXb,
                    const __grid_constant__ CUtensorMap tmYa, const __grid_constant__ CUtensorMap tmYb, const GramArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kGramStages * kGramStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kGramStages;
  uint64_t* tmem_full = bars + 2 * kGramStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
  const int Dz = p.Dx + p.Dy;
  const int units = p.n_tiles * p.n_ranges;
  constexpr int kKb = kGramChunk / 64;  // k-blocks per accumulation run

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmXa);
    tma_prefetch_desc(&tmXb);
    tma_prefetch_desc(&tmYa);
    tma_prefetch_desc(&tmYb);
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
  const uint32_t tmem_base = warp_uniform(*tmem_slot);

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
          const CUtensorMap* ma = acol < p.Dx ? &tmXa : &tmYa;      // box: 2 groups of 64 features
          const CUtensorMap* mb = bcol < p.Dx ? &tmXb : &tmYb;      // box: 4 groups
          const int ag = (acol < p.Dx ? acol : acol - p.Dx) / 64, bg = (bcol < p.Dx ? bcol : bcol - p.Dx) / 64;
          const int row_end = c1 * kGramChunk;
          // Pacing window.  The CTAs of a range stream the same rows, and each row window is meant to come from HBM once and
          // from L2 for the other ~70 CTAs.  Nothing holds the pack together: over the 1400 runs of a C3-sized job the CTAs
          // drift apart by more rows than L2 holds and every tile ends up streaming its own copy from HBM (measured: 2.3x
          // slower than the same kernel on a fifth of the rows).  So a CTA may start run lc only when every CTA of its range
          // has issued run lc - kGramPaceWindow: one atomic and one polled load per 2048 rows.  (Needs one unit per CTA, all
          // CTAs resident: the host enables it only then.)
          int* pace = p.pace ? p.pace + (u / p.n_tiles) * ((p.total_chunks + p.n_ranges - 1) / p.n_ranges + 1) : nullptr;
          for (int c = c0; c < c1; ++c) {
            if (pace != nullptr && c - c0 >= kGramPaceWindow) {
              const int* slot = pace + (c - c0 - kGramPaceWindow);
              int seen;
              const long long t0 = clock64();
              for (;;) {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(slot) : "memory");
                if (seen >= p.n_tiles) break;
                __nanosleep(200);
                if (clock64() - t0 > kGramPaceTimeoutCycles) {
                  // pacing is a hint, never a dependency: if the rest of the range does not show up in ~1 ms (CTAs not
                  // co-resident because something else holds SMs), this CTA stops pacing instead of waiting
                  pace = nullptr;
                  break;
                }
              }
            }
            for (int kb = 0; kb < kKb; ++kb) {
              const int row = c * kGramChunk + kb * 64;
              // the CTAs that lead the pack touch rows nobody has read yet: ask L2 for them kGramPrefetch k-blocks ahead
              const int prow = row + p.prefetch * 64;
              if (p.prefetch > 0 && prow < row_end) {
                tma_prefetch_3d(ma, 0, prow, ag);
                tma_prefetch_3d(mb, 0, prow, bg);
              }
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = stage_base + stage * kGramStageBytes;
              if (p.debug & 2) {  // timing experiment: no loads, the MMA runs on whatever is in shared memory
                mbar_arrive(&full_bar[stage]);
                if (++stage == kGramStages) { stage = 0; phase ^= 1; }
                continue;
              }
              mbar_arrive_expect_tx(&full_bar[stage], kGramStageBytes);
              tma_load_3d(sa, ma, &full_bar[stage], 0, row, ag);
              tma_load_3d(sa + kGramABytes, mb, &full_bar[stage], 0, row, bg);
              if (++stage == kGramStages) { stage = 0; phase ^= 1; }
            }
            if (p.pace != nullptr)   // (arrivals continue even if this CTA gave up waiting: the others still count on them)
              atomicAdd(p.pace + (u / p.n_tiles) * ((p.total_chunks + p.n_ranges - 1) / p.n_ranges + 1) + (c - c0), 1);
          }
        }
      }
    } else if (warp == 1) {
      // ===================================================================== MMA issuer
      constexpr uint32_t kDescHi = smem_desc_hi_sw128(1024);               // SBO: 8-row (K) groups 1 KB apart
      // (debug & 4: timing experiment with K-major descriptors on the same bytes — wrong results, same traffic)
      const uint32_t kIdesc = (p.debug & 4) ? make_idesc_16(128, 256, 0, 0, false) : make_idesc_16(128, 256, 1, 1, false);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(stage_base), (p.debug & 4) ? 16 : 8192);     // LBO: 64-feature groups 8 KB apart
      const uint32_t b_lo0 = smem_desc_lo(smem_u32(stage_base) + kGramABytes, (p.debug & 4) ? 16 : 8192);
      const uint32_t kstep = (p.debug & 4) ? 2 : (2048 >> 4);
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
                umma_bf16_ss(d_tmem, smem_desc_join(kDescHi, a_lo + k * kstep), smem_desc_join(kDescHi, b_lo + k * kstep),
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
      // thread = (group of 8 features, group of 8 rows): one LDS.128 per row brings 8 features; the 8 threads of an LDS
      // phase read the 8 different 16-byte chunks of one row (conflict-free under the 128-byte swizzle)
      const int t = threadIdx.x - 4 * 32;
      const int fg = t & 15, rg = t >> 4;
      const uint32_t off = (fg >> 3) * 8192 + rg * 1024;
      const uint32_t chunk16 = fg & 7;
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int c0, c1, ti, tj;
        unit_chunks(u, &c0, &c1, &ti, &tj);
        const bool owner = tj == (ti >> 1);
        for (int c = c0; c < c1; ++c) {
          float s8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) s8[e] = 0.f;
          for (int kb = 0; kb < kKb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            if (owner) {
              const uint32_t base = smem_u32(stage_base + stage * kGramStageBytes) + off;
#pragma unroll
              for (int r = 0; r < 8; ++r) {
                uint32_t w0, w1, w2, w3;
                ld_shared_v4(base + r * 128 + ((chunk16 ^ r) << 4), w0, w1, w2, w3);
                const uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  s8[2 * q] += __uint_as_float(w[q] << 16);
                  s8[2 * q + 1] += __uint_as_float(w[q] & 0xffff0000u);
                }
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == kGramStages) { stage = 0; phase ^= 1; }
          }
          if (owner) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&p.colsum[ti * 128 + fg * 8 + e], static_cast<double>(s8[e]));
          }
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
      const long long t_start = clock64();
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
      if ((p.debug & 1) && threadIdx.x == 8 * 32) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        printf("gram unit %d tile (%d,%d) chunks %d cycles %lld sm %u\n", u, ti, tj, c1 - c0, clock64() - t_start, smid);
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
