// A score (A_score/compute.py:12-15,54-72): per pair b
//     rowmax[b, i] = max_j cos(other[b, i, :], ref[b, j, :])
// as a batched tcgen05 GEMM  other[b] · ref[b]^T  with the inverse norms of the `other` rows computed from the SAME
// shared-memory tiles the tensor core reads: `other` is read from HBM once.  (Round 1 ran a norm pass over both inputs
// first: 10.4 ms at C3.  A version of this kernel that also took the `ref` norms from the tiles had to give each CTA a whole
// pair — 148 pairs x 2.4 MB in flight, more than L2 holds, so the tile re-reads went to HBM: 12.3 ms, measured.  The `ref`
// norms therefore stay a small pre-pass over `ref` alone and the m tiles of a pair run on neighbouring CTAs, which share the
// pair's tiles through L2.)
//
// Unit = (pair b, 128-row tile mt of `other`), units dealt round-robin to persistent CTAs; per unit every BN-row tile of `ref`
// (nt) streams K/64 k-blocks through a 4-stage TMA ring.
//   warp 0        TMA producer (3-D maps: k, row, pair; rows past x / y are zero-filled)
//   warp 1        MMA issuer, 128 x BN x 16 per instruction, fp32 accumulators double-buffered in TMEM
//   warps 2..5    epilogue: tcgen05.ld (thread = row of `other`), acc * inv_ref[b, col], running row maximum in a register
//                 across the n tiles, * inv_other[row] at the end of the unit
//   warps 6..9    norm warps: wait on the same `full` barrier as the MMA issuer and, while nt == 0, read the A half of the
//                 stage with swizzle-aware, conflict-free LDS.128 (thread = tile row): 1/n_tiles of the stages, 16 KB each
// Barriers: full[s] (TMA -> MMA + norm), empty[s] (count 1 + 4: tcgen05.commit + one arrive per norm warp),
// tmem_full/tmem_empty[acc] as in gemm.cuh, aux_full[acc] (count 4): the norm warps are done with tile (mt, nt) — they
// arrive once per tile, so its phase follows the accumulator's; the row scales of a unit are published with its first tile,
// in a slot that alternates with the unit's parity.  Needs K/64 > stages (4, or 5 for BN = 192) (the phase argument: the norm warps cannot run a
// whole tile ahead of the MMA issuer).
#pragma once
#include "gemm.cuh"
#include "scores.cuh"

namespace lovr {

constexpr int kAsThreads = 64 + 4 * 32 + 4 * 32;


struct AScoreArgs {
  int x, y, K, n_pairs;
  const float* inv_ref;  // [n_pairs, y] 1 / ||ref|| with the reference's epsilons (rownorm_inv_bf16_kernel)
  int prefetch;          // k-blocks of L2 prefetch ahead of the TMA loads (0 = off, the default: see AScoreSmem)
  float* rowmax;         // [n_pairs, x]
};

template <int BN, int STAGES = 4>
struct AScoreSmem {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = STAGES;
  static constexpr int kTotal = kStages * kStageBytes + 2 * kBM * 4 /*inv_other*/ + 256 /*barriers*/ + 1024 /*align*/;
};

// sum of squares of one 128-byte (64 x bf16) swizzled tile row; chunk order is irrelevant for a sum, so thread t starts at
// chunk (t & 7): the 8 threads of an LDS.128 phase hit 8 different bank groups
__device__ __forceinline__ float row_sumsq_128B(uint32_t row_addr, int t) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t a, b, c, d;
    ld_shared_v4(row_addr + (((j ^ t) & 7) << 4), a, b, c, d);
    const uint32_t w[4] = {a, b, c, d};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float lo = __uint_as_float(w[k] << 16), hi = __uint_as_float(w[k] & 0xffff0000u);
      s0 = fmaf(lo, lo, s0);
      s1 = fmaf(hi, hi, s1);
    }
  }
  return s0 + s1;
}

template <int BN, int STAGES = 4>
__global__ void __launch_bounds__(kAsThreads, 1)
ascore_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const AScoreArgs p) {
  using L = AScoreSmem<BN, STAGES>;
  constexpr uint32_t kTmemCols = (2 * BN <= 256) ? 256 : 512;
  constexpr uint32_t kIdesc = make_idesc_16(kBM, BN, 0, 0, false);
  constexpr int kAsStages = L::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int m_tiles = (p.x + kBM - 1) / kBM;
  const int n_tiles = (p.y + BN - 1) / BN;
  const int k_blocks = p.K / kBK;
  const int units = p.n_pairs * m_tiles;
  uint8_t* stage_base = smem;
  float* inv_other = reinterpret_cast<float*>(smem + kAsStages * L::kStageBytes);  // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(inv_other + 2 * kBM);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kAsStages;
  uint64_t* tmem_full = bars + 2 * kAsStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* aux_full = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_full + 2);

  const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kAsStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1 + 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
      mbar_init(&aux_full[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // optional L2 prefetch p.prefetch steps ahead of the loads, across tile and unit boundaries (first touches — the A tile
      // on its first pass; the B tiles once per pair — come from HBM).  Off by default: measured slower, the pairs in flight
      // already fill most of L2.
      // (two odometers instead of a flat index: a 64-bit division per step costs more than the step)
      struct Pos { int u, nt, kb; };
      auto advance = [&](Pos& q) {
        if (++q.kb == k_blocks) {
          q.kb = 0;
          if (++q.nt == n_tiles) { q.nt = 0; q.u += static_cast<int>(gridDim.x); }
        }
      };
      Pos cur{static_cast<int>(blockIdx.x), 0, 0}, pf = cur;
      for (int i = 0; i < p.prefetch && pf.u < units; ++i) advance(pf);
      for (; cur.u < units; advance(cur)) {
        if (p.prefetch > 0 && pf.u < units) {
          const int pb = pf.u / m_tiles, pmt = pf.u % m_tiles;
          if (pf.nt == 0) tma_prefetch_3d(&tmA, pf.kb * kBK, pmt * kBM, pb);
          if (pmt == 0 || pf.u < static_cast<int>(gridDim.x)) tma_prefetch_3d(&tmB, pf.kb * kBK, pf.nt * BN, pb);
          advance(pf);
        }
        const int b = cur.u / m_tiles, mt = cur.u % m_tiles;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = stage_base + stage * L::kStageBytes;
        mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
        tma_load_3d(sa, &tmA, &full_bar[stage], cur.kb * kBK, mt * kBM, b);
        tma_load_3d(sa + L::kABytes, &tmB, &full_bar[stage], cur.kb * kBK, cur.nt * BN, b);
        if (++stage == kAsStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t kDescHi = smem_desc_hi_sw128(1024);
    const uint32_t a_lo0 = smem_desc_lo(smem_u32(stage_base), 16);
    const uint32_t b_lo0 = smem_desc_lo(smem_u32(stage_base) + L::kABytes, 16);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x)
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_lo = a_lo0 + stage * (L::kStageBytes >> 4);
            const uint32_t b_lo = b_lo0 + stage * (L::kStageBytes >> 4);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16_ss(d_tmem, smem_desc_join(kDescHi, a_lo + 2 * k), smem_desc_join(kDescHi, b_lo + 2 * k), kIdesc,
                           (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (kb == k_blocks - 1) umma_commit(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == kAsStages) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
  } else if (warp < 6) {
    // ===================================================================== epilogue: running row maximum
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    int k = 0;
    const int rt = quarter * 32 + lane;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++k) {
      const int b = u / m_tiles, mt = u % m_tiles;
      const float* cs = p.inv_ref + static_cast<size_t>(b) * p.y;
      float run_max = -INFINITY;
      float rs = 0.f;
      for (int nt = 0; nt < n_tiles; ++nt) {
        mbar_wait(&tmem_full[acc], acc_phase);
        mbar_wait(&aux_full[acc], acc_phase);
        tc_fence_after();
        // this unit's row scale was published with its first tile; taken into a register now, while the accumulator
        // hand-shake still keeps the norm warps from reusing the slot
        if (nt == 0) rs = inv_other[(k & 1) * kBM + rt];
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int col0 = nt * BN + c * 32;
          if (col0 >= p.y) break;  // warp-uniform
          uint32_t r0[32];
          tmem_ld_32x32b_x32(t_row + c * 32, r0);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float v = __uint_as_float(r0[e]) * __ldg(cs + min(col0 + e, p.y - 1));  // same address in every lane
            if (col0 + e < p.y) run_max = fmaxf(run_max, v);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      const int row = mt * kBM + rt;
      if (row < p.x) p.rowmax[static_cast<size_t>(b) * p.x + row] = run_max * rs;
    }
  } else {
    // ===================================================================== norm warps
    const int t = threadIdx.x - 6 * 32;  // 0..127: tile row
    int stage = 0, acc = 0;
    uint32_t phase = 0;
    int k = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++k)
      for (int nt = 0; nt < n_tiles; ++nt) {
        float ss = 0.f;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (nt == 0) ss += row_sumsq_128B(smem_u32(stage_base + stage * L::kStageBytes) + t * 128, t);
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[stage]);
          if (++stage == kAsStages) { stage = 0; phase ^= 1; }
        }
        if (nt == 0) inv_other[(k & 1) * kBM + t] = inv_norm_scale(ss);
        __syncwarp();
        if (lane == 0) mbar_arrive(&aux_full[acc]);  // release: the scales above are visible to the epilogue's acquire
        acc ^= 1;
      }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace lovr
