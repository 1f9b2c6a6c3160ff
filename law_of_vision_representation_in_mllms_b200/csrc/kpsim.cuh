// C score, keypoint rows of the similarity matrix on tcgen05 (bf16 descriptors):
//     out[b, k, t] = n(d1[b, idx[b,k], :]) . n(d2[b, t, :]),   n(x) = x / (||x|| + 1e-10)      (C_score/pck_train.py:24-29,
//     utils_correspondence.py:360: only the K source-keypoint rows of the P^2 x P^2 matrix are consumed by the callers)
// The work is one read of the target descriptors (7.4 MB per P = 60 pair): HBM-bound.  The CUDA-core kernel
// (kp_sim_rows4_kernel) feeds its fp32 FMAs from shared memory and reaches 0.16 of the copy bandwidth; here the dot products
// go to the tensor core as a skinny GEMM — M = 128 target rows per tile, N = 32 (keypoints, zero-padded), K = D — so the SM
// only has to keep TMA loads in flight.
//
// Unit = (pair b, group of kKpTilesPerUnit consecutive 128-row target tiles), dealt round-robin to persistent CTAs.
//   warps 6..9  (128 threads) gather the K source rows of the pair into shared memory in the K-major 128-byte-swizzled
//               UMMA layout (D/64 k-blocks of 32 rows x 128 B: 64 KB at D = 1024, resident for the whole unit) and their
//               inverse norms; then, per pipeline stage, the sum of squares of the 128 target rows from the stage's tile
//               (thread = row, conflict-free LDS.128 as in ascore.cuh) -> inverse norms of the targets
//   warp 0      TMA producer: target tiles 128 rows x 64 columns through a 6-stage ring (3-D map: column, row, pair)
//   warp 1      MMA issuer: 128 x 32 x 16, accumulators (32 TMEM columns) double-buffered per tile
//   warps 2..5  epilogue: thread = target row; out[b, k, t] = acc[k] * inv_src[k] * inv_tgt[t] for k < K (coalesced over t)
// Barriers: full/empty per stage (empty: tcgen05.commit + 4 norm warps), tmem_full/tmem_empty per accumulator, aux_full per
// accumulator (norm warps: the tile's target norms are published), b_ready (norm warps -> MMA issuer: the unit's source rows
// are in place; generic-proxy writes are fenced for the async proxy), b_free (tcgen05.commit after the unit's last MMA ->
// norm warps: the source rows may be overwritten).
#pragma once
#include "gemm.cuh"
#include "scores.cuh"
#include "ascore.cuh"

namespace lovr {

constexpr int kKpN = 32;
#ifndef LOVR_KP_STAGES
#define LOVR_KP_STAGES 6
#endif
#ifndef LOVR_KP_TILES_PER_UNIT
#define LOVR_KP_TILES_PER_UNIT 8
#endif
// -DLOVR_KP_STAGES / -DLOVR_KP_TILES_PER_UNIT: A/B builds (build.py --out).  Measured at 512 P=60 pairs (us per pair): tiles per
// unit 4 / 8 / 10 / 15 / 29 = 1.58 / 1.46 / 1.44 / 1.43 / 1.48 (each unit re-gathers the pair's K source rows; 8 keeps ~7 waves of
// units at the 256-pair chunks the job uses); stages 6 / 8 / 9 = 1.58 / 1.57 / 1.57 (6 keeps D = 2048 inside 227 KB).
constexpr int kKpStages = LOVR_KP_STAGES;
constexpr int kKpTilesPerUnit = LOVR_KP_TILES_PER_UNIT;
constexpr int kKpThreads = 64 + 4 * 32 + 4 * 32;
constexpr int kKpABytes = kBM * kBK * 2;

struct KpArgs {
  int PP, D, K, n_pairs;
  const uint16_t* d1;   // [n_pairs, PP, D] bf16
  const int* idx;       // [n_pairs, K]
  float* out;           // [n_pairs, K, PP]
};

inline size_t kp_smem_bytes(int D) {
  return static_cast<size_t>(kKpStages) * kKpABytes + static_cast<size_t>(D / 64) * kKpN * 128 + (4 * kKpN + 2 * kBM) * 4 + 256 + 1024;
}

__global__ void __launch_bounds__(kKpThreads, 1)
kpsim_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const KpArgs p) {
  constexpr uint32_t kTmemCols = 64;  // two accumulators of 32 columns
  constexpr uint32_t kIdesc = make_idesc_16(kBM, kKpN, 0, 0, false);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int k_blocks = p.D / kBK;
  const int m_tiles = (p.PP + kBM - 1) / kBM;
  const int groups = (m_tiles + kKpTilesPerUnit - 1) / kKpTilesPerUnit;
  const int units = p.n_pairs * groups;
  uint8_t* stage_base = smem;
  uint8_t* sB = smem + kKpStages * kKpABytes;                                   // [k_blocks][32 rows][128 B], swizzled
  // inverse source norms: a ring of four units — the epilogue may still be reading unit k's scales while the norm warps,
  // which only wait for the MMAs, already gather unit k + 2 (one-tile units: P <= 11)
  float* inv_src = reinterpret_cast<float*>(sB + static_cast<size_t>(k_blocks) * kKpN * 128);  // [4][32]
  float* inv_tgt = inv_src + 4 * kKpN;                                          // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(inv_tgt + 2 * kBM);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kKpStages;
  uint64_t* tmem_full = bars + 2 * kKpStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* aux_full = tmem_empty + 2;
  uint64_t* b_ready = aux_full + 2;
  uint64_t* b_free = b_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_free + 1);

  const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    for (int i = 0; i < kKpStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1 + 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
      mbar_init(&aux_full[i], 4);
    }
    mbar_init(b_ready, 4);
    mbar_init(b_free, 1);
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

  auto unit_tiles = [&](int u, int* b, int* t0, int* t1) {
    *b = u / groups;
    *t0 = (u % groups) * kKpTilesPerUnit;
    *t1 = min(m_tiles, *t0 + kKpTilesPerUnit);
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int b, t0, t1;
        unit_tiles(u, &b, &t0, &t1);
        for (int mt = t0; mt < t1; ++mt)
          for (int kb = 0; kb < k_blocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], kKpABytes);
            tma_load_3d(stage_base + stage * kKpABytes, &tmA, &full_bar[stage], kb * kBK, mt * kBM, b);
            if (++stage == kKpStages) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t kDescHi = smem_desc_hi_sw128(1024);
    const uint32_t a_lo0 = smem_desc_lo(smem_u32(stage_base), 16);
    const uint32_t b_lo0 = smem_desc_lo(smem_u32(sB), 16);
    int stage = 0, acc = 0, k = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++k) {
      int b, t0, t1;
      unit_tiles(u, &b, &t0, &t1);
      mbar_wait(b_ready, k & 1);  // this unit's source rows are in shared memory
      tc_fence_after();
      for (int mt = t0; mt < t1; ++mt) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kKpN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_lo = a_lo0 + stage * (kKpABytes >> 4);
            const uint32_t b_lo = b_lo0 + kb * ((kKpN * 128) >> 4);
#pragma unroll
            for (int kk = 0; kk < kBK / 16; ++kk)
              umma_bf16_ss(d_tmem, smem_desc_join(kDescHi, a_lo + 2 * kk), smem_desc_join(kDescHi, b_lo + 2 * kk), kIdesc,
                           (kb | kk) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (kb == k_blocks - 1) {
              umma_commit(&tmem_full[acc]);
              if (mt == t1 - 1) umma_commit(b_free);  // every MMA that reads this unit's source rows has been issued
            }
          }
          __syncwarp();
          if (++stage == kKpStages) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp < 6) {
    // ===================================================================== epilogue
    const int quarter = warp & 3;
    const int rt = quarter * 32 + lane;
    int acc = 0, k = 0;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++k) {
      int b, t0, t1;
      unit_tiles(u, &b, &t0, &t1);
      for (int mt = t0; mt < t1; ++mt) {
        mbar_wait(&tmem_full[acc], acc_phase);
        mbar_wait(&aux_full[acc], acc_phase);
        tc_fence_after();
        uint32_t r0[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kKpN, r0);
        tmem_ld_wait();
        const float it = inv_tgt[acc * kBM + rt];
        float is[kKpN];
#pragma unroll
        for (int e = 0; e < kKpN; ++e) is[e] = inv_src[(k & 3) * kKpN + e];
        tc_fence_before();
        __syncwarp();
        const int t = mt * kBM + rt;
        if (t < p.PP) {
          float* dst = p.out + static_cast<size_t>(b) * p.K * p.PP + t;
#pragma unroll
          for (int e = 0; e < kKpN; ++e)
            if (e < p.K) dst[static_cast<size_t>(e) * p.PP] = __uint_as_float(r0[e]) * is[e] * it;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);  // after the inv_tgt / inv_src reads: the norm warps reuse the slots
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================================================================== gather + norm warps
    const int t = threadIdx.x - 6 * 32;  // 0..127
    const int nw = warp - 6;             // 0..3
    int stage = 0, acc = 0, k = 0;
    uint32_t phase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++k) {
      int b, t0, t1;
      unit_tiles(u, &b, &t0, &t1);
      // ---- source rows of this pair -> sB (the previous unit's MMAs must be done with it)
      if (k > 0) mbar_wait(b_free, (k - 1) & 1);
      const int chunks = p.D / 8;  // 16-byte chunks per row
      for (int r = nw; r < kKpN; r += 4) {   // warp per row
        float ss = 0.f;
        const bool live = r < p.K;
        const uint4* src = nullptr;
        if (live) src = reinterpret_cast<const uint4*>(p.d1 + (static_cast<size_t>(b) * p.PP + __ldg(p.idx + static_cast<size_t>(b) * p.K + r)) * p.D);
        for (int q = lane; q < chunks; q += 32) {
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (live) v = __ldg(src + q);
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float lo = __uint_as_float(w[i] << 16), hi = __uint_as_float(w[i] & 0xffff0000u);
            ss = fmaf(lo, lo, ss);
            ss = fmaf(hi, hi, ss);
          }
          const int kb = q >> 3, c = q & 7;
          st_shared_v4(smem_u32(sB) + kb * (kKpN * 128) + r * 128 + ((c ^ (r & 7)) << 4), v.x, v.y, v.z, v.w);
        }
        ss = warp_sum(ss);
        if (lane == 0) inv_src[(k & 3) * kKpN + r] = live ? 1.0f / (sqrtf(ss) + 1e-10f) : 0.f;
      }
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's (async proxy) operand reads
      __syncwarp();
      if (lane == 0) mbar_arrive(b_ready);
      // ---- target norms, tile by tile
      for (int mt = t0; mt < t1; ++mt) {
        float ss = 0.f;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          ss += row_sumsq_128B(smem_u32(stage_base + stage * kKpABytes) + t * 128, t);
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[stage]);
          if (++stage == kKpStages) { stage = 0; phase ^= 1; }
        }
        inv_tgt[acc * kBM + t] = 1.0f / (sqrtf(ss) + 1e-10f);
        __syncwarp();
        if (lane == 0) mbar_arrive(&aux_full[acc]);
        acc ^= 1;
      }
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
