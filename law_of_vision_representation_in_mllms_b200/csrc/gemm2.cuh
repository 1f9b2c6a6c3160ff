// 2-CTA (tcgen05 cta_group::2) variant of the dense-layer GEMM:  C[M,N] = epilogue(A[M,K] · W[N,K]^T).
//
// A cluster of two CTAs on one TPC computes a 256 x 256 output tile with UMMA_M = 256: each CTA TMA-loads ITS 128 rows of
// A and ITS half (128 of the 256 rows) of the W tile per k-block, and owns the 128 x 256 fp32 accumulator of its rows in
// its own TMEM.  Only the leader CTA (cluster rank 0) issues `tcgen05.mma.cta_group::2`; the tensor cores of both SMs read
// both CTAs' shared memory, so per SM the operand traffic drops from 48 KB to 32 KB per k-block (stage) and six stages fit
// where the 1-CTA kernel has four.  Synchronisation:
//   full[stage]    (leader only)  armed by the leader's producer with the bytes of BOTH CTAs; every TMA load (either CTA)
//                                 signals it (cp.async.bulk.tensor ... .cta_group::2 with the leader's cluster address)
//   empty[stage]   (both CTAs)    tcgen05.commit.cta_group::2 ... multicast::cluster from the leader's MMA thread
//   tmem_full[acc] (both CTAs)    same multicast commit after the last k-block
//   tmem_empty[acc](leader only)  count 8: the four epilogue warps of both CTAs arrive (the peer's remotely)
// Epilogues are the same code as the 1-CTA kernel (epilogue_tile).
#pragma once
#include "gemm.cuh"

namespace lovr {

// The residual epilogue trades one pipeline stage for a 4-deep staging ring per epilogue warp (two residual loads and two
// stores in flight): it was the epilogue, not the main loop, that bounded the K=1024 out-projection.
// EW = epilogue warps: 4 (one per TMEM lane quarter, 256 columns each) or 8 (two per quarter, 128 columns each).
template <int EPI, int EW>
struct Gemm2Smem {
  static constexpr int kStages = (EPI == EPI_RESID_F32 || EW == 8) ? 5 : 6;
  static constexpr int kResidBufs = (EPI == EPI_RESID_F32 && EW == 4) ? 4 : 2;
  static constexpr int kABytes = kBM * kBK * 2;        // 128 x 64 x 2
  static constexpr int kBBytes = 128 * kBK * 2;        // this CTA's half of the 256-row W tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = EW * kResidBufs * kStageBufBytes;
  static constexpr int kTotal = kStages * kStageBytes + kStagingBytes + 512 + 1024;
  static constexpr int kThreads = 64 + EW * 32;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on an mbarrier of the PEER CTA (cluster address).  Default semantics (release at CTA scope), the form CUTLASS uses
// for consumer-release signals across a cluster: what the waiter consumes after it (TMEM accumulators already drained with
// tcgen05.wait::ld, shared memory filled by TMA) is not this thread's generic-proxy data.  The `.release.cluster` form
// compiles to MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of every arrive (SASS) — ~1 us on the issuing thread; with it
// on the MMA thread the 2-CTA A-score kernel ran at 20.9 ms instead of 9.5 (measured).  -DLOVR_REMOTE_ARRIVE_RELEASE_CLUSTER=1
// restores it (A/B timing).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
#if defined(LOVR_REMOTE_ARRIVE_RELEASE_CLUSTER)
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#else
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#endif
}
// TMA load whose completion bytes are credited to an mbarrier that may live in the PEER CTA (cluster address).
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t mbar_cluster_addr, int c0,
                                                int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this shared offset in BOTH CTAs of the pair once all prior MMAs of this thread retire
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}

template <int EPI, int ACT, bool F16, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + EW * 32, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const GemmArgs p) {
  using L = Gemm2Smem<EPI, EW>;
  constexpr int CW = 256 / (EW / 4);  // columns per epilogue warp
  constexpr int kStages2 = L::kStages;
  constexpr int RB = L::kResidBufs;
  constexpr int BN = 256;
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kIdesc = make_idesc_16(256, BN, 0, 0, F16);  // UMMA_M = 256 across the CTA pair

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* staging = smem + kStages2 * L::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + L::kStagingBytes);
  uint64_t* full_bar = bars;                       // [kStages2]  (used in the leader)
  uint64_t* empty_bar = bars + kStages2;           // [kStages2]
  uint64_t* tmem_full = bars + 2 * kStages2;       // [2]
  uint64_t* tmem_empty = bars + 2 * kStages2 + 2;  // [2]         (used in the leader)
  uint64_t* resid_bar = bars + 2 * kStages2 + 4;   // [kEpiWarps][RB]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages2 + 4 + RB * EW);

  const int warp = uniform_warp_index();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  // units: (batch b, 256-row tile mt, 256-column tile nt), nt fastest.  Dense layers: batch 1, N a multiple of 256.  Batched
  // similarity GEMMs (C score): one problem per pair through the third tensor-map coordinate; rows / columns past M / N are
  // zero-filled on load and clipped on store.
  const int m_tiles = (p.M + 255) / 256;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int units = (p.batch > 0 ? p.batch : 1) * m_tiles * n_tiles;
  const int k_blocks = p.K / kBK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int i = 0; i < kStages2; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * EW);
    }
    for (int i = 0; i < RB * EW; ++i) mbar_init(&resid_bar[i], 1);
    fence_barrier_init();
  }
  cluster_sync_all();  // both CTAs' barriers exist before any remote signal
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);

  if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = pair; unit < units; unit += num_pairs) {
        const int nt = unit % n_tiles;
        const int mt = (unit / n_tiles) % m_tiles;
        const int b = unit / (n_tiles * m_tiles);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          tma_load_3d_2sm(sa, &tmA, leader_full, kb * kBK, mt * 256 + static_cast<int>(rank) * 128, b);
          tma_load_3d_2sm(sb, &tmB, leader_full, kb * kBK, nt * BN + static_cast<int>(rank) * 128, b);
          if (++stage == kStages2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA only)
    if (rank == 0) {
      constexpr uint32_t kDescHi = smem_desc_hi_sw128(1024);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(stage_base), 16);
      const uint32_t b_lo0 = smem_desc_lo(smem_u32(stage_base) + L::kABytes, 16);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = pair; unit < units; unit += num_pairs) {
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
              umma_bf16_ss_2sm(d_tmem, smem_desc_join(kDescHi, a_lo + 2 * k), smem_desc_join(kDescHi, b_lo + 2 * k), kIdesc,
                               (kb | k) != 0 ? 1u : 0u);
            umma_commit_2sm(&empty_bar[stage]);
            if (kb == k_blocks - 1) umma_commit_2sm(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == kStages2) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================================================================== epilogue warps (both CTAs)
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int chalf = ew >> 2;  // which CW-column part of the tile this warp owns (always 0 when EW == 4)
    uint8_t* my_stage = staging + ew * RB * kStageBufBytes;
    uint64_t* my_rbar = resid_bar + ew * RB;
    const uint32_t row_sw = (lane & 7);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t g = 0;
    float run_max = 0.f;
    for (int unit = pair; unit < units; unit += num_pairs) {
      const int nt = unit % n_tiles;
      const int mt = (unit / n_tiles) % m_tiles;
      const int b = unit / (n_tiles * m_tiles);
      const int row0 = mt * 256 + static_cast<int>(rank) * 128 + quarter * 32;
      float rs = 1.0f, rm = 0.0f;  // folded LayerNorm: the row's statistics are merged while the tile's MMAs still run
      if constexpr (EPI == EPI_BF16) ln_row_scales(p, min(row0 + lane, p.M - 1), rs, rm);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
      epilogue_tile<CW, EPI, ACT, F16, RB>(t_row + chalf * CW, nt * BN + chalf * CW, row0, b, p, tmC, my_stage, my_rbar, lane,
                                          row_sw, g, run_max, rs, rm);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0));  // leader's barrier
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  cluster_sync_all();  // nobody exits (or frees TMEM) while the peer may still signal or read this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
}

}  // namespace lovr
