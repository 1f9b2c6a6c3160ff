// Fused non-causal ViT attention on tcgen05, fourth generation: attention3's three softmax pipelines, PERSISTENT.
//
// attention3 runs one CTA per (image, head) and one CTA per SM (224 KB of shared memory), so every CTA pays ~5 us that
// nothing overlaps (launch, TMEM allocation, the first K/Q loads, the last tile's epilogue, teardown: 14 % of the CTA at
// T = 577, 40-54 % at T = 257 / 196), and its third pipeline idles while the 4th and 5th query tiles run (5 tiles on 3
// pipelines).  Here one CTA per SM walks a list of items (image, head) and its query tiles form ONE rolling sequence,
// tile s -> pipeline s % 3, across item boundaries:
//   * K and V stay single-buffered (160 KB), but every 128-row block has its own full/free mbarrier pair: a tile's MMA
//     issuer commits `k_free[j]` / `v_free[j]` once it is past block j, and when all tiles of the item have done so the
//     TMA warp loads the NEXT item's block j into the same shared memory.  The next item's first tile therefore starts
//     while the current item's last tiles are still running (trailing them by ~3 units), and no pipeline idles.
//   * the TMA warp is one polling loop over every pending request (Q tile per pipeline, next K block, next V block), so
//     no request waits behind another.
//   * TMEM, barriers and setmaxnreg are set up once per SM.
// Up to three items can be in flight when an item has two query tiles (T <= 256); the block-full barriers therefore
// alternate between two sets (see below).  T <= 128 (one tile per item, one K/V block) keeps the attention3 kernel.
// Register split (setmaxnreg): 80 for the four control warps, 144 for the twelve softmax warps.  Measured per 64 images
// at T = 577 with control budgets 40 / 56 / 80 / 104 / 128 (no rebalancing): 0.240 / 0.217 / 0.209 / 0.240 / 0.314 ms — at 40
// registers the MMA issue loop spills to local memory, and the issuer's loop time is on the critical path.
// Tried and dropped (measured slower): letting a softmax warp whose 32 rows all lie past T (fourth warp of the last tile at
// T = 577) skip its loads / exponentials / P store (0.195 -> 0.203 ms per 64 images); moving every 4th / 2nd exponential to the FMA pipe (cubic 2^f, +5 % / +15 %: the
// kernel is issue- and latency-bound, not MUFU-bound); ex2.approx.f16x2 (compiles to two MUFU.EX2.F16); 32-column half units.
#pragma once
#include "attention3.cuh"

namespace lovr {

constexpr int kAttn4ControlRegs = 80;

template <bool F16, int RC>  // RC = registers of the four control warps; the softmax warps get (65536 - 128 RC) / 384
__global__ void __launch_bounds__(kAttn3Threads, 1)
attention4_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NP = 3;
  const int T = p.T, H = p.H, nb = p.nb, nu = p.nu, last_w = p.last_w64;
  const float c1 = p.scale * 1.4426950408889634f;
  uint16_t* const out = p.out;
  uint8_t* sK = smem;
  uint8_t* sV = sK + nb * 16384;
  uint8_t* sQ = sV + nb * 16384;  // [3 pipelines] x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sQ + NP * 16384);
  // barriers per 128-row block j: k_full[j] (TMA -> QK^T issuers), k_free[j] (every tile of the item has retired its QK^T
  // on the block: the next item's block may land), v_full[j], v_free[j]; then per pipeline g (base kPipeBar + 8 g):
  // q_full, q_empty, s_full, s_empty, p_full, p_empty, o_full, o_empty
  // Three consecutive tiles are in flight.  With >= 3 tiles per item they span two items and the parity of a block's fill
  // identifies the item.  With two tiles per item (T <= 256) they can span three: a tile of item i can reach a block
  // while item i-1's fill is still pending and the block holds item i-2, whose "fill complete" has the parity it waits
  // for.  Then the full barriers alternate between two sets (even / odd items): the fill of item i is phase i/2 of set
  // i%2 and the nearest earlier fill with the same parity is four items back.
  // Layout note: the placement of these 8-byte objects is measurable.  With the first set, the free barriers and the
  // pipeline groups packed as below the kernel runs 0.230 ms per 64 images at T = 577; with the two full sets placed first
  // (or every pipeline group padded to its own 128-byte line) 0.242 ms — same code otherwise.  The second set, only used
  // when an item has two tiles, therefore sits behind the pipeline groups.
  constexpr int kPipeBar = 4 * kAttnMaxBlocks;
  constexpr int kSet1 = kPipeBar + 8 * NP;        // second set of full barriers: k at kSet1 + j, v at kSet1 + 5 + j
  uint64_t* k_full = bars + 0;
  uint64_t* k_free = bars + kAttnMaxBlocks;
  uint64_t* v_full = bars + 2 * kAttnMaxBlocks;
  uint64_t* v_free = bars + 3 * kAttnMaxBlocks;
  uint64_t* k_full1 = bars + kSet1;
  uint64_t* v_full1 = bars + kSet1 + kAttnMaxBlocks;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kSet1 + 2 * kAttnMaxBlocks);

  // warp index through a shuffle: the compiler then KNOWS it is warp-uniform and keeps everything derived from it (pipeline
  // index, TMEM / shared-memory addresses, descriptors) on the uniform datapath — tcgen05.mma takes uniform-register operands,
  // and values it cannot prove uniform go through an ELECT / R2UR loop per instruction (~150 cycles per MMA, measured)
  const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
  const int D = H * 64;
  const int q_tiles = (T + 127) / 128;
  // this CTA's items: blockIdx.x, blockIdx.x + gridDim.x, ...; its tiles in one rolling sequence, tile s -> pipeline s % 3
  const int my_items = (p.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int total_tiles = my_items * q_tiles;
  const bool two_sets = q_tiles < 3;  // see the barrier comment above: three tiles in flight span three items only then

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int j = 0; j < kAttnMaxBlocks; ++j) {
      mbar_init(k_full + j, 1);
      mbar_init(v_full + j, 1);
      mbar_init(k_full1 + j, 1);
      mbar_init(v_full1 + j, 1);
    }
    for (int j = 0; j < kAttnMaxBlocks; ++j) {
      mbar_init(k_free + j, q_tiles);  // one tcgen05.commit per tile of the item
      mbar_init(v_free + j, q_tiles);
    }
    for (int g = 0; g < NP; ++g) {
      uint64_t* b = bars + kPipeBar + 8 * g;
      mbar_init(b + 0, 1);  // q_full
      mbar_init(b + 1, 1);  // q_empty
      mbar_init(b + 2, 1);  // s_full
      mbar_init(b + 3, 4);  // s_empty
      mbar_init(b + 4, 4);  // p_full
      mbar_init(b + 5, 1);  // p_empty
      mbar_init(b + 6, 1);  // o_full
      mbar_init(b + 7, 4);  // o_empty
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

  if (warp < 4) {
    if constexpr (RC < 128) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RC));
    if (warp == 0) {
      // ===================================================================== TMA
      if (lane == 0) {
        // One polling loop serves every pending request so that no pipeline waits behind another one's load:
        //   Q tile of pipeline g (needs the QK^T of that pipeline's previous tile retired),
        //   next K block / next V block in (item, block) order (needs the previous item's tiles past that block).
        int q_next[NP], q_cnt[NP];
        for (int g = 0; g < NP; ++g) { q_next[g] = g; q_cnt[g] = 0; }
        int k_item = 0, k_blk = 0, v_item = 0, v_blk = 0;
        const long long t0 = clock64();
        for (;;) {
          bool pending = false;
          for (int g = 0; g < NP; ++g) {
            const int s = q_next[g];
            if (s >= total_tiles) continue;
            pending = true;
            uint64_t* b = bars + kPipeBar + 8 * g;
            if (q_cnt[g] == 0 || mbar_try_wait(b + 1, (q_cnt[g] - 1) & 1)) {
              const int item = static_cast<int>(blockIdx.x) + (s / q_tiles) * static_cast<int>(gridDim.x);
              mbar_arrive_expect_tx(b + 0, 16384);
              tma_load_3d(sQ + g * 16384, &tmQKV, b + 0, (item % H) * 64, (s % q_tiles) * 128, item / H);
              q_next[g] += NP;
              ++q_cnt[g];
            }
          }
          if (k_item < my_items) {
            pending = true;
            if (k_item == 0 || mbar_try_wait(k_free + k_blk, (k_item - 1) & 1)) {
              const int item = static_cast<int>(blockIdx.x) + k_item * static_cast<int>(gridDim.x);
              uint64_t* full = ((two_sets && (k_item & 1)) ? k_full1 : k_full) + k_blk;
              mbar_arrive_expect_tx(full, 16384);
              tma_load_3d(sK + k_blk * 16384, &tmQKV, full, D + (item % H) * 64, k_blk * 128, item / H);
              if (++k_blk == nb) { k_blk = 0; ++k_item; }
            }
          }
          if (v_item < my_items) {
            pending = true;
            if (v_item == 0 || mbar_try_wait(v_free + v_blk, (v_item - 1) & 1)) {
              const int item = static_cast<int>(blockIdx.x) + v_item * static_cast<int>(gridDim.x);
              uint64_t* full = ((two_sets && (v_item & 1)) ? v_full1 : v_full) + v_blk;
              mbar_arrive_expect_tx(full, 16384);
              tma_load_3d(sV + v_blk * 16384, &tmQKV, full, 2 * D + (item % H) * 64, v_blk * 128, item / H);
              if (++v_blk == nb) { v_blk = 0; ++v_item; }
            }
          }
          if (!pending) break;
          if (clock64() - t0 > 4 * LOVR_SPIN_LIMIT_CYCLES) {
            printf("lovr: attention4 TMA poll timed out (block %d)\n", blockIdx.x);
            __trap();
          }
        }
      }
    } else {
      // ===================================================================== MMA issuer of pipeline g = warp - 1
      const int g = warp - 1;
      uint64_t* b = bars + kPipeBar + 8 * g;
      uint64_t* const q_full = b + 0;
      uint64_t* const q_empty = b + 1;
      uint64_t* const s_full = b + 2;
      uint64_t* const s_empty = b + 3;
      uint64_t* const p_full = b + 4;
      uint64_t* const p_empty = b + 5;
      uint64_t* const o_full = b + 6;
      uint64_t* const o_empty = b + 7;
      const uint32_t tS = tmem_base + g * kAttn3PipeCols;
      const uint32_t tO = tS + 64;
      const uint32_t tP = tS + 128;
      constexpr uint32_t kHi = smem_desc_hi_sw128(1024);
      const uint32_t q_lo = smem_desc_lo(smem_u32(sQ + g * 16384), 16);
      const uint32_t k_lo0 = smem_desc_lo(smem_u32(sK), 16);
      const uint32_t v_lo0 = smem_desc_lo(smem_u32(sV), 16384);
      const uint32_t idesc_o = make_idesc_16(128, 64, 0, 1, F16);  // B (= V) is MN-major
      const uint32_t idesc_s64 = make_idesc_16(128, 64, 0, 0, F16);
      const uint32_t idesc_slast = make_idesc_16(128, last_w, 0, 0, F16);
      uint32_t n_s = 0;  // QK^T units issued (S fills)
      uint32_t n_p = 0;  // P units consumed
      int k = 0;
      for (int s = g; s < total_tiles; s += NP, ++k) {
        const int il = s / q_tiles;                                  // this CTA's il-th item
        const bool set1 = two_sets && (il & 1);                      // barrier set and parity of its K/V block fills
        uint64_t* const kf = set1 ? k_full1 : k_full;
        uint64_t* const vf = set1 ? v_full1 : v_full;
        const uint32_t kv_par = static_cast<uint32_t>(two_sets ? (il >> 1) : il) & 1u;
        mbar_wait(q_full, k & 1);
        for (int u = 0; u <= nu; ++u) {
          // QK^T(u) goes first: S(u-1) is already in the softmax warps' registers, so it runs while they exponentiate
          if (u < nu) {
            mbar_wait(kf + (u >> 1), kv_par);  // this item's K block u/2 has landed
            mbar_wait(s_empty, (n_s & 1) ^ 1);
            tc_fence_after();
            if (lane == 0) {
              const uint32_t k_lo = k_lo0 + u * (8192 >> 4);  // unit u = rows [64u, 64u+64) of the resident K
              const uint32_t idesc = (u == nu - 1) ? idesc_slast : idesc_s64;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_bf16_ss(tS, smem_desc_join(kHi, q_lo + 2 * kk), smem_desc_join(kHi, k_lo + 2 * kk), idesc, kk != 0);
              umma_commit(s_full);
              if (u == nu - 1) umma_commit(q_empty);  // every QK^T of this tile issued: Q may be overwritten once they retire
              if ((u & 1) || u == nu - 1) umma_commit(k_free + (u >> 1));  // this tile is past K block u/2
            }
            __syncwarp();
            ++n_s;
          }
          if (u == 0) {
            mbar_wait(o_empty, (k & 1) ^ 1);  // epilogue of this pipeline's previous tile has read O
            tc_fence_after();
          } else {
            // P·V(u-1)
            const int up = u - 1;
            mbar_wait(vf + (up >> 1), kv_par);
            mbar_wait(p_full, n_p & 1);
            tc_fence_after();
            if (lane == 0) {
              const uint32_t v_lo = v_lo0 + up * (8192 >> 4);
              const int steps = ((up == nu - 1) ? last_w : 64) / 16;
              for (int s = 0; s < steps; ++s)  // A = P from TMEM: lane = row, 8 columns of packed 16-bit pairs per K=16 step
                umma_bf16_ts(tO, tP + s * 8, smem_desc_join(kHi, v_lo + s * (2048 >> 4)), idesc_o,
                             (up == 0 && s == 0) ? 0u : 1u);
              umma_commit(p_empty);
              if ((up & 1) || up == nu - 1) umma_commit(v_free + (up >> 1));
              if (up == nu - 1) umma_commit(o_full);
            }
            __syncwarp();
            ++n_p;
          }
        }
      }
    }
  } else {
    // ===================================================================== softmax pipelines
    if constexpr (RC < 128) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"((65536 - 128 * RC) / 384 / 8 * 8));
    const int g = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    uint64_t* b = bars + kPipeBar + 8 * g;
    uint64_t* const s_full = b + 2;
    uint64_t* const s_empty = b + 3;
    uint64_t* const p_full = b + 4;
    uint64_t* const p_empty = b + 5;
    uint64_t* const o_full = b + 6;
    uint64_t* const o_empty = b + 7;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + g * kAttn3PipeCols + lane_sel;
    const uint32_t tO = tS + 64;
    const uint32_t tPw = tS + 128;
    uint32_t n_s = 0;  // S units consumed
    uint32_t n_p = 0;  // P units produced
    int k = 0;
    uint32_t buf[2][32];
    uint32_t pk[32];
    // hand P(u) to the P·V issuer: the tcgen05.st must have completed before the arrive
    auto hand_off = [&]() {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    };
    // O (and the running sums) *= alpha for the rows whose maximum moved; every P·V issued so far must have landed
    auto rescale = [&](bool need, float hm, float& m_used, float& l0, float& l1, bool o_valid) {
      const float alpha = need ? ex2_approx((m_used - hm) * c1) : 1.0f;
      if (need) m_used = hm;
      l0 *= alpha;
      l1 *= alpha;
      if (o_valid) {
        mbar_wait(p_empty, (n_p & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int piece = 0; piece < 2; ++piece) {
          uint32_t tmp[32];
          tmem_ld_32x32b_x32(tO + piece * 32, tmp);
          tmem_ld_wait_dep(tmp);
#pragma unroll
          for (int e = 0; e < 32; ++e) tmp[e] = __float_as_uint(__uint_as_float(tmp[e]) * alpha);
          tmem_st_32x32b_x32(tO + piece * 32, tmp);
        }
        tmem_st_wait();
        tc_fence_before();
      }
      return alpha;
    };
    for (int s = g; s < total_tiles; s += NP, ++k) {
      const int item = static_cast<int>(blockIdx.x) + (s / q_tiles) * static_cast<int>(gridDim.x);
      const int n = item / H, h = item % H, t = s % q_tiles;
      float m_used = 0.f;  // row maximum (raw score units) the exponentials are currently taken against
      float l0 = 0.f, l1 = 0.f;
      bool pending = false;  // P(u-1) stored to TMEM but not yet handed off
      for (int u = 0; u < nu; ++u) {
        const int w = (u == nu - 1) ? last_w : 64;
        const int valid = min(64, T - u * 64);
        mbar_wait(s_full, n_s & 1);
        tc_fence_after();
        tmem_ld_32x32b_x32(tS, buf[0]);
        if (w == 64) tmem_ld_32x32b_x32(tS + 32, buf[1]);
        if (pending) hand_off();  // overlaps the load latency
        tmem_ld_wait_dep(buf[0]);
        {
          if (w == 64) tmem_ld_wait_dep(buf[1]);
          // S unit is in registers: give the buffer back so the next QK^T can start
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty);
          ++n_s;
          // ---- reference maximum.  The exact row maximum is only taken over the tile's FIRST unit; afterwards the
          // exponentials keep using it and the guard is the unit's own row sum (already computed, all terms >= 0): as
          // long as it stays below kSumLimit no P can overflow its 16-bit format, and a larger-than-one P is harmless —
          // O and l carry the same factor and terms 2^-24 below the largest are below fp32 resolution anyway.  Only when
          // the sum trips the limit (or is inf/NaN) the unit is redone: exact maximum, O and l rescaled, exponentials
          // recomputed from the scores still in registers.  This removes the 64-element max reduction and the vote
          // from nine of ten units (21 % of the softmax warps' instructions).
          constexpr float kSumLimit = F16 ? 32768.0f : 1.0e18f;
          auto unit_max = [&]() {
            if (valid == 64) return fmaxf(attn3_max_half<false>(buf[0], 32), attn3_max_half<false>(buf[1], 32));
            return (w == 64) ? fmaxf(attn3_max_half<false>(buf[0], 32), attn3_max_half<true>(buf[1], valid - 32))
                             : attn3_max_half<true>(buf[0], valid);
          };
          auto exp_unit = [&](float mc) {
            if (valid == 64) {
              attn3_exp_half<F16, false>(buf[0], 32, c1, mc, l0, l1, pk);
              attn3_exp_half<F16, false>(buf[1], 32, c1, mc, l0, l1, pk + 16);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) pk[i] = 0u;
              if (valid >= 32) attn3_exp_half<F16, false>(buf[0], 32, c1, mc, l0, l1, pk);
              else attn3_exp_half<F16, true>(buf[0], valid, c1, mc, l0, l1, pk);
              if (valid > 32) attn3_exp_half<F16, true>(buf[1], valid - 32, c1, mc, l0, l1, pk + 16);
            }
          };
          if (u == 0) m_used = unit_max();
          const float lp0 = l0, lp1 = l1;
          exp_unit(m_used * c1);
          const bool bad = !((l0 - lp0) + (l1 - lp1) < kSumLimit);
          if (u > 0 && __any_sync(0xffffffffu, bad)) {
            l0 = lp0;
            l1 = lp1;
            rescale(bad, unit_max(), m_used, l0, l1, true);
            exp_unit(m_used * c1);
          }
        }
        // ---- hand the unit to the P·V issuer
        mbar_wait(p_empty, (n_p & 1) ^ 1);
        tmem_st_32x32b_x32(tPw, pk);  // 64 P values of this row as 32 packed columns
        ++n_p;
        pending = true;
      }
      if (pending) hand_off();
      // ---------------- O / l -> global
      mbar_wait(o_full, k & 1);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32b_x32(tO, o0);
      tmem_ld_32x32b_x32(tO + 32, o1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      const int tok = t * 128 + r;
      if (tok < T) {
        const float inv = 1.0f / (l0 + l1);
        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(n) * T + tok) * D + h * 64);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint4 w4;
          const uint32_t* src = (q < 4) ? &o0[q * 8] : &o1[(q - 4) * 8];
          w4.x = pack16x2<F16>(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
          w4.y = pack16x2<F16>(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
          w4.z = pack16x2<F16>(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
          w4.w = pack16x2<F16>(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
          dst[q] = w4;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace lovr
