// Fused non-causal ViT attention on tcgen05, third generation: the three-pipeline single-pass kernel of attention2.cuh
// rebuilt from what its ncu source-level profile showed (profiles/r01_attention3_ncu.md):
//   * every per-unit trace point read `p.debug` from LOCAL memory (the argument struct was captured by reference in
//     lambdas, so it lived on the stack) and the fill/consume counters were stack arrays: ~8 % of the softmax warps'
//     samples were long-scoreboard stalls on LDL.  Here the arguments are plain registers, counters are scalars and
//     there is no tracing in the kernel: 0.285 -> 0.250 ms per 64 images, bit-identical output.
//   * timing diagnostics (compile-time removal of one piece at a time, same launch): no exponentials -1.6 %, no
//     tcgen05.st of P -2.4 %, no tcgen05.ld of S -10 %: no single unit is saturated; the kernel is bound by the
//     per-unit dependency chain of a softmax warp (barrier wake-up -> TMEM load -> max -> exp -> TMEM store -> arrive).
//     (ex2.approx.f16x2 does not help: it compiles to two MUFU.EX2.F16, one per half.)
//   * PIPE = 1 (default) shortens that chain a little: the hand-off of P(u) (tcgen05.wait::st, fence, arrive) is deferred
//     until the TMEM load of S(u+1) has been issued, so the store's completion latency overlaps the load's.  (Splitting the
//     unit into two 32-column halves to overlap the second load with the first half's exponentials measured 5 % SLOWER: it
//     delays the release of S and with it the next QK^T.)
//   * 13-18 % of the softmax warps' samples sat in the first s_full wait of the CTA: one CTA per SM (224 KB of shared
//     memory), so nothing overlaps the 208 KB of K, V, Q loads.  K and V now complete per 128-row block (one mbarrier
//     each, loads issued in order of first use), so the first QK^T starts after 32 KB instead of 80 KB and the remaining
//     blocks stream in under the first tile's compute.
// Structure (unchanged): one CTA per (image, head); K, V resident in shared memory; query tile t -> pipeline t % 3;
// per pipeline: S (64 columns, single buffer), O (64), P (32 packed columns) in TMEM; thread = query row.
//   warp 0        TMA        warps 1-3   tcgen05.mma issuers of pipelines 0-2        warps 4-15  softmax (4 per pipeline)
#pragma once
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace lovr {

constexpr int kAttnMaxBlocks = 5;  // T <= 640: K and V of one (image, head) stay resident in shared memory
constexpr float kAttnRescaleLog2 = 8.0f;

struct AttnArgs {
  int T;        // tokens per image
  int H;        // heads
  int nb;       // ceil(T / 128)
  int last_w;   // width of the last KV block, multiple of 32
  int nu;       // ceil(T / 64) KV units
  int last_w64; // width of the last 64-column unit: 32 or 64
  float scale;  // head_dim^-0.5
  uint16_t* out;  // [N*T, H*64], 16-bit operand format
  long long* trace;  // debug & 8: clock64 timeline of CTA 0, pipeline 0 (diagnostics only)
  int n_items;    // (persistent kernel) images x heads
  int debug;      // timing experiments only (LOVR_ATTN_DEBUG)
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace lovr

namespace lovr {

constexpr int kAttn3Threads = 16 * 32;
constexpr int kAttn3PipeCols = 160;  // S [0,64) | O [64,128) | P [128,160)
constexpr int kAttn3DefaultPipe = 1;

inline size_t attn3_smem_bytes(int nb) { return static_cast<size_t>(2 * nb + 3) * 16384 + 512 + 1024; }

// P = exp2(s*c1 - mc) for one 32-column half held in registers (thread = row); row sums into l0/l1; 16 packed pairs.
// Packed fp32 (FFMA2 / FADD2, sm_100): the scale-subtract and the row-sum update of a column PAIR are one instruction each —
// 5 instead of 7 issue slots per pair in a loop that is issue- and latency-bound; bit-identical to the scalar form.
template <bool F16, bool MASK>
__device__ __forceinline__ void attn3_exp_half(const uint32_t (&buf)[32], int valid, float c1, float mc, float& l0,
                                               float& l1, uint32_t* pk) {
  const float2 c2 = make_float2(c1, c1), m2 = make_float2(-mc, -mc);
  float2 l = make_float2(l0, l1);
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const float2 t = __ffma2_rn(make_float2(__uint_as_float(buf[i]), __uint_as_float(buf[i + 1])), c2, m2);
    float2 e = make_float2(ex2_approx(t.x), ex2_approx(t.y));
    if constexpr (MASK) {
      e.x = (i < valid) ? e.x : 0.f;
      e.y = (i + 1 < valid) ? e.y : 0.f;
    }
    l = __fadd2_rn(l, e);
    pk[i >> 1] = pack16x2<F16>(e.x, e.y);
  }
  l0 = l.x;
  l1 = l.y;
}

template <bool MASK>
__device__ __forceinline__ float attn3_max_half(const uint32_t (&buf)[32], int valid) {
  float a0 = -INFINITY, a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
  for (int e = 0; e < 32; e += 4) {
    if (!MASK || e < valid) a0 = fmaxf(a0, __uint_as_float(buf[e]));
    if (!MASK || e + 1 < valid) a1 = fmaxf(a1, __uint_as_float(buf[e + 1]));
    if (!MASK || e + 2 < valid) a2 = fmaxf(a2, __uint_as_float(buf[e + 2]));
    if (!MASK || e + 3 < valid) a3 = fmaxf(a3, __uint_as_float(buf[e + 3]));
  }
  return fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
}

template <bool F16, int PIPE>
__global__ void __launch_bounds__(kAttn3Threads, 1)
attention3_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs p) {
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
  // barriers: k_full[5] (one per 128-row K block, so the first QK^T starts after 16 KB, not 80 KB), v_full[5], then per
  // pipeline g (base kPipeBar + 8 g): q_full, q_empty, s_full, s_empty, p_full, p_empty, o_full, o_empty
  constexpr int kPipeBar = 2 * kAttnMaxBlocks;
  uint64_t* k_full = bars + 0;
  uint64_t* v_full = bars + kAttnMaxBlocks;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kPipeBar + 8 * NP);

  const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
  const int n = blockIdx.x / H, h = blockIdx.x % H;
  const int D = H * 64;
  const int q_tiles = (T + 127) / 128;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int j = 0; j < 2 * kAttnMaxBlocks; ++j) mbar_init(bars + j, 1);
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
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(80));  // 40 spills the MMA issue loop: see attention4.cuh
    if (warp == 0) {
      // ===================================================================== TMA
      if (lane == 0) {
        // issue order = order of first use: Q0, K0, Q1, Q2, then K and V blocks alternating
        auto load_q = [&](int g, int t) {
          mbar_arrive_expect_tx(bars + kPipeBar + 8 * g, 16384);
          tma_load_3d(sQ + g * 16384, &tmQKV, bars + kPipeBar + 8 * g, h * 64, t * 128, n);
        };
        auto load_k = [&](int j) {
          mbar_arrive_expect_tx(k_full + j, 16384);
          tma_load_3d(sK + j * 16384, &tmQKV, k_full + j, D + h * 64, j * 128, n);
        };
        auto load_v = [&](int j) {
          mbar_arrive_expect_tx(v_full + j, 16384);
          tma_load_3d(sV + j * 16384, &tmQKV, v_full + j, 2 * D + h * 64, j * 128, n);
        };
        load_q(0, 0);
        load_k(0);
        for (int g = 1; g < NP; ++g)
          if (g < q_tiles) load_q(g, g);
        load_v(0);
        for (int j = 1; j < nb; ++j) {
          load_k(j);
          load_v(j);
        }
        // remaining Q tiles in tile order: tile t waits for the QK^T of tile t - 3 (same pipeline) to retire
        for (int t = NP; t < q_tiles; ++t) {
          const int g = t % NP, k = t / NP;
          mbar_wait(bars + kPipeBar + 8 * g + 1, (k - 1) & 1);
          load_q(g, t);
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
      for (int t = g; t < q_tiles; t += NP, ++k) {
        mbar_wait(q_full, k & 1);
        for (int u = 0; u <= nu; ++u) {
          // QK^T(u) goes first: S(u-1) is already in the softmax warps' registers, so it runs while they exponentiate
          if (u < nu) {
            if (k == 0) mbar_wait(k_full + (u >> 1), 0);  // first tile of the CTA: K block u/2 has landed
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
            if (k == 0) mbar_wait(v_full + (up >> 1), 0);
            mbar_wait(p_full, n_p & 1);
            tc_fence_after();
            if (lane == 0) {
              const uint32_t v_lo = v_lo0 + up * (8192 >> 4);
              const int steps = ((up == nu - 1) ? last_w : 64) / 16;
              for (int s = 0; s < steps; ++s)  // A = P from TMEM: lane = row, 8 columns of packed 16-bit pairs per K=16 step
                umma_bf16_ts(tO, tP + s * 8, smem_desc_join(kHi, v_lo + s * (2048 >> 4)), idesc_o,
                             (up == 0 && s == 0) ? 0u : 1u);
              umma_commit(p_empty);
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
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(144));
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
    for (int t = g; t < q_tiles; t += NP, ++k) {
      float m_used = 0.f;  // row maximum (raw score units) the exponentials are currently taken against
      float l0 = 0.f, l1 = 0.f;
      bool pending = false;  // PIPE >= 1: P(u-1) stored to TMEM but not yet handed off
      for (int u = 0; u < nu; ++u) {
        const int w = (u == nu - 1) ? last_w : 64;
        const int valid = min(64, T - u * 64);
        mbar_wait(s_full, n_s & 1);
        tc_fence_after();
        tmem_ld_32x32b_x32(tS, buf[0]);
        if (w == 64) tmem_ld_32x32b_x32(tS + 32, buf[1]);
        if (PIPE >= 1 && pending) hand_off();  // overlaps the load latency
        tmem_ld_wait_dep(buf[0]);
        {
          if (w == 64) tmem_ld_wait_dep(buf[1]);
          // S unit is in registers: give the buffer back so the next QK^T can start
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty);
          ++n_s;
          // ---- unit maximum
          float hm;
          if (valid == 64) hm = fmaxf(attn3_max_half<false>(buf[0], 32), attn3_max_half<false>(buf[1], 32));
          else hm = (w == 64) ? fmaxf(attn3_max_half<false>(buf[0], 32), attn3_max_half<true>(buf[1], valid - 32))
                              : attn3_max_half<true>(buf[0], valid);
          // ---- online maximum; rescale O only when the maximum moves by more than 2^kAttnRescaleLog2
          if (u == 0) {
            m_used = hm;
          } else {
            const bool need = (hm - m_used) * c1 > kAttnRescaleLog2;
            if (__any_sync(0xffffffffu, need)) rescale(need, hm, m_used, l0, l1, true);
          }
          const float mc = m_used * c1;
          // ---- P = exp2(s*c1 - m*c1), row sum
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
        }
        // ---- hand the unit to the P·V issuer
        mbar_wait(p_empty, (n_p & 1) ^ 1);
        tmem_st_32x32b_x32(tPw, pk);  // 64 P values of this row as 32 packed columns
        ++n_p;
        if constexpr (PIPE >= 1) pending = true;
        else hand_off();
      }
      if (PIPE >= 1 && pending) hand_off();
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
