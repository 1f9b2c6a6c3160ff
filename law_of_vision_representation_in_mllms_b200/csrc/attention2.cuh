// Fused non-causal ViT attention on tcgen05, second generation.
//
// Measured facts that shape it (B200, scripts/micro/tmem_bw.cu and profiles/): tcgen05.ld moves ~64 B/cycle per SM
// regardless of how many warps issue it, and one x32 load has ~250-300 cycles of latency under load.  With head_dim 64
// the score tile S (fp32, 4 B per element) is therefore the bottleneck resource: the first-generation kernel read every
// S element twice (max pass + exp pass: 3.4 MB of TMEM reads per (image, head) = 52k of its 88k cycles).  This kernel
// reads S exactly once:
//   * single-pass ONLINE softmax over 64-column KV units, thread = query row;
//   * the running row maximum is only advanced when a unit exceeds it by more than 2^8 (exp2 domain), so the O
//     accumulator in TMEM is rescaled (tcgen05.ld -> scale -> tcgen05.st, by the softmax warp that owns the rows) only
//     on the rare units where that happens; P values are then bounded by 2^8 and the final O / l is exact;
//   * TWO independent pipelines per CTA (query tile t -> pipeline t & 1), each with its own Q/P smem, TMEM region
//     (S_a, S_b double-buffered 64-column units + O), mbarriers and single-thread MMA issuer, so TMEM-load latency,
//     mbarrier wake-ups and MMA round trips of one pipeline hide behind the other.
//
//   warp 0        TMA: K blocks, V blocks (resident for the whole CTA), Q tiles of both pipelines
//   warp 1 / 2    tcgen05.mma issuer of pipeline 0 / 1;  S_u = Q·K_u^T (K-major x K-major, N = 64),
//                 O += P_u·V_u (P from smem K-major, V as MN-major B operand: no transpose of V anywhere)
//   warps 4-7     softmax of pipeline 0      warps 8-11   softmax of pipeline 1   (warp 3 idle)
// Registers are rebalanced with setmaxnreg: the control warpgroup drops to 56, the softmax warpgroups rise to 224.
// TMEM columns of pipeline g: [g*256 + 0, +64) S_a, [+64, +128) S_b, [+128, +192) O, [+192, +224) P (PT variant).
#pragma once
#include "attention.cuh"

namespace lovr {

// NP = number of softmax pipelines per CTA (query tile t -> pipeline t % NP).  NP = 2: S double-buffered per pipeline.
// NP = 3: S single-buffered, P through TMEM only; with 5 query tiles (T = 577) the per-CTA critical path is 2 tile times
// instead of 3, and a third softmax warp per SM sub-partition fills more issue slots.
constexpr int attn2_threads(int np) { return (4 + 4 * np) * 32; }  // warpgroup 0: TMA + MMA issuers; then NP softmax warpgroups
constexpr float kAttnRescaleLog2 = 8.0f;

inline size_t attn2_smem_bytes(int nb, int np, bool pt) {
  return static_cast<size_t>(2 * nb + np + (pt ? 0 : np)) * 16384 + 512 + 1024;
}

#define LOVR_TRACE(cond, idx) do { if ((p.debug & 8) && blockIdx.x == 0 && (cond)) p.trace[(idx)] = clock64(); } while (0)

template <bool F16, bool PT, int NP>  // PT: P handed to the P·V MMA through TMEM (A operand from TMEM) instead of smem
__global__ void __launch_bounds__(attn2_threads(NP), 1)
attention2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nb = p.nb;
  uint8_t* sK = smem;
  uint8_t* sV = sK + nb * 16384;
  static_assert(NP == 2 || (NP == 3 && PT), "3 pipelines need P in TMEM (shared memory is full)");
  constexpr int NBUF = (NP == 2) ? 2 : 1;            // S buffers per pipeline
  constexpr int kPipeCols = (NP == 2) ? 256 : 160;   // TMEM columns per pipeline
  constexpr int kOCol = NBUF * 64, kPCol = NBUF * 64 + 64;
  uint8_t* sQ = sV + nb * 16384;  // [NP pipelines] x 16 KB
  uint8_t* sP = sQ + NP * 16384;  // [NP pipelines] x 16 KB (one K-major atom: 128 rows x 64 columns), unused when PT
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + (PT ? 0 : NP) * 16384);
  uint64_t* k_full = bars + 0;
  uint64_t* v_full = bars + 1;
  auto q_full = [&](int g) { return bars + 2 + 10 * g + 0; };
  auto q_empty = [&](int g) { return bars + 2 + 10 * g + 1; };
  auto s_full = [&](int g, int b) { return bars + 2 + 10 * g + 2 + b; };
  auto s_empty = [&](int g, int b) { return bars + 2 + 10 * g + 4 + b; };
  auto p_full = [&](int g) { return bars + 2 + 10 * g + 6; };
  auto p_empty = [&](int g) { return bars + 2 + 10 * g + 7; };
  auto o_full = [&](int g) { return bars + 2 + 10 * g + 8; };
  auto o_empty = [&](int g) { return bars + 2 + 10 * g + 9; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 10 * NP);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int D = p.H * 64;
  const int q_tiles = (p.T + 127) / 128;
  const int nu = p.nu;  // 64-column KV units
  auto width_of = [&](int u) { return (u == nu - 1) ? p.last_w64 : 64; };

  LOVR_TRACE(threadIdx.x == 0, 4000);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(k_full, 1);
    mbar_init(v_full, 1);
    for (int g = 0; g < NP; ++g) {
      mbar_init(q_full(g), 1);
      mbar_init(q_empty(g), 1);
      for (int b = 0; b < 2; ++b) {
        mbar_init(s_full(g, b), 1);
        mbar_init(s_empty(g, b), 4);
      }
      mbar_init(p_full(g), 4);
      mbar_init(p_empty(g), 1);
      mbar_init(o_full(g), 1);
      mbar_init(o_empty(g), 4);
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

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(NP == 2 ? 56 : 40));
  if (warp == 0) {
    // ===================================================================== TMA
    if (lane == 0) {
      mbar_arrive_expect_tx(k_full, nb * 16384);
      for (int j = 0; j < nb; ++j) tma_load_3d(sK + j * 16384, &tmQKV, k_full, D + h * 64, j * 128, n);
      int next_k[NP] = {};  // per pipeline: index of its next tile (tile = NP*k + g)
      for (int g = 0; g < NP; ++g) {
        if (g < q_tiles) {
          mbar_arrive_expect_tx(q_full(g), 16384);
          tma_load_3d(sQ + g * 16384, &tmQKV, q_full(g), h * 64, g * 128, n);
          next_k[g] = 1;
        }
      }
      mbar_arrive_expect_tx(v_full, nb * 16384);
      for (int j = 0; j < nb; ++j) tma_load_3d(sV + j * 16384, &tmQKV, v_full, 2 * D + h * 64, j * 128, n);
      // remaining Q tiles: poll both pipelines so neither blocks the other
      const long long t0 = clock64();
      for (;;) {
        bool pending = false;
        for (int g = 0; g < NP; ++g) {
          const int t = NP * next_k[g] + g;
          if (t >= q_tiles) continue;
          pending = true;
          if (mbar_try_wait(q_empty(g), (next_k[g] - 1) & 1)) {  // previous tile's QK^T all retired
            mbar_arrive_expect_tx(q_full(g), 16384);
            tma_load_3d(sQ + g * 16384, &tmQKV, q_full(g), h * 64, t * 128, n);
            ++next_k[g];
          }
        }
        if (!pending) break;
        if (clock64() - t0 > LOVR_SPIN_LIMIT_CYCLES) {
          printf("lovr: attention2 TMA poll timed out (block %d)\n", blockIdx.x);
          __trap();
        }
      }
    }
  } else if (warp >= 1 && warp <= NP) {
    // ===================================================================== MMA issuer of pipeline g
    // The whole warp walks the schedule and waits on the mbarriers (warp-uniform control flow, uniform registers);
    // lane 0 issues tcgen05.mma / tcgen05.commit.  Descriptors are stepped with 32-bit adds.
    {
      const int g = warp - 1;
      const uint32_t tS[2] = {tmem_base + g * kPipeCols, tmem_base + g * kPipeCols + (NBUF - 1) * 64};
      const uint32_t tO = tmem_base + g * kPipeCols + kOCol;
      const uint32_t tP = tmem_base + g * kPipeCols + kPCol;
      constexpr uint32_t kHi = smem_desc_hi_sw128(1024);
      const uint32_t q_lo = smem_desc_lo(smem_u32(sQ + g * 16384), 16);
      const uint32_t p_lo = smem_desc_lo(smem_u32(sP + (PT ? 0 : g) * 16384), 16);
      const uint32_t k_lo0 = smem_desc_lo(smem_u32(sK), 16);
      const uint32_t v_lo0 = smem_desc_lo(smem_u32(sV), 16384);
      const uint32_t idesc_o = make_idesc_16(128, 64, 0, 1, F16);  // B (=V) is MN-major
      const uint32_t idesc_s64 = make_idesc_16(128, 64, 0, 0, F16);
      const uint32_t idesc_slast = make_idesc_16(128, p.last_w64, 0, 0, F16);
      uint32_t us[2] = {0, 0};  // fills issued per S buffer
      uint32_t up = 0;          // P units consumed
      int trk = 0;
      auto issue_qk = [&](int u) {
        const int b = u & (NBUF - 1);
        LOVR_TRACE(g == 0 && trk < 2 && lane == 0, 2000 + (trk * 16 + u) * 6 + 0);
        mbar_wait(s_empty(g, b), (us[b] & 1) ^ 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t k_lo = k_lo0 + u * (8192 >> 4);  // unit u = rows [64u, 64u+64) of the resident K
          const uint32_t idesc = (u == nu - 1) ? idesc_slast : idesc_s64;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tS[b], smem_desc_join(kHi, q_lo + 2 * k), smem_desc_join(kHi, k_lo + 2 * k), idesc, k != 0);
          umma_commit(s_full(g, b));
        }
        __syncwarp();
        LOVR_TRACE(g == 0 && trk < 2 && lane == 0, 2000 + (trk * 16 + u) * 6 + 1);
        ++us[b];
      };
      auto issue_pv = [&](int u) {
        LOVR_TRACE(g == 0 && trk < 2 && lane == 0, 2000 + (trk * 16 + u) * 6 + 2);
        mbar_wait(p_full(g), up & 1);
        LOVR_TRACE(g == 0 && trk < 2 && lane == 0, 2000 + (trk * 16 + u) * 6 + 3);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t v_lo = v_lo0 + u * (8192 >> 4);
          const int steps = width_of(u) / 16;
          for (int s = 0; s < steps; ++s) {
            const uint64_t vdesc = smem_desc_join(kHi, v_lo + s * (2048 >> 4));
            if constexpr (PT)  // A = P from TMEM: lane = row, 8 columns of packed 16-bit pairs per K=16 step
              umma_bf16_ts(tO, tP + s * 8, vdesc, idesc_o, (u == 0 && s == 0) ? 0u : 1u);
            else
              umma_bf16_ss(tO, smem_desc_join(kHi, p_lo + 2 * s), vdesc, idesc_o, (u == 0 && s == 0) ? 0u : 1u);
          }
          umma_commit(p_empty(g));
        }
        __syncwarp();
        LOVR_TRACE(g == 0 && trk < 2 && lane == 0, 2000 + (trk * 16 + u) * 6 + 4);
        ++up;
      };
      LOVR_TRACE(g == 0 && lane == 0, 4001);
      mbar_wait(k_full, 0);
      LOVR_TRACE(g == 0 && lane == 0, 4002);
      int k = 0;
      for (int t = g; t < q_tiles; t += NP, ++k) {
        mbar_wait(q_full(g), k & 1);
        int issued = 0;
        issue_qk(issued++);
        if (NBUF == 2 && nu > 1) issue_qk(issued++);
        if (issued == nu && lane == 0) umma_commit(q_empty(g));
        if (k == 0) mbar_wait(v_full, 0);
        mbar_wait(o_empty(g), (k & 1) ^ 1);  // epilogue of this pipeline's previous tile has read O
        tc_fence_after();
        for (int u = 0; u < nu; ++u) {
          // single S buffer: S(u) is already in the softmax warps' registers, so QK^T(u+1) goes first and runs while
          // they exponentiate; double buffer: QK^T(u+2) follows P·V(u)
          if (NBUF == 1 && issued < nu) {
            issue_qk(issued++);
            if (issued == nu && lane == 0) umma_commit(q_empty(g));
          }
          issue_pv(u);
          if (NBUF == 2 && issued < nu) {
            issue_qk(issued++);
            if (issued == nu && lane == 0) umma_commit(q_empty(g));  // every QK^T of this tile issued
          }
        }
        if (lane == 0) umma_commit(o_full(g));
        __syncwarp();
        ++trk;
      }
    }
  }
  } else {
    // ===================================================================== softmax pipelines
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(NP == 2 ? 224 : 152));
    const int g = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS[2] = {tmem_base + g * kPipeCols + lane_sel, tmem_base + g * kPipeCols + (NBUF - 1) * 64 + lane_sel};
    const uint32_t tO = tmem_base + g * kPipeCols + kOCol + lane_sel;
    const uint32_t tPw = tmem_base + g * kPipeCols + kPCol + lane_sel;
    const uint32_t p_row = smem_u32(sP + (PT ? 0 : g) * 16384) + r * 128;
    const uint32_t row_sw = r & 7;
    const float c1 = p.scale * 1.4426950408889634f;
    uint32_t cs[2] = {0, 0};  // S fills consumed per buffer
    uint32_t cp = 0;          // P units produced
    int k = 0;
    for (int t = g; t < q_tiles; t += NP, ++k) {
      float m_used = 0.f;  // row maximum (raw score units) the exponentials are currently taken against
      float l0 = 0.f, l1 = 0.f;
      for (int u = 0; u < nu; ++u) {
        const int b = u & (NBUF - 1);
        const int w = width_of(u);
        const int valid = min(64, p.T - u * 64);
        const bool tr = (g == 0 && warp == 4 && lane == 0 && k < 2);
        LOVR_TRACE(tr, (k * 16 + u) * 8 + 0);
        mbar_wait(s_full(g, b), cs[b] & 1);
        LOVR_TRACE(tr, (k * 16 + u) * 8 + 1);
        tc_fence_after();
        uint32_t buf[2][32];
        tmem_ld_32x32b_x32(tS[b], buf[0]);
        if (w == 64) tmem_ld_32x32b_x32(tS[b] + 32, buf[1]);
        tmem_ld_wait_dep(buf[0]);
        if (w == 64) tmem_ld_wait_dep(buf[1]);
        LOVR_TRACE(tr, (k * 16 + u) * 8 + 2);
        // S unit is in registers: give the buffer back so the next-but-one QK^T can start
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty(g, b));
        ++cs[b];
        // ---- unit maximum
        float hm;
        if (valid == 64) {
          float a0 = -INFINITY, a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            a0 = fmaxf(a0, fmaxf(__uint_as_float(buf[0][e]), __uint_as_float(buf[1][e])));
            a1 = fmaxf(a1, fmaxf(__uint_as_float(buf[0][e + 1]), __uint_as_float(buf[1][e + 1])));
            a2 = fmaxf(a2, fmaxf(__uint_as_float(buf[0][e + 2]), __uint_as_float(buf[1][e + 2])));
            a3 = fmaxf(a3, fmaxf(__uint_as_float(buf[0][e + 3]), __uint_as_float(buf[1][e + 3])));
          }
          hm = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
        } else {
          hm = -INFINITY;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (e < valid) hm = fmaxf(hm, __uint_as_float(buf[0][e]));
            if (w == 64 && 32 + e < valid) hm = fmaxf(hm, __uint_as_float(buf[1][e]));
          }
        }
        // ---- online maximum; rescale O only when the maximum moves by more than 2^kAttnRescaleLog2
        if (u == 0) {
          m_used = hm;
        } else {
          const bool need = (hm - m_used) * c1 > kAttnRescaleLog2;
          if (__any_sync(0xffffffffu, need)) {
            mbar_wait(p_empty(g), (cp & 1) ^ 1);  // every P·V issued so far has landed in O
            tc_fence_after();
            const float alpha = need ? ex2_approx((m_used - hm) * c1) : 1.0f;
            if (need) m_used = hm;
            l0 *= alpha;
            l1 *= alpha;
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
        }
        const float mc = m_used * c1;
        // ---- P = exp2(s*c1 - m*c1), row sum
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) pk[i] = 0u;
        if (valid == 64) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float e0 = fmaf(__uint_as_float(buf[c][i]), c1, -mc);
              float e1 = fmaf(__uint_as_float(buf[c][i + 1]), c1, -mc);
              if (!(p.debug & 2)) { e0 = ex2_approx(e0); e1 = ex2_approx(e1); }
              l0 += e0;
              l1 += e1;
              pk[c * 16 + (i >> 1)] = pack16x2<F16>(e0, e1);
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (c * 32 < w) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                float e0 = ex2_approx(fmaf(__uint_as_float(buf[c][i]), c1, -mc));
                float e1 = ex2_approx(fmaf(__uint_as_float(buf[c][i + 1]), c1, -mc));
                e0 = (c * 32 + i < valid) ? e0 : 0.f;
                e1 = (c * 32 + i + 1 < valid) ? e1 : 0.f;
                l0 += e0;
                l1 += e1;
                pk[c * 16 + (i >> 1)] = pack16x2<F16>(e0, e1);
              }
            }
          }
        }
        // ---- hand the unit to the P·V issuer
        LOVR_TRACE(tr, (k * 16 + u) * 8 + 3);
        mbar_wait(p_empty(g), (cp & 1) ^ 1);
        LOVR_TRACE(tr, (k * 16 + u) * 8 + 4);
        if constexpr (PT) {
          tmem_st_32x32b_x32(tPw, pk);  // 64 P values of this row as 32 packed columns
          tmem_st_wait();
          tc_fence_before();
        } else {
          const int nq = w >> 3;  // 16-byte chunks
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q < nq) st_shared_v4(p_row + ((q ^ row_sw) << 4), pk[q * 4 + 0], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
          fence_proxy_async_smem();  // measured ~700 cycles per unit: the reason for the TMEM variant
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(g));
        LOVR_TRACE(tr, (k * 16 + u) * 8 + 5);
        ++cp;
      }
      // ---------------- O / l -> global
      LOVR_TRACE(g == 0 && warp == 4 && lane == 0 && k < 2, (k * 16 + 15) * 8 + 6);
      mbar_wait(o_full(g), k & 1);
      LOVR_TRACE(g == 0 && warp == 4 && lane == 0 && k < 2, (k * 16 + 15) * 8 + 7);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32b_x32(tO, o0);
      tmem_ld_32x32b_x32(tO + 32, o1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty(g));
      const int tok = t * 128 + r;
      if (tok < p.T) {
        const float inv = 1.0f / (l0 + l1);
        uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(n) * p.T + tok) * D + h * 64);
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
  LOVR_TRACE(threadIdx.x == 0, 4003);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace lovr
