// Fused non-causal ViT attention on tcgen05, second generation: TWO independent softmax pipelines per CTA.
//
// One CTA per (image, head); K and V of the head stay resident in shared memory.  Query tiles of 128 rows alternate
// between two warpgroup pipelines (tile t -> pipeline t & 1); each pipeline has its own Q buffer, P buffer, TMEM
// region, mbarriers and its own single-thread tcgen05.mma issuer, so the tensor-core round trips, mbarrier wake-ups
// and the MUFU-bound exp work of one pipeline overlap with the other's (the first-generation kernel, one pipeline,
// was latency-bound: XU pipe 26 %, tensor pipe 19 %).
//
//   warp 0        TMA: K blocks, V blocks, Q tiles of both pipelines
//   warp 1 / 2    MMA issuer of pipeline 0 / 1 (one thread each); warp 1 also owns TMEM alloc/dealloc
//   warps 3-6     softmax of pipeline 0 (thread = query row via tcgen05.ld 32x32b)
//   warps 7-10    softmax of pipeline 1
//
// Per pipeline, TMEM columns [g*256, g*256+256): S0 = +0..127, S1 = +128..255, O = +128..191 (aliases S1).
//   pass 1 (exact row max): S double-buffered (S0/S1) — only max is taken, so the QK^T latency must be hidden
//   pass 2 (P = exp2((s-m)*scale*log2e), O += P·V): S single-buffered in S0 (exp-bound), O accumulates in +128..191;
//           P goes through one 16 KB swizzled smem atom per pipeline in 64-column halves.
// Two-pass softmax needs no accumulator rescaling; the doubled QK^T MMAs are free in a MUFU-bound kernel.
#pragma once
#include "attention.cuh"

namespace lovr {

constexpr int kAttn2Threads = 11 * 32;

inline size_t attn2_smem_bytes(int nb) { return static_cast<size_t>(2 * nb + 2 + 2) * 16384 + 512 + 1024; }

template <bool F16>
__global__ void __launch_bounds__(kAttn2Threads, 1)
attention2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nb = p.nb;
  uint8_t* sK = smem;
  uint8_t* sV = sK + nb * 16384;
  uint8_t* sQ = sV + nb * 16384;  // [2 pipelines] x 16 KB
  uint8_t* sP = sQ + 2 * 16384;   // [2 pipelines] x 16 KB (one K-major atom: 128 rows x 64 columns)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * 16384);
  uint64_t* k_full = bars + 0;
  uint64_t* v_full = bars + 1;
  // per pipeline g: 10 barriers starting at bars + 2 + 10*g
  auto q_full = [&](int g) { return bars + 2 + 10 * g + 0; };
  auto q_empty = [&](int g) { return bars + 2 + 10 * g + 1; };
  auto s_full = [&](int g, int b) { return bars + 2 + 10 * g + 2 + b; };
  auto s_empty = [&](int g, int b) { return bars + 2 + 10 * g + 4 + b; };
  auto p_full = [&](int g) { return bars + 2 + 10 * g + 6; };
  auto p_empty = [&](int g) { return bars + 2 + 10 * g + 7; };
  auto o_full = [&](int g) { return bars + 2 + 10 * g + 8; };
  auto o_empty = [&](int g) { return bars + 2 + 10 * g + 9; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int D = p.H * 64;
  const int q_tiles = (p.T + 127) / 128;
  auto width_of = [&](int j) { return (j == nb - 1) ? p.last_w : 128; };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(k_full, 1);
    mbar_init(v_full, 1);
    for (int g = 0; g < 2; ++g) {
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

  if (warp == 0) {
    // ===================================================================== TMA
    if (lane == 0) {
      mbar_arrive_expect_tx(k_full, nb * 16384);
      for (int j = 0; j < nb; ++j) tma_load_3d(sK + j * 16384, &tmQKV, k_full, D + h * 64, j * 128, n);
      int next_k[2] = {0, 0};  // per pipeline: index of its next tile (tile = 2*k + g)
      for (int g = 0; g < 2; ++g) {
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
        for (int g = 0; g < 2; ++g) {
          const int t = 2 * next_k[g] + g;
          if (t >= q_tiles) continue;
          pending = true;
          // the (next_k-1)-th use of the buffer has been released when q_empty completed next_k times
          if (mbar_try_wait(q_empty(g), (next_k[g] - 1) & 1)) {
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
  } else if (warp == 1 || warp == 2) {
    // ===================================================================== MMA issuer of pipeline g
    if (lane == 0) {
      const int g = warp - 1;
      const uint32_t tS[2] = {tmem_base + g * 256, tmem_base + g * 256 + 128};
      const uint32_t tO = tmem_base + g * 256 + 128;
      const uint32_t qa = smem_u32(sQ + g * 16384), pa = smem_u32(sP + g * 16384);
      const uint32_t idesc_o = make_idesc_16(128, 64, 0, 1, F16);  // B (=V) is MN-major
      uint32_t us[2] = {0, 0};  // fills issued per S buffer
      uint32_t up = 0;          // P halves consumed
      auto issue_qk = [&](int b, int j) {
        mbar_wait(s_empty(g, b), (us[b] & 1) ^ 1);
        tc_fence_after();
        const uint32_t ka = smem_u32(sK + j * 16384);
        const uint32_t idesc = make_idesc_16(128, width_of(j), 0, 0, F16);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tS[b], make_smem_desc_sw128(qa + k * 32, 16, 1024), make_smem_desc_sw128(ka + k * 32, 16, 1024),
                       idesc, k != 0);
        umma_commit(s_full(g, b));
        ++us[b];
      };
      auto issue_pv = [&](int j, int half, bool first) {
        mbar_wait(p_full(g), up & 1);
        tc_fence_after();
        const int cols = min(64, width_of(j) - 64 * half);
        const uint32_t va = smem_u32(sV + j * 16384) + half * 8192;
        for (int s = 0; s < cols / 16; ++s)
          umma_bf16_ss(tO, make_smem_desc_sw128(pa + s * 32, 16, 1024), make_smem_desc_sw128(va + s * 2048, 16384, 1024),
                       idesc_o, (first && s == 0) ? 0u : 1u);
        umma_commit(p_empty(g));
        ++up;
      };
      mbar_wait(k_full, 0);
      int k = 0;
      for (int t = g; t < q_tiles; t += 2, ++k) {
        mbar_wait(q_full(g), k & 1);
        // ---- pass 1: S double-buffered; buffer 1 aliases O of the previous tile -> wait for its epilogue
        if (p.debug & 1) {
          mbar_wait(o_empty(g), (k & 1) ^ 1);
        } else {
          for (int j = 0; j < nb; ++j) {
            if (j == 1) mbar_wait(o_empty(g), (k & 1) ^ 1);
            issue_qk(j & 1, j);
          }
        }
        // ---- pass 2
        if (k == 0) mbar_wait(v_full, 0);
        issue_qk(0, 0);
        // O aliases S1: every pass-1 read of S1 must be complete before the first P·V lands there
        mbar_wait(s_empty(g, 1), (us[1] & 1) ^ 1);
        if (nb == 1) mbar_wait(o_empty(g), (k & 1) ^ 1);  // (no pass-1 use of S1 happened: still wait for the epilogue)
        tc_fence_after();
        for (int j = 0; j < nb; ++j) {
          issue_pv(j, 0, j == 0);
          if (j + 1 < nb) issue_qk(0, j + 1);
          else umma_commit(q_empty(g));  // all QK^T of this tile issued: Q buffer reusable when they retire
          if (width_of(j) > 64) issue_pv(j, 1, false);
        }
        umma_commit(o_full(g));
      }
    }
  } else {
    // ===================================================================== softmax pipelines
    const int g = (warp - 3) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS[2] = {tmem_base + g * 256 + lane_sel, tmem_base + g * 256 + 128 + lane_sel};
    const uint32_t tO = tmem_base + g * 256 + 128 + lane_sel;
    const uint32_t p_row = smem_u32(sP + g * 16384) + r * 128;
    const uint32_t row_sw = r & 7;
    const float c1 = p.scale * 1.4426950408889634f;
    uint32_t cs[2] = {0, 0};  // S fills consumed per buffer
    uint32_t cp = 0;          // P halves produced
    int k = 0;
    for (int t = g; t < q_tiles; t += 2, ++k) {
      // ---------------- pass 1: exact row maximum of the raw scores
      float m0 = (p.debug & 1) ? 0.f : -INFINITY, m1 = m0, m2 = m0, m3 = m0;
      for (int j = 0; j < ((p.debug & 1) ? 0 : nb); ++j) {
        const int b = j & 1;
        mbar_wait(s_full(g, b), cs[b] & 1);
        tc_fence_after();
        const int nch = width_of(j) >> 5;
        const int valid = min(128, p.T - j * 128);
        uint32_t buf[2][32];
        tmem_ld_32x32b_x32(tS[b], buf[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nch) {
            tmem_ld_wait_dep(buf[c & 1]);
            if (c + 1 < nch) tmem_ld_32x32b_x32(tS[b] + (c + 1) * 32, buf[(c + 1) & 1]);
            if (c * 32 + 32 <= valid) {  // warp-uniform: no per-element masking away from the sequence tail
#pragma unroll
              for (int e = 0; e < 32; e += 4) {  // four independent chains instead of one 4-cycle-latency chain
                m0 = fmaxf(m0, __uint_as_float(buf[c & 1][e]));
                m1 = fmaxf(m1, __uint_as_float(buf[c & 1][e + 1]));
                m2 = fmaxf(m2, __uint_as_float(buf[c & 1][e + 2]));
                m3 = fmaxf(m3, __uint_as_float(buf[c & 1][e + 3]));
              }
            } else {
#pragma unroll
              for (int e = 0; e < 32; ++e)
                if (c * 32 + e < valid) m0 = fmaxf(m0, __uint_as_float(buf[c & 1][e]));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty(g, b));
        ++cs[b];
      }
      const float mc = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * c1;
      // ---------------- pass 2: P = exp2(s*c1 - m*c1) in 64-column halves, row sum
      float l0 = 0.f, l1 = 0.f;
      for (int j = 0; j < nb; ++j) {
        mbar_wait(s_full(g, 0), cs[0] & 1);
        tc_fence_after();
        const int nch = width_of(j) >> 5;
        const int valid = min(128, p.T - j * 128);
        uint32_t buf[2][32];
        uint32_t pk[32];  // one 64-column half, packed 16-bit pairs
        tmem_ld_32x32b_x32(tS[0], buf[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nch) {
            tmem_ld_wait_dep(buf[c & 1]);
            if (c + 1 < nch) {
              tmem_ld_32x32b_x32(tS[0] + (c + 1) * 32, buf[(c + 1) & 1]);
            } else {
              // every TMEM read of this S block has completed: release it so the next QK^T overlaps the exps
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(s_empty(g, 0));
            }
            if (c * 32 + 32 <= valid) {  // warp-uniform fast path: FFMA + MUFU.EX2 + FADD + half a pack per element
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                float e0 = fmaf(__uint_as_float(buf[c & 1][i]), c1, -mc);
                float e1 = fmaf(__uint_as_float(buf[c & 1][i + 1]), c1, -mc);
                if (!(p.debug & 2)) { e0 = ex2_approx(e0); e1 = ex2_approx(e1); }
                l0 += e0;
                l1 += e1;
                pk[(c & 1) * 16 + (i >> 1)] = pack16x2<F16>(e0, e1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                float e0 = ex2_approx(fmaf(__uint_as_float(buf[c & 1][i]), c1, -mc));
                float e1 = ex2_approx(fmaf(__uint_as_float(buf[c & 1][i + 1]), c1, -mc));
                e0 = (c * 32 + i < valid) ? e0 : 0.f;
                e1 = (c * 32 + i + 1 < valid) ? e1 : 0.f;
                l0 += e0;
                l1 += e1;
                pk[(c & 1) * 16 + (i >> 1)] = pack16x2<F16>(e0, e1);
              }
            }
            if ((c & 1) == 1 || c == nch - 1) {  // a half is complete -> hand it to the P·V issuer
              mbar_wait(p_empty(g), (cp & 1) ^ 1);
              const int nq = ((c & 1) == 1) ? 8 : 4;  // 16-byte chunks in this half
#pragma unroll
              for (int q = 0; q < 8; ++q)
                if (q < nq)
                  st_shared_v4(p_row + ((q ^ row_sw) << 4), pk[q * 4 + 0], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) mbar_arrive(p_full(g));
              ++cp;
            }
          }
        }
        ++cs[0];
      }
      // ---------------- O / l -> global
      mbar_wait(o_full(g), k & 1);
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
          uint4 w;
          const uint32_t* src = (q < 4) ? &o0[q * 8] : &o1[(q - 4) * 8];
          w.x = pack16x2<F16>(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
          w.y = pack16x2<F16>(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
          w.z = pack16x2<F16>(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
          w.w = pack16x2<F16>(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
          dst[q] = w;
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
