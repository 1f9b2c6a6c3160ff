// Fused non-causal attention for ViT towers on tcgen05 (head_dim 64, T <= 640 tokens per image).
//
// One CTA per (image, head).  K and V of the head stay resident in shared memory (TMA, 128-B swizzle,
// rows beyond T zero-filled by the 3-D tensor map); query tiles of 128 rows stream through a 2-deep ring.
//   warp 0 (1 thread)   TMA: K blocks, V blocks, Q tiles
//   warp 1 (1 thread)   tcgen05.mma issuer:  S = Q·K_j^T (K-major x K-major)  and  O += P_j·V_j
//                       (P from smem K-major, V as MN-major B operand -> no transpose of V anywhere)
//   warps 2..5          softmax, thread = query row (tcgen05.ld 32x32b): no cross-lane reductions
// Softmax is two-pass per query tile: pass 1 recomputes S blocks only for the exact row maximum, pass 2
// recomputes them for P = exp2((s - m)·scale·log2e).  For T <= 640 the kernel is MUFU(ex2)-bound, so the
// doubled QK^T MMAs are free and no accumulator rescaling (tcgen05.st / correction warps) is needed.
// S lives in TMEM (2 x 128 columns), O in TMEM (2 x 64 columns), P (bf16) in 32 KB of swizzled smem.
#pragma once
#include "ptx.cuh"

namespace lovr {

constexpr int kAttnThreads = 192;
constexpr int kAttnMaxBlocks = 5;  // T <= 640

struct AttnArgs {
  int T;        // tokens per image
  int H;        // heads
  int nb;       // ceil(T / 128)
  int last_w;   // width of the last KV block, multiple of 32
  int nu;       // (second-generation kernel) ceil(T / 64) KV units
  int last_w64; // width of the last 64-column unit: 32 or 64
  float scale;  // head_dim^-0.5
  uint16_t* out;  // [N*T, H*64], 16-bit operand format
  long long* trace;  // debug & 8: clock64 timeline of CTA 0, pipeline 0 (diagnostics only)
  int n_items;    // (persistent kernel) images x heads
  int debug;      // timing experiments only (LOVR_ATTN_DEBUG): 1 = skip pass 1, 2 = no ex2, 4 = skip max/exp math
};

inline size_t attn_smem_bytes(int nb) {
  return static_cast<size_t>(2 * nb + 2 + 2) * 16384 + 256 + 1024;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool F16>
__global__ void __launch_bounds__(kAttnThreads, 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nb = p.nb;
  uint8_t* sK = smem;
  uint8_t* sV = sK + nb * 16384;
  uint8_t* sQ = sV + nb * 16384;  // 2 x 16 KB
  uint8_t* sP = sQ + 2 * 16384;   // 32 KB: two K-major atoms of [128 rows x 64 cols]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * 16384);
  uint64_t* k_full = bars + 0;
  uint64_t* v_full = bars + 1;
  uint64_t* q_full = bars + 2;    // [2]
  uint64_t* q_empty = bars + 4;   // [2]
  uint64_t* s_full = bars + 6;    // [2]
  uint64_t* s_empty = bars + 8;   // [2]
  uint64_t* p_full = bars + 10;
  uint64_t* p_empty = bars + 11;
  uint64_t* o_full = bars + 12;   // [2]
  uint64_t* o_empty = bars + 14;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int D = p.H * 64;
  const int q_tiles = (p.T + 127) / 128;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(k_full, 1);
    mbar_init(v_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);
    }
    mbar_init(p_full, 4);
    mbar_init(p_empty, 1);
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
  const uint32_t tmem_S = tmem_base;        // 2 x 128 columns
  const uint32_t tmem_O = tmem_base + 256;  // 2 x 64 columns

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(k_full, nb * 16384);
      for (int j = 0; j < nb; ++j) tma_load_3d(sK + j * 16384, &tmQKV, k_full, D + h * 64, j * 128, n);
      // first Q tile before V so pass 1 can start as early as possible
      mbar_arrive_expect_tx(&q_full[0], 16384);
      tma_load_3d(sQ, &tmQKV, &q_full[0], h * 64, 0, n);
      mbar_arrive_expect_tx(v_full, nb * 16384);
      for (int j = 0; j < nb; ++j) tma_load_3d(sV + j * 16384, &tmQKV, v_full, 2 * D + h * 64, j * 128, n);
      for (int t = 1; t < q_tiles; ++t) {
        const int qb = t & 1;
        mbar_wait(&q_empty[qb], ((t >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[qb], 16384);
        tma_load_3d(sQ + qb * 16384, &tmQKV, &q_full[qb], h * 64, t * 128, n);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s_full = make_idesc_16(128, 128, 0, 0, F16);
      const uint32_t idesc_s_last = make_idesc_16(128, p.last_w, 0, 0, F16);
      const uint32_t idesc_o = make_idesc_16(128, 64, 0, 1, F16);  // B (=V) is MN-major
      uint32_t sc = 0, pc = 0;
      mbar_wait(k_full, 0);
      auto issue_qk = [&](int qb, int j) {
        const uint32_t sb = sc & 1;
        mbar_wait(&s_empty[sb], ((sc >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(sQ + qb * 16384), ka = smem_u32(sK + j * 16384);
        const uint32_t idesc = (j == nb - 1) ? idesc_s_last : idesc_s_full;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_S + sb * 128, make_smem_desc_sw128(qa + k * 32, 16, 1024),
                       make_smem_desc_sw128(ka + k * 32, 16, 1024), idesc, k != 0);
        umma_commit(&s_full[sb]);
        ++sc;
      };
      auto issue_pv = [&](int ob, int j) {
        mbar_wait(p_full, pc & 1);
        tc_fence_after();
        const int steps = (j == nb - 1) ? p.last_w / 16 : 8;
        const uint32_t pa = smem_u32(sP), va = smem_u32(sV + j * 16384);
        for (int s = 0; s < steps; ++s)
          umma_bf16_ss(tmem_O + ob * 64, make_smem_desc_sw128(pa + (s >> 2) * 16384 + (s & 3) * 32, 16, 1024),
                       make_smem_desc_sw128(va + s * 2048, 16384, 1024), idesc_o, (j | s) != 0);
        umma_commit(p_empty);
        ++pc;
      };
      for (int t = 0; t < q_tiles; ++t) {
        const int qb = t & 1, ob = t & 1;
        mbar_wait(&q_full[qb], (t >> 1) & 1);
        for (int j = 0; j < nb; ++j) issue_qk(qb, j);  // pass 1: row maxima
        if (t == 0) mbar_wait(v_full, 0);
        mbar_wait(&o_empty[ob], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int j = 0; j < nb; ++j) {                 // pass 2: P and O
          issue_qk(qb, j);
          if (j == nb - 1) umma_commit(&q_empty[qb]);
          if (j >= 1) issue_pv(ob, j - 1);
        }
        issue_pv(ob, nb - 1);
        umma_commit(&o_full[ob]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row within the query tile
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const float c1 = p.scale * 1.4426950408889634f;
    const uint32_t row_sw = r & 7;
    uint32_t sc = 0, pc = 0;
    for (int t = 0; t < q_tiles; ++t) {
      const int ob = t & 1;
      // ---------------- pass 1: exact row maximum of the raw scores
      float m = -INFINITY;
      for (int j = 0; j < nb; ++j, ++sc) {
        const uint32_t sb = sc & 1;
        mbar_wait(&s_full[sb], (sc >> 1) & 1);
        tc_fence_after();
        const int nch = ((j == nb - 1) ? p.last_w : 128) >> 5;  // 32-column chunks
        const int valid = min(128, p.T - j * 128);
        const uint32_t ts = tmem_S + lane_sel + sb * 128;
        uint32_t buf[2][32];
        tmem_ld_32x32b_x32(ts, buf[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nch) {  // warp-uniform
            tmem_ld_wait_dep(buf[c & 1]);
            if (c + 1 < nch) tmem_ld_32x32b_x32(ts + (c + 1) * 32, buf[(c + 1) & 1]);  // in flight during the max
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (c * 32 + e < valid) m = fmaxf(m, __uint_as_float(buf[c & 1][e]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[sb]);
      }
      const float mc = m * c1;
      // ---------------- pass 2: P = exp2(s*c1 - m*c1), row sum, P -> smem (bf16, K-major swizzled)
      float l = 0.f;
      for (int j = 0; j < nb; ++j, ++sc, ++pc) {
        const uint32_t sb = sc & 1;
        mbar_wait(&s_full[sb], (sc >> 1) & 1);
        tc_fence_after();
        const int width = (j == nb - 1) ? p.last_w : 128;
        const int nch = width >> 5;
        const int valid = min(128, p.T - j * 128);
        const uint32_t ts = tmem_S + lane_sel + sb * 128;
        uint32_t pk[64];  // the whole P row of this block, packed bf16x2
        uint32_t buf[2][32];
        tmem_ld_32x32b_x32(ts, buf[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nch) {  // warp-uniform
            tmem_ld_wait_dep(buf[c & 1]);
            if (c + 1 < nch) tmem_ld_32x32b_x32(ts + (c + 1) * 32, buf[(c + 1) & 1]);  // overlaps the exps below
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float e0 = ex2_approx(fmaf(__uint_as_float(buf[c & 1][i]), c1, -mc));
              float e1 = ex2_approx(fmaf(__uint_as_float(buf[c & 1][i + 1]), c1, -mc));
              e0 = (c * 32 + i < valid) ? e0 : 0.f;
              e1 = (c * 32 + i + 1 < valid) ? e1 : 0.f;
              l += e0 + e1;
              pk[(c * 32 + i) >> 1] = pack16x2<F16>(e0, e1);
            }
          }
        }
        // S buffer fully read: hand it back before waiting for the P buffer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[sb]);
        mbar_wait(p_empty, (pc & 1) ^ 1);  // previous P·V finished reading the P buffer
#pragma unroll
        for (int cidx = 0; cidx < 16; ++cidx) {  // 16-byte chunk index within the 128-column block
          if (cidx * 8 < width) {
            const uint32_t addr = smem_u32(sP) + (cidx >> 3) * 16384 + r * 128 + (((cidx & 7) ^ row_sw) << 4);
            st_shared_v4(addr, pk[cidx * 4 + 0], pk[cidx * 4 + 1], pk[cidx * 4 + 2], pk[cidx * 4 + 3]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      // ---------------- O / l -> global
      mbar_wait(&o_full[ob], (t >> 1) & 1);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32b_x32(tmem_O + lane_sel + ob * 64, o0);
      tmem_ld_32x32b_x32(tmem_O + lane_sel + ob * 64 + 32, o1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[ob]);
      const int tok = t * 128 + r;
      if (tok < p.T) {
        const float inv = 1.0f / l;
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
