// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] · W[N,K]^T)
//
//   warp 0 (1 thread)  TMA producer: A tile 128x64 and W tile BNx64 (bf16, K-major, 128-B swizzle)
//                      into a kStages-deep smem ring, completion on `full` mbarriers.
//   warp 1 (1 thread)  MMA issuer: tcgen05.mma cta_group::1 kind::f16, 128 x BN x 16 per instruction,
//                      fp32 accumulator in TMEM, double-buffered (2 x BN columns); tcgen05.commit
//                      releases smem stages (`empty`) and publishes accumulators (`tmem_full`).
//   warps 2..5         epilogue: tcgen05.ld (thread = accumulator row), fused bias / activation /
//                      LayerScale / residual / row-max, 128-B-swizzled smem staging, TMA store.
//
// Tiles are visited unit-major: unit = (batch, m_tile, n_group); each unit runs `n_inner` consecutive
// n tiles (n_inner = 1 for the dense layers; = all n tiles for the row-max epilogue so the running
// maximum stays in registers).  Rows beyond M are zero-filled by TMA on load and clipped on store.
#pragma once
#include "ptx.cuh"

namespace lovr {

enum : int { ACT_NONE = 0, ACT_QUICK_GELU = 1, ACT_GELU_ERF = 2, ACT_GELU_TANH = 3 };
enum : int {
  EPI_BF16 = 0,      // C 16-bit (operand format) = act(acc + bias)
  EPI_RESID_F32 = 1, // C fp32 += colscale * (acc + bias)          (in place residual stream)
  EPI_F32 = 2,       // C fp32  = acc + bias
  EPI_ROWMAX = 3     // rowmax[b,row] = rowscale[b,row] * max_col(acc * colscale[b,col]), cols < n_valid
};

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kStages = 4;
constexpr int kEpiWarps = 4;
constexpr int kGemmThreads = 64 + kEpiWarps * 32;
constexpr int kStageBufBytes = 32 * 128;  // one epilogue staging box: 32 rows x 128 B

struct GemmArgs {
  int M;            // rows of A per batch entry
  int N;            // columns of C (= rows of W) per batch entry, multiple of BN unless EPI_ROWMAX
  int K;            // multiple of 64
  int batch;        // >= 1; >1 uses the 3rd tensor-map coordinate
  int n_inner;      // n tiles per unit
  const float* bias;      // [N] or nullptr
  const float* colscale;  // EPI_RESID_F32: [N] LayerScale or nullptr;  EPI_ROWMAX: [batch, N] 1/||ref||
  const float* rowscale;  // EPI_ROWMAX: [batch, M] 1/||other||
  float* rowmax;          // EPI_ROWMAX: [batch, M]
  int n_valid;            // EPI_ROWMAX: real number of columns (<= N)
  int scale_ld;           // EPI_F32 with colscale/rowscale: per-batch stride of both scale vectors (padded)
  void* c_ptr;            // EPI_RESID_F32: base of C (fp32), row stride ldc elements, batch stride c_batch_stride
  long long ldc;
  long long c_batch_stride;
  // ---- LayerNorm folded into the GEMMs around it (no LayerNorm kernel, no second pass over the residual stream):
  //   LN(x) W^T + b  =  rstd (x^ W'^T - mu cw) + bw,   W' = W diag(gamma) (16-bit), cw = row sums of W', bw = W beta + b
  // consumer (EPI_BF16; A = x^, the 16-bit copy of the residual stream, W = W', bias = bw):
  const float* ln_stats;  // [M][N_x / 128][4] per-row partial statistics of x or nullptr (see kLnPartCols)
  const float* ln_cw;     // [N]
  int ln_parts;           // N_x / 128
  float ln_eps;
  // producer (EPI_RESID_F32): besides the fp32 update of x also emit x^ and the partial statistics of the new x
  float* stats_out;       // [M][N / 128][4] or nullptr
  void* xhat_out;         // 16-bit [M, N], row stride ld_xhat elements, or nullptr
  long long ld_xhat;
};

// Row statistics travel as one slot per 128 columns: (shift, sum(x - shift), sum((x - shift)^2), unused) with shift = the
// slot's first element — sums about a value inside the data instead of about zero, so the variance does not cancel when a
// row has a large mean; slots are merged with the pairwise (Chan) update.
constexpr int kLnPartCols = 128;

__device__ __forceinline__ void ln_row_scales(const GemmArgs& p, int row, float& rs, float& rm) {
  rs = 1.0f;
  rm = 0.0f;
  if (p.ln_stats == nullptr) return;
  const float4* st = reinterpret_cast<const float4*>(p.ln_stats) + static_cast<size_t>(row) * p.ln_parts;
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (int q = 0; q < p.ln_parts; ++q) {
    const float4 v = __ldg(st + q);
    const float nb = static_cast<float>(kLnPartCols);
    const float mb = v.x + v.y * (1.0f / nb);
    const float m2b = v.z - v.y * v.y * (1.0f / nb);
    const float delta = mb - mean;
    const float nn = n + nb;
    mean += delta * (nb / nn);
    m2 += m2b + delta * delta * (n * nb / nn);
    n = nn;
  }
  const float rstd = rsqrtf(m2 / n + p.ln_eps);
  rs = rstd;
  rm = -rstd * mean;
}

template <int BN>
struct GemmSmem {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = kEpiWarps * 2 * kStageBufBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kStages * kStageBytes + kStagingBytes + kBarBytes + 1024 /*align slack*/;
};

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Activations of the fc1 epilogue.  sigmoid(z) = (1 + tanh(z/2)) / 2, so quick_gelu and gelu_tanh each cost ONE MUFU
// (tanh.approx, relative error 2^-11, below the 16-bit output rounding) instead of ex2 + rcp: the epilogue of the fc1
// GEMM was XU-bound (38 % of the pipe).
template <int act>
__device__ __forceinline__ float apply_act(float x) {
  if constexpr (act == ACT_QUICK_GELU) {
    // x * sigmoid(1.702 x) = 0.5 x (1 + tanh(0.851 x))
    const float h = 0.5f * x;
    return fmaf(h, tanh_approx(0.851f * x), h);
  } else if constexpr (act == ACT_GELU_ERF) {
    // x * Phi(x), Phi(-|x|) = erfc(|x|/sqrt2)/2 by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 on erf; the tail is formed
    // as polynomial * exp, so small Phi keeps its relative accuracy): 2 MUFU + ~12 FMA-pipe instructions instead of the
    // ~25-instruction branchy erff — the fc1 epilogue of DINOv2 / OpenCLIP was FMA-bound.
    const float z = fabsf(x) * 0.70710678118654752440f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float poly = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
    poly = fmaf(poly, t, 0.5f * 1.421413741f);
    poly = fmaf(poly, t, 0.5f * -0.284496736f);
    poly = fmaf(poly, t, 0.5f * 0.254829592f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
    const float q = poly * t * e;                 // Phi(-|x|)
    return x * (x < 0.f ? q : 1.0f - q);
  } else if constexpr (act == ACT_GELU_TANH) {
    const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
    const float h = 0.5f * x;
    return fmaf(h, tanh_approx(u), h);
  }
  return x;
}

// Epilogue of one 128-row x BN-column accumulator tile for one epilogue warp (its 32 TMEM lanes = 32 output rows):
// shared by the 1-CTA and the 2-CTA (cta_group::2) kernels.
template <int BN, int EPI, int ACT, bool F16, int RB = 2>  // RB: staging buffers per warp used by the residual epilogue
__device__ __forceinline__ void epilogue_tile(const uint32_t t_row, const int col_tile, const int row0, const int b,
                                              const GemmArgs& p, const CUtensorMap& tmC, uint8_t* my_stage,
                                              uint64_t* my_rbar, const int lane, const uint32_t row_sw, uint32_t& g,
                                              float& run_max, const float rs = 1.0f, const float rm = 0.0f) {
    if constexpr (EPI == EPI_BF16) {
      constexpr int CH = 64;
      // folded LayerNorm: x = rs * acc + (rm * cw + bias); without it rs = 1, rm = 0 and the result is acc + bias exactly.
      // (rs, rm come from the caller, which merges the row's statistics slots BEFORE it waits for the accumulator: the
      // slot loads then overlap the tile's MMAs instead of sitting at the head of the epilogue)
#pragma unroll 1
      for (int c = 0; c < BN / CH; ++c, ++g) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(t_row + c * CH, r0);
        tmem_ld_32x32b_x32(t_row + c * CH + 32, r1);
        uint8_t* buf = my_stage + (g & 1) * kStageBufBytes;
        if (lane == 0) tma_store_wait_read<1>();  // the store issued 2 chunks ago has drained this buffer
        __syncwarp();
        tmem_ld_wait();
        const int col0 = col_tile + c * CH;
        const uint32_t buf_row = smem_u32(buf) + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v[8];
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
          if (p.bias) {
            b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j * 8));
            b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j * 8 + 4));
          }
          float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
          if (p.ln_cw) {
            w0 = __ldg(reinterpret_cast<const float4*>(p.ln_cw + col0 + j * 8));
            w1 = __ldg(reinterpret_cast<const float4*>(p.ln_cw + col0 + j * 8 + 4));
          }
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          const float2 rs2 = make_float2(rs, rs), rm2 = make_float2(rm, rm);
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const int cc = j * 8 + e;
            const float2 a2 = make_float2(__uint_as_float(cc < 32 ? r0[cc] : r1[cc - 32]),
                                          __uint_as_float(cc < 32 ? r0[cc + 1] : r1[cc + 1 - 32]));
            const float2 x2 = __ffma2_rn(rs2, a2, __ffma2_rn(rm2, make_float2(ww[e], ww[e + 1]), make_float2(bb[e], bb[e + 1])));
            v[e] = apply_act<ACT>(x2.x);
            v[e + 1] = apply_act<ACT>(x2.y);
          }
          st_shared_v4(buf_row + ((j ^ row_sw) << 4), pack16x2<F16>(v[0], v[1]), pack16x2<F16>(v[2], v[3]),
                       pack16x2<F16>(v[4], v[5]), pack16x2<F16>(v[6], v[7]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmC, buf, col0, row0, b);
          tma_store_commit();
        }
      }
    } else if constexpr (EPI == EPI_F32) {
      constexpr int CH = 32;
#pragma unroll 1
      for (int c = 0; c < BN / CH; ++c, ++g) {
        uint32_t r0[32];
        tmem_ld_32x32b_x32(t_row + c * CH, r0);
        uint8_t* buf = my_stage + (g & 1) * kStageBufBytes;
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        tmem_ld_wait();
        const int col0 = col_tile + c * CH;
        const uint32_t buf_row = smem_u32(buf) + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v[4];
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias) b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j * 4));
          const float bb[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = __uint_as_float(r0[j * 4 + e]) + bb[e];
          if (p.colscale) {  // similarity GEMMs: C = rowscale[b,row] * colscale[b,col] * acc (padded to BN)
            const float4 s0 =
                __ldg(reinterpret_cast<const float4*>(p.colscale + static_cast<size_t>(b) * p.scale_ld + col0 + j * 4));
            const float rs = __ldg(p.rowscale + static_cast<size_t>(b) * p.scale_ld + min(row0 + lane, p.M - 1));
            v[0] *= s0.x * rs;
            v[1] *= s0.y * rs;
            v[2] *= s0.z * rs;
            v[3] *= s0.w * rs;
          }
          st_shared_v4(buf_row + ((j ^ row_sw) << 4), __float_as_uint(v[0]), __float_as_uint(v[1]),
                       __float_as_uint(v[2]), __float_as_uint(v[3]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmC, buf, col0, row0, b);
          tma_store_commit();
        }
      }
    } else if constexpr (EPI == EPI_RESID_F32) {
      // x[row, col] += colscale * (acc + bias) in place.  The residual chunk (32 rows x 128 B) is TMA-loaded into the
      // same swizzled staging buffer the result is stored from, PD chunks ahead of its use; with RB staging buffers per
      // warp, PD = RB/2 loads and RB/2 stores are in flight.
      constexpr int CH = 32;
      constexpr int NCH = BN / CH;
      constexpr int PD = RB / 2;
      static_assert(RB == 2 || RB == 4, "residual staging depth");
      if (lane == 0) {
        tma_store_wait_read<0>();  // tile boundary: every staging buffer of this warp is free again
#pragma unroll
        for (int i = 0; i < PD; ++i) {
          const uint32_t gi = g + i;
          uint8_t* buf = my_stage + (gi % RB) * kStageBufBytes;
          mbar_arrive_expect_tx(&my_rbar[gi % RB], kStageBufBytes);
          tma_load_3d(buf, &tmC, &my_rbar[gi % RB], col_tile + i * CH, row0, b);
        }
      }
      float2 st_nshift = make_float2(0.f, 0.f), st_s1 = st_nshift, st_s2 = st_nshift;  // two lanes of partial sums
      const bool row_ok = row0 + lane < p.M;
#pragma unroll 1
      for (int c = 0; c < NCH; ++c, ++g) {
        uint32_t r0[32];
        tmem_ld_32x32b_x32(t_row + c * CH, r0);
        if (lane == 0 && c + PD < NCH) {
          // buffer (g+PD) % RB was last read by the store of chunk g+PD-RB: allow the PD-1 younger stores to stay in flight
          tma_store_wait_read<PD - 1>();
          const uint32_t gi = g + PD;
          uint8_t* nbuf = my_stage + (gi % RB) * kStageBufBytes;
          mbar_arrive_expect_tx(&my_rbar[gi % RB], kStageBufBytes);
          tma_load_3d(nbuf, &tmC, &my_rbar[gi % RB], col_tile + (c + PD) * CH, row0, b);
        }
        uint8_t* buf = my_stage + (g % RB) * kStageBufBytes;
        mbar_wait(&my_rbar[g % RB], (g / RB) & 1);
        tmem_ld_wait();
        const int col0 = col_tile + c * CH;
        const uint32_t buf_row = smem_u32(buf) + lane * 128;
        uint32_t xh[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t q0, q1, q2, q3;
          const uint32_t addr = buf_row + ((j ^ row_sw) << 4);
          ld_shared_v4(addr, q0, q1, q2, q3);
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), s0 = make_float4(1.f, 1.f, 1.f, 1.f);
          if (p.bias) b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j * 4));
          if (p.colscale) s0 = __ldg(reinterpret_cast<const float4*>(p.colscale + col0 + j * 4));
          // packed fp32 (FADD2 / FFMA2): two elements per instruction, the same two roundings per element as the scalar form
          float2 ra = make_float2(__uint_as_float(q0), __uint_as_float(q1)), rb = make_float2(__uint_as_float(q2), __uint_as_float(q3));
          ra = __ffma2_rn(__fadd2_rn(make_float2(__uint_as_float(r0[j * 4]), __uint_as_float(r0[j * 4 + 1])), make_float2(b0.x, b0.y)),
                          make_float2(s0.x, s0.y), ra);
          rb = __ffma2_rn(__fadd2_rn(make_float2(__uint_as_float(r0[j * 4 + 2]), __uint_as_float(r0[j * 4 + 3])), make_float2(b0.z, b0.w)),
                          make_float2(s0.z, s0.w), rb);
          st_shared_v4(addr, __float_as_uint(ra.x), __float_as_uint(ra.y), __float_as_uint(rb.x), __float_as_uint(rb.y));
          if (p.stats_out != nullptr) {  // (warp-uniform) statistics of the NEW residual row for the next folded LayerNorm
            if ((c & 3) == 0 && j == 0) st_nshift = make_float2(-ra.x, -ra.x);
            const float2 da = __fadd2_rn(ra, st_nshift), db = __fadd2_rn(rb, st_nshift);
            st_s1 = __fadd2_rn(__fadd2_rn(st_s1, da), db);
            st_s2 = __ffma2_rn(db, db, __ffma2_rn(da, da, st_s2));
            xh[2 * j] = pack16x2<F16>(ra.x, ra.y);
            xh[2 * j + 1] = pack16x2<F16>(rb.x, rb.y);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmC, buf, col0, row0, b);
          tma_store_commit();
        }
        if (p.stats_out != nullptr && row_ok) {
          // x^: this thread's 32 consecutive 16-bit values (64 contiguous bytes, whole 32-byte sectors)
          uint4* dst = reinterpret_cast<uint4*>(static_cast<uint16_t*>(p.xhat_out) + static_cast<size_t>(row0 + lane) * p.ld_xhat + col0);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q] = make_uint4(xh[4 * q], xh[4 * q + 1], xh[4 * q + 2], xh[4 * q + 3]);
          if ((c & 3) == 3) {  // 128 columns done: one slot
            reinterpret_cast<float4*>(p.stats_out)[static_cast<size_t>(row0 + lane) * (p.N / kLnPartCols) + col0 / kLnPartCols] =
                make_float4(-st_nshift.x, st_s1.x + st_s1.y, st_s2.x + st_s2.y, 0.f);
            st_s1 = make_float2(0.f, 0.f);
            st_s2 = st_s1;
          }
        }
      }
    // (A variant that replaced the TMA load/store of the residual chunk by a coalesced ld.global/st.global pass through
    //  a shared-memory transpose — no async-proxy fences — was measured SLOWER: out-proj 685 vs 855 TFLOP/s, because the
    //  residual loads then sit on the warp's critical path; the TMA version prefetches them one chunk ahead.)
    } else {  // EPI_ROWMAX
      constexpr int CH = 32;
      const float* cs = p.colscale + static_cast<size_t>(b) * p.N;
#pragma unroll 1
      for (int c = 0; c < BN / CH; ++c) {
        const int col0 = col_tile + c * CH;
        if (col0 >= p.n_valid) break;  // warp-uniform
        uint32_t r0[32];
        tmem_ld_32x32b_x32(t_row + c * CH, r0);
        tmem_ld_wait();
#pragma unroll
        for (int cc = 0; cc < 32; ++cc) {
          const float x = __uint_as_float(r0[cc]) * __ldg(cs + min(col0 + cc, p.N - 1));
          if (col0 + cc < p.n_valid) run_max = fmaxf(run_max, x);
        }
      }
    }
}

template <int BN, int EPI, int ACT, bool F16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const GemmArgs p) {
  using L = GemmSmem<BN>;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N constraint for M=128");
  constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  constexpr uint32_t kIdesc = make_idesc_16(kBM, BN, 0, 0, F16);  // operands fp16 or bf16, fp32 accumulate

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* staging = smem + kStages * L::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + L::kStagingBytes);
  uint64_t* full_bar = bars;                    // [kStages]
  uint64_t* empty_bar = bars + kStages;         // [kStages]
  uint64_t* tmem_full = bars + 2 * kStages;     // [2]
  uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]
  uint64_t* resid_bar = bars + 2 * kStages + 4;   // [kEpiWarps][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4 + 2 * kEpiWarps);

  const int warp = uniform_warp_index();
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + kBM - 1) / kBM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int n_groups = n_tiles / p.n_inner;
  const int units = p.batch * m_tiles * n_groups;
  const int k_blocks = p.K / kBK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (EPI != EPI_ROWMAX) tma_prefetch_desc(&tmC);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiWarps);
    }
    for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&resid_bar[i], 1);
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
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const int ng = unit % n_groups;
        const int mt = (unit / n_groups) % m_tiles;
        const int b = unit / (n_groups * m_tiles);
        for (int ni = 0; ni < p.n_inner; ++ni) {
          const int nt = ng * p.n_inner + ni;
          for (int kb = 0; kb < k_blocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = stage_base + stage * L::kStageBytes;
            uint8_t* sb = sa + L::kABytes;
            mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
            tma_load_3d(sa, &tmA, &full_bar[stage], kb * kBK, mt * kBM, b);
            tma_load_3d(sb, &tmB, &full_bar[stage], kb * kBK, nt * BN, b);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // The whole warp walks the (warp-uniform) schedule and waits on the mbarriers; lane 0 alone issues tcgen05.mma
    // and tcgen05.commit.  Descriptors are stepped with one 32-bit add: a single divergent thread rebuilding 64-bit
    // descriptors per MMA cost ~150 issue cycles per instruction (measured), i.e. more than the 128 cycles the MMA
    // itself occupies the tensor pipe.
    constexpr uint32_t kDescHi = smem_desc_hi_sw128(1024);
    const uint32_t a_lo0 = smem_desc_lo(smem_u32(stage_base), 16);
    const uint32_t b_lo0 = smem_desc_lo(smem_u32(stage_base) + L::kABytes, 16);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
      for (int ni = 0; ni < p.n_inner; ++ni) {
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
            umma_commit(&empty_bar[stage]);  // smem stage reusable once these MMAs retire
            if (kb == k_blocks - 1) umma_commit(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================================================================== epilogue warps
    const int ew = warp - 2;       // staging / barrier slot
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    uint8_t* my_stage = staging + ew * 2 * kStageBufBytes;
    uint64_t* my_rbar = resid_bar + ew * 2;
    const uint32_t row_sw = (lane & 7);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t g = 0;  // running chunk counter -> staging buffer / barrier parity
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
      const int ng = unit % n_groups;
      const int mt = (unit / n_groups) % m_tiles;
      const int b = unit / (n_groups * m_tiles);
      const int row0 = mt * kBM + quarter * 32;
      float run_max = -INFINITY;
      for (int ni = 0; ni < p.n_inner; ++ni) {
        const int nt = ng * p.n_inner + ni;
        const int col_tile = nt * BN;
        float rs = 1.0f, rm = 0.0f;
        if constexpr (EPI == EPI_BF16) ln_row_scales(p, min(row0 + lane, p.M - 1), rs, rm);
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;

        epilogue_tile<BN, EPI, ACT, F16>(t_row, col_tile, row0, b, p, tmC, my_stage, my_rbar, lane, row_sw, g, run_max, rs, rm);
        // all TMEM reads of this accumulator are complete (tcgen05.wait::ld above)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if constexpr (EPI == EPI_ROWMAX) {
        const int row = row0 + lane;
        if (row < p.M) {
          const size_t idx = static_cast<size_t>(b) * p.M + row;
          p.rowmax[idx] = run_max * __ldg(p.rowscale + idx);
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace lovr
