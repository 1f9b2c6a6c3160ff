// Host side of liblovr_b200.so: tensor-map construction, kernel launchers, the tower forward schedule
// and the extern "C" entry points declared in include/lovr_b200.h.  No torch types, no CPU fallback.
#include "../../include/lovr_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "attention3.cuh"
#include "attention4.cuh"
#include "attention_simt.cuh"
#include "ascore.cuh"
#include "cka.cuh"
#include "cscore.cuh"
#include "gemm.cuh"
#include "gemm2.cuh"
#include "gram.cuh"
#include "kpsim.cuh"
#include "preprocess.cuh"
#include "rowwise.cuh"
#include "scores.cuh"

namespace {

thread_local std::string g_err;
thread_local uint64_t g_launches = 0;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess) return fail(LOVR_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

#define LAUNCH_CHECK(name)                                                                         \
  do {                                                                                             \
    ++g_launches;                                                                                  \
    cudaError_t e_ = cudaGetLastError();                                                           \
    if (e_ != cudaSuccess) return fail(LOVR_E_CUDA, "launch %s: %s", name, cudaGetErrorString(e_)); \
  } while (0)

// ------------------------------------------------------------------------------------ device info
struct DeviceInfo {
  int sms = 0;
  int cc_major = 0;
  bool ok = false;
  std::string why;
};
const DeviceInfo& device_info() {
  static DeviceInfo info = [] {
    DeviceInfo d;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
      d.why = std::string("no CUDA device: ") + cudaGetErrorString(e);
      return d;
    }
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) {
      d.why = cudaGetErrorString(e);
      return d;
    }
    d.sms = p.multiProcessorCount;
    d.cc_major = p.major;
    if (p.major != 10) {
      d.why = "lovr_b200 kernels are sm_100a-only (tcgen05/TMEM); device is sm_" + std::to_string(p.major) +
              std::to_string(p.minor);
      return d;
    }
    d.ok = true;
    return d;
  }();
  return info;
}
int require_device() {
  const DeviceInfo& d = device_info();
  if (!d.ok) return fail(LOVR_E_UNSUPPORTED, "%s", d.why.c_str());
  return 0;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute of a kernel: remember the largest value set for
// (device, kernel) under a mutex, so several GPUs in one process and several host threads both work.
int ensure_smem(const void* func, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> set;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> g(mu);
  size_t& cur = set[std::make_pair(dev, func)];
  if (bytes > cur) {
    CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    cur = bytes;
  }
  return 0;
}

// ------------------------------------------------------------------------------------ tensor maps
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 3-D row-major tensor {dim0 (contiguous), dim1, dim2}; strides in elements; 128-byte swizzle;
// box {box0, box1, box2}; out-of-bounds reads return zero, out-of-bounds writes are dropped.
int make_tmap(CUtensorMap* m, const void* ptr, int elem_bytes, uint64_t dim0, uint64_t dim1, uint64_t dim2,
              uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box0, uint32_t box1, uint32_t box2 = 1) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(LOVR_E_CUDA, "cuTensorMapEncodeTiled entry point not found");
  if (box0 * elem_bytes != 128) return fail(LOVR_E_ARG, "tensor-map inner box must span 128 bytes");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (stride1_elems * elem_bytes) % 16 || (stride2_elems * elem_bytes) % 16)
    return fail(LOVR_E_ARG, "tensor-map base/strides must be 16-byte aligned");
  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t dims[3] = {dim0, dim1, dim2};
  cuuint64_t strides[2] = {stride1_elems * elem_bytes, stride2_elems * elem_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, dt, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(LOVR_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return 0;
}

// ------------------------------------------------------------------------------------ live profiling
// bench.py's roofline leg: CUDA events around every tcgen05 GEMM launch, on the launch stream, so the
// per-launch duration is measured inside the timed region itself (two event records per launch).
struct ProfRec {
  cudaEvent_t a, b;
  double work;   // FLOPs (tensor-bound kinds) or bytes (HBM-bound kinds) of this launch
  int kind;      // LOVR_PROF_*
};
struct Prof {
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;
  std::mutex mu;
};
Prof& prof() {
  static Prof p;
  return p;
}
bool prof_begin(cudaStream_t s, double work, size_t* idx, int kind = LOVR_PROF_GEMM) {
  Prof& p = prof();
  if (!p.on) return false;
  std::lock_guard<std::mutex> g(p.mu);
  ProfRec r;
  if (!p.pool.empty()) {
    r.a = p.pool.back().first;
    r.b = p.pool.back().second;
    p.pool.pop_back();
  } else {
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return false;
  }
  r.work = work;
  r.kind = kind;
  cudaEventRecord(r.a, s);
  p.recs.push_back(r);
  *idx = p.recs.size() - 1;
  return true;
}
void prof_end(cudaStream_t s, size_t idx) {
  Prof& p = prof();
  std::lock_guard<std::mutex> g(p.mu);
  cudaEventRecord(p.recs[idx].b, s);
}

// scope guard: times everything enqueued on `s` between construction and destruction as one record of `kind`
struct ProfScope {
  cudaStream_t s;
  size_t idx = 0;
  bool on;
  ProfScope(cudaStream_t stream, int kind, double work) : s(stream) { on = prof_begin(s, work, &idx, kind); }
  ~ProfScope() { if (on) prof_end(s, idx); }
};

// ------------------------------------------------------------------------------------ GEMM launch
template <int BN, int EPI, int ACT, bool F16>
int launch_gemm_inst(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const lovr::GemmArgs& args,
                     cudaStream_t s) {
  auto kern = lovr::gemm_tcgen05_kernel<BN, EPI, ACT, F16>;
  const int smem = lovr::GemmSmem<BN>::kTotal;
  if (int rc = ensure_smem(reinterpret_cast<const void*>(kern), smem)) return rc;
  const int m_tiles = (args.M + lovr::kBM - 1) / lovr::kBM;
  const int n_tiles = (args.N + BN - 1) / BN;
  const int units = args.batch * m_tiles * (n_tiles / args.n_inner);
  const int grid = units < device_info().sms ? units : device_info().sms;
  if (grid <= 0) return 0;
  size_t pidx = 0;
  const bool profiled = prof_begin(s, 2.0 * args.M * args.N * static_cast<double>(args.K) * args.batch, &pidx);
  kern<<<grid, lovr::kGemmThreads, smem, s>>>(a, b, c, args);
  if (profiled) prof_end(s, pidx);
  LAUNCH_CHECK("gemm_tcgen05_kernel");
  return 0;
}

int launch_gemm(int bn, int epi, int act, bool f16, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c,
                const lovr::GemmArgs& args, cudaStream_t s) {
  using namespace lovr;
  if (args.K % kBK) return fail(LOVR_E_ARG, "gemm: K=%d must be a multiple of 64", args.K);
#define LOVR_GEMM_CASE(BN_, EPI_, ACT_)                                                   \
  if (bn == BN_ && epi == EPI_ && act == ACT_)                                            \
    return f16 ? launch_gemm_inst<BN_, EPI_, ACT_, true>(a, b, c, args, s)                \
               : launch_gemm_inst<BN_, EPI_, ACT_, false>(a, b, c, args, s);
  LOVR_GEMM_CASE(256, EPI_BF16, ACT_NONE)
  LOVR_GEMM_CASE(256, EPI_BF16, ACT_QUICK_GELU)
  LOVR_GEMM_CASE(256, EPI_BF16, ACT_GELU_ERF)
  LOVR_GEMM_CASE(256, EPI_BF16, ACT_GELU_TANH)
  LOVR_GEMM_CASE(256, EPI_RESID_F32, ACT_NONE)
  LOVR_GEMM_CASE(256, EPI_F32, ACT_NONE)
#undef LOVR_GEMM_CASE
  if (!f16 && epi == EPI_ROWMAX && bn == 128) return launch_gemm_inst<128, EPI_ROWMAX, ACT_NONE, false>(a, b, c, args, s);
  if (!f16 && epi == EPI_ROWMAX && bn == 192) return launch_gemm_inst<192, EPI_ROWMAX, ACT_NONE, false>(a, b, c, args, s);
  if (!f16 && epi == EPI_ROWMAX && bn == 256) return launch_gemm_inst<256, EPI_ROWMAX, ACT_NONE, false>(a, b, c, args, s);
  return fail(LOVR_E_UNSUPPORTED, "gemm: no instance for BN=%d epi=%d act=%d f16=%d", bn, epi, act, int(f16));
}

constexpr int kDenseBN = 256;

template <int EPI, int ACT, bool F16, int EW>
int launch_gemm2_inst(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const lovr::GemmArgs& args,
                      cudaStream_t s) {
  auto kern = lovr::gemm2_tcgen05_kernel<EPI, ACT, F16, EW>;
  const int smem = lovr::Gemm2Smem<EPI, EW>::kTotal;
  if (int rc = ensure_smem(reinterpret_cast<const void*>(kern), smem)) return rc;
  const long long units_ll = static_cast<long long>(args.batch > 0 ? args.batch : 1) * ((args.M + 255) / 256) * ((args.N + 255) / 256);
  const int units = units_ll > 0x3fffffff ? 0x3fffffff : static_cast<int>(units_ll);
  int grid = 2 * units < device_info().sms ? 2 * units : (device_info().sms & ~1);
  if (grid <= 0) return 0;
  size_t pidx = 0;
  const bool profiled = prof_begin(s, 2.0 * args.M * args.N * static_cast<double>(args.K) * (args.batch > 0 ? args.batch : 1), &pidx);
  kern<<<grid, lovr::Gemm2Smem<EPI, EW>::kThreads, smem, s>>>(a, b, c, args);  // __cluster_dims__(2,1,1) on the kernel
  if (profiled) prof_end(s, pidx);
  LAUNCH_CHECK("gemm2_tcgen05_kernel");
  return 0;
}

int launch_gemm2(int epi, int act, bool f16, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c,
                 const lovr::GemmArgs& args, cudaStream_t s) {
  using namespace lovr;
  // epilogue warps: 8 (two per TMEM lane quarter) for the residual epilogue, whose per-chunk TMA round trip and proxy
  // fence bound the K=1024 out-projection; LOVR_GEMM_EPI_WARPS=4|8 overrides for experiments
  static const int ew_env = [] { const char* e = getenv("LOVR_GEMM_EPI_WARPS"); return e ? atoi(e) : 0; }();
  const int ew = ew_env ? ew_env : ((epi == EPI_RESID_F32 || act == ACT_GELU_ERF) ? 8 : 4);  // erf-GELU: +9 % with 8
#define LOVR_GEMM2_CASE(EPI_, ACT_)                                                                          \
  if (epi == EPI_ && act == ACT_) {                                                                          \
    if (ew == 8)                                                                                             \
      return f16 ? launch_gemm2_inst<EPI_, ACT_, true, 8>(a, b, c, args, s)                                  \
                 : launch_gemm2_inst<EPI_, ACT_, false, 8>(a, b, c, args, s);                                \
    return f16 ? launch_gemm2_inst<EPI_, ACT_, true, 4>(a, b, c, args, s)                                    \
               : launch_gemm2_inst<EPI_, ACT_, false, 4>(a, b, c, args, s);                                  \
  }
  LOVR_GEMM2_CASE(EPI_BF16, ACT_NONE)
  LOVR_GEMM2_CASE(EPI_BF16, ACT_QUICK_GELU)
  LOVR_GEMM2_CASE(EPI_BF16, ACT_GELU_ERF)
  LOVR_GEMM2_CASE(EPI_BF16, ACT_GELU_TANH)
  LOVR_GEMM2_CASE(EPI_RESID_F32, ACT_NONE)
  LOVR_GEMM2_CASE(EPI_F32, ACT_NONE)
#undef LOVR_GEMM2_CASE
  return fail(LOVR_E_UNSUPPORTED, "gemm2: no instance for epi=%d act=%d", epi, act);
}

// Dense layers default to the CTA-pair kernel (cta_group::2; measured +6 % on qkv, +5 % on fc1/out-proj, +3.5 % on the
// whole tower forward); LOVR_GEMM_2CTA=0 selects the 1-CTA kernel, which also serves M <= 128 and the batched score GEMMs.
int gemm_use_pairs() {
  const char* e = getenv("LOVR_GEMM_2CTA");
  return e ? atoi(e) : 1;
}

// LayerNorm folded into the GEMMs around it (gemm.cuh): consumer side (stats, cw, eps) for the 16-bit epilogue, producer
// side (stats_out, xhat_out) for the residual epilogue.
struct LnFold {
  const float* stats = nullptr;
  const float* cw = nullptr;
  int parts = 0;
  float eps = 0.f;
  float* stats_out = nullptr;
  void* xhat_out = nullptr;
  int64_t ld_xhat = 0;
};

// Dense layer: A [M,K] bf16 (lda), W [N,K] bf16 (ldw), C [M,N] (ldc; bf16 or fp32 by epilogue).
int dense_gemm(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
               int epi, int act, bool f16, const float* bias, const float* colscale, cudaStream_t s,
               const LnFold* ln = nullptr) {
  if (N % kDenseBN) return fail(LOVR_E_ARG, "gemm: N=%d must be a multiple of %d", N, kDenseBN);
  if (M <= 0) return 0;
  CUtensorMap ta, tb, tc;
  int rc;
  if ((rc = make_tmap(&ta, A, 2, K, M, 1, lda, static_cast<uint64_t>(lda) * M, 64, lovr::kBM))) return rc;
  const bool pairs = gemm_use_pairs() != 0 && M > 128;
  if ((rc = make_tmap(&tb, W, 2, K, N, 1, ldw, static_cast<uint64_t>(ldw) * N, 64, pairs ? 128 : kDenseBN))) return rc;
  if (epi == lovr::EPI_BF16) {
    if ((rc = make_tmap(&tc, C, 2, N, M, 1, ldc, static_cast<uint64_t>(ldc) * M, 64, 32))) return rc;
  } else {
    if ((rc = make_tmap(&tc, C, 4, N, M, 1, ldc, static_cast<uint64_t>(ldc) * M, 32, 32))) return rc;
  }
  lovr::GemmArgs g{};
  g.M = M; g.N = N; g.K = K; g.batch = 1; g.n_inner = 1;
  g.bias = bias; g.colscale = colscale; g.rowscale = nullptr; g.rowmax = nullptr; g.n_valid = N; g.scale_ld = 0;
  g.c_ptr = C; g.ldc = ldc; g.c_batch_stride = 0;
  if (epi == lovr::EPI_F32) g.colscale = nullptr;  // (similarity scaling is only used by the C-score path)
  if (ln != nullptr) {
    if (epi == lovr::EPI_BF16 && ln->stats) {
      g.ln_stats = ln->stats; g.ln_cw = ln->cw; g.ln_parts = ln->parts; g.ln_eps = ln->eps;
    }
    if (epi == lovr::EPI_RESID_F32 && ln->stats_out) {
      if (N % lovr::kLnPartCols) return fail(LOVR_E_ARG, "gemm: folded LayerNorm needs N %% 128 == 0");
      g.stats_out = ln->stats_out; g.xhat_out = ln->xhat_out; g.ld_xhat = ln->ld_xhat;
    }
  }
  if (pairs) return launch_gemm2(epi, act, f16, ta, tb, tc, g, s);
  return launch_gemm(kDenseBN, epi, act, f16, ta, tb, tc, g, s);
}

// ------------------------------------------------------------------------------------ row-wise launches
template <typename F4, typename F3>
int dispatch_width(int D, F4 f1024, F3 f768) {
  if (D == 1024) return f1024();
  if (D == 768) return f768();
  return fail(LOVR_E_UNSUPPORTED, "hidden size %d not instantiated (1024, 768)", D);
}

int launch_layernorm(const float* x, const float* g, const float* b, float eps, void* y, int64_t rows, int D, bool f16,
                     cudaStream_t s) {
  if (rows <= 0) return 0;
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((rows + wpb - 1) / wpb);
  auto y16 = static_cast<uint16_t*>(y);
  ProfScope prof_scope(s, LOVR_PROF_LAYERNORM, 6.0 * static_cast<double>(rows) * D);  // 4 B read + 2 B written per element
  int rc = dispatch_width(
      D,
      [&] {
        if (f16) lovr::layernorm_bf16_kernel<8, true><<<grid, wpb * 32, 0, s>>>(x, g, b, eps, y16, rows);
        else lovr::layernorm_bf16_kernel<8, false><<<grid, wpb * 32, 0, s>>>(x, g, b, eps, y16, rows);
        return 0;
      },
      [&] {
        if (f16) lovr::layernorm_bf16_kernel<6, true><<<grid, wpb * 32, 0, s>>>(x, g, b, eps, y16, rows);
        else lovr::layernorm_bf16_kernel<6, false><<<grid, wpb * 32, 0, s>>>(x, g, b, eps, y16, rows);
        return 0;
      });
  if (rc) return rc;
  LAUNCH_CHECK("layernorm_bf16_kernel");
  return 0;
}

int launch_attention(const void* qkv, void* out, int n_img, int T, int H, int impl, bool f16, cudaStream_t s) {
  if (n_img <= 0) return 0;
  const int D = H * 64;
  ProfScope prof_scope(s, LOVR_PROF_ATTENTION, 4.0 * T * static_cast<double>(T) * 64.0 * n_img * H);  // QK^T + PV
  if (impl == LOVR_ATTN_SIMT_CHECK) {
    const size_t smem = lovr::attn_simt_smem_bytes(T);
    if (int rc = ensure_smem(reinterpret_cast<const void*>(lovr::attention_simt_kernel<true>), smem)) return rc;
    if (int rc = ensure_smem(reinterpret_cast<const void*>(lovr::attention_simt_kernel<false>), smem)) return rc;
    if (f16)
      lovr::attention_simt_kernel<true><<<n_img * H, lovr::kSimtAttnWarps * 32, smem, s>>>(
          static_cast<const uint16_t*>(qkv), static_cast<uint16_t*>(out), T, H, 0.125f);
    else
      lovr::attention_simt_kernel<false><<<n_img * H, lovr::kSimtAttnWarps * 32, smem, s>>>(
          static_cast<const uint16_t*>(qkv), static_cast<uint16_t*>(out), T, H, 0.125f);
    LAUNCH_CHECK("attention_simt_kernel");
    return 0;
  }
  const int nb = (T + 127) / 128;
  if (nb > lovr::kAttnMaxBlocks) return fail(LOVR_E_UNSUPPORTED, "attention: T=%d > 640 tokens", T);
  const int rem = T - 128 * (nb - 1);
  lovr::AttnArgs a{};
  a.T = T; a.H = H; a.nb = nb; a.last_w = (rem + 31) / 32 * 32; a.scale = 0.125f;
  a.nu = (T + 63) / 64;
  a.last_w64 = (T - 64 * (a.nu - 1) + 31) / 32 * 32;
  a.out = static_cast<uint16_t*>(out);
  a.n_items = n_img * H;
  {
    const char* dbg = getenv("LOVR_ATTN_DEBUG");  // timing experiments only
    a.debug = dbg ? atoi(dbg) : 0;
    a.trace = nullptr;
  }
  CUtensorMap tq;
  int rc;
  if ((rc = make_tmap(&tq, qkv, 2, 3 * D, T, n_img, 3 * D, static_cast<uint64_t>(3) * D * T, 64, 128))) return rc;
  if (impl == LOVR_ATTN_AUTO) impl = LOVR_ATTN_TCGEN05_V4;  // fastest at T = 577, 257 and 196 (profiles/r01_attention_ncu.md)
  if (impl == LOVR_ATTN_TCGEN05_V4 && nb < 2) impl = LOVR_ATTN_TCGEN05_V3;  // one query tile per item: see attention4.cuh
  if (impl == LOVR_ATTN_TCGEN05_V4) {
    const size_t smem4 = lovr::attn3_smem_bytes(nb);
    const int grid = a.n_items < device_info().sms ? a.n_items : device_info().sms;
#define LOVR_ATTN4_LAUNCH(F16_, RC_)                                                                                  \
  do {                                                                                                               \
    auto kern = lovr::attention4_tcgen05_kernel<F16_, RC_>;                                                             \
    if (int rc_ = ensure_smem(reinterpret_cast<const void*>(kern), smem4)) return rc_;                                 \
    kern<<<grid, lovr::kAttn3Threads, smem4, s>>>(tq, a);                                                            \
  } while (0)
    if (f16) LOVR_ATTN4_LAUNCH(true, lovr::kAttn4ControlRegs); else LOVR_ATTN4_LAUNCH(false, lovr::kAttn4ControlRegs);
#undef LOVR_ATTN4_LAUNCH
    LAUNCH_CHECK("attention4_tcgen05_kernel");
    return 0;
  }
  if (impl == LOVR_ATTN_TCGEN05_V3) {
    const size_t smem3 = lovr::attn3_smem_bytes(nb);
    // LOVR_ATTN3_PIPE (A/B switch for measurements): 0 = plain, 1 = deferred P hand-off (default)
    const char* e_pipe = getenv("LOVR_ATTN3_PIPE");
    const int pipe = e_pipe ? atoi(e_pipe) : lovr::kAttn3DefaultPipe;
#define LOVR_ATTN3_LAUNCH(F16_, PIPE_)                                                                                \
  do {                                                                                                               \
    auto kern = lovr::attention3_tcgen05_kernel<F16_, PIPE_>;                                                        \
    if (int rc_ = ensure_smem(reinterpret_cast<const void*>(kern), smem3)) return rc_;                                 \
    kern<<<n_img * H, lovr::kAttn3Threads, smem3, s>>>(tq, a);                                                       \
  } while (0)
    switch (pipe) {
      case 0: if (f16) LOVR_ATTN3_LAUNCH(true, 0); else LOVR_ATTN3_LAUNCH(false, 0); break;
      case 1: if (f16) LOVR_ATTN3_LAUNCH(true, 1); else LOVR_ATTN3_LAUNCH(false, 1); break;
      default: return fail(LOVR_E_ARG, "LOVR_ATTN3_PIPE=%d: use 0 or 1", pipe);
    }
#undef LOVR_ATTN3_LAUNCH
    LAUNCH_CHECK("attention3_tcgen05_kernel");
    return 0;
  }
  return fail(LOVR_E_UNSUPPORTED, "attention: implementation %d is not in this build (the first two tcgen05 generations were "
                                  "removed in round 2: use LOVR_ATTN_AUTO, _V3, _V4 or LOVR_ATTN_SIMT_CHECK)", impl);
}

int launch_im2col(const void* pixels, int dt, void* cols, int n_img, int R, int P, int Kpad, bool f16, cudaStream_t s) {
  if (n_img <= 0) return 0;
  const int G = R / P;
  // 16-bit pixels already in the operand format: word-copy kernel (no conversion, constant divisors)
  const bool same16 = (dt == LOVR_DT_BF16 && !f16) || (dt == LOVR_DT_F16 && f16);
  if (same16 && (P == 14 || P == 16) && R % 2 == 0 && Kpad % 2 == 0 && (reinterpret_cast<uintptr_t>(pixels) & 3) == 0) {
    const size_t rows = static_cast<size_t>(n_img) * G * G;
    const unsigned g16 = static_cast<unsigned>((rows + 7) / 8);
    if (P == 14)
      lovr::im2col16_kernel<14><<<g16, 256, 0, s>>>(static_cast<const uint32_t*>(pixels), static_cast<uint32_t*>(cols), n_img, R, G, Kpad);
    else
      lovr::im2col16_kernel<16><<<g16, 256, 0, s>>>(static_cast<const uint32_t*>(pixels), static_cast<uint32_t*>(cols), n_img, R, G, Kpad);
    LAUNCH_CHECK("im2col16_kernel");
    return 0;
  }
  const size_t total = static_cast<size_t>(n_img) * G * G * Kpad;
  const unsigned grid = static_cast<unsigned>(std::min<size_t>((total + 255) / 256, 148 * 32));
  lovr::im2col_kernel<<<grid, 256, 0, s>>>(pixels, dt, static_cast<uint16_t*>(cols), n_img, R, P, G, Kpad, f16 ? 1 : 0);
  LAUNCH_CHECK("im2col_kernel");
  return 0;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

// ====================================================================================== tower
struct lovr_tower {
  lovr_tower_config cfg;
  lovr_tower_weights w;
  std::vector<lovr_layer_weights> layers;
  int max_batch;
  int T, G;
  // workspace carve-up
  float* x;      // [B*T, D] fp32 residual stream
  void* ln;      // [B*T, D] bf16: LayerNorm output / attention output
  void* big;     // [B*T, I] bf16: qkv (first 3D columns' worth) or MLP hidden; patch-embed scratch
  void* xh;      // [B*T, D] 16-bit copy of the residual stream (cfg.ln_fold)
  float* stats;  // [B*T, D/128, 4] row statistics of the residual stream (cfg.ln_fold)
  cudaStream_t aux_stream = nullptr;  // second half-batch (created on first use)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

namespace {
struct WsLayout {
  size_t x, ln, big, xh, stats, total;
};
WsLayout ws_layout(const lovr_tower_config& c, int B) {
  const int G = c.image_size / c.patch_size;
  const size_t T = static_cast<size_t>(G) * G + (c.has_cls ? 1 : 0);
  const size_t M = T * B;
  WsLayout l;
  l.x = 0;
  size_t off = align_up(M * c.hidden_size * 4, 1024);
  l.ln = off;
  off += align_up(M * c.hidden_size * 2, 1024);
  l.big = off;
  size_t big = M * static_cast<size_t>(c.intermediate_size) * 2;
  const size_t qkv = M * 3 * static_cast<size_t>(c.hidden_size) * 2;
  const size_t pe = align_up(static_cast<size_t>(B) * G * G * c.patch_k_padded * 2, 1024) +
                    static_cast<size_t>(B) * G * G * c.hidden_size * 4;
  if (qkv > big) big = qkv;
  if (pe > big) big = pe;
  off += align_up(big, 1024) + 2048;  // slack: the second half-batch's slice starts at a 1024-aligned offset
  l.xh = l.stats = 0;
  if (c.ln_fold) {
    l.xh = off;
    off += align_up(M * c.hidden_size * 2, 1024);
    l.stats = off;
    off += align_up(M * (c.hidden_size / lovr::kLnPartCols) * 16, 1024);
  }
  l.total = off;
  return l;
}
}  // namespace

extern "C" {

int lovr_abi_version(void) { return LOVR_ABI_VERSION; }
const char* lovr_last_error(void) { return g_err.c_str(); }
uint64_t lovr_launch_count(void) { return g_launches; }

void lovr_prof_enable(int32_t on) {
  Prof& p = prof();
  std::lock_guard<std::mutex> g(p.mu);
  p.on = on != 0;
}

int lovr_prof_read_kinds(double* ms, double* work, uint64_t* launches) {
  Prof& p = prof();
  std::lock_guard<std::mutex> g(p.mu);
  for (int k = 0; k < LOVR_PROF_KINDS; ++k) {
    if (ms) ms[k] = 0;
    if (work) work[k] = 0;
    if (launches) launches[k] = 0;
  }
  for (ProfRec& r : p.recs) {
    cudaError_t e = cudaEventSynchronize(r.b);
    if (e != cudaSuccess) return fail(LOVR_E_CUDA, "prof_read: %s", cudaGetErrorString(e));
    float t = 0;
    e = cudaEventElapsedTime(&t, r.a, r.b);
    if (e != cudaSuccess) return fail(LOVR_E_CUDA, "prof_read: %s", cudaGetErrorString(e));
    if (ms) ms[r.kind] += t;
    if (work) work[r.kind] += r.work;
    if (launches) launches[r.kind] += 1;
    p.pool.emplace_back(r.a, r.b);
  }
  p.recs.clear();
  return 0;
}

int lovr_prof_read(double* gemm_ms, double* gemm_flops, uint64_t* gemm_launches) {
  double ms[LOVR_PROF_KINDS], work[LOVR_PROF_KINDS];
  uint64_t n[LOVR_PROF_KINDS];
  if (int rc = lovr_prof_read_kinds(ms, work, n)) return rc;
  if (gemm_ms) *gemm_ms = ms[LOVR_PROF_GEMM];
  if (gemm_flops) *gemm_flops = work[LOVR_PROF_GEMM];
  if (gemm_launches) *gemm_launches = n[LOVR_PROF_GEMM];
  return 0;
}

size_t lovr_tower_workspace_bytes(const lovr_tower_config* cfg, int32_t max_batch) {
  if (!cfg || max_batch <= 0) return 0;
  return ws_layout(*cfg, max_batch).total;
}

int lovr_tower_create(lovr_tower_t** out, const lovr_tower_config* cfg, const lovr_tower_weights* w, void* workspace,
                      size_t workspace_bytes, int32_t max_batch) {
  if (!out || !cfg || !w || !workspace) return fail(LOVR_E_ARG, "tower_create: null argument");
  if (int rc = require_device()) return rc;
  if (cfg->hidden_size != cfg->num_heads * 64) return fail(LOVR_E_UNSUPPORTED, "head_dim must be 64");
  if (cfg->hidden_size != 1024 && cfg->hidden_size != 768) return fail(LOVR_E_UNSUPPORTED, "hidden size %d", cfg->hidden_size);
  if (cfg->image_size % cfg->patch_size) return fail(LOVR_E_ARG, "image size not a multiple of the patch size");
  if (cfg->patch_k_padded % 64 || cfg->patch_k_padded < 3 * cfg->patch_size * cfg->patch_size)
    return fail(LOVR_E_ARG, "patch_k_padded must be a multiple of 64 covering 3*P*P");
  if (cfg->intermediate_size % 256 || cfg->hidden_size % 256) return fail(LOVR_E_UNSUPPORTED, "widths must be multiples of 256");
  if (cfg->operand_dtype != LOVR_DT_BF16 && cfg->operand_dtype != LOVR_DT_F16)
    return fail(LOVR_E_ARG, "operand_dtype must be LOVR_DT_BF16 or LOVR_DT_F16");
  const WsLayout l = ws_layout(*cfg, max_batch);
  if (workspace_bytes < l.total) return fail(LOVR_E_NOMEM, "workspace %zu < required %zu", workspace_bytes, l.total);
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(LOVR_E_ARG, "workspace must be 1024-byte aligned");
  auto* t = new lovr_tower();
  t->cfg = *cfg;
  t->w = *w;
  t->layers.assign(w->layers_host, w->layers_host + cfg->num_layers);
  t->w.layers_host = t->layers.data();
  t->max_batch = max_batch;
  t->G = cfg->image_size / cfg->patch_size;
  t->T = t->G * t->G + (cfg->has_cls ? 1 : 0);
  if (t->T > 640) {
    delete t;
    return fail(LOVR_E_UNSUPPORTED, "%d tokens per image > 640", t->T);
  }
  auto base = static_cast<uint8_t*>(workspace);
  t->x = reinterpret_cast<float*>(base + l.x);
  t->ln = base + l.ln;
  t->big = base + l.big;
  t->xh = cfg->ln_fold ? base + l.xh : nullptr;
  t->stats = cfg->ln_fold ? reinterpret_cast<float*>(base + l.stats) : nullptr;
  if (cfg->ln_fold) {
    for (int i = 0; i < cfg->num_layers; ++i)
      if (((cfg->ln_fold & 1) && !t->layers[i].qkv_cw) || ((cfg->ln_fold & 2) && !t->layers[i].fc1_cw)) {
        delete t;
        return fail(LOVR_E_ARG, "tower_create: ln_fold needs qkv_cw / fc1_cw for every folded LayerNorm (lovr_ln_fold_weights)");
      }
  }
  *out = t;
  return 0;
}

void lovr_tower_destroy(lovr_tower_t* t) {
  if (!t) return;
  if (t->aux_stream) cudaStreamDestroy(t->aux_stream);
  if (t->ev_fork) cudaEventDestroy(t->ev_fork);
  if (t->ev_join) cudaEventDestroy(t->ev_join);
  delete t;
}

namespace {
// One contiguous range of images of a forward call, with its own slice of the workspace and its own stream.
struct Range {
  int n0, n;
  cudaStream_t s;
  const void* pixels;
  void* out;
  float* x;
  void* ln;
  void* big;
  void* xh;
  float* stats;
};
size_t dtype_bytes(int dt) { return dt == LOVR_DT_F32 ? 4 : 2; }

int tower_embed(lovr_tower_t* t, const Range& r, int pixel_dtype) {
  using namespace lovr;
  const lovr_tower_config& c = t->cfg;
  const int D = c.hidden_size, T = t->T, G = t->G;
  const bool f16 = c.operand_dtype == LOVR_DT_F16;
  const int M = r.n * T, Mp = r.n * G * G;
  int rc;
  void* cols = r.big;
  float* patch_out = reinterpret_cast<float*>(static_cast<uint8_t*>(r.big) +
                                              align_up(static_cast<size_t>(Mp) * c.patch_k_padded * 2, 1024));
  if ((rc = launch_im2col(r.pixels, pixel_dtype, cols, r.n, c.image_size, c.patch_size, c.patch_k_padded, f16, r.s))) return rc;
  if ((rc = dense_gemm(cols, c.patch_k_padded, t->w.patch_w, c.patch_k_padded, patch_out, D, Mp, D, c.patch_k_padded,
                       EPI_F32, ACT_NONE, f16, t->w.patch_b, nullptr, r.s)))
    return rc;
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((static_cast<size_t>(M) + wpb - 1) / wpb);
  uint16_t* xh = (c.ln_fold & 1) ? static_cast<uint16_t*>(r.xh) : nullptr;
#define LOVR_EMBED_LAUNCH(V4_)                                                                                                  \
  do {                                                                                                                          \
    if (f16)                                                                                                                    \
      embed_finalize_kernel<V4_, true><<<grid, wpb * 32, 0, r.s>>>(patch_out, t->w.cls, t->w.pos, t->w.pre_ln_w, t->w.pre_ln_b, \
                                                                   c.ln_eps, r.x, r.n, T, c.has_cls, xh, r.stats);             \
    else                                                                                                                        \
      embed_finalize_kernel<V4_, false><<<grid, wpb * 32, 0, r.s>>>(patch_out, t->w.cls, t->w.pos, t->w.pre_ln_w, t->w.pre_ln_b, \
                                                                    c.ln_eps, r.x, r.n, T, c.has_cls, xh, r.stats);            \
  } while (0)
  rc = dispatch_width(
      D, [&] { LOVR_EMBED_LAUNCH(8); return 0; }, [&] { LOVR_EMBED_LAUNCH(6); return 0; });
#undef LOVR_EMBED_LAUNCH
  if (rc) return rc;
  LAUNCH_CHECK("embed_finalize_kernel");
  return 0;
}

// step 0..6 of one encoder layer for one range (so two ranges can be enqueued op by op on their streams)
int tower_layer_step(lovr_tower_t* t, const Range& r, int li, int step, int attn_impl) {
  using namespace lovr;
  const lovr_tower_config& c = t->cfg;
  const int D = c.hidden_size, I = c.intermediate_size, T = t->T;
  const bool f16 = c.operand_dtype == LOVR_DT_F16;
  const int M = r.n * T;
  const lovr_layer_weights& lw = t->layers[li];
  if (c.ln_fold) {
    // LayerNorm folded into the GEMMs around it (bit 0: LN1 -> qkv, produced by fc2 / embed_finalize; bit 1: LN2 -> fc1,
    // produced by the out-projection): the producing residual epilogue keeps x^ and the row statistics up to date, the
    // consuming 16-bit epilogue applies them (gemm.cuh); a LayerNorm that is not folded runs as the kernel it was
    const bool f1 = c.ln_fold & 1, f2 = c.ln_fold & 2;
    LnFold consume, produce;
    consume.stats = r.stats; consume.parts = D / kLnPartCols; consume.eps = c.ln_eps;
    produce.stats_out = r.stats; produce.xhat_out = r.xh; produce.ld_xhat = D;
    switch (step) {
      case 0: return f1 ? 0 : launch_layernorm(r.x, lw.ln1_w, lw.ln1_b, c.ln_eps, r.ln, M, D, f16, r.s);
      case 1: consume.cw = lw.qkv_cw;
              return dense_gemm(f1 ? r.xh : r.ln, D, lw.qkv_w, D, r.big, 3 * D, M, 3 * D, D, EPI_BF16, ACT_NONE, f16, lw.qkv_b, nullptr,
                                r.s, f1 ? &consume : nullptr);
      case 2: return launch_attention(r.big, r.ln, r.n, T, c.num_heads, attn_impl, f16, r.s);
      case 3: return dense_gemm(r.ln, D, lw.o_w, D, r.x, D, M, D, D, EPI_RESID_F32, ACT_NONE, f16, lw.o_b, lw.ls1, r.s,
                                f2 ? &produce : nullptr);
      case 4: return f2 ? 0 : launch_layernorm(r.x, lw.ln2_w, lw.ln2_b, c.ln_eps, r.ln, M, D, f16, r.s);
      case 5: consume.cw = lw.fc1_cw;
              return dense_gemm(f2 ? r.xh : r.ln, D, lw.fc1_w, D, r.big, I, M, I, D, EPI_BF16, c.act, f16, lw.fc1_b, nullptr, r.s,
                                f2 ? &consume : nullptr);
      default: return dense_gemm(r.big, I, lw.fc2_w, I, r.x, D, M, D, I, EPI_RESID_F32, ACT_NONE, f16, lw.fc2_b, lw.ls2, r.s,
                                 f1 ? &produce : nullptr);
    }
  }
  switch (step) {
    case 0: return launch_layernorm(r.x, lw.ln1_w, lw.ln1_b, c.ln_eps, r.ln, M, D, f16, r.s);
    case 1: return dense_gemm(r.ln, D, lw.qkv_w, D, r.big, 3 * D, M, 3 * D, D, EPI_BF16, ACT_NONE, f16, lw.qkv_b, nullptr, r.s);
    case 2: return launch_attention(r.big, r.ln, r.n, T, c.num_heads, attn_impl, f16, r.s);
    case 3: return dense_gemm(r.ln, D, lw.o_w, D, r.x, D, M, D, D, EPI_RESID_F32, ACT_NONE, f16, lw.o_b, lw.ls1, r.s);
    case 4: return launch_layernorm(r.x, lw.ln2_w, lw.ln2_b, c.ln_eps, r.ln, M, D, f16, r.s);
    case 5: return dense_gemm(r.ln, D, lw.fc1_w, D, r.big, I, M, I, D, EPI_BF16, c.act, f16, lw.fc1_b, nullptr, r.s);
    default: return dense_gemm(r.big, I, lw.fc2_w, I, r.x, D, M, D, I, EPI_RESID_F32, ACT_NONE, f16, lw.fc2_b, lw.ls2, r.s);
  }
}

int tower_finish(lovr_tower_t* t, const Range& r, int drop, int out_dtype, int64_t out_ld, int64_t out_col_off) {
  using namespace lovr;
  const int D = t->cfg.hidden_size, T = t->T;
  const size_t rows = static_cast<size_t>(r.n) * (T - drop);
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((rows + wpb - 1) / wpb);
  int rc = dispatch_width(
      D,
      [&] {
        select_cast_kernel<8><<<grid, wpb * 32, 0, r.s>>>(r.x, r.out, out_dtype, r.n, T, drop, static_cast<int>(out_ld),
                                                          static_cast<int>(out_col_off));
        return 0;
      },
      [&] {
        select_cast_kernel<6><<<grid, wpb * 32, 0, r.s>>>(r.x, r.out, out_dtype, r.n, T, drop, static_cast<int>(out_ld),
                                                          static_cast<int>(out_col_off));
        return 0;
      });
  if (rc) return rc;
  LAUNCH_CHECK("select_cast_kernel");
  return 0;
}
}  // namespace

int lovr_tower_forward(lovr_tower_t* t, const void* pixels, int32_t pixel_dtype, int32_t n_img, int32_t layers_to_run,
                       int32_t drop_first, void* out, int32_t out_dtype, int64_t out_ld, int64_t out_col_off,
                       int32_t attn_impl, void* stream) {
  if (!t || !pixels || !out) return fail(LOVR_E_ARG, "tower_forward: null argument");
  if (n_img <= 0 || n_img > t->max_batch) return fail(LOVR_E_ARG, "n_img=%d outside [1,%d]", n_img, t->max_batch);
  if (layers_to_run < 0 || layers_to_run > t->cfg.num_layers) return fail(LOVR_E_ARG, "layers_to_run=%d", layers_to_run);
  const lovr_tower_config& c = t->cfg;
  cudaStream_t s0 = static_cast<cudaStream_t>(stream);
  const int D = c.hidden_size, I = c.intermediate_size, T = t->T;
  const int drop = drop_first ? 1 : 0;
  // Optional (LOVR_TOWER_STREAMS=2): two half-batches on two streams, so the HBM-bound kernels of one half (LayerNorm,
  // im2col, select) can co-run with the persistent tensor-core GEMM of the other half.  Measured on B200: no gain
  // (2326 vs 2322 images/s) — the step runs at the 1000 W power cap (sw_power_cap, ~1.55 GHz), so overlapping more work
  // only lowers the clock.  Default is therefore one stream, which also keeps the live GEMM timing of bench.py clean.
  static const int kStreams = [] { const char* e = getenv("LOVR_TOWER_STREAMS"); return e ? atoi(e) : 1; }();
  const int nr = (kStreams >= 2 && n_img >= 16) ? 2 : 1;
  if (nr == 2 && !t->aux_stream) {
    CUDA_TRY(cudaStreamCreateWithFlags(&t->aux_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&t->ev_join, cudaEventDisableTiming));
  }
  Range rg[2];
  const size_t px_img = static_cast<size_t>(3) * c.image_size * c.image_size * dtype_bytes(pixel_dtype);
  const size_t out_img = static_cast<size_t>(T - drop) * out_ld * dtype_bytes(out_dtype);
  for (int i = 0; i < nr; ++i) {
    const int n0 = i == 0 ? 0 : n_img / 2;
    const int n1 = (nr == 1) ? n_img : (i == 0 ? n_img / 2 : n_img);
    rg[i].n0 = n0; rg[i].n = n1 - n0;
    rg[i].s = i == 0 ? s0 : t->aux_stream;
    rg[i].pixels = static_cast<const uint8_t*>(pixels) + n0 * px_img;
    rg[i].out = static_cast<uint8_t*>(out) + n0 * out_img;
    rg[i].x = t->x + static_cast<size_t>(n0) * T * D;
    rg[i].ln = static_cast<uint8_t*>(t->ln) + static_cast<size_t>(n0) * T * D * 2;
    rg[i].big = static_cast<uint8_t*>(t->big) + align_up(static_cast<size_t>(n0) * T * I * 2, 1024);
    rg[i].xh = c.ln_fold ? static_cast<uint8_t*>(t->xh) + static_cast<size_t>(n0) * T * D * 2 : nullptr;
    rg[i].stats = c.ln_fold ? t->stats + static_cast<size_t>(n0) * T * (D / lovr::kLnPartCols) * 4 : nullptr;
  }
  if (nr == 2) {
    CUDA_TRY(cudaEventRecord(t->ev_fork, s0));
    CUDA_TRY(cudaStreamWaitEvent(t->aux_stream, t->ev_fork, 0));
  }
  int rc;
  for (int i = 0; i < nr; ++i)
    if ((rc = tower_embed(t, rg[i], pixel_dtype))) return rc;
  for (int li = 0; li < layers_to_run; ++li)
    for (int step = 0; step < 7; ++step)
      for (int i = 0; i < nr; ++i)
        if ((rc = tower_layer_step(t, rg[i], li, step, attn_impl))) return rc;
  for (int i = 0; i < nr; ++i)
    if ((rc = tower_finish(t, rg[i], drop, out_dtype, out_ld, out_col_off))) return rc;
  if (nr == 2) {
    CUDA_TRY(cudaEventRecord(t->ev_join, t->aux_stream));
    CUDA_TRY(cudaStreamWaitEvent(s0, t->ev_join, 0));
  }
  return 0;
}

// ====================================================================================== building blocks
int lovr_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int32_t M, int32_t N,
                   int32_t K, int32_t epi, int32_t act, int32_t operand_dtype, const float* bias, const float* colscale,
                   void* stream) {
  if (int rc = require_device()) return rc;
  if (!A || !W || !C) return fail(LOVR_E_ARG, "gemm: null pointer");
  if (epi < 0 || epi > LOVR_EPI_F32) return fail(LOVR_E_ARG, "gemm: epilogue %d", epi);
  if (epi != LOVR_EPI_BF16 && act != LOVR_ACT_NONE) return fail(LOVR_E_ARG, "gemm: activation only with the bf16 epilogue");
  return dense_gemm(A, lda, W, ldw, C, ldc, M, N, K, epi, act, operand_dtype == LOVR_DT_F16, bias, colscale,
                    static_cast<cudaStream_t>(stream));
}

int lovr_gemm_lnfold(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int32_t M, int32_t N,
                     int32_t K, int32_t epi, int32_t act, int32_t operand_dtype, const float* bias, const float* colscale,
                     const float* ln_stats, const float* ln_cw, int32_t ln_cols, float ln_eps, float* stats_out, void* xhat_out,
                     int64_t ld_xhat, void* stream) {
  if (int rc = require_device()) return rc;
  if (!A || !W || !C) return fail(LOVR_E_ARG, "gemm: null pointer");
  if (epi != LOVR_EPI_BF16 && epi != LOVR_EPI_RESID_F32) return fail(LOVR_E_ARG, "gemm_lnfold: epilogue %d", epi);
  if (epi != LOVR_EPI_BF16 && act != LOVR_ACT_NONE) return fail(LOVR_E_ARG, "gemm: activation only with the bf16 epilogue");
  LnFold ln;
  if (epi == LOVR_EPI_BF16) {
    if (!ln_stats || !ln_cw || ln_cols <= 0 || ln_cols % lovr::kLnPartCols)
      return fail(LOVR_E_ARG, "gemm_lnfold: the 16-bit epilogue needs ln_stats, ln_cw and ln_cols %% 128 == 0");
    ln.stats = ln_stats; ln.cw = ln_cw; ln.parts = ln_cols / lovr::kLnPartCols; ln.eps = ln_eps;
  } else {
    if (!stats_out || !xhat_out) return fail(LOVR_E_ARG, "gemm_lnfold: the residual epilogue needs stats_out and xhat_out");
    ln.stats_out = stats_out; ln.xhat_out = xhat_out; ln.ld_xhat = ld_xhat;
  }
  return dense_gemm(A, lda, W, ldw, C, ldc, M, N, K, epi, act, operand_dtype == LOVR_DT_F16, bias, colscale,
                    static_cast<cudaStream_t>(stream), &ln);
}

int lovr_ln_fold_weights(const float* W, const float* gamma, const float* beta, const float* bias, int32_t N, int32_t K,
                         int32_t operand_dtype, void* W16, float* cw, float* bw, void* stream) {
  if (int rc = require_device()) return rc;
  if (!W || !gamma || !beta || !W16 || !cw || !bw || N <= 0 || K <= 0) return fail(LOVR_E_ARG, "ln_fold_weights: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>((N + 7) / 8);
  if (operand_dtype == LOVR_DT_F16)
    lovr::ln_fold_weights_kernel<true><<<grid, 256, 0, s>>>(W, gamma, beta, bias, N, K, static_cast<uint16_t*>(W16), cw, bw);
  else
    lovr::ln_fold_weights_kernel<false><<<grid, 256, 0, s>>>(W, gamma, beta, bias, N, K, static_cast<uint16_t*>(W16), cw, bw);
  LAUNCH_CHECK("ln_fold_weights_kernel");
  return 0;
}

int lovr_layernorm_bf16(const float* x, const float* gamma, const float* beta, float eps, void* y, int64_t rows,
                        int32_t D, int32_t operand_dtype, void* stream) {
  if (int rc = require_device()) return rc;
  return launch_layernorm(x, gamma, beta, eps, y, rows, D, operand_dtype == LOVR_DT_F16, static_cast<cudaStream_t>(stream));
}

int lovr_attention(const void* qkv, void* out, int32_t n_img, int32_t T, int32_t H, int32_t impl, int32_t operand_dtype,
                   void* stream) {
  if (int rc = require_device()) return rc;
  return launch_attention(qkv, out, n_img, T, H, impl, operand_dtype == LOVR_DT_F16, static_cast<cudaStream_t>(stream));
}

int lovr_im2col(const void* pixels, int32_t pixel_dtype, void* cols, int32_t n_img, int32_t image_size, int32_t patch,
                int32_t k_padded, int32_t operand_dtype, void* stream) {
  if (int rc = require_device()) return rc;
  return launch_im2col(pixels, pixel_dtype, cols, n_img, image_size, patch, k_padded, operand_dtype == LOVR_DT_F16,
                       static_cast<cudaStream_t>(stream));
}

// ====================================================================================== A score
namespace {
struct AScoreLayout {
  size_t inv_o, inv_r, rowmax, split_o, split_r, total;
};
AScoreLayout a_layout(int dtype, int64_t n, int x, int y, int D) {
  AScoreLayout l{};
  size_t off = 0;
  l.inv_o = off; off += align_up(static_cast<size_t>(n) * x * 4, 1024);
  l.inv_r = off; off += align_up(static_cast<size_t>(n) * y * 4, 1024);
  l.rowmax = off; off += align_up(static_cast<size_t>(n) * x * 4, 1024);
  if (dtype == LOVR_DT_F32) {
    l.split_o = off; off += align_up(static_cast<size_t>(n) * x * 3 * D * 2, 1024);
    l.split_r = off; off += align_up(static_cast<size_t>(n) * y * 3 * D * 2, 1024);
  }
  l.total = off;
  return l;
}
int pick_bn(int y) {
  int best = 128;
  long best_waste = -1;
  for (int bn : {128, 192, 256}) {
    const long waste = static_cast<long>((y + bn - 1) / bn) * bn - y;
    if (best_waste < 0 || waste < best_waste || (waste == best_waste && bn > best)) {
      best = bn;
      best_waste = waste;
    }
  }
  return best;
}
}  // namespace

size_t lovr_a_score_workspace_bytes(int32_t dtype, int64_t n_pairs, int32_t x, int32_t y, int32_t D) {
  return a_layout(dtype, n_pairs, x, y, D).total;
}

int lovr_a_score_maxcos(const void* other, const void* ref, int32_t dtype, int64_t n_pairs, int32_t x, int32_t y,
                        int32_t D, float* pair_score, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace lovr;
  if (int rc = require_device()) return rc;
  if (!other || !ref || !pair_score || !workspace) return fail(LOVR_E_ARG, "a_score: null pointer");
  if (dtype != LOVR_DT_BF16 && dtype != LOVR_DT_F32) return fail(LOVR_E_UNSUPPORTED, "a_score: dtype %d (bf16 or fp32)", dtype);
  if (D % 64 || x <= 0 || y <= 0 || n_pairs <= 0) return fail(LOVR_E_ARG, "a_score: need D %% 64 == 0 and positive sizes");
  if (n_pairs > 0x7fffffff / ((x + 127) / 128)) return fail(LOVR_E_ARG, "a_score: too many pairs for one call");
  const AScoreLayout l = a_layout(dtype, n_pairs, x, y, D);
  if (workspace_bytes < l.total) return fail(LOVR_E_NOMEM, "a_score workspace %zu < %zu", workspace_bytes, l.total);
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(LOVR_E_ARG, "workspace must be 1024-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto base = static_cast<uint8_t*>(workspace);
  float* inv_o = reinterpret_cast<float*>(base + l.inv_o);
  float* inv_r = reinterpret_cast<float*>(base + l.inv_r);
  float* rowmax = reinterpret_cast<float*>(base + l.rowmax);
  const size_t rows_o = static_cast<size_t>(n_pairs) * x, rows_r = static_cast<size_t>(n_pairs) * y;
  const int wpb = 8;
  const void* a_ptr = other;
  const void* b_ptr = ref;
  int K = D;
  const int bn = pick_bn(y);
  // bf16 features: the `other` norms come out of the shared-memory tiles of the GEMM itself (ascore.cuh), so `other` is read
  // from HBM once; only `ref` gets a norm pre-pass.  Needs more k-blocks than pipeline stages (its barrier-phase argument).
  static const int fuse_env = [] { const char* e = getenv("LOVR_A_SCORE_FUSED"); return e ? atoi(e) : 1; }();
  if (dtype == LOVR_DT_BF16 && fuse_env && D / kBK > 5) {
    rownorm_inv_bf16_kernel<<<static_cast<unsigned>((rows_r + wpb - 1) / wpb), wpb * 32, 0, s>>>(
        static_cast<const __nv_bfloat16*>(ref), inv_r, rows_r, D);
    LAUNCH_CHECK("rownorm_inv_bf16_kernel");
    size_t smem = 0;
    const void* kern = nullptr;
    // BN = 192 (y = 576): five 40 KB stages fit and measure 1.5 % faster than four (9.22 vs 9.36 ms at C3); LOVR_A_STAGES=4 for A/B
    static const int stages5 = [] { const char* e = getenv("LOVR_A_STAGES"); return e ? atoi(e) == 5 : 1; }();
    if (bn == 192 && stages5) {
      smem = AScoreSmem<192, 5>::kTotal;
      kern = reinterpret_cast<const void*>(ascore_tcgen05_kernel<192, 5>);
    } else
    if (bn == 128) { smem = AScoreSmem<128>::kTotal; kern = reinterpret_cast<const void*>(ascore_tcgen05_kernel<128>); }
    else if (bn == 192) { smem = AScoreSmem<192>::kTotal; kern = reinterpret_cast<const void*>(ascore_tcgen05_kernel<192>); }
    else { smem = AScoreSmem<256>::kTotal; kern = reinterpret_cast<const void*>(ascore_tcgen05_kernel<256>); }
    CUtensorMap ta, tb;
    int rc;
    if ((rc = make_tmap(&ta, other, 2, D, x, n_pairs, D, static_cast<uint64_t>(D) * x, 64, kBM))) return rc;
    if ((rc = make_tmap(&tb, ref, 2, D, y, n_pairs, D, static_cast<uint64_t>(D) * y, 64, bn))) return rc;
    if ((rc = ensure_smem(kern, smem))) return rc;
    AScoreArgs a{};
    a.x = x; a.y = y; a.K = D; a.n_pairs = static_cast<int>(n_pairs); a.inv_ref = inv_r; a.rowmax = rowmax;
    { const char* e = getenv("LOVR_A_PREFETCH"); a.prefetch = e ? atoi(e) : 0; }
    const long long units = static_cast<long long>(n_pairs) * ((x + kBM - 1) / kBM);
    const int grid = units < device_info().sms ? static_cast<int>(units) : device_info().sms;
    {
      ProfScope prof_scope(s, LOVR_PROF_OTHER, 2.0 * x * static_cast<double>(y) * D * n_pairs);
      if (bn == 192 && stages5) ascore_tcgen05_kernel<192, 5><<<grid, kAsThreads, smem, s>>>(ta, tb, a);
      else if (bn == 128) ascore_tcgen05_kernel<128><<<grid, kAsThreads, smem, s>>>(ta, tb, a);
      else if (bn == 192) ascore_tcgen05_kernel<192><<<grid, kAsThreads, smem, s>>>(ta, tb, a);
      else ascore_tcgen05_kernel<256><<<grid, kAsThreads, smem, s>>>(ta, tb, a);
    }
    LAUNCH_CHECK("ascore_tcgen05_kernel");
    pair_mean_kernel<<<static_cast<unsigned>(n_pairs), 256, 0, s>>>(rowmax, pair_score, x);
    LAUNCH_CHECK("pair_mean_kernel");
    return 0;
  }
  if (dtype == LOVR_DT_BF16) {
    rownorm_inv_bf16_kernel<<<static_cast<unsigned>((rows_o + wpb - 1) / wpb), wpb * 32, 0, s>>>(
        static_cast<const __nv_bfloat16*>(other), inv_o, rows_o, D);
    LAUNCH_CHECK("rownorm_inv_bf16_kernel");
    rownorm_inv_bf16_kernel<<<static_cast<unsigned>((rows_r + wpb - 1) / wpb), wpb * 32, 0, s>>>(
        static_cast<const __nv_bfloat16*>(ref), inv_r, rows_r, D);
    LAUNCH_CHECK("rownorm_inv_bf16_kernel");
  } else {
    auto so = reinterpret_cast<__nv_bfloat16*>(base + l.split_o);
    auto sr = reinterpret_cast<__nv_bfloat16*>(base + l.split_r);
    split3_kernel<<<static_cast<unsigned>((rows_o + wpb - 1) / wpb), wpb * 32, 0, s>>>(
        static_cast<const float*>(other), so, inv_o, rows_o, D, 0);
    LAUNCH_CHECK("split3_kernel");
    split3_kernel<<<static_cast<unsigned>((rows_r + wpb - 1) / wpb), wpb * 32, 0, s>>>(
        static_cast<const float*>(ref), sr, inv_r, rows_r, D, 1);
    LAUNCH_CHECK("split3_kernel");
    a_ptr = so;
    b_ptr = sr;
    K = 3 * D;
  }
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_tmap(&ta, a_ptr, 2, K, x, n_pairs, K, static_cast<uint64_t>(K) * x, 64, kBM))) return rc;
  if ((rc = make_tmap(&tb, b_ptr, 2, K, y, n_pairs, K, static_cast<uint64_t>(K) * y, 64, bn))) return rc;
  GemmArgs g{};
  g.M = x; g.N = y; g.K = K; g.batch = static_cast<int>(n_pairs);
  g.n_inner = (y + bn - 1) / bn;
  g.bias = nullptr; g.colscale = inv_r; g.rowscale = inv_o; g.rowmax = rowmax; g.n_valid = y;
  if ((rc = launch_gemm(bn, EPI_ROWMAX, ACT_NONE, false, ta, tb, ta /*unused*/, g, s))) return rc;
  pair_mean_kernel<<<static_cast<unsigned>(n_pairs), 256, 0, s>>>(rowmax, pair_score, x);
  LAUNCH_CHECK("pair_mean_kernel");
  return 0;
}

// ====================================================================================== C score
namespace {
struct CScoreLayout {
  size_t inv1, inv2, sim, split1, split2, total;
  int scale_ld, sim_ld;
};
CScoreLayout c_layout(int dtype, int64_t n, int P, int D) {
  CScoreLayout l{};
  const int PP = P * P;
  l.scale_ld = (PP + 255) / 256 * 256 + 256;  // float4 epilogue loads may run past P*P inside the last N tile
  l.sim_ld = (PP + 3) / 4 * 4;
  size_t off = 0;
  l.inv1 = off; off += align_up(static_cast<size_t>(n) * l.scale_ld * 4, 1024);
  l.inv2 = off; off += align_up(static_cast<size_t>(n) * l.scale_ld * 4, 1024);
  l.sim = off; off += align_up(static_cast<size_t>(n) * PP * l.sim_ld * 4, 1024);
  if (dtype == LOVR_DT_F32) {
    l.split1 = off; off += align_up(static_cast<size_t>(n) * PP * 3 * D * 2, 1024);
    l.split2 = off; off += align_up(static_cast<size_t>(n) * PP * 3 * D * 2, 1024);
  }
  l.total = off;
  return l;
}
int check_c_args(const void* d1, const void* d2, int dtype, int64_t n, int P, int D) {
  if (!d1 || !d2) return fail(LOVR_E_ARG, "c_score: null pointer");
  if (dtype != LOVR_DT_BF16 && dtype != LOVR_DT_F32) return fail(LOVR_E_UNSUPPORTED, "c_score: dtype %d (bf16 or fp32)", dtype);
  if (P < 2 || P > lovr::kMaxGrid) return fail(LOVR_E_UNSUPPORTED, "c_score: grid side P=%d outside [2,%d]", P, lovr::kMaxGrid);
  if (n <= 0 || D <= 0) return fail(LOVR_E_ARG, "c_score: sizes must be positive");
  return 0;
}
int launch_flow_rows(const float* sim, size_t ld, size_t rows, int P, int window, int soft, float stride, float* flow,
                     int* argmax, cudaStream_t s) {
  const int wpb = 8;
  lovr::flow_rows_kernel<<<static_cast<unsigned>((rows + wpb - 1) / wpb), wpb * 32, 0, s>>>(
      sim, ld, rows, P, window, 0.02f, soft ? 1 : 0, stride, flow, argmax);
  LAUNCH_CHECK("flow_rows_kernel");
  return 0;
}
}  // namespace

size_t lovr_c_score_workspace_bytes(int32_t dtype, int64_t n_pairs, int32_t P, int32_t D, int32_t K) {
  if (K > 0) return align_up(static_cast<size_t>(n_pairs) * K * P * P * 4, 1024);  // keypoint rows only
  return c_layout(dtype, n_pairs, P, D).total;
}

namespace {
// n(desc1) n(desc2)^T for every pair into the workspace (layout l): inverse norms, optional fp32 -> 3 x bf16 split,
// batched tcgen05 GEMM with row/column scales.  Shared by the dense flow and the mutual-NN distance.
int c_dense_similarity(const void* desc1, const void* desc2, int32_t dtype, int64_t n_pairs, int32_t P, int32_t D,
                       uint8_t* base, const CScoreLayout& l, cudaStream_t s) {
  using namespace lovr;
  float* inv1 = reinterpret_cast<float*>(base + l.inv1);
  float* inv2 = reinterpret_cast<float*>(base + l.inv2);
  float* sim = reinterpret_cast<float*>(base + l.sim);
  const int PP = P * P;
  const size_t rows = static_cast<size_t>(n_pairs) * PP;
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((rows + wpb - 1) / wpb);
  // padded tails of the scale vectors are read by the last N tile's float4 loads: keep them finite
  CUDA_TRY(cudaMemsetAsync(inv1, 0, static_cast<size_t>(n_pairs) * l.scale_ld * 4, s));
  CUDA_TRY(cudaMemsetAsync(inv2, 0, static_cast<size_t>(n_pairs) * l.scale_ld * 4, s));
  const void* a_ptr = desc1;
  const void* b_ptr = desc2;
  int K = D;
  if (dtype == LOVR_DT_BF16) {
    rownorm_inv_bf16_kernel<<<grid, wpb * 32, 0, s>>>(static_cast<const __nv_bfloat16*>(desc1), inv1, rows, D, 1, PP, l.scale_ld);
    LAUNCH_CHECK("rownorm_inv_bf16_kernel");
    rownorm_inv_bf16_kernel<<<grid, wpb * 32, 0, s>>>(static_cast<const __nv_bfloat16*>(desc2), inv2, rows, D, 1, PP, l.scale_ld);
    LAUNCH_CHECK("rownorm_inv_bf16_kernel");
  } else {
    auto s1 = reinterpret_cast<__nv_bfloat16*>(base + l.split1);
    auto s2 = reinterpret_cast<__nv_bfloat16*>(base + l.split2);
    split3_kernel<<<grid, wpb * 32, 0, s>>>(static_cast<const float*>(desc1), s1, inv1, rows, D, 0, 1, PP, l.scale_ld);
    LAUNCH_CHECK("split3_kernel");
    split3_kernel<<<grid, wpb * 32, 0, s>>>(static_cast<const float*>(desc2), s2, inv2, rows, D, 1, 1, PP, l.scale_ld);
    LAUNCH_CHECK("split3_kernel");
    a_ptr = s1;
    b_ptr = s2;
    K = 3 * D;
  }
  CUtensorMap ta, tb, tc;
  int rc;
  if ((rc = make_tmap(&ta, a_ptr, 2, K, PP, n_pairs, K, static_cast<uint64_t>(K) * PP, 64, kBM))) return rc;
  // CTA-pair kernel (each SM loads half of the desc2 tile; the 1-CTA 128 x 256 tile is bound by shared-memory bandwidth at
  // ~67 % of the tensor pipe) when a pair has more than one 128-row tile; LOVR_CSIM_2CTA=0 keeps the 1-CTA kernel (A/B)
  const char* pairs_s = getenv("LOVR_CSIM_2CTA");
  const bool pairs = (pairs_s ? atoi(pairs_s) != 0 : true) && gemm_use_pairs() != 0 && PP > 128;
  if ((rc = make_tmap(&tb, b_ptr, 2, K, PP, n_pairs, K, static_cast<uint64_t>(K) * PP, 64, pairs ? 128 : 256))) return rc;
  if ((rc = make_tmap(&tc, sim, 4, PP, PP, n_pairs, l.sim_ld, static_cast<uint64_t>(l.sim_ld) * PP, 32, 32))) return rc;
  GemmArgs g{};
  g.M = PP; g.N = PP; g.K = K; g.batch = static_cast<int>(n_pairs); g.n_inner = 1;
  g.bias = nullptr; g.colscale = inv2; g.rowscale = inv1; g.rowmax = nullptr; g.n_valid = PP; g.scale_ld = l.scale_ld;
  if (pairs) return launch_gemm2(EPI_F32, ACT_NONE, false, ta, tb, tc, g, s);
  return launch_gemm(256, EPI_F32, ACT_NONE, false, ta, tb, tc, g, s);
}
struct MnnLayout {
  size_t row_d2, row_arg, col_d2, col_arg, total;
};
MnnLayout mnn_layout(const CScoreLayout& l, int64_t n, int P) {
  MnnLayout m{};
  const size_t PP = static_cast<size_t>(P) * P;
  size_t off = l.total;
  m.row_d2 = off; off += align_up(static_cast<size_t>(n) * PP * 4, 1024);
  m.row_arg = off; off += align_up(static_cast<size_t>(n) * PP * 4, 1024);
  m.col_d2 = off; off += align_up(static_cast<size_t>(n) * lovr::kMnnSlabs * PP * 4, 1024);
  m.col_arg = off; off += align_up(static_cast<size_t>(n) * lovr::kMnnSlabs * PP * 4, 1024);
  m.total = off;
  return m;
}
}  // namespace

int lovr_c_score_flow(const void* desc1, const void* desc2, int32_t dtype, int64_t n_pairs, int32_t P, int32_t D,
                      int32_t window, int32_t soft, float* flow, int32_t* argmax, void* workspace,
                      size_t workspace_bytes, void* stream) {
  if (int rc = require_device()) return rc;
  if (int rc = check_c_args(desc1, desc2, dtype, n_pairs, P, D)) return rc;
  if (!flow || !workspace) return fail(LOVR_E_ARG, "c_score_flow: null pointer");
  if (D % 64) return fail(LOVR_E_ARG, "c_score_flow: D=%d must be a multiple of 64 (pad channels with zeros)", D);
  const CScoreLayout l = c_layout(dtype, n_pairs, P, D);
  if (workspace_bytes < l.total) return fail(LOVR_E_NOMEM, "c_score workspace %zu < %zu", workspace_bytes, l.total);
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(LOVR_E_ARG, "workspace must be 1024-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto base = static_cast<uint8_t*>(workspace);
  if (int rc = c_dense_similarity(desc1, desc2, dtype, n_pairs, P, D, base, l, s)) return rc;
  return launch_flow_rows(reinterpret_cast<float*>(base + l.sim), l.sim_ld, static_cast<size_t>(n_pairs) * P * P, P,
                          window, soft, 0.f, flow, argmax, s);
}

size_t lovr_c_score_mutual_nn_workspace_bytes(int32_t dtype, int64_t n_pairs, int32_t P, int32_t D) {
  return mnn_layout(c_layout(dtype, n_pairs, P, D), n_pairs, P).total;
}

int lovr_c_score_mutual_nn(const void* desc1, const void* desc2, int32_t dtype, int64_t n_pairs, int32_t P, int32_t D,
                           float* out, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace lovr;
  if (int rc = require_device()) return rc;
  if (int rc = check_c_args(desc1, desc2, dtype, n_pairs, P, D)) return rc;
  if (!out || !workspace) return fail(LOVR_E_ARG, "c_score_mutual_nn: null pointer");
  if (D % 64) return fail(LOVR_E_ARG, "c_score_mutual_nn: D=%d must be a multiple of 64 (pad channels with zeros)", D);
  if (n_pairs > 65535) return fail(LOVR_E_ARG, "c_score_mutual_nn: at most 65535 pairs per call");
  const CScoreLayout l = c_layout(dtype, n_pairs, P, D);
  const MnnLayout m = mnn_layout(l, n_pairs, P);
  if (workspace_bytes < m.total) return fail(LOVR_E_NOMEM, "c_score_mutual_nn workspace %zu < %zu", workspace_bytes, m.total);
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(LOVR_E_ARG, "workspace must be 1024-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto base = static_cast<uint8_t*>(workspace);
  if (int rc = c_dense_similarity(desc1, desc2, dtype, n_pairs, P, D, base, l, s)) return rc;
  const float* sim = reinterpret_cast<const float*>(base + l.sim);
  const float* inv1 = reinterpret_cast<const float*>(base + l.inv1);
  const float* inv2 = reinterpret_cast<const float*>(base + l.inv2);
  float* row_d2 = reinterpret_cast<float*>(base + m.row_d2);
  int* row_arg = reinterpret_cast<int*>(base + m.row_arg);
  float* col_d2 = reinterpret_cast<float*>(base + m.col_d2);
  int* col_arg = reinterpret_cast<int*>(base + m.col_arg);
  const int PP = P * P;
  const size_t rows = static_cast<size_t>(n_pairs) * PP;
  mnn_rows_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(sim, l.sim_ld, inv1, inv2, l.scale_ld, PP, rows,
                                                                        row_d2, row_arg);
  LAUNCH_CHECK("mnn_rows_kernel");
  mnn_cols_kernel<<<dim3((PP + 127) / 128, kMnnSlabs, static_cast<unsigned>(n_pairs)), 128, 0, s>>>(
      sim, l.sim_ld, inv1, inv2, l.scale_ld, PP, col_d2, col_arg);
  LAUNCH_CHECK("mnn_cols_kernel");
  mnn_reduce_kernel<<<static_cast<unsigned>(n_pairs), 256, 0, s>>>(row_d2, row_arg, col_d2, col_arg, PP, out);
  LAUNCH_CHECK("mnn_reduce_kernel");
  return 0;
}

int lovr_c_score_keypoints(const void* desc1, const void* desc2, int32_t dtype, int64_t n_pairs, int32_t P, int32_t D,
                           const int32_t* patch_idx, int32_t K, int32_t window, int32_t soft, float anno_size,
                           float* kps_out, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace lovr;
  if (int rc = require_device()) return rc;
  if (int rc = check_c_args(desc1, desc2, dtype, n_pairs, P, D)) return rc;
  if (!patch_idx || !kps_out || !workspace) return fail(LOVR_E_ARG, "c_score_keypoints: null pointer");
  if (K <= 0 || K > kMaxKps) return fail(LOVR_E_UNSUPPORTED, "c_score_keypoints: K=%d outside [1,%d]", K, kMaxKps);
  if (D % 32) return fail(LOVR_E_ARG, "c_score_keypoints: D=%d must be a multiple of 32", D);
  const int PP = P * P;
  const size_t need = static_cast<size_t>(n_pairs) * K * PP * 4;
  if (workspace_bytes < need) return fail(LOVR_E_NOMEM, "c_score_keypoints workspace %zu < %zu", workspace_bytes, need);
  if (n_pairs > 65535) return fail(LOVR_E_ARG, "c_score_keypoints: at most 65535 pairs per call");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* rows = static_cast<float*>(workspace);
  // bf16 descriptors: the keypoint rows as a skinny tcgen05 GEMM (kpsim.cuh) — the SM only keeps TMA loads of the target
  // descriptors in flight.  Needs more k-blocks than pipeline stages and the source rows resident in shared memory.
  static const int kp_tc_env = [] { const char* e = getenv("LOVR_KP_TENSOR"); return e ? atoi(e) : 1; }();
  if (kp_tc_env && dtype == LOVR_DT_BF16 && D % 64 == 0 && D / 64 > kKpStages && kp_smem_bytes(D) <= 227 * 1024) {
    CUtensorMap ta;
    if (int rc = make_tmap(&ta, desc2, 2, D, PP, n_pairs, D, static_cast<uint64_t>(D) * PP, 64, kBM)) return rc;
    const size_t smem = kp_smem_bytes(D);
    if (int rc = ensure_smem(reinterpret_cast<const void*>(kpsim_tcgen05_kernel), smem)) return rc;
    KpArgs a{};
    a.PP = PP; a.D = D; a.K = K; a.n_pairs = static_cast<int>(n_pairs);
    a.d1 = static_cast<const uint16_t*>(desc1); a.idx = patch_idx; a.out = rows;
    const int m_tiles = (PP + kBM - 1) / kBM;
    const long long units = static_cast<long long>(n_pairs) * ((m_tiles + kKpTilesPerUnit - 1) / kKpTilesPerUnit);
    const int grid = units < device_info().sms ? static_cast<int>(units) : device_info().sms;
    {
      ProfScope prof_scope(s, LOVR_PROF_OTHER, static_cast<double>(n_pairs) * (PP + K) * D * 2.0);  // bytes: HBM-bound
      kpsim_tcgen05_kernel<<<grid, kKpThreads, smem, s>>>(ta, a);
    }
    LAUNCH_CHECK("kpsim_tcgen05_kernel");
    return launch_flow_rows(rows, PP, static_cast<size_t>(n_pairs) * K, P, window, soft, anno_size / P, kps_out, nullptr, s);
  }
  // The K normalised source rows live in shared memory (K * D floats).  SPair-71k pads to 30 keypoints and the two-model
  // concat has D = 2048: 240 KB, above the 227 KB a block may opt in to — so K is split into equal chunks that fit.
  constexpr size_t kKpSmemBudget = 200 * 1024;
  const int kc_max = static_cast<int>(kKpSmemBudget / (static_cast<size_t>(D) * 4));
  if (kc_max < 1) return fail(LOVR_E_UNSUPPORTED, "c_score_keypoints: D=%d: one source row exceeds the shared-memory budget", D);
  const int n_chunks = (K + kc_max - 1) / kc_max;
  const int kc = (K + n_chunks - 1) / n_chunks;
  const int rpb = 64;
  dim3 grid((PP + rpb - 1) / rpb, static_cast<unsigned>(n_pairs));
  const float* f1 = static_cast<const float*>(desc1);
  const float* f2 = static_cast<const float*>(desc2);
  const __nv_bfloat16* h1 = static_cast<const __nv_bfloat16*>(desc1);
  const __nv_bfloat16* h2 = static_cast<const __nv_bfloat16*>(desc2);
  const bool f32 = dtype == LOVR_DT_F32;
  for (int k0 = 0; k0 < K; k0 += kc) {
    const int kn = K - k0 < kc ? K - k0 : kc;
    const size_t smem = static_cast<size_t>(kn) * D * 4;
#define LOVR_KP_LAUNCH(KERN_, ...)                                                        \
  do {                                                                                   \
    if (int rc_ = ensure_smem(reinterpret_cast<const void*>(&KERN_), smem)) return rc_;  \
    KERN_<<<grid, 256, smem, s>>>(__VA_ARGS__);                                          \
  } while (0)
    if (D == 1024) {
      if (f32) LOVR_KP_LAUNCH((kp_sim_rows4_kernel<float, 8>), f1, f2, patch_idx, K, k0, kn, PP, rpb, rows);
      else LOVR_KP_LAUNCH((kp_sim_rows4_kernel<__nv_bfloat16, 8>), h1, h2, patch_idx, K, k0, kn, PP, rpb, rows);
    } else if (D == 768) {
      if (f32) LOVR_KP_LAUNCH((kp_sim_rows4_kernel<float, 6>), f1, f2, patch_idx, K, k0, kn, PP, rpb, rows);
      else LOVR_KP_LAUNCH((kp_sim_rows4_kernel<__nv_bfloat16, 6>), h1, h2, patch_idx, K, k0, kn, PP, rpb, rows);
    } else if (D == 2048) {
      if (f32) LOVR_KP_LAUNCH((kp_sim_rows4_kernel<float, 16>), f1, f2, patch_idx, K, k0, kn, PP, rpb, rows);
      else LOVR_KP_LAUNCH((kp_sim_rows4_kernel<__nv_bfloat16, 16>), h1, h2, patch_idx, K, k0, kn, PP, rpb, rows);
    } else {
      if (f32) LOVR_KP_LAUNCH(kp_sim_rows_kernel<float>, f1, f2, patch_idx, K, k0, kn, PP, D, rpb, rows);
      else LOVR_KP_LAUNCH(kp_sim_rows_kernel<__nv_bfloat16>, h1, h2, patch_idx, K, k0, kn, PP, D, rpb, rows);
    }
#undef LOVR_KP_LAUNCH
    LAUNCH_CHECK("kp_sim_rows_kernel");
  }
  return launch_flow_rows(rows, PP, static_cast<size_t>(n_pairs) * K, P, window, soft, anno_size / P, kps_out, nullptr, s);
}

int lovr_c_score_concat_two(const void* a, const void* b, int32_t dtype, int64_t rows, int32_t Da, int32_t Db, void* out,
                            void* stream) {
  if (int rc = require_device()) return rc;
  if (!a || !b || !out) return fail(LOVR_E_ARG, "c_score_concat_two: null pointer");
  if (dtype != LOVR_DT_BF16 && dtype != LOVR_DT_F32) return fail(LOVR_E_UNSUPPORTED, "c_score_concat_two: dtype %d (bf16 or fp32)", dtype);
  if (rows <= 0 || Da <= 0 || Db <= 0) return fail(LOVR_E_ARG, "c_score_concat_two: sizes must be positive");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  if (dtype == LOVR_DT_F32)
    lovr::concat_normalize_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(a), static_cast<const float*>(b),
                                                              static_cast<size_t>(rows), Da, Db, static_cast<float*>(out));
  else
    lovr::concat_normalize_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(
        static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), static_cast<size_t>(rows), Da, Db,
        static_cast<__nv_bfloat16*>(out));
  LAUNCH_CHECK("concat_normalize_kernel");
  return 0;
}

int lovr_pck(const float* pred, const float* gt, const uint8_t* visible, const float* thresholds, int64_t n_pairs,
             int32_t K, float* out, void* stream) {
  if (int rc = require_device()) return rc;
  if (!pred || !gt || !visible || !thresholds || !out) return fail(LOVR_E_ARG, "pck: null pointer");
  if (n_pairs <= 0) return 0;
  lovr::pck_kernel<<<static_cast<unsigned>((n_pairs + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, gt, visible, thresholds, static_cast<int>(n_pairs), K, out);
  LAUNCH_CHECK("pck_kernel");
  return 0;
}

// ====================================================================================== linear CKA
namespace {
// One wave of persistent CTAs: unit = (row range, upper-triangle tile of the [X|Y] Gram), see gram.cuh.  The layout depends
// only on the sizes (148 SMs assumed for the range count; any grid size runs it correctly).
struct CkaLayout {
  size_t pace, pace_bytes, total;
  int n_tiles, n_ranges, total_chunks;
};
CkaLayout cka_layout(int64_t R, int Dx, int Dy) {
  CkaLayout l{};
  l.n_tiles = lovr::gram_tile_count(Dx + Dy);
  l.total_chunks = static_cast<int>((R + lovr::kGramChunk - 1) / lovr::kGramChunk);
  l.n_ranges = 148 / l.n_tiles;
  if (l.n_ranges < 1) l.n_ranges = 1;
  if (l.n_ranges > l.total_chunks) l.n_ranges = l.total_chunks > 0 ? l.total_chunks : 1;
  l.pace = align_up(static_cast<size_t>(l.n_tiles) * l.n_ranges * lovr::kGramTileElems * 4, 1024);
  l.pace_bytes = static_cast<size_t>(l.n_ranges) * ((l.total_chunks + l.n_ranges - 1) / l.n_ranges + 1) * 4;
  l.total = l.pace + align_up(l.pace_bytes, 1024);
  return l;
}
}  // namespace

size_t lovr_cka_workspace_bytes(int64_t rows, int32_t Dx, int32_t Dy) { return cka_layout(rows, Dx, Dy).total; }
size_t lovr_cka_gram_elems(int32_t Dx, int32_t Dy) {
  return static_cast<size_t>(Dx) * (Dx + Dy) + static_cast<size_t>(Dy) * Dy + Dx + Dy + 1;  // last element: row count
}

int lovr_cka_accumulate(const void* X, const void* Y, int64_t rows, int32_t Dx, int32_t Dy, double* gram, void* workspace,
                        size_t workspace_bytes, void* stream) {
  using namespace lovr;
  if (int rc = require_device()) return rc;
  if (!X || !Y || !gram || !workspace) return fail(LOVR_E_ARG, "cka: null pointer");
  if (rows <= 0) return 0;
  if (Dx % 256 || Dy % 256) return fail(LOVR_E_ARG, "cka: feature widths must be multiples of 256 (got %d, %d)", Dx, Dy);
  if (rows > 0x7fffffff - kGramChunk) return fail(LOVR_E_ARG, "cka: at most 2^31 rows per call");
  const CkaLayout l = cka_layout(rows, Dx, Dy);
  if (workspace_bytes < l.total) return fail(LOVR_E_NOMEM, "cka workspace %zu < %zu", workspace_bytes, l.total);
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(LOVR_E_ARG, "workspace must be 1024-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int Dz = Dx + Dy;
  const size_t na = static_cast<size_t>(Dx) * Dz, nb = static_cast<size_t>(Dy) * Dy;
  // features viewed as [groups of 64][row][64]: ONE TMA box {64 features, 64 rows, g groups} lands as g consecutive 8 KB
  // MN-major slabs — the A operand (2 groups) and the B operand (4 groups) of a k-block are one load each
  CUtensorMap txa, txb, tya, tyb;
  int rc;
  if ((rc = make_tmap(&txa, X, 2, 64, rows, Dx / 64, Dx, 64, 64, 64, 2))) return rc;
  if ((rc = make_tmap(&txb, X, 2, 64, rows, Dx / 64, Dx, 64, 64, 64, 4))) return rc;
  if ((rc = make_tmap(&tya, Y, 2, 64, rows, Dy / 64, Dy, 64, 64, 64, 2))) return rc;
  if ((rc = make_tmap(&tyb, Y, 2, 64, rows, Dy / 64, Dy, 64, 64, 64, 4))) return rc;
  GramArgs g{};
  g.R = rows; g.Dx = Dx; g.Dy = Dy; g.n_tiles = l.n_tiles; g.n_ranges = l.n_ranges; g.total_chunks = l.total_chunks;
  g.partial = static_cast<float*>(workspace);
  g.colsum = gram + na + nb;
  { const char* e = getenv("LOVR_GRAM_DEBUG"); g.debug = e ? atoi(e) : 0; }
  { const char* e = getenv("LOVR_GRAM_PREFETCH"); g.prefetch = e ? atoi(e) : kGramPrefetch; }
  if ((rc = ensure_smem(reinterpret_cast<const void*>(gram_tcgen05_kernel), kGramSmem))) return rc;
  const int units = l.n_tiles * l.n_ranges;
  const int grid = units < device_info().sms ? units : device_info().sms;
  g.pace = nullptr;
  static const int pace_env = [] { const char* e = getenv("LOVR_GRAM_PACE"); return e ? atoi(e) : 1; }();
  if (pace_env && units <= grid) {  // one unit per CTA, every CTA resident (1 CTA per SM by shared memory): the window cannot deadlock
    g.pace = reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + l.pace);
    CUDA_TRY(cudaMemsetAsync(g.pace, 0, l.pace_bytes, s));
  }
  {
    // executed FLOPs: only the upper-triangle tiles (the algorithmic count of SURVEY.md §8d is 2 R (Dx Dz + Dy^2))
    ProfScope prof_scope(s, LOVR_PROF_OTHER, 2.0 * static_cast<double>(rows) * l.n_tiles * kGramTileElems);
    gram_tcgen05_kernel<<<grid, kGramThreads, kGramSmem, s>>>(txa, txb, tya, tyb, g);
  }
  LAUNCH_CHECK("gram_tcgen05_kernel");
  // the row count travels in the same buffer (its last element), so the multi-GPU path needs ONE all-reduce
  const size_t n = static_cast<size_t>(l.n_tiles) * kGramTileElems;
  gram_reduce_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(g.partial, l.n_tiles, l.n_ranges, Dx, Dy, gram, gram + na,
                                                                        gram + na + nb + Dz, static_cast<double>(rows));
  LAUNCH_CHECK("gram_reduce_kernel");
  return 0;
}

int lovr_cka_finalize(const double* gram, double n_rows, int32_t Dx, int32_t Dy, double* out3, void* stream) {
  if (int rc = require_device()) return rc;
  if (!gram || !out3) return fail(LOVR_E_ARG, "cka_finalize: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUDA_TRY(cudaMemsetAsync(out3, 0, 3 * sizeof(double), s));
  const size_t na = static_cast<size_t>(Dx) * (Dx + Dy), nb = static_cast<size_t>(Dy) * Dy;
  // n_rows <= 0: use the count accumulated in the buffer's last element (summed over ranks by the same all-reduce)
  lovr::cka_finalize_kernel<<<296, 256, 0, s>>>(gram, gram + na, gram + na + nb, n_rows, gram + na + nb + Dx + Dy, Dx, Dy, out3);
  LAUNCH_CHECK("cka_finalize_kernel");
  return 0;
}

// ====================================================================================== image preprocessing
namespace {
struct PrePlan {
  std::vector<lovr::PreImg> imgs;
  size_t lut_off = 0, coef_off = 0, tmp_off = 0, total = 0;
  int max_rows = 0;
};
// Resample.c precompute_coeffs bounds of output xx (same double operations as resample_coeffs_kernel; this file is
// compiled with -ffp-contract=off so the host never fuses them)
void resample_bounds(int in_size, int out_size, int xx, int* xmin, int* xmax, int* ksize) {
  const double scale = static_cast<double>(in_size) / static_cast<double>(out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  const double center = (static_cast<double>(xx) + 0.5) * scale;
  int lo = static_cast<int>(center - support + 0.5);
  if (lo < 0) lo = 0;
  int hi = static_cast<int>(center + support + 0.5);
  if (hi > in_size) hi = in_size;
  *xmin = lo;
  *xmax = hi;
  *ksize = static_cast<int>(ceil(support)) * 2 + 1;
}
int pre_plan(const lovr_preprocess_config* cfg, int32_t n_img, const int32_t* hw, const int64_t* offsets, PrePlan* plan) {
  if (!cfg || !hw || n_img <= 0 || n_img > 65535)
    return fail(LOVR_E_ARG, "preprocess: null config / sizes, or n_img outside [1, 65535]");
  const int R = cfg->crop_size, S = cfg->resize_size;
  if (R <= 0 || S <= 0 || S < R) return fail(LOVR_E_ARG, "preprocess: need 0 < crop_size <= resize_size (got %d, %d)", R, S);
  if (cfg->resize_mode != LOVR_RESIZE_SHORTEST_EDGE && cfg->resize_mode != LOVR_RESIZE_EXACT)
    return fail(LOVR_E_ARG, "preprocess: unknown resize_mode %d", cfg->resize_mode);
  if (cfg->pixel_map != LOVR_PIX_HF_RESCALE_NORMALIZE && cfg->pixel_map != LOVR_PIX_CSCORE)
    return fail(LOVR_E_ARG, "preprocess: unknown pixel_map %d", cfg->pixel_map);
  plan->imgs.resize(n_img);
  size_t coef_ints = 0, tmp_bytes = 0;
  for (int i = 0; i < n_img; ++i) {
    lovr::PreImg& im = plan->imgs[i];
    im.H = hw[2 * i];
    im.W = hw[2 * i + 1];
    if (im.H <= 0 || im.W <= 0 || im.H > 32768 || im.W > 32768)
      return fail(LOVR_E_ARG, "preprocess: image %d has unsupported size %dx%d", i, im.H, im.W);
    im.src_off = offsets ? offsets[i] : 0;
    im.vH = im.H;
    im.vW = im.W;
    im.pad_top = im.pad_left = 0;
    if (cfg->pad_square && im.H != im.W) {  // mm_utils.py:63-75: paste at ((side - w)//2, 0) or (0, (side - h)//2)
      const int side = im.H > im.W ? im.H : im.W;
      im.pad_top = (side - im.H) / 2;
      im.pad_left = (side - im.W) / 2;
      im.vH = im.vW = side;
    }
    if (cfg->resize_mode == LOVR_RESIZE_EXACT) {
      im.newH = im.newW = S;
    } else {  // transformers get_resize_output_image_size(default_to_square=False)
      const int shrt = im.vW <= im.vH ? im.vW : im.vH, lng = im.vW <= im.vH ? im.vH : im.vW;
      const int new_long = static_cast<int>(static_cast<double>(static_cast<long long>(S) * lng) / shrt);
      im.newH = im.vW <= im.vH ? new_long : S;
      im.newW = im.vW <= im.vH ? S : new_long;
    }
    if (im.newH < R || im.newW < R)
      return fail(LOVR_E_ARG, "preprocess: image %d resizes to %dx%d, smaller than the crop %d (zero-padding crop is not on this path)",
                  i, im.newH, im.newW, R);
    im.crop_top = (im.newH - R) / 2;
    im.crop_left = (im.newW - R) / 2;
    int lo, hi, k, lo2, hi2;
    resample_bounds(im.vW, im.newW, im.crop_left, &lo, &hi, &k);
    im.kx = k;
    resample_bounds(im.vH, im.newH, im.crop_top, &lo, &hi, &k);
    resample_bounds(im.vH, im.newH, im.crop_top + R - 1, &lo2, &hi2, &k);
    im.ky = k;
    im.row0 = lo;
    im.rows = hi2 - lo;
    im.coef_off = static_cast<long long>(coef_ints);
    im.tmp_off = static_cast<long long>(tmp_bytes);
    coef_ints += static_cast<size_t>(R) * (2 + im.kx) + static_cast<size_t>(R) * (2 + im.ky);
    tmp_bytes += align_up(static_cast<size_t>(3) * im.rows * R, 16);
    if (im.rows > plan->max_rows) plan->max_rows = im.rows;
  }
  plan->lut_off = align_up(static_cast<size_t>(n_img) * sizeof(lovr::PreImg), 256);
  plan->coef_off = plan->lut_off + align_up(768 * sizeof(float), 256);
  plan->tmp_off = plan->coef_off + align_up(coef_ints * 4, 256);
  plan->total = plan->tmp_off + align_up(tmp_bytes, 256);
  return 0;
}
}  // namespace

size_t lovr_preprocess_workspace_bytes(const lovr_preprocess_config* cfg, int32_t n_img, const int32_t* hw_host) {
  PrePlan plan;
  if (pre_plan(cfg, n_img, hw_host, nullptr, &plan)) return 0;
  return plan.total;
}

int lovr_preprocess(const uint8_t* images, const int64_t* offsets_host, const int32_t* hw_host, int32_t n_img,
                    const lovr_preprocess_config* cfg, void* out, int32_t out_dtype, void* workspace,
                    size_t workspace_bytes, void* stream) {
  if (int rc = require_device()) return rc;
  if (n_img == 0) return 0;
  if (!images || !offsets_host || !out || !workspace) return fail(LOVR_E_ARG, "preprocess: null pointer");
  if (out_dtype != LOVR_DT_F32 && out_dtype != LOVR_DT_BF16 && out_dtype != LOVR_DT_F16)
    return fail(LOVR_E_ARG, "preprocess: unsupported out_dtype %d", out_dtype);
  PrePlan plan;
  if (int rc = pre_plan(cfg, n_img, hw_host, offsets_host, &plan)) return rc;
  if (workspace_bytes < plan.total)
    return fail(LOVR_E_NOMEM, "preprocess: workspace %zu < required %zu bytes", workspace_bytes, plan.total);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const lovr::PreImg* d_imgs = reinterpret_cast<const lovr::PreImg*>(ws);
  int* d_coefs = reinterpret_cast<int*>(ws + plan.coef_off);
  uint8_t* d_tmp = ws + plan.tmp_off;
  // pageable source: the runtime stages the descriptors before returning, so `plan` may die at scope exit
  CUDA_TRY(cudaMemcpyAsync(ws, plan.imgs.data(), plan.imgs.size() * sizeof(lovr::PreImg), cudaMemcpyHostToDevice, s));
  const int R = cfg->crop_size;
  lovr::resample_coeffs_kernel<<<dim3((R + 127) / 128, 2, n_img), 128, 0, s>>>(d_imgs, d_coefs, n_img, R);
  LAUNCH_CHECK("resample_coeffs_kernel");
  float* d_lut = reinterpret_cast<float*>(ws + plan.lut_off);
  lovr::PixelMap pm;
  pm.mode = cfg->pixel_map;
  for (int c = 0; c < 3; ++c) {
    pm.mean[c] = cfg->mean[c];
    pm.stdv[c] = cfg->std[c];
  }
  lovr::pixel_lut_kernel<<<3, 256, 0, s>>>(d_lut, pm);
  LAUNCH_CHECK("pixel_lut_kernel");
  const uint32_t bg = cfg->pad_rgb[0] | (cfg->pad_rgb[1] << 8) | (cfg->pad_rgb[2] << 16);
  const dim3 hgrid((plan.max_rows + lovr::kPreRowsPerBlock - 1) / lovr::kPreRowsPerBlock, n_img);
  if (cfg->pad_square)
    lovr::resample_h_kernel<true><<<hgrid, 128, 0, s>>>(images, d_imgs, d_coefs, d_tmp, R, bg);
  else
    lovr::resample_h_kernel<false><<<hgrid, 128, 0, s>>>(images, d_imgs, d_coefs, d_tmp, R, bg);
  LAUNCH_CHECK("resample_h_kernel");
  if (R % 4 == 0) {
    const dim3 vgrid((R * (R / 4) + 255) / 256, n_img);
    if (out_dtype == LOVR_DT_F32)
      lovr::resample_v4_kernel<lovr::DT_F32><<<vgrid, 256, 0, s>>>(d_tmp, d_imgs, d_coefs, d_lut, out, R);
    else if (out_dtype == LOVR_DT_BF16)
      lovr::resample_v4_kernel<lovr::DT_BF16><<<vgrid, 256, 0, s>>>(d_tmp, d_imgs, d_coefs, d_lut, out, R);
    else
      lovr::resample_v4_kernel<lovr::DT_F16><<<vgrid, 256, 0, s>>>(d_tmp, d_imgs, d_coefs, d_lut, out, R);
  } else {
    lovr::resample_v_kernel<<<dim3(R, n_img), 256, 0, s>>>(d_tmp, d_imgs, d_coefs, d_lut, out, out_dtype, R);
  }
  LAUNCH_CHECK("resample_v_kernel");
  return 0;
}

}  // extern "C"
