/* lovr_b200 — C ABI of the B200-native vision-representation scoring path.
 *
 * The reference (bronyayang/Law_of_Vision_Representation_in_MLLMs) has NO native/FFI boundary on this path
 * (SURVEY.md §0 F5): its "operator API" is the Python nn.Module contract of the tower classes plus two
 * score scripts.  This header is the boundary a native replacement needs; every entry point cites the
 * reference interface it stands behind.  The Python mirror of the reference classes
 * (law_of_vision_representation_in_mllms_b200/towers.py etc.) binds these symbols with ctypes.
 *
 * Conventions: plain pointers and sizes only; all pointers are DEVICE pointers unless named *_host;
 * the caller owns every buffer (the library never allocates or frees caller memory; its only internal
 * allocations are small host-side plan objects); all work is enqueued on `stream` (a CUstream /
 * cudaStream_t handle passed as void*) with no host synchronisation; return 0 on success, a negative
 * LOVR_E_* code on failure with a message available from lovr_last_error() (thread-local).
 * There is no CPU fallback: without an sm_100 device every compute entry point fails.
 */
#ifndef LOVR_B200_H_
#define LOVR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LOVR_ABI_VERSION 2

enum { LOVR_OK = 0, LOVR_E_ARG = -1, LOVR_E_CUDA = -2, LOVR_E_UNSUPPORTED = -3, LOVR_E_NOMEM = -4 };
enum { LOVR_DT_F32 = 0, LOVR_DT_BF16 = 1, LOVR_DT_F16 = 2 };
/* hidden_act of the HF config: quick_gelu (openai CLIP), gelu (OpenCLIP, DINOv2), gelu_pytorch_tanh (SigLIP) */
enum { LOVR_ACT_NONE = 0, LOVR_ACT_QUICK_GELU = 1, LOVR_ACT_GELU_ERF = 2, LOVR_ACT_GELU_TANH = 3 };
enum { LOVR_EPI_BF16 = 0, LOVR_EPI_RESID_F32 = 1, LOVR_EPI_F32 = 2 };
/* attention implementations.  Values 0, 2, 3, 4 (the first two tcgen05 generations) were removed from the build in round 2:
 * they return LOVR_E_UNSUPPORTED.  1 = CUDA-core cross-check used only by tests. */
enum { LOVR_ATTN_SIMT_CHECK = 1,
       LOVR_ATTN_AUTO = 5 /* the product default: currently 7 (6 when T <= 128) */,
       LOVR_ATTN_TCGEN05_V3 = 6 /* one CTA per (image, head): three softmax pipelines, P through TMEM, K/V streamed per block */,
       LOVR_ATTN_TCGEN05_V4 = 7 /* persistent: one CTA per SM, tiles of consecutive (image, head) items roll through the three pipelines */ };

/* Threading: every entry point may be called from any host thread on any stream; lovr_last_error() and lovr_launch_count()
 * are per host thread; per-kernel attributes are cached per device under a mutex (several GPUs in one process work).
 * A lovr_tower_t owns ONE workspace: calls on the same handle must be stream-ordered (one handle per concurrent stream).
 * The live profiler (lovr_prof_*) is a process-wide measurement aid, mutex-guarded. */
int lovr_abi_version(void);
const char* lovr_last_error(void);
/* number of kernels this library has launched on the calling thread since load (bench.py's gpu_launches) */
uint64_t lovr_launch_count(void);
/* Live roofline measurement: while enabled, every tcgen05 GEMM launch is bracketed by CUDA events on its launch
 * stream; lovr_prof_read() synchronises on them, returns the summed duration (ms), the summed algorithmic FLOPs
 * (2*M*N*K per launch) and the launch count since the last read, and resets the accumulator. */
void lovr_prof_enable(int32_t on);
int lovr_prof_read(double* gemm_ms, double* gemm_flops, uint64_t* gemm_launches);
/* The same accumulator per kernel class: ms / work / launches are arrays of LOVR_PROF_KINDS entries.  `work` is the
 * algorithmic FLOP count for the tensor-bound classes (GEMM: 2*M*N*K; attention: 4*T*T*64 per (image, head)) and the
 * algorithmic byte count for the HBM-bound ones (LayerNorm: 6 B per element).  Reading resets the accumulator. */
enum { LOVR_PROF_GEMM = 0, LOVR_PROF_ATTENTION = 1, LOVR_PROF_LAYERNORM = 2, LOVR_PROF_OTHER = 3, LOVR_PROF_KINDS = 4 };
int lovr_prof_read_kinds(double* ms, double* work, uint64_t* launches);

/* ---------------------------------------------------------------------------------------------------
 * Vision tower (replaces the arithmetic behind CLIPVisionTower.forward / DinoV2VisionTower.forward /
 * SigLipVisionTower.forward: llava/model/multimodal_encoder/clip_encoder.py:39-51,
 * dinov2_encoder.py:32-54, siglip_encoder.py:30-52, i.e. HF CLIPVisionModel / Dinov2Model /
 * SiglipVisionModel with output_hidden_states=True followed by feature_select).
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t hidden_size;       /* 1024 | 768 (multiple of 128) */
  int32_t intermediate_size; /* 4096 | 3072 */
  int32_t num_layers;        /* layers present in the weights */
  int32_t num_heads;         /* head_dim must be 64 */
  int32_t image_size;        /* input resolution this instance runs at */
  int32_t patch_size;
  int32_t act;               /* LOVR_ACT_* */
  int32_t has_cls;
  int32_t pre_ln;            /* CLIP pre_layrnorm */
  int32_t layerscale;        /* DINOv2 LayerScale */
  float ln_eps;
  int32_t patch_k_padded;    /* 3*patch*patch rounded up to a multiple of 64 (im2col K) */
  int32_t operand_dtype;     /* LOVR_DT_BF16 or LOVR_DT_F16: 16-bit format of the tensor-core operands (activations AND
                                the weight matrices below); accumulation, residual stream, LayerNorm, softmax are fp32 */
  int32_t ln_fold;           /* bit mask: 1 = LayerNorm-1 folded into the qkv GEMM, 2 = LayerNorm-2 folded into fc1 (no LayerNorm
                                kernel for a folded one): the folded GEMM's weight holds W diag(gamma), its bias W beta + b and
                                qkv_cw / fc1_cw the row sums of the folded weight — all three produced by
                                lovr_ln_fold_weights(); the folded LayerNorm's ln*_w / ln*_b are unused */
} lovr_tower_config;

typedef struct {
  const float* ln1_w; const float* ln1_b;
  const void* qkv_w;  /* 16-bit (operand_dtype) [3D, D] rows = q,k,v output features */
  const float* qkv_b; /* [3D] */
  const void* o_w;    /* 16-bit [D, D] */
  const float* o_b;
  const float* ls1;   /* [D] or NULL */
  const float* ln2_w; const float* ln2_b;
  const void* fc1_w;  /* 16-bit [I, D] */
  const float* fc1_b;
  const void* fc2_w;  /* 16-bit [D, I] */
  const float* fc2_b;
  const float* ls2;   /* [D] or NULL */
  const float* qkv_cw; /* [3D] (ln_fold) */
  const float* fc1_cw; /* [I]  (ln_fold) */
} lovr_layer_weights;

typedef struct {
  const void* patch_w;   /* 16-bit [D, patch_k_padded], k = c*P*P + py*P + px, zero padded */
  const float* patch_b;  /* [D] or NULL (CLIP has no conv bias) */
  const float* cls;      /* [D] or NULL */
  const float* pos;      /* [T, D] (already interpolated for image_size: DINOv2) */
  const float* pre_ln_w; const float* pre_ln_b; /* or NULL */
  const lovr_layer_weights* layers_host; /* HOST array [num_layers] of device pointers */
} lovr_tower_weights;

typedef struct lovr_tower lovr_tower_t;

/* bytes of device workspace needed to run up to max_batch images per forward call */
size_t lovr_tower_workspace_bytes(const lovr_tower_config* cfg, int32_t max_batch);
int lovr_tower_create(lovr_tower_t** out, const lovr_tower_config* cfg, const lovr_tower_weights* w,
                      void* workspace, size_t workspace_bytes, int32_t max_batch);
void lovr_tower_destroy(lovr_tower_t* t);
/* pixels [n,3,R,R] in pixel_dtype (NCHW, already normalised like the HF processors do) ->
 * out[n, p, out_col_off : out_col_off + D], row stride out_ld elements, dtype out_dtype.
 * layers_to_run = select_layer mod (L+1) (hidden_states[select_layer]); drop_cls=1 for 'patch'.
 * attn_impl: LOVR_ATTN_AUTO (product) or LOVR_ATTN_SIMT_CHECK (cross-check kernel for tests). */
int lovr_tower_forward(lovr_tower_t* t, const void* pixels, int32_t pixel_dtype, int32_t n_img,
                       int32_t layers_to_run, int32_t drop_first_token, void* out, int32_t out_dtype,
                       int64_t out_ld, int64_t out_col_off, int32_t attn_impl, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Building blocks (exported so each kernel is parity-tested in isolation through the same ABI).
 * ------------------------------------------------------------------------------------------------- */
/* operand_dtype (LOVR_DT_BF16 | LOVR_DT_F16) = the 16-bit format of A, W and of 16-bit outputs; fp32 accumulate.
 * C[M,N] = epilogue(A[M,K] · W[N,K]^T): nn.Linear semantics. lda/ldw/ldc in elements.
 * LOVR_EPI_BF16: C 16-bit = act(acc+bias). LOVR_EPI_F32: C fp32 = acc+bias.
 * LOVR_EPI_RESID_F32: C fp32 += colscale*(acc+bias) in place. */
int lovr_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int32_t M,
                   int32_t N, int32_t K, int32_t epi, int32_t act, int32_t operand_dtype, const float* bias,
                   const float* colscale, void* stream);
/* LayerNorm folded into the GEMMs on both sides of it:  LN(x) W^T + b = rstd (x^ W'^T - mu cw) + bw.
 * lovr_ln_fold_weights (weight ingest, once): W [N,K] fp32, gamma/beta [K], bias [N] or NULL ->
 *   W16 = 16-bit(W diag(gamma)), cw[j] = sum_k W16[j,k] (of the rounded values), bw = W beta + bias.
 * lovr_gemm_lnfold, LOVR_EPI_BF16: A = x^ (16-bit copy of the residual stream [M, ln_cols]), W = W16, bias = bw; ln_stats
 *   [M][ln_cols/128][4] = per 128 columns of x (shift, sum(x-shift), sum((x-shift)^2), 0); C 16-bit = act(LN(x) W^T + b).
 * lovr_gemm_lnfold, LOVR_EPI_RESID_F32: C fp32 += colscale*(acc+bias) as lovr_gemm_bf16, and additionally writes the 16-bit
 *   copy of the updated C to xhat_out (row stride ld_xhat) and its statistics slots to stats_out [M][N/128][4]. */
int lovr_ln_fold_weights(const float* W, const float* gamma, const float* beta, const float* bias, int32_t N, int32_t K,
                         int32_t operand_dtype, void* W16, float* cw, float* bw, void* stream);
int lovr_gemm_lnfold(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int32_t M, int32_t N,
                     int32_t K, int32_t epi, int32_t act, int32_t operand_dtype, const float* bias, const float* colscale,
                     const float* ln_stats, const float* ln_cw, int32_t ln_cols, float ln_eps, float* stats_out,
                     void* xhat_out, int64_t ld_xhat, void* stream);
int lovr_layernorm_bf16(const float* x, const float* gamma, const float* beta, float eps, void* y_16bit,
                        int64_t rows, int32_t D, int32_t operand_dtype, void* stream);
/* qkv 16-bit [n_img*T, 3*H*64] -> out 16-bit [n_img*T, H*64]; softmax(q k^T / 8) v per (image, head) */
int lovr_attention(const void* qkv, void* out, int32_t n_img, int32_t T, int32_t H, int32_t impl,
                   int32_t operand_dtype, void* stream);
int lovr_im2col(const void* pixels, int32_t pixel_dtype, void* cols_16bit, int32_t n_img, int32_t image_size,
                int32_t patch, int32_t k_padded, int32_t operand_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * A score (replaces the per-pair arithmetic of A_score/compute.py:54-72):
 *   pair_score[b] = mean_i max_j cos(n(other[b,i,:]), n(ref[b,j,:])),  n(x)=x/(||x||+1e-10), cos eps 1e-8.
 * other [n_pairs, x, D], ref [n_pairs, y, D] contiguous, dtype bf16 (exact products, fp32 accumulate) or
 * fp32 (split into bf16 hi/lo terms on the fly, |error| ~1e-6).  The 100-image and (336,224) averaging of
 * compute.py:75-81 is host arithmetic on pair_score (see a_score.py).
 * ------------------------------------------------------------------------------------------------- */
size_t lovr_a_score_workspace_bytes(int32_t dtype, int64_t n_pairs, int32_t x, int32_t y, int32_t D);
int lovr_a_score_maxcos(const void* other, const void* ref, int32_t dtype, int64_t n_pairs, int32_t x, int32_t y,
                        int32_t D, float* pair_score, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * C score (replaces C_score/pck_train.py:24-29 normalize_feats + C_score/utils/utils_correspondence.py:345-382
 * calculate_keypoint_transformation / :297-337 get_flow / :234-256 soft_argmax, and pck_train.py:150-163 PCK).
 * desc1/desc2 [n_pairs, P*P, D] UN-normalised patch descriptors (bf16 or fp32), rows in row-major patch order.
 *   lovr_c_score_flow:      the dense path — normalise, P^2 x P^2 similarity on tcgen05 (fp32 inputs are split into
 *                           bf16 hi/lo terms: |error| ~1e-6), row argmax, windowed soft-argmax over every source row.
 *                           flow [n_pairs, P*P, 2] = (x, y) in patch units == get_flow(...).reshape(-1, 2);
 *                           argmax [n_pairs, P*P] (may be NULL) == sim.argmax(-1).  soft=0 gives the argmax coordinates.
 *   lovr_c_score_keypoints: what the reference's callers consume — only the K source-keypoint rows
 *                           (patch_idx [n_pairs, K] from kpts_to_patch_idx :384-389), similarity in plain fp32;
 *                           kps_out [n_pairs, K, 2] = (x, y) pixels = clamp(flow)*stride + stride//2, stride=anno/P.
 *   lovr_pck:               out [n_pairs, 4] = PCK@0.10, @0.05, @0.01 over visible keypoints, #visible.
 * Workspace: lovr_c_score_workspace_bytes(dtype, n_pairs, P, D, K) with K=0 for the dense path.
 * ------------------------------------------------------------------------------------------------- */
size_t lovr_c_score_workspace_bytes(int32_t dtype, int64_t n_pairs, int32_t P, int32_t D, int32_t K);
int lovr_c_score_flow(const void* desc1, const void* desc2, int32_t dtype, int64_t n_pairs, int32_t P, int32_t D,
                      int32_t window, int32_t soft, float* flow, int32_t* argmax, void* workspace,
                      size_t workspace_bytes, void* stream);
int lovr_c_score_keypoints(const void* desc1, const void* desc2, int32_t dtype, int64_t n_pairs, int32_t P, int32_t D,
                           const int32_t* patch_idx, int32_t K, int32_t window, int32_t soft, float anno_size,
                           float* kps_out, void* workspace, size_t workspace_bytes, void* stream);
/* Adaptive-flip test (C_score/utils/utils_correspondence.py:55-73 get_distance_mutual_nn; caller pck_train.py:111-126):
 * out[b] = {mean over mutual nearest neighbours i <-> j of ||n(d1_i) - n(d2_j)||, number of mutual pairs}
 * (NaN mean when there is none, like torch's mean of an empty selection).  Same dense similarity as lovr_c_score_flow;
 * dist^2 = |n1|^2 + |n2|^2 - 2 n1.n2 (torch.cdist's matmul form). */
size_t lovr_c_score_mutual_nn_workspace_bytes(int32_t dtype, int64_t n_pairs, int32_t P, int32_t D);
int lovr_c_score_mutual_nn(const void* desc1, const void* desc2, int32_t dtype, int64_t n_pairs, int32_t P, int32_t D,
                           float* out /* [n_pairs, 2] */, void* workspace, size_t workspace_bytes, void* stream);
/* Two-model descriptors (C_score/pck_train_two.py:24-35): out[r, :] = [ n(a[r, :]) , n(b[r, :]) ], n(x) = x / (||x|| + 1e-10);
 * a [rows, Da], b [rows, Db], out [rows, Da + Db], all `dtype` (bf16 or fp32).  The renormalisation of the concatenation
 * that follows in the reference is fused into the similarity kernels above. */
int lovr_c_score_concat_two(const void* a, const void* b, int32_t dtype, int64_t rows, int32_t Da, int32_t Db, void* out,
                            void* stream);
int lovr_pck(const float* pred, const float* gt, const uint8_t* visible, const float* thresholds, int64_t n_pairs,
             int32_t K, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Linear CKA (the metric BASELINE.json north_star names; ABSENT from the reference — builder-defined, parity pinned only
 * to the fp64 restatement in oracle/a_score_oracle.py:linear_cka).  X [rows, Dx], Y [rows, Dy] bf16, rows = all
 * N*P tokens.  lovr_cka_accumulate ADDS this shard's uncentred sums into `gram` (fp64, lovr_cka_gram_elems(Dx,Dy)
 * elements: X^T[X|Y] [Dx, Dx+Dy], Y^T Y [Dy, Dy], column sums [Dx+Dy], row count [1]; zero it first) — that ONE buffer
 * is what the single all-reduce of the multi-GPU path sums over ranks (the row count rides in its last element).
 * lovr_cka_finalize applies the centring G - s s^T / n (n = n_rows if > 0, else the count in the buffer) and
 * returns out3 = {||Xc^T Yc||_F^2, ||Xc^T Xc||_F^2, ||Yc^T Yc||_F^2};  CKA = out3[0] / sqrt(out3[1]*out3[2]).
 * ------------------------------------------------------------------------------------------------- */
size_t lovr_cka_workspace_bytes(int64_t rows, int32_t Dx, int32_t Dy);
size_t lovr_cka_gram_elems(int32_t Dx, int32_t Dy);
int lovr_cka_accumulate(const void* X, const void* Y, int64_t rows, int32_t Dx, int32_t Dy, double* gram,
                        void* workspace, size_t workspace_bytes, void* stream);
int lovr_cka_finalize(const double* gram, double n_rows, int32_t Dx, int32_t Dy, double* out3, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Image preprocessing in front of the tower.  Replaces, bit for bit, the CPU path
 *   llava/mm_utils.py:63-95  process_images / expand2square -> HF CLIPImageProcessor.preprocess
 *                            (transformers 4.31 slow processor: PIL BICUBIC resize of the shortest edge, centre crop,
 *                            float32(float64(v) * 1/255), (x - mean) / std in float32),
 *   llava/model/multimodal_encoder/dinov2_encoder.py:24-26 (BitImageProcessor 256/224 and the 336x336 override),
 *   SiglipImageProcessor (exact 224x224), and
 *   C_score/extract_feature.py:65-67  img.resize((S,S)) [PIL default = BICUBIC] + (PILToTensor()(img)/255.0 - 0.5)*2.
 * images: DEVICE buffer of n_img packed 8-bit RGB images (HWC, 3 bytes per pixel, rows contiguous);
 * offsets_host[i] = byte offset of image i, hw_host[2*i], hw_host[2*i+1] = its height and width (HOST arrays: the
 * library derives every per-image size from them and enqueues a small descriptor upload on `stream`).
 * out [n_img, 3, crop_size, crop_size] in out_dtype (LOVR_DT_*).  Pillow's two-pass 8-bit resampling
 * (libImaging/Resample.c; restated in oracle/preprocess_oracle.py) is integer arithmetic: results equal Pillow's.
 * ------------------------------------------------------------------------------------------------- */
enum { LOVR_RESIZE_SHORTEST_EDGE = 0, LOVR_RESIZE_EXACT = 1 };
enum { LOVR_PIX_HF_RESCALE_NORMALIZE = 0, LOVR_PIX_CSCORE = 1 };
typedef struct {
  int32_t resize_mode;   /* LOVR_RESIZE_SHORTEST_EDGE: shortest edge -> resize_size, aspect kept (int(size*long/short));
                            LOVR_RESIZE_EXACT: (resize_size, resize_size) */
  int32_t resize_size;
  int32_t crop_size;     /* centre crop (HF center_crop: top = (h - crop)//2); must be <= both resized sides */
  int32_t pad_square;    /* 1: expand2square first (mm_utils.py:63-75, image_aspect_ratio == 'pad') */
  uint8_t pad_rgb[4];    /* background colour of the square canvas: int(mean*255) per channel */
  int32_t pixel_map;     /* LOVR_PIX_* : 8-bit value -> float */
  float mean[3];
  float std[3];
} lovr_preprocess_config;

size_t lovr_preprocess_workspace_bytes(const lovr_preprocess_config* cfg, int32_t n_img, const int32_t* hw_host);
int lovr_preprocess(const uint8_t* images, const int64_t* offsets_host, const int32_t* hw_host, int32_t n_img,
                    const lovr_preprocess_config* cfg, void* out, int32_t out_dtype, void* workspace,
                    size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LOVR_B200_H_ */
