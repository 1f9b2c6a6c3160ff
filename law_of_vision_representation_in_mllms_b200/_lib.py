"""ctypes binding of liblovr_b200.so (the C ABI in include/lovr_b200.h).

The shared library is built in-tree by `build.py` (nvcc, sm_100a).  There is no fallback: if the
library is missing or a call fails, the Python layer raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LOVR_B200_LIB: load an experiment build of the same ABI instead (build.py --out ...); never a different implementation
LIB_PATH = os.environ.get("LOVR_B200_LIB") or os.path.join(_HERE, "liblovr_b200.so")

OK, E_ARG, E_CUDA, E_UNSUPPORTED, E_NOMEM = 0, -1, -2, -3, -4
DT_F32, DT_BF16, DT_F16 = 0, 1, 2
EPI_BF16, EPI_RESID_F32, EPI_F32 = 0, 1, 2
ATTN_SIMT_CHECK = 1
ATTN_AUTO = 5
ATTN_TCGEN05_V3 = 6
ATTN_TCGEN05_V4 = 7
RESIZE_SHORTEST_EDGE, RESIZE_EXACT = 0, 1
PIX_HF_RESCALE_NORMALIZE, PIX_CSCORE = 0, 1

# every symbol include/lovr_b200.h declares (tests/test_abi.py checks the library exports all of them)
EXPORTS = [
    "lovr_abi_version", "lovr_last_error", "lovr_launch_count", "lovr_prof_enable", "lovr_prof_read", "lovr_prof_read_kinds",
    "lovr_tower_workspace_bytes", "lovr_tower_create", "lovr_tower_destroy", "lovr_tower_forward",
    "lovr_gemm_bf16", "lovr_gemm_lnfold", "lovr_ln_fold_weights", "lovr_layernorm_bf16", "lovr_attention", "lovr_im2col",
    "lovr_a_score_workspace_bytes", "lovr_a_score_maxcos",
    "lovr_c_score_workspace_bytes", "lovr_c_score_flow", "lovr_c_score_keypoints", "lovr_pck",
    "lovr_c_score_mutual_nn_workspace_bytes", "lovr_c_score_mutual_nn", "lovr_c_score_concat_two",
    "lovr_cka_workspace_bytes", "lovr_cka_gram_elems", "lovr_cka_accumulate", "lovr_cka_finalize",
    "lovr_preprocess_workspace_bytes", "lovr_preprocess",
]


class TowerConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("hidden_size", "intermediate_size", "num_layers", "num_heads", "image_size",
                                         "patch_size", "act", "has_cls", "pre_ln", "layerscale")] + \
               [("ln_eps", C.c_float), ("patch_k_padded", C.c_int32), ("operand_dtype", C.c_int32), ("ln_fold", C.c_int32)]


class LayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "o_w", "o_b", "ls1", "ln2_w", "ln2_b",
                                          "fc1_w", "fc1_b", "fc2_w", "fc2_b", "ls2", "qkv_cw", "fc1_cw")]


class TowerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("patch_w", "patch_b", "cls", "pos", "pre_ln_w", "pre_ln_b")] + \
               [("layers_host", C.POINTER(LayerWeights))]


class PreprocessConfig(C.Structure):
    _fields_ = [("resize_mode", C.c_int32), ("resize_size", C.c_int32), ("crop_size", C.c_int32),
                ("pad_square", C.c_int32), ("pad_rgb", C.c_uint8 * 4), ("pixel_map", C.c_int32),
                ("mean", C.c_float * 3), ("std", C.c_float * 3)]


class LovrError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the shared library once; raise loudly if it is absent (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LovrError(
            f"{LIB_PATH} not found: build it with `python -m law_of_vision_representation_in_mllms_b200.build` "
            "(nvcc, sm_100a). This package has no CPU or PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
    L.lovr_abi_version.restype = C.c_int
    L.lovr_last_error.restype = C.c_char_p
    L.lovr_launch_count.restype = C.c_uint64
    L.lovr_prof_enable.restype = None
    L.lovr_prof_enable.argtypes = [i32]
    L.lovr_prof_read.restype = C.c_int
    L.lovr_prof_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.lovr_prof_read_kinds.restype = C.c_int
    L.lovr_prof_read_kinds.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.lovr_tower_workspace_bytes.restype = sz
    L.lovr_tower_workspace_bytes.argtypes = [C.POINTER(TowerConfig), i32]
    L.lovr_tower_create.restype = C.c_int
    L.lovr_tower_create.argtypes = [C.POINTER(vp), C.POINTER(TowerConfig), C.POINTER(TowerWeights), vp, sz, i32]
    L.lovr_tower_destroy.restype = None
    L.lovr_tower_destroy.argtypes = [vp]
    L.lovr_tower_forward.restype = C.c_int
    L.lovr_tower_forward.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, i64, i64, i32, vp]
    L.lovr_gemm_bf16.restype = C.c_int
    L.lovr_gemm_bf16.argtypes = [vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    L.lovr_gemm_lnfold.restype = C.c_int
    L.lovr_gemm_lnfold.argtypes = [vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, i32, f32, vp, vp,
                                   i64, vp]
    L.lovr_ln_fold_weights.restype = C.c_int
    L.lovr_ln_fold_weights.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]
    L.lovr_layernorm_bf16.restype = C.c_int
    L.lovr_layernorm_bf16.argtypes = [vp, vp, vp, f32, vp, i64, i32, i32, vp]
    L.lovr_attention.restype = C.c_int
    L.lovr_attention.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    L.lovr_im2col.restype = C.c_int
    L.lovr_im2col.argtypes = [vp, i32, vp, i32, i32, i32, i32, i32, vp]
    L.lovr_a_score_workspace_bytes.restype = sz
    L.lovr_a_score_workspace_bytes.argtypes = [i32, i64, i32, i32, i32]
    L.lovr_a_score_maxcos.restype = C.c_int
    L.lovr_a_score_maxcos.argtypes = [vp, vp, i32, i64, i32, i32, i32, vp, vp, sz, vp]
    L.lovr_c_score_workspace_bytes.restype = sz
    L.lovr_c_score_workspace_bytes.argtypes = [i32, i64, i32, i32, i32]
    L.lovr_c_score_flow.restype = C.c_int
    L.lovr_c_score_flow.argtypes = [vp, vp, i32, i64, i32, i32, i32, i32, vp, vp, vp, sz, vp]
    L.lovr_c_score_keypoints.restype = C.c_int
    L.lovr_c_score_keypoints.argtypes = [vp, vp, i32, i64, i32, i32, vp, i32, i32, i32, f32, vp, vp, sz, vp]
    L.lovr_c_score_mutual_nn_workspace_bytes.restype = sz
    L.lovr_c_score_mutual_nn_workspace_bytes.argtypes = [i32, i64, i32, i32]
    L.lovr_c_score_mutual_nn.restype = C.c_int
    L.lovr_c_score_mutual_nn.argtypes = [vp, vp, i32, i64, i32, i32, vp, vp, sz, vp]
    L.lovr_c_score_concat_two.restype = C.c_int
    L.lovr_c_score_concat_two.argtypes = [vp, vp, i32, i64, i32, i32, vp, vp]
    L.lovr_pck.restype = C.c_int
    L.lovr_pck.argtypes = [vp, vp, vp, vp, i64, i32, vp, vp]
    L.lovr_cka_workspace_bytes.restype = sz
    L.lovr_cka_workspace_bytes.argtypes = [i64, i32, i32]
    L.lovr_cka_gram_elems.restype = sz
    L.lovr_cka_gram_elems.argtypes = [i32, i32]
    L.lovr_cka_accumulate.restype = C.c_int
    L.lovr_cka_accumulate.argtypes = [vp, vp, i64, i32, i32, vp, vp, sz, vp]
    L.lovr_cka_finalize.restype = C.c_int
    L.lovr_cka_finalize.argtypes = [vp, C.c_double, i32, i32, vp, vp]
    L.lovr_preprocess_workspace_bytes.restype = sz
    L.lovr_preprocess_workspace_bytes.argtypes = [C.POINTER(PreprocessConfig), i32, C.POINTER(C.c_int32)]
    L.lovr_preprocess.restype = C.c_int
    L.lovr_preprocess.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int32), i32, C.POINTER(PreprocessConfig), vp,
                                  i32, vp, sz, vp]
    _lib = L
    return L


def check(rc, what=""):
    if rc != OK:
        msg = lib().lovr_last_error().decode("utf-8", "replace")
        raise LovrError(f"{what or 'lovr call'} failed ({rc}): {msg}")


def launch_count():
    return int(lib().lovr_launch_count())


def prof_enable(on):
    lib().lovr_prof_enable(1 if on else 0)


def prof_read():
    """-> (gemm_ms, gemm_flops, gemm_launches) accumulated since the last read."""
    ms, fl, n = C.c_double(), C.c_double(), C.c_uint64()
    check(lib().lovr_prof_read(C.byref(ms), C.byref(fl), C.byref(n)), "lovr_prof_read")
    return ms.value, fl.value, int(n.value)


PROF_KINDS = ("gemm", "attention", "layernorm", "other")


def prof_read_kinds():
    """-> {kind: (ms, work, launches)} accumulated since the last read; work = FLOPs (gemm, attention) or bytes."""
    n = len(PROF_KINDS)
    ms, work, cnt = (C.c_double * n)(), (C.c_double * n)(), (C.c_uint64 * n)()
    check(lib().lovr_prof_read_kinds(ms, work, cnt), "lovr_prof_read_kinds")
    return {k: (ms[i], work[i], int(cnt[i])) for i, k in enumerate(PROF_KINDS)}


def torch_dtype_code(dtype):
    import torch
    try:
        return {torch.float32: DT_F32, torch.bfloat16: DT_BF16, torch.float16: DT_F16}[dtype]
    except KeyError:
        raise LovrError(f"unsupported dtype {dtype}: the device path takes float32, bfloat16 or float16") from None


def stream_ptr(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
