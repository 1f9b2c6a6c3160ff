"""Helpers for the -m gpu tests: thin Python wrappers over the C ABI building blocks."""
import ctypes as C

import torch

from law_of_vision_representation_in_mllms_b200 import _lib


def gemm(A, W, epi, act=0, bias=None, colscale=None, C_inout=None):
    od = _lib.torch_dtype_code(A.dtype)
    """A [M,K] bf16, W [N,K] bf16 -> C [M,N] (bf16 for EPI_BF16 else fp32; in place for EPI_RESID_F32)."""
    M, K = A.shape
    N = W.shape[0]
    if C_inout is None:
        C_inout = torch.empty(M, N, device=A.device, dtype=A.dtype if epi == _lib.EPI_BF16 else torch.float32)
    L = _lib.lib()
    _lib.check(L.lovr_gemm_bf16(_lib.ptr(A), A.stride(0), _lib.ptr(W), W.stride(0), _lib.ptr(C_inout),
                                C_inout.stride(0), M, N, K, epi, act, od, _lib.ptr(bias), _lib.ptr(colscale),
                                _lib.stream_ptr()), "lovr_gemm_bf16")
    return C_inout


def layernorm(x, g, b, eps, out_dtype=torch.bfloat16):
    y = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    _lib.check(_lib.lib().lovr_layernorm_bf16(_lib.ptr(x), _lib.ptr(g), _lib.ptr(b), eps, _lib.ptr(y), x.shape[0],
                                              x.shape[1], _lib.torch_dtype_code(out_dtype), _lib.stream_ptr()), "lovr_layernorm_bf16")
    return y


def attention(qkv, n_img, T, H, impl):
    out = torch.empty(n_img * T, H * 64, device=qkv.device, dtype=qkv.dtype)
    _lib.check(_lib.lib().lovr_attention(_lib.ptr(qkv), _lib.ptr(out), n_img, T, H, impl,
                                         _lib.torch_dtype_code(qkv.dtype), _lib.stream_ptr()),
               "lovr_attention")
    return out


def im2col(pixels, patch, kpad, out_dtype=torch.bfloat16):
    n, _, R, _ = pixels.shape
    G = R // patch
    cols = torch.empty(n * G * G, kpad, device=pixels.device, dtype=out_dtype)
    _lib.check(_lib.lib().lovr_im2col(_lib.ptr(pixels), _lib.torch_dtype_code(pixels.dtype), _lib.ptr(cols), n, R,
                                      patch, kpad, _lib.torch_dtype_code(out_dtype), _lib.stream_ptr()), "lovr_im2col")
    return cols


def rel_err(got, want):
    got, want = got.double(), want.double()
    return ((got - want).norm() / want.norm().clamp_min(1e-30)).item()
