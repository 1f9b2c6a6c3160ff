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


def ln_fold_weights(W, gamma, beta, bias, op_dtype):
    """lovr_ln_fold_weights: W [N,K] fp32 (device) -> (W16 = 16-bit(W diag(gamma)), cw [N], bw [N])."""
    N, K = W.shape
    W16 = torch.empty(N, K, device=W.device, dtype=op_dtype)
    cw = torch.empty(N, device=W.device, dtype=torch.float32)
    bw = torch.empty(N, device=W.device, dtype=torch.float32)
    _lib.check(_lib.lib().lovr_ln_fold_weights(_lib.ptr(W.contiguous()), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(bias), N, K,
                                               _lib.torch_dtype_code(op_dtype), _lib.ptr(W16), _lib.ptr(cw), _lib.ptr(bw),
                                               _lib.stream_ptr()), "lovr_ln_fold_weights")
    return W16, cw, bw


def gemm_resid_lnfold(A, W, x_inout, bias=None, colscale=None):
    """x += colscale * (A W^T + bias) in place; -> (xhat 16-bit [M,N], stats [M, N/128, 4])."""
    M, K = A.shape
    N = W.shape[0]
    xhat = torch.empty(M, N, device=A.device, dtype=A.dtype)
    stats = torch.zeros(M, N // 128, 4, device=A.device, dtype=torch.float32)
    _lib.check(_lib.lib().lovr_gemm_lnfold(_lib.ptr(A), A.stride(0), _lib.ptr(W), W.stride(0), _lib.ptr(x_inout), x_inout.stride(0),
                                           M, N, K, _lib.EPI_RESID_F32, 0, _lib.torch_dtype_code(A.dtype), _lib.ptr(bias),
                                           _lib.ptr(colscale), None, None, 0, 0.0, _lib.ptr(stats), _lib.ptr(xhat), xhat.stride(0),
                                           _lib.stream_ptr()), "lovr_gemm_lnfold")
    return xhat, stats


def gemm_consume_lnfold(xhat, stats, W16, cw, bw, eps, act=0):
    """act(LN(x) W^T + b) from x^ / statistics / folded weights -> 16-bit [M, N]."""
    M, K = xhat.shape
    N = W16.shape[0]
    C_ = torch.empty(M, N, device=xhat.device, dtype=xhat.dtype)
    _lib.check(_lib.lib().lovr_gemm_lnfold(_lib.ptr(xhat), xhat.stride(0), _lib.ptr(W16), W16.stride(0), _lib.ptr(C_), C_.stride(0),
                                           M, N, K, _lib.EPI_BF16, act, _lib.torch_dtype_code(xhat.dtype), _lib.ptr(bw), None,
                                           _lib.ptr(stats), _lib.ptr(cw), K, eps, None, None, 0, _lib.stream_ptr()),
               "lovr_gemm_lnfold")
    return C_
