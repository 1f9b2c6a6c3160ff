"""Linear CKA between two representations on the device (metric named by BASELINE.json; the reference has no CKA —
see SURVEY.md §0 F1 — so this is builder-defined and pinned only to the fp64 restatement in oracle/a_score_oracle.py).

    CKA(X, Y) = ||Yc^T Xc||_F^2 / (||Xc^T Xc||_F ||Yc^T Yc||_F),   Xc = X - mean_rows(X),  rows = all N*P tokens

`cka_gram` accumulates the uncentred fp64 sums of one shard (rows may be processed in slabs) and the shard's row count
into ONE buffer; the multi-GPU path all-reduces that buffer once (distributed.cka_reduce) and calls `cka_from_gram`.
"""
import ctypes as C

import torch

from . import _lib

_SLAB_ROWS = 1 << 30        # rows per native call (the native limit is 2^31 - 2048 rows)


def _as_rows(t):
    t = t.reshape(-1, t.shape[-1])
    if t.dtype != torch.bfloat16:
        t = t.to(torch.bfloat16)
    return t.contiguous()


def _pad256(t):
    D = t.shape[1]
    if D % 256:
        t = torch.nn.functional.pad(t, (0, 256 - D % 256))      # zero features: zero rows/cols of the centred Grams
    return t


def cka_gram(X, Y, gram=None):
    """X [..., Dx], Y [..., Dy] (same leading shape) on the device -> (gram fp64 buffer, n_rows, Dx_padded, Dy_padded)."""
    if not X.is_cuda:
        raise _lib.LovrError("cka needs CUDA tensors: this package has no CPU path")
    X, Y = _pad256(_as_rows(X)), _pad256(_as_rows(Y))
    if X.shape[0] != Y.shape[0]:
        raise ValueError(f"row counts differ: {X.shape[0]} vs {Y.shape[0]}")
    R, Dx, Dy = X.shape[0], X.shape[1], Y.shape[1]
    L = _lib.lib()
    if gram is None:
        gram = torch.zeros(L.lovr_cka_gram_elems(Dx, Dy), dtype=torch.float64, device=X.device)
    slab = min(R, _SLAB_ROWS)
    nbytes = L.lovr_cka_workspace_bytes(slab, Dx, Dy)
    ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=X.device)
    base = C.c_void_p((ws.data_ptr() + 1023) // 1024 * 1024)
    with torch.cuda.device(X.device):
        st = _lib.stream_ptr(X.device)
        for s in range(0, R, slab):
            e = min(R, s + slab)
            _lib.check(L.lovr_cka_accumulate(_lib.ptr(X[s:e]), _lib.ptr(Y[s:e]), e - s, Dx, Dy, _lib.ptr(gram), base,
                                             nbytes, st), "lovr_cka_accumulate")
    return gram, R, Dx, Dy


def cka_from_gram(gram, n_rows, Dx, Dy):
    """n_rows None -> the row count accumulated in the buffer's last element (what the multi-GPU path all-reduces)."""
    out = torch.empty(3, dtype=torch.float64, device=gram.device)
    with torch.cuda.device(gram.device):
        _lib.check(_lib.lib().lovr_cka_finalize(_lib.ptr(gram), float(n_rows) if n_rows else 0.0, Dx, Dy, _lib.ptr(out),
                                                _lib.stream_ptr(gram.device)), "lovr_cka_finalize")
    xy, xx, yy = out.cpu().tolist()
    return xy / (xx * yy) ** 0.5


def linear_cka(X, Y):
    gram, n, Dx, Dy = cka_gram(X, Y)
    return cka_from_gram(gram, n, Dx, Dy)
