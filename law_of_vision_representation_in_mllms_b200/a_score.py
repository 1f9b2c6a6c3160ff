"""A score on the device (reference: A_score/compute.py — max-cosine matching, NOT CKA).

    a(o, r)  = mean_i max_j cos(n(o_i), n(r_j))            compute.py:54-72   -> lovr_a_score_maxcos
    A(rep)   = ( mean_k a(o_k, clip336_k) + mean_k a(o_k, clip224_k) ) / 2     compute.py:75-81 (host, float64)

`compute_a_scores(base_folder, subfolders)` mirrors the script's behaviour (same folder layout, same skip
rules, same printed lines) with the two hard-coded module constants (:7, :10) turned into arguments.
"""
import ctypes as C
import os
from collections import defaultdict

import torch

from . import _lib

_MAX_PAIRS_PER_CALL = 4096


def a_score_pairs(other, ref):
    """other [B,x,D], ref [B,y,D] (same dtype: bf16 or fp32; fp16 is widened to fp32) -> [B] fp32 on the device."""
    if other.dim() != 3 or ref.dim() != 3 or other.shape[0] != ref.shape[0] or other.shape[2] != ref.shape[2]:
        raise ValueError(f"expected [B,x,D] and [B,y,D]; got {tuple(other.shape)} and {tuple(ref.shape)}")
    if not other.is_cuda:
        raise _lib.LovrError("a_score_pairs needs CUDA tensors: this package has no CPU path")
    if other.dtype not in (torch.bfloat16, torch.float32):
        other, ref = other.float(), ref.float()
    ref = ref.to(other.dtype)
    B, x, D = other.shape
    y = ref.shape[1]
    if D % 64:
        pad = 64 - D % 64          # zero channels change neither dot products nor norms
        other = torch.nn.functional.pad(other, (0, pad))
        ref = torch.nn.functional.pad(ref, (0, pad))
        D += pad
    other, ref = other.contiguous(), ref.contiguous()
    L = _lib.lib()
    dt = _lib.torch_dtype_code(other.dtype)
    out = torch.empty(B, dtype=torch.float32, device=other.device)
    step = min(B, _MAX_PAIRS_PER_CALL)
    ws_bytes = L.lovr_a_score_workspace_bytes(dt, step, x, y, D)
    ws = torch.empty(ws_bytes + 1024, dtype=torch.uint8, device=other.device)
    base = (ws.data_ptr() + 1023) // 1024 * 1024
    with torch.cuda.device(other.device):
        st = _lib.stream_ptr(other.device)
        for s in range(0, B, step):
            e = min(B, s + step)
            _lib.check(L.lovr_a_score_maxcos(_lib.ptr(other[s:e]), _lib.ptr(ref[s:e]), dt, e - s, x, y, D,
                                             _lib.ptr(out[s:e]), C.c_void_p(base), ws_bytes, st),
                       "lovr_a_score_maxcos")
    return out


def a_pairs_list(others, refs, device="cuda"):
    """Per-image a(o_k, r_k) for lists of [T,D] tensors whose shapes may differ between images
    (compute.py zips three lists).  Images are grouped by shape and each group runs as one batch."""
    groups = defaultdict(list)
    for k, (o, r) in enumerate(zip(others, refs)):
        groups[(tuple(o.shape), tuple(r.shape), o.dtype)].append(k)
    res = [None] * min(len(others), len(refs))
    for idx in groups.values():
        o = torch.stack([others[k] for k in idx]).to(device)
        r = torch.stack([refs[k] for k in idx]).to(device)
        vals = a_score_pairs(o, r).double().cpu().tolist()
        for k, v in zip(idx, vals):
            res[k] = v
    return res


def a_score(others, refs336, refs224, device="cuda"):
    """compute.py:75-81: average per reference over the images, then the mean of the two references."""
    s336 = a_pairs_list(others, refs336, device)
    s224 = a_pairs_list(others, refs224, device)
    avg336 = sum(s336) / len(s336)
    avg224 = sum(s224) / len(s224)
    return (avg336 + avg224) / 2, s336, s224


def load_tensors(base_folder, subfolder, count=100):
    """compute.py:18-28 — tensor_1.pt .. tensor_100.pt; any load error -> [] (the representation is skipped)."""
    tensors = []
    for i in range(1, count + 1):
        tensor_path = os.path.join(base_folder, subfolder, f"tensor_{i}.pt")
        try:
            tensors.append(torch.load(tensor_path, map_location="cpu"))
        except Exception as e:  # noqa: BLE001 — the reference swallows every error the same way
            print(f"Error loading {tensor_path}: {e}")
            return []
    return tensors


def compute_a_scores(base_folder, subfolders, device="cuda", verbose=True):
    """Folder-level driver with the behaviour of A_score/compute.py:31-85."""
    clip336 = load_tensors(base_folder, "clip336")
    clip224 = load_tensors(base_folder, "clip224")
    if not clip336 or not clip224:
        raise ValueError("Failed to load tensors from 'clip336' or 'clip224' subfolder")
    results = {}
    for sub in subfolders:
        other = load_tensors(base_folder, sub)
        if not other:
            print(f"Skipping {sub} due to loading error.")
            continue
        results[sub], _, _ = a_score(other, clip336, clip224, device)
    if verbose:
        for sub, v in results.items():
            print(f"Average cosine similarity between clip224+clip336 and {sub}: {v}")
    return results
