"""CPU-only checks of the C-ABI boundary: the library builds/loads, exports every symbol the header declares,
and refuses to compute without an sm_100 device (no fallback)."""
import os
import re

import pytest
import torch

from law_of_vision_representation_in_mllms_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "lovr_b200.h")).read()
    declared = set(re.findall(r"\b(lovr_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.lovr_abi_version() == 1


def test_workspace_queries_need_no_gpu():
    L = _lib.lib()
    cfg = _lib.TowerConfig(hidden_size=1024, intermediate_size=4096, num_layers=24, num_heads=16, image_size=336,
                           patch_size=14, act=1, has_cls=1, pre_ln=1, layerscale=0, ln_eps=1e-5, patch_k_padded=640,
                           operand_dtype=_lib.DT_BF16)
    import ctypes as C
    b1 = L.lovr_tower_workspace_bytes(C.byref(cfg), 1)
    b8 = L.lovr_tower_workspace_bytes(C.byref(cfg), 8)
    M = 577
    assert b1 >= M * 1024 * 4 + M * 1024 * 2 + M * 4096 * 2
    assert 7.9 * b1 < b8 < 8.1 * b1
    assert L.lovr_a_score_workspace_bytes(_lib.DT_BF16, 10, 576, 576, 1024) >= 3 * 10 * 576 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    L = _lib.lib()
    x = torch.zeros(128, 64, dtype=torch.bfloat16)
    out = torch.zeros(128, 256)
    rc = L.lovr_gemm_bf16(_lib.ptr(x), 64, _lib.ptr(torch.zeros(256, 64, dtype=torch.bfloat16)), 64, _lib.ptr(out), 256,
                          128, 256, 64, _lib.EPI_F32, 0, _lib.DT_BF16, None, None, None)
    assert rc in (_lib.E_UNSUPPORTED, _lib.E_CUDA)
    assert L.lovr_last_error()
