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
    assert L.lovr_abi_version() == 2


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


def test_header_is_plain_c_and_cpp(tmp_path):
    """The boundary header must be consumable by a C compiler (cgo / JNI / ctypes-style hosts) and a C++ one."""
    import shutil
    import subprocess
    hdr = os.path.join(ROOT, "include", "lovr_b200.h")
    src_c = tmp_path / "use.c"
    src_c.write_text('#include "lovr_b200.h"\nint main(void) { lovr_tower_config c; lovr_preprocess_config p; (void)c; (void)p; '
                     'return lovr_abi_version() == LOVR_ABI_VERSION ? 0 : 1; }\n')
    for cc, std, src in (("gcc", "-std=c99", src_c), ("g++", "-std=c++11", src_c)):
        if shutil.which(cc) is None:
            pytest.skip(f"{cc} not installed")
        extra = ["-x", "c++"] if cc == "g++" else []
        r = subprocess.run([cc, std, "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.dirname(hdr)] + extra + [str(src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_c_program_links_and_calls_the_library(tmp_path):
    """A C host (what a cgo / JNI shim compiles to) links liblovr_b200.so and calls device-free entry points."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not installed")
    _lib.lib()      # builds nothing, but fails loudly if the library is absent
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include "lovr_b200.h"
int main(void) {
  lovr_preprocess_config cfg = {LOVR_RESIZE_SHORTEST_EDGE, 336, 336, 0, {0, 0, 0, 0}, LOVR_PIX_HF_RESCALE_NORMALIZE,
                                {0.5f, 0.5f, 0.5f}, {0.5f, 0.5f, 0.5f}};
  int32_t hw[4] = {480, 640, 500, 333};
  size_t need = lovr_preprocess_workspace_bytes(&cfg, 2, hw);
  cfg.crop_size = 1000;                                   /* larger than the resized image: an argument error */
  size_t bad = lovr_preprocess_workspace_bytes(&cfg, 2, hw);
  printf("%d %zu %zu %s\n", lovr_abi_version(), need, bad, lovr_last_error());
  return 0;
}
''')
    exe = tmp_path / "host"
    libdir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-llovr_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    ver, need, bad, msg = out.stdout.split(maxsplit=3)
    assert int(ver) == 2 and int(need) > 2 * 3 * 336 * 336 and int(bad) == 0 and "crop_size" in msg
