"""Timing experiments on the attention kernels (LOVR_ATTN_DEBUG toggles) — diagnostics only."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gpu_util as G
from law_of_vision_representation_in_mllms_b200 import _lib
from bench_kernels import timeit
n, T, H = 64, 577, 16
qkv = torch.randn(n * T, 3 * H * 64, device="cuda").bfloat16()
for impl in (0, 2, 3, 4):
    for dbg in (0, 2):
        if impl == 0 and dbg: continue
        os.environ["LOVR_ATTN_DEBUG"] = str(dbg)
        ms = timeit(lambda: G.attention(qkv, n, T, H, impl))
        print(json.dumps(dict(impl=impl, debug=dbg, ms=ms)), flush=True)
for T2, n2 in ((257, 128), (196, 256), (128, 256), (640, 64)):
    os.environ["LOVR_ATTN_DEBUG"] = "0"
    q2 = torch.randn(n2 * T2, 3 * H * 64, device="cuda").bfloat16()
    for impl in (0, 2, 3, 4):
        ms = timeit(lambda: G.attention(q2, n2, T2, H, impl))
        print(json.dumps(dict(impl=impl, T=T2, n=n2, ms=ms, tflops=4 * n2 * H * T2 * T2 * 64 / ms / 1e9)), flush=True)
