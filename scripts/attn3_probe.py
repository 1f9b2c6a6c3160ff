"""attention3 A/B timings (LOVR_ATTN3_PIPE) against attention2 — diagnostics only."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gpu_util as G
from law_of_vision_representation_in_mllms_b200 import _lib
from bench_kernels import timeit
H = 16
for n, T in ((64, 577), (128, 577), (128, 257), (256, 196), (37, 577), (3, 300)):
    for dt in (torch.bfloat16, torch.float16):
        qkv = torch.randn(n * T, 3 * H * 64, device="cuda").to(dt)
        ref = G.attention(qkv, n, T, H, _lib.ATTN_SIMT_CHECK).float()
        for impl in (3, 4):
            out = G.attention(qkv, n, T, H, impl).float()
            print(json.dumps(dict(n=n, T=T, dt=str(dt), impl=impl, ms=timeit(lambda: G.attention(qkv, n, T, H, impl)),
                                  rel_vs_simt=((out - ref).norm() / ref.norm()).item())), flush=True)
        out = G.attention(qkv, n, T, H, _lib.ATTN_TCGEN05_V4).float()
        torch.cuda.synchronize()
        print(json.dumps(dict(n=n, T=T, dt=str(dt), impl=7, ms=timeit(lambda: G.attention(qkv, n, T, H, _lib.ATTN_TCGEN05_V4)),
                              rel_vs_simt=((out - ref).norm() / ref.norm()).item())), flush=True)
        for pipe in ("1",):
            os.environ["LOVR_ATTN3_PIPE"] = pipe
            try:
                out = G.attention(qkv, n, T, H, _lib.ATTN_TCGEN05_V3).float()
                torch.cuda.synchronize()
                err = ((out - ref).norm() / ref.norm()).item()
                ms = timeit(lambda: G.attention(qkv, n, T, H, _lib.ATTN_TCGEN05_V3))
                print(json.dumps(dict(n=n, T=T, dt=str(dt), impl=6, pipe=pipe, ms=ms, rel_vs_simt=err)), flush=True)
            except Exception as ex:
                print(json.dumps(dict(impl=6, pipe=pipe, error=str(ex)[:200])), flush=True)
