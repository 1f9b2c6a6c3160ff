"""attention4 A/B timings (LOVR_ATTN4_POLY) — diagnostics only."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gpu_util as G
from law_of_vision_representation_in_mllms_b200 import _lib
from bench_kernels import timeit
H = 16
for n, T in ((64, 577), (128, 577), (128, 257), (256, 196)):
    for dt in (torch.bfloat16, torch.float16):
        qkv = torch.randn(n * T, 3 * H * 64, device="cuda").to(dt)
        ref = G.attention(qkv, n, T, H, _lib.ATTN_SIMT_CHECK).float()
        r = dict(n=n, T=T, dt=str(dt)[6:])
        for poly in ("0", "4", "2"):
            os.environ["LOVR_ATTN4_POLY"] = poly
            out = G.attention(qkv, n, T, H, 7).float()
            r[f"poly{poly}"] = (round(timeit(lambda: G.attention(qkv, n, T, H, 7), iters=20), 4),
                                round(((out - ref).norm() / ref.norm()).item(), 6))
        print(json.dumps(r), flush=True)
