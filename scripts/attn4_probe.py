"""attention4 (persistent) vs attention3 timings over the control-warp register budget — diagnostics only."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gpu_util as G
from law_of_vision_representation_in_mllms_b200 import _lib
from bench_kernels import timeit
H = 16
for n, T in ((64, 577), (128, 577), (128, 257), (256, 196)):
    qkv = torch.randn(n * T, 3 * H * 64, device="cuda").bfloat16()
    ref = G.attention(qkv, n, T, H, 6)
    r = dict(n=n, T=T, v3=round(timeit(lambda: G.attention(qkv, n, T, H, 6), iters=20), 4))
    for rc in os.environ.get("RCS", "40,56,80,104,128").split(","):
        os.environ["LOVR_ATTN4_RC"] = rc
        ok = torch.equal(G.attention(qkv, n, T, H, 7), ref)
        r[f"v4_rc{rc}"] = (round(timeit(lambda: G.attention(qkv, n, T, H, 7), iters=20), 4), ok)
    print(json.dumps(r), flush=True)
