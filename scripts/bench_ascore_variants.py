"""A score kernel forms side by side (same process, env switches are read per call): 1-CTA / 2-CTA, rotated tile order on / off,
at 576 tokens (C3) and 256 tokens (224-pixel towers).  Median of 5, CUDA events."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from law_of_vision_representation_in_mllms_b200 import a_score as A   # noqa: E402

dev = torch.device("cuda", 0)


def med(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[reps // 2], out


for T, n in ((576, 10000), (256, 22500)):
    g = torch.Generator(device=dev).manual_seed(11)
    po, pr = [], []
    for b in range(0, n, 250):
        o = torch.randn(min(250, n - b), T, 1024, generator=g, device=dev)
        r = 0.5 * o + torch.randn(o.shape, generator=g, device=dev)
        po.append(o.bfloat16()); pr.append(r.bfloat16())
        del o, r
    o, r = torch.cat(po), torch.cat(pr)
    del po, pr
    for two in ("0", "1"):
        for rot in ("0", "1"):
            os.environ["LOVR_A_SCORE_2CTA"], os.environ["LOVR_A_ROTATE"] = two, rot
            ms, s = med(lambda: A.a_score_pairs(o, r).double().sum())
            print(f"tokens={T} pairs={n} 2cta={two} rotate={rot}: {ms:.3f} ms  mean_a {s.item() / n:.6f}  "
                  f"{2.0 * n * T * T * 1024 / ms / 1e9:.0f} TFLOP/s  {2.0 * n * T * 1024 * 2 / ms / 1e6:.0f} GB/s", flush=True)
    del o, r
    torch.cuda.empty_cache()
