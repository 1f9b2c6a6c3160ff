"""A score and linear CKA at the C3 shape (10k x 576 x 1024 bf16 pairs) on one GPU, median of 5, CUDA events."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from law_of_vision_representation_in_mllms_b200 import a_score as A, cka as CK   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(11)
parts_o, parts_r = [], []
for b in range(0, n, 250):
    o = torch.randn(min(250, n - b), 576, 1024, generator=g, device=dev)
    r = 0.5 * o + torch.randn(o.shape, generator=g, device=dev)
    parts_o.append(o.bfloat16()); parts_r.append(r.bfloat16())
    del o, r
o, r = torch.cat(parts_o), torch.cat(parts_r)
del parts_o, parts_r


def med(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[reps // 2], ts, out


ms, ts, s = med(lambda: A.a_score_pairs(o, r).double().sum())
print(f"a_score n={n}: {ms:.3f} ms  runs {['%.2f' % t for t in ts]}  mean_a {s.item() / n:.6f}  "
      f"{2.0 * n * 576 * 576 * 1024 / ms / 1e9:.0f} TFLOP/s  {2.0 * n * 576 * 1024 * 2 / ms / 1e6:.0f} GB/s", flush=True)
ms, ts, v = med(lambda: CK.linear_cka(o, r))
print(f"linear_cka rows={n * 576}: {ms:.3f} ms  runs {['%.2f' % t for t in ts]}  value {v:.8f}  "
      f"{2.0 * n * 576 * 3 * 1024 * 1024 / ms / 1e9:.0f} TFLOP/s algorithmic", flush=True)
