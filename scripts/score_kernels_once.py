"""One pass over every score kernel at a moderate size, for ncu captures (`ncu -k regex:... python scripts/score_kernels_once.py`).
Sizes are large enough that each kernel runs many waves (inputs larger than L2) and small enough for ncu's replays."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from law_of_vision_representation_in_mllms_b200 import a_score as A, c_score as CS, cka as CK   # noqa: E402

which = set(sys.argv[1:]) or {"a", "cka", "dense", "kp"}
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
if "a" in which or "cka" in which:
    n = 1024
    o = torch.randn(n, 576, 1024, generator=g, device=dev).to(torch.bfloat16)
    r = (0.5 * o.float() + torch.randn(n, 576, 1024, generator=g, device=dev)).to(torch.bfloat16)
    if "a" in which:
        for _ in range(2):
            A.a_score_pairs(o, r)
    if "cka" in which:
        for _ in range(2):
            CK.linear_cka(o, r)
    del o, r
if "dense" in which or "kp" in which:
    n, P, K = 64, 60, 12
    d1 = torch.randn(n, P * P, 1024, generator=g, device=dev).to(torch.bfloat16)
    d2 = torch.randn(n, P * P, 1024, generator=g, device=dev).to(torch.bfloat16)
    idx = np.random.default_rng(0).integers(0, P * P, (n, K))
    for _ in range(2):
        if "dense" in which:
            CS.get_flow(d1[:16], d2[:16])
        if "kp" in which:
            CS.keypoint_transformation(d1, d2, idx)
torch.cuda.synchronize()
print("done")
