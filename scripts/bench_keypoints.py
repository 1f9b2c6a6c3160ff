"""C-score keypoint rows (kpsim_tcgen05_kernel + the windowed soft-argmax): 512 P=60 pairs of bf16 descriptors, K = 12,
median of 7 with CUDA events; prints us per pair and the bandwidth over the 7.4 MB per pair the path has to read."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from law_of_vision_representation_in_mllms_b200 import c_score as CS   # noqa: E402

dev = torch.device("cuda", 0)
P, D, K, n = 60, 1024, 12, 512
g = torch.Generator(device=dev).manual_seed(5)
d1 = torch.randn(n, P * P, D, generator=g, device=dev).bfloat16()
d2 = torch.randn(n, P * P, D, generator=g, device=dev).bfloat16()
idx = torch.randint(0, P * P, (n, K), generator=torch.Generator().manual_seed(6))
ts = []
for i in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = CS.keypoint_transformation(d1, d2, idx); e1.record(); torch.cuda.synchronize()
    if i >= 3:
        ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
print(f"keypoints {n} pairs P={P} D={D} K={K}: {ms:.3f} ms  {ms / n * 1e3:.3f} us/pair  "
      f"{(P * P + K) * D * 2.0 * n / ms / 1e6:.0f} GB/s  checksum {out.double().sum().item():.6f}", flush=True)
