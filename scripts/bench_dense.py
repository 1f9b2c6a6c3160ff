"""Dense C-score flow at P = 60, D = 1024 (bf16): microseconds per pair, CUDA events."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from law_of_vision_representation_in_mllms_b200 import c_score as CS
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
n, P = 64, 60
d1 = torch.randn(n, P * P, 1024, generator=g, device=dev).to(torch.bfloat16)
d2 = torch.randn(n, P * P, 1024, generator=g, device=dev).to(torch.bfloat16)
CS.get_flow(d1, d2); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    CS.get_flow(d1, d2)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("LOVR_CSIM_2CTA", os.environ.get("LOVR_CSIM_2CTA", "1"), "dense flow us per pair", 1e3 * ms / n,
      "TFLOP/s", 2.0 * P ** 4 * 1024 * n / ms / 1e9)
