import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from law_of_vision_representation_in_mllms_b200 import cka as CK
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
n = 2048
o = torch.randn(n * 576, 1024, generator=g, device=dev).to(torch.bfloat16)
r = torch.randn(n * 576, 1024, generator=g, device=dev).to(torch.bfloat16)
for dbg, pf in (("0", "12"), ("0", "0"), ("0", "4"), ("0", "32"), ("2", "0")):
    os.environ["LOVR_GRAM_DEBUG"] = dbg
    os.environ["LOVR_GRAM_PREFETCH"] = pf
    CK.cka_gram(o, r); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); CK.cka_gram(o, r); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"debug={dbg} prefetch={pf}: {ms:.3f} ms  executed {72 * 128 * 256 * 2.0 * n * 576 / ms / 1e9:.0f} TFLOP/s", flush=True)
