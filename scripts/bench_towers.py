"""Resident-input throughput of every tower family of SURVEY §8 rows a3-a5 (images/s, one GPU, bf16 operands) — for profiles/."""
import json, os, sys, types
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from law_of_vision_representation_in_mllms_b200 import specs as S, towers as T

CASES = [("openai/clip-vit-large-patch14-336", 336, 256), ("openai/clip-vit-large-patch14", 224, 512),
         ("laion/CLIP-ViT-L-14-laion2B-s32B-b82K", 224, 512), ("facebook/dinov2-large", 224, 512),
         ("facebook/dinov2-large", 336, 256), ("google/siglip-base-patch16-224", 224, 1024)]
FLOPS = {336: 366.03e9, 224: 155.29e9}
for name, R, n in CASES:
    cfg = types.SimpleNamespace(mm_vision_tower=name, mm_vision_select_layer=-2, mm_vision_select_feature="patch",
                                lovr_max_batch=n, lovr_image_size=R)
    tw = T.build_towers_from_selector(cfg)
    tw = tw[0] if isinstance(tw, torch.nn.ModuleList) else tw
    px = [torch.randn(n, 3, R, R, device="cuda").bfloat16() for _ in range(2)]
    out = torch.empty(n, tw.out_tokens(R), tw.hidden_size, dtype=torch.bfloat16, device="cuda")
    for i in range(3):
        tw.forward_into(px[i & 1], out, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(6):
        tw.forward_into(px[i & 1], out, 0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 6
    r = dict(tower=name, image_size=R, batch=n, tokens=out.shape[1], hidden=out.shape[2], ms_per_batch=round(ms, 3),
             images_per_s=round(n / ms * 1e3, 1))
    if tw.hidden_size == 1024:
        r["tflops"] = round(n / ms * 1e3 * FLOPS[R] / 1e12, 1)
    print(json.dumps(r), flush=True)
    del tw, px, out
    torch.cuda.empty_cache()
