"""One lovr_preprocess call on 256 device-resident 480x640 images (for ncu captures)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from law_of_vision_representation_in_mllms_b200 import preprocess as PP
rng = np.random.default_rng(0)
imgs = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(8)] * 32
host, offsets, hw = PP.pack_images(imgs)
dev = host.cuda()
cfg = PP.B200ImageProcessor({"shortest_edge": 336}, 336)._cfg()
out = torch.empty(len(imgs), 3, 336, 336, dtype=torch.bfloat16, device="cuda")
for _ in range(2):
    PP.run_packed(dev, offsets, hw, cfg, out=out)
torch.cuda.synchronize()
