import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from law_of_vision_representation_in_mllms_b200.specs import TOWER_SPECS
from law_of_vision_representation_in_mllms_b200.weights import canonical_weights, synthetic_images, synthetic_state_dict
from oracle import vit_oracle
spec = TOWER_SPECS["openai/clip-vit-large-patch14-336"]
t0 = time.perf_counter(); cw = canonical_weights(spec, synthetic_state_dict(spec, seed=0, style="hf_init")); print("weights", time.perf_counter() - t0, flush=True)
for thr, n in ((128, 1), (64, 1), (32, 1), (16, 1), (32, 4), (64, 4)):
    torch.set_num_threads(thr)
    imgs = synthetic_images(n, 336, seed=1)
    t0 = time.perf_counter(); vit_oracle.tower_forward(cw, spec, imgs); dt = time.perf_counter() - t0
    print(f"threads {thr} images {n}: {dt:.1f} s -> {n/dt:.3f} img/s", flush=True)
