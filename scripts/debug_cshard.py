import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from law_of_vision_representation_in_mllms_b200 import c_score as CS
dev = torch.device("cuda", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
torch.cuda.set_device(dev)
# dirty the caching allocator like the bench does
junk = torch.full((6, 1 << 30), 0x7f, dtype=torch.uint8, device=dev); del junk
junk = torch.randn(2, 1 << 28, device=dev) * 1e30; del junk
g = torch.Generator(device=dev).manual_seed(4242)
npair, P, K, D = 48, 60, 12, 1024
d1 = torch.randn(npair, P * P, D, generator=g, device=dev).to(torch.bfloat16)
d2 = (0.6 * torch.roll(d1.float(), 3, dims=1) + 0.8 * torch.randn(npair, P * P, D, generator=g, device=dev)).to(torch.bfloat16)
gc = torch.Generator().manual_seed(4243)
k1 = torch.cat([torch.rand(npair, K, 2, generator=gc) * 839.0, torch.ones(npair, K, 1)], -1)
idx = CS.kpts_to_patch_idx(k1.reshape(-1, 3), P).reshape(npair, -1)
whole = CS.keypoint_transformation(d1, d2, idx)
for rep in range(3):
    for lo, hi in ((0, 24), (24, 48), (0, 20), (20, 21), (21, 48)):
        junk = torch.randn(1 << 26, device=dev) * 1e30; del junk
        part = CS.keypoint_transformation(d1[lo:hi], d2[lo:hi], idx[lo:hi])
        diff = (part - whole[lo:hi]).abs()
        bad = (diff > 1e-3).nonzero()
        print(rep, lo, hi, "max diff", diff.max().item(), "n bad", len(bad), bad[:3].tolist() if len(bad) else "", flush=True)
w2 = CS.keypoint_transformation(d1, d2, idx)
print("whole repeat max diff", (w2 - whole).abs().max().item())
