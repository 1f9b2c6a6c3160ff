"""Seeded synthetic inputs shared by oracle/make_golden.py (which feeds them to the reference) and
tests (which feed the same tensors to the CUDA path).  Test infrastructure only."""
import torch


def _bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


# ---- A score: 100 (clip336, clip224, other) triples per case, like A_score/compute.py:18-28 expects
A_SCORE_CASES = {
    # rectangular x != y everywhere, tiny: the whole literal reference runs in seconds
    "small": dict(seed=11, n=100, D=64, T336=49, T224=25, Tother=36, big=0, others=("dino", "openclip")),
    # images 1..3 at the real shapes (576/256 tokens, D=1024), the other 97 tiny
    "mixed": dict(seed=12, n=100, D=32, T336=9, T224=4, Tother=9, big=3, bigD=1024, bigT336=576, bigT224=256,
                  bigTother=576, others=("dino",)),
}


def a_score_features(cfg):
    """-> {'clip336': [..100 tensors [T,D]], 'clip224': [...], <other>: [...]}, fp32 holding
    bf16-representable values (so a bf16 device copy is exact)."""
    g = torch.Generator().manual_seed(cfg["seed"])
    out = {"clip336": [], "clip224": []}
    for o in cfg["others"]:
        out[o] = []
    for i in range(cfg["n"]):
        big = i < cfg["big"]
        D = cfg["bigD"] if big else cfg["D"]
        t336 = cfg["bigT336"] if big else cfg["T336"]
        t224 = cfg["bigT224"] if big else cfg["T224"]
        to = cfg["bigTother"] if big else cfg["Tother"]
        r336 = torch.randn(t336, D, generator=g)
        mix = torch.randn(t224, t336, generator=g).softmax(dim=-1)
        r224 = mix @ r336 + 0.5 * torch.randn(t224, D, generator=g)
        out["clip336"].append(_bf16r(r336))
        out["clip224"].append(_bf16r(r224))
        for k, o in enumerate(cfg["others"]):
            sel = torch.randint(0, t336, (to,), generator=g)
            other = (0.3 + 0.3 * k) * r336[sel] + torch.randn(to, D, generator=g)
            other = other * (0.1 + 3.0 * torch.rand(to, 1, generator=g))     # norms vary per token
            out[o].append(_bf16r(other))
    return out


# ---- C score: one (desc1, desc2, keypoints) pair per case
C_SCORE_CASES = {
    "p16_rand": dict(seed=21, P=16, D=1024, K=12, window=5, corr=0.0),
    "p16_corr": dict(seed=22, P=16, D=1024, K=12, window=5, corr=0.7),
    "p24_corr": dict(seed=23, P=24, D=1024, K=12, window=5, corr=0.5),
    "p14_d768": dict(seed=24, P=14, D=768, K=9, window=5, corr=0.4),
    "p60_rand": dict(seed=25, P=60, D=1024, K=12, window=5, corr=0.0),
    "p60_corr": dict(seed=26, P=60, D=1024, K=30, window=5, corr=0.6),
    "p24_two": dict(seed=27, P=24, D=2048, K=12, window=5, corr=0.5),
}


def c_score_pair(cfg):
    """-> desc1 [1,P*P,D], desc2 [1,P*P,D] (un-normalised fp32), kps [K,3] (x, y, visible) in [0,840)."""
    g = torch.Generator().manual_seed(cfg["seed"])
    P, D, K = cfg["P"], cfg["D"], cfg["K"]
    d1 = torch.randn(1, P * P, D, generator=g)
    # desc2 = spatially shifted copy of desc1 (+noise): a genuine dense correspondence
    grid = d1.view(P, P, D)
    shifted = torch.roll(grid, shifts=(2, -3), dims=(0, 1)).reshape(1, P * P, D)
    d2 = cfg["corr"] * shifted + (1 - cfg["corr"]) * torch.randn(1, P * P, D, generator=g)
    d1 = d1 * (0.5 + torch.rand(1, P * P, 1, generator=g))
    kps = torch.rand(K, 3, generator=g) * 839.0
    kps[:, 2] = 1.0
    return d1, d2, kps
