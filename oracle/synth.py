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


# ---- image preprocessing: seeded 8-bit RGB images (smooth structure + texture + noise, so resampling is non-trivial)
PREPROCESS_IMAGE_SIZES = [(480, 640), (500, 333), (224, 224), (336, 500), (97, 61), (700, 1000)]   # (H, W)

# processor configurations as the checkpoints' preprocessor_config.json state them (size, crop, mean/std family, exact)
PREPROCESS_CASES = {
    "clip336": dict(kind="hf", size=336, crop=336, norm="clip", exact=False, images=[0]),
    "clip336_pad": dict(kind="hf", size=336, crop=336, norm="clip", exact=False, pad=True, images=[0, 1, 2, 4]),
    "clip224": dict(kind="hf", size=224, crop=224, norm="clip", exact=False, images=[1]),
    "clip224_upscale": dict(kind="hf", size=224, crop=224, norm="clip", exact=False, images=[4]),
    "dinov2_256_224": dict(kind="hf", size=256, crop=224, norm="imagenet", exact=False, images=[3]),
    "dinov2_336_override": dict(kind="hf", size=336, crop=336, norm="imagenet", exact=True, images=[0]),
    "siglip224": dict(kind="hf", size=224, crop=224, norm="half", exact=True, images=[5]),
    "cscore224": dict(kind="cscore", size=224, images=[0]),
    "cscore224_b": dict(kind="cscore", size=224, images=[5]),
}
PREPROCESS_NORMS = {
    "clip": ([0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]),
    "imagenet": ([0.485, 0.456, 0.406], [0.229, 0.224, 0.225]),
    "half": ([0.5, 0.5, 0.5], [0.5, 0.5, 0.5]),
}


def synthetic_rgb_image(h, w, seed):
    """-> HWC uint8 numpy array."""
    import numpy as np
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.empty((h, w, 3), np.float64)
    for c in range(3):
        fx, fy, ph = rng.uniform(0.01, 0.3), rng.uniform(0.01, 0.3), rng.uniform(0, 6.28)
        img[..., c] = 127 + 90 * np.sin(fx * xx + fy * yy + ph) + 40 * np.sign(np.sin(0.11 * (c + 1) * xx) * np.sin(0.07 * yy))
    img += rng.normal(0, 12, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def preprocess_images():
    return [synthetic_rgb_image(h, w, 4000 + i) for i, (h, w) in enumerate(PREPROCESS_IMAGE_SIZES)]


# ---- adaptive flip: (desc1, desc1 of the mirrored image, desc2, keypoints with left/right groups) per case
FLIP_CASES = {
    # desc2 is built from the MIRRORED source -> the flipped descriptors are closer: predictions of swapped groups are taken
    "p16_flip_wins": dict(seed=31, P=16, D=256, K=10, corr=0.7, from_flip=True,
                          permute=[0, [1, 2], 3, [4, 5], [6, 7, 8, 9]], invisible=(7,)),
    "p16_orig_wins": dict(seed=32, P=16, D=256, K=10, corr=0.7, from_flip=False,
                          permute=[0, [1, 2], 3, [4, 5], [6, 7, 8, 9]], invisible=()),
    "p24_flip_wins": dict(seed=33, P=24, D=1024, K=8, corr=0.5, from_flip=True,
                          permute=[[0, 1], [2, 3], 4, 5, [6, 7]], invisible=(2,)),
}


def flip_case(cfg):
    """-> desc1, desc1_flip, desc2 [1,P*P,D] un-normalised fp32; kps1, kps2 [K,3] (x, y, visible)."""
    g = torch.Generator().manual_seed(cfg["seed"])
    P, D, K = cfg["P"], cfg["D"], cfg["K"]
    grid = torch.randn(P, P, D, generator=g)
    d1 = grid.reshape(1, P * P, D)
    # the mirrored image's features: mirrored grid plus a perturbation (a tower is not exactly flip-equivariant)
    d1f = (torch.flip(grid, dims=(1,)) + 0.3 * torch.randn(P, P, D, generator=g)).reshape(1, P * P, D)
    src = d1f if cfg["from_flip"] else d1
    d2 = cfg["corr"] * src + (1 - cfg["corr"]) * torch.randn(1, P * P, D, generator=g)
    d1 = d1 * (0.5 + torch.rand(1, P * P, 1, generator=g))
    kps1 = torch.rand(K, 3, generator=g) * 839.0
    kps2 = torch.rand(K, 3, generator=g) * 839.0
    kps1[:, 2] = 1.0
    kps2[:, 2] = 1.0
    for i in cfg["invisible"]:
        kps2[i, 2] = 0.0
    return d1, d1f, d2, kps1, kps2


# ----------------------------------------------------------------------------------- mm_projector
PROJECTOR_CASES = {
    # name: (mm_projector_type, mm_hidden_size, hidden_size, tokens)
    "mlp2x_gelu_1024": ("mlp2x_gelu", 1024, 4096, 24),
    "mlp2x_gelu_2048": ("mlp2x_gelu", 2048, 4096, 16),
    "linear_1024": ("linear", 1024, 4096, 24),
    "mlp3x_gelu_768": ("mlp3x_gelu", 768, 1024, 24),
}


def projector_inputs(name):
    """Seeded fp32 weights (NOT rounded to 16 bits) for every Linear of the projector, in build order, and the input."""
    ptype, din, dh, tokens = PROJECTOR_CASES[name]
    depth = 1 if ptype == "linear" else int(ptype[3])
    g = torch.Generator().manual_seed(500 + sum(map(ord, name)))
    dims = [din] + [dh] * depth
    linears = [(torch.randn(dims[i + 1], dims[i], generator=g) * dims[i] ** -0.5, torch.randn(dims[i + 1], generator=g) * 0.1)
               for i in range(depth)]
    x = torch.randn(2, tokens, din, generator=g)
    return ptype, din, dh, linears, x
