"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the oracle cannot run them in
seconds) plus oracle spot checks on sampled units:

  C2  tower over a large batch     per-image independence of batch composition (bit-exact), oracle on one sampled image
  C3  A score, 10k x 576 x 1024    self-similarity = 1, invariance to reference-token permutation and to per-token
                                   power-of-two scaling (bit-exact), oracle on sampled pairs
  C3  linear CKA, 5.76M rows       CKA(X, XQ) = 1 for orthogonal Q, shard additivity
  C5  C score, P = 60              identity pair -> identity correspondence; rolled grid -> rolled argmax
  preprocessing, 1000 images       oracle on sampled images, identity resize = value map only
"""
import types

import numpy as np
import pytest
import torch

from law_of_vision_representation_in_mllms_b200 import a_score as A
from law_of_vision_representation_in_mllms_b200 import c_score as CS
from law_of_vision_representation_in_mllms_b200 import cka as CK
from law_of_vision_representation_in_mllms_b200 import preprocess as PP
from law_of_vision_representation_in_mllms_b200 import specs as S
from law_of_vision_representation_in_mllms_b200 import towers as T
from oracle import a_score_oracle as ao
from oracle import preprocess_oracle as po
from oracle import synth

pytestmark = pytest.mark.gpu


def test_a_score_c3_size_properties():
    n, P, D = 10000, 576, 1024
    g = torch.Generator(device="cuda").manual_seed(11)
    o = torch.randn(n, P, D, generator=g, device="cuda").to(torch.bfloat16)
    r = torch.randn(n, P, D, generator=g, device="cuda").to(torch.bfloat16)
    r[:, : P // 2] += o[:, P // 2:]                     # correlated half: the maximum is not degenerate
    base = A.a_score_pairs(o, r)
    assert base.shape == (n,) and torch.isfinite(base).all()
    # oracle on sampled pairs (fp32 restatement of compute.py), 1e-5
    for i in (0, 4999, 9999):
        assert abs(float(base[i]) - ao.a_pair(o[i].float().cpu(), r[i].float().cpu())) < 1e-5
    # a(o, o) = 1
    self_sim = A.a_score_pairs(o[:2000], o[:2000])
    assert (self_sim - 1).abs().max() < 1e-5
    # the row maximum does not depend on the order of the reference tokens: bit-exact
    perm = torch.randperm(P, device="cuda")
    assert torch.equal(A.a_score_pairs(o[:2000], r[:2000][:, perm]), base[:2000])
    # cosine ignores positive per-token scales; powers of two keep every product exact
    scale = torch.tensor([0.25, 0.5, 2.0, 4.0], device="cuda")[torch.randint(0, 4, (2000, P, 1), device="cuda")]
    scaled = A.a_score_pairs((o[:2000].float() * scale).to(torch.bfloat16), r[:2000])
    assert (scaled - base[:2000]).abs().max() < 2e-6


def test_linear_cka_c3_size_properties():
    rows, D = 10000 * 576, 1024
    g = torch.Generator(device="cuda").manual_seed(3)
    X = torch.randn(rows, D, generator=g, device="cuda").to(torch.bfloat16)
    Q, _ = torch.linalg.qr(torch.randn(D, D, generator=g, device="cuda"))
    Y = torch.empty_like(X)
    for s in range(0, rows, 1 << 20):                   # Y = 1.5 X Q + shift, slab-wise
        Y[s:s + (1 << 20)] = (1.5 * (X[s:s + (1 << 20)].float() @ Q) + 0.25).to(torch.bfloat16)
    whole = CK.linear_cka(X, Y)
    assert abs(whole - 1.0) < 1e-4                       # bf16 rounding of Y is the only deviation from an exact rotation
    assert abs(CK.linear_cka(X, X) - 1.0) < 1e-5
    # shards add up: two halves accumulated into one Gram == one pass
    gram, n1, dx, dy = CK.cka_gram(X[: rows // 2], Y[: rows // 2])
    gram, n2, _, _ = CK.cka_gram(X[rows // 2:], Y[rows // 2:], gram=gram)
    # run boundaries differ between the sharded and the unsharded job and runs are summed in fp32 registers (gram.cuh):
    # ~1e-8 on CKA (measured 3e-8), three orders below the 1e-5 tolerance
    assert abs(CK.cka_from_gram(gram, n1 + n2, dx, dy) - whole) < 1e-6
    Z = torch.randn(rows // 8, D, generator=g, device="cuda").to(torch.bfloat16)
    assert CK.linear_cka(X[: rows // 8], Z) < 0.01       # independent features


def test_c_score_c5_shape_properties():
    P, D, B = 60, 1024, 6
    g = torch.Generator(device="cuda").manual_seed(7)
    d1 = torch.randn(B, P * P, D, generator=g, device="cuda")
    flow, amax = CS.get_flow(d1, d1, window=5, return_argmax=True)
    ident = torch.arange(P * P, device="cuda")
    assert torch.equal(amax.long(), ident.expand(B, -1))
    grid = torch.stack([ident % P, ident // P], dim=-1).float()
    assert (flow - grid).abs().max() < 1e-3              # soft-argmax of a one-hot-like row: the patch itself
    rolled = torch.roll(d1.view(B, P, P, D), shifts=(3, -2), dims=(1, 2)).reshape(B, P * P, D)
    _, amax2 = CS.get_flow(d1, rolled, window=5, return_argmax=True)
    ty, tx = (ident // P + 3) % P, (ident % P - 2) % P
    assert torch.equal(amax2.long(), (ty * P + tx).expand(B, -1))
    dist = CS.get_distance_mutual_nn(d1, rolled)          # every patch has its exact copy: distance 0, all mutual
    assert dist.abs().max() < 5e-3                       # sqrt of a 1e-6-accurate squared distance


def test_tower_large_batch_independence():
    spec = S.TOWER_SPECS["openai/clip-vit-large-patch14-336"]
    args = types.SimpleNamespace(mm_vision_select_layer=-2, mm_vision_select_feature="patch", lovr_max_batch=128)
    tower = T.CLIPVisionTower("openai/clip-vit-large-patch14-336", args)
    g = torch.Generator(device="cuda").manual_seed(1234)
    px = torch.randn(1000, 3, 336, 336, generator=g, device="cuda").to(torch.bfloat16)
    out = tower(px)
    assert out.shape == (1000, 576, 1024) and torch.isfinite(out.float()).all()
    # an image's features do not depend on which batch, chunk or position it was computed in
    idx = torch.tensor([999, 0, 517, 128, 127, 640], device="cuda")
    again = tower(px[idx])
    assert torch.equal(again, out[idx])
    assert torch.equal(tower(px[517:518]), out[517:518])


def test_preprocess_large_batch():
    rng = np.random.default_rng(11)
    sizes = [(int(h), int(w)) for h, w in rng.integers(200, 720, (1000, 2))]
    base = [synth.synthetic_rgb_image(h, w, 9000 + i) for i, (h, w) in enumerate(sizes[:40])]
    imgs = []                                             # 1000 crops of 40 base images: up- and down-scaling cases
    for i in range(1000):
        b = base[i % 40]
        imgs.append(b[: 64 + sizes[i][0] % (b.shape[0] - 64), : 64 + sizes[i][1] % (b.shape[1] - 64)])
    proc = PP.processor_for("openai/clip-vit-large-patch14-336")
    out = proc.preprocess(imgs)["pixel_values"]
    assert out.shape == (1000, 3, 336, 336)
    for i in (0, 333, 999):
        want = po.hf_preprocess(np.ascontiguousarray(imgs[i]), 336, 336, proc.image_mean, proc.image_std)
        assert np.array_equal(out[i].cpu().numpy(), want), i
    # resize to the size the image already has = the value map alone
    sq = [synth.synthetic_rgb_image(224, 224, 5)]
    got = PP.c_score_pixels(sq, 224)[0].cpu()
    want = (torch.from_numpy(sq[0]).permute(2, 0, 1).float() / 255.0 - 0.5) * 2
    assert torch.equal(got, want)
