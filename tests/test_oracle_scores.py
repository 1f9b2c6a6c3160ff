"""A-score / C-score oracle restatements against fixtures produced by the reference's own code
(A_score/compute.py exec'd; C_score/utils/utils_correspondence.py imported) — CPU only."""
import os

import pytest
import torch

from oracle import a_score_oracle as ao
from oracle import c_score_oracle as co
from oracle import synth


@pytest.fixture(scope="module")
def a_gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "a_score.pt"))


@pytest.fixture(scope="module")
def c_gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "c_score.pt"))


@pytest.mark.parametrize("cname", list(synth.A_SCORE_CASES))
def test_a_score_oracle(cname, a_gold):
    feats = synth.a_score_features(synth.A_SCORE_CASES[cname])
    for sub, rec in a_gold[cname].items():
        res, s336, s224 = ao.a_score(feats[sub], feats["clip336"], feats["clip224"])
        assert abs(res - rec["result"]) < 1e-6
        assert max(abs(a - b) for a, b in zip(s336, rec["per_image_336"])) < 1e-6
        assert max(abs(a - b) for a, b in zip(s224, rec["per_image_224"])) < 1e-6


def test_a_pair_literal_equals_gemm_form():
    feats = synth.a_score_features(synth.A_SCORE_CASES["small"])
    for i in range(5):
        lit = ao.a_pair_literal(feats["dino"][i], feats["clip224"][i])
        assert abs(lit - ao.a_pair(feats["dino"][i], feats["clip224"][i])) < 1e-6


@pytest.mark.parametrize("cname", list(synth.C_SCORE_CASES))
def test_c_score_oracle(cname, c_gold):
    cfg = synth.C_SCORE_CASES[cname]
    d1, d2, kps = synth.c_score_pair(cfg)
    rec = c_gold[cname]
    P = cfg["P"]
    d1n, d2n = co.normalize_feats(d1), co.normalize_feats(d2)
    idx = co.kpts_to_patch_idx(kps, P)
    assert torch.equal(torch.as_tensor(idx), rec["patch_idx"])
    sim = torch.matmul(d1n, d2n.permute(0, 2, 1))[0]
    assert torch.equal(sim.argmax(-1), rec["argmax"])
    flow = co.get_flow(sim, cfg["window"], P)
    # relative tolerance in patch units (fp32 ulp at ~800 px is 6e-5)
    assert (flow - rec["flow"]).abs().max() < 2e-5 * P
    soft = co.keypoint_transformation(d1n, d2n, idx, P, soft=True, window=cfg["window"])
    hard = co.keypoint_transformation(d1n, d2n, idx, P, soft=False)
    assert torch.allclose(soft, rec["soft"], rtol=1e-5, atol=2e-4)
    assert torch.equal(hard, rec["hard"])


def test_linear_cka_properties():
    g = torch.Generator().manual_seed(0)
    X = torch.randn(500, 32, generator=g)
    Q, _ = torch.linalg.qr(torch.randn(32, 32, generator=g))
    assert abs(ao.linear_cka(X, X) - 1.0) < 1e-12
    assert abs(ao.linear_cka(X, 3.0 * X @ Q + 5.0) - 1.0) < 1e-9      # invariant to rotation, scale, shift
    assert ao.linear_cka(X, torch.randn(500, 32, generator=g)) < 0.2


@pytest.mark.parametrize("cname", list(synth.FLIP_CASES))
def test_adaptive_flip_oracle(cname, golden_dir):
    """Oracle restatement of the ADAPT_FLIP + MUTUAL_NN block against the reference's own functions (c_score_flip.pt)."""
    rec = torch.load(os.path.join(golden_dir, "c_score_flip.pt"))[cname]
    cfg = synth.FLIP_CASES[cname]
    d1, d1f, d2, kps1, kps2 = synth.flip_case(cfg)
    P = cfg["P"]
    n1, n1f, n2 = co.normalize_feats(d1), co.normalize_feats(d1f), co.normalize_feats(d2)
    vis = kps1[:, 2] * kps2[:, 2] > 0
    k1f = co.flip_keypoints(kps1, 840, co.permute_indices(cfg["permute"], vis))
    assert torch.equal(k1f, rec["kps1_flip"])
    idx_f = co.kpts_to_patch_idx(k1f, P)
    assert torch.equal(torch.as_tensor(idx_f), rec["idx_flip"])
    od, _ = co.mutual_nn_distance(n1, n2)
    fd, _ = co.mutual_nn_distance(n1f, n2)
    assert abs(od - rec["original_dist"]) < 2e-4 and abs(fd - rec["flip_dist"]) < 2e-4   # reference: fp32 matmul cdist
    kps = co.keypoint_transformation(n1, n2, co.kpts_to_patch_idx(kps1, P), P)
    kps_f = co.keypoint_transformation(n1f, n2, idx_f, P)
    final = co.optimized_kps_1_to_2(kps, kps_f, kps1, kps2, fd, od, vis, cfg["permute"])
    assert torch.allclose(final, rec["final"], rtol=1e-5, atol=2e-4)
    assert (fd < od) == cfg["from_flip"]
