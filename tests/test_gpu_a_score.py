"""A score on the device against (a) fixtures produced by exec'ing the reference's A_score/compute.py and
(b) the CPU oracle, tolerance 1e-5 (north_star)."""
import os

import pytest
import torch

from law_of_vision_representation_in_mllms_b200 import a_score as A
from oracle import a_score_oracle as ao
from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def a_gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "a_score.pt"))


@pytest.mark.parametrize("cname", list(synth.A_SCORE_CASES))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_a_score_matches_reference_script(cname, dtype, a_gold):
    feats = synth.a_score_features(synth.A_SCORE_CASES[cname])      # bf16-representable values
    cast = lambda ts: [t.to(dtype) for t in ts]
    for sub, rec in a_gold[cname].items():
        res, s336, s224 = A.a_score(cast(feats[sub]), cast(feats["clip336"]), cast(feats["clip224"]))
        assert abs(res - rec["result"]) < TOL
        assert max(abs(a - b) for a, b in zip(s336, rec["per_image_336"])) < TOL
        assert max(abs(a - b) for a, b in zip(s224, rec["per_image_224"])) < TOL


@pytest.mark.parametrize("x,y,D", [(576, 576, 1024), (576, 256, 1024), (256, 576, 4096), (196, 576, 768), (37, 301, 200)])
def test_a_pairs_fp32_general_shapes(x, y, D):
    g = torch.Generator().manual_seed(x * 7 + y)
    o = torch.randn(3, x, D, generator=g) * torch.rand(3, x, 1, generator=g)
    r = torch.randn(3, y, D, generator=g) + 0.2 * torch.randn(3, 1, D, generator=g)
    got = A.a_score_pairs(o.cuda(), r.cuda()).cpu()
    want = torch.tensor([ao.a_pair(o[i], r[i]) for i in range(3)])
    assert (got.double() - want.double()).abs().max() < TOL


def test_zero_rows_and_self_similarity():
    g = torch.Generator().manual_seed(3)
    o = torch.randn(2, 64, 128, generator=g)
    o[0, 5] = 0                                              # a zero token: cosine 0 like the reference (eps paths)
    got = A.a_score_pairs(o.cuda(), o.cuda()).cpu()
    want = torch.tensor([ao.a_pair(o[i], o[i]) for i in range(2)])
    assert (got.double() - want.double()).abs().max() < TOL
    assert abs(got[1].item() - 1.0) < TOL


def test_folder_driver_mirrors_compute_py(tmp_path, a_gold):
    cfg = synth.A_SCORE_CASES["small"]
    feats = synth.a_score_features(cfg)
    for sub, ts in feats.items():
        os.makedirs(tmp_path / sub)
        for i, t in enumerate(ts):
            torch.save(t, tmp_path / sub / f"tensor_{i + 1}.pt")
    res = A.compute_a_scores(str(tmp_path), ["dino", "missing_rep", "openclip"], verbose=False)
    assert set(res) == {"dino", "openclip"}                 # load failure -> skipped, like compute.py:44-46
    for k, v in res.items():
        assert abs(v - a_gold["small"][k]["result"]) < TOL
