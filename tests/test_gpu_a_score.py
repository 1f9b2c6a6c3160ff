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


@pytest.mark.parametrize("n,x,y,D", [(3, 576, 576, 1024), (3, 576, 256, 1024), (2, 256, 576, 4096), (5, 130, 700, 512),
                                     (4, 1, 1, 320), (3, 577, 129, 384), (2, 100, 1500, 1024)])
def test_a_pairs_bf16_single_pass_kernel(n, x, y, D):
    """bf16 features with D > 256 take the one-pass kernel (norms from the GEMM's own shared-memory tiles, ascore.cuh):
    ragged tiles on both sides, every BN instance (y = 576 -> 192, 256 -> 128/256, 700 -> 128 ...), zero rows."""
    g = torch.Generator().manual_seed(n * 1000 + x + y)
    o = (torch.randn(n, x, D, generator=g) * (0.2 + torch.rand(n, x, 1, generator=g))).bfloat16()
    r = (torch.randn(n, y, D, generator=g) + 0.3 * torch.randn(n, 1, D, generator=g)).bfloat16()
    if x > 5:
        o[0, 5] = 0
    got = A.a_score_pairs(o.cuda(), r.cuda()).cpu()
    want = torch.tensor([ao.a_pair(o[i].float(), r[i].float()) for i in range(n)])
    assert (got.double() - want.double()).abs().max() < TOL


def test_a_pairs_bf16_single_pass_many_pairs_persistent_loop():
    """More pairs than SMs: every CTA walks several pairs (scale buffers alternate by pair parity); bitwise repeatable and
    independent of how the pairs are batched."""
    g = torch.Generator().manual_seed(77)
    n, x, y, D = 700, 150, 200, 320
    o = torch.randn(n, x, D, generator=g).bfloat16().cuda()
    r = (0.5 * o[:, :1].float().cpu() + torch.randn(n, y, D, generator=g)).bfloat16().cuda()
    a1, a2 = A.a_score_pairs(o, r), A.a_score_pairs(o, r)
    assert torch.equal(a1, a2)
    assert torch.equal(a1[100:433], A.a_score_pairs(o[100:433], r[100:433]))
    idx = [0, 147, 148, 149, 699]
    want = torch.tensor([ao.a_pair(o[i].float().cpu(), r[i].float().cpu()) for i in idx])
    assert (a1[idx].cpu().double() - want.double()).abs().max() < TOL
