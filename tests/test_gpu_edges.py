"""Edge cases of the hot path (reference tests none; these pin the behaviours §8b lists): tiny / ragged shapes, size
limits, error codes surfaced as the reference's exception types, determinism."""
import types

import pytest
import torch

from law_of_vision_representation_in_mllms_b200 import _lib, a_score as A, c_score as CS, specs as S, towers
from law_of_vision_representation_in_mllms_b200.weights import synthetic_images
from oracle import a_score_oracle as ao
from oracle import c_score_oracle as co

pytestmark = pytest.mark.gpu


def test_a_score_degenerate_shapes():
    g = torch.Generator().manual_seed(1)
    for x, y, D in [(1, 1, 64), (1, 700, 64), (129, 1, 128), (577, 577, 64)]:
        o, r = torch.randn(2, x, D, generator=g), torch.randn(2, y, D, generator=g)
        got = A.a_score_pairs(o.cuda(), r.cuda()).cpu()
        want = torch.tensor([ao.a_pair(o[i], r[i]) for i in range(2)])
        assert (got.double() - want.double()).abs().max() < 1e-5, (x, y, D)


def test_a_score_many_pairs_chunked_and_deterministic():
    g = torch.Generator().manual_seed(2)
    o = torch.randn(5000, 16, 64, generator=g).bfloat16().cuda()      # > _MAX_PAIRS_PER_CALL: two native calls
    r = torch.randn(5000, 24, 64, generator=g).bfloat16().cuda()
    a1, a2 = A.a_score_pairs(o, r), A.a_score_pairs(o, r)
    assert torch.equal(a1, a2)                                          # fixed-order reductions: bitwise repeatable
    idx = [0, 4095, 4096, 4999]
    want = torch.tensor([ao.a_pair(o[i].float().cpu(), r[i].float().cpu()) for i in idx])
    assert (a1[idx].cpu().double() - want.double()).abs().max() < 1e-5


def test_a_score_rejects_cpu_and_bad_shapes():
    with pytest.raises(_lib.LovrError):
        A.a_score_pairs(torch.zeros(1, 4, 64), torch.zeros(1, 4, 64))
    with pytest.raises(ValueError):
        A.a_score_pairs(torch.zeros(1, 4, 64).cuda(), torch.zeros(2, 4, 64).cuda())


def test_c_score_small_grid_window_larger_than_grid_and_index_errors():
    g = torch.Generator().manual_seed(3)
    P, D = 4, 64                                                         # window 5 covers the whole 4x4 grid: no zero tail
    d1, d2 = torch.randn(1, P * P, D, generator=g), torch.randn(1, P * P, D, generator=g)
    idx = [[0, 5, 15]]
    got = CS.keypoint_transformation(d1.cuda(), d2.cuda(), idx, soft=True, window=5)[0].cpu()
    want = co.keypoint_transformation(co.normalize_feats(d1), co.normalize_feats(d2), idx[0], P, soft=True, window=5)
    assert (got - want).abs().max() < 1e-2
    with pytest.raises(IndexError):
        CS.keypoint_transformation(d1.cuda(), d2.cuda(), [[0, 16]])
    with pytest.raises(ValueError):
        CS.get_flow(torch.zeros(1, 15, 64).cuda(), torch.zeros(1, 15, 64).cuda())   # 15 patches is not a square grid


def test_tower_limits_and_single_image():
    spec = S.tiny_spec(S.TOWER_SPECS["facebook/dinov2-large"], 2)
    args = types.SimpleNamespace(mm_vision_select_layer=-1, mm_vision_select_feature="cls_patch", lovr_spec=spec,
                                 lovr_max_batch=2, lovr_image_size=224)
    tower = towers.DinoV2VisionTower("facebook/dinov2-large", args)
    one = tower(synthetic_images(1, 224).cuda())
    assert tuple(one.shape) == (1, 257, 1024) and torch.isfinite(one).all()
    three = tower(synthetic_images(3, 224).cuda())                      # 3 > max_batch=2: chunked, same first image
    assert torch.equal(three[:1], one)
    with pytest.raises(_lib.LovrError, match="tokens per image"):      # native 518px = 1370 tokens > 640
        tower(synthetic_images(1, 518).cuda())
    with pytest.raises(ValueError):
        tower(torch.zeros(1, 1, 224, 224).cuda())                       # not RGB


def test_selector_grammar_matches_reference():
    cfg = types.SimpleNamespace(mm_vision_tower="no/such-tower", mm_vision_select_layer=-2)
    with pytest.raises(KeyError):
        towers.build_towers_from_selector(cfg)                          # llava_arch.py:53 dict lookup
    cfg = types.SimpleNamespace(mm_vision_tower="runwayml/stable-diffusion-v1-5", mm_vision_select_layer=-2)
    with pytest.raises(NotImplementedError):
        towers.build_towers_from_selector(cfg)                          # diffusion towers: out of scope
    assert towers.build_towers_from_selector(types.SimpleNamespace(mm_vision_tower="some_feature")) == "feature"
    with pytest.raises(ValueError, match="Unknown vision tower"):       # builder.py:14
        towers.build_vision_tower(types.SimpleNamespace(mm_vision_tower="facebook/dinov2-large", mm_vision_select_layer=-2))
