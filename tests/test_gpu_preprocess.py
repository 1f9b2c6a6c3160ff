"""GPU: lovr_preprocess (through the C ABI, via the reference-shaped Python layer) against the oracle and the fixtures made
by the reference's own process_images.  Integer/byte work: the bar is BIT-EXACT float32 outputs."""
import hashlib
import os
import types

import numpy as np
import pytest
import torch

from law_of_vision_representation_in_mllms_b200 import preprocess as PP
from oracle import preprocess_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu


def _proc(cfg):
    mean, std = synth.PREPROCESS_NORMS[cfg["norm"]]
    size = {"height": cfg["size"], "width": cfg["size"]} if cfg["exact"] else {"shortest_edge": cfg["size"]}
    return PP.B200ImageProcessor(size, cfg["crop"], mean, std)


def _device_case(cfg, imgs):
    sel = [imgs[i] for i in cfg["images"]]
    if cfg["kind"] == "cscore":
        return PP.c_score_pixels(sel, cfg["size"])
    mcfg = types.SimpleNamespace(image_aspect_ratio="pad" if cfg.get("pad") else None)
    out = PP.process_images(sel, _proc(cfg), mcfg)
    return out[None] if out.dim() == 3 else out


def test_fixtures_from_reference_process_images(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "preprocess.pt"))
    imgs = synth.preprocess_images()
    for cname, cfg in synth.PREPROCESS_CASES.items():
        got = _device_case(cfg, imgs).float().cpu().contiguous()
        assert tuple(got.shape) == gold[cname]["shape"], cname
        assert hashlib.sha256(got.numpy().tobytes()).hexdigest() == gold[cname]["sha256"], cname


def test_random_sizes_bit_exact_vs_oracle():
    rng = np.random.default_rng(7)
    imgs, want = [], []
    mean, std = synth.PREPROCESS_NORMS["clip"]
    for i in range(24):
        h, w = (int(v) for v in rng.integers(20, 900, 2))
        imgs.append(synth.synthetic_rgb_image(h, w, 500 + i))
    # crop 70 / 45: not a multiple of 4 -> the one-column-per-thread vertical kernel; the others take the 4-column one
    for size, crop, exact, pad in ((64, 64, False, False), (96, 80, False, False), (112, 112, True, False),
                                   (72, 72, False, True), (90, 70, False, False), (45, 45, True, True)):
        proc = PP.B200ImageProcessor({"height": size, "width": size} if exact else {"shortest_edge": size}, crop, mean, std)
        got = proc.preprocess(imgs, return_tensors="pt", pad_square=pad)["pixel_values"].cpu().numpy()
        for k, a in enumerate(imgs):
            ref = O.hf_preprocess(a, size, crop, mean, std, exact=exact, pad_square=pad)
            assert np.array_equal(got[k], ref), (size, crop, exact, pad, a.shape)


def test_edge_shapes_and_dtypes():
    mean, std = synth.PREPROCESS_NORMS["imagenet"]
    proc = PP.B200ImageProcessor({"shortest_edge": 32}, 32, mean, std)
    cases = [synth.synthetic_rgb_image(32, 32, 1),        # identity resize
             synth.synthetic_rgb_image(1, 1, 2),          # single pixel, 32x upscale
             synth.synthetic_rgb_image(3, 200, 3),        # extreme aspect ratio
             synth.synthetic_rgb_image(1200, 40, 4),      # 37x downscale on one axis (75 taps)
             np.zeros((50, 70, 3), np.uint8), np.full((70, 50, 3), 255, np.uint8)]
    got = proc.preprocess(cases)["pixel_values"].cpu().numpy()
    for k, a in enumerate(cases):
        assert np.array_equal(got[k], O.hf_preprocess(a, 32, 32, mean, std)), a.shape
    # 16-bit outputs are the float32 values rounded once (what `.to(dtype)` does in the reference)
    for dt in (torch.bfloat16, torch.float16):
        lo = proc.preprocess(cases, dtype=dt)["pixel_values"].cpu()
        assert torch.equal(lo, torch.from_numpy(got).to(dt))
    # c-score map
    cs = PP.c_score_pixels(cases[:3], 48).cpu().numpy()
    for k in range(3):
        assert np.array_equal(cs[k], O.c_score_pixels(cases[k], 48))


def test_errors():
    from law_of_vision_representation_in_mllms_b200 import _lib
    proc = PP.B200ImageProcessor({"shortest_edge": 32}, 64)
    with pytest.raises(_lib.LovrError):
        proc.preprocess([synth.synthetic_rgb_image(40, 50, 1)])
    with pytest.raises(ValueError):
        PP.B200ImageProcessor({"shortest_edge": 32}, 32).preprocess([np.zeros((4, 4), np.uint8)])


def test_extractor_from_raw_images_equals_preprocess_then_tower():
    """FeatureExtractor.run_images (8-bit images in, GPU preprocessing per tower, features out) == preprocessing with each
    tower's processor and calling the towers, bit for bit — single tower and the clip336.dinov2 dual tower."""
    from law_of_vision_representation_in_mllms_b200 import extract as X
    from law_of_vision_representation_in_mllms_b200 import specs as S
    rng = np.random.default_rng(3)
    imgs = [synth.synthetic_rgb_image(int(h), int(w), 700 + i) for i, (h, w) in enumerate(rng.integers(150, 600, (11, 2)))]
    for selector in ("openai/clip-vit-large-patch14-336", "openai/clip-vit-large-patch14-336.facebook/dinov2-large"):
        ex = X.build_extractor(selector, chunk=4, lovr_max_batch=4, lovr_image_size=336,
                               lovr_spec=None)
        got = ex.run_images(imgs)
        torch.cuda.synchronize()
        parts = []
        for t in ex.towers:
            px = t.image_processor.preprocess(imgs, dtype=torch.bfloat16)["pixel_values"]
            parts.append(torch.cat([t(px[s:s + 4]) for s in range(0, len(imgs), 4)]))
        want = torch.cat(parts, dim=-1).cpu()
        assert got.shape == want.shape and torch.equal(got, want), selector
