"""Dataset-level extraction on the device (extract.extract_folder): files in the reference's two layouts equal the
direct tower call on the same pixels, a re-run launches nothing, and the extractor's host tensors are complete on return."""
import os
import types

import numpy as np
import pytest
import torch
from PIL import Image

from law_of_vision_representation_in_mllms_b200 import _lib, extract as X, preprocess as PP, specs as S, towers
from law_of_vision_representation_in_mllms_b200.weights import synthetic_state_dict

pytestmark = pytest.mark.gpu


def _tower(key, image_size, layers=2):
    spec = S.tiny_spec(S.TOWER_SPECS[key], layers)
    sd = synthetic_state_dict(spec, seed=3, style="hf_init")
    args = types.SimpleNamespace(mm_vision_select_layer=-1, mm_vision_select_feature="patch", lovr_spec=spec,
                                 lovr_state_dict=sd, lovr_image_size=image_size, lovr_max_batch=4)
    cls = towers.DinoV2VisionTower if "dinov2" in key else towers.CLIPVisionTower
    return cls(key, args)


def _tree(root, n=5):
    rng = np.random.default_rng(0)
    paths = []
    for cls in ("bus", "dog"):
        os.makedirs(os.path.join(root, cls))
        for i in range(n):
            h, w = int(rng.integers(60, 200)), int(rng.integers(60, 200))
            p = os.path.join(root, cls, f"2008_{i:04d}.png")
            Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).save(p)
            paths.append(p)
    return paths


def test_extract_folder_c_score_layout_matches_direct_call_and_rerun_is_free(tmp_path):
    src, dst = str(tmp_path / "JPEGImages"), str(tmp_path / "features")
    paths = _tree(src)
    tower = _tower("facebook/dinov2-large", 336)
    ex = X.FeatureExtractor(tower, chunk=4)
    st = X.extract_folder(ex, src, dst, layout="c_score", model_tag="dino336", img_size=336, batch=6, barrier=False)
    assert (st["total"], st["written"], st["skipped"]) == (10, 10, 0) and st["launches"] > 0
    # the reference's per-image recipe: C-score pixels (extract_feature.py:65-67), batch-1 forward, permute + view (:86-90)
    for p in (paths[0], paths[7]):
        img = Image.open(p).convert("RGB")
        f = tower(PP.c_score_pixels([img], 336))
        want = f.permute(0, 2, 1).reshape(1, 1024, 24, 24).cpu()
        cls, stem = os.path.basename(os.path.dirname(p)), os.path.splitext(os.path.basename(p))[0]
        got = torch.load(os.path.join(dst, cls, f"{stem}_dino336.pt"))
        assert got.dtype == torch.float32 and tuple(got.shape) == (1, 1024, 24, 24)
        assert torch.equal(got, want)                       # batch position / chunking do not change a bit
    before = _lib.launch_count()
    st = X.extract_folder(ex, src, dst, layout="c_score", model_tag="dino336", img_size=336, batch=6, barrier=False)
    assert st["skipped"] == 10 and st["written"] == 0 and st["launches"] == 0 and _lib.launch_count() == before


def test_extract_folder_tokens_layout_and_host_tensor_complete_on_return(tmp_path):
    src, dst = str(tmp_path / "images"), str(tmp_path / "obelisc")
    paths = _tree(src, n=3)
    tower = _tower("openai/clip-vit-large-patch14", 224)
    ex = X.FeatureExtractor(tower, chunk=4)
    st = X.extract_folder(ex, src, dst, layout="tokens", batch=4, barrier=False)
    assert st["written"] == 6
    imgs = [Image.open(p).convert("RGB") for p in paths]
    # run_images returns a host tensor the caller may read IMMEDIATELY (no torch.cuda.synchronize() here: ADVICE r1 high)
    feats = ex.run_images(imgs, pixel_dtype=torch.bfloat16)
    snapshot = feats.clone()
    torch.cuda.synchronize()
    assert torch.equal(snapshot, feats)
    px = tower.image_processor.preprocess(imgs, return_tensors="pt", dtype=torch.bfloat16)["pixel_values"]
    assert torch.equal(feats, tower(px).cpu())
    got = torch.load(os.path.join(dst, "dog", "2008_0001.pt"))
    assert got.dtype == torch.bfloat16 and tuple(got.shape) == (256, 1024)
    assert torch.equal(got, feats[4])
