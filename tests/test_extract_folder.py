"""Host logic of the dataset-level extraction driver (extract.extract_folder) on CPU with a stub extractor: the walk, the
contiguous rank shards, the two file layouts of the reference (C_score/extract_feature.py:110-130,
llava/feature/extract.py:222-231), skip-if-exists (:225) and that a finished job loads and runs nothing."""
import os
import types

import numpy as np
import torch
from PIL import Image

from law_of_vision_representation_in_mllms_b200 import extract as X


class _StubTower:
    hidden_size = 8

    def _default_image_size(self):
        return 28


class _StubExtractor:
    """Stands for FeatureExtractor: features = f(first pixel of the image), 4 tokens, records every call."""

    def __init__(self):
        self.towers = [_StubTower()]
        self.calls = []

    def run_images(self, images, pixel_dtype=torch.bfloat16, pre_cfgs=None, host_out=None, **kw):
        self.calls.append((len(images), pixel_dtype, None if pre_cfgs is None else int(pre_cfgs[0].crop_size)))
        out = torch.zeros(len(images), 4, 8, dtype=pixel_dtype)
        for i, im in enumerate(images):
            out[i] = float(np.asarray(im)[0, 0, 0]) + torch.arange(32, dtype=torch.float32).reshape(4, 8).to(pixel_dtype)
        return out


def _make_tree(root, n_per_class=3):
    k = 0
    for cls in ("aeroplane", "cat"):
        os.makedirs(os.path.join(root, cls))
        for i in range(n_per_class):
            Image.fromarray(np.full((5, 7, 3), k, dtype=np.uint8)).save(os.path.join(root, cls, f"img{i}.png"))
            k += 1
    open(os.path.join(root, "cat", "notes.txt"), "w").write("not an image")


def test_c_score_layout_shards_and_skip(tmp_path):
    src, dst = str(tmp_path / "JPEGImages"), str(tmp_path / "features")
    _make_tree(src)
    items = X.list_images(src)
    assert [(c, s) for c, s, _ in items] == [("aeroplane", "img0"), ("aeroplane", "img1"), ("aeroplane", "img2"),
                                            ("cat", "img0"), ("cat", "img1"), ("cat", "img2")]
    ex = _StubExtractor()
    st0 = X.extract_folder(ex, src, dst, layout="c_score", model_tag="dino336", rank=0, world=2, batch=2, barrier=False)
    assert (st0["total"], st0["shard"], st0["skipped"], st0["written"]) == (6, 3, 0, 3)
    assert ex.calls == [(2, torch.float32, 28), (1, torch.float32, 28)]                  # rank 0: first half, batches of 2
    assert sorted(os.listdir(os.path.join(dst, "aeroplane"))) == [f"img{i}_dino336.pt" for i in range(3)]
    assert not os.path.exists(os.path.join(dst, "cat"))
    f = torch.load(os.path.join(dst, "aeroplane", "img1_dino336.pt"))
    assert tuple(f.shape) == (1, 8, 2, 2) and f.dtype == torch.float32                   # [1, C, h, w]
    want = (1.0 + torch.arange(32, dtype=torch.float32).reshape(1, 4, 8)).permute(0, 2, 1).reshape(1, 8, 2, 2)
    assert torch.equal(f, want)
    st1 = X.extract_folder(ex, src, dst, layout="c_score", model_tag="dino336", rank=1, world=2, batch=8, barrier=False)
    assert st1["written"] == 3 and sorted(os.listdir(os.path.join(dst, "cat"))) == [f"img{i}_dino336.pt" for i in range(3)]
    # finished job: nothing is loaded, nothing runs
    n_calls = len(ex.calls)
    loads = []
    st = X.extract_folder(ex, src, dst, layout="c_score", model_tag="dino336", rank=0, world=1, barrier=False,
                          loader=lambda p: loads.append(p))
    assert st["skipped"] == 6 and st["written"] == 0 and len(ex.calls) == n_calls and loads == []
    # one file removed: exactly that image is redone
    os.remove(os.path.join(dst, "cat", "img1_dino336.pt"))
    st = X.extract_folder(ex, src, dst, layout="c_score", model_tag="dino336", rank=0, world=1, barrier=False)
    assert st["skipped"] == 5 and st["written"] == 1 and ex.calls[-1][0] == 1
    assert not [f for _, _, fs in os.walk(dst) for f in fs if ".tmp" in f]                # no temporaries left behind


def test_tokens_layout(tmp_path):
    src, dst = str(tmp_path / "images"), str(tmp_path / "obelisc")
    _make_tree(src, n_per_class=2)
    ex = _StubExtractor()
    st = X.extract_folder(ex, src, dst, layout="tokens", rank=0, world=1, barrier=False)
    assert st["written"] == 4 and ex.calls == [(4, torch.bfloat16, None)]                 # the tower's own processor, bf16
    f = torch.load(os.path.join(dst, "cat", "img1.pt"))                                   # <out>/<parent dir>/<stem>.pt
    assert tuple(f.shape) == (4, 8) and f.dtype == torch.bfloat16                         # out_list[0].squeeze().cpu()
