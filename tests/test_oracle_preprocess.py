"""CPU: the numpy restatement of Pillow's 8-bit bicubic resampling + the HF / C-score value maps (oracle/preprocess_oracle.py)
against (a) Pillow itself, (b) the committed fixtures produced by the reference's own `process_images` (oracle/make_golden.py),
and the host side of the preprocessing boundary (argument validation, workspace query) — none of which needs a GPU."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest
import torch

from law_of_vision_representation_in_mllms_b200 import _lib
from oracle import preprocess_oracle as O
from oracle import synth


def oracle_case(cfg, imgs):
    sel = [imgs[i] for i in cfg["images"]]
    if cfg["kind"] == "cscore":
        return np.stack([O.c_score_pixels(a, cfg["size"]) for a in sel])
    mean, std = synth.PREPROCESS_NORMS[cfg["norm"]]
    return np.stack([O.hf_preprocess(a, cfg["size"], cfg["crop"], mean, std, exact=cfg["exact"],
                                     pad_square=bool(cfg.get("pad"))) for a in sel])


def test_oracle_matches_reference_fixtures(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "preprocess.pt"))
    imgs = synth.preprocess_images()
    for cname, cfg in synth.PREPROCESS_CASES.items():
        got = oracle_case(cfg, imgs)
        assert tuple(got.shape) == gold[cname]["shape"], cname
        assert hashlib.sha256(np.ascontiguousarray(got, np.float32).tobytes()).hexdigest() == gold[cname]["sha256"], cname
        assert np.array_equal(got[0, :, ::37, ::41], gold[cname]["sample"].numpy()), cname


def test_oracle_matches_pillow_bit_exact():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(5)
    for i in range(25):
        h, w = (int(v) for v in rng.integers(5, 400, 2))
        nh, nw = (int(v) for v in rng.integers(3, 300, 2))
        if i % 5 == 0:
            nw = w
        if i % 7 == 0:
            nh = h
        img = synth.synthetic_rgb_image(h, w, 100 + i) if i % 2 else rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = np.array(Image.fromarray(img).resize((nw, nh), resample=Image.BICUBIC))
        assert np.array_equal(O.pil_bicubic_resize(img, nw, nh), want), (h, w, nh, nw)


def _cfg(size=336, crop=336, mode=_lib.RESIZE_SHORTEST_EDGE, pix=_lib.PIX_HF_RESCALE_NORMALIZE):
    cfg = _lib.PreprocessConfig()
    cfg.resize_mode, cfg.resize_size, cfg.crop_size, cfg.pixel_map = mode, size, crop, pix
    for i in range(3):
        cfg.mean[i], cfg.std[i] = 0.5, 0.5
    return cfg


def test_workspace_query_and_validation_need_no_gpu():
    L = _lib.lib()
    hw = np.array([[480, 640], [500, 333]], dtype=np.int32)
    hw_p = hw.ctypes.data_as(C.POINTER(C.c_int32))
    need = L.lovr_preprocess_workspace_bytes(C.byref(_cfg()), 2, hw_p)
    # temp image: 3 x rows x 336 bytes per image, rows <= source rows
    assert 2 * 3 * 336 * 336 < need < 2 * 3 * 336 * 520 + (1 << 20)
    # crop larger than the resized image, unknown modes, empty images are argument errors (reported by the query as 0)
    assert L.lovr_preprocess_workspace_bytes(C.byref(_cfg(size=224, crop=336)), 2, hw_p) == 0
    assert b"crop_size" in L.lovr_last_error()
    assert L.lovr_preprocess_workspace_bytes(C.byref(_cfg(mode=7)), 2, hw_p) == 0
    bad = np.array([[0, 10]], dtype=np.int32)
    assert L.lovr_preprocess_workspace_bytes(C.byref(_cfg()), 1, bad.ctypes.data_as(C.POINTER(C.c_int32))) == 0


def test_oracle_matches_pillow_property_based():
    """Hypothesis-driven sizes (incl. 1-pixel sides, identity, extreme aspect ratios): restatement == Pillow, bit for bit."""
    Image = pytest.importorskip("PIL.Image")
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=40, deadline=None, derandomize=True)
    @hyp.given(st.integers(1, 260), st.integers(1, 260), st.integers(1, 200), st.integers(1, 200), st.integers(0, 2 ** 31 - 1))
    def check(h, w, nh, nw, seed):
        img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = np.array(Image.fromarray(img).resize((nw, nh), resample=Image.BICUBIC))
        assert np.array_equal(O.pil_bicubic_resize(img, nw, nh), want), (h, w, nh, nw)

    check()


def test_hf_pipeline_restated_against_transformers_pil_backend():
    """The whole processor pipeline (shortest-edge resize, centre crop, rescale, normalise) against transformers' own
    PIL-backend CLIP processor — the descendant of the 4.31 slow processor the reference pins."""
    Image = pytest.importorskip("PIL.Image")
    mod = pytest.importorskip("transformers.models.clip.image_processing_pil_clip")
    proc = mod.CLIPImageProcessorPil(size={"shortest_edge": 224}, crop_size={"height": 224, "width": 224})
    rng = np.random.default_rng(2)
    for h, w in ((300, 450), (512, 230), (224, 224), (100, 90)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = proc.preprocess(Image.fromarray(img), return_tensors="np")["pixel_values"][0]
        got = O.hf_preprocess(img, 224, 224, proc.image_mean, proc.image_std)
        assert np.array_equal(want, got), (h, w)
