"""Vision towers on the device, through the reference-shaped Python interface (towers.py -> C ABI), against
(a) fixtures produced by the reference wrappers themselves (tests/golden/vit_*.pt) and (b) the CPU oracle.
Error measure: ||got-want||_F / ||want||_F against the fp32 reference.

Tolerances (written here once):
  * REL_TOL = 1e-3 — BASELINE.json north_star ("patch features within 1e-3 relative (bf16)").  With bf16 operands the
    benchmark weights (HF-init-style, cases *_hf) land at ~3e-3 over 23 layers (the reference's own bf16 path: 1.26e-2);
    operand rounding 2^-9 per GEMM input makes 1e-3 unreachable for ANY bf16-operand pipeline at this depth.  The same
    kernels with fp16 operands (2^-12) meet REL_TOL; both modes are tested.
  * "stress" weights (wide q/k -> sharp softmax, full LayerNorm affine/bias/LayerScale) amplify bf16 operand
    rounding: the reference's OWN bf16 path (tower.to(bfloat16), llava/feature/extract.py:228) deviates 1.1e-2
    (3 layers) to 3.3e-2 (23 layers) from its fp32 path on these weights — recorded per case in the fixtures as
    `ref_bf16_rel` by oracle/make_golden.py.  No bf16-operand implementation can meet 1e-3 there; the bar for these
    cases is to be strictly better than the reference's bf16 path (we keep residual/LN/softmax in fp32).
"""
import os
import types

import pytest
import torch

from law_of_vision_representation_in_mllms_b200 import _lib, specs as S, towers
from law_of_vision_representation_in_mllms_b200.weights import canonical_weights, synthetic_images
from oracle import vit_oracle
from oracle.make_golden import VIT_CASES, case_pos_interp, case_style, vit_case_inputs
import gpu_util as G

pytestmark = pytest.mark.gpu
REL_TOL = 1e-3

_CLS = {S.FAMILY_CLIP: towers.CLIPVisionTower, S.FAMILY_DINOV2: towers.DinoV2VisionTower,
        S.FAMILY_SIGLIP: towers.SigLipVisionTower}


OUTLIER_CHANNELS = (7, 300)      # weights._outlier_state_dict


def tolerance_for(name, rec):
    """bf16 operands (see module docstring): a fraction of the reference's OWN bf16-path deviation from its fp32 path."""
    ref = rec.get("ref_bf16_rel")
    if case_style(name) in ("hf_init", "hf_init_raw"):
        # measured 3.1e-3 ... 3.5e-3 (pre-rounded weights) and 4.6e-3 ... 5.6e-3 (`_raw`: bf16 weight quantisation inside)
        frac = 0.4 if case_style(name) == "hf_init" else 0.5
        return frac * ref if ref is not None else 6e-3      # SigLIP: the reference's wrapper does not run on the CPU
    return 0.75 * ref if ref is not None else 1.2e-2


def _build(case, attn_impl=_lib.ATTN_AUTO, operand_dtype="bf16", ln_fold=0):
    name, key, nl, size, n, sl, sf = case
    spec, sd, imgs = vit_case_inputs(case)
    args = types.SimpleNamespace(mm_vision_select_layer=sl, mm_vision_select_feature=sf, lovr_spec=spec,
                                 lovr_state_dict=sd, lovr_image_size=size, lovr_max_batch=4,
                                 lovr_operand_dtype=operand_dtype, lovr_pos_interp=case_pos_interp(name), lovr_ln_fold=ln_fold)
    tower = _CLS[spec.family](key, args)
    tower._attn_impl = attn_impl
    return tower, spec, sd, imgs


@pytest.mark.parametrize("case", VIT_CASES, ids=[c[0] for c in VIT_CASES])
def test_tower_matches_reference_fixture(case, golden_dir):
    name = case[0]
    rec = torch.load(os.path.join(golden_dir, f"vit_{name}.pt"))
    tower, spec, sd, imgs = _build(case)
    out = tower(imgs.cuda())
    assert out.dtype == imgs.dtype and tuple(out.shape) == tuple(rec["shape"])
    got = out[:, rec["rows"]].cpu()
    rel = G.rel_err(got, rec["values"])
    tol = tolerance_for(name, rec)
    print(f"{name}: rel {rel:.3e} (tolerance {tol:.3e}, reference bf16 path {rec.get('ref_bf16_rel')})")
    assert rel < tol, (rel, tol)
    assert torch.allclose(out.float().flatten(1).norm(dim=1).cpu(), rec["fro"], rtol=2e-3)


FULL_CASES = [c for c in VIT_CASES if c[2] is None]


@pytest.mark.parametrize("case", FULL_CASES, ids=[c[0] for c in FULL_CASES])
def test_tower_fp16_operands_meet_north_star_tolerance(case, golden_dir):
    """fp16 tensor-core operands — the mode bench.py's headline runs in, and the dtype the reference's eval path casts towers
    to (llava/model/builder.py:143; SigLIP hard-codes it, siglip_encoder.py:59-60).  Every family at full depth:
      * benchmark weights (`_hf`) and the same draw NOT pre-rounded to 16 bits (`_raw`: weight quantisation is inside the
        measured error): within the north_star 1e-3 of the reference's fp32 path;
      * stress / outlier weights (sharp softmax, |x| > 100 in the residual stream): finite, reported, and several times
        closer to the fp32 path than the reference's OWN fp16 path (`ref_fp16_rel` in the fixture) — the CPU emulation of
        fp16 operand rounding (oracle/vit_oracle.py, scripts/predict_tower_error.py) predicts 0.8e-3 ... 3.7e-3 there: a property of 2^-12 operand
        rounding under those weights, not of this implementation."""
    name = case[0]
    rec = torch.load(os.path.join(golden_dir, f"vit_{name}.pt"))
    tower, spec, sd, imgs = _build(case, operand_dtype="fp16")
    assert tower.dtype == torch.float16
    out = tower(imgs.cuda())
    assert torch.isfinite(out).all()
    got = out[:, rec["rows"]].cpu()
    rel = G.rel_err(got, rec["values"])
    style = case_style(name)
    extra = ""
    if style == "outlier":
        keep = [c for c in range(got.shape[-1]) if c not in OUTLIER_CHANNELS]
        rel_rest = G.rel_err(got[..., keep], rec["values"][..., keep])
        extra = f", excluding the outlier channels {rel_rest:.3e}, |x|max {rec['abs_max']:.0f}"
        assert rec["abs_max"] > 100
    print(f"{name} fp16 operands: rel {rel:.3e}{extra} (reference bf16 path {rec.get('ref_bf16_rel')}, "
          f"reference fp16 path {rec.get('ref_fp16_rel')})")
    if style in ("hf_init", "hf_init_raw"):
        assert rel < REL_TOL, rel
    else:
        ref16 = rec.get("ref_fp16_rel")
        assert rel < (0.5 * ref16 if ref16 is not None and style != "outlier" else 6e-3), (rel, ref16)
        if style == "outlier":
            assert rel_rest < 1.2e-2        # measured 7.6e-3 (CLIP; 3.7e-3 over all channels), reference fp16 path 2.8e-2


FOLD_CASES = ["clip336_l3", "dinov2_336_l3", "siglip224_l3", "clip224_l3_clspatch", "clip336_full_hf", "dinov2_336_full_raw",
              "siglip224_full_hf", "clip336_full_outlier", "clip336_full"]


@pytest.mark.parametrize("fold", [3, 1])
@pytest.mark.parametrize("cname", FOLD_CASES)
def test_tower_with_layernorm_folded_into_the_gemms(cname, fold, golden_dir):
    """`lovr_ln_fold` (bit mask: 1 = LayerNorm-1 into qkv, 2 = LayerNorm-2 into fc1; 3 = both): no LayerNorm kernel — the residual epilogues emit the 16-bit residual copy and row statistics, the qkv /
    fc1 epilogues apply them (include/lovr_b200.h `ln_fold`).  Same fixtures, same bars as the unfolded path (fp16 operands:
    1e-3 on benchmark / un-rounded weights); the CPU emulation predicts errors equal to the unfolded path on those weights and
    ~20 % larger on the stress weights (gamma is folded into the rounded weight matrix)."""
    case = next(c for c in VIT_CASES if c[0] == cname)
    rec = torch.load(os.path.join(golden_dir, f"vit_{cname}.pt"))
    launches0 = _lib.launch_count()
    res = {}
    for od in ("fp16", "bf16"):
        tower, spec, sd, imgs = _build(case, operand_dtype=od, ln_fold=fold)
        out = tower(imgs.cuda())
        assert torch.isfinite(out).all() and tuple(out.shape) == tuple(rec["shape"])
        res[od] = G.rel_err(out[:, rec["rows"]].cpu(), rec["values"])
    style = case_style(cname)
    print(f"{cname} folded LayerNorm (mask {fold}): fp16 {res['fp16']:.3e}, bf16 {res['bf16']:.3e}")
    if style in ("hf_init", "hf_init_raw"):
        assert res["fp16"] < REL_TOL
    elif rec.get("ref_fp16_rel") is not None and style != "outlier":
        assert res["fp16"] < 0.5 * rec["ref_fp16_rel"]
    else:
        assert res["fp16"] < 6e-3
    assert res["bf16"] < (tolerance_for(cname, rec) if style != "stress" else 1.2 * tolerance_for(cname, rec))
    assert _lib.launch_count() > launches0


@pytest.mark.parametrize("cname", ["clip336_l3", "dinov2_336_l3", "siglip224_l3"])
def test_tower_matches_oracle_full_tensor_and_simt_attention(cname):
    case = next(c for c in VIT_CASES if c[0] == cname)
    name, key, nl, size, n, sl, sf = case
    tower, spec, sd, imgs = _build(case)
    cw = canonical_weights(spec, sd, image_size=size, pos_interp=case_pos_interp(name))
    want = vit_oracle.tower_forward(cw, spec.at_resolution(size), imgs, sl, tower.select_feature)
    out = tower(imgs.cuda()).cpu()
    e1 = G.rel_err(out, want)
    errs = {}
    for name_, impl in (("simt", _lib.ATTN_SIMT_CHECK), ("tcgen05_v3", _lib.ATTN_TCGEN05_V3),
                        ("tcgen05_v4_persistent", _lib.ATTN_TCGEN05_V4)):
        tower._attn_impl = impl
        errs[name_] = G.rel_err(tower(imgs.cuda()).cpu(), want)
    print(f"{cname}: default attention {e1:.3e}, others {errs}")
    assert e1 < 9e-3 and max(errs.values()) < 9e-3          # stress weights, 2 layers run (see module docstring)


def test_caller_dtype_list_input_and_batch_chunking():
    case = next(c for c in VIT_CASES if c[0] == "clip336_l3")
    tower, spec, sd, _ = _build(case)
    imgs = synthetic_images(6, 336, seed=5)           # > lovr_max_batch=4 -> two native calls
    ref = tower(imgs.cuda())
    for dt in (torch.bfloat16, torch.float16):
        o = tower(imgs.cuda().to(dt))
        assert o.dtype == dt
        assert G.rel_err(o, ref) < 6e-3
    lst = tower([im for im in imgs.cuda()])
    assert isinstance(lst, list) and tuple(lst[0].shape) == (1, 576, 1024)
    assert torch.equal(torch.cat(lst), ref)          # batch-size independent results


def test_select_feature_and_errors():
    case = next(c for c in VIT_CASES if c[0] == "clip224_l3_clspatch")
    tower, spec, sd, imgs = _build(case)
    tower.select_feature = "bogus"
    with pytest.raises(ValueError, match="Unexpected select feature"):
        tower(imgs.cuda())
    tower.select_feature = "patch"
    with pytest.raises(ValueError):                    # CLIP is fixed-resolution (HF raises on mismatch)
        tower(synthetic_images(1, 336).cuda())


def test_dual_tower_concat_matches_separate_towers():
    c1 = next(c for c in VIT_CASES if c[0] == "clip336_l3")
    c2 = next(c for c in VIT_CASES if c[0] == "dinov2_336_l3")
    t1, _, _, im1 = _build(c1)
    t2, _, _, im2 = _build(c2)
    both = torch.nn.ModuleList([t1, t2])
    fused = towers.encode_images(both, [im1.cuda(), im2.cuda()])
    assert tuple(fused.shape) == (2, 576, 2048)
    assert torch.equal(fused[..., :1024], t1(im1.cuda()))
    assert torch.equal(fused[..., 1024:], t2(im2.cuda()))
