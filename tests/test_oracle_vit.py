"""The ViT oracle restatement (oracle/vit_oracle.py) against fixtures produced by the reference's own
wrapper classes over HF transformers (oracle/make_golden.py).  CPU only."""
import os

import pytest
import torch

from law_of_vision_representation_in_mllms_b200.weights import canonical_weights
from oracle import vit_oracle
from oracle.make_golden import VIT_CASES, case_pos_interp, vit_case_inputs

# full-depth cases are slow on CPU; two of them are enough to pin depth handling
FAST = [c for c in VIT_CASES if c[2] is not None] + [c for c in VIT_CASES if c[0] in ("siglip224_full", "clip224_full")]


@pytest.mark.parametrize("case", FAST, ids=[c[0] for c in FAST])
def test_vit_oracle_matches_reference_wrapper(case, golden_dir):
    name, key, nl, size, n, sl, sf = case
    rec = torch.load(os.path.join(golden_dir, f"vit_{name}.pt"))
    spec, sd, imgs = vit_case_inputs(case)
    if name == "clip224_full":
        imgs = imgs[:1]
    cw = canonical_weights(spec, sd, image_size=size, pos_interp=case_pos_interp(name))
    out = vit_oracle.tower_forward(cw, spec.at_resolution(size), imgs, sl, sf)
    assert tuple(out.shape[1:]) == tuple(rec["shape"][1:])
    got = out[:, rec["rows"]]
    want = rec["values"][: got.shape[0]]
    rel = (got - want).norm() / want.norm()
    assert rel < 2e-5, rel
    assert torch.allclose(out.flatten(1).norm(dim=1), rec["fro"][: got.shape[0]], rtol=1e-4)


def test_list_input_convention_recorded(golden_dir):
    rec = torch.load(os.path.join(golden_dir, "vit_clip336_l3.pt"))
    assert rec["list_equal_batch"] is True
    assert rec["list_shapes"] == [(1, 576, 1024), (1, 576, 1024)]


@pytest.mark.parametrize("mode", ["transformers-5.5", "transformers-4.38"])
@pytest.mark.parametrize("new_grid", [16, 24])
def test_dinov2_position_interpolation_forms(mode, new_grid):
    """weights.dinov2_interpolated_pos (host-side weight ingest) against the numpy restatement of ATen's bicubic with the
    source scale of each transformers version; the two forms differ (the reference pins the scale_factor + 0.1 one)."""
    from law_of_vision_representation_in_mllms_b200.weights import dinov2_interpolated_pos
    g = torch.Generator().manual_seed(new_grid)
    pos = torch.randn(1, 1 + 37 * 37, 32, generator=g)
    got = dinov2_interpolated_pos(pos, 37, new_grid, mode)
    want = vit_oracle.dinov2_pos_restated(pos, 37, new_grid, mode)
    assert tuple(got.shape) == (1, 1 + new_grid * new_grid, 32)
    assert (got - want).abs().max() < 5e-5          # torch computes the source coordinates in fp32, the restatement in fp64
    other = dinov2_interpolated_pos(pos, 37, new_grid, "transformers-5.5" if mode != "transformers-5.5" else "transformers-4.38")
    assert (got - other).abs().max() > 1e-2          # not interchangeable


def test_dinov2_position_interpolation_55_form_is_hf(golden_dir):
    """The 5.5 form is what THIS image's transformers computes (pins the restatement's kernel and coordinate convention to
    the third-party code); identity when the grid is the stored one."""
    import transformers as T
    from law_of_vision_representation_in_mllms_b200.weights import dinov2_interpolated_pos
    cfg = T.Dinov2Config(hidden_size=32, num_hidden_layers=1, num_attention_heads=1, image_size=518, patch_size=14)
    m = T.Dinov2Model(cfg).eval()
    g = torch.Generator().manual_seed(1)
    pos = torch.randn(1, 1 + 37 * 37, 32, generator=g)
    with torch.no_grad():
        m.embeddings.position_embeddings.copy_(pos)
        hf = m.embeddings.interpolate_pos_encoding(torch.zeros(1, 1 + 24 * 24, 32), 336, 336)
    assert torch.equal(hf, dinov2_interpolated_pos(pos, 37, 24, "transformers-5.5"))
    assert dinov2_interpolated_pos(pos, 37, 37, "transformers-4.38") is pos


def test_operand_rounding_emulation_predicts_16bit_error():
    """vit_oracle's operand-rounding emulation (used to set the GPU tolerances without a GPU): fp16 operands land ~8x
    closer to the fp32 forward than bf16 operands (2^-12 vs 2^-9 per GEMM input)."""
    from law_of_vision_representation_in_mllms_b200 import specs as S
    from law_of_vision_representation_in_mllms_b200.weights import synthetic_images, synthetic_state_dict
    spec = S.tiny_spec(S.TOWER_SPECS["openai/clip-vit-large-patch14"], 3)
    sd = synthetic_state_dict(spec, seed=1, style="hf_init_raw")
    cw = canonical_weights(spec, sd)
    imgs = synthetic_images(1, 224, seed=2)
    want = vit_oracle.tower_forward(cw, spec, imgs)
    e16 = (vit_oracle.tower_forward(cw, spec, imgs, operand_dtype=torch.float16) - want).norm() / want.norm()
    eb16 = (vit_oracle.tower_forward(cw, spec, imgs, operand_dtype=torch.bfloat16) - want).norm() / want.norm()
    assert e16 < 5e-4 and 4 * e16 < eb16 < 16 * e16, (e16, eb16)
