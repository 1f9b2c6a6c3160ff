"""The ViT oracle restatement (oracle/vit_oracle.py) against fixtures produced by the reference's own
wrapper classes over HF transformers (oracle/make_golden.py).  CPU only."""
import os

import pytest
import torch

from law_of_vision_representation_in_mllms_b200.weights import canonical_weights
from oracle import vit_oracle
from oracle.make_golden import VIT_CASES, vit_case_inputs

# full-depth cases are slow on CPU; two of them are enough to pin depth handling
FAST = [c for c in VIT_CASES if c[2] is not None] + [c for c in VIT_CASES if c[0] in ("siglip224_full", "clip224_full")]


@pytest.mark.parametrize("case", FAST, ids=[c[0] for c in FAST])
def test_vit_oracle_matches_reference_wrapper(case, golden_dir):
    name, key, nl, size, n, sl, sf = case
    rec = torch.load(os.path.join(golden_dir, f"vit_{name}.pt"))
    spec, sd, imgs = vit_case_inputs(case)
    if name == "clip224_full":
        imgs = imgs[:1]
    cw = canonical_weights(spec, sd, image_size=size)
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
