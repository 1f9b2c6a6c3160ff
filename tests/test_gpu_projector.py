"""mm_projector (mlp2x_gelu / linear) on the device against nn.Sequential(Linear, GELU, Linear) in fp32 torch, and the
full encode_images(tower(s), projector) path of llava_arch.py:260-286."""
import types

import pytest
import torch
import torch.nn as nn

import gpu_util as G
from law_of_vision_representation_in_mllms_b200 import projector as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ptype,din", [("mlp2x_gelu", 1024), ("mlp2x_gelu", 2048), ("linear", 1024), ("mlp3x_gelu", 768)])
def test_projector_matches_torch_sequential(ptype, din):
    torch.manual_seed(0)
    depth = 1 if ptype == "linear" else int(ptype[3])
    mods = [nn.Linear(din, 4096)]
    for _ in range(1, depth):
        mods += [nn.GELU(), nn.Linear(4096, 4096)]
    ref = nn.Sequential(*mods)
    sd = {k: v.detach().to(torch.bfloat16).float() for k, v in ref.state_dict().items()}   # bf16-representable weights
    ref.load_state_dict(sd)
    cfg = types.SimpleNamespace(mm_projector_type=ptype, mm_hidden_size=din, hidden_size=4096,
                                lovr_projector_state_dict=sd)
    proj = P.build_vision_projector(cfg)
    x = torch.randn(3, 577, din).to(torch.bfloat16)
    want = ref(x.float())
    got = proj(x.cuda())
    assert got.dtype == torch.bfloat16 and tuple(got.shape) == (3, 577, 4096)
    assert G.rel_err(got.cpu(), want) < 4e-3                      # bf16 hidden + bf16 output rounding
    got32 = proj(x.cuda().float())                                # caller dtype fp32 -> fp32 out
    assert got32.dtype == torch.float32
    assert G.rel_err(got32.cpu(), want) < 2.5e-3


@pytest.mark.parametrize("name", ["mlp2x_gelu_1024", "mlp2x_gelu_2048", "linear_1024", "mlp3x_gelu_768"])
@pytest.mark.parametrize("od,tol", [("bf16", 5e-3), ("fp16", 7e-4)])     # measured 2.3e-3 ... 4.0e-3 / 2.9e-4 ... 5.1e-4
def test_projector_matches_reference_builder_fixture(name, od, tol, golden_dir):
    """tests/golden/projector.pt: outputs of the module the reference's own build_vision_projector
    (llava/model/multimodal_projector/builder.py:34-59, loaded by path in oracle/make_golden.py) builds, with seeded fp32
    weights; the weights enter here through the reference module's own state_dict keys ('linear' = bare 'weight'/'bias')."""
    import os
    from oracle import synth
    rec = torch.load(os.path.join(golden_dir, "projector.pt"))[name]
    ptype, din, dh, linears, x = synth.projector_inputs(name)
    flat = [t for wb in linears for t in wb]
    sd = dict(zip(rec["keys"], flat))
    cfg = types.SimpleNamespace(mm_projector_type=ptype, mm_hidden_size=din, hidden_size=dh, lovr_projector_state_dict=sd,
                                lovr_operand_dtype=od)
    got = P.build_vision_projector(cfg)(x.cuda())
    assert got.dtype == torch.float32 and tuple(got.shape) == tuple(rec["out"].shape)
    rel = G.rel_err(got.cpu(), rec["out"])
    print(f"projector {name} {od}: rel {rel:.2e}")
    assert rel < tol, rel


def test_projector_type_errors():
    with pytest.raises(ValueError, match="Unknown projector type"):
        P.build_vision_projector(types.SimpleNamespace(mm_projector_type="bogus", mm_hidden_size=1024, hidden_size=4096))
    assert isinstance(P.build_vision_projector(types.SimpleNamespace(mm_projector_type="identity")), P.IdentityMap)
