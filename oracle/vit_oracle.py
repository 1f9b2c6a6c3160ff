"""ORACLE (test infrastructure — never imported by the product path).

CPU fp32 restatement of the vision-tower forward that the reference delegates to the
third-party `transformers` package (pinned 4.31.0 in /root/reference/pyproject.toml:16,
4.38.2 in run.sh:5; NOT vendored under /root/reference).  The published algorithm restated
here is that of transformers' `modeling_clip.py`, `modeling_dinov2.py`, `modeling_siglip.py`
as driven by the reference wrappers:

  * CLIPVisionTower.forward / feature_select   llava/model/multimodal_encoder/clip_encoder.py:29-51
  * DinoV2VisionTower.forward                  llava/model/multimodal_encoder/dinov2_encoder.py:32-54
  * SigLipVisionTower.forward                  llava/model/multimodal_encoder/siglip_encoder.py:30-52

Pinning: `oracle/make_golden.py` runs the *reference wrapper classes themselves* (loaded by file
path from /root/reference, over transformers 5.5.0 — the version present here) on seeded
weights/images and stores slices in tests/golden/vit_*.pt; tests/test_oracle_vit.py checks this
restatement against those slices.  The reference ships no tests or golden vectors for this
boundary, so those generated fixtures are the only pin.
"""
import math

import torch
import torch.nn.functional as F

ACT_NONE, ACT_QUICK_GELU, ACT_GELU_ERF, ACT_GELU_TANH = 0, 1, 2, 3


def activation(x, act):
    if act == ACT_QUICK_GELU:      # transformers activations.QuickGELUActivation
        return x * torch.sigmoid(1.702 * x)
    if act == ACT_GELU_ERF:        # nn.GELU()
        return F.gelu(x)
    if act == ACT_GELU_TANH:       # gelu_pytorch_tanh
        return F.gelu(x, approximate="tanh")
    return x


def embeddings(cw, spec, pixels):
    """conv patch-embed -> flatten -> [cls] -> +pos -> [pre_layrnorm]   (hidden_states[0])."""
    D, P = spec.hidden_size, spec.patch_size
    x = F.conv2d(pixels.float(), cw["patch_w"], cw["patch_b"], stride=P)      # [N,D,g,g]
    x = x.flatten(2).transpose(1, 2)                                          # [N,g*g,D] row-major grid
    if cw["cls"] is not None:
        cls = cw["cls"].reshape(1, 1, D).expand(x.shape[0], 1, D)
        x = torch.cat([cls, x], dim=1)
    x = x + cw["pos"].unsqueeze(0)
    if cw["pre_ln_w"] is not None:
        x = F.layer_norm(x, (D,), cw["pre_ln_w"], cw["pre_ln_b"], spec.ln_eps)
    return x


def attention(x, lw, num_heads):
    N, T, D = x.shape
    hd = D // num_heads
    qkv = F.linear(x, lw["qkv_w"], lw["qkv_b"]).reshape(N, T, 3, num_heads, hd)
    q, k, v = qkv[:, :, 0].transpose(1, 2), qkv[:, :, 1].transpose(1, 2), qkv[:, :, 2].transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) * (hd ** -0.5)
    p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, v).transpose(1, 2).reshape(N, T, D)
    return F.linear(o, lw["o_w"], lw["o_b"])


def encoder_layer(x, lw, spec):
    D = spec.hidden_size
    h = attention(F.layer_norm(x, (D,), lw["ln1_w"], lw["ln1_b"], spec.ln_eps), lw, spec.num_heads)
    if lw["ls1"] is not None:
        h = h * lw["ls1"]
    x = x + h
    m = F.layer_norm(x, (D,), lw["ln2_w"], lw["ln2_b"], spec.ln_eps)
    m = F.linear(activation(F.linear(m, lw["fc1_w"], lw["fc1_b"]), spec.act), lw["fc2_w"], lw["fc2_b"])
    if lw["ls2"] is not None:
        m = m * lw["ls2"]
    return x + m


@torch.no_grad()
def tower_forward(cw, spec, pixels, select_layer=-2, select_feature="patch"):
    """`hidden_states[select_layer]` then feature_select (clip_encoder.py:29-37)."""
    n = spec.num_layers + 1
    nrun = select_layer % n
    x = embeddings(cw, spec, pixels)
    for i in range(nrun):
        x = encoder_layer(x, cw["layers"][i], spec)
    if select_feature == "patch":
        x = x[:, 1:]
    elif select_feature != "cls_patch":
        raise ValueError(f"Unexpected select feature: {select_feature}")
    return x
