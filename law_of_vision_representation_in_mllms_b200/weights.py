"""Weight ingest for the towers.

Two sources, one format:

* `synthetic_state_dict(spec, seed)` — seeded random weights under the HuggingFace parameter
  names of the family, because no checkpoints exist offline.  GEMM
  matrices are rounded to bf16-representable values so the fp32 oracle and the bf16 device
  path consume *identical* weights (the reference casts the tower to bf16/fp16 the same way:
  llava/feature/extract.py:228, llava/model/builder.py:143).
* a real HF `state_dict()` — same names, so real checkpoints drop in.

`canonical_weights(spec, state_dict)` maps either to a family-independent dict of fp32 CPU
tensors that both `oracle/vit_oracle.py` and the device packer (`towers.py`) read.
"""
import math

import torch
import torch.nn.functional as F

from .specs import FAMILY_CLIP, FAMILY_DINOV2, FAMILY_SIGLIP, TowerSpec


def _bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def synthetic_state_dict(spec: TowerSpec, seed: int = 0, style: str = "stress"):
    """Seeded random weights with HF names for `spec` (CPU fp32).

    style="stress": wide q/k projections (sharp softmax), non-trivial LayerNorm affine, biases and
    LayerScale everywhere — every term of the arithmetic matters and rounding errors are amplified.
    style="hf_init": the distribution HF `_init_weights` draws (std 0.02 linears, unit LayerNorm, zero
    biases), i.e. the "HF-initialised random weights" SURVEY.md §8d prescribes for benchmarks."""
    if style == "hf_init":
        return _hf_init_state_dict(spec, seed)
    if style == "hf_init_raw":      # same draw, weights NOT pre-rounded to bf16: weight quantisation is inside the measured error
        return _hf_init_state_dict(spec, seed, round_weights=False)
    if style == "outlier":
        return _outlier_state_dict(spec, seed)
    if style != "stress":
        raise ValueError(f"unknown weight style {style!r}")
    g = torch.Generator(device="cpu").manual_seed(seed)
    D, I, L, P = spec.hidden_size, spec.intermediate_size, spec.num_layers, spec.patch_size
    sd = {}

    def rn(*shape, std=1.0, mean=0.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std + mean

    w_std = 0.03
    proj_std = 0.03
    npos = spec.pos_grid * spec.pos_grid + (1 if spec.has_cls else 0)
    patch_w = _bf16_round(rn(D, 3, P, P, std=0.03))
    if spec.family == FAMILY_CLIP:
        e = "vision_model.embeddings."
        sd[e + "class_embedding"] = rn(D, std=0.5)
        sd[e + "patch_embedding.weight"] = patch_w
        sd[e + "position_embedding.weight"] = rn(npos, D, std=0.3)
        sd["vision_model.pre_layrnorm.weight"] = rn(D, std=0.1, mean=1.0)
        sd["vision_model.pre_layrnorm.bias"] = rn(D, std=0.1)
        lp = "vision_model.encoder.layers.{}."
        names = dict(ln1="layer_norm1", ln2="layer_norm2", q="self_attn.q_proj", k="self_attn.k_proj",
                     v="self_attn.v_proj", o="self_attn.out_proj", fc1="mlp.fc1", fc2="mlp.fc2")
        sd["vision_model.post_layernorm.weight"] = torch.ones(D)
        sd["vision_model.post_layernorm.bias"] = torch.zeros(D)
    elif spec.family == FAMILY_SIGLIP:
        e = "vision_model.embeddings."
        sd[e + "patch_embedding.weight"] = patch_w
        sd[e + "patch_embedding.bias"] = rn(D, std=0.1)
        sd[e + "position_embedding.weight"] = rn(npos, D, std=0.3)
        lp = "vision_model.encoder.layers.{}."
        names = dict(ln1="layer_norm1", ln2="layer_norm2", q="self_attn.q_proj", k="self_attn.k_proj",
                     v="self_attn.v_proj", o="self_attn.out_proj", fc1="mlp.fc1", fc2="mlp.fc2")
        sd["vision_model.post_layernorm.weight"] = torch.ones(D)
        sd["vision_model.post_layernorm.bias"] = torch.zeros(D)
    elif spec.family == FAMILY_DINOV2:
        e = "embeddings."
        sd[e + "cls_token"] = rn(1, 1, D, std=0.5)
        sd[e + "mask_token"] = torch.zeros(1, D)
        sd[e + "position_embeddings"] = rn(1, npos, D, std=0.3)
        sd[e + "patch_embeddings.projection.weight"] = patch_w
        sd[e + "patch_embeddings.projection.bias"] = rn(D, std=0.1)
        lp = "encoder.layer.{}."
        names = dict(ln1="norm1", ln2="norm2", q="attention.attention.query", k="attention.attention.key",
                     v="attention.attention.value", o="attention.output.dense", fc1="mlp.fc1", fc2="mlp.fc2")
        sd["layernorm.weight"] = torch.ones(D)
        sd["layernorm.bias"] = torch.zeros(D)
    else:
        raise ValueError(spec.family)

    for i in range(L):
        p = lp.format(i)
        for ln in ("ln1", "ln2"):
            sd[p + names[ln] + ".weight"] = rn(D, std=0.1, mean=1.0)
            sd[p + names[ln] + ".bias"] = rn(D, std=0.1)
        for n in ("q", "k", "v"):
            # q/k a bit larger so the softmax is far from uniform
            s = 0.06 if n in ("q", "k") else w_std
            sd[p + names[n] + ".weight"] = _bf16_round(rn(D, D, std=s))
            sd[p + names[n] + ".bias"] = rn(D, std=0.1)
        sd[p + names["o"] + ".weight"] = _bf16_round(rn(D, D, std=proj_std))
        sd[p + names["o"] + ".bias"] = rn(D, std=0.05)
        sd[p + names["fc1"] + ".weight"] = _bf16_round(rn(I, D, std=w_std))
        sd[p + names["fc1"] + ".bias"] = rn(I, std=0.1)
        sd[p + names["fc2"] + ".weight"] = _bf16_round(rn(D, I, std=proj_std / 2))
        sd[p + names["fc2"] + ".bias"] = rn(D, std=0.05)
        if spec.layerscale:
            sd[p + "layer_scale1.lambda1"] = rn(D, std=0.2, mean=1.0)
            sd[p + "layer_scale2.lambda1"] = rn(D, std=0.2, mean=1.0)
    return sd


def _hf_init_state_dict(spec, seed, round_weights=True):
    rnd = _bf16_round if round_weights else (lambda t: t)
    sd = synthetic_state_dict(spec, seed, style="stress")   # same keys, then overwrite the distributions
    g = torch.Generator(device="cpu").manual_seed(seed + 7919)
    for k, v in sd.items():
        if k.endswith("lambda1"):
            sd[k] = torch.ones_like(v)
        elif "norm" in k and k.endswith(".weight") and v.dim() == 1:
            sd[k] = torch.ones_like(v)
        elif k.endswith(".bias"):
            sd[k] = torch.zeros_like(v)
        elif v.dim() >= 2 and "position" not in k and "cls_token" not in k and "mask_token" not in k:
            sd[k] = rnd(torch.randn(v.shape, generator=g) * 0.02)
        else:                                   # class / cls / position embeddings
            sd[k] = torch.randn(v.shape, generator=g) * 0.02
    return sd


POS_INTERP_MODES = ("transformers-4.38", "transformers-5.5")
DEFAULT_POS_INTERP = "transformers-4.38"     # the reference pins transformers 4.31.0 (pyproject.toml:16) / 4.38.2 (run.sh:5)


def _outlier_state_dict(spec, seed):
    """Activation statistics of trained checkpoints that random weights lack (VERDICT r1 weak #1b): a few residual-stream
    channels carry MASSIVE activations (hundreds; CLIP/DINOv2 ViT-L are known for them), LayerNorm gains are far from one,
    attention logits are sharp.  fp32 (un-rounded) weights; built on the hf_init draw:
      * position table: +120 / -90 on two channels for every token, so the residual holds |x| > 100 from layer 0 on,
      * every fc2 bias pushes the first of them further (+8 per layer: ~300 at depth 23),
      * LayerNorm gains drawn in [0.25, 2.5] with the outlier channels' gains small (what trained models do), biases non-zero,
      * q/k projections x3 (sharp softmax), all biases ~0.05."""
    sd = _hf_init_state_dict(spec, seed, round_weights=False)
    g = torch.Generator(device="cpu").manual_seed(seed + 104729)
    ch = (7 % spec.hidden_size, 300 % spec.hidden_size)
    for k, v in list(sd.items()):
        if "position_embedding" in k:
            v = v.clone()
            v[..., ch[0]] += 120.0
            v[..., ch[1]] -= 90.0
            sd[k] = v
        elif "norm" in k and k.endswith(".weight") and v.dim() == 1 and "post_layernorm" not in k and not k.startswith("layernorm"):
            w = 0.25 + 2.25 * torch.rand(v.shape, generator=g)
            w[ch[0]] = 0.05
            w[ch[1]] = 0.05
            sd[k] = w
        elif k.endswith("fc2.bias"):
            b = 0.05 * torch.randn(v.shape, generator=g)
            b[ch[0]] = 8.0
            sd[k] = b
        elif k.endswith(".bias") and v.dim() == 1 and "post_layernorm" not in k and not k.startswith("layernorm"):
            sd[k] = 0.05 * torch.randn(v.shape, generator=g)
        elif any(t in k for t in ("q_proj.weight", "k_proj.weight", "attention.query.weight", "attention.key.weight")):
            sd[k] = v * 3.0
    return sd


def dinov2_interpolated_pos(pos, pos_grid, new_grid, mode=DEFAULT_POS_INTERP):
    """HF `Dinov2Embeddings.interpolate_pos_encoding`: bicubic, align_corners=False, fp32.  `pos` is [1, 1+g*g, D].

    Two published forms exist and they give DIFFERENT tables (max deviation ~1e-2 of the table's scale at 37 -> 24):
      * "transformers-4.38" (4.31.0 ... 4.38.2, the versions the reference pins; modeling_dinov2.py `interpolate_pos_encoding`):
        `height, width = height + 0.1, width + 0.1`, then `F.interpolate(scale_factor=(height / sqrt(N), width / sqrt(N)))`.
        PyTorch uses the GIVEN scale factor for the source coordinates (src = (dst + 0.5) / scale_factor - 0.5) and
        floor(in * scale_factor) for the output size, so the sampling grid is stretched by (g + 0.1) / g against ...
      * "transformers-5.5" (the version in this image; modeling_dinov2.py:57-95): `F.interpolate(size=(g, g))`, i.e.
        src = (dst + 0.5) * in / out - 0.5.
    The default follows the reference's pin so that a drop-in next to its stack reproduces its DINOv2 features."""
    if new_grid == pos_grid:
        return pos
    if mode not in POS_INTERP_MODES:
        raise ValueError(f"pos_interp must be one of {POS_INTERP_MODES}, got {mode!r}")
    cls_pos, patch_pos = pos[:, :1], pos[:, 1:]
    D = pos.shape[-1]
    patch_pos = patch_pos.reshape(1, pos_grid, pos_grid, D).permute(0, 3, 1, 2)
    if mode == "transformers-5.5":
        patch_pos = F.interpolate(patch_pos.float(), size=(new_grid, new_grid), mode="bicubic", align_corners=False)
    else:
        sf = float((new_grid + 0.1) / math.sqrt(pos_grid * pos_grid))
        patch_pos = F.interpolate(patch_pos.float(), scale_factor=(sf, sf), mode="bicubic", align_corners=False)
        if patch_pos.shape[-1] != new_grid or patch_pos.shape[-2] != new_grid:
            raise ValueError("Width or height does not match with the interpolated position embeddings")
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, D)
    return torch.cat([cls_pos, patch_pos], dim=1)


def canonical_weights(spec: TowerSpec, sd, image_size=None, pos_interp=DEFAULT_POS_INTERP):
    """HF names -> family-independent names, fp32 CPU.

    Keys: patch_w [D,3,P,P], patch_b [D]|None, cls [D]|None, pos [T,D] (already resized for
    `image_size`), pre_ln_w/b|None, and per layer i: ln1_w, ln1_b, qkv_w [3D,D], qkv_b [3D],
    o_w, o_b, ls1 [D]|None, ln2_w, ln2_b, fc1_w, fc1_b, fc2_w, fc2_b, ls2|None.
    """
    image_size = spec.image_size if image_size is None else image_size
    spec_r = spec.at_resolution(image_size)
    f = lambda k: sd[k].detach().to("cpu", torch.float32)
    D = spec.hidden_size
    cw = {"layers": []}
    if spec.family == FAMILY_CLIP:
        e = "vision_model.embeddings."
        cw["patch_w"], cw["patch_b"] = f(e + "patch_embedding.weight"), None
        cw["cls"] = f(e + "class_embedding").reshape(D)
        cw["pos"] = f(e + "position_embedding.weight")
        cw["pre_ln_w"], cw["pre_ln_b"] = f("vision_model.pre_layrnorm.weight"), f("vision_model.pre_layrnorm.bias")
        lp = "vision_model.encoder.layers.{}."
        n = dict(ln1="layer_norm1", ln2="layer_norm2", q="self_attn.q_proj", k="self_attn.k_proj",
                 v="self_attn.v_proj", o="self_attn.out_proj")
    elif spec.family == FAMILY_SIGLIP:
        e = "vision_model.embeddings."
        cw["patch_w"], cw["patch_b"] = f(e + "patch_embedding.weight"), f(e + "patch_embedding.bias")
        cw["cls"] = None
        cw["pos"] = f(e + "position_embedding.weight")
        cw["pre_ln_w"] = cw["pre_ln_b"] = None
        lp = "vision_model.encoder.layers.{}."
        n = dict(ln1="layer_norm1", ln2="layer_norm2", q="self_attn.q_proj", k="self_attn.k_proj",
                 v="self_attn.v_proj", o="self_attn.out_proj")
    else:
        e = "embeddings."
        cw["patch_w"] = f(e + "patch_embeddings.projection.weight")
        cw["patch_b"] = f(e + "patch_embeddings.projection.bias")
        cw["cls"] = f(e + "cls_token").reshape(D)
        pos = f(e + "position_embeddings")
        cw["pos"] = dinov2_interpolated_pos(pos, spec.pos_grid, image_size // spec.patch_size, pos_interp)[0]
        cw["pre_ln_w"] = cw["pre_ln_b"] = None
        lp = "encoder.layer.{}."
        n = dict(ln1="norm1", ln2="norm2", q="attention.attention.query", k="attention.attention.key",
                 v="attention.attention.value", o="attention.output.dense")
    if cw["pos"].shape[0] != spec_r.tokens():
        raise ValueError(f"position table has {cw['pos'].shape[0]} rows, expected {spec_r.tokens()}")
    for i in range(spec.num_layers):
        p = lp.format(i)
        lw = dict(
            ln1_w=f(p + n["ln1"] + ".weight"), ln1_b=f(p + n["ln1"] + ".bias"),
            qkv_w=torch.cat([f(p + n[x] + ".weight") for x in "qkv"], 0),
            qkv_b=torch.cat([f(p + n[x] + ".bias") for x in "qkv"], 0),
            o_w=f(p + n["o"] + ".weight"), o_b=f(p + n["o"] + ".bias"),
            ln2_w=f(p + n["ln2"] + ".weight"), ln2_b=f(p + n["ln2"] + ".bias"),
            fc1_w=f(p + "mlp.fc1.weight"), fc1_b=f(p + "mlp.fc1.bias"),
            fc2_w=f(p + "mlp.fc2.weight"), fc2_b=f(p + "mlp.fc2.bias"),
            ls1=f(p + "layer_scale1.lambda1") if spec.layerscale else None,
            ls2=f(p + "layer_scale2.lambda1") if spec.layerscale else None,
        )
        cw["layers"].append(lw)
    return cw


def synthetic_images(n, image_size, seed=1234, dtype=torch.float32):
    """Post-normalisation pixel statistics: randn, rounded to bf16-representable
    values because every reference caller casts pixels to the tower dtype before the conv
    (clip_encoder.py:48)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(n, 3, image_size, image_size, generator=g, dtype=torch.float32)
    return _bf16_round(x).to(dtype)
