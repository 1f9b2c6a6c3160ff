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


def _ident(t):
    return t


def operand_rounder(operand_dtype):
    """Emulation hook: round a tensor to the 16-bit tensor-core operand format and back (None -> exact fp32 oracle).
    Used only to PREDICT the device path's operand-rounding error on the CPU (tests/test_oracle_vit.py); the parity
    reference is always the un-rounded fp32 forward."""
    if operand_dtype is None:
        return _ident
    return lambda t: t.to(operand_dtype).to(torch.float32)


def attention(x, lw, num_heads, rnd=_ident):
    N, T, D = x.shape
    hd = D // num_heads
    qkv = rnd(F.linear(x, rnd(lw["qkv_w"]), lw["qkv_b"])).reshape(N, T, 3, num_heads, hd)
    q, k, v = qkv[:, :, 0].transpose(1, 2), qkv[:, :, 1].transpose(1, 2), qkv[:, :, 2].transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) * (hd ** -0.5)
    if rnd is _ident:
        p = torch.softmax(s, dim=-1)
        o = torch.matmul(p, v)
    else:       # device form: unnormalised 16-bit P into the P.V product, fp32 row sums, one division at the end
        e = torch.exp(s - s.amax(dim=-1, keepdim=True))
        o = torch.matmul(rnd(e), v) / e.sum(dim=-1, keepdim=True)
    o = rnd(o.transpose(1, 2).reshape(N, T, D))
    return F.linear(o, rnd(lw["o_w"]), lw["o_b"])


def encoder_layer(x, lw, spec, rnd=_ident):
    D = spec.hidden_size
    h = attention(rnd(F.layer_norm(x, (D,), lw["ln1_w"], lw["ln1_b"], spec.ln_eps)), lw, spec.num_heads, rnd)
    if lw["ls1"] is not None:
        h = h * lw["ls1"]
    x = x + h
    m = rnd(F.layer_norm(x, (D,), lw["ln2_w"], lw["ln2_b"], spec.ln_eps))
    m = F.linear(rnd(activation(F.linear(m, rnd(lw["fc1_w"]), lw["fc1_b"]), spec.act)), rnd(lw["fc2_w"]), lw["fc2_b"])
    if lw["ls2"] is not None:
        m = m * lw["ls2"]
    return x + m


@torch.no_grad()
def tower_forward(cw, spec, pixels, select_layer=-2, select_feature="patch", operand_dtype=None):
    """`hidden_states[select_layer]` then feature_select (clip_encoder.py:29-37).  `operand_dtype` (torch.bfloat16 /
    torch.float16) switches on the operand-rounding emulation described at `operand_rounder`."""
    n = spec.num_layers + 1
    nrun = select_layer % n
    rnd = operand_rounder(operand_dtype)
    if operand_dtype is not None:
        cw = dict(cw, patch_w=rnd(cw["patch_w"]))
        pixels = rnd(pixels.float())
    x = embeddings(cw, spec, pixels)
    for i in range(nrun):
        x = encoder_layer(x, cw["layers"][i], spec, rnd)
    if select_feature == "patch":
        x = x[:, 1:]
    elif select_feature != "cls_patch":
        raise ValueError(f"Unexpected select feature: {select_feature}")
    return x


# ----------------------------------------------------------------------------------- DINOv2 position interpolation
def _cubic_coeffs(t, A=-0.75):
    """Keys cubic-convolution weights for taps at offsets -1, 0, +1, +2 (ATen UpSample.h `get_cubic_upsample_coefficients`)."""
    def c1(x):      # |x| <= 1
        return ((A + 2.0) * x - (A + 3.0)) * x * x + 1.0

    def c2(x):      # 1 < |x| < 2
        return ((A * x - 5.0 * A) * x + 8.0 * A) * x - 4.0 * A
    return [c2(t + 1.0), c1(t), c1(1.0 - t), c2(2.0 - t)]


def bicubic_resize_restated(table, out_size, scale):
    """Separable bicubic resampling of `table` [g, g, D] (fp32 torch or array) to [out, out, D] the way
    `F.interpolate(mode="bicubic", align_corners=False)` does it: src = scale * (dst + 0.5) - 0.5 (NOT clamped at 0 for
    cubic), taps floor(src) - 1 .. floor(src) + 2 with indices clamped to the border, A = -0.75.
    `scale` = in/out for the `size=` form, 1/scale_factor when a scale factor is given (ATen `compute_scales_value`)."""
    import numpy as np
    x = np.asarray(table, dtype=np.float64)
    g = x.shape[0]

    def axis_weights():
        W = np.zeros((out_size, g), dtype=np.float64)
        for i in range(out_size):
            src = scale * (i + 0.5) - 0.5
            ix = math.floor(src)
            t = src - ix
            for k, c in enumerate(_cubic_coeffs(t)):
                W[i, min(max(ix - 1 + k, 0), g - 1)] += c
        return W
    W = axis_weights()
    y = np.einsum("ih,hwd->iwd", W, x)
    return np.einsum("jw,iwd->ijd", W, y)


def dinov2_pos_restated(pos, pos_grid, new_grid, mode):
    """`Dinov2Embeddings.interpolate_pos_encoding` of the un-vendored `transformers`:
      mode "transformers-4.38" (4.31.0 … 4.38.2, the reference's pins pyproject.toml:16 / run.sh:5): `height + 0.1`,
           `scale_factor = (height + 0.1) / sqrt(num_positions)` -> source scale = 1 / scale_factor
      mode "transformers-5.5": `size=(g, g)` -> source scale = pos_grid / new_grid.
    pos [1, 1 + g*g, D] -> [1, 1 + new*new, D] (float64 arithmetic, returned as fp32)."""
    import numpy as np
    if new_grid == pos_grid:
        return pos
    D = pos.shape[-1]
    table = pos[0, 1:].reshape(pos_grid, pos_grid, D).double().numpy()
    scale = pos_grid / new_grid if mode == "transformers-5.5" else 1.0 / ((new_grid + 0.1) / math.sqrt(pos_grid * pos_grid))
    if mode != "transformers-5.5":
        assert math.floor(pos_grid * ((new_grid + 0.1) / pos_grid)) == new_grid
    out = bicubic_resize_restated(table, new_grid, scale).reshape(1, new_grid * new_grid, D)
    return torch.cat([pos[:, :1].float(), torch.from_numpy(np.ascontiguousarray(out)).float()], dim=1)
