"""ORACLE (test infrastructure — never imported by the product path).

Loads the REFERENCE's own tower wrapper classes by file path and drives them on the CPU:
  llava/model/multimodal_encoder/{clip,dinov2,siglip}_encoder.py
either where they lie (`/root/reference`, this container only) or from `oracle/_ref/` — the git-ignored directory
`oracle/build_ref.py` fills with those few files so that they travel to the GPU box, where `bench.py --impl reference` and
the `cpu_baseline` leg time the reference's own code (`kind: "reference"`).  Only the `from_pretrained` entry points (network /
checkpoints) are replaced: by constructors that build the HF model from a config and load a seeded synthetic state dict.
"""
import importlib.util
import os
import types

import torch

from law_of_vision_representation_in_mllms_b200 import specs as S

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ENCODERS = "/root/reference/llava/model/multimodal_encoder"
LOCAL_REF = os.path.join(HERE, "_ref")


def encoder_dir():
    """The reference checkout when present (build container), else the files oracle/build_ref.py put under oracle/_ref."""
    if os.path.isdir(REF_ENCODERS):
        return REF_ENCODERS
    if os.path.exists(os.path.join(LOCAL_REF, "clip_encoder.py")):
        return LOCAL_REF
    return None


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Args:
    def __init__(self, layer=-2, feature="patch"):
        self.mm_vision_select_layer = layer
        self.mm_vision_select_feature = feature


def _patch_dinov2_interp_438(emb):
    """Replace transformers 5.5's `Dinov2Embeddings.interpolate_pos_encoding` by the form of the versions the reference
    pins (4.31.0 pyproject.toml:16 / 4.38.2 run.sh:5; modeling_dinov2.py `interpolate_pos_encoding`, restated — that source
    is not in this image): `height + 0.1`, `F.interpolate(scale_factor=height / sqrt(num_positions))`."""
    import math

    def interpolate_pos_encoding(embeddings, height, width):
        num_patches = embeddings.shape[1] - 1
        num_positions = emb.position_embeddings.shape[1] - 1
        if num_patches == num_positions and height == width:
            return emb.position_embeddings
        class_pos_embed = emb.position_embeddings[:, 0]
        patch_pos_embed = emb.position_embeddings[:, 1:]
        dim = embeddings.shape[-1]
        h = height // emb.config.patch_size + 0.1
        w = width // emb.config.patch_size + 0.1
        side = int(math.sqrt(num_positions))
        patch_pos_embed = patch_pos_embed.reshape(1, side, side, dim).permute(0, 3, 1, 2)
        target_dtype = patch_pos_embed.dtype
        patch_pos_embed = torch.nn.functional.interpolate(
            patch_pos_embed.to(torch.float32), scale_factor=(float(h / math.sqrt(num_positions)), float(w / math.sqrt(num_positions))),
            mode="bicubic", align_corners=False).to(dtype=target_dtype)
        if int(h) != patch_pos_embed.shape[-2] or int(w) != patch_pos_embed.shape[-1]:
            raise ValueError("Width or height does not match with the interpolated position embeddings")
        patch_pos_embed = patch_pos_embed.permute(0, 2, 3, 1).reshape(1, -1, dim)
        return torch.cat((class_pos_embed.unsqueeze(0), patch_pos_embed), dim=1)
    emb.interpolate_pos_encoding = interpolate_pos_encoding


def reference_tower(spec, sd, select_layer, select_feature, pos_interp="transformers-5.5", enc=None):
    """Build the reference wrapper for `spec`, with from_pretrained stubbed to (config, sd)."""
    import transformers as T
    enc = enc or encoder_dir()
    if enc is None:
        raise FileNotFoundError("reference tower wrappers: neither /root/reference nor oracle/_ref is present")
    if spec.family == S.FAMILY_CLIP:
        mod = _load(os.path.join(enc, "clip_encoder.py"), "ref_clip_encoder")
        cfg = T.CLIPVisionConfig(hidden_size=spec.hidden_size, intermediate_size=spec.intermediate_size,
                                 num_hidden_layers=spec.num_layers, num_attention_heads=spec.num_heads,
                                 image_size=spec.image_size, patch_size=spec.patch_size,
                                 hidden_act={S.ACT_QUICK_GELU: "quick_gelu", S.ACT_GELU_ERF: "gelu"}[spec.act],
                                 layer_norm_eps=spec.ln_eps)

        def build(*a, **k):
            m = T.CLIPVisionModel(cfg)
            m.load_state_dict(sd, strict=True)
            return m.eval()
        mod.CLIPVisionModel = types.SimpleNamespace(from_pretrained=build)
        mod.CLIPImageProcessor = types.SimpleNamespace(from_pretrained=lambda *a, **k: None)
        return mod.CLIPVisionTower(spec.name, _Args(select_layer, select_feature))
    if spec.family == S.FAMILY_DINOV2:
        mod = _load(os.path.join(enc, "dinov2_encoder.py"), "ref_dinov2_encoder")
        cfg = T.Dinov2Config(hidden_size=spec.hidden_size, num_hidden_layers=spec.num_layers,
                             num_attention_heads=spec.num_heads, mlp_ratio=4, image_size=518,
                             patch_size=spec.patch_size, hidden_act="gelu", layer_norm_eps=spec.ln_eps)

        def build(*a, **k):
            m = T.Dinov2Model(cfg)
            m.load_state_dict(sd, strict=True)
            if pos_interp == "transformers-4.38":
                _patch_dinov2_interp_438(m.embeddings)
            return m.eval()
        mod.AutoModel = types.SimpleNamespace(from_pretrained=build)
        mod.AutoConfig = types.SimpleNamespace(from_pretrained=lambda *a, **k: cfg)
        mod.AutoImageProcessor = types.SimpleNamespace(from_pretrained=lambda *a, **k: types.SimpleNamespace())
        return mod.DinoV2VisionTower(spec.name, _Args(select_layer, select_feature))
    if spec.family == S.FAMILY_SIGLIP:
        # siglip_encoder.py hard-codes dtype fp16 / device "cuda" (:59-64) and under transformers 5.5
        # `.vision_model(..., output_hidden_states=True).hidden_states` is None, so the
        # wrapper cannot run on CPU.  Drive HF SiglipVisionModel directly and apply the wrapper's
        # feature_select (:30-38) with its hard-coded 'cls_patch'.
        cfg = T.SiglipVisionConfig(hidden_size=spec.hidden_size, intermediate_size=spec.intermediate_size,
                                   num_hidden_layers=spec.num_layers, num_attention_heads=spec.num_heads,
                                   image_size=spec.image_size, patch_size=spec.patch_size,
                                   hidden_act="gelu_pytorch_tanh", layer_norm_eps=spec.ln_eps)
        m = T.SiglipVisionModel(cfg)
        missing = m.load_state_dict(sd, strict=False)
        assert all(k.startswith("vision_model.head.") for k in missing.missing_keys), missing
        m.eval()

        class _Sig(torch.nn.Module):
            def forward(self, images):
                hs = m(images.float(), output_hidden_states=True).hidden_states
                return hs[select_layer].to(images.dtype)
        return _Sig()
    raise ValueError(spec.family)


