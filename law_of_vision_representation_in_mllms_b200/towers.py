"""Host-side mirror of the reference's vision-tower interface, running on liblovr_b200.so.

Mirrors, name for name:
  * `CLIPVisionTower`, `DinoV2VisionTower`, `SigLipVisionTower`
      (llava/model/multimodal_encoder/clip_encoder.py:7-78, dinov2_encoder.py:8-83, siglip_encoder.py:7-79):
      ctor `(vision_tower, args, delay_load=False)`, `load_model()`, `feature_select` semantics via
      `select_layer` / `select_feature`, `forward(images)` taking a 4-D tensor or a list of 3-D tensors and
      returning features in the caller's dtype, properties `dtype, device, config, hidden_size,
      num_patches, dummy_feature`, attributes `is_loaded, vision_tower_name, image_processor`.
  * `build_vision_tower`, `build_dinov2_vision_tower`, `build_siglip_vision_tower`
      (llava/model/multimodal_encoder/builder.py:8-28) and `build_function_mapping`
      (llava/model/llava_arch.py:29-40), plus the `--vision_tower` selector grammar of
      llava_arch.py:49-84 in `build_towers_from_selector`.
  * `encode_images` (llava/model/llava_arch.py:260-286) including the dual-tower `a.b` channel concat,
      which is fused here: each tower writes its slice of one [N,P,D1+D2] tensor.

The arithmetic (HF CLIPVisionModel / Dinov2Model / SiglipVisionModel forward) runs in hand-written
sm_100a kernels; nothing here falls back to PyTorch ops for compute.
"""
import ctypes as C
import os
import types

import torch
import torch.nn as nn

from . import _lib
from .specs import (FAMILY_CLIP, FAMILY_DINOV2, FAMILY_SIGLIP, TOWER_SPECS, TowerSpec, layers_to_run)
from .weights import DEFAULT_POS_INTERP, canonical_weights, synthetic_state_dict

DEFAULT_MAX_BATCH = int(os.environ.get("LOVR_MAX_BATCH", "128"))   # images per native call; 8.3 MB of workspace each at ViT-L/336


class _NativeTower:
    """One (weights, resolution) instance of the native tower: packed device weights + workspace + handle."""

    def __init__(self, spec: TowerSpec, cw, device, max_batch, operand_dtype=torch.bfloat16, ln_fold=0):
        self.spec = spec
        self.ln_fold = int(ln_fold)       # bit mask: 1 = LayerNorm-1 into qkv, 2 = LayerNorm-2 into fc1
        self.operand_dtype = operand_dtype
        self.device = torch.device(device)
        self.max_batch = max_batch
        self._keep = []   # device tensors referenced by raw pointers
        D, P = spec.hidden_size, spec.patch_size
        kpad = (3 * P * P + 63) // 64 * 64

        def dev(t, dtype=torch.float32):
            if t is None:
                return None
            d = t.detach().to(self.device, dtype).contiguous()
            self._keep.append(d)
            return d

        pw = torch.zeros(D, kpad, dtype=torch.float32)
        pw[:, : 3 * P * P] = cw["patch_w"].reshape(D, -1)
        self.cfg = _lib.TowerConfig(
            hidden_size=D, intermediate_size=spec.intermediate_size, num_layers=spec.num_layers,
            num_heads=spec.num_heads, image_size=spec.image_size, patch_size=P, act=spec.act,
            has_cls=int(spec.has_cls), pre_ln=int(spec.pre_ln), layerscale=int(spec.layerscale),
            ln_eps=spec.ln_eps, patch_k_padded=kpad, operand_dtype=_lib.torch_dtype_code(operand_dtype),
            ln_fold=self.ln_fold)
        layers = (_lib.LayerWeights * spec.num_layers)()
        bf = operand_dtype      # 16-bit tensor-core operand format of the weight matrices
        for i, lw in enumerate(cw["layers"]):
            L = layers[i]
            folded = {}
            # LayerNorm folded into the qkv / fc1 GEMMs (include/lovr_b200.h `ln_fold`): W diag(gamma) in the operand
            # format, its row sums and W beta + b come from the library's own ingest kernel, from the fp32 weights
            if self.ln_fold & 1:
                folded.update(self._fold(lw["qkv_w"], lw["ln1_w"], lw["ln1_b"], lw["qkv_b"], "qkv", bf))
            if self.ln_fold & 2:
                folded.update(self._fold(lw["fc1_w"], lw["ln2_w"], lw["ln2_b"], lw["fc1_b"], "fc1", bf))
            for name, dt in (("ln1_w", None), ("ln1_b", None), ("qkv_w", bf), ("qkv_b", None), ("o_w", bf),
                             ("o_b", None), ("ls1", None), ("ln2_w", None), ("ln2_b", None), ("fc1_w", bf),
                             ("fc1_b", None), ("fc2_w", bf), ("fc2_b", None), ("ls2", None), ("qkv_cw", None), ("fc1_cw", None)):
                if name in folded:
                    t = folded[name]
                else:
                    t = dev(lw.get(name), dt or torch.float32)
                setattr(L, name, t.data_ptr() if t is not None else None)
        self._layers = layers
        tw = _lib.TowerWeights()
        for name, t in (("patch_w", dev(pw, bf)), ("patch_b", dev(cw["patch_b"])), ("cls", dev(cw["cls"])),
                        ("pos", dev(cw["pos"])), ("pre_ln_w", dev(cw["pre_ln_w"])), ("pre_ln_b", dev(cw["pre_ln_b"]))):
            setattr(tw, name, t.data_ptr() if t is not None else None)
        tw.layers_host = C.cast(layers, C.POINTER(_lib.LayerWeights))
        self._tw = tw
        L = _lib.lib()
        ws_bytes = L.lovr_tower_workspace_bytes(C.byref(self.cfg), max_batch)
        self.workspace = torch.empty(ws_bytes + 1024, dtype=torch.uint8, device=self.device)
        base = (self.workspace.data_ptr() + 1023) // 1024 * 1024
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.lovr_tower_create(C.byref(self.handle), C.byref(self.cfg), C.byref(tw), C.c_void_p(base),
                                           ws_bytes, max_batch), "lovr_tower_create")

    def _fold(self, W, gamma, beta, bias, prefix, op_dtype):
        """lovr_ln_fold_weights on the device -> {prefix_w: 16-bit W diag(gamma), prefix_cw: its row sums, prefix_b: W beta + b}."""
        L = _lib.lib()
        N, K = W.shape
        with torch.cuda.device(self.device):
            w32 = W.to(self.device, torch.float32).contiguous()
            g, b = gamma.to(self.device, torch.float32).contiguous(), beta.to(self.device, torch.float32).contiguous()
            bi = bias.to(self.device, torch.float32).contiguous() if bias is not None else None
            w16 = torch.empty(N, K, dtype=op_dtype, device=self.device)
            cw_ = torch.empty(N, dtype=torch.float32, device=self.device)
            bw = torch.empty(N, dtype=torch.float32, device=self.device)
            _lib.check(L.lovr_ln_fold_weights(_lib.ptr(w32), _lib.ptr(g), _lib.ptr(b), _lib.ptr(bi), N, K,
                                              _lib.torch_dtype_code(op_dtype), _lib.ptr(w16), _lib.ptr(cw_), _lib.ptr(bw),
                                              _lib.stream_ptr(self.device)), "lovr_ln_fold_weights")
            torch.cuda.current_stream(self.device).synchronize()      # w32 / g / b / bi are freed on return
        self._keep += [w16, cw_, bw]
        return {prefix + "_w": w16, prefix + "_cw": cw_, prefix + "_b": bw}

    def __del__(self):
        h = getattr(self, "handle", None)
        if h and h.value:
            _lib.lib().lovr_tower_destroy(h)
            self.handle = C.c_void_p()

    def forward_into(self, pixels, out, n_run, drop_cls, col_off=0, attn_impl=_lib.ATTN_AUTO):
        """pixels [n,3,R,R] on device; out [n,P,Dtot] on device; writes out[:, :, col_off:col_off+D]."""
        L = _lib.lib()
        n = pixels.shape[0]
        pixels = pixels.contiguous()
        assert out.is_contiguous() and out.shape[0] == n
        pd, od = _lib.torch_dtype_code(pixels.dtype), _lib.torch_dtype_code(out.dtype)
        with torch.cuda.device(self.device):
            st = _lib.stream_ptr(self.device)
            for s in range(0, n, self.max_batch):
                e = min(n, s + self.max_batch)
                _lib.check(L.lovr_tower_forward(self.handle, _lib.ptr(pixels[s:e]), pd, e - s, n_run, int(drop_cls),
                                                _lib.ptr(out[s:e]), od, out.shape[-1], col_off, attn_impl, st),
                           "lovr_tower_forward")


class _B200VisionTower(nn.Module):
    """Shared implementation; subclasses only fix the family quirks the reference wrappers have."""

    _forced_select_feature = None

    def __init__(self, vision_tower, args, delay_load=False):
        super().__init__()
        self.is_loaded = False
        self.vision_tower_name = vision_tower
        self.select_layer = args.mm_vision_select_layer
        self.select_feature = self._forced_select_feature or getattr(args, "mm_vision_select_feature", "patch")
        self._args = args
        self._spec = self._resolve_spec(vision_tower, args)
        self._native = {}          # image_size -> _NativeTower
        self._state_dict = None
        self.image_processor = None
        self._attn_impl = _lib.ATTN_AUTO   # single-pass online softmax, P through TMEM, 3 or 2 pipelines by token count
        if not delay_load:
            self.load_model()

    # ---- weights ------------------------------------------------------------------------------
    @staticmethod
    def _resolve_spec(name, args):
        spec = getattr(args, "lovr_spec", None)
        if spec is not None:
            return spec
        if name in TOWER_SPECS:
            return TOWER_SPECS[name]
        raise KeyError(name)

    def load_model(self, state_dict=None, device=None):
        """Reference: `*.from_pretrained(self.vision_tower_name)` (clip_encoder.py:22-27).  Offline there are
        no checkpoints, so weights come from, in order: the `state_dict` argument, `args.lovr_state_dict`,
        `$LOVR_WEIGHTS_DIR/<name with / -> __>.pt` (an HF state_dict), else seeded synthetic weights."""
        if state_dict is None:
            state_dict = getattr(self._args, "lovr_state_dict", None)
        self.weights_source = "state_dict"
        if state_dict is None:
            wdir = os.environ.get("LOVR_WEIGHTS_DIR")
            path = os.path.join(wdir, self.vision_tower_name.replace("/", "__") + ".pt") if wdir else None
            if path and os.path.exists(path):
                state_dict = torch.load(path, map_location="cpu")
                self.weights_source = path
            else:
                style = getattr(self._args, "lovr_weight_style", "hf_init")
                state_dict = synthetic_state_dict(self._spec, seed=getattr(self._args, "lovr_seed", 0), style=style)
                self.weights_source = f"synthetic:{style}"
        self._state_dict = state_dict
        self._device = torch.device(device if device is not None else getattr(self._args, "lovr_device", "cuda"))
        self._max_batch = int(getattr(self._args, "lovr_max_batch", DEFAULT_MAX_BATCH))
        od = getattr(self._args, "lovr_operand_dtype", os.environ.get("LOVR_OPERAND_DTYPE", "bf16"))
        self._operand_dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "f16": torch.float16,
                               torch.bfloat16: torch.bfloat16, torch.float16: torch.float16}[od]
        # DINOv2 position-table interpolation form: the reference's pinned transformers (4.31/4.38, default) or this image's 5.5
        self._pos_interp = getattr(self._args, "lovr_pos_interp", os.environ.get("LOVR_DINOV2_POS_INTERP", DEFAULT_POS_INTERP))
        # LayerNorm folding (lovr_ln_fold / LOVR_LN_FOLD, bit mask: 1 = LayerNorm-1 into qkv, 2 = LayerNorm-2 into fc1): a folded
        # LayerNorm has no kernel and costs one pass less over the residual stream.  Folding both was measured 3.4 % SLOWER end
        # to end (2387 vs 2470 images/s): the K=1024 out-projection is bound by its residual epilogue, which producing
        # LayerNorm-2's inputs makes heavier
        self._ln_fold = int(getattr(self._args, "lovr_ln_fold", os.environ.get("LOVR_LN_FOLD", "0")))
        self._native.clear()
        self._instance(self._default_image_size())
        self.image_processor = self._make_image_processor()
        self.is_loaded = True

    def _default_image_size(self):
        return getattr(self._args, "lovr_image_size", None) or (
            self._spec.image_size if self._spec.fixed_resolution else 224)

    def _instance(self, image_size):
        inst = self._native.get(image_size)
        if inst is None:
            spec_r = self._spec.at_resolution(image_size)     # raises like HF for fixed-resolution towers
            cw = canonical_weights(self._spec, self._state_dict, image_size=image_size, pos_interp=self._pos_interp)
            inst = _NativeTower(spec_r, cw, self._device, self._max_batch, self._operand_dtype, self._ln_fold)
            self._native[image_size] = inst
        return inst

    def _make_image_processor(self):
        """Reference: `CLIPImageProcessor.from_pretrained` (clip_encoder.py:23) / `AutoImageProcessor` (dinov2_encoder.py:24,
        siglip_encoder.py:24).  Here: the GPU processor with the checkpoint's preprocessor_config values."""
        from .preprocess import processor_for
        return processor_for(self.vision_tower_name, self._default_image_size(), self._device)

    # ---- forward --------------------------------------------------------------------------------
    def _select(self):
        n_run = layers_to_run(self._spec, self.select_layer)
        if self.select_feature == "patch":
            drop = True
        elif self.select_feature == "cls_patch":
            drop = False
        else:
            raise ValueError(f"Unexpected select feature: {self.select_feature}")
        return n_run, drop

    def out_tokens(self, image_size):
        n_run, drop = self._select()
        return self._spec.at_resolution(image_size).tokens() - (1 if drop else 0)

    def forward_into(self, images, out, col_off=0):
        """Write features for a 4-D batch into out[:, :, col_off:col_off+hidden] (fused dual-tower concat)."""
        n_run, drop = self._select()
        if images.shape[-1] != images.shape[-2]:
            raise ValueError("square inputs only")
        inst = self._instance(int(images.shape[-1]))
        inst.forward_into(images.to(self._device), out, n_run, drop, col_off, self._attn_impl)

    @torch.no_grad()
    def forward(self, images):
        if type(images) is list:
            image_features = []
            for image in images:
                image_features.append(self._forward_batch(image.unsqueeze(0)))
            return image_features
        return self._forward_batch(images)

    def _forward_batch(self, images):
        if images.dim() != 4 or images.shape[1] != 3:
            raise ValueError(f"expected [N,3,H,W] pixels, got {tuple(images.shape)}")
        out = torch.empty(images.shape[0], self.out_tokens(int(images.shape[-1])), self.hidden_size,
                          dtype=images.dtype, device=self._device)
        self.forward_into(images, out, 0)
        return out

    # ---- properties (clip_encoder.py:53-78) ---------------------------------------------------------
    @property
    def dummy_feature(self):
        return torch.zeros(1, self.hidden_size, device=self.device, dtype=self.dtype)

    @property
    def dtype(self):
        """16-bit format of the tensor-core operands: bf16 (the reference's extraction dtype, extract.py:228) or
        fp16 (the reference's eval dtype for towers, llava/model/builder.py:143); accumulation/residual are fp32."""
        return getattr(self, "_operand_dtype", torch.bfloat16)

    @property
    def device(self):
        return self._device if self.is_loaded else torch.device("cpu")

    @property
    def config(self):
        s = self._spec
        return types.SimpleNamespace(hidden_size=s.hidden_size, image_size=s.image_size, patch_size=s.patch_size,
                                     num_hidden_layers=s.num_layers, num_attention_heads=s.num_heads,
                                     intermediate_size=s.intermediate_size)

    @property
    def hidden_size(self):
        return self._spec.hidden_size

    @property
    def num_patches(self):
        return (self.config.image_size // self.config.patch_size) ** 2


class CLIPVisionTower(_B200VisionTower):
    """llava/model/multimodal_encoder/clip_encoder.py:7 (openai CLIP and laion OpenCLIP ViT-L/14)."""


class DinoV2VisionTower(_B200VisionTower):
    """llava/model/multimodal_encoder/dinov2_encoder.py:8.  `num_patches` keeps the reference's
    config.image_size=518 arithmetic (1369) — wrong but unused there too (dinov2_encoder.py:81-83)."""


class SigLipVisionTower(_B200VisionTower):
    """llava/model/multimodal_encoder/siglip_encoder.py:7 — select_feature hard-coded to 'cls_patch' (:15),
    hidden_size 768 (:74-75)."""
    _forced_select_feature = "cls_patch"


# ---- builders (llava/model/multimodal_encoder/builder.py:8-28) --------------------------------------
def _tower_name(cfg):
    return getattr(cfg, "mm_vision_tower", getattr(cfg, "vision_tower", None))


def build_vision_tower(vision_tower_cfg, **kwargs):
    vision_tower = _tower_name(vision_tower_cfg)
    is_absolute_path_exists = os.path.exists(vision_tower)
    if is_absolute_path_exists or vision_tower.startswith("openai") or vision_tower.startswith("laion"):
        return CLIPVisionTower(vision_tower, args=vision_tower_cfg, **kwargs)
    raise ValueError(f"Unknown vision tower: {vision_tower}")


def build_dinov2_vision_tower(vision_tower_cfg, **kwargs):
    return DinoV2VisionTower(_tower_name(vision_tower_cfg), args=vision_tower_cfg, **kwargs)


def build_siglip_vision_tower(vision_tower_cfg, **kwargs):
    return SigLipVisionTower(_tower_name(vision_tower_cfg), args=vision_tower_cfg, **kwargs)


def build_feature(vision_tower_cfg, **kwargs):
    return "feature"


def _out_of_scope(vision_tower_cfg, **kwargs):
    raise NotImplementedError(
        f"{_tower_name(vision_tower_cfg)}: diffusion towers (llava/model/multimodal_encoder/diffLVLM) are outside "
        "the scoring hot path this package implements (SURVEY.md §2 row 6)")


# llava/model/llava_arch.py:29-40 — same keys; diffusion entries are present but out of scope here
build_function_mapping = {
    "openai/clip-vit-large-patch14-336": build_vision_tower,
    "google/siglip-base-patch16-224": build_siglip_vision_tower,
    "laion/CLIP-ViT-L-14-laion2B-s32B-b82K": build_vision_tower,
    "stabilityai/stable-diffusion-2-1": _out_of_scope,
    "runwayml/stable-diffusion-v1-5": _out_of_scope,
    "lambdalabs/sd-image-variations-diffusers": _out_of_scope,
    "facebook/dinov2-large": build_dinov2_vision_tower,
    "stabilityai/stable-diffusion-xl-base-1.0": _out_of_scope,
    "feature": build_feature,
    "facebook/DiT-XL-2-512": _out_of_scope,
    "stabilityai/stable-diffusion-3-medium-diffusers": _out_of_scope,
    "openai/clip-vit-large-patch14": build_vision_tower,
}


def build_towers_from_selector(config):
    """The `--vision_tower` grammar of LlavaMetaModel.__init__ (llava_arch.py:49-84): a plain name builds one
    tower; `a.b` builds an nn.ModuleList whose features are concatenated on the channel dim; `a,b` builds the
    'MoF' list; an entry containing 'noise' is a None slot; a name containing 'feature' selects precomputed
    features.  Unknown names raise KeyError exactly as the dict lookup in the reference does."""
    sel = config.mm_vision_tower
    if "," not in sel and "." not in sel:
        if "feature" in sel:
            return build_function_mapping["feature"](config)
        return build_function_mapping[sel](config)
    sep = "," if "," in sel else "."
    towers = nn.ModuleList([])
    ori = config.mm_vision_tower
    try:
        for v in sel.split(sep):
            config.mm_vision_tower = v
            if "noise" in v:
                towers.append(None)
            else:
                towers.append(build_function_mapping[v](config))
    finally:
        config.mm_vision_tower = ori
    return towers


@torch.no_grad()
def encode_images(vision_tower, images, mm_projector=None):
    """LlavaMetaForCausalLM.encode_images (llava_arch.py:260-286).  Single tower + tensor: proj(tower(images)).
    ModuleList + list of per-tower pixel tensors: proj(cat([v(images[i])], -1)); the concat is fused — every
    tower writes its channel slice of one output tensor."""
    if type(images) is not list:
        feats = vision_tower(images)
    else:
        if not isinstance(vision_tower, nn.ModuleList):
            raise TypeError("a list of pixel tensors needs an nn.ModuleList of towers (selector 'a.b')")
        towers = list(vision_tower)
        n = images[0].shape[0]
        toks = {v.out_tokens(int(images[i].shape[-1])) for i, v in enumerate(towers)}
        if len(toks) != 1:
            raise RuntimeError(f"Sizes of tensors must match except in dimension 2; token counts {sorted(toks)}")
        widths = [v.hidden_size for v in towers]
        feats = torch.empty(n, toks.pop(), sum(widths), dtype=images[0].dtype, device=towers[0].device)
        off = 0
        for i, v in enumerate(towers):
            v.forward_into(images[i], feats, off)
            off += widths[i]
    if mm_projector is not None:
        feats = mm_projector(feats)
    return feats
