"""Tower specifications for the in-scope `--vision_tower` keys.

Mirrors what the reference gets from HuggingFace configs when it builds a tower
(reference: llava/model/multimodal_encoder/clip_encoder.py:22-27,
dinov2_encoder.py:21-30, siglip_encoder.py:20-27) and the arithmetic facts listed
in SURVEY.md Appendix A.  Only the fields the scoring hot path needs are kept.
"""
from dataclasses import dataclass, replace

# activation codes shared with csrc/lovr_b200.h (LOVR_ACT_*)
ACT_NONE, ACT_QUICK_GELU, ACT_GELU_ERF, ACT_GELU_TANH = 0, 1, 2, 3
# family codes
FAMILY_CLIP, FAMILY_DINOV2, FAMILY_SIGLIP = "clip", "dinov2", "siglip"


@dataclass(frozen=True)
class TowerSpec:
    name: str
    family: str
    hidden_size: int
    intermediate_size: int
    num_layers: int
    num_heads: int
    image_size: int          # native resolution (== config.image_size in HF)
    patch_size: int
    act: int
    ln_eps: float
    has_cls: bool
    patch_bias: bool
    pre_ln: bool             # CLIP `pre_layrnorm` (sic)
    layerscale: bool         # DINOv2 LayerScale
    fixed_resolution: bool   # CLIP/SigLIP raise on a size mismatch; DINOv2 interpolates
    pos_grid: int            # side of the stored position grid

    @property
    def head_dim(self):
        return self.hidden_size // self.num_heads

    def tokens(self, image_size=None):
        s = self.image_size if image_size is None else image_size
        g = s // self.patch_size
        return g * g + (1 if self.has_cls else 0)

    def at_resolution(self, image_size):
        """A spec for the same weights run at another input size (DINOv2 only)."""
        if image_size == self.image_size:
            return self
        if self.fixed_resolution:
            raise ValueError(
                f"Input image size ({image_size}*{image_size}) doesn't match model "
                f"({self.image_size}*{self.image_size}).")
        return replace(self, image_size=image_size)


_CLIP_L = dict(family=FAMILY_CLIP, hidden_size=1024, intermediate_size=4096, num_layers=24,
               num_heads=16, patch_size=14, ln_eps=1e-5, has_cls=True, patch_bias=False,
               pre_ln=True, layerscale=False, fixed_resolution=True)

TOWER_SPECS = {
    "openai/clip-vit-large-patch14-336": TowerSpec(
        name="openai/clip-vit-large-patch14-336", image_size=336, pos_grid=24, act=ACT_QUICK_GELU, **_CLIP_L),
    "openai/clip-vit-large-patch14": TowerSpec(
        name="openai/clip-vit-large-patch14", image_size=224, pos_grid=16, act=ACT_QUICK_GELU, **_CLIP_L),
    "laion/CLIP-ViT-L-14-laion2B-s32B-b82K": TowerSpec(
        name="laion/CLIP-ViT-L-14-laion2B-s32B-b82K", image_size=224, pos_grid=16, act=ACT_GELU_ERF, **_CLIP_L),
    "facebook/dinov2-large": TowerSpec(
        name="facebook/dinov2-large", family=FAMILY_DINOV2, hidden_size=1024, intermediate_size=4096,
        num_layers=24, num_heads=16, image_size=518, patch_size=14, act=ACT_GELU_ERF, ln_eps=1e-6,
        has_cls=True, patch_bias=True, pre_ln=False, layerscale=True, fixed_resolution=False, pos_grid=37),
    "google/siglip-base-patch16-224": TowerSpec(
        name="google/siglip-base-patch16-224", family=FAMILY_SIGLIP, hidden_size=768, intermediate_size=3072,
        num_layers=12, num_heads=12, image_size=224, patch_size=16, act=ACT_GELU_TANH, ln_eps=1e-6,
        has_cls=False, patch_bias=True, pre_ln=False, layerscale=False, fixed_resolution=True, pos_grid=14),
}


def tiny_spec(spec: TowerSpec, num_layers: int) -> TowerSpec:
    """Same widths, fewer layers: used by the fast parity tests (kernels are width-specialised,
    not depth-specialised)."""
    return replace(spec, num_layers=num_layers)


def layers_to_run(spec: TowerSpec, select_layer: int) -> int:
    """`hidden_states[select_layer]` has L+1 entries; entry i is the output of encoder layer i
    (entry 0 = embeddings).  Reference: clip_encoder.py:29-30."""
    n = spec.num_layers + 1
    if not -n <= select_layer < n:
        raise IndexError("tuple index out of range")
    return select_layer % n
