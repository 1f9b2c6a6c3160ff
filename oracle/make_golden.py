"""Generate tests/golden/*.pt by running the REFERENCE's own code (from /root/reference, read-only)
on seeded synthetic inputs.  Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py [vit] [ascore] [cscore] [preprocess] [flip]

What runs from the reference, unmodified:
  * llava/model/multimodal_encoder/{clip,dinov2,siglip}_encoder.py — loaded by file path; only the
    `from_pretrained` entry points (which need network/checkpoints) are replaced by constructors
    that build the HF model from a config and load our seeded synthetic state dict.
  * A_score/compute.py — exec'd with only its two hard-coded path/list lines (:7, :10) substituted.
  * C_score/utils/utils_correspondence.py — imported by file path.
  * llava/mm_utils.py (process_images, expand2square) — loaded by file path with a stub `llava` package holding the
    reference's own llava/constants.py; driven with transformers' PIL-backend processors (the descendants of the 4.31 slow
    processors the reference pins) and Pillow itself; the three C-score lines of C_score/extract_feature.py:65-67 are
    executed verbatim (the script itself needs a GPU and checkpoints).
Inputs are regenerated from seeds by tests (law_of_vision_representation_in_mllms_b200.weights /
oracle.synth), so fixtures hold only outputs.
"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")

from law_of_vision_representation_in_mllms_b200 import specs as S          # noqa: E402
from law_of_vision_representation_in_mllms_b200.weights import (           # noqa: E402
    synthetic_state_dict, synthetic_images)
from oracle import synth                                                   # noqa: E402


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Args:
    def __init__(self, layer=-2, feature="patch"):
        self.mm_vision_select_layer = layer
        self.mm_vision_select_feature = feature


# ----------------------------------------------------------------------------------- ViT towers
def case_pos_interp(name):
    """Fixtures come from transformers 5.5 (this image) unless the case name ends in `_t438` (the reference's pinned form)."""
    return "transformers-4.38" if name.endswith("_t438") else "transformers-5.5"


from oracle.ref_loader import reference_tower as _reference_tower   # noqa: E402


VIT_CASES = [
    # name, tower key, layers kept (None = full depth), image size, n images, select_layer, select_feature
    ("clip224_full", "openai/clip-vit-large-patch14", None, 224, 2, -2, "patch"),
    ("clip336_full", "openai/clip-vit-large-patch14-336", None, 336, 1, -2, "patch"),
    ("clip336_l3", "openai/clip-vit-large-patch14-336", 3, 336, 2, -2, "patch"),
    ("clip224_l3_clspatch", "openai/clip-vit-large-patch14", 3, 224, 2, -1, "cls_patch"),
    ("openclip224_l3", "laion/CLIP-ViT-L-14-laion2B-s32B-b82K", 3, 224, 2, -2, "patch"),
    ("dinov2_224_l3", "facebook/dinov2-large", 3, 224, 2, -2, "patch"),
    ("dinov2_336_l3", "facebook/dinov2-large", 3, 336, 2, -2, "patch"),
    ("dinov2_336_full", "facebook/dinov2-large", None, 336, 1, -2, "patch"),
    ("siglip224_l3", "google/siglip-base-patch16-224", 3, 224, 2, -2, "cls_patch"),
    ("siglip224_full", "google/siglip-base-patch16-224", None, 224, 2, -2, "cls_patch"),
    # "_hf" suffix: HF-init-style weights (weights.synthetic_state_dict(style="hf_init")), the benchmark weights
    ("clip336_full_hf", "openai/clip-vit-large-patch14-336", None, 336, 1, -2, "patch"),
    ("clip224_full_hf", "openai/clip-vit-large-patch14", None, 224, 2, -2, "patch"),
    ("dinov2_336_full_hf", "facebook/dinov2-large", None, 336, 1, -2, "patch"),
    # "_t438": DINOv2 position interpolation of the reference's pinned transformers (scale_factor + 0.1 form)
    ("dinov2_336_l3_t438", "facebook/dinov2-large", 3, 336, 2, -2, "patch"),
    ("dinov2_224_l3_t438", "facebook/dinov2-large", 3, 224, 2, -2, "patch"),
    # round 2: every family at full depth on the benchmark (HF-init) weights ...
    ("siglip224_full_hf", "google/siglip-base-patch16-224", None, 224, 2, -2, "cls_patch"),
    ("openclip224_full_hf", "laion/CLIP-ViT-L-14-laion2B-s32B-b82K", None, 224, 1, -2, "patch"),
    # ... "_raw": the same draw NOT pre-rounded to bf16 (weight quantisation inside the measured error) ...
    ("clip336_full_raw", "openai/clip-vit-large-patch14-336", None, 336, 1, -2, "patch"),
    ("dinov2_336_full_raw", "facebook/dinov2-large", None, 336, 1, -2, "patch"),
    ("siglip224_full_raw", "google/siglip-base-patch16-224", None, 224, 2, -2, "cls_patch"),
    # ... "_outlier": massive-activation channels (|x| > 100 in the residual), wide LayerNorm gains, sharp softmax, fp32 weights
    ("clip336_full_outlier", "openai/clip-vit-large-patch14-336", None, 336, 1, -2, "patch"),
    ("dinov2_336_full_outlier", "facebook/dinov2-large", None, 336, 1, -2, "patch"),
]
VIT_TOKENS = 12   # token rows kept per image


def case_style(name):
    for suffix, style in (("_hf", "hf_init"), ("_raw", "hf_init_raw"), ("_outlier", "outlier")):
        if name.endswith(suffix):
            return style
    return "stress"


def vit_case_inputs(case):
    name, key, nl, size, n, sl, sf = case
    spec = S.TOWER_SPECS[key]
    if nl is not None:
        spec = S.tiny_spec(spec, nl)
    seed = 100 + sum(map(ord, name)) % 1000
    sd = synthetic_state_dict(spec, seed=seed, style=case_style(name))
    imgs = synthetic_images(n, size, seed=seed + 1)
    return spec, sd, imgs


def token_rows(T):
    g = torch.Generator().manual_seed(T)
    rows = torch.randperm(T, generator=g)[:VIT_TOKENS - 2].tolist()
    return sorted(set(rows + [0, T - 1]))


def make_vit(only=None):
    for case in VIT_CASES:
        name, key, nl, size, n, sl, sf = case
        if only and name not in only:
            continue
        spec, sd, imgs = vit_case_inputs(case)
        tower = _reference_tower(spec, sd, sl, sf, case_pos_interp(name))
        with torch.no_grad():
            out = tower(imgs)                       # [N,P,D] fp32
            out_list = None
            if name == "clip336_l3":                # the list-of-3D-tensors calling convention
                out_list = tower([im for im in imgs])
        rows = token_rows(out.shape[1])
        ref_bf16_rel = ref_fp16_rel = None
        if spec.family != S.FAMILY_SIGLIP:
            # the reference's own reduced-precision paths: tower cast to bf16 (llava/feature/extract.py:228) and to
            # fp16 (the eval path, llava/model/builder.py:143)
            with torch.no_grad():
                tower.vision_tower.to(torch.bfloat16)
                o16 = tower(imgs[:1].to(torch.bfloat16)).float()
                ref_bf16_rel = ((o16 - out[:1]).norm() / out[:1].norm()).item()
                try:
                    tower.vision_tower.to(torch.float16)
                    o16 = tower(imgs[:1].to(torch.float16)).float()
                    ref_fp16_rel = ((o16 - out[:1]).norm() / out[:1].norm()).item()
                except Exception as ex:      # an op without a CPU half kernel: recorded as absent
                    print("  (reference fp16 path not runnable on this CPU:", str(ex)[:80], ")")
                tower.vision_tower.to(torch.float32)
        rec = dict(case=case, shape=tuple(out.shape), rows=rows, values=out[:, rows].clone(),
                   fro=out.flatten(1).norm(dim=1), mean=out.flatten(1).mean(dim=1),
                   abs_mean=out.abs().flatten(1).mean(dim=1), ref_bf16_rel=ref_bf16_rel, ref_fp16_rel=ref_fp16_rel,
                   abs_max=out.abs().max().item(), weight_style=case_style(name), pos_interp=case_pos_interp(name))
        if out_list is not None:
            rec["list_equal_batch"] = bool(all(torch.allclose(o[0], out[i], atol=1e-5) for i, o in enumerate(out_list)))
            rec["list_shapes"] = [tuple(o.shape) for o in out_list]
        torch.save(rec, os.path.join(GOLD, f"vit_{name}.pt"))
        print("vit", name, tuple(out.shape), "fro", rec["fro"].tolist(), "ref bf16 rel", ref_bf16_rel, "ref fp16 rel",
              ref_fp16_rel, flush=True)


# ----------------------------------------------------------------------------------- A score
def make_ascore():
    src = open(os.path.join(REF, "A_score/compute.py")).read()
    out = {}
    for cname, cfg in synth.A_SCORE_CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            feats = synth.a_score_features(cfg)
            for sub, tensors in feats.items():
                os.makedirs(os.path.join(tmp, sub))
                for i, t in enumerate(tensors):
                    torch.save(t, os.path.join(tmp, sub, f"tensor_{i + 1}.pt"))
            res = {}
            for sub in feats:                       # one exec per subfolder to read the per-image lists
                s = src.replace("base_folder = '/any/path/mmbench'", f"base_folder = {tmp!r}")
                s = s.replace("subfolders = ['clip336', 'clip224', 'dino', 'dit', 'imsd', 'openclip', 'sd1.5', "
                              "'sd2.1', 'sd3', 'sdxl']", f"subfolders = [{sub!r}]")
                assert s != src
                ns = {"__name__": "ref_a_score"}
                exec(compile(s, "A_score/compute.py", "exec"), ns)
                res[sub] = dict(result=ns["results"][sub], per_image_336=list(ns["cosine_similarities_336"]),
                                per_image_224=list(ns["cosine_similarities_224"]))
                print("ascore", cname, sub, res[sub]["result"])
            out[cname] = res
    torch.save(out, os.path.join(GOLD, "a_score.pt"))


# ----------------------------------------------------------------------------------- C score
def make_cscore():
    uc = _load(os.path.join(REF, "C_score/utils/utils_correspondence.py"), "ref_utils_correspondence")
    out = {}
    for cname, cfg in synth.C_SCORE_CASES.items():
        d1, d2, kps = synth.c_score_pair(cfg)
        P = cfg["P"]
        # normalize_feats restated from C_score/pck_train.py:24-29 (module not importable offline)
        nf = lambda f: f / (torch.linalg.norm(f, dim=-1)[:, :, None] + 1e-10)
        d1n, d2n = nf(d1), nf(d2)
        rec = {}
        for soft in (True, False):
            args = types.SimpleNamespace(ANNO_SIZE=840, SOFT_EVAL=soft, SOFT_EVAL_WINDOW=cfg["window"])
            idx = uc.kpts_to_patch_idx(args, kps, P)
            k12 = uc.calculate_keypoint_transformation(args, d1n, d2n, torch.as_tensor(idx, dtype=torch.long), P)
            rec["soft" if soft else "hard"] = k12.clone()
            rec["patch_idx"] = torch.as_tensor(idx)
        sim = torch.matmul(d1n, d2n.permute(0, 2, 1))[0]
        rec["flow"] = uc.get_flow(sim, cfg["window"], P).reshape(-1, 2).clone()
        rec["argmax"] = sim.argmax(dim=-1)
        out[cname] = rec
        print("cscore", cname, rec["soft"][:2].tolist())
    torch.save(out, os.path.join(GOLD, "c_score.pt"))


# ----------------------------------------------------------------------------------- adaptive flip
def make_flip():
    """The reference's pck_train.py:103-126 block (ADAPT_FLIP + MUTUAL_NN) driven with its own functions:
    utils_correspondence.{kpts_to_patch_idx, calculate_keypoint_transformation, get_distance_mutual_nn} and
    utils_geoware.{permute_indices, flip_keypoints, optimized_kps_1_to_2}.  Those call `.cuda()` on their tensors:
    here (no GPU) `Tensor.cuda` is made the identity for the duration of the run."""
    uc = _load(os.path.join(REF, "C_score/utils/utils_correspondence.py"), "ref_utils_correspondence")
    ug = _load(os.path.join(REF, "C_score/utils/utils_geoware.py"), "ref_utils_geoware")
    nf = lambda f: f / (torch.linalg.norm(f, dim=-1)[:, :, None] + 1e-10)      # pck_train.py:24-29
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    out = {}
    try:
        for cname, cfg in synth.FLIP_CASES.items():
            d1, d1f, d2, kps1, kps2 = synth.flip_case(cfg)
            P = cfg["P"]
            args = types.SimpleNamespace(ANNO_SIZE=840, SOFT_EVAL=True, SOFT_EVAL_WINDOW=5)
            img1_desc, img1_desc_flip, img2_desc = nf(d1), nf(d1f), nf(d2)
            vis = kps1[:, 2] * kps2[:, 2] > 0                                                        # :100
            idx = uc.kpts_to_patch_idx(args, kps1, P)                                                # :106
            kps_1_to_2 = uc.calculate_keypoint_transformation(args, img1_desc, img2_desc, torch.as_tensor(idx).long(), P)
            permute_list = cfg["permute"]
            kps1_flip = ug.flip_keypoints(kps1, args.ANNO_SIZE, ug.permute_indices(permute_list, vis))   # :114
            idx_flip = uc.kpts_to_patch_idx(args, kps1_flip, P)                                      # :115
            kps_1_to_2_flip = uc.calculate_keypoint_transformation(args, img1_desc_flip, img2_desc,
                                                                   torch.as_tensor(idx_flip).long(), P)
            original_dist = uc.get_distance_mutual_nn(img1_desc, img2_desc)                          # :120
            flip_dist = uc.get_distance_mutual_nn(img1_desc_flip, img2_desc)                         # :121
            final = ug.optimized_kps_1_to_2(args, kps_1_to_2.clone(), kps_1_to_2_flip, kps1, kps2, flip_dist,
                                            original_dist, vis, permute_list)                         # :126
            out[cname] = dict(original_dist=float(original_dist), flip_dist=float(flip_dist), kps=kps_1_to_2.clone(),
                              kps_flip=kps_1_to_2_flip.clone(), final=final.clone(), idx_flip=torch.as_tensor(idx_flip),
                              kps1_flip=kps1_flip.clone())
            print("flip", cname, out[cname]["original_dist"], out[cname]["flip_dist"])
    finally:
        torch.Tensor.cuda = real_cuda
    torch.save(out, os.path.join(GOLD, "c_score_flip.pt"))


# ----------------------------------------------------------------------------------- mm_projector
def make_projector():
    """llava/model/multimodal_projector/builder.py:34-59 `build_vision_projector` itself builds the module (loaded by path
    inside a stub package: its one relative import, perceiver_helpers, is only needed by the out-of-scope 'perceiver'
    branch and is stubbed when it cannot be imported); seeded weights are loaded through the module's OWN state_dict keys,
    so the key naming ('0.weight' / bare 'weight' for 'linear') is pinned as well.  Fixture: fp32 outputs + the key list."""
    pkg = types.ModuleType("ref_mm_projector")
    pkg.__path__ = [os.path.join(REF, "llava/model/multimodal_projector")]
    sys.modules["ref_mm_projector"] = pkg
    try:
        sys.modules["ref_mm_projector.perceiver_helpers"] = _load(
            os.path.join(REF, "llava/model/multimodal_projector/perceiver_helpers.py"), "ref_mm_projector.perceiver_helpers")
    except Exception as ex:
        print("  (perceiver_helpers not importable here, stubbed:", str(ex)[:80], ")")
        sys.modules["ref_mm_projector.perceiver_helpers"] = types.SimpleNamespace(PerceiverResampler=None)
    b = _load(os.path.join(REF, "llava/model/multimodal_projector/builder.py"), "ref_mm_projector.builder")
    out = {}
    for name in synth.PROJECTOR_CASES:
        ptype, din, dh, linears, x = synth.projector_inputs(name)
        cfg = types.SimpleNamespace(mm_projector_type=ptype, mm_hidden_size=din, hidden_size=dh)
        mod = b.build_vision_projector(cfg)
        lin = [m for m in mod.modules() if isinstance(m, torch.nn.Linear)]
        assert len(lin) == len(linears)
        with torch.no_grad():
            for m, (w, bias) in zip(lin, linears):
                m.weight.copy_(w)
                m.bias.copy_(bias)
            y = mod(x)
        out[name] = dict(keys=list(mod.state_dict().keys()), out=y.clone())
        print("projector", name, out[name]["keys"], tuple(y.shape), y.norm().item())
    torch.save(out, os.path.join(GOLD, "projector.pt"))


# ----------------------------------------------------------------------------------- preprocessing
def make_preprocess():
    """Outputs are bit-exact quantities: the fixture stores their SHA-256 (+ shape and a few samples), not the tensors."""
    import hashlib
    from PIL import Image
    from torchvision.transforms import PILToTensor
    from transformers.models.bit.image_processing_pil_bit import BitImageProcessorPil
    from transformers.models.clip.image_processing_pil_clip import CLIPImageProcessorPil
    from transformers.models.siglip.image_processing_pil_siglip import SiglipImageProcessorPil
    pkg = types.ModuleType("llava")
    pkg.__path__ = []
    sys.modules["llava"] = pkg
    sys.modules["llava.constants"] = _load(os.path.join(REF, "llava/constants.py"), "llava.constants")
    mm = _load(os.path.join(REF, "llava/mm_utils.py"), "ref_mm_utils")
    imgs = [Image.fromarray(a) for a in synth.preprocess_images()]
    out = {}
    for cname, cfg in synth.PREPROCESS_CASES.items():
        sel = [imgs[i] for i in cfg["images"]]
        if cfg["kind"] == "cscore":
            res = []
            for img in sel:
                img_size = cfg["size"]
                img = img.convert('RGB')                                   # extract_feature.py:65
                img = img.resize((img_size, img_size))                     # :66
                img_tensor = (PILToTensor()(img) / 255.0 - 0.5) * 2        # :67
                res.append(img_tensor)
            t = torch.stack(res)
        else:
            mean, std = synth.PREPROCESS_NORMS[cfg["norm"]]
            size = {"height": cfg["size"], "width": cfg["size"]} if cfg["exact"] else {"shortest_edge": cfg["size"]}
            crop = {"height": cfg["crop"], "width": cfg["crop"]}
            cls = {"clip": CLIPImageProcessorPil, "imagenet": BitImageProcessorPil, "half": SiglipImageProcessorPil}[cfg["norm"]]
            kw = dict(size=size, image_mean=mean, image_std=std, resample=3)
            if cls is not SiglipImageProcessorPil:
                kw["crop_size"] = crop
            proc = cls(**kw)
            mcfg = types.SimpleNamespace(image_aspect_ratio="pad" if cfg.get("pad") else None)
            t = mm.process_images(sel, proc, mcfg)                         # the reference's own function
            t = torch.as_tensor(t)
            if t.dim() == 3:
                t = t[None]
        t = t.contiguous().float()
        out[cname] = dict(shape=tuple(t.shape), sha256=hashlib.sha256(t.numpy().tobytes()).hexdigest(),
                          sample=t[0, :, ::37, ::41].clone())
        print("preprocess", cname, out[cname]["shape"], out[cname]["sha256"][:16])
    torch.save(out, os.path.join(GOLD, "preprocess.pt"))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    what = sys.argv[1:] or ["vit", "ascore", "cscore", "preprocess", "flip", "projector"]
    if any(w.startswith("vit:") for w in what):      # `vit:<case>` regenerates single ViT cases
        what.append("vit")
    torch.set_num_threads(os.cpu_count())
    if "projector" in what:
        make_projector()
    if "preprocess" in what:
        make_preprocess()
    if "flip" in what:
        make_flip()
    if "ascore" in what:
        make_ascore()
    if "cscore" in what:
        make_cscore()
    if "vit" in what:
        make_vit(only=[w[4:] for w in what if w.startswith("vit:")] or None)
