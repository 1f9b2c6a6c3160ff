"""Image preprocessing on the GPU, mirroring the reference's CPU preprocessing bit for bit.

Reference interfaces mirrored here (names, argument meaning, outputs):
  * `process_images(images, image_processor, model_cfg)` and `expand2square`      llava/mm_utils.py:63-95
  * the HF processor objects the towers expose as `.image_processor`
    (`CLIPImageProcessor.from_pretrained`, clip_encoder.py:23; `AutoImageProcessor`, dinov2_encoder.py:24-26;
    siglip_encoder.py:24): `.image_mean`, `.image_std`, `.size`, `.crop_size`, `.preprocess(img, return_tensors='pt')['pixel_values']`
  * `img.resize((S,S))` + `(PILToTensor()(img) / 255.0 - 0.5) * 2`                 C_score/extract_feature.py:65-67

The arithmetic runs in `lovr_preprocess` (csrc/preprocess.cuh): Pillow's 8-bit two-pass BICUBIC resampling is integer
work, so pixel values equal the CPU path's exactly.  Images are accepted as PIL images, HWC uint8 numpy arrays or HWC
uint8 torch tensors; they are packed into one pinned buffer, copied once, and resampled/cropped/normalised on the
device into `[N,3,R,R]`.  No CPU fallback: without the library / device the call raises.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

OPENAI_CLIP_MEAN = [0.48145466, 0.4578275, 0.40821073]
OPENAI_CLIP_STD = [0.26862954, 0.26130258, 0.27577711]
IMAGENET_DEFAULT_MEAN = [0.485, 0.456, 0.406]
IMAGENET_DEFAULT_STD = [0.229, 0.224, 0.225]
IMAGENET_STANDARD_MEAN = [0.5, 0.5, 0.5]
IMAGENET_STANDARD_STD = [0.5, 0.5, 0.5]


def _as_hwc_uint8(image):
    """PIL image / numpy / torch -> contiguous HWC uint8 numpy array (RGB).  `convert('RGB')` like do_convert_rgb."""
    if hasattr(image, "convert") and hasattr(image, "size") and not isinstance(image, (np.ndarray, torch.Tensor)):
        image = np.asarray(image.convert("RGB"))
    elif isinstance(image, torch.Tensor):
        image = image.detach().cpu().numpy()
    image = np.asarray(image)
    if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
        raise ValueError(f"expected an RGB image as HWC uint8, got {image.dtype} {image.shape}")
    return np.ascontiguousarray(image)


def pack_images(images, staging=None, grow=None):
    """-> (pinned uint8 buffer holding the images back to back as HWC RGB, int64 byte offsets [N], int32 (H, W) [N,2]).
    `staging`: a pinned uint8 tensor to pack into when it is large enough (pinning 0.5 GB per call costs more than the copy);
    the returned buffer is then a view of it.  `grow(nbytes)`: called for a larger staging tensor when it is not."""
    arrs = [_as_hwc_uint8(im) for im in images]
    if not arrs:
        raise ValueError("no images")
    hw = np.array([[a.shape[0], a.shape[1]] for a in arrs], dtype=np.int32)
    sizes = [a.size for a in arrs]
    offsets = np.zeros(len(arrs), dtype=np.int64)
    offsets[1:] = np.cumsum(sizes[:-1])
    total = int(sum(sizes))
    if (staging is None or staging.numel() < total) and grow is not None:
        staging = grow(total)
    if staging is not None and staging.numel() >= total:
        host = staging[:total]
    else:
        host = torch.empty(total, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    hv = host.numpy()
    for a, o in zip(arrs, offsets):
        hv[o:o + a.size] = a.reshape(-1)
    return host, offsets, hw


def workspace_bytes(cfg, hw):
    L = _lib.lib()
    need = L.lovr_preprocess_workspace_bytes(C.byref(cfg), len(hw), hw.ctypes.data_as(C.POINTER(C.c_int32)))
    if need == 0:
        raise _lib.LovrError("lovr_preprocess_workspace_bytes: " + L.lovr_last_error().decode("utf-8", "replace"))
    return need


def run_packed(dev_images, offsets, hw, cfg, out_dtype=torch.float32, out=None, ws=None):
    """One `lovr_preprocess` call on images already resident on the device (packed like `pack_images`)."""
    n = len(hw)
    device = dev_images.device
    need = workspace_bytes(cfg, hw)
    with torch.cuda.device(device):
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=device)
        if out is None:
            out = torch.empty(n, 3, cfg.crop_size, cfg.crop_size, dtype=out_dtype, device=device)
        _lib.check(_lib.lib().lovr_preprocess(_lib.ptr(dev_images), offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                                              hw.ctypes.data_as(C.POINTER(C.c_int32)), n, C.byref(cfg), _lib.ptr(out),
                                              _lib.torch_dtype_code(out.dtype), _lib.ptr(ws), ws.numel(),
                                              _lib.stream_ptr(device)), "lovr_preprocess")
    return out


def _run(images, cfg, out_dtype, device):
    host, offsets, hw = pack_images(images)
    workspace_bytes(cfg, hw)                     # argument errors surface before any copy
    device = torch.device(device)
    with torch.cuda.device(device):
        dev = host.to(device, non_blocking=True)
    # dev / the workspace go back to torch's caching allocator on return; reuse is ordered after the kernels on this stream
    return run_packed(dev, offsets, hw, cfg, out_dtype)


class B200ImageProcessor:
    """Stands where the reference's HF image processor stands (`tower.image_processor`).

    size: int = shortest edge (CLIP / DINOv2 defaults, aspect ratio kept) or {'height': S, 'width': S} /
    {'shortest_edge': S} like the HF config; crop_size: int or {'height': R, 'width': R}."""

    def __init__(self, size, crop_size=None, image_mean=None, image_std=None, device="cuda"):
        if isinstance(size, dict):
            if "shortest_edge" in size:
                self._exact, s = False, int(size["shortest_edge"])
            else:
                if size["height"] != size["width"]:
                    raise ValueError("only square target sizes are on this path")
                self._exact, s = True, int(size["height"])
        else:
            self._exact, s = False, int(size)
        self._s = s
        self.size = dict(size) if isinstance(size, dict) else {"shortest_edge": s}
        if crop_size is None:
            crop_size = s
        r = int(crop_size["height"]) if isinstance(crop_size, dict) else int(crop_size)
        self.crop_size = {"height": r, "width": r}
        self.image_mean = list(image_mean if image_mean is not None else OPENAI_CLIP_MEAN)
        self.image_std = list(image_std if image_std is not None else OPENAI_CLIP_STD)
        self.device = device

    def _cfg(self, pad_square=False):
        cfg = _lib.PreprocessConfig()
        cfg.resize_mode = _lib.RESIZE_EXACT if self._exact else _lib.RESIZE_SHORTEST_EDGE
        cfg.resize_size = self._s
        cfg.crop_size = self.crop_size["height"]
        cfg.pad_square = 1 if pad_square else 0
        bg = [int(x * 255) for x in self.image_mean]      # mm_utils.py:84
        for i in range(3):
            cfg.pad_rgb[i] = bg[i]
            cfg.mean[i] = float(np.float32(self.image_mean[i]))
            cfg.std[i] = float(np.float32(self.image_std[i]))
        cfg.pixel_map = _lib.PIX_HF_RESCALE_NORMALIZE
        return cfg

    def preprocess(self, images, return_tensors="pt", pad_square=False, dtype=torch.float32):
        """-> {'pixel_values': float tensor [N,3,R,R] on the device} (the HF call returns the same values on the CPU)."""
        if return_tensors not in ("pt", None):
            raise ValueError("only return_tensors='pt' is supported")
        if not isinstance(images, (list, tuple)):
            images = [images]
        return {"pixel_values": _run(images, self._cfg(pad_square), dtype, self.device)}

    __call__ = preprocess


def process_images(images, image_processor, model_cfg):
    """llava/mm_utils.py:78-95.  `image_aspect_ratio == 'pad'`: every image is squared on the mean colour (`expand2square`
    :63-75 never materialises: the horizontal pass reads the background colour outside the pasted image) and preprocessed
    -> stacked [N,3,R,R]; otherwise only images[0] is processed (the reference's behaviour), -> [3,R,R]
    (or a list of [1,3,R,R], one per processor, when `image_processor` is a list: the dual-tower case)."""
    image_aspect_ratio = getattr(model_cfg, "image_aspect_ratio", None)
    if image_aspect_ratio == "pad":
        return image_processor.preprocess(list(images), return_tensors="pt", pad_square=True)["pixel_values"]
    if type(image_processor) is list:
        return [p.preprocess(images[0], return_tensors="pt")["pixel_values"] for p in image_processor]
    return image_processor.preprocess(images[0], return_tensors="pt")["pixel_values"][0]


def c_score_pixels(images, img_size, device="cuda", dtype=torch.float32):
    """C_score/extract_feature.py:65-67 for a list of images -> [N,3,S,S]: `img.resize((S,S))` then `(x/255 - 0.5)*2`."""
    if not isinstance(images, (list, tuple)):
        images = [images]
    cfg = _lib.PreprocessConfig()
    cfg.resize_mode = _lib.RESIZE_EXACT
    cfg.resize_size = cfg.crop_size = int(img_size)
    cfg.pad_square = 0
    cfg.pixel_map = _lib.PIX_CSCORE
    for i in range(3):
        cfg.mean[i], cfg.std[i] = 0.0, 1.0
    return _run(images, cfg, dtype, device)


def processor_for(vision_tower_name, image_size=None, device="cuda"):
    """The processor configuration each checkpoint ships (preprocessor_config.json), with the reference's 336 override for
    DINOv2 (dinov2_encoder.py:25-26) when the tower is run at 336."""
    name = vision_tower_name
    if "dinov2" in name:
        if image_size and image_size != 224:
            return B200ImageProcessor({"height": image_size, "width": image_size}, image_size, IMAGENET_DEFAULT_MEAN,
                                      IMAGENET_DEFAULT_STD, device)
        return B200ImageProcessor({"shortest_edge": 256}, 224, IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD, device)
    if "siglip" in name:
        return B200ImageProcessor({"height": 224, "width": 224}, 224, IMAGENET_STANDARD_MEAN, IMAGENET_STANDARD_STD, device)
    r = 336 if "336" in name else 224
    return B200ImageProcessor({"shortest_edge": r}, r, OPENAI_CLIP_MEAN, OPENAI_CLIP_STD, device)
