"""CPU oracle (TEST INFRASTRUCTURE, never imported by the product) of the image preprocessing either side of the tower:

  * `process_images` / `expand2square`                         llava/mm_utils.py:63-95
      -> HF `CLIPImageProcessor.preprocess` (transformers==4.31.0, pyproject.toml:16; un-vendored): convert RGB,
         resize shortest edge with PIL BICUBIC, centre crop, x * (1/255) in float64 -> float32, (x - mean) / std in float32
  * DINOv2 `AutoImageProcessor` (BitImageProcessor: 256 / crop 224, ImageNet mean/std) and the 336 override
      llava/model/multimodal_encoder/dinov2_encoder.py:24-26 (size = crop_size = {'width': 336, 'height': 336})
  * C-score `img.resize((S, S))` (PIL default filter = BICUBIC) + `(PILToTensor()(img) / 255.0 - 0.5) * 2`
      C_score/extract_feature.py:65-67

The resampling arithmetic lives in the third-party Pillow C library (`src/libImaging/Resample.c`, not under /root/reference):
it is restated here in numpy integer arithmetic — separable two-pass convolution, horizontal pass first, 8-bit rounding
after EACH pass, double-precision coefficients normalised and then fixed to 22 fractional bits — and pinned bit-exactly
against Pillow itself and against transformers' PIL-backend CLIP processor in tests/test_oracle_preprocess.py.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size, out_size, support=2.0, filt=_bicubic):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc: -> (xmin[out], count[out], int32 k[out, ksize])."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmins = np.zeros(out_size, np.int32)
    counts = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        xmins[xx], counts[xx] = xmin, xmax
    return xmins, counts, kk


def _pass(img, out_size, axis):
    """One 8-bit resampling pass along `axis` (0 = vertical, 1 = horizontal) of an HWC uint8 image."""
    in_size = img.shape[axis]
    xmins, counts, kk = resample_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        seg = src[xmins[xx]:xmins[xx] + counts[xx]]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[xx, :counts[xx]].astype(np.int64), seg, axes=(0, 0))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_bicubic_resize(img, new_w, new_h):
    """== np.array(PIL.Image.fromarray(img).resize((new_w, new_h), resample=BICUBIC)) for an HWC uint8 RGB image."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    if new_w != w:            # ImagingResample: a pass whose size does not change is skipped
        img = _pass(img, new_w, 1)
    if new_h != h:
        img = _pass(img, new_h, 0)
    return img


def expand2square(img, background_rgb):
    """llava/mm_utils.py:63-75 on an HWC uint8 array."""
    h, w = img.shape[:2]
    if w == h:
        return img
    side = max(w, h)
    out = np.empty((side, side, 3), np.uint8)
    out[:] = np.asarray(background_rgb, np.uint8)
    if w > h:
        top = (w - h) // 2
        out[top:top + h] = img
    else:
        left = (h - w) // 2
        out[:, left:left + w] = img
    return out


def shortest_edge_size(h, w, size):
    """transformers.image_transforms.get_resize_output_image_size(default_to_square=False) -> (new_h, new_w)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def hf_preprocess(img, size, crop, mean, std, exact=False, pad_square=False):
    """One image through the HF slow (PIL) processor pipeline -> float32 [3, crop, crop].
    exact=False: size = shortest edge (CLIP / DINOv2 defaults); exact=True: resize to (size, size) (the 336 override)."""
    if pad_square:
        img = expand2square(img, tuple(int(x * 255) for x in mean))
    h, w = img.shape[:2]
    new_h, new_w = (size, size) if exact else shortest_edge_size(h, w, size)
    img = pil_bicubic_resize(img, new_w, new_h)
    top, left = (new_h - crop) // 2, (new_w - crop) // 2
    assert top >= 0 and left >= 0, "crop larger than the resized image is outside the supported path"
    img = img[top:top + crop, left:left + crop]
    x = (img.astype(np.float64) * (1 / 255)).astype(np.float32)
    x = (x - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def c_score_pixels(img, size):
    """C_score/extract_feature.py:65-67 -> float32 [3, size, size]."""
    img = pil_bicubic_resize(img, size, size)
    x = img.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)
    return np.ascontiguousarray((x - np.float32(0.5)) * np.float32(2))
