"""CPU: host-side logic of the product layer that needs no device — the adaptive-flip index logic against the fixtures
produced by the reference's own utils_geoware.py, image packing and the per-checkpoint processor configurations."""
import os

import numpy as np
import torch

from law_of_vision_representation_in_mllms_b200 import c_score as CS
from law_of_vision_representation_in_mllms_b200 import preprocess as PP
from oracle import synth


def test_flip_index_logic_matches_reference_fixtures(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "c_score_flip.pt"))
    for cname, cfg in synth.FLIP_CASES.items():
        _, _, _, kps1, kps2 = synth.flip_case(cfg)
        rec = gold[cname]
        vis = kps1[:, 2] * kps2[:, 2] > 0
        k1f = CS.flip_keypoints(kps1, 840, CS.permute_indices(cfg["permute"], vis))
        assert torch.equal(k1f, rec["kps1_flip"])
        assert torch.equal(torch.as_tensor(CS.kpts_to_patch_idx(k1f, cfg["P"])), rec["idx_flip"])
        # optimized_kps_1_to_2 fed with the reference's own predictions and distances reproduces its choice
        final = CS.optimized_kps_1_to_2(rec["kps"], rec["kps_flip"], kps1, kps2, rec["flip_dist"], rec["original_dist"], vis,
                                        cfg["permute"])
        assert torch.equal(final, rec["final"])


def test_permute_indices_examples_of_the_reference_docstring():
    # utils_geoware.py:151-189 documents these two cases
    assert CS.permute_indices([1, 2, [0, 3]], [True, True, True, False]) == [0, 1, 2, 3]
    assert CS.permute_indices([1, 2, [0, 3]], None) == [3, 1, 2, 0]
    assert CS.permute_indices([[0, 1], [2, 3, 5, 6], [4, 7]]) == [1, 0, 3, 5, 7, 6, 2, 4]


def test_pack_images_layout_and_type_checks():
    a = synth.synthetic_rgb_image(5, 7, 1)
    b = synth.synthetic_rgb_image(3, 2, 2)
    host, offsets, hw = PP.pack_images([a, torch.from_numpy(b)])
    assert host.dtype == torch.uint8 and host.numel() == a.size + b.size
    assert offsets.tolist() == [0, a.size] and hw.tolist() == [[5, 7], [3, 2]]
    assert np.array_equal(host.numpy()[: a.size].reshape(a.shape), a)
    assert np.array_equal(host.numpy()[a.size:].reshape(b.shape), b)
    for bad in (np.zeros((4, 4), np.uint8), np.zeros((4, 4, 3), np.float32), np.zeros((4, 4, 4), np.uint8)):
        try:
            PP.pack_images([bad])
        except ValueError:
            continue
        raise AssertionError("expected ValueError")


def test_processor_configurations_follow_the_checkpoints():
    clip = PP.processor_for("openai/clip-vit-large-patch14-336")
    assert clip.size == {"shortest_edge": 336} and clip.crop_size == {"height": 336, "width": 336}
    assert clip.image_mean == PP.OPENAI_CLIP_MEAN and clip.image_std == PP.OPENAI_CLIP_STD
    assert PP.processor_for("laion/CLIP-ViT-L-14-laion2B-s32B-b82K").crop_size["height"] == 224
    dino = PP.processor_for("facebook/dinov2-large")
    assert dino.size == {"shortest_edge": 256} and dino.crop_size["height"] == 224 and dino.image_mean == PP.IMAGENET_DEFAULT_MEAN
    dino336 = PP.processor_for("facebook/dinov2-large", 336)          # dinov2_encoder.py:25-26 override
    assert dino336.size == {"height": 336, "width": 336} and dino336.crop_size["height"] == 336
    sig = PP.processor_for("google/siglip-base-patch16-224")
    assert sig.size == {"height": 224, "width": 224} and sig.image_mean == [0.5, 0.5, 0.5]
    cfg = clip._cfg(pad_square=True)
    assert (cfg.resize_mode, cfg.resize_size, cfg.crop_size, cfg.pad_square) == (0, 336, 336, 1)
    assert list(cfg.pad_rgb)[:3] == [int(m * 255) for m in PP.OPENAI_CLIP_MEAN]      # mm_utils.py:84


def test_pil_inputs_are_converted_to_rgb_like_do_convert_rgb():
    from PIL import Image
    gray = Image.fromarray(np.arange(48, dtype=np.uint8).reshape(6, 8), mode="L")
    rgba = Image.fromarray(np.full((4, 5, 4), 200, np.uint8), mode="RGBA")
    host, offsets, hw = PP.pack_images([gray, rgba])
    assert hw.tolist() == [[6, 8], [4, 5]] and host.numel() == 6 * 8 * 3 + 4 * 5 * 3
    g = host.numpy()[: 6 * 8 * 3].reshape(6, 8, 3)
    assert np.array_equal(g, np.asarray(gray.convert("RGB")))
    assert np.array_equal(host.numpy()[6 * 8 * 3:].reshape(4, 5, 3), np.asarray(rgba.convert("RGB")))


def test_extractor_chunk_schedule_covers_every_image_once():
    from law_of_vision_representation_in_mllms_b200.extract import chunk_schedule
    for n in (1, 2, 3, 5, 100, 128, 129, 255, 256, 257, 300, 511, 512, 513, 1000, 10000):
        for c in (1, 4, 64, 128, 256):
            ch = chunk_schedule(n, c)
            assert ch[0][0] == 0 and ch[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
            assert all(0 < hi - lo <= c for lo, hi in ch)
    assert [hi - lo for lo, hi in chunk_schedule(512, 256)] == [128, 256, 128]
