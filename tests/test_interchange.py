"""File-layout interchange with the reference (CPU only): tensor_{i}.pt folders, C-score feature files, CSV row."""
import os

import torch

from law_of_vision_representation_in_mllms_b200 import a_score as A
from law_of_vision_representation_in_mllms_b200 import c_score as CS
from law_of_vision_representation_in_mllms_b200 import interchange as IO


def test_tensor_folder_roundtrip_matches_reference_reader(tmp_path):
    folder = tmp_path / "mmbench" / "clip336"
    done = False
    for i in range(5):
        done = IO.save_tensor_to_folder(torch.full((3, 8), float(i)), str(folder), max_tensors=4)
    assert done and sorted(os.listdir(folder)) == [f"tensor_{i}.pt" for i in range(1, 5)]   # 5th write dropped
    ts = A.load_tensors(str(tmp_path / "mmbench"), "clip336", count=4)
    assert [t[0, 0].item() for t in ts] == [0.0, 1.0, 2.0, 3.0]
    assert A.load_tensors(str(tmp_path / "mmbench"), "missing", count=4) == []            # compute.py:25-27


def test_c_score_feature_file_layout(tmp_path):
    f = torch.arange(1 * 576 * 4, dtype=torch.float32).reshape(1, 576, 4)
    path = IO.save_c_score_feature(f, 24, 24, str(tmp_path / "features"), "cat", "2008_000001.jpg", "dino336")
    assert path.endswith(os.path.join("features", "cat", "2008_000001_dino336.pt"))
    fmap = torch.load(path)
    assert tuple(fmap.shape) == (1, 4, 24, 24)
    assert torch.equal(CS.desc_from_feature_map(fmap), f)


def test_policy_row_schema():
    row = IO.policy_row("clip336", {"mmbench_en": (0.8, 0.7)}, 15.66, perf={"mmbench_en": 64.3})
    assert row == {"model": "clip336", "mmbench_en": 64.3, "mmbench_en_mm_align224": 0.7, "mmbench_en_mm_align336": 0.8,
                   "mmbench_en_average": 0.75, "corres": 15.66}
