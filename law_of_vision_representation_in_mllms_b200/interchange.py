"""On-disk interchange with the reference: the same file layouts, so features written here are
scored by the reference's scripts and vice versa, and scores drop into `policy/ablations_t.csv`.

  tensor_{i}.pt   [T, D] one file per benchmark image, i = 1..100     llava/model/llava_arch.py:229-248 (writer hook),
                                                                       A_score/compute.py:18-28 (reader)
  <features>/<category>/<image>_<MODEL>.pt   [1, C, h, w]              C_score/extract_feature.py:110-130 (writer),
                                                                       C_score/pck_train.py:31-38 (reader)
  ablations_t.csv row: `{bench}_mm_align336`, `{bench}_mm_align224`, `{bench}_average`, `corres`   policy/fit.py:33-41
"""
import os

import torch

from .extract import to_c_score_layout


def save_tensor_to_folder(tensor, folder_path, max_tensors=100, exit_when_full=False):
    """llava_arch.py:229-248: write `tensor_{count+1}.pt` while fewer than `max_tensors` files exist.  The reference
    calls `exit()` once the folder is full; here that is opt-in (`exit_when_full=True`) and the function returns True
    when the folder has reached `max_tensors` so a caller can stop its loop."""
    os.makedirs(folder_path, exist_ok=True)
    tensor_files = [f for f in os.listdir(folder_path) if f.endswith(".pt")]
    tensor_count = len(tensor_files)
    if tensor_count < max_tensors:
        tensor_filename = os.path.join(folder_path, f"tensor_{tensor_count + 1}.pt")
        torch.save(tensor.detach().cpu(), tensor_filename)
    full = tensor_count + 1 >= max_tensors
    if full and exit_when_full:
        raise SystemExit(f"Tensor count has reached {max_tensors}. Exiting the program.")
    return full


def c_score_feature_path(features_root, category, image_name, model_tag):
    """pck_train.py:33-37: img_path.replace('JPEGImages','features').replace('.jpg','') + f'_{MODEL}.pt'."""
    stem = os.path.splitext(os.path.basename(image_name))[0]
    return os.path.join(features_root, category, f"{stem}_{model_tag}.pt")


def save_c_score_feature(features, h, w, features_root, category, image_name, model_tag):
    """extract_feature.py:82,125-129: tower output [1,P,C] -> [1,C,h,w] -> torch.save under <root>/<category>/."""
    path = c_score_feature_path(features_root, category, image_name, model_tag)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save(to_c_score_layout(features, h, w).detach().cpu(), path)
    return path


def policy_row(model_name, a_scores, c_score, perf=None):
    """One row in the schema of policy/ablations_t.csv that policy/fit.py reads (:33-41): for every benchmark
    `{bench}_mm_align336`, `{bench}_mm_align224`, `{bench}_average` (A score) and the global `corres` (C score, PCK@0.10 in
    percent).  `a_scores[bench] = (avg336, avg224)` as returned per reference by a_score.a_score; `perf[bench]` optional."""
    row = {"model": model_name}
    for bench, (a336, a224) in a_scores.items():
        if perf and bench in perf:
            row[bench] = perf[bench]
        row[f"{bench}_mm_align224"] = a224
        row[f"{bench}_mm_align336"] = a336
        row[f"{bench}_average"] = (a336 + a224) / 2
    row["corres"] = c_score
    return row
