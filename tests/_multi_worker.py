"""Worker of tests/test_gpu_multi.py (launched under torchrun, one process per GPU): the sharded A score / C score / linear CKA
with their single NCCL all-reduce against the unsharded values on the same seeded data (bench.multi_gpu_parity), plus the
dataset-level extraction driver on a shared directory (each rank writes its shard, final barrier)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    with bench._stdout_to_stderr():
        dist.init_process_group("nccl", device_id=dev)
    res = bench.multi_gpu_parity(dev, rank, world)
    # dataset-level driver: every rank walks the same tree, extracts its contiguous shard, barrier at the end
    out_dir = os.environ.get("LOVR_MULTI_OUT")
    if out_dir:
        import types
        from law_of_vision_representation_in_mllms_b200 import extract as X, specs as S, towers
        from law_of_vision_representation_in_mllms_b200.weights import synthetic_state_dict
        spec = S.tiny_spec(S.TOWER_SPECS["openai/clip-vit-large-patch14"], 2)
        args = types.SimpleNamespace(mm_vision_select_layer=-1, mm_vision_select_feature="patch", lovr_spec=spec,
                                     lovr_state_dict=synthetic_state_dict(spec, seed=3, style="hf_init"), lovr_max_batch=4,
                                     lovr_device=dev)
        ex = X.FeatureExtractor(towers.CLIPVisionTower("openai/clip-vit-large-patch14", args), chunk=4)
        st = X.extract_folder(ex, os.path.join(out_dir, "images"), os.path.join(out_dir, "features"), layout="tokens", batch=4)
        res["extract"] = st
        n_files = sum(len(f) for _, _, f in os.walk(os.path.join(out_dir, "features")))      # after the driver's barrier
        res["files_after_barrier"] = n_files
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
