"""Data-parallel plumbing: one process per GPU under torchrun, `torch.distributed` (NCCL over NVLink/NVSwitch on the
GPU box, gloo in the CPU tests).  The reference's DP pattern (llava/feature/extract.py:193-231,259) shards images with a
DistributedSampler, runs independent per-rank forwards and synchronises only with a final barrier; scoring needs one
sum over ranks of small partial-sum buffers:

    extraction : contiguous image shard per rank, NO collective
    A score    : [sum_336, sum_224, n_images]                      -> one all-reduce(sum)
    C score    : [n_categories, 4] category partial sums             -> one all-reduce(sum)
    linear CKA : fp64 uncentred Gram + column sums + row count in ONE buffer (~25 MB @ D=1024) -> one all-reduce(sum)
"""
import os

import torch
import torch.distributed as dist

from .extract import shard_range  # noqa: F401  (re-export)


def init_from_env(backend=None, device=None):
    """Initialise the default process group from torchrun's env (RANK / WORLD_SIZE / MASTER_*).  Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def all_reduce_sum(t):
    """In-place sum over ranks (no-op for a single process).  Issued on the tensor's current stream, right after the
    kernels that produced it — no host synchronisation in between."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def a_score_reduce(per_image_336, per_image_224, device=None):
    """compute.py:75-81 across ranks: each rank holds the per-image a(o,r) of its image shard."""
    t = torch.tensor([float(sum(per_image_336)), float(sum(per_image_224)), float(len(per_image_336))],
                     dtype=torch.float64, device=device)
    all_reduce_sum(t)
    s336, s224, n = t.tolist()
    return (s336 / n + s224 / n) / 2


def c_score_reduce(category_sums):
    from .c_score import weighted_c_score
    return weighted_c_score(all_reduce_sum(category_sums.clone()))


def cka_reduce(gram, Dx, Dy):
    """One all-reduce: the per-shard Gram sums, column sums AND row count live in the same fp64 buffer (cka.cka_gram)."""
    from .cka import cka_from_gram
    all_reduce_sum(gram)
    return cka_from_gram(gram, None, Dx, Dy)
