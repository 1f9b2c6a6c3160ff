"""World-size-2 gloo runs of the data-parallel host logic (sharding + the single all-reduce per score) on CPU, with the
per-shard arithmetic supplied by the oracle (no GPU here).  The sharded result must equal the unsharded one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from law_of_vision_representation_in_mllms_b200 import c_score as CS
from law_of_vision_representation_in_mllms_b200 import distributed as D
from law_of_vision_representation_in_mllms_b200.extract import shard_range
from oracle import a_score_oracle as ao
from oracle import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cka_data():
    g = torch.Generator().manual_seed(5)
    X = torch.randn(1001, 48, generator=g) + 0.3
    Y = 0.5 * X[:, :32] + torch.randn(1001, 32, generator=g)
    return X, Y


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w = D.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    feats = synth.a_score_features(synth.A_SCORE_CASES["small"])
    lo, hi = shard_range(100, rank, world)
    s336 = [ao.a_pair(o, r_) for o, r_ in zip(feats["dino"][lo:hi], feats["clip336"][lo:hi])]
    s224 = [ao.a_pair(o, r_) for o, r_ in zip(feats["dino"][lo:hi], feats["clip224"][lo:hi])]
    a = D.a_score_reduce(s336, s224)
    # C score: 7 pairs in 3 categories, sharded; partial category sums all-reduced
    g = torch.Generator().manual_seed(0)
    pck = torch.rand(7, 4, generator=g)
    cats = [0, 1, 1, 2, 0, 2, 2]
    lo, hi = shard_range(7, rank, world)
    sums = CS.category_sums(pck[lo:hi], cats[lo:hi], 3)
    c = D.c_score_reduce(sums)
    # linear CKA: rows sharded, ONE all-reduce of the partial-sum buffer (Gram sums + column sums + row count)
    X, Y = _cka_data()
    lo, hi = shard_range(X.shape[0], rank, world)
    buf = D.all_reduce_sum(ao.cka_gram_buffer(X[lo:hi], Y[lo:hi]))
    cka = ao.cka_from_gram_buffer(buf, X.shape[1], Y.shape[1])
    torch.save(dict(a=a, c=c, cka=cka, cka_rows=buf[-1].item()), os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    for n in (0, 1, 7, 100, 10000):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_world2_gloo_scores_equal_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    feats = synth.a_score_features(synth.A_SCORE_CASES["small"])
    want_a, _, _ = ao.a_score(feats["dino"], feats["clip336"], feats["clip224"])
    g = torch.Generator().manual_seed(0)
    pck = torch.rand(7, 4, generator=g)
    cats = np.array([0, 1, 1, 2, 0, 2, 2])
    means = [pck[cats == k, 0].double().mean().item() for k in range(3)]
    want_c = float(np.average(means, weights=[2, 2, 3]))
    X, Y = _cka_data()
    want_cka = ao.linear_cka(X, Y)
    for r in res:
        assert r["cka_rows"] == 1001 and abs(r["cka"] - want_cka) < 1e-12
        assert abs(r["a"] - want_a) < 1e-12
        assert abs(r["c"][0] - want_c) < 1e-12
