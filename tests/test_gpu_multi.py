"""Hardware multi-GPU parity (needs >= 2 GPUs; skipped on a one-GPU box): under torchrun with NCCL, the sharded A score /
C score / linear CKA equal the unsharded values on the same seeded data, and the dataset-level extraction driver writes every
image exactly once across ranks.  `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_sharded_scores_equal_unsharded_over_nccl(tmp_path):
    world = min(torch.cuda.device_count(), 8)
    rng = np.random.default_rng(0)
    for cls in ("a", "b", "c"):
        os.makedirs(tmp_path / "images" / cls)
        for i in range(5):
            Image.fromarray(rng.integers(0, 256, (80 + 7 * i, 120, 3), dtype=np.uint8)).save(tmp_path / "images" / cls / f"im{i}.png")
    env = dict(os.environ, LOVR_MULTI_OUT=str(tmp_path))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(ROOT, "tests", "_multi_worker.py")], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    print(json.dumps(res))
    assert res["world"] == world and res["ok"], res
    assert res["a_score"]["abs_diff"] < 1e-9 and res["c_score"]["abs_diff"] < 1e-9 and res["linear_cka"]["abs_diff"] < 1e-5
    assert res["files_after_barrier"] == 15                  # every image exactly once, visible to rank 0 after the barrier
    assert res["extract"]["total"] == 15 and res["extract"]["written"] == res["extract"]["shard"]
