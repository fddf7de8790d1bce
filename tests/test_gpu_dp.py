"""Data-parallel parity on real GPUs (needs >= 2): tests/dp_check.py trains the tiny golden configuration for
three epochs through Trainer.train_step -- once in one process, once under torchrun with two ranks that split
every batch (sharded optimiser step, NCCL collectives captured in the CUDA graph) -- and requires the global
losses to agree within 1e-4 and the final weights within 1e-3.  Skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_training_matches_one_rank():
    script = os.path.join(HERE, "dp_check.py")
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, script], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", script], env=env, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0 and "dp2:" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
