"""Multi-GPU tests (skipped on a single-GPU box): the row-sharded DLRM forward over NVLink peer
memory and the replica-parallel bench launch, both under torchrun with one rank per GPU."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _torchrun(n, script_args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", *script_args]
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_dlrm_matches_unsharded():
    n = min(torch.cuda.device_count(), 8)
    r = _torchrun(n, ["tests/dist_sharded_check.py"])
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_data_parallel_training_matches_single_gpu_twin():
    r = _torchrun(2, ["tests/dist_train_check.py"])
    assert r.returncode == 0 and "TRAIN_DP_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_bench_replicas_two_gpus():
    r = _torchrun(2, ["bench.py", "--gpus", "2", "--steps", "5", "--warmup", "3", "--batch", "8192"], timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["value"] > 0 and line["scaling"] == "weak"
