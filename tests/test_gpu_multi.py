"""Needs at least two GPUs (skipped otherwise): the N>1 paths over NCCL, one process per GPU (tests/multi_gpu_worker.py),
and bench.py's 16k_stream workload at a small frame size."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _torchrun(n, script_args, env=None, timeout=600):
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + script_args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=dict(os.environ, **(env or {})))


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_nccl_gatherv_allgather_and_banded_frame():
    n = min(_gpus(), 4)
    p = _torchrun(n, [os.path.join(ROOT, "tests", "multi_gpu_worker.py")])
    assert p.returncode == 0 and "multi-gpu worker ok" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_stream_workload_small():
    n = min(_gpus(), 2)
    p = _torchrun(n, [os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--workload", "16k_stream", "--steps", "3", "--warmup", "3"],
                  env={"HAPB200_STREAM_SIZE": "2048"})
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == n and line["delivered_frame_verified"] is True and line["nvlink_bytes_per_step"] > 0 and line["fps"] > 0
