"""N>1 host logic on CPU: world_size-2 gloo processes exercise the frame sharding and the variable-size
gather (hap_b200/sharding.py), and bench.py's rank-0-only reference arm."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hap_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partitions_cover_everything_once():
    for n in (1, 2, 3, 8):
        for total in (0, 1, 7, 64):
            seen = sorted(f for r in range(n) for f in sharding.frames_for_rank(total, n, r))
            assert seen == list(range(total))
        for chunks in (1, 8, 13, 64):
            bands = [sharding.chunk_band_for_rank(chunks, n, r) for r in range(n)]
            assert bands[0][0] == 0 and bands[-1][1] == chunks
            assert all(bands[i][1] == bands[i + 1][0] for i in range(n - 1))
            assert max(b - a for a, b in bands) - min(b - a for a, b in bands) <= 1


def _worker(rank, world, port, frame_count, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.frames_for_rank(frame_count, world, rank)
    stride = 64
    frames = torch.zeros((len(mine), stride), dtype=torch.uint8)
    used = torch.zeros(len(mine), dtype=torch.int64)
    for i, f in enumerate(mine):
        n = 5 + (f * 7) % 50
        frames[i, :n] = torch.arange(n, dtype=torch.uint8) + f
        used[i] = n
    got = sharding.gather_encoded_frames(frames, used, frame_count)
    ok = len(got) == frame_count
    for f, g in enumerate(got):
        n = 5 + (f * 7) % 50
        ok = ok and g.numel() == n and torch.equal(g, (torch.arange(n, dtype=torch.uint8) + f))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_variable_size_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_bench_reference_arm_prints_once_under_two_ranks():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "1"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["value"] > 0 and j["cpu_baseline"]["cores"] >= 1 and j["n_gpus"] == 2
