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


# ---- one frame, bands of whole chunks: the container splice is host logic, checked with the CPU codecs

def _band_textures(n_bands, blocks_per_band, seed):
    """YCoCg-DXT5-sized payloads per band: picture-like (repeats), flat, and noise (-> raw fallback inside a band)"""
    import numpy as np
    rng = np.random.default_rng(seed)
    out = []
    for b in range(n_bands):
        kind = b % 3
        if kind == 0:
            base = rng.integers(0, 256, 16 * 64, dtype=np.uint8)
            t = np.tile(base, blocks_per_band // 64 + 1)[: 16 * blocks_per_band].copy()
            t[:: 97] ^= 0x5A
        elif kind == 1:
            t = np.full(16 * blocks_per_band, 0x33, np.uint8)
        else:
            t = rng.integers(0, 256, 16 * blocks_per_band, dtype=np.uint8)
        out.append(t.tobytes())
    return out


def test_banded_frame_assembly_decodes_in_oracle_reference_and_library():
    import hap_b200
    import oracles
    from hap_b200.abi import HapCompressorNone, HapCompressorSnappy, HapTextureFormat_YCoCg_DXT5 as YC
    lib = hap_b200.load()
    bands = _band_textures(5, 960, 1)
    chunk_counts = [3, 1, 2, 4, 1]
    for codec in filter(None, (oracles.oracle_abi(), oracles.ref_abi())):
        frames = []
        for i, (t, k) in enumerate(zip(bands, chunk_counts)):
            r, f = codec.encode([t], [YC], [HapCompressorNone if i == 3 else HapCompressorSnappy], [k])
            assert r == 0
            frames.append(f)
        whole = sharding.assemble_banded_frame(frames)
        want = b"".join(bands)
        # band 2 is noise (its encoder falls back to a verbatim section -> one raw chunk), band 3 was stored verbatim
        assert lib.texture_count(whole) == (0, 1) and lib.texture_format(whole, 0) == (0, YC)
        assert lib.chunk_count(whole, 0) == (0, 3 + 1 + 1 + 1 + 1)
        for dec in filter(None, (oracles.oracle_abi(), oracles.ref_abi())):
            r, tex, fmt, _ = dec.decode(whole, 0, len(want))
            assert (r, fmt) == (0, YC) and tex == want
    import pytest
    with pytest.raises(ValueError):
        sharding.assemble_banded_frame([])
    with pytest.raises(ValueError):
        sharding.assemble_banded_frame([b"\x05\x00\x00"])


def _band_worker(rank, world, port, q):
    import oracles
    from hap_b200.abi import HapCompressorSnappy, HapTextureFormat_YCoCg_DXT5 as YC
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chunks = 5
    bands = _band_textures(world, 1200, 7)
    first, last = sharding.chunk_band_for_rank(chunks, world, rank)
    # (the test's stand-in for the per-rank GPU encode: the CPU oracle encodes this rank's band)
    r, f = oracles.oracle_abi().encode([bands[rank]], [YC], [HapCompressorSnappy], [last - first])
    buf = torch.zeros(len(f) + 100, dtype=torch.uint8)
    buf[: len(f)] = torch.frombuffer(bytearray(f), dtype=torch.uint8)
    got = sharding.gather_band_frames(buf, len(f), dst=0)
    ok = r == 0
    if rank == 0:
        whole = sharding.assemble_banded_frame(got)
        rr, tex, fmt, _ = oracles.oracle_abi().decode(whole, 0, sum(len(b) for b in bands))
        ok = ok and rr == 0 and tex == b"".join(bands) and oracles.oracle_abi().chunk_count(whole, 0) == (0, chunks)
    else:
        ok = ok and got is None
    q.put((rank, ok))
    dist.destroy_process_group()


def test_band_gather_and_assembly_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_band_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _gatherv_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, stride = 3, 96
    frames = torch.zeros((n, stride), dtype=torch.uint8)
    used = torch.zeros(n, dtype=torch.int64)
    for i in range(n):
        f = i * world + rank                     # round-robin stream position
        ln = 9 + (f * 11) % 80
        frames[i, :ln] = (torch.arange(ln) * 3 + f).to(torch.uint8)
        used[i] = ln
    ring, lengths = sharding.gatherv_frames_to_root(frames, used, dst=0)
    ok = lengths.shape == (world, n)
    if rank == 0:
        for r in range(world):
            for i in range(n):
                f = i * world + r
                ln = 9 + (f * 11) % 80
                ok = ok and int(lengths[r, i]) == ln and torch.equal(ring[r, i, :ln], (torch.arange(ln) * 3 + f).to(torch.uint8))
    else:
        ok = ok and ring is None
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gatherv_to_root_world2_gloo():
    """The stream case of SURVEY.md 8e: exact-length frames delivered to rank 0 by grouped point-to-point transfers."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 500
    procs = [ctx.Process(target=_gatherv_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
