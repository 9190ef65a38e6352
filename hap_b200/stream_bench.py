"""bench.py --workload 16k_stream: BASELINE.json configs[4] -- a 16384 x 16384 RGBA stream, Hap Q (scaled YCoCg DXT5,
64 chunks, Snappy), frames sharded round-robin over the GPUs of one box, the ENCODED frames delivered to rank 0 (where a
muxer would sit) inside the timed region.

One process per GPU (torchrun), NCCL over NVLink / NVSwitch.  One STEP on every rank:
    encode B device-resident frames (HapB200EncodeRGBABatch)             -- the hot path, no communication
    all-gather of the B encoded lengths                                  -- 8 bytes per frame
    grouped ncclSend/ncclRecv of exactly the encoded bytes to rank 0      -- sharding.gatherv_frames_to_root
Timing: CUDA events on the stream around K steps, barrier + synchronize on both sides, max over ranks.  `value` = RGBA bytes
of all frames encoded and delivered per second; fps next to it.  The only host synchronisation inside a step is the read-back
of the lengths (message sizes must be known on the host).  Outside the timed region rank 0 decodes one frame that came from
another rank and compares it with a checksum of that rank's own texture.
"""
from __future__ import annotations

import os
import time


def run(args, rank, local_rank, world, emit, ClockSampler, measured_peak_hbm):
    import torch
    import torch.distributed as dist

    import hap_b200
    from hap_b200 import sharding, synth
    from hap_b200.lib import HapB200Codec_HapY

    W = H = int(os.environ.get("HAPB200_STREAM_SIZE", "16384"))
    CH = 64
    B = int(os.environ.get("HAPB200_STREAM_BATCH", "2"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = hap_b200.load()
    lib.set_option(lib.OPTION_WRITE_INDEX, 0 if args.no_index else 1)
    codec = HapB200Codec_HapY
    rgba_bytes = 4 * W * H
    tex_bytes = lib.texture_bytes(W, H, codec)
    cap = (lib.max_encoded_length_rgba(W, H, codec, CH) + 255) // 256 * 256
    peak, peak_src = measured_peak_hbm()

    # B distinct frames per rank, built tile by tile (a 16K frame of int64 intermediates would not fit otherwise)
    rgba = torch.empty((B, H, W, 4), dtype=torch.uint8, device=dev)
    th = min(2048, H)
    for b in range(B):
        for ty in range(0, H, th):
            for tx in range(0, W, 2 * th):
                tw = min(2 * th, W - tx)
                rgba[b, ty:ty + th, tx:tx + tw] = synth.frame(tw, th, (rank * B + b) * 64 + (ty // th) * 8 + tx // (2 * th), device=dev)
    # two sets of buffers: the delivery of batch i (NCCL's own stream) overlaps the encode of batch i+1
    frames2 = [torch.empty((B, cap), dtype=torch.uint8, device=dev) for _ in range(2)]
    used2 = [torch.zeros(B, dtype=torch.int64, device=dev) for _ in range(2)]
    ring2 = [torch.empty((world, B, cap), dtype=torch.uint8, device=dev) for _ in range(2)] if rank == 0 else [None, None]
    frames, used, ring = frames2[0], used2[0], ring2[0]
    stream = torch.cuda.Stream(device=dev)
    sp = stream.cuda_stream
    moved = {"bytes": 0}
    state = {"n": 0, "works": [[], []]}

    def finish(k):
        for w in state["works"][k]:
            w.wait()
        state["works"][k] = []

    def step():
        k = state["n"] & 1
        state["n"] += 1
        finish(k)                       # the transfers that read / wrote buffer set k two batches ago
        r = lib.encode_rgba_batch(rgba.data_ptr(), B, rgba_bytes, W, H, codec, 1, CH, frames2[k].data_ptr(), cap, used2[k].data_ptr(), stream=sp)
        assert r == 0, r
        if world > 1:
            _, lengths, works = sharding.gatherv_frames_to_root(frames2[k], used2[k], 0, ring2[k], wait=False)
            state["works"][k] = works
            moved["bytes"] = int(lengths.sum() - lengths[0].sum())
        else:
            lengths = used2[k].cpu().view(1, B)
            for i in range(B):
                ring2[k][0, i, : int(lengths[0, i])].copy_(frames2[k][i, : int(lengths[0, i])], non_blocking=True)
        state["last"] = k
        return lengths

    def drain():
        finish(0)
        finish(1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            lengths = step()
        drain()
        barrier()
        launches0 = lib.launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wall0 = time.time()
        e0.record(stream)
        for _ in range(args.steps):
            lengths = step()
        drain()                         # every frame of the K batches has arrived on rank 0 inside the timed region
        e1.record(stream)
        barrier()
        wall1 = time.time()
        launches = lib.launches() - launches0
        clocks = sampler.stop(wall0, wall1) if rank == 0 else None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # encode alone (no delivery), for the share of the step the exchange costs
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for _ in range(args.steps):
            r = lib.encode_rgba_batch(rgba.data_ptr(), B, rgba_bytes, W, H, codec, 1, CH, frames2[0].data_ptr(), cap, used2[0].data_ptr(), stream=sp)
            assert r == 0
        e3.record(stream)
        e3.synchronize()
        ms_enc = torch.tensor([e2.elapsed_time(e3)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms_enc, op=dist.ReduceOp.MAX)

    # ---- verification (outside the timed region): a frame that came from the LAST rank decodes on rank 0 to that rank's texture
    tex = torch.empty(tex_bytes, dtype=torch.uint8, device=dev)
    assert lib.block_encode_batch(rgba.data_ptr(), 1, rgba_bytes, W, H, codec, tex.data_ptr(), tex_bytes) == 0
    mine = torch.stack([tex.view(torch.int32).sum(dtype=torch.int64), tex[::4099].to(torch.int64).sum()])
    sums = torch.zeros((world, 2), dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_gather_into_tensor(sums.view(-1), mine)
    else:
        sums[0] = mine
    verified = None
    if rank == 0:
        src_rank = world - 1
        n = int(lengths[src_rank, 0])
        back = torch.empty(tex_bytes, dtype=torch.uint8, device=dev)
        bu, bf, br = torch.zeros(1, dtype=torch.int64, device=dev), torch.zeros(1, dtype=torch.int32, device=dev), torch.full((1,), 9, dtype=torch.int32, device=dev)
        ln = torch.tensor([n], dtype=torch.int64, device=dev)
        ring = ring2[state["last"]]
        assert lib.decode_batch(ring[src_rank, 0].data_ptr(), 1, cap, ln.data_ptr(), 0, CH, back.data_ptr(), tex_bytes, bu.data_ptr(), bf.data_ptr(), br.data_ptr()) == 0
        got = torch.stack([back.view(torch.int32).sum(dtype=torch.int64), back[::4099].to(torch.int64).sum()])
        verified = bool(br.tolist() == [0] and bu.tolist() == [tex_bytes] and torch.equal(got, sums[src_rank]))
        assert verified, "a frame delivered to rank 0 does not decode to its sender's texture"

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    ms_step = float(ms.item()) / args.steps
    ms_enc_step = float(ms_enc.item()) / args.steps
    frames_per_step = world * B
    fps = frames_per_step / (ms_step * 1e-3)
    mean_frame = float(lengths.double().mean())
    line = {
        "metric": "hapq_16k_stream_encode_deliver_rgba_GBps", "value": frames_per_step * rgba_bytes / (ms_step * 1e-3) / 1e9, "unit": "GB/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"hap_q_16k_stream({W}x{H},YCoCg-DXT5,snappy,{CH}chunks), frames round-robin over {world} gpu(s), encoded frames "
                               "delivered to rank 0 inside the timed region", "frames_per_gpu_per_step": B,
                   "l2": f"inputs larger than L2 ({B * rgba_bytes / 1e9:.2f} GB RGBA per step per GPU)",
                   "parallelism": f"dp{world}: all-gather of lengths + grouped ncclSend/ncclRecv gatherv to rank 0, "
                                  "the delivery of batch i overlapping the encode of batch i+1"},
        "fps": fps, "fps_target_of_config": 60,
        "encode_only_ms_per_step": ms_enc_step, "delivery_share_of_step": max(0.0, 1.0 - ms_enc_step / ms_step),
        "nvlink_bytes_per_step": moved["bytes"], "nvlink_GBps_into_rank0": moved["bytes"] / (ms_step * 1e-3) / 1e9,
        "compression_ratio": mean_frame / tex_bytes, "delivered_frame_verified": verified,
        "roofline": {"bound": "hbm", "kernel": "encode pipeline (RGBA read)", "achieved": B * rgba_bytes / (ms_enc_step * 1e-3) / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": B * rgba_bytes / (ms_enc_step * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src},
        "clocks": clocks, "gpu_launches": launches,
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
