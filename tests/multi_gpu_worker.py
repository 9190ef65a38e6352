"""TEST INFRASTRUCTURE: body of tests/test_gpu_multi.py, one process per GPU under torchrun (NCCL).  Exercises, on real
GPUs and over NCCL, what tests/test_sharding_gloo.py exercises on CPU: the gatherv of encoded frames to rank 0, the padded
all-gather, and the banded single-frame encode; every result is checked by decoding."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import hap_b200
    import hap_b200.lib as L
    import oracles
    from hap_b200 import sharding, synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = hap_b200.load()
    codec, w, h, k = L.HapB200Codec_HapY, 1024, 512, 4
    n = 3                                                  # frames per rank
    imgs = torch.stack([synth.frame(w, h, i * world + rank, device=dev) for i in range(n)])
    cap = (lib.max_encoded_length_rgba(w, h, codec, k) + 255) // 256 * 256
    frames = torch.zeros((n, cap), dtype=torch.uint8, device=dev)
    used = torch.zeros(n, dtype=torch.int64, device=dev)
    assert lib.encode_rgba_batch(imgs.data_ptr(), n, w * h * 4, w, h, codec, 1, k, frames.data_ptr(), cap, used.data_ptr()) == 0
    tex_n = lib.texture_bytes(w, h, codec)

    # 1. gatherv to rank 0 over NCCL: every delivered frame decodes (reference build on the host) to the texture of the
    #    picture its stream position names
    ring, lengths = sharding.gatherv_frames_to_root(frames, used, 0)
    if rank == 0:
        ref = oracles.ref_abi() or oracles.oracle_abi()
        host = ring.cpu().numpy()
        for r in range(world):
            for i in range(n):
                fr = host[r, i, : int(lengths[r, i])].tobytes()
                img = synth.frame(w, h, i * world + r, device=dev)
                tex = torch.zeros(tex_n, dtype=torch.uint8, device=dev)
                assert lib.block_encode_batch(img.data_ptr(), 1, img.numel(), w, h, codec, tex.data_ptr(), tex_n) == 0
                rr, data, fmt, _ = ref.decode(fr, 0, tex_n)
                assert rr == 0 and data == tex.cpu().numpy().tobytes(), (r, i)
        assert int(lengths.sum() - lengths[0].sum()) > 0

    # 2. the padded all-gather (every rank gets the whole stream)
    got = sharding.gather_encoded_frames(frames, used, n * world)
    assert len(got) == n * world
    for f, g in enumerate(got):
        r, i = f % world, f // world
        assert g.numel() == int(lengths[r, i])
    mine = [got[i * world + rank] for i in range(n)]
    assert all(torch.equal(m, frames[i, : m.numel()]) for i, m in enumerate(mine))

    # 3. one picture cut into bands of whole chunks, one band per rank, band frames gathered and spliced on rank 0
    W2, H2, K2 = 1024, 128 * world, 2 * world
    pic = synth.frame(W2, H2, 77, device=dev)
    c0, c1 = sharding.chunk_band_for_rank(K2, world, rank)
    rows_per_chunk = H2 // K2
    band = pic[c0 * rows_per_chunk: c1 * rows_per_chunk].contiguous()
    bh = band.shape[0]
    bcap = (lib.max_encoded_length_rgba(W2, bh, codec, c1 - c0) + 255) // 256 * 256
    bf = torch.zeros(bcap, dtype=torch.uint8, device=dev)
    bu = torch.zeros(1, dtype=torch.int64, device=dev)
    assert lib.encode_rgba_batch(band.data_ptr(), 1, band.numel(), W2, bh, codec, 1, c1 - c0, bf.data_ptr(), bcap, bu.data_ptr()) == 0
    parts = sharding.gather_band_frames(bf, int(bu[0]), 0)
    if rank == 0:
        whole = sharding.assemble_banded_frame(parts)
        n2 = lib.texture_bytes(W2, H2, codec)
        tex = torch.zeros(n2, dtype=torch.uint8, device=dev)
        assert lib.block_encode_batch(pic.data_ptr(), 1, pic.numel(), W2, H2, codec, tex.data_ptr(), n2) == 0
        rr, data, fmt, calls = lib.decode(whole, 0, n2)
        assert rr == 0 and data == tex.cpu().numpy().tobytes() and calls == [K2]
        ro, od, _, _ = (oracles.ref_abi() or oracles.oracle_abi()).decode(whole, 0, n2)
        assert ro == 0 and od == data

    # 4. one process, several devices: a call whose buffers live on ANOTHER GPU runs there and leaves the caller's device alone
    if rank == 0 and torch.cuda.device_count() >= 2:
        other = torch.device("cuda", (local + 1) % torch.cuda.device_count())
        img_o = synth.frame(256, 128, 5, device=other)
        cap_o = (lib.max_encoded_length_rgba(256, 128, codec, 2) + 255) // 256 * 256
        out_o = torch.zeros(cap_o, dtype=torch.uint8, device=other)
        used_o = torch.zeros(1, dtype=torch.int64, device=other)
        assert lib.encode_rgba_batch(img_o.data_ptr(), 1, img_o.numel(), 256, 128, codec, 1, 2, out_o.data_ptr(), cap_o, used_o.data_ptr()) == 0
        assert torch.cuda.current_device() == local
        fr = out_o[: int(used_o[0])].cpu().numpy().tobytes()
        img_l = img_o.to(dev)
        out_l = torch.zeros(cap_o, dtype=torch.uint8, device=dev)
        used_l = torch.zeros(1, dtype=torch.int64, device=dev)
        assert lib.encode_rgba_batch(img_l.data_ptr(), 1, img_l.numel(), 256, 128, codec, 1, 2, out_l.data_ptr(), cap_o, used_l.data_ptr()) == 0
        assert fr == out_l[: int(used_l[0])].cpu().numpy().tobytes()
    # 5. delivery ring: every rank's encoder writes its frames STRAIGHT into rank 0's memory (CUDA IPC peer mapping), lengths
    #    into the slot header, a release store publishes the slot; rank 0 waits on the flags and decodes every frame (reference
    #    build on the host) to the texture of the picture its stream position names
    HEADER = 4096
    slot_bytes = HEADER + n * cap
    handle = torch.zeros(lib.RING_HANDLE_BYTES, dtype=torch.uint8, device=dev)
    ring_ptr = 0
    if rank == 0:
        rr, ring_ptr, hb = lib.ring_create(local, world * slot_bytes)
        assert rr == 0
        handle.copy_(torch.frombuffer(bytearray(hb), dtype=torch.uint8))
    dist.broadcast(handle, 0)
    if rank != 0:
        rr, ring_ptr = lib.ring_open(local, handle.cpu().numpy().tobytes())
        assert rr == 0 and ring_ptr != 0
    mine_slot = ring_ptr + rank * slot_bytes
    st = torch.cuda.Stream(device=dev)
    assert lib.encode_rgba_batch(imgs.data_ptr(), n, w * h * 4, w, h, codec, 1, k, mine_slot + HEADER, cap, mine_slot + 64, stream=st.cuda_stream) == 0
    assert lib.ring_publish(local, mine_slot, 1, stream=st.cuda_stream) == 0
    assert torch.cuda.current_device() == local
    if rank == 0:
        cs = torch.cuda.Stream(device=dev)
        for r in range(world):
            assert lib.ring_wait(local, ring_ptr + r * slot_bytes, 1, 30000 if r & 1 else 0, stream=cs.cuda_stream) == 0   # both kinds of wait
        cs.synchronize()

        class _Ext:
            def __init__(self, ptr, nb):
                self.__cuda_array_interface__ = {"shape": (nb,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        ring_t = torch.as_tensor(_Ext(ring_ptr, world * slot_bytes), device=dev)
        host = ring_t.cpu().numpy()
        ref = oracles.ref_abi() or oracles.oracle_abi()
        for r in range(world):
            o = r * slot_bytes
            assert int(np.frombuffer(host[o: o + 4].tobytes(), np.uint32)[0]) == 1
            lens = np.frombuffer(host[o + 64: o + 64 + 8 * n].tobytes(), np.uint64)
            assert [int(x) for x in lens] == [int(x) for x in lengths[r]], (r, lens, lengths[r])     # the same frames the gatherv moved
            for i in range(n):
                fr = host[o + HEADER + i * cap: o + HEADER + i * cap + int(lens[i])].tobytes()
                img = synth.frame(w, h, i * world + r, device=dev)
                tex = torch.zeros(tex_n, dtype=torch.uint8, device=dev)
                assert lib.block_encode_batch(img.data_ptr(), 1, img.numel(), w, h, codec, tex.data_ptr(), tex_n) == 0
                rr, data, fmt, _ = ref.decode(fr, 0, tex_n)
                assert rr == 0 and data == tex.cpu().numpy().tobytes(), ("ring", r, i)
        del ring_t
        # a flag nobody publishes: the wait gives up after its timeout instead of holding the GPU
        assert lib.ring_wait(local, ring_ptr + 8, 7, 50, stream=cs.cuda_stream) == 0
        cs.synchronize()
    st.synchronize()
    dist.barrier()
    if rank != 0:
        assert lib.ring_close(local, ring_ptr) == 0
    dist.barrier()
    if rank == 0:
        assert lib.ring_destroy(local, ring_ptr) == 0
        # same process, two devices: attach instead of open
        if torch.cuda.device_count() >= 2:
            od = (local + 1) % torch.cuda.device_count()
            other = torch.device("cuda", od)
            rr, rp, _ = lib.ring_create(local, slot_bytes)
            assert rr == 0 and lib.ring_attach(od, local) == 0
            img_o = imgs.to(other)
            so = torch.cuda.Stream(device=other)
            assert lib.encode_rgba_batch(img_o.data_ptr(), n, w * h * 4, w, h, codec, 1, k, rp + HEADER, cap, rp + 64, stream=so.cuda_stream) == 0
            assert lib.ring_publish(od, rp, 5, stream=so.cuda_stream) == 0
            assert lib.ring_wait(local, rp, 5, 30000) == 0          # legacy stream: returns when the flag is there
            torch.cuda.synchronize(dev)

            class _Ext2:
                def __init__(self, ptr, nb):
                    self.__cuda_array_interface__ = {"shape": (nb,), "typestr": "|u1", "data": (ptr, False), "version": 2}
            rt = torch.as_tensor(_Ext2(rp, slot_bytes), device=dev)
            lens = rt[64: 64 + 8 * n].view(torch.int64).cpu()
            assert [int(x) for x in lens] == [int(x) for x in lengths[0]]
            for i in range(n):
                assert torch.equal(rt[HEADER + i * cap: HEADER + i * cap + int(lens[i])], frames[i, : int(lens[i])])
            del rt
            so.synchronize()
            assert lib.ring_destroy(local, rp) == 0
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("multi-gpu worker ok: world", world)


if __name__ == "__main__":
    main()
