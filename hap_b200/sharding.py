"""Multi-GPU sharding of the Hap path: one process per GPU.

Frames of a stream are independent (HapVideoDRAFT.md:29-34), so frame f goes to rank f mod N and no
collective sits on the data path.  When a consumer needs the encoded stream in one place, the
variable-size frames are gathered with one all-gather of lengths plus one padded all-gather of bytes
(`gather_encoded_frames`); it works on NCCL (device tensors) and gloo (CPU tensors) alike.
A single very large frame can instead be split into bands of whole chunks (`chunk_band_for_rank`).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def frames_for_rank(frame_count: int, world: int, rank: int) -> List[int]:
    """Round-robin: the frame indices rank `rank` of `world` encodes/decodes."""
    return list(range(rank, frame_count, world))


def chunk_band_for_rank(chunks: int, world: int, rank: int) -> Tuple[int, int]:
    """[first, last) chunk indices of one frame owned by `rank` (contiguous, sizes differ by <= 1).
    Chunks are independent Snappy streams over contiguous block ranges (hap.c:433, :450), so a band of
    whole chunks can be produced without looking at its neighbours."""
    base, extra = divmod(chunks, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def gather_encoded_frames(local_frames: torch.Tensor, local_used: torch.Tensor, frame_count: int):
    """local_frames: (n_local, stride) uint8, local_used: (n_local,) int64 for this rank's frames
    (the ones `frames_for_rank` lists, in that order).  Returns on EVERY rank the list of `frame_count`
    byte tensors in stream order.  One all_gather of lengths, one all_gather of padded payloads."""
    world, rank = dist.get_world_size(), dist.get_rank()
    n_max = (frame_count + world - 1) // world
    stride = local_frames.shape[1] if local_frames.numel() else 0
    st = torch.tensor([stride], dtype=torch.int64, device=local_used.device)
    dist.all_reduce(st, op=dist.ReduceOp.MAX)
    stride = int(st.item())
    used = torch.zeros(n_max, dtype=torch.int64, device=local_used.device)
    used[: local_used.numel()] = local_used
    all_used = [torch.empty_like(used) for _ in range(world)]
    dist.all_gather(all_used, used)
    pad = torch.zeros((n_max, stride), dtype=torch.uint8, device=local_frames.device)
    if local_frames.numel():
        pad[: local_frames.shape[0], : local_frames.shape[1]] = local_frames
    all_frames = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(all_frames, pad)
    out = []
    for f in range(frame_count):
        r, i = f % world, f // world
        out.append(all_frames[r][i, : int(all_used[r][i])])
    return out
