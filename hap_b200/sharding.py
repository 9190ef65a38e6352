"""Multi-GPU sharding of the Hap path: one process per GPU.

Frames of a stream are independent (HapVideoDRAFT.md:29-34), so frame f goes to rank f mod N and no
collective sits on the data path.  When a consumer needs the encoded stream in one place, the
variable-size frames are gathered with one all-gather of lengths plus one padded all-gather of bytes
(`gather_encoded_frames`); it works on NCCL (device tensors) and gloo (CPU tensors) alike.
A single very large frame can instead be split into bands of whole chunks (`chunk_band_for_rank`): every rank
encodes its band with the ordinary single-frame call, one gather brings the band frames together
(`gather_band_frames`) and `assemble_banded_frame` splices them into the frame of the whole picture.
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


# ---- one huge frame cut into bands of whole chunks (second case) ------------------------------------
# Every rank runs the ordinary single-texture encode (HapEncode / HapB200EncodeRGBA) on ITS band of block rows with
# ITS share of the chunk count; that yields a complete little Hap frame per band.  Chunks are independent Snappy
# streams laid out back to back behind two tables (hap.c:430-476), so the frame of the whole picture is: one section
# header, one Decode Instructions container whose tables are the bands' tables one after the other, the bands' chunk
# payloads one after the other.  `assemble_banded_frame` does exactly that on the gathered band frames (host bytes);
# `gather_band_frames` is the one exchange step (variable-size gather, NCCL or gloo).

_SEC_DI, _SEC_COMPRESSORS, _SEC_SIZES, _SEC_OFFSETS = 0x01, 0x02, 0x03, 0x04
_CHUNK_RAW, _CHUNK_SNAPPY, _COMPLEX = 0x0A, 0x0B, 0x0C


def _section(buf: bytes, off: int, end: int):
    """(header bytes, body length, type) of the section at off (hap.c:137-187); raises ValueError when it does not fit."""
    if end - off < 4:
        raise ValueError("truncated section header")
    length = buf[off] | (buf[off + 1] << 8) | (buf[off + 2] << 16)
    hdr = 4
    if length == 0:
        if end - off < 8:
            raise ValueError("truncated 8-byte section header")
        length = int.from_bytes(buf[off + 4: off + 8], "little")
        hdr = 8
    if length > end - off - hdr:
        raise ValueError("section longer than its buffer")
    return hdr, length, buf[off + 3]


def _put_section_header(length: int, typ: int, force8: bool = False) -> bytes:
    if length < (1 << 24) and not force8 and length != 0:
        return bytes([length & 255, (length >> 8) & 255, (length >> 16) & 255, typ])
    return bytes([0, 0, 0, typ]) + length.to_bytes(4, "little")


def split_single_texture_frame(frame: bytes):
    """One single-texture Hap frame -> (format nibble, [(compressor 0x0A|0x0B, chunk payload bytes), ...]).
    A verbatim section (0xA?) comes back as one raw chunk; an 0xB? section as one Snappy chunk."""
    hdr, length, typ = _section(frame, 0, len(frame))
    comp, fmt = typ >> 4, typ & 0xF
    body = frame[hdr: hdr + length]
    if comp == _CHUNK_RAW or comp == _CHUNK_SNAPPY:
        return fmt, [(comp, bytes(body))]
    if comp != _COMPLEX:
        raise ValueError("not a single-texture Hap frame")
    dh, dlen, dtyp = _section(body, 0, len(body))
    if dtyp != _SEC_DI:
        raise ValueError("complex section without Decode Instructions")
    compressors = sizes = offsets = None
    off, end = dh, dh + dlen
    while off < end:
        sh, slen, styp = _section(body, off, end)
        data = body[off + sh: off + sh + slen]
        if styp == _SEC_COMPRESSORS:
            compressors = list(data)
        elif styp == _SEC_SIZES:
            sizes = [int.from_bytes(data[i: i + 4], "little") for i in range(0, len(data) - 3, 4)]
        elif styp == _SEC_OFFSETS:
            offsets = [int.from_bytes(data[i: i + 4], "little") for i in range(0, len(data) - 3, 4)]
        off += sh + slen
    if compressors is None or sizes is None or len(compressors) != len(sizes) or (offsets is not None and len(offsets) != len(sizes)):
        raise ValueError("inconsistent Decode Instructions tables")
    payload = body[end:]
    chunks, at = [], 0
    for i, (c, n) in enumerate(zip(compressors, sizes)):
        if offsets is not None:
            at = offsets[i]
        if c not in (_CHUNK_RAW, _CHUNK_SNAPPY) or at + n > len(payload):
            raise ValueError("bad chunk table entry")
        chunks.append((c, bytes(payload[at: at + n])))
        at += n
    return fmt, chunks


def assemble_banded_frame(band_frames: List[bytes]) -> bytes:
    """Single-texture Hap frames of consecutive bands of one picture -> the Hap frame of the whole picture
    (a Complex section: hap.c:430-476; chunk i of the result is chunk i of the concatenated bands)."""
    fmt, chunks = None, []
    for bf in band_frames:
        f, ch = split_single_texture_frame(bytes(bf))
        if fmt is not None and f != fmt:
            raise ValueError("bands of different texture formats")
        fmt = f
        chunks.extend(ch)
    if fmt is None or not chunks:
        raise ValueError("no bands")
    k = len(chunks)
    comp_table = _put_section_header(k, _SEC_COMPRESSORS) + bytes(c for c, _ in chunks)
    size_table = _put_section_header(4 * k, _SEC_SIZES) + b"".join(len(p).to_bytes(4, "little") for _, p in chunks)
    di = _put_section_header(len(comp_table) + len(size_table), _SEC_DI) + comp_table + size_table
    body = di + b"".join(p for _, p in chunks)
    return _put_section_header(len(body), (_COMPLEX << 4) | fmt) + body


def gather_band_frames(band_frame: torch.Tensor, used: int, dst: int = 0):
    """band_frame: uint8 tensor holding this rank's band frame in its first `used` bytes (device tensor with NCCL,
    CPU tensor with gloo).  Returns on rank `dst` the list of all ranks' band frames as bytes, in rank order (= band
    order); None elsewhere.  One all_gather of lengths, one all_gather of padded payloads."""
    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([used], dtype=torch.int64, device=band_frame.device)
    lens = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(lens, n)
    cap = max(int(x.item()) for x in lens)
    pad = torch.zeros(cap, dtype=torch.uint8, device=band_frame.device)
    pad[:used] = band_frame[:used]
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    if rank != dst:
        return None
    return [bytes(parts[r][: int(lens[r].item())].cpu().numpy().tobytes()) for r in range(world)]


# ---- a stream of frames delivered to one rank (first case; BASELINE.json configs[4]) ------------------------------
# Every rank encodes its share of the frames; the consumer (a muxer, a file writer) sits on rank `dst`.  Frames have
# different lengths, so the exchange is a GATHERV: one all-gather of the lengths (8 bytes per frame), then one group of
# point-to-point transfers that move exactly the encoded bytes -- ncclSend/ncclRecv under NCCL, which on an NVSwitch box go
# GPU to GPU over NVLink.  No padding travels (the padded all-gather of gather_encoded_frames moves world x worst-case
# bytes to EVERY rank; this moves the sum of the real lengths to one).

def gatherv_frames_to_root(local_frames: torch.Tensor, local_used: torch.Tensor, dst: int = 0, ring: torch.Tensor = None,
                           wait: bool = True):
    """local_frames: (n, stride) uint8 -- this rank's n encoded frames, frame i in its first local_used[i] bytes (every
    rank passes the same n and stride).  On rank `dst`: returns (ring, lengths) where ring is a (world, n, stride) uint8
    tensor holding rank r's frame i in ring[r, i, :lengths[r, i]] (pass `ring` to reuse a buffer) -- stream order for
    round-robin sharding is (i, r).  Elsewhere: (None, lengths).  lengths is a (world, n) int64 CPU tensor on every
    rank; total bytes moved over the interconnect = lengths.sum() - lengths[dst].sum().
    wait=False returns (ring, lengths, works) without waiting for the transfers: the caller overlaps them with the next
    batch's encode and calls w.wait() on every work before it reuses `local_frames` / reads `ring`."""
    world, rank = dist.get_world_size(), dist.get_rank()
    n, stride = int(local_frames.shape[0]), int(local_frames.shape[1])
    lens_dev = torch.empty((world, n), dtype=torch.int64, device=local_used.device)
    dist.all_gather_into_tensor(lens_dev.view(-1), local_used.contiguous().view(-1))
    lengths = lens_dev.cpu()     # the message sizes must be known on the host: the one synchronisation of the step
    ops = []
    if rank == dst:
        if ring is None:
            ring = torch.empty((world, n, stride), dtype=torch.uint8, device=local_frames.device)
        for i in range(n):
            ring[dst, i, : int(lengths[dst, i])].copy_(local_frames[i, : int(lengths[dst, i])], non_blocking=True)
        for r in range(world):
            if r == dst:
                continue
            for i in range(n):
                ops.append(dist.P2POp(dist.irecv, ring[r, i, : int(lengths[r, i])], r))
    else:
        for i in range(n):
            ops.append(dist.P2POp(dist.isend, local_frames[i, : int(lengths[rank, i])], dst))
    works = dist.batch_isend_irecv(ops) if ops else []
    if not wait:
        return (ring if rank == dst else None), lengths, works
    for w in works:
        w.wait()
    return (ring if rank == dst else None), lengths
