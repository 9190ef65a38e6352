"""Loads tests/golden/hap_golden.json (outputs of the unmodified reference, see make_golden.py)."""
import base64
import functools
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


@functools.lru_cache(None)
def golden():
    with open(os.path.join(HERE, "golden", "hap_golden.json")) as fh:
        return json.load(fh)


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def b64(s):
    return base64.b64decode(s)


def noise_bytes(n, seed=1):
    import torch
    from hap_b200 import synth
    v = synth.pcg_hash(torch.arange(n, dtype=torch.int64) ^ seed)
    return (v >> 24).to(torch.uint8).numpy().tobytes()


def kat_c_bytes(n=33177600):
    i = np.arange(n, dtype=np.int64)
    return (((i // 4096) * 3) & 0xFF).astype(np.uint8).tobytes()


def container_checks(lib, decode_kw=None):
    """Runs every golden container/Snappy known-answer case against `lib` (a HapABI).  Shared by the
    oracle tests (CPU) and the CUDA library tests (-m gpu).  Returns a list of mismatch strings."""
    from hap_b200.abi import HapCompressorNone, HapCompressorSnappy
    G = golden()
    bad = []
    kw = decode_kw or {}

    def chk(name, got, want):
        if got != want:
            bad.append(f"{name}: got {got!r} want {want!r}")

    # decode of reference frames
    x = bytes([0x55]) * 64
    for key in ("kat_a", "kat_a_none"):
        r, data, fmt, calls = lib.decode(bytes.fromhex(G[key]["frame"]), 0, 64, **kw)
        chk(key + ".decode", (r, data, fmt, calls), (0, x, 0x83F0, []))
    f = bytes.fromhex(G["kat_d"]["frame"])
    chk("kat_d.count", lib.texture_count(f), (0, 2))
    chk("kat_d.formats", [lib.texture_format(f, i)[1] for i in (0, 1)], G["kat_d"]["formats"])
    chk("kat_d.chunks", [lib.chunk_count(f, i)[1] for i in (0, 1)], G["kat_d"]["chunk_counts"])
    r, data, fmt, calls = lib.decode(f, 0, 4096, **kw)
    chk("kat_d.tex0", (r, data, fmt, calls), (0, kat_c_bytes(4096), 0x01, [2]))
    r, data, fmt, calls = lib.decode(f, 1, 2048, **kw)
    chk("kat_d.tex1", (r, data, fmt, calls), (0, bytes([0x55]) * 2048, 0x8DBB, [2]))
    for name, e in G["ref_frames"].items():
        fr = b64(e["frame_b64"])
        if name == "hapm":
            for i in (0, 1):
                r, data, fmt, calls = lib.decode(fr, i, e["payload_len"][i], **kw)
                chk(f"ref_frames.hapm[{i}]", (r, sha(data or b"")), (0, e["payload_sha256"][i]))
        else:
            r, data, fmt, calls = lib.decode(fr, 0, e["payload_len"], **kw)
            chk(f"ref_frames.{name}", (r, sha(data or b""), fmt), (0, e["payload_sha256"], e["format"]))
            chk(f"ref_frames.{name}.chunks", lib.chunk_count(fr, 0), (0, e["chunks"]))
    # raw snappy streams
    for name, e in G["snappy_streams"].items():
        s = bytes.fromhex(e["stream"])
        n = len(s)
        whole = (n.to_bytes(3, "little") + b"\xbe" + s) if n else (b"\0\0\0\xbe" + (0).to_bytes(4, "little"))
        r, data, fmt, calls = lib.decode(whole, 0, 4096, **kw)
        w = e["whole_section"]
        chk(f"snappy.{name}.whole", (r, sha(data) if data is not None else None), (w["result"], w["out_sha256"]))
        di = b"\x01\0\0\x02\x0b" + b"\x04\0\0\x03" + n.to_bytes(4, "little")
        body = len(di).to_bytes(3, "little") + b"\x01" + di + s
        cx = len(body).to_bytes(3, "little") + b"\xce" + body
        r, data, fmt, calls = lib.decode(cx, 0, 4096, **kw)
        w = e["complex_1chunk"]
        chk(f"snappy.{name}.complex", (r, sha(data) if data is not None else None), (w["result"], w["out_sha256"]))
    # container cases
    for name, e in G["container_cases"].items():
        fr = bytes.fromhex(e["frame"])
        r, data, fmt, calls = lib.decode(fr, e["index"], e["cap"], **kw)
        chk(f"case.{name}", (r, data.hex() if data is not None else None, calls), (e["result"], e["out"], e["callback_counts"]))
        if r == 0:
            chk(f"case.{name}.fmt", fmt, e["format"])
        if e["texture_count"] is not None:
            got = lib.texture_count(fr)
            if e["texture_count"][0] == 0:
                chk(f"case.{name}.count", list(got), e["texture_count"])
            else:
                chk(f"case.{name}.count.result", got[0], e["texture_count"][0])
        if e["texture_format"] is not None:
            got = lib.texture_format(fr, e["index"])
            chk(f"case.{name}.format", got[0] if e["texture_format"][0] else list(got), e["texture_format"][0] if e["texture_format"][0] else e["texture_format"])
        if e["chunk_count"] is not None:
            chk(f"case.{name}.chunkcount", list(lib.chunk_count(fr, e["index"])), e["chunk_count"])
    return bad
