"""CPU-side checks of libhap_b200.so: it loads, exports every symbol the headers declare, and its
host logic (size bounds, header walks, verbatim frames, argument validation) matches the reference's
golden results.  No compute entry point is exercised successfully here: without a GPU they must fail."""
import os
import re

import pytest

import hap_b200
from golden_util import golden, kat_c_bytes, sha
from hap_b200.abi import (HapCompressorNone, HapCompressorSnappy, HapTextureFormat_A_RGTC1,
                          HapTextureFormat_RGB_DXT1, HapTextureFormat_YCoCg_DXT5)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DXT1, YCOCG, RGTC1 = HapTextureFormat_RGB_DXT1, HapTextureFormat_YCoCg_DXT5, HapTextureFormat_A_RGTC1


@pytest.fixture(scope="module")
def lib():
    from hap_b200 import build
    build.build()
    return hap_b200.load()


def have_gpu():
    import torch
    return torch.cuda.is_available()


def test_exports_every_declared_symbol(lib):
    names = set()
    for hdr in ("hap.h", "hap_b200.h", "hap_mov.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        names |= set(re.findall(r"\b(Hap(?:B200)?[A-Z][A-Za-z0-9]*)\s*\(", text))
    names -= {"HapDecodeWorkFunction", "HapDecodeCallback", "HapMaxEncodedLength()"}
    assert {"HapEncode", "HapDecode", "HapMaxEncodedLength", "HapGetFrameTextureCount", "HapGetFrameTextureFormat",
            "HapGetFrameTextureChunkCount", "HapB200EncodeRGBABatch", "HapB200DecodeBatch", "HapB200MovOpen",
            "HapB200MovWriteFrame"} <= names
    for n in sorted(names):
        assert hasattr(lib.lib, n), n


def test_max_encoded_length_matches_reference(lib):
    G = golden()
    assert lib.max_encoded_length([1036800], [DXT1], [1]) == G["max_len_1080p_dxt1"]
    assert lib.max_encoded_length([64], [DXT1], [1]) == G["kat_a"]["max_len"]
    assert lib.max_encoded_length([4096, 2048], [YCOCG, RGTC1], [2, 2]) == G["kat_d"]["max_len"]
    assert lib.max_encoded_length([33177600], [YCOCG], [8]) == G["kat_c"]["max_len"]
    assert lib.max_encoded_length([64], [DXT1], [0]) == 0
    assert lib.max_encoded_length([64, 64, 64], [DXT1] * 3, [1] * 3) == 0
    # SURVEY.md section 8 table
    assert lib.max_encoded_length([4147200], [DXT1], [1]) == 4838465
    assert lib.max_encoded_length([8294400], [YCOCG], [8]) == 9677124


def test_header_walks_match_reference(lib):
    G = golden()
    f = bytes.fromhex(G["kat_d"]["frame"])
    assert lib.texture_count(f) == (0, 2)
    assert [lib.texture_format(f, i)[1] for i in (0, 1)] == G["kat_d"]["formats"]
    assert [lib.chunk_count(f, i)[1] for i in (0, 1)] == G["kat_d"]["chunk_counts"]
    for name, e in G["container_cases"].items():
        fr = bytes.fromhex(e["frame"])
        if e["texture_count"] is not None:
            got = lib.texture_count(fr)
            assert (list(got) == e["texture_count"]) if e["texture_count"][0] == 0 else got[0] == e["texture_count"][0], name
        if e["texture_format"] is not None:
            got = lib.texture_format(fr, e["index"])
            assert (list(got) == e["texture_format"]) if e["texture_format"][0] == 0 else got[0] == e["texture_format"][0], name
        if e["chunk_count"] is not None:
            assert list(lib.chunk_count(fr, e["index"])) == e["chunk_count"], name


def test_verbatim_frames_are_byte_identical_to_reference(lib):
    G = golden()
    x = bytes([0x55]) * 64
    assert lib.encode([x], [DXT1], [HapCompressorNone], [1])[1].hex() == G["kat_a_none"]["frame"]
    r, f = lib.encode([kat_c_bytes(4096), bytes([0x55]) * 2048], [YCOCG, RGTC1], [0, 0], [2, 2])
    assert r == 0 and (sha(f), len(f), f[:12].hex()) == (G["kat_d_none"]["frame_sha256"], G["kat_d_none"]["frame_len"], G["kat_d_none"]["header"])
    # verbatim textures decode without any compute as well
    assert lib.decode(f, 1, 2048)[:3] == (0, bytes([0x55]) * 2048, RGTC1)


def test_argument_validation_matches_reference(lib):
    G = golden()["encode_results"]
    p16 = bytes(16)
    assert lib.encode([p16], [DXT1], [1], [0])[0] == G["chunk_0"]
    assert lib.encode([p16], [0x1234], [1], [1])[0] == G["bad_format"]
    assert lib.encode([p16], [DXT1], [7], [1])[0] == G["bad_compressor"]
    assert lib.encode([p16], [DXT1], [1], [1], out_capacity=20)[0] == G["small_buffer"]
    assert lib.encode([p16, p16], [DXT1, DXT1], [1, 1], [1, 1])[0] == G["two_dxt1"]
    f = bytes.fromhex(golden()["kat_a"]["frame"])
    assert lib.decode(f, 0, 64, callback=None)[0] == 1
    assert lib.decode(f, 2, 64)[0] == 1
    assert lib.decode(f, 1, 64)[0] == 1


def test_rgba_geometry_helpers(lib):
    from hap_b200.lib import HapB200Codec_Hap1, HapB200Codec_HapM, HapB200Codec_HapY
    assert lib.texture_bytes(3840, 2160, HapB200Codec_HapY) == 8294400
    assert lib.texture_bytes(3840, 2160, HapB200Codec_Hap1) == 4147200
    assert lib.texture_bytes(7680, 4320, HapB200Codec_HapM, 1) == 16588800
    assert lib.max_encoded_length_rgba(3840, 2160, HapB200Codec_HapY, 8) == 9677124
    assert lib.texture_bytes(3841, 2160, HapB200Codec_HapY) == 0


@pytest.mark.skipif(have_gpu(), reason="only meaningful where there is no CUDA device")
def test_compute_paths_fail_loudly_without_a_gpu(lib):
    x = bytes([0x55]) * 64
    assert lib.encode([x], [DXT1], [HapCompressorSnappy], [1])[0] == 4  # HapResult_Internal_Error, no CPU fallback
    assert lib.decode(bytes.fromhex(golden()["kat_a"]["frame"]), 0, 64)[0] == 4


def _sec(typ, body):
    assert len(body) < (1 << 24)
    return len(body).to_bytes(3, "little") + bytes([typ]) + body


def test_decode_instruction_tables_must_all_describe_the_same_chunks(lib):
    """a compressor table of length 0 (or a size table shorter than 4 bytes) leaves the chunk count
    to the other table in hap.c:709-716; the per-chunk reads then run past the short table.  Here: Bad_Frame, on the
    header walk and on HapDecode alike, before anything is launched."""
    data = bytes(64)
    sizes3 = b"".join((16).to_bytes(4, "little") for _ in range(3))
    cases = {
        "empty_compressor_table": _sec(0x02, b"") + _sec(0x03, sizes3),
        "short_size_table": _sec(0x02, bytes([0x0A] * 3)) + _sec(0x03, b"\x10\x00"),
        "short_offset_table": _sec(0x02, bytes([0x0A] * 3)) + _sec(0x03, sizes3) + _sec(0x04, b"\x00\x00"),
        "offset_table_of_two": _sec(0x02, bytes([0x0A] * 3)) + _sec(0x03, sizes3) + _sec(0x04, bytes(8)),
    }
    for name, tables in cases.items():
        frame = _sec(0xCB, _sec(0x01, tables) + data)
        assert lib.chunk_count(frame, 0)[0] == 3, name
        assert lib.decode(frame, 0, 64)[0] == 3, name
    # the well-formed sibling (three raw chunks) is accepted by the header walk; an unknown section is skipped
    good = _sec(0xCB, _sec(0x01, _sec(0x02, bytes([0x0A] * 3)) + _sec(0x7B, b"private") + _sec(0x03, sizes3)) + bytes(48))
    assert lib.chunk_count(good, 0) == (0, 3)
