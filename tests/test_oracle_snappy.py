"""Pins oracle/snappy_oracle.c against genuine Google Snappy (pyarrow-bundled) in both directions."""
import numpy as np
import pytest

import oracles
from golden_util import golden, sha

pa = pytest.importorskip("pyarrow")


def payloads():
    rng = np.random.default_rng(7)
    yield b""
    yield b"a"
    yield b"hello " * 5 + b"hello"
    yield bytes(100000)
    yield rng.integers(0, 256, 70000, dtype=np.uint8).tobytes()
    blk = rng.integers(0, 256, 16, dtype=np.uint8)
    a = np.tile(blk, 9000)
    a[::37] ^= 0x5A
    yield a.tobytes()
    yield (np.arange(200000) // 3 % 251).astype(np.uint8).tobytes()


@pytest.mark.parametrize("i", range(7))
def test_cross_decode(i):
    p = list(payloads())[i]
    ours = oracles.snappy_compress(p)
    assert len(ours) <= 32 + len(p) + len(p) // 6
    if p:
        assert pa.decompress(ours, decompressed_size=len(p), codec="snappy", asbytes=True) == p
        theirs = pa.compress(p, codec="snappy", asbytes=True)
        assert oracles.snappy_uncompress(theirs) == (0, p)
    assert oracles.snappy_uncompress(ours) == (0, p)


def test_known_bytes():
    # SURVEY.md Appendix B example, produced by Google Snappy
    assert oracles.snappy_uncompress(bytes.fromhex("2314" + b"hello ".hex() + "720600")) == (0, b"hello " * 5 + b"hello")


def test_handmade_streams_match_reference_results():
    for name, e in golden()["snappy_streams"].items():
        st, out = oracles.snappy_uncompress(bytes.fromhex(e["stream"]), cap=4096)
        want = e["complex_1chunk"]
        if want["result"] == 0:
            assert st == 0 and sha(out) == want["out_sha256"], name
        else:
            assert st == 1, name  # SNAPPY_INVALID_INPUT -> Bad_Frame (hap.c:617-628)


def test_buffer_too_small():
    c = oracles.snappy_compress(bytes(1000))
    assert oracles.snappy_uncompress(c, cap=999)[0] == 2
