"""Block-codec oracle: decoders checked against Pillow's independent bcn decoder; the cluster-fit
encoder (the quality bar) checked for sanity (decodable, PSNR floor, beats Pillow's encoder)."""
import io

import numpy as np
import pytest

import oracles
from hap_b200 import synth

PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402


def pil_decode(kind, blocks, w, h):
    if kind == "bc1":
        return np.asarray(Image.frombytes("RGBA", (w, h), blocks, "bcn", (1, "DXT1")))
    if kind == "bc3":
        return np.asarray(Image.frombytes("RGBA", (w, h), blocks, "bcn", (3, "DXT5")))
    return np.asarray(Image.frombytes("L", (w, h), blocks, "bcn", (4, "BC4U")))


@pytest.mark.parametrize("kind", ["bc1", "bc3", "bc4"])
def test_decoders_match_pillow_on_random_blocks(kind):
    rng = np.random.default_rng(3)
    w, h = 64, 32
    n = (w // 4) * (h // 4) * (8 if kind in ("bc1", "bc4") else 16)
    blocks = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    ours = oracles.bc_decode(kind, blocks, w, h)
    theirs = pil_decode(kind, blocks, w, h)
    if kind == "bc1":
        # Pillow returns transparent black for index 3 of 3-colour blocks, as we do
        assert np.array_equal(ours, theirs)
    else:
        assert np.array_equal(ours, theirs)


def test_clusterfit_quality_floor():
    img = synth.frame(128, 128, 0).numpy()
    for kind, floor in (("bc1", 30.0), ("bc3", 30.0), ("ycocg", 33.0)):
        blk = oracles.bc_encode_clusterfit(kind, img, 8)
        dec = oracles.bc_decode(kind, blk, 128, 128)
        assert oracles.psnr(img, dec) > floor, kind
        one = oracles.bc_encode_clusterfit(kind, img, 1)
        if kind != "ycocg":  # the YCoCg fit minimises error in (Co,Cg) space, not RGB
            assert oracles.psnr(img, oracles.bc_decode(kind, one, 128, 128)) <= oracles.psnr(img, dec) + 1e-9


def test_clusterfit_beats_pillow_encoder():
    img = synth.frame(128, 128, 0).numpy()
    buf = io.BytesIO()
    Image.fromarray(img, "RGBA").save(buf, "DDS", pixel_format="DXT1")
    pil_blocks = buf.getvalue()[128:]
    p_pil = oracles.psnr(img, oracles.bc_decode("bc1", pil_blocks, 128, 128))
    p_orc = oracles.psnr(img, oracles.bc_decode("bc1", oracles.bc_encode_clusterfit("bc1", img, 8), 128, 128))
    assert p_orc >= p_pil - 0.05


def test_bc4_squish_fit_is_exact_on_two_level_blocks():
    img = np.zeros((8, 8, 4), np.uint8)
    img[..., 3] = 40
    img[:, 4:, 3] = 200
    blk = oracles.bc_encode_clusterfit("bc4", img)
    assert np.array_equal(oracles.bc_decode("bc4", blk, 8, 8), img[..., 3])


def test_ycocg_roundtrip_of_gray_is_close():
    img = np.zeros((16, 16, 4), np.uint8)
    img[..., :3] = np.arange(16, dtype=np.uint8).reshape(1, 16, 1) * 16
    img[..., 3] = 255
    dec = oracles.bc_decode("ycocg", oracles.bc_encode_clusterfit("ycocg", img, 8), 16, 16)
    assert np.abs(dec[..., :3].astype(int) - img[..., :3].astype(int)).max() <= 4  # 128 is not on the 5:6:5 grid, so gray carries a small chroma bias


@pytest.mark.parametrize("kind", ["bc1", "bc3", "ycocg", "bc4"])
def test_kernel_block_decoder_source_is_bit_exact_against_oracle(kind):
    """The decoder math the CUDA kernel compiles (byte-permute palette look-ups), built for the host, on random
    blocks (every mode: 3- and 4-colour BC1, 6- and 8-value BC4, every YCoCg scale code) and on real textures."""
    import twin
    rng = np.random.default_rng(11)
    w, h = 256, 64
    per = 8 if kind in ("bc1", "bc4") else 16
    blocks = rng.integers(0, 256, (w // 4) * (h // 4) * per, dtype=np.uint8).tobytes()
    got = twin.decode(kind, blocks, w, h)
    want = oracles.bc_decode(kind, blocks, w, h)
    if kind == "bc4":
        assert np.array_equal(got[..., 0], want) and np.array_equal(got[..., 2], want) and (got[..., 3] == 255).all()
    else:
        assert np.array_equal(got, want)
