"""Block codecs on the GPU (K1-K4, K8): device result bit-identical to the host build of the same
source, decodable by the oracle decoders (themselves checked against Pillow), PSNR at the oracle's bar."""
import numpy as np
import pytest
import torch

import hap_b200
import hap_b200.lib as L
import oracles
import twin
from hap_b200 import synth

pytestmark = pytest.mark.gpu
KINDS = [("bc1", L.HapB200Codec_Hap1), ("bc3", L.HapB200Codec_Hap5), ("ycocg", L.HapB200Codec_HapY), ("bc4", L.HapB200Codec_HapA)]


@pytest.fixture(scope="module")
def lib():
    return hap_b200.load()


def gpu_blocks(lib, img, codec):
    h, w = img.shape[:2]
    d = torch.from_numpy(np.ascontiguousarray(img)).cuda()
    n = lib.texture_bytes(w, h, codec, 0) + lib.texture_bytes(w, h, codec, 1)
    out = torch.zeros((n + 15) // 16 * 16, dtype=torch.uint8, device="cuda")
    assert lib.block_encode_batch(d.data_ptr(), 1, d.numel(), w, h, codec, out.data_ptr(), out.numel()) == 0
    return out[:n].cpu().numpy().tobytes()


@pytest.mark.parametrize("name,codec", KINDS)
@pytest.mark.parametrize("kind", ["video", "noise", "flat"])
def test_device_blocks_equal_host_build_of_same_source(lib, name, codec, kind):
    img = synth.frame(256, 192, 1, kind=kind, alpha="ramp" if kind == "video" else "opaque").numpy()
    assert gpu_blocks(lib, img, codec) == twin.encode(name, img)


@pytest.mark.parametrize("kind", ["video", "gradient", "edges"])
def test_chroma_refine_option_device_equals_host_build(lib, kind):
    """HAPB200_OPTION_CHROMA_REFINE selects the <.., true> kernels: bytes equal the host build with the refinement, differ from the
    default path where it matters, and the option goes back off."""
    img = synth.frame(512, 256, 1, kind=kind, alpha="ramp").numpy()
    plain = gpu_blocks(lib, img, L.HapB200Codec_HapY)
    assert lib.set_option(lib.OPTION_CHROMA_REFINE, 1) == 0
    try:
        fine = gpu_blocks(lib, img, L.HapB200Codec_HapY)
        both = gpu_blocks(lib, img, L.HapB200Codec_HapM)
    finally:
        lib.set_option(lib.OPTION_CHROMA_REFINE, 0)
    assert fine == twin.encode("ycocg_refine", img) and plain == twin.encode("ycocg", img)
    n0 = lib.texture_bytes(512, 256, L.HapB200Codec_HapM, 0)
    assert both[:n0] == fine and both[n0:] == twin.encode("bc4", img)
    if kind == "gradient":
        assert fine != plain
    assert gpu_blocks(lib, img, L.HapB200Codec_HapY) == plain


def test_hapm_writes_both_planes(lib):
    img = synth.frame(128, 64, 0, alpha="ramp").numpy()
    b = gpu_blocks(lib, img, L.HapB200Codec_HapM)
    n0 = lib.texture_bytes(128, 64, L.HapB200Codec_HapM, 0)
    assert b[:n0] == twin.encode("ycocg", img) and b[n0:] == twin.encode("bc4", img)


@pytest.mark.parametrize("name,codec", KINDS)
def test_psnr_bar_on_1080p(lib, name, codec):
    img = synth.frame(1920, 1080, 0, alpha="ramp").numpy()
    blk = gpu_blocks(lib, img, codec)
    if name == "bc4":
        assert oracles.psnr(img[..., 3:4], oracles.bc_decode("bc4", blk, 1920, 1080)[..., None], (0,)) > 45
        return
    crop = img[256:512, 512:1024]
    ours = oracles.psnr(img, oracles.bc_decode(name, blk, 1920, 1080))
    # oracle cluster fit on a crop (it is slow), ours on the same crop
    ob = oracles.bc_encode_clusterfit(name, crop, 8)
    tb = twin.encode(name, crop)
    po = oracles.psnr(crop, oracles.bc_decode(name, ob, 512, 256))
    pt = oracles.psnr(crop, oracles.bc_decode(name, tb, 512, 256))
    assert ours > 38 and pt >= po - 0.10, (name, ours, pt, po)


@pytest.mark.parametrize("name,codec", KINDS)
def test_block_decoder_bit_exact_against_oracle(lib, name, codec):
    rng = np.random.default_rng(2)
    w, h = 256, 128
    n = lib.texture_bytes(w, h, codec, 0)
    blocks = rng.integers(0, 256, n, dtype=np.uint8)
    d = torch.from_numpy(blocks).cuda()
    pad = torch.zeros((n + 15) // 16 * 16, dtype=torch.uint8, device="cuda")
    pad[:n] = d
    out = torch.zeros(h * w * 4, dtype=torch.uint8, device="cuda")
    assert lib.block_decode_batch(pad.data_ptr(), 1, pad.numel(), w, h, codec, out.data_ptr(), out.numel()) == 0
    got = out.cpu().numpy().reshape(h, w, 4)
    want = oracles.bc_decode(name, blocks.tobytes(), w, h)
    if name == "bc4":
        assert np.array_equal(got[..., 0], want) and np.array_equal(got[..., 1], want) and (got[..., 3] == 255).all()
    else:
        assert np.array_equal(got, want)


@pytest.mark.parametrize("name,codec,pixel_format", [("bc1", L.HapB200Codec_Hap1, "DXT1"), ("bc3", L.HapB200Codec_Hap5, "DXT5")])
def test_block_decoder_on_blocks_an_independent_encoder_wrote(lib, name, codec, pixel_format):
    """Blocks compressed by Pillow's own S3TC encoder (shares nothing with this repo or the oracle) decode on the GPU to exactly
    the pixels Pillow's own decoder gives."""
    import io
    from PIL import Image
    w, h = 512, 256
    img = np.ascontiguousarray(synth.frame(w, h, 3, alpha="ramp").numpy())
    if name == "bc1":
        img[..., 3] = 255
    buf = io.BytesIO()
    Image.fromarray(img, "RGBA").save(buf, format="DDS", pixel_format=pixel_format)
    dds = buf.getvalue()
    n = lib.texture_bytes(w, h, codec, 0)
    blocks = np.frombuffer(dds[len(dds) - n:], np.uint8)
    want = np.asarray(Image.open(io.BytesIO(dds)).convert("RGBA"))
    pad = torch.zeros((n + 15) // 16 * 16, dtype=torch.uint8, device="cuda")
    pad[:n] = torch.from_numpy(blocks.copy()).cuda()
    out = torch.zeros(h * w * 4, dtype=torch.uint8, device="cuda")
    assert lib.block_decode_batch(pad.data_ptr(), 1, pad.numel(), w, h, codec, out.data_ptr(), out.numel()) == 0
    got = out.cpu().numpy().reshape(h, w, 4)
    ch = (0, 1, 2) if name == "bc1" else (0, 1, 2, 3)
    assert np.array_equal(got[..., ch], want[..., ch])


def test_rgba_roundtrip_single_frame_host_pointers(lib):
    img = synth.frame(512, 256, 2, alpha="ramp").numpy()
    for codec, ch in ((L.HapB200Codec_Hap1, (0, 1, 2)), (L.HapB200Codec_Hap5, (0, 1, 2, 3)), (L.HapB200Codec_HapY, (0, 1, 2)),
                      (L.HapB200Codec_HapM, (0, 1, 2, 3))):
        r, frame = lib.encode_rgba(img, 512, 256, codec, 1, 4)
        assert r == 0
        r, rgba = lib.decode_rgba(frame, 512, 256)
        assert r == 0
        assert oracles.psnr(img, np.frombuffer(rgba, np.uint8).reshape(256, 512, 4), ch) > 33
