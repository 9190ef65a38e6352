"""Parity at BASELINE.json's full sizes on PICTURE content, closing the holes of the round-1 review:
  (a) frames of configs 1-4 encoded by the REFERENCE (hap.c + Google Snappy: byte-granular, multi-window, 60 KB
      offsets) and by the oracle port, decoded by HapDecode and HapB200DecodeBatch -> bytes equal to the payload;
  (b) the 16K frame encoded on the GPU, decoded by the reference build (not by the CUDA path itself);
  (c) the block encoders' PSNR against the oracle cluster fit on four content classes and full 1080p frames.
Bit-exact for (a) and (b); (c) is the one floating-point comparison of the repo, bars written per class below."""
import numpy as np
import pytest
import torch

import hap_b200
import hap_b200.lib as L
import oracles
from hap_b200 import synth

pytestmark = pytest.mark.gpu

CONFIGS = [  # name, w, h, codec, chunks, alpha
    ("1080p_dxt1_x1", 1920, 1080, L.HapB200Codec_Hap1, 1, "opaque"),
    ("4k_dxt1_x1", 3840, 2160, L.HapB200Codec_Hap1, 1, "opaque"),
    ("4k_ycocg_x8", 3840, 2160, L.HapB200Codec_HapY, 8, "opaque"),
    ("8k_hapm_x32", 7680, 4320, L.HapB200Codec_HapM, 32, "ramp"),
]
FMT = {L.HapB200Codec_Hap1: [hap_b200.HapTextureFormat_RGB_DXT1], L.HapB200Codec_HapY: [hap_b200.HapTextureFormat_YCoCg_DXT5],
       L.HapB200Codec_HapM: [hap_b200.HapTextureFormat_YCoCg_DXT5, hap_b200.HapTextureFormat_A_RGTC1]}


@pytest.fixture(scope="module")
def lib():
    return hap_b200.load()


def textures_of(lib, w, h, codec, alpha, index=0):
    """DXT textures of one synthetic picture (block encoder on the GPU; here they are just realistic payload bytes)."""
    img = synth.frame(w, h, index, alpha=alpha, device="cuda")
    n0, n1 = lib.texture_bytes(w, h, codec, 0), lib.texture_bytes(w, h, codec, 1)
    buf = torch.zeros((n0 + n1 + 15) // 16 * 16, dtype=torch.uint8, device="cuda")
    assert lib.block_encode_batch(img.data_ptr(), 1, img.numel(), w, h, codec, buf.data_ptr(), buf.numel()) == 0
    host = buf.cpu().numpy()
    return [host[:n0].tobytes()] + ([host[n0:n0 + n1].tobytes()] if n1 else [])


@pytest.mark.parametrize("name,w,h,codec,k,alpha", CONFIGS)
@pytest.mark.parametrize("encoder", ["reference", "oracle"])
def test_foreign_full_size_frames_decode_to_their_payload(lib, name, w, h, codec, k, alpha, encoder):
    enc = oracles.ref_abi() if encoder == "reference" else oracles.oracle_abi()
    if enc is None:
        pytest.skip("reference build not available")
    tex = textures_of(lib, w, h, codec, alpha)
    fmts = FMT[codec]
    r, frame = enc.encode(tex, fmts, [1] * len(tex), [k] * len(tex))
    assert r == 0 and len(frame) < sum(len(t) for t in tex)        # picture content compresses: Complex storage, Snappy chunks
    # host-pointer HapDecode, every texture
    for i, t in enumerate(tex):
        r, data, fmt, calls = lib.decode(frame, i, len(t))
        assert (r, fmt) == (0, fmts[i]) and data == t, (name, encoder, i)
        assert calls == ([k] if k > 1 else [])
    # device-resident batch of 3 copies (one of them truncated: its neighbours must be unaffected)
    F = 3
    cap = (len(frame) + 63) // 64 * 64
    buf = torch.zeros(F * cap, dtype=torch.uint8, device="cuda")
    fr = torch.frombuffer(bytearray(frame), dtype=torch.uint8).cuda()
    for f in range(F):
        buf[f * cap: f * cap + len(frame)] = fr
    used = torch.tensor([len(frame), len(frame) - 9, len(frame)], dtype=torch.int64, device="cuda")
    for i, t in enumerate(tex):
        n = len(t)
        stride = (n + 15) // 16 * 16
        out = torch.zeros(F * stride, dtype=torch.uint8, device="cuda")
        o_used = torch.zeros(F, dtype=torch.int64, device="cuda")
        o_fmt = torch.zeros(F, dtype=torch.int32, device="cuda")
        res = torch.full((F,), 9, dtype=torch.int32, device="cuda")
        assert lib.decode_batch(buf.data_ptr(), F, cap, used.data_ptr(), i, k, out.data_ptr(), stride, o_used.data_ptr(),
                                o_fmt.data_ptr(), res.data_ptr()) == 0
        rl = res.tolist()
        assert rl[0] == 0 and rl[2] == 0 and rl[1] != 0, rl
        want = torch.frombuffer(bytearray(t), dtype=torch.uint8).cuda()
        assert torch.equal(out[:n], want) and torch.equal(out[2 * stride: 2 * stride + n], want), (name, encoder, i)


def test_16k_frame_encoded_here_decodes_in_the_reference(lib):
    """Config 5's frame: 16384 x 16384 RGBA -> Hap Q, 64 chunks on the GPU; the 150 MB frame is decoded by the
    unmodified reference (all 64 chunks) and compared with the texture the block encoder wrote."""
    ref = oracles.ref_abi() or oracles.oracle_abi()
    w = h = 16384
    codec, k = L.HapB200Codec_HapY, 64
    free, _ = torch.cuda.mem_get_info()
    if free < 6 * 2 ** 30:
        pytest.skip("not enough free device memory")
    img = torch.empty((h, w, 4), dtype=torch.uint8, device="cuda")
    for ty in range(0, h, 2048):
        for tx in range(0, w, 4096):
            img[ty:ty + 2048, tx:tx + 4096] = synth.frame(4096, 2048, (ty // 2048) * 4 + tx // 4096, device="cuda")
    n = lib.texture_bytes(w, h, codec)
    blocks = torch.zeros(n, dtype=torch.uint8, device="cuda")
    assert lib.block_encode_batch(img.data_ptr(), 1, img.numel(), w, h, codec, blocks.data_ptr(), n) == 0
    cap = (lib.max_encoded_length_rgba(w, h, codec, k) + 15) // 16 * 16
    out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    used = torch.zeros(1, dtype=torch.int64, device="cuda")
    assert lib.encode_rgba_batch(img.data_ptr(), 1, img.numel(), w, h, codec, 1, k, out.data_ptr(), cap, used.data_ptr()) == 0
    del img
    frame = out[: int(used[0])].cpu().numpy()
    assert lib.chunk_count(frame, 0) == (0, 64)
    host_out = np.zeros(n, np.uint8)
    r, got, fmt, calls = ref.decode(frame, 0, n, out=host_out)
    assert (r, got, fmt, calls) == (0, n, hap_b200.HapTextureFormat_YCoCg_DXT5, [64])
    assert np.array_equal(host_out, blocks.cpu().numpy())


# (c) PSNR against the oracle cluster fit (8 iterations), full 1080p frames, per content class.
# north_star bar: within 0.1 dB.  Bars that are wider than that are the MEASURED deficits of the current encoders
# (DESIGN.md section 4 lists them); they are written here so that a regression shows, not to claim the target.
CLASSES = ["video", "gradient", "texture", "edges"]
# measured on these frames with the host build of the encoder source (device bytes == host bytes, tests/test_gpu_blocks.py):
#   ycocg: video +0.61, gradient -0.47 (!), texture +0.15, edges -0.13;  bc1/bc3: +0.03, -0.06, -0.08, -0.01;  bc4: equal
BAR = {("ycocg", "video"): 0.10, ("ycocg", "gradient"): 0.55, ("ycocg", "texture"): 0.10, ("ycocg", "edges"): 0.20,
       ("bc1", "video"): 0.10, ("bc1", "gradient"): 0.10, ("bc1", "texture"): 0.10, ("bc1", "edges"): 0.10,
       ("bc3", "video"): 0.10, ("bc3", "gradient"): 0.10, ("bc3", "texture"): 0.10, ("bc3", "edges"): 0.10,
       ("bc4", "video"): 0.10, ("bc4", "gradient"): 0.10, ("bc4", "texture"): 0.10, ("bc4", "edges"): 0.10}
KIND_CODEC = {"bc1": L.HapB200Codec_Hap1, "bc3": L.HapB200Codec_Hap5, "ycocg": L.HapB200Codec_HapY, "bc4": L.HapB200Codec_HapA}


def _psnr(kind, img, blocks, w, h):
    if kind == "bc4":
        ref = img[..., 3].astype(np.float64)
        d = oracles.bc_decode("bc4", blocks, w, h).astype(np.float64) - ref
        return 10 * np.log10(255.0 ** 2 / max((d * d).mean(), 1e-12))
    return oracles.psnr(img, oracles.bc_decode(kind, blocks, w, h), (0, 1, 2, 3) if kind == "bc3" else (0, 1, 2))


@pytest.mark.parametrize("cls", CLASSES)
@pytest.mark.parametrize("kind", ["ycocg", "bc1", "bc3", "bc4"])
def test_psnr_bar_full_1080p_per_content_class(lib, kind, cls):
    from concurrent.futures import ThreadPoolExecutor
    w, h = 1920, 1080
    img = synth.frame(w, h, 1, kind=cls, alpha="ramp").numpy()
    d = torch.from_numpy(np.ascontiguousarray(img)).cuda()
    codec = KIND_CODEC[kind]
    n = lib.texture_bytes(w, h, codec, 0)
    out = torch.zeros((n + 15) // 16 * 16, dtype=torch.uint8, device="cuda")
    assert lib.block_encode_batch(d.data_ptr(), 1, d.numel(), w, h, codec, out.data_ptr(), out.numel()) == 0
    ours = out[:n].cpu().numpy().tobytes()
    rows = [(y, min(y + 40, h)) for y in range(0, h, 40)]
    with ThreadPoolExecutor() as pool:
        theirs = b"".join(pool.map(lambda r: oracles.bc_encode_clusterfit(kind, np.ascontiguousarray(img[r[0]:r[1]]), 8), rows))
    pa, pb = _psnr(kind, img, ours, w, h), _psnr(kind, img, theirs, w, h)
    assert pa >= pb - BAR[(kind, cls)], (kind, cls, pa, pb)
    if kind == "ycocg":
        # with HAPB200_OPTION_CHROMA_REFINE every class is inside the north-star bar except `edges` (-0.13 dB: three-valued
        # luma blocks where BC4's 6-value mode happens to fit better, which the oracle tries and this encoder does not)
        assert lib.set_option(lib.OPTION_CHROMA_REFINE, 1) == 0
        try:
            assert lib.block_encode_batch(d.data_ptr(), 1, d.numel(), w, h, codec, out.data_ptr(), out.numel()) == 0
        finally:
            lib.set_option(lib.OPTION_CHROMA_REFINE, 0)
        pr = _psnr(kind, img, out[:n].cpu().numpy().tobytes(), w, h)
        assert pr >= pb - (0.20 if cls == "edges" else 0.10) and pr >= pa - 0.005, (kind, cls, pr, pa, pb)
