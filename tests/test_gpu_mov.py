"""GPU-encoded Hap streams through the container layer, checked by a third-party decoder: frames written by
HapB200EncodeRGBA (Snappy, chunked) go into a QuickTime file (include/hap_mov.h); FFmpeg -- its own mov demuxer,
its own Snappy decoder, its own Hap section parser and block decoder, reached through OpenCV -- must decode them to
the picture our decode path produces."""
import numpy as np
import pytest

import hap_b200
import hap_b200.lib as L
import oracles
from hap_b200 import mov, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return hap_b200.load()


@pytest.mark.parametrize("codec,cc,kind,chunks", [(L.HapB200Codec_Hap1, "Hap1", "bc1", 1), (L.HapB200Codec_Hap5, "Hap5", "bc3", 4),
                                                  (L.HapB200Codec_HapY, "HapY", "ycocg", 8), (L.HapB200Codec_HapM, "HapM", "ycocg", 4),
                                                  (L.HapB200Codec_HapA, "HapA", "bc4", 2)])
@pytest.mark.parametrize("with_index", [0, 1])
def test_ffmpeg_decodes_gpu_encoded_movie(lib, tmp_path, codec, cc, kind, chunks, with_index):
    """with_index = 1: the frames carry the trailing fragment index section (hap_index.h); FFmpeg must not mind."""
    cv2 = pytest.importorskip("cv2")
    lib.set_option(lib.OPTION_WRITE_INDEX, with_index)
    try:
        _ffmpeg_decodes_gpu_encoded_movie(cv2, lib, tmp_path, codec, cc, kind, chunks)
    finally:
        lib.set_option(lib.OPTION_WRITE_INDEX, 0)


def _ffmpeg_decodes_gpu_encoded_movie(cv2, lib, tmp_path, codec, cc, kind, chunks):
    w, h, n = 512, 256, 4
    path = str(tmp_path / f"{cc}.mov")
    imgs, frames = [], []
    with mov.MovWriter(path, cc, w, h, 600) as wr:
        for i in range(n):
            img = synth.frame(w, h, i, alpha="ramp").numpy()
            r, f = lib.encode_rgba(img, w, h, codec, hap_b200.HapCompressorSnappy, chunks)
            assert r == 0 and mov.fourcc_for_frame(f) == (0, cc)
            assert lib.chunk_count(f, 0) == (0, chunks)
            assert wr.write(f, 20) == 0
            imgs.append(img)
            frames.append(f)
    # our own reader + decoder
    with mov.MovReader(path) as rd:
        assert (rd.fourcc, rd.width, rd.height, rd.frames) == (cc, w, h, n)
        texs = []
        for i in range(n):
            r, f, _ = rd.read(i)
            assert r == 0 and f == frames[i]
            r, tex, _, _ = lib.decode(f, 0, lib.texture_bytes(w, h, codec, 0))
            assert r == 0
            texs.append(tex)
    # FFmpeg
    cap = cv2.VideoCapture(path, cv2.CAP_FFMPEG)
    if not cap.isOpened():
        pytest.skip("this OpenCV build cannot open QuickTime files")
    try:
        for i in range(n):
            ok, bgr = cap.read()
            if not ok and i == 0:
                pytest.skip("this OpenCV/FFmpeg build has no Hap decoder")
            assert ok
            if kind == "bc4":   # alpha-only: FFmpeg shows the one channel as grey
                ref = oracles.bc_decode("bc4", texs[i], w, h)
                assert np.abs(bgr[..., 0].astype(int) - ref.astype(int)).max() <= 1, (cc, i)
                continue
            ref = oracles.bc_decode(kind, texs[i], w, h)[..., :3][..., ::-1]
            assert np.abs(bgr.astype(int) - ref.astype(int)).max() <= 3, (cc, i)
            assert oracles.psnr(np.ascontiguousarray(imgs[i][..., :3][..., ::-1]), bgr) > 30
    finally:
        cap.release()
