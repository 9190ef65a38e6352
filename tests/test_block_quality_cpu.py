"""Quality of the block-encoder algorithm (the source the CUDA kernels compile, run on the host) against
the oracle's squish-HIGH-class cluster fit.  north_star bar: PSNR within 0.1 dB.  All four formats are held to it."""
import numpy as np
import pytest

import oracles
import twin
from hap_b200 import synth

# (kind, allowed deficit in dB versus the oracle on the 512x512 synthetic video frame)
BARS = [("ycocg", 0.10), ("bc4", 0.10), ("bc1", 0.10), ("bc3", 0.10)]


@pytest.mark.parametrize("kind,deficit", BARS)
def test_psnr_against_cluster_fit(kind, deficit):
    img = synth.frame(512, 512, 0, alpha="ramp").numpy()
    ours = twin.encode(kind, img)
    theirs = oracles.bc_encode_clusterfit(kind, img, 8)
    if kind == "bc4":
        a, b = oracles.bc_decode("bc4", ours, 512, 512), oracles.bc_decode("bc4", theirs, 512, 512)
        ref = img[..., 3].astype(np.float64)
        p = lambda x: 10 * np.log10(255.0 ** 2 / max(((x - ref) ** 2).mean(), 1e-12))
        assert p(a) >= p(b) - deficit
        return
    ch = (0, 1, 2, 3) if kind == "bc3" else (0, 1, 2)
    pa = oracles.psnr(img, oracles.bc_decode(kind, ours, 512, 512), ch)
    pb = oracles.psnr(img, oracles.bc_decode(kind, theirs, 512, 512), ch)
    assert pa >= pb - deficit, (kind, pa, pb)


def test_flat_and_noise_blocks_are_sane():
    flat = synth.frame(64, 64, 0, kind="flat").numpy()
    for kind in ("bc1", "bc3", "ycocg"):
        dec = oracles.bc_decode(kind, twin.encode(kind, flat), 64, 64)
        assert np.abs(dec[..., :3].astype(int) - flat[..., :3].astype(int)).max() <= 4, kind
    noise = synth.frame(64, 64, 3, kind="noise").numpy()
    for kind in ("bc1", "bc3", "ycocg", "bc4"):
        blk = twin.encode(kind, noise)
        assert len(blk) == 256 * (8 if kind in ("bc1", "bc4") else 16)
