#!/usr/bin/env python3
"""Block codec kernels alone on 4K frames (K1-K4 encode, K8 decode), per flavour: ms per frame and GB/s of
RGBA + texture traffic, from the library's own CUDA-event stage timer.  python tools/measure_block_codecs.py [frames]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hap_b200  # noqa: E402
import hap_b200.lib as L  # noqa: E402
from hap_b200 import synth  # noqa: E402

W, H = 3840, 2160


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    lib = hap_b200.load()
    dev = torch.device("cuda", 0)
    rgba = torch.empty((F, H, W, 4), dtype=torch.uint8, device=dev)
    for i in range(F):
        rgba[i] = synth.frame(W, H, i, device=dev, alpha="ramp")
    out = {}
    for name, codec in (("Hap1", L.HapB200Codec_Hap1), ("Hap5", L.HapB200Codec_Hap5), ("HapY", L.HapB200Codec_HapY),
                        ("HapM", L.HapB200Codec_HapM), ("HapA", L.HapB200Codec_HapA), ("HapY+chroma_refine", L.HapB200Codec_HapY)):
        lib.set_option(lib.OPTION_CHROMA_REFINE, 1 if name.endswith("refine") else 0)
        nb = lib.texture_bytes(W, H, codec, 0) + lib.texture_bytes(W, H, codec, 1)
        stride = (nb + 15) // 16 * 16
        blocks = torch.empty(F * stride, dtype=torch.uint8, device=dev)
        back = torch.empty_like(rgba)
        for _ in range(2):
            assert lib.block_encode_batch(rgba.data_ptr(), F, 4 * W * H, W, H, codec, blocks.data_ptr(), stride) == 0
            assert lib.block_decode_batch(blocks.data_ptr(), F, stride, W, H, codec, back.data_ptr(), 4 * W * H) == 0
        torch.cuda.synchronize()
        lib.set_stage_timing(True)
        lib.stage_times()
        for _ in range(3):
            assert lib.block_encode_batch(rgba.data_ptr(), F, 4 * W * H, W, H, codec, blocks.data_ptr(), stride) == 0
            assert lib.block_decode_batch(blocks.data_ptr(), F, stride, W, H, codec, back.data_ptr(), 4 * W * H) == 0
        st = lib.stage_times()
        lib.set_stage_timing(False)
        enc, dec = st["bc_encode"][0] / st["bc_encode"][1] / F, st["bc_decode"][0] / st["bc_decode"][1] / F
        traffic = 4 * W * H + nb
        out[name] = {"encode_us_per_frame": round(enc * 1e3, 2), "encode_GBps": round(traffic / (enc * 1e-3) / 1e9, 1),
                     "decode_us_per_frame": round(dec * 1e3, 2), "decode_GBps": round(traffic / (dec * 1e-3) / 1e9, 1)}
    lib.set_option(lib.OPTION_CHROMA_REFINE, 0)
    print(json.dumps({"what": "block codec kernels on 3840x2160 frames, %d frames per launch" % F, "per_flavour": out}))


if __name__ == "__main__":
    main()
