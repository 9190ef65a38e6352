"""Debug helper (GPU): bench.py's extra-config leg for 16K, alone and after the other configurations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import hap_b200
from hap_b200.lib import HapB200Codec_HapY, HapB200Codec_HapM
lib = hap_b200.load()
lib.set_option(lib.OPTION_WRITE_INDEX, 1)
dev = torch.device("cuda")
stream = torch.cuda.Stream()
sp = stream.cuda_stream


def one(w, h, codec, chunks, F, alpha="opaque", on_stream=True, reps=4):
    rt = bench.Roundtrip(lib, dev, w, h, codec, chunks, F, alpha=alpha)
    for rep in range(reps):
        rt.used.zero_()
        torch.cuda.synchronize()
        if on_stream:
            with torch.cuda.stream(stream):
                rt.encode(sp)
        else:
            rt.encode(None)
        torch.cuda.synchronize()
        print(f"  {w}x{h} F={F} stream={on_stream} rep {rep}: used {rt.used.tolist()} free {torch.cuda.mem_get_info()[0] >> 20} MiB", flush=True)
    with torch.cuda.stream(stream):
        rt.decode(sp)
    torch.cuda.synchronize()
    print("   decode res", rt.res.tolist(), "sizes", rt.tex_used.tolist(), flush=True)
    del rt
    torch.cuda.empty_cache()


print("16K alone, default stream"); one(16384, 16384, HapB200Codec_HapY, 64, 3, on_stream=False)
print("16K alone, side stream"); one(16384, 16384, HapB200Codec_HapY, 64, 3)
print("8K then 16K"); one(7680, 4320, HapB200Codec_HapM, 32, 12, alpha="ramp"); one(16384, 16384, HapB200Codec_HapY, 64, 3)
