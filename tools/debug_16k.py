"""Debug helper (GPU): why does a batch of 16K frames of synth content not decode in bench.py's extra leg?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, hap_b200
from hap_b200 import synth
from hap_b200.lib import HapB200Codec_HapY
lib = hap_b200.load()
W = H = 16384
INDEX = int(os.environ.get("DEBUG_INDEX", "1"))
lib.set_option(lib.OPTION_WRITE_INDEX, INDEX)
print("write index", INDEX, flush=True)
for F in (1, 3):
    dev = torch.device("cuda")
    rgba = torch.empty((F, H, W, 4), dtype=torch.uint8, device=dev)
    for i in range(F):
        rgba[i] = synth.frame(W, H, i, device=dev)
    n = lib.texture_bytes(W, H, HapB200Codec_HapY)
    cap = (lib.max_encoded_length_rgba(W, H, HapB200Codec_HapY, 64) + 15) // 16 * 16
    frames = torch.zeros(F * cap, dtype=torch.uint8, device=dev)
    used = torch.zeros(F, dtype=torch.int64, device=dev)
    r = lib.encode_rgba_batch(rgba.data_ptr(), F, 4 * W * H, W, H, HapB200Codec_HapY, 1, 64, frames.data_ptr(), cap, used.data_ptr())
    torch.cuda.synchronize()
    print("F", F, "encode r", r, "cap", cap, "used", used.tolist(), flush=True)
    for f in range(F):
        head = frames[f * cap: f * cap + 32].cpu().numpy().tobytes()
        print(" frame", f, head.hex(), "chunks", lib.chunk_count(frames[f * cap: f * cap + int(used[f])].cpu().numpy(), 0), flush=True)
    tex = torch.zeros(F * n, dtype=torch.uint8, device=dev)
    tu = torch.zeros(F, dtype=torch.int64, device=dev)
    tf = torch.zeros(F, dtype=torch.int32, device=dev)
    res = torch.full((F,), 9, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream()
    for rep in range(3):
        tu.zero_(); res.fill_(9)
        torch.cuda.synchronize()
        r = lib.decode_batch(frames.data_ptr(), F, cap, used.data_ptr(), 0, 64, tex.data_ptr(), n, tu.data_ptr(), tf.data_ptr(), res.data_ptr(),
                             stream=st.cuda_stream if rep else None)
        torch.cuda.synchronize()
        print(" decode rep", rep, "r", r, "res", res.tolist(), "used", tu.tolist(), "fmt", tf.tolist(), flush=True)
    blocks = torch.zeros(n, dtype=torch.uint8, device=dev)
    for f in range(F):
        lib.block_encode_batch(rgba[f].data_ptr(), 1, 4 * W * H, W, H, HapB200Codec_HapY, blocks.data_ptr(), n)
        print("  frame", f, "texture equal", bool(torch.equal(blocks, tex[f * n:(f + 1) * n])), flush=True)
    del rgba, frames, tex, blocks
    torch.cuda.empty_cache()
