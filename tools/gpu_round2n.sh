#!/bin/bash
# 8-GPU call: the 16K stream at N=8 and N=4 (ring delivery as value, NCCL gatherv as the baseline in the same line)
TAG=${1:-r02n}
mkdir -p gpurun_out
for N in 8 4; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29550 + N)) bench.py --gpus $N --workload 16k_stream --steps 8 --warmup 4 > gpurun_out/${TAG}_stream_n$N.json 2> gpurun_out/${TAG}_stream_n$N.err
done
python - <<PY
import json
for n in (4,8):
    try:
        d=json.load(open("gpurun_out/${TAG}_stream_n%d.json"%n)); b=d.get("nccl_gatherv_baseline") or {}
        print(n, "ring fps %.0f ms %.3f enc-only %.3f nvlink %.0f GB/s | nccl fps %.0f" % (d["fps"], d["ms_per_step"], d["encode_only_ms_per_step"], d["nvlink_GBps_into_rank0"], b.get("fps",0)))
    except Exception as e: print(n, "failed", e)
PY
tail -3 gpurun_out/${TAG}_stream_n8.err
