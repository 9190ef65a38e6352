#!/bin/bash
# 8-GPU call: the 16K stream at N=4 and N=8 (ring delivery + NCCL baseline) and the default bench at N=8 (short; delivery legs)
TAG=${1:-r02n}
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/${TAG}_gpus.txt
for N in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29550 + N)) bench.py --gpus $N --workload 16k_stream --steps 8 --warmup 4 > gpurun_out/${TAG}_stream_n$N.json 2> gpurun_out/${TAG}_stream_n$N.err
done
timeout 200 python bench.py --workload 16k_stream --steps 8 --warmup 4 > gpurun_out/${TAG}_stream_n1.json 2> gpurun_out/${TAG}_stream_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/${TAG}_bench_n8.json 2> gpurun_out/${TAG}_bench_n8.err
python - <<PY
import json
for n in (1,2,4,8):
    try:
        d=json.load(open("gpurun_out/${TAG}_stream_n%d.json"%n)); b=d.get("nccl_gatherv_baseline") or {}
        print(n, "ring fps %.0f ms %.3f enc-only %.3f nvlink %.0f GB/s | nccl fps %.0f" % (d["fps"], d["ms_per_step"], d["encode_only_ms_per_step"], d["nvlink_GBps_into_rank0"], b.get("fps",0)))
    except Exception as e: print(n, "failed", e)
try:
    d=json.load(open("gpurun_out/${TAG}_bench_n8.json")); print(d["value"], d["e2e"]["value"], json.dumps(d["extra"].get("delivery_to_rank0"))[:1200])
except Exception as e: print("bench n8 failed", e)
PY
tail -3 gpurun_out/${TAG}_stream_n8.err
