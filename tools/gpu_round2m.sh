#!/bin/bash
# 2-GPU call: the multi-GPU tests, the 16K stream at N=2 (ring delivery + NCCL baseline), the default bench at N=2 (delivery legs)
TAG=${1:-r02m}
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 300) > gpurun_out/${TAG}_tests.log 2>&1
tail -4 gpurun_out/${TAG}_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload 16k_stream --steps 8 --warmup 3 > gpurun_out/${TAG}_stream_n2.json 2> gpurun_out/${TAG}_stream_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
for f in stream_n2 bench_n2; do head -c 300 gpurun_out/${TAG}_$f.json; echo; tail -3 gpurun_out/${TAG}_$f.err; done
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_stream_n2.json")); print({k:d[k] for k in ("value","fps","ms_per_step","encode_only_ms_per_step","nvlink_GBps_into_rank0")}, {k:v for k,v in d["nccl_gatherv_baseline"].items() if k!="what"})
d=json.load(open("gpurun_out/${TAG}_bench_n2.json")); print(d["value"], json.dumps(d["extra"].get("delivery_to_rank0"))[:1500])
PY
