#!/bin/bash
# 2-GPU call: all GPU tests (the multi-GPU ones run with 2 devices), the 16K stream at N=1 and N=2 (ring delivery + NCCL baseline),
# the default bench at N=2 (delivery leg)
TAG=${1:-r02j}
mkdir -p gpurun_out
(time timeout 1700 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_tests.log 2>&1
tail -5 gpurun_out/${TAG}_tests.log
timeout 300 python bench.py --workload 16k_stream --steps 8 --warmup 3 > gpurun_out/${TAG}_stream_n1.json 2> gpurun_out/${TAG}_stream_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload 16k_stream --steps 8 --warmup 3 > gpurun_out/${TAG}_stream_n2.json 2> gpurun_out/${TAG}_stream_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
for f in stream_n1 stream_n2 bench_n2; do head -c 400 gpurun_out/${TAG}_$f.json; echo; tail -3 gpurun_out/${TAG}_$f.err; done
