#!/bin/bash
# 8-GPU call: the 16k_stream workload at N = 8, 4, 2 (N = 1 was measured on a 1-GPU box)
TAG=${1:-r02g}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --workload 16k_stream --steps 8 --warmup 3 > gpurun_out/${TAG}_stream_n$N.json 2> gpurun_out/${TAG}_stream_n$N.err
done
timeout 300 python bench.py --workload 16k_stream --steps 8 --warmup 3 > gpurun_out/${TAG}_stream_n1.json 2> gpurun_out/${TAG}_stream_n1.err
for N in 8 4 2 1; do head -c 300 gpurun_out/${TAG}_stream_n$N.json; echo; done
