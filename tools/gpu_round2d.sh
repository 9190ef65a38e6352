#!/bin/bash
# 2-GPU call: full GPU test suite (incl. the NCCL tests), the 16K debug helper, bench at N=1 and N=2, the 16k_stream workload at N=1 and N=2
TAG=${1:-r02d}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_tests.log 2>&1
timeout 600 python tools/debug_16k.py > gpurun_out/${TAG}_debug16k.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
python bench.py --workload 16k_stream --steps 6 --warmup 3 > gpurun_out/${TAG}_stream_n1.json 2> gpurun_out/${TAG}_stream_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload 16k_stream --steps 6 --warmup 3 > gpurun_out/${TAG}_stream_n2.json 2> gpurun_out/${TAG}_stream_n2.err
tail -4 gpurun_out/${TAG}_tests.log
tail -12 gpurun_out/${TAG}_debug16k.log
head -c 400 gpurun_out/${TAG}_stream_n2.json
