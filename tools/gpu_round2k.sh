#!/bin/bash
# 2-GPU call: GPU tests, block codecs (default build + register-capped variants), the 16K stream at N=1 and N=2
TAG=${1:-r02k}
mkdir -p gpurun_out
(time timeout 1700 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_tests.log 2>&1
tail -5 gpurun_out/${TAG}_tests.log
python tools/measure_block_codecs.py 32 > gpurun_out/${TAG}_block_codecs.json 2> gpurun_out/${TAG}_block_codecs.err
for v in mb6 mb5; do
  HAPB200_LIBRARY=$PWD/hap_b200/libhap_b200_$v.so python tools/measure_block_codecs.py 32 > gpurun_out/${TAG}_block_codecs_$v.json 2>> gpurun_out/${TAG}_block_codecs.err
done
timeout 300 python bench.py --workload 16k_stream --steps 8 --warmup 3 > gpurun_out/${TAG}_stream_n1.json 2> gpurun_out/${TAG}_stream_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload 16k_stream --steps 8 --warmup 3 > gpurun_out/${TAG}_stream_n2.json 2> gpurun_out/${TAG}_stream_n2.err
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_block_codecs*.json
for f in stream_n1 stream_n2 bench; do head -c 300 gpurun_out/${TAG}_$f.json; echo; tail -2 gpurun_out/${TAG}_$f.err; done
