#!/bin/bash
# 2-GPU call: NCCL point-to-point channel settings for the stream delivery; ncu of the decode kernels on reference-made streams
TAG=${1:-r02h}
mkdir -p gpurun_out
run_stream() {  # name, env...
  local name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29530 + RANDOM % 50)) bench.py --gpus 2 --workload 16k_stream --steps 8 --warmup 3 > gpurun_out/${TAG}_stream_n2_$name.json 2> gpurun_out/${TAG}_stream_n2_$name.err
}
run_stream default HAPB200_DUMMY=1
run_stream ch16 NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=16
run_stream ch32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"snappy_execute|snappy_index" -c 2 \
    -o gpurun_out/${TAG}_prof_ref python tests/measure_ref_decode.py --frames 64 > gpurun_out/${TAG}_ncu_ref.log 2>&1
python tests/measure_ref_decode.py --frames 64 > gpurun_out/${TAG}_ref_decode.json 2> gpurun_out/${TAG}_ref_decode.err
for n in default ch16 ch32; do head -c 260 gpurun_out/${TAG}_stream_n2_$n.json; echo; done
cat gpurun_out/${TAG}_ref_decode.json | head -c 600
