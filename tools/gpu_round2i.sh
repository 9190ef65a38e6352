#!/bin/bash
# 1-GPU call: tests, bench with the chroma refinement (default build) and without it (variant build), block codecs alone
TAG=${1:-r02i}
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_tests.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
HAPB200_LIBRARY=$PWD/hap_b200/libhap_b200_norefine.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_norefine.json 2> gpurun_out/${TAG}_bench_norefine.err
python tools/measure_block_codecs.py 32 > gpurun_out/${TAG}_block_codecs.json 2> gpurun_out/${TAG}_block_codecs.err
HAPB200_LIBRARY=$PWD/hap_b200/libhap_b200_norefine.so python tools/measure_block_codecs.py 32 > gpurun_out/${TAG}_block_codecs_norefine.json 2>> gpurun_out/${TAG}_block_codecs.err
tail -3 gpurun_out/${TAG}_tests.log
head -c 300 gpurun_out/${TAG}_bench.json; echo
head -c 300 gpurun_out/${TAG}_bench_norefine.json; echo
cat gpurun_out/${TAG}_block_codecs.json gpurun_out/${TAG}_block_codecs_norefine.json
