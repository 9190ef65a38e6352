#!/bin/bash
TAG=${1:-r02f}
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_tests.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 900 python tools/debug_16k_bench.py > gpurun_out/${TAG}_debug16k_bench.log 2>&1
tail -3 gpurun_out/${TAG}_tests.log
tail -25 gpurun_out/${TAG}_debug16k_bench.log
