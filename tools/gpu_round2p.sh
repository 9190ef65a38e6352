#!/bin/bash
# last 1-GPU check of the round: GPU tests, smoke(), a short bench
TAG=${1:-r02p}
mkdir -p gpurun_out
(time timeout 400 python -m pytest tests -m gpu -x -q --timeout 200) > gpurun_out/${TAG}_tests.log 2>&1
tail -4 gpurun_out/${TAG}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
head -c 300 gpurun_out/${TAG}_bench.json; echo; tail -2 gpurun_out/${TAG}_bench.err
