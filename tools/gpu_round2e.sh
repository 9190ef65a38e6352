#!/bin/bash
TAG=${1:-r02e}
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_tests.log 2>&1
DEBUG_INDEX=1 timeout 600 python tools/debug_16k.py > gpurun_out/${TAG}_debug16k.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --profile --frames 64 --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"snappy_encode_fragments|snappy_execute|bc_encode|hap_place" -c 8 \
    -o gpurun_out/${TAG}_prof python bench.py --profile --frames 64 --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"snappy_execute|snappy_index" -c 2 \
    -o gpurun_out/${TAG}_prof_noindex python bench.py --profile --no-index --frames 64 --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_full_noindex.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "fragment_index and Hap5 or offset_table" > gpurun_out/${TAG}_memcheck.log 2>&1
tail -3 gpurun_out/${TAG}_tests.log
tail -8 gpurun_out/${TAG}_debug16k.log
