#!/bin/bash
# 1-GPU call, end of round 2: tests, bench, launch list of the bench command, block codecs, the 16K stream at N=1
TAG=${1:-r02o}
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -x -q --timeout 300) > gpurun_out/${TAG}_tests.log 2>&1
tail -4 gpurun_out/${TAG}_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python tools/measure_block_codecs.py 32 > gpurun_out/${TAG}_block_codecs.json 2> gpurun_out/${TAG}_block_codecs.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bc_|hap_|snappy_" -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --profile --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 200 python bench.py --workload 16k_stream --steps 8 --warmup 4 > gpurun_out/${TAG}_stream_n1.json 2> gpurun_out/${TAG}_stream_n1.err
head -c 300 gpurun_out/${TAG}_bench.json; echo
cat gpurun_out/${TAG}_block_codecs.json; echo
head -c 400 gpurun_out/${TAG}_stream_n1.json; echo
grep -c bc_encode gpurun_out/${TAG}_launches.csv
