#!/bin/bash
# One GPU call of round 2: tests, bench (with and without the fragment index), ncu launch list and full captures of the two
# decode kernels and the fragment compressor, a sanitizer pass on a small case.  Everything lands in gpurun_out/.
TAG=${1:-r02b}
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_tests.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --steps 10 --warmup 3 --no-index --no-extra --no-cpu-baseline > gpurun_out/${TAG}_bench_noindex.json 2> gpurun_out/${TAG}_bench_noindex.err
# every launch with its device time (cold, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --profile --frames 64 --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_launches.log 2>&1
# full capture of the decode kernels and the fragment compressor (3 launches each)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"snappy_execute|snappy_index|snappy_encode_fragments" -s 6 -c 6 \
    -o gpurun_out/${TAG}_prof python bench.py --profile --frames 64 --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"snappy_execute|snappy_index" -s 4 -c 4 \
    -o gpurun_out/${TAG}_prof_noindex python bench.py --profile --no-index --frames 64 --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_full_noindex.log 2>&1
# memcheck + racecheck on one small indexed round trip
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "fragment_index and Hap5" > gpurun_out/${TAG}_memcheck.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "fragment_index and Hap5" > gpurun_out/${TAG}_racecheck.log 2>&1
tail -3 gpurun_out/${TAG}_tests.log
head -c 600 gpurun_out/${TAG}_bench.json
