#!/bin/bash
TAG=${1:-r02c}
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_tests.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --steps 6 --warmup 3 --no-index --no-extra --no-cpu-baseline > gpurun_out/${TAG}_bench_noindex.json 2> gpurun_out/${TAG}_bench_noindex.err
for v in v4 v3s; do
  HAPB200_LIBRARY=$PWD/hap_b200/libhap_b200_$v.so python bench.py --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/${TAG}_bench_$v.json 2> gpurun_out/${TAG}_bench_$v.err
  HAPB200_LIBRARY=$PWD/hap_b200/libhap_b200_$v.so python bench.py --steps 6 --warmup 3 --no-index --no-extra --no-cpu-baseline > gpurun_out/${TAG}_bench_${v}_noindex.json 2> gpurun_out/${TAG}_bench_${v}_noindex.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"snappy_execute|snappy_index" -s 2 -c 2 \
    -o gpurun_out/${TAG}_prof python bench.py --profile --frames 64 --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"snappy_execute|snappy_index" -s 2 -c 2 \
    -o gpurun_out/${TAG}_prof_noindex python bench.py --profile --no-index --frames 64 --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_full_noindex.log 2>&1
tail -3 gpurun_out/${TAG}_tests.log
head -c 300 gpurun_out/${TAG}_bench.json
