#!/usr/bin/env bash
# What is run on a B200 box before a round is closed (through `gpurun -- 'bash tools/gpu_validate.sh'`): the GPU test
# suite, the smoke entry, the default bench line, the ncu launch list of the bench command, one `ncu --set full`
# capture of the main kernels (55 frames per launch) and the reference-stream decode measurement.  Everything lands
# in gpurun_out/; tools/ncu_summary.py, tools/src_hotspots.py turn the captures into the files kept under profiles/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; head -c 400 gpurun_out/bench_default.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:hapb200|hap_|snappy_|bc_' -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --profile --no-overlap > gpurun_out/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:snappy_decode_chunks|snappy_encode_fragments|bc_encode_kernel|hap_place_fragments' -c 4 -f -o gpurun_out/prof_final python bench.py --steps 1 --warmup 3 --frames 55 --profile --no-overlap > gpurun_out/ncu_final.log 2>&1
timeout 200 python tests/measure_ref_decode.py > gpurun_out/ref_stream_decode.json 2> gpurun_out/ref_stream_decode.err
echo done
