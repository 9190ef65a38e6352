set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 6 --warmup 3 --profile > gpurun_out/b2.json 2> gpurun_out/b2.err
cat gpurun_out/b2.json | head -c 1500
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:snappy_decode_chunks|snappy_encode_fragments|bc_encode_kernel' -c 3 -f -o gpurun_out/prof_r01b python bench.py --steps 1 --warmup 3 --profile --no-overlap > gpurun_out/ncu2.log 2>&1
tail -5 gpurun_out/ncu2.log
ls -la gpurun_out
