set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for fr in 55 222; do
  timeout 300 python bench.py --steps 6 --warmup 3 --frames $fr --profile > gpurun_out/ab_walk_$fr.json 2> gpurun_out/ab_walk_$fr.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_walk_$fr.json")); r=d["roofline"]
    print("RESULT walk $fr", "value %.1f ms/step %.3f"%(d["value"], d["ms_per_step"]), {k:round(v*55/$fr,3) for k,v in r["stage_ms_per_step"].items() if v>0.02}, r["decode_phase_share"], r["decode_counts"])
except Exception as e: print("RESULT FAILED", e)
PY
done
timeout 200 python tools/measure_ref_decode.py > gpurun_out/ref_stream_decode.json 2> gpurun_out/ref_stream_decode.err; cat gpurun_out/ref_stream_decode.json
