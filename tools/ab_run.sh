set -x
mkdir -p gpurun_out
for th in 8 16 32; do
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-threads $th --e2e-frames 64 > gpurun_out/e2e_$th.json 2> gpurun_out/e2e_$th.err
python - <<PY
import json
d=json.load(open("gpurun_out/e2e_$th.json")); print("threads $th value %.1f e2e %.2f"%(d["value"], d["e2e"]["value"]), {k:round(v,3) for k,v in d["roofline"]["stage_ms_per_step"].items() if v>0.05})
PY
done
