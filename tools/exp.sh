timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/b.json"))
print("value", d["value"], "ms", d["ms_per_step"], "roof", d["roofline"]["frac"], d["roofline"]["step_frac"], d["roofline"]["per_level_ms"])
print("inverse", d["inverse"]["value"], d["inverse"]["step_frac_of_hbm_peak"])
PY
python tools/config_bench.py --steps 10 2>&1 | cut -c1-250
