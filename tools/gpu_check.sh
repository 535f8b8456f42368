# full GPU check: parity suite, smoke, default bench (both arms)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo bench rc=$?
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_full.json"))
print("value", d["value"], "ms", d["ms_per_step"], "roof", d["roofline"]["frac"], d["roofline"]["step_frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], d["clocks"])
print("inverse", d["inverse"]["value"], d["inverse"]["step_frac_of_hbm_peak"])
r=json.load(open("gpurun_out/bench_ref.json")); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
lscpu | grep -E "Model name|^CPU\(s\)|Socket|Thread" 
