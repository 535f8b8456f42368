timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ring4.json 2> gpurun_out/bench_ring4.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_ring4.json"))
print("value", d["value"], "ms", d["ms_per_step"], "roof", d["roofline"]["frac"], d["roofline"]["step_frac"], d["roofline"]["per_level_ms"])
print("inverse", d["inverse"]); print("e2e", d["e2e"])
PY
tail -3 gpurun_out/bench_ring4.err
J2K_RING_INV_DISABLE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2>gpurun_out/b.err; python -c "
import json;d=json.load(open('gpurun_out/b.json'));print('old inverse path', d['inverse'])"
