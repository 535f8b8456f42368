# quick GPU iteration: parity suite, resident bench (no CPU leg), per-config table
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/b.json 2> gpurun_out/b.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/b.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"],4), "kernel-alone frac", round(d["roofline"]["frac"],4), "step frac", round(d["roofline"]["step_frac"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4))
if d.get("inverse"): print("inverse", round(d["inverse"]["value"]), round(d["inverse"]["step_frac_of_hbm_peak"],4))
print("e2e", d["e2e"])
PY
[ -n "$CFG" ] && python tools/config_bench.py --steps 10 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config'][:40], 'fwd', round(d['fwd_frac_hbm'],3), 'inv', round(d['inv_frac_hbm'],3), 'fwd_ms', round(d['fwd_ms'],4), 'inv_ms', round(d['inv_ms'],4))
"
true
