run() { # label, env...
  label=$1; shift
  env "$@" > gpurun_out/b.json 2> gpurun_out/b.err
  python - "$label" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/b.json")); print(sys.argv[1], "ms/step %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["step_frac"], {k:round(v,4) for k,v in d["roofline"]["per_level_ms"].items()})
except Exception as e:
    print(sys.argv[1], "FAILED", open("gpurun_out/b.err").read()[-300:])
PY
}
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run s1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --streams 1
run s2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --streams 2
run s3 timeout 300 python bench.py --steps 21 --warmup 3 --no-cpu-baseline --no-e2e --streams 3
run s2f16 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --streams 2 --frames 16
