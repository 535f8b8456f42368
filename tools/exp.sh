run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err
  python - "$label" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/b.json")); print("%-28s"%sys.argv[1], "fwd step %.4f alone %.4f (frac %.3f) inv %.4f"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["inverse"]["ms_per_step"]), {k:round(v,4) for k,v in d["roofline"]["per_level_ms"].items()})
except Exception as e:
    print(sys.argv[1], "FAILED", open("gpurun_out/b.err").read()[-300:])
PY
}
run base A=1
run chunk128 J2K_RING_CHUNK=128
run chunk96 J2K_RING_CHUNK=96
run chunk128deep64 J2K_RING_CHUNK=128 J2K_RING_CHUNK_DEEP=64
run chunk48 J2K_RING_CHUNK=48
run perlevel J2K_RING_PER_LEVEL=1
