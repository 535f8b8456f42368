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
run perlevel J2K_RING_PER_LEVEL=1
run target2368 J2K_RING_TARGET_JOBS=2368
run target1776 J2K_RING_TARGET_JOBS=1776
run target1184 J2K_RING_TARGET_JOBS=1184
run target592 J2K_RING_TARGET_JOBS=592
run t1776min16 J2K_RING_TARGET_JOBS=1776 J2K_RING_CHUNK_MIN=16
run t1184min16 J2K_RING_TARGET_JOBS=1184 J2K_RING_CHUNK_MIN=16
run t1184min32 J2K_RING_TARGET_JOBS=1184 J2K_RING_CHUNK_MIN=32
run chunk128 J2K_RING_CHUNK=128
run chunk128t1184 J2K_RING_CHUNK=128 J2K_RING_TARGET_JOBS=1184 J2K_RING_CHUNK_MIN=16
run frames32 BENCH_FRAMES=32
