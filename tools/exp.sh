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
run max96al32 J2K_RING_STRIP_MAX=96 J2K_RING_STRIP_ALIGN=32
run max112al16 J2K_RING_STRIP_MAX=112 J2K_RING_STRIP_ALIGN=16
run max120al8 J2K_RING_STRIP_MAX=120 J2K_RING_STRIP_ALIGN=8
run max64al32 J2K_RING_STRIP_MAX=64 J2K_RING_STRIP_ALIGN=32
run max96al32_pl J2K_RING_STRIP_MAX=96 J2K_RING_STRIP_ALIGN=32 J2K_RING_PER_LEVEL=1
