# launch-tail experiment: chunk height of the deep levels / job target (env knobs of ring_chunks), resident legs only
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err
  python - "$label" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/b.json")); print("%-22s"%sys.argv[1], "fwd step %.4f alone %.4f (frac %.3f) single %.4f inv %.4f"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], sorted(d["roofline"]["kernel_ms_single_launches"])[len(d["roofline"]["kernel_ms_single_launches"])//2], d["inverse"]["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "FAILED", open("gpurun_out/b.err").read()[-300:])
PY
}
run base A=1
run deep32 J2K_RING_CHUNK_DEEP=32
run deep16 J2K_RING_CHUNK_DEEP=16
run deep8 J2K_RING_CHUNK_DEEP=8 J2K_RING_CHUNK_MIN=4
run tgt2x J2K_RING_TARGET_JOBS=9472
run tgt4x_deep32 J2K_RING_TARGET_JOBS=18944 J2K_RING_CHUNK_DEEP=32
run chunk48 J2K_RING_CHUNK=48
run chunk96 J2K_RING_CHUNK=96
run base2 A=1
