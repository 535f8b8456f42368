# sustained-load A/B of env knobs (power-capped regime): chunk height of big batches
run() { label=$1; shift
  env "$@" timeout 120 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-e2e --sustained-seconds 3 > gpurun_out/b.json 2> gpurun_out/b.err
  python - "$label" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/b.json")); s=d["sustained"]; print("%-14s burst %.4f alone %.4f inv %.4f sustained %.4f ms  %s MHz %s"%(sys.argv[1], d["ms_per_step"], d["roofline"]["kernel_ms"], d["inverse"]["ms_per_step"], s["ms_per_step"], s["clocks"]["sm_mhz"], s["clocks"]["reasons"]))
except Exception as e:
    print(sys.argv[1], "FAILED", open("gpurun_out/b.err").read()[-300:])
PY
}
run base A=1
run chunk128 J2K_RING_CHUNK=128
run chunk192 J2K_RING_CHUNK=192
run chunk256 J2K_RING_CHUNK=256
for c in 64 128 256; do echo CHUNK=$c; J2K_RING_CHUNK=$c timeout 200 python tools/config_bench.py --steps 10 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(' ', d['config'][:32], 'fwd', round(d['fwd_frac_hbm'],3), 'inv', round(d['inv_frac_hbm'],3), d['lossless_roundtrip_identical'])
"; done
