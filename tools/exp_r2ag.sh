# third sweep: job-count target / chunk cap for the 512-wide 5/3 configs and the general-alignment variants
run() {  # env, config
  env $1 timeout 300 python tools/config_bench.py --steps 20 --only $2 2>gpurun_out/r2ag.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d['key'] != '$2': continue
    print('$1', d['key'], d['frames'], 'fwd', round(d['fwd_frac_hbm'],4), 'inv', round(d['inv_frac_hbm'],4))
"
}
for v in "J2K_X=default" "J2K_RING_TARGET_JOBS=2368" "J2K_RING_TARGET_JOBS=3552" "J2K_RING_TARGET_JOBS=7104" "J2K_RING_CHUNK=64" "J2K_RING_CHUNK=32"; do
  run "$v" C1; run "$v" DX; run "$v" CR; run "$v" C5
done
