# fourth sweep: fewer / taller jobs for C2 x32 (the bench workload) and for the inverse of the RGB 9/7 configs
run() {  # env, config, frames
  env $1 timeout 300 python tools/config_bench.py --steps 20 --only $2 ${3:+--frames $3} 2>gpurun_out/r2ai.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d['key'] != '$2': continue
    print('$1', d['key'], d['frames'], 'fwd', round(d['fwd_frac_hbm'],4), 'inv', round(d['inv_frac_hbm'],4))
"
}
for v in "J2K_X=default" "J2K_RING_TARGET_JOBS=2368" "J2K_RING_TARGET_JOBS=3552" "J2K_RING_CHUNK=192" "J2K_RING_CHUNK=256" "J2K_RING_CHUNK_DEEP=64" "J2K_RING_CHUNK_DEEP=32"; do
  run "$v" C2 32; run "$v" C3i; run "$v" C5
done
