# job-count target / chunk cap of ring_chunks against the launch size: C3(ii) x32, C2 x16 / x32, C1 x256 (forward and inverse)
run() {  # env, config, frames
  env $1 timeout 300 python tools/config_bench.py --steps 20 --only $2 ${3:+--frames $3} 2>gpurun_out/r2ad.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d['key'] != '$2': continue
    print('$1', d['key'], d['frames'], 'fwd', round(d['fwd_frac_hbm'],4), 'inv', round(d['inv_frac_hbm'],4))
"
}
for v in "J2K_X=default" "J2K_RING_TARGET_JOBS=9472" "J2K_RING_TARGET_JOBS=18944" "J2K_RING_CHUNK=64" "J2K_RING_CHUNK=32" "J2K_RING_CHUNK=96"; do
  run "$v" C3ii; run "$v" C2 32; run "$v" C1; run "$v" C3i
done
