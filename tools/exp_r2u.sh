# fwd3w_kernel: batch-size asymptote of C3(i) (frames per launch) and chunk height on C5, against the component-split jobs
run() {  # env, config, frames
  env $1 timeout 300 python tools/config_bench.py --steps 20 --only $2 --frames $3 2>gpurun_out/r2u.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d['key'] != '$2': continue
    print('$1', d['key'], 'frames', d['frames'], 'fwd', round(d['fwd_frac_hbm'],4), 'fwd_ms', round(d['fwd_ms'],4))
"
}
for f in 8 16 32; do
  run "J2K_FWD3W=0" C3i $f
  run "J2K_FWD3W=1 J2K_FWD3W_TDIV=4" C3i $f
  run "J2K_FWD3W=1 J2K_FWD3W_TDIV=8" C3i $f
  run "J2K_FWD3W=1 J2K_FWD3W_TDIV=8 J2K_RING_CHUNK=256" C3i $f
done
run "J2K_FWD3W=1 J2K_FWD3W_TDIV=8 J2K_RING_CHUNK=256" C5 1
run "J2K_FWD3W=1 J2K_FWD3W_TDIV=8 J2K_RING_CHUNK=512" C5 1
run "J2K_FWD3W=1 J2K_FWD3W_TDIV=8 J2K_RING_CHUNK=96" C5 1
