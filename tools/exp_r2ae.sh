# second sweep: chunk cap for the RCT + 5/3 RGB kernels (C3(ii) at 8 and 32 frames), job-count target of the coarser levels
# behind fwd3w_kernel (C3(i) x32, C5)
run() {  # env, config, frames
  env $1 timeout 300 python tools/config_bench.py --steps 20 --only $2 ${3:+--frames $3} 2>gpurun_out/r2ae.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d['key'] != '$2': continue
    print('$1', d['key'], d['frames'], 'fwd', round(d['fwd_frac_hbm'],4), 'inv', round(d['inv_frac_hbm'],4))
"
}
for c in 16 24 32 48; do run "J2K_RING_CHUNK=$c" C3ii 32; run "J2K_RING_CHUNK=$c" C3ii 8; done
for t in 4736 7104 9472 14208; do run "J2K_RING_TARGET_JOBS=$t" C3i 32; run "J2K_RING_TARGET_JOBS=$t" C5; done
run "J2K_RING_TARGET_JOBS=9472 J2K_RING_CHUNK_DEEP=64" C3i 32
run "J2K_RING_TARGET_JOBS=9472 J2K_RING_CHUNK_DEEP=64" C5
