# launch-size threshold of fwd3w_kernel after the job-size policies: C3(i) at 8 / 12 / 16 / 24 frames, forced (2) against the component jobs (0)
for f in 8 12 16 24; do for m in 0 2; do
  J2K_FWD3W=$m timeout 300 python tools/config_bench.py --steps 20 --only C3i --frames $f 2>gpurun_out/r2ak.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d['key']=='C3i': print('frames $f J2K_FWD3W=$m fwd', round(d['fwd_frac_hbm'],4), 'fwd_ms', round(d['fwd_ms'],4))
"
done; done
