# fwd3w_kernel (one-producer ICT + 9/7 forward): parity, then C3(i) / C5 forward against the component-split jobs (J2K_FWD3W=0)
# and the job-height knob, interleaved twice on one box.
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "one_producer or pipeline or tall_chunks or tiles" 2>&1 | tail -3
for rep in 1; do
for v in "J2K_FWD3W=0" "J2K_FWD3W=1 J2K_FWD3W_TDIV=4" "J2K_FWD3W=1 J2K_FWD3W_TDIV=2" "J2K_FWD3W=1 J2K_FWD3W_TDIV=1" "J2K_FWD3W=1 J2K_FWD3W_TDIV=8"; do
  for c in C3i C5; do
    env $v timeout 300 python tools/config_bench.py --steps 20 --only $c 2>gpurun_out/r2t.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$v', d['key'], 'fwd', round(d['fwd_frac_hbm'],4), 'inv', round(d['inv_frac_hbm'],4), 'fwd_ms', round(d['fwd_ms'],4), 'inv_ms', round(d['inv_ms'],4))
"
  done
done
done
