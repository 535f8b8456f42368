# full GPU parity suite with fwd3w_kernel as the default ICT + 9/7 forward, then the config table (C3(i) at 8 / 32 frames)
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/config_bench.py --steps 20 2>gpurun_out/r2v.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['key'], d['frames'], 'fwd', round(d['fwd_frac_hbm'],4), 'inv', round(d['inv_frac_hbm'],4), 'fwd_ms', round(d['fwd_ms'],4), 'inv_ms', round(d['inv_ms'],4))
"
for f in 16 32; do python tools/config_bench.py --steps 20 --only C3i --frames $f 2>>gpurun_out/r2v.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d['key']=='C3i': print(d['key'], d['frames'], 'fwd', round(d['fwd_frac_hbm'],4), 'inv', round(d['inv_frac_hbm'],4), 'fwd_ms', round(d['fwd_ms'],4), 'inv_ms', round(d['inv_ms'],4))
"; done
