# chunk policies (RGB 5/3 cap 32, 1.5 x deep jobs behind fwd3w_kernel) as defaults: full GPU suite, config table, tiled RGB 5/3 probe
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/config_bench.py --steps 20 2>gpurun_out/r2af.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['key'], d['frames'], 'fwd', round(d['fwd_frac_hbm'],4), 'inv', round(d['inv_frac_hbm'],4), 'fwd_ms', round(d['fwd_ms'],4), 'inv_ms', round(d['inv_ms'],4))
"
