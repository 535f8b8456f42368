# round-2 experiment N: validation of the build that is going to ship (parity suite, smoke, config lines, full bench both arms)
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python tools/config_bench.py --steps 20 2> gpurun_out/cfg.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config'][:40], 'fwd %.3f inv %.3f  ms %.4f %.4f'%(d['fwd_frac_hbm'], d['inv_frac_hbm'], d['fwd_ms'], d['inv_ms']))
"
timeout 900 python bench.py > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err; echo bench rc=$?; tail -2 gpurun_out/bench_r02_b.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r02_b_ref.json 2> gpurun_out/bench_r02_b_ref.err; echo ref rc=$?
