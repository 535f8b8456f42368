# round-2 experiment Q: group-pipelined job order (J2K_RING_LAG) per config and direction -- the inverse runs its small coarse levels
# FIRST, a dependency chain most warps wait on (8 % of the DX inverse's instructions are the polling loop)
for lag in 0 1 2 3; do
for only in DX "C2" "C1" "C3(i)" "C5" CR; do
  J2K_RING_LAG=$lag timeout 200 python tools/config_bench.py --steps 20 --only "$only" 2> gpurun_out/cfg.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('lag $lag', d['config'][:28], 'fwd %.3f inv %.3f  ms %.4f %.4f'%(d['fwd_frac_hbm'], d['inv_frac_hbm'], d['fwd_ms'], d['inv_ms']))
"; tail -2 gpurun_out/cfg.err
done
done
