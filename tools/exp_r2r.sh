# round-2 experiment R: shorter chunks at the small levels (J2K_RING_CHUNK_MIN): the inverse starts with a dependency chain through its
# coarse levels that nothing can overlap (the previous launch's CTAs are all busy with level 1 until they exit)
for cm in 8 4 2; do
for only in DX "C2" "C1" "C3(i)" "C5" CR; do
  J2K_RING_CHUNK_MIN=$cm timeout 200 python tools/config_bench.py --steps 20 --only "$only" 2> gpurun_out/cfg.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('chunk_min $cm', d['config'][:28], 'fwd %.3f inv %.3f  ms %.4f %.4f'%(d['fwd_frac_hbm'], d['inv_frac_hbm'], d['fwd_ms'], d['inv_ms']))
"; tail -2 gpurun_out/cfg.err
done
done
