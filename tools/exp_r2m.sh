# round-2 experiment M: three-producer inverse with round-robin triples (no CTA barrier), reverted HF / inverse-UA fetch
cfg() { label=$1; only=$2; shift 2
  env "$@" timeout 300 python tools/config_bench.py --steps 20 --only "$only" 2> gpurun_out/cfg.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('%-12s'%'$label', d['config'][:28], 'fwd %.3f inv %.3f  ms %.4f %.4f'%(d['fwd_frac_hbm'], d['inv_frac_hbm'], d['fwd_ms'], d['inv_ms']))
"; tail -2 gpurun_out/cfg.err; }
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for rep in 1 2; do
cfg all "" A=1
cfg old "C3(i)" J2K_INV3W=0
cfg old "C5" J2K_INV3W=0
done
