# three-producer inverse with taller chunks / fewer jobs (the CTA barrier per job is its main stall)
cfg() { label=$1; only=$2; shift 2
  env "$@" timeout 300 python tools/config_bench.py --steps 20 --only "$only" 2> gpurun_out/cfg.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('%-22s'%'$label', d['config'][:24], 'fwd %.3f inv %.3f  ms %.4f %.4f'%(d['fwd_frac_hbm'], d['inv_frac_hbm'], d['fwd_ms'], d['inv_ms']))
"; tail -2 gpurun_out/cfg.err; }
for only in "C3(i)" "C5"; do
cfg x3 "$only" J2K_INV3W=1
cfg x3_t1184 "$only" J2K_INV3W=1 J2K_RING_TARGET_JOBS=1184
cfg x3_t1184_c256 "$only" J2K_INV3W=1 J2K_RING_TARGET_JOBS=1184 J2K_RING_CHUNK=256
cfg x3_t592_c256 "$only" J2K_INV3W=1 J2K_RING_TARGET_JOBS=592 J2K_RING_CHUNK=256
cfg old "$only" J2K_INV3W=0
done
cfg dx "DX" A=1
cfg cr "CR" A=1
cfg c1 "C1" A=1
cfg c4 "C4" A=1
