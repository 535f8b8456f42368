# level-1 chunk height by estimated makespan (J2K_RING_MAKESPAN=1, bubble 8 / 16 / 0) against the job-target rule: every config + C2 x32
tbl() {
  env $1 python tools/config_bench.py --steps 20 2>gpurun_out/r2ah.err | python -c "
import sys,json
out=[]
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    out.append('%s %.3f/%.3f' % (d['key'], d['fwd_frac_hbm'], d['inv_frac_hbm']))
print('$1 |', ' | '.join(out))
"
  env $1 python tools/config_bench.py --steps 20 --only C2 --frames 32 2>>gpurun_out/r2ah.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d['key']=='C2': print('   C2x32 %.4f/%.4f' % (d['fwd_frac_hbm'], d['inv_frac_hbm']))
"
}
tbl "J2K_RING_MAKESPAN=0"
tbl "J2K_RING_MAKESPAN=1"
tbl "J2K_RING_MAKESPAN=1 J2K_RING_BUBBLE=16"
tbl "J2K_RING_MAKESPAN=1 J2K_RING_BUBBLE=0"
tbl "J2K_RING_MAKESPAN=0"
