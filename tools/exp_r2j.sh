# round-2 experiment J: three-producer level-1 inverse (inv3w_kernel) against the one-warp-per-strip variant
cfg() { label=$1; only=$2; shift 2
  env "$@" timeout 300 python tools/config_bench.py --steps 20 --only "$only" 2> gpurun_out/cfg.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('%-12s'%'$label', d['config'][:34], 'fwd %.3f inv %.3f  ms %.4f %.4f'%(d['fwd_frac_hbm'], d['inv_frac_hbm'], d['fwd_ms'], d['inv_ms']), d.get('lossy_roundtrip_max_abs_error'), d.get('launches_per_forward_call'))
"; tail -2 gpurun_out/cfg.err; }
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_ict_fast.py tests/test_gpu_fullsize.py -m gpu -q -x -k "pipeline or tiles or c3_full or c5_full or tie or interop or random" 2>&1 | tail -4
for rep in 1 2; do
cfg x3 "C3(i)" A=1
cfg x3 "C5" A=1
cfg old "C3(i)" J2K_INV3W=0
cfg old "C5" J2K_INV3W=0
done
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
bash tools/ncu_cfg.sh "C3(i)" r02_x3
