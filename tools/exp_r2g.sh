# round-2 experiment G: general-alignment ring variant on the GPU (parity + DX / CR config lines, ring on and off)
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
cfg() { label=$1; only=$2; shift 2
  env "$@" timeout 300 python tools/config_bench.py --steps 20 --only "$only" 2> gpurun_out/cfg.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('%-12s'%'$label', d['config'][:34], 'fwd %.3f inv %.3f  ms %.4f %.4f'%(d['fwd_frac_hbm'], d['inv_frac_hbm'], d['fwd_ms'], d['inv_ms']), d.get('lossless_roundtrip_identical'), d.get('launches_per_forward_call'))
"; tail -2 gpurun_out/cfg.err; }
for rep in 1 2; do
cfg ua "DX" A=1
cfg ua "CR" A=1
cfg no_ua "DX" J2K_RING_UA=0
cfg no_ua "CR" J2K_RING_UA=0
done
cfg all "C" A=1
timeout 300 compute-sanitizer --tool memcheck python tools/config_bench.py --steps 2 --only "CR" --frames 2 2>&1 | tail -8
