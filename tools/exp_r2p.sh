# round-2 experiment P: general-alignment ring kernels staged by 1-D tensor-map copies (UTMALDG) instead of bulk copies + row phases
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "general_alignment or random_geometry or hybrid or wavelet_api" 2>&1 | tail -4
for only in DX CR C2; do
  timeout 200 python tools/config_bench.py --steps 20 --only "$only" 2> gpurun_out/cfg.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('   ', d['config'][:28], 'fwd %.3f inv %.3f  ms %.4f %.4f'%(d['fwd_frac_hbm'], d['inv_frac_hbm'], d['fwd_ms'], d['inv_ms']))
"; tail -2 gpurun_out/cfg.err
done
