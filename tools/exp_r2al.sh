# general-alignment forward: unmasked stores for interior strips (default build) against the masked stores everywhere (libj2kb200_base.so)
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "general_alignment or random_geometry or hybrid or wavelet_api" 2>&1 | tail -2
B=go-dicom-codec_b200/csrc/build/libj2kb200_base.so
for rep in 1 2; do for v in "J2K_X=new" "J2K_B200_LIB=$B"; do for c in DX CR; do
  env $v timeout 300 python tools/config_bench.py --steps 20 --only $c 2>gpurun_out/r2al.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d['key']=='$c': print('$v'[:20], d['key'], 'fwd', round(d['fwd_frac_hbm'],4), 'inv', round(d['inv_frac_hbm'],4), 'fwd_ms', round(d['fwd_ms'],4))
"
done; done; done
