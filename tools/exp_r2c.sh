# round-2 experiment C: halo-free 5/3 inverse strips (default build), 4 CTAs/SM for every ring kernel (minb4), 4 CTAs/SM for the 5/3 inverse only
B=go-dicom-codec_b200/csrc/build
cfg() { # label lib only env...
  label=$1; lib=$2; only=$3; shift 3
  env J2K_B200_LIB=$lib "$@" timeout 300 python tools/config_bench.py --steps 20 --only "$only" 2> gpurun_out/cfg.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('%-22s'%'$label', d['config'][:28], 'fwd %.3f inv %.3f  ms %.4f %.4f'%(d['fwd_frac_hbm'], d['inv_frac_hbm'], d['fwd_ms'], d['inv_ms']))
"
}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
cfg default $B/libj2kb200.so "C"
cfg minb4 $B/libj2kb200_minb4.so "C"
cfg inv53b4 $B/libj2kb200_inv53b4.so "C1"
cfg inv53b4 $B/libj2kb200_inv53b4.so "C4"
cfg inv53b4 $B/libj2kb200_inv53b4.so "C3(ii)"
done
J2K_B200_LIB=$B/libj2kb200_minb4.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipeline or tiles or interop or c5_full or full_size or random" 2>&1 | tail -2
