# round-2 experiment D: full parity suite on the default build (halo-free 5/3 inverse at 4 CTAs/SM, tie-robust float32 ICT), barrier-free group claiming
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
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
for rep in 1 2; do
cfg default $B/libj2kb200.so "C"
cfg claimq $B/libj2kb200_claimq.so "C"
done
J2K_B200_LIB=$B/libj2kb200_claimq.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipeline or tiles or interop or c5_full or full_size or random" 2>&1 | tail -2
bash tools/ncu_cfg.sh "C3(i)" r02b
