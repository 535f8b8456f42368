# round-2 experiment B: float32 inverse ICT fast path, per-warp claiming in the forward kernel, 4 CTAs/SM for the 5/3 inverse
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
for rep in 1 2; do
cfg default $B/libj2kb200.so "C"
cfg noict $B/libj2kb200_noict.so "C3(i)"
cfg noict $B/libj2kb200_noict.so "C5"
cfg warpclaim $B/libj2kb200_warpclaim.so "C"
cfg inv53b4 $B/libj2kb200_inv53b4.so "C1"
cfg inv53b4 $B/libj2kb200_inv53b4.so "C4"
cfg inv53b4 $B/libj2kb200_inv53b4.so "C2"
done
cfg default_t9472 $B/libj2kb200.so "C3" J2K_RING_TARGET_JOBS=9472
cfg default_32fr $B/libj2kb200.so "C3(i)" 
timeout 300 python tools/config_bench.py --steps 20 --only "C3(i)" --frames 32 2>&1 | cut -c1-300
# parity of the default build (fast ICT) on the RGB cases + random sweep
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipeline or tiles or interop or c5_full or full_size or random" 2>&1 | tail -2
# ncu of the C3i inverse with the fast path
bash tools/ncu_cfg.sh "C3(i)" r02a
