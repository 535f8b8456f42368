# round-2 experiment A: 3-component 9/7 inverse with NP = 4 (2 CTAs/SM), single- and two-copy body, against the default build
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
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
for rep in 1 2; do
cfg base $B/libj2kb200.so "C"
cfg inv4 $B/libj2kb200_inv4.so "C3i"
cfg inv4 $B/libj2kb200_inv4.so "C5"
cfg inv4b $B/libj2kb200_inv4b.so "C3i"
cfg inv4b $B/libj2kb200_inv4b.so "C5"
done
cfg base_t2368 $B/libj2kb200.so "C3i" J2K_RING_TARGET_JOBS=2368
cfg base_t1184 $B/libj2kb200.so "C3i" J2K_RING_TARGET_JOBS=1184
cfg inv4_t2368 $B/libj2kb200_inv4.so "C3i" J2K_RING_TARGET_JOBS=2368
cfg inv4_t1184 $B/libj2kb200_inv4.so "C3i" J2K_RING_TARGET_JOBS=1184
# parity of the variants on the RGB cases
for v in inv4 inv4b; do
J2K_B200_LIB=$B/libj2kb200_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipeline or tiles or interop or c5_full or full_size" 2>&1 | tail -2
done
