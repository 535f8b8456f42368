# round-2 experiment I: UA variant with the alignment variants hoisted out of the job loop; RGB kernels with the two-copy body
B=go-dicom-codec_b200/csrc/build
cfg() { label=$1; lib=$2; only=$3; shift 3
  env J2K_B200_LIB=$lib "$@" timeout 300 python tools/config_bench.py --steps 20 --only "$only" 2> gpurun_out/cfg.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('%-12s'%'$label', d['config'][:34], 'fwd %.3f inv %.3f  ms %.4f %.4f'%(d['fwd_frac_hbm'], d['inv_frac_hbm'], d['fwd_ms'], d['inv_ms']), d.get('lossless_roundtrip_identical'), d.get('launches_per_forward_call'))
"; tail -2 gpurun_out/cfg.err; }
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for rep in 1 2; do
cfg default $B/libj2kb200.so "DX"
cfg default $B/libj2kb200.so "CR"
cfg default $B/libj2kb200.so "C3(i)"
cfg default $B/libj2kb200.so "C5"
cfg rgb2copy $B/libj2kb200_rgb2copy.so "C3(i)"
cfg rgb2copy $B/libj2kb200_rgb2copy.so "C5"
done
bash tools/ncu_cfg.sh "DX" r02_dx2
