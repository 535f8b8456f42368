# round-2 experiment E: full parity suite, config bench with the T1 round trip in front of the inverse legs (ICT fast path on / off), full bench line
B=go-dicom-codec_b200/csrc/build
cfg() { # label lib only env...
  label=$1; lib=$2; only=$3; shift 3
  env J2K_B200_LIB=$lib "$@" timeout 300 python tools/config_bench.py --steps 20 --only "$only" 2> gpurun_out/cfg.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('%-22s'%'$label', d['config'][:28], 'fwd %.3f inv %.3f  ms %.4f %.4f'%(d['fwd_frac_hbm'], d['inv_frac_hbm'], d['fwd_ms'], d['inv_ms']), d.get('lossy_roundtrip_max_abs_error'))
"
  tail -3 gpurun_out/cfg.err
}
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
for rep in 1 2; do
cfg default $B/libj2kb200.so "C"
cfg noict $B/libj2kb200_noict.so "C3(i)"
cfg noict $B/libj2kb200_noict.so "C5"
done
timeout 900 python bench.py > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; echo bench rc=$?; tail -3 gpurun_out/bench_r02_a.err
timeout 600 python bench.py --impl reference --steps 6 --warmup 3 > gpurun_out/bench_r02_a_ref.json 2> gpurun_out/bench_r02_a_ref.err; echo ref rc=$?
bash tools/ncu_cfg.sh "C3(i)" r02c
