# inv3q_kernel (four quads per CTA, setmaxnreg) against inv3w_kernel: parity subset, then C3(i) x8 / x32 and C5 inverse
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "pipeline or tiles or c3_full or c5 or tall_chunks or interop" 2>&1 | tail -3
run() {  # env, config, frames
  env $1 timeout 300 python tools/config_bench.py --steps 20 --only $2 --frames $3 2>gpurun_out/r2w.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d['key'] != '$2': continue
    print('$1', d['key'], 'frames', d['frames'], 'fwd', round(d['fwd_frac_hbm'],4), 'inv', round(d['inv_frac_hbm'],4), 'inv_ms', round(d['inv_ms'],4))
"
}
Q=go-dicom-codec_b200/csrc/build/libj2kb200_q152.so
for v in "J2K_INV3Q=0" "J2K_INV3Q=1" "J2K_INV3Q=1 J2K_B200_LIB=$Q" "J2K_INV3Q=1 J2K_INV3W_TDIV=2" "J2K_INV3Q=1 J2K_INV3W_TDIV=8"; do
  run "$v" C3i 8; run "$v" C3i 32; run "$v" C5 1
done
