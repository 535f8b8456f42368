# warp-to-warp ring waits that sleep between polls (default build) against the tight poll loop (libj2kb200_nosleep.so):
# C3(i) x32 and C5 in both directions, interleaved twice; parity subset first
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "one_producer or pipeline or tiles or c3_full or c5" 2>&1 | tail -2
N=go-dicom-codec_b200/csrc/build/libj2kb200_nosleep.so
for rep in 1 2; do
for v in "J2K_X=sleep" "J2K_B200_LIB=$N"; do
  for c in C3i C5; do
    env $v timeout 300 python tools/config_bench.py --steps 20 --only $c 2>gpurun_out/r2y.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d['key'] != '$c': continue
    print('$v'[:24], d['key'], d['frames'], 'fwd', round(d['fwd_frac_hbm'],4), 'inv', round(d['inv_frac_hbm'],4), 'fwd_ms', round(d['fwd_ms'],4), 'inv_ms', round(d['inv_ms'],4))
"
  done
done
done
