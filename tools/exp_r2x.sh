# inv3w_kernel: job triples through a shared-memory queue (J2K_INV3W_QUEUE=1) against the CTA barrier per claim; failed-device test
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "pipeline or tiles or c3_full or c5 or tall_chunks or interop or failed_device" 2>&1 | tail -3
run() {  # env, config, frames
  env $1 timeout 300 python tools/config_bench.py --steps 20 --only $2 --frames $3 2>gpurun_out/r2x.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d['key'] != '$2': continue
    print('$1', d['key'], 'frames', d['frames'], 'inv', round(d['inv_frac_hbm'],4), 'inv_ms', round(d['inv_ms'],4))
"
}
for v in "J2K_INV3W_QUEUE=0" "J2K_INV3W_QUEUE=1" "J2K_INV3W_QUEUE=1 J2K_INV3W_TDIV=2" "J2K_INV3W_QUEUE=1 J2K_INV3W_TDIV=1" "J2K_INV3W_QUEUE=1 J2K_INV3W_TDIV=8"; do
  run "$v" C3i 8; run "$v" C3i 32; run "$v" C5 1
done
