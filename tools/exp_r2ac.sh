# deferred release of the forward kernels (default build) against the immediate one (libj2kb200_nodefer.so): forward parity
# subset, then bench resident legs + config table, interleaved twice
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "pipeline or tiles or one_producer or tall_chunks or c3_full or c5 or batch or series or hybrid or general_alignment or pipelined" 2>&1 | tail -2
N=go-dicom-codec_b200/csrc/build/libj2kb200_nodefer.so
for rep in 1 2; do
for v in "J2K_X=defer" "J2K_B200_LIB=$N"; do
  env $v timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-ht --sustained-seconds 0 > gpurun_out/ac.json 2> gpurun_out/ac.err
  python - "$v" <<'PY'
import json,sys
d=json.load(open("gpurun_out/ac.json"))
print(sys.argv[1][:22], "C2 fwd step %.4f ms (frac %.4f) alone %.4f ms (frac %.4f)" % (d["ms_per_step"], d["roofline"]["step_frac"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
PY
  env $v timeout 300 python tools/config_bench.py --steps 20 2>gpurun_out/r2ac.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('   ', d['key'], d['frames'], 'fwd', round(d['fwd_frac_hbm'],4), 'fwd_ms', round(d['fwd_ms'],4))
"
done
done
