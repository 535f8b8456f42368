run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err
  python - "$label" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/b.json")); print(sys.argv[1], "ms/step %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["step_frac"], {k:round(v,4) for k,v in d["roofline"]["per_level_ms"].items()})
except Exception as e:
    print(sys.argv[1], "FAILED", open("gpurun_out/b.err").read()[-300:])
PY
}
B=go-dicom-codec_b200/csrc/build
run mb3_17k J2K_B200_LIB=$B/libj2kb200_mb3_17k.so
run mb3_17k_pl J2K_B200_LIB=$B/libj2kb200_mb3_17k.so J2K_RING_PER_LEVEL=1
run mb2_17k J2K_B200_LIB=$B/libj2kb200_mb2_17k.so
run mb2_17k_pl J2K_B200_LIB=$B/libj2kb200_mb2_17k.so J2K_RING_PER_LEVEL=1
run mb3_17k_c128 J2K_B200_LIB=$B/libj2kb200_mb3_17k.so J2K_RING_CHUNK=128
run mb3_c128 J2K_RING_CHUNK=128
