# A/B kernel experiments on ONE box: every go-dicom-codec_b200/csrc/build/libj2kb200_*.so variant plus the default build,
# resident bench only, interleaved twice so that box-to-box variation cancels.
libs="go-dicom-codec_b200/csrc/build/libj2kb200.so $(ls go-dicom-codec_b200/csrc/build/libj2kb200_*.so 2>/dev/null)"
for rep in 1 2; do
for lib in $libs; do
  J2K_B200_LIB=$lib timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - "$lib" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/ab.json"))
    print("%-70s fwd step %.4f ms (frac %.3f) alone %.4f ms (frac %.3f) | inv %.4f ms (frac %.3f)" % (sys.argv[1].split("/")[-1], d["ms_per_step"], d["roofline"]["step_frac"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["inverse"]["ms_per_step"], d["inverse"]["step_frac_of_hbm_peak"]))
except Exception as e:
    print(sys.argv[1], "FAILED", open("gpurun_out/ab.err").read()[-300:])
PY
done
done
