# round-2 experiment O: forward ring kernels at 13-16 warps per SM (register cap x warps per CTA), forward columns only
# (the inverse kernels keep their own launch bounds; their columns are meaningless for the 5- and 7-warp builds)
libs="go-dicom-codec_b200/csrc/build/libj2kb200.so $(ls go-dicom-codec_b200/csrc/build/libj2kb200_*.so 2>/dev/null)"
for rep in 1 2; do
for lib in $libs; do
  echo "== $(basename $lib)"
  for only in C2 "C3(i)" "C5" DX; do
  J2K_B200_LIB=$lib timeout 200 python tools/config_bench.py --steps 20 --only "$only" 2> gpurun_out/cfg.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('   ', d['config'][:28], 'fwd %.3f inv %.3f  ms %.4f %.4f'%(d['fwd_frac_hbm'], d['inv_frac_hbm'], d['fwd_ms'], d['inv_ms']))
"
  done
done
done
