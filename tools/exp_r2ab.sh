# j2k_forward_ht with ramped sub-batches (J2K_HT_RAMP=1, default) against uniform ones: HT GPU tests, then the blocking call
# on 8 and 32 C2 frames, interleaved twice
timeout 900 python -m pytest tests/test_ht_gpu.py -x -q 2>&1 | tail -2
for rep in 1 2; do for f in 8 32; do for r in 0 1; do
  echo "frames $f ramp $r: $(J2K_HT_RAMP=$r timeout 300 python tools/ht_enc_probe.py $f 5 2>/dev/null | tail -2 | tr '\n' ' ')"
done; done; done
for m in 96 128; do echo "frames 32 ramp 1 sub $m: $(J2K_HT_SUBBATCH_MSAMPLES=$m timeout 300 python tools/ht_enc_probe.py 32 5 2>/dev/null | tail -2 | tr '\n' ' ')"; done
