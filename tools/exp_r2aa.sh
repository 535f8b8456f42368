# j2k_forward_ht: sub-batch size against frames per call (blocking call, pinned buffers), then the bench line with the new
# step-sized HT encode leg
for f in 8 32; do for m in 16 32 64; do
  echo "frames $f subbatch_msamples $m: $(J2K_HT_SUBBATCH_MSAMPLES=$m timeout 300 python tools/ht_enc_probe.py $f 4 2>/dev/null | tail -2 | tr '\n' ' ')"
done; done
timeout 900 python bench.py > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02c.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), "ht_enc e2e", round(d["ht_encode"]["e2e"]["value"]), "step batch", d["ht_encode"]["e2e_step_batch"])
PY
