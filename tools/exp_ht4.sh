timeout 900 python bench.py --steps 20 --no-configs --sustained-seconds 0 --no-cpu-baseline > gpurun_out/bench_ht.json 2> gpurun_out/bench_ht.err; echo bench rc=$?; tail -5 gpurun_out/bench_ht.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_ht.json"))
print(json.dumps(d.get("ht_encode"), indent=1))
h=d.get("ht_decode"); print("decode:", h["ht_decode_Mpixel_s"], h["e2e"]["value"], h["e2e_planes"]["value"], h["pixels_identical"])
print("e2e", d["e2e"]["value"], "value", d["value"])
PY
