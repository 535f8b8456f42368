# HT block decoder, two-kernel build: parity suite, bench tool, the bench line with the ht_decode leg
timeout 900 python -m pytest tests/test_ht_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 900 python tools/ht_bench.py --size 2048 --frames 8 > gpurun_out/ht_bench_2k.json 2> gpurun_out/ht_bench_2k.err; echo rc=$?; tail -3 gpurun_out/ht_bench_2k.err; cat gpurun_out/ht_bench_2k.json
timeout 900 python bench.py --steps 20 --no-configs --sustained-seconds 0 --no-cpu-baseline > gpurun_out/bench_ht.json 2> gpurun_out/bench_ht.err; echo bench rc=$?; tail -5 gpurun_out/bench_ht.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_ht.json"))
print(json.dumps(d.get("ht_decode"), indent=1))
print("e2e", d["e2e"]["value"], "value", d["value"])
PY
