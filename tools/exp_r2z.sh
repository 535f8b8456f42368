# two real GPUs: the in-process sharding tests after the run_frames refactor (failed-device handling), the HT multi-device
# probe, and bench.py the way the driver launches it at N = 2
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q -k "multi_device or failed_device or sharded or tickets or pageable" 2>&1 | tail -3
timeout 600 python tools/ht_multi_probe.py > gpurun_out/ht_multi_r02b.json 2> gpurun_out/ht_multi.err; echo probe rc=$?; tail -c 600 gpurun_out/ht_multi_r02b.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "configs", [(c["key"], round(c["fwd_frac"],3), round(c["inv_frac"],3)) for c in d.get("configs", [])])
PY
