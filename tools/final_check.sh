timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo bench rc=$?
timeout 300 python tools/config_bench.py --steps 10 > gpurun_out/config_bench.jsonl 2> gpurun_out/cfg.err; echo cfg rc=$?
