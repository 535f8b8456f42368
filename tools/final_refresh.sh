bash tools/ncu_capture.sh r01g
timeout 600 python tools/config_bench.py --steps 10 > gpurun_out/config_bench.jsonl 2> gpurun_out/cfg.err; echo cfg rc=$?
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo bench rc=$?
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?
