# HT block decoder: first GPU run (parity suite, then the bench tool)
timeout 900 python -m pytest tests/test_ht_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 600 python tools/ht_bench.py --size 1024 --frames 8 > gpurun_out/ht_bench_1k.json 2> gpurun_out/ht_bench_1k.err; echo rc=$?; tail -3 gpurun_out/ht_bench_1k.err; cat gpurun_out/ht_bench_1k.json
timeout 900 python tools/ht_bench.py --size 2048 --frames 8 > gpurun_out/ht_bench_2k.json 2> gpurun_out/ht_bench_2k.err; echo rc=$?; tail -3 gpurun_out/ht_bench_2k.err; cat gpurun_out/ht_bench_2k.json
