# HT block encoder + decoder on the GPU: parity suite
timeout 1200 python -m pytest tests/test_ht_gpu.py -m gpu -x -q 2>&1 | tail -12
