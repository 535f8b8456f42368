# round-2 experiment H: GPU parity of the general-alignment variant after the right-border fix; ncu of the DX config
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
bash tools/ncu_cfg.sh "DX" r02_dx
