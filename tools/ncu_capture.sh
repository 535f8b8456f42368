# ncu evidence for the round: one --set full capture of the forward and the inverse ring kernel (after warm-up),
# plus the launch list of the same bench command.  Numbers printed under ncu are never bench values.
tag=${1:-r01}
common="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --streams 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_ring_kernel --launch-skip 4 --launch-count 1 \
  -f -o gpurun_out/prof_fwd_ring_$tag $common --no-inverse > gpurun_out/ncu_fwd_$tag.log 2>&1; echo fwd rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inv_ring_kernel --launch-skip 4 --launch-count 1 \
  -f -o gpurun_out/prof_inv_ring_$tag $common > gpurun_out/ncu_inv_$tag.log 2>&1; echo inv rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $common > gpurun_out/ncu_list_$tag.log 2>&1; echo list rc=$?
