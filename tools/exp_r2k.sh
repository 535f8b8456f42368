# ncu of the three-producer inverse kernel (C3i, C5)
for only in "C3(i)" "C5"; do
tag=$(echo $only | tr -d '()')
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inv3w_kernel --launch-skip 3 --launch-count 1 \
  -f -o gpurun_out/prof_inv_r02_x3_$tag python tools/config_bench.py --steps 2 --only "$only" > gpurun_out/ncu_inv_r02_x3_$tag.log 2>&1; echo $only rc=$?
done
