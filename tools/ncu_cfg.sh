# ncu --set full of the forward and inverse ring kernels of one config_bench config: tools/ncu_cfg.sh "C3i" tag
only=$1; tag=$2
for dir in fwd inv; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${dir}_ring_kernel --launch-skip 3 --launch-count 1 \
  -f -o gpurun_out/prof_${dir}_$tag python tools/config_bench.py --steps 2 --only "$only" > gpurun_out/ncu_${dir}_$tag.log 2>&1; echo $dir rc=$?
done
