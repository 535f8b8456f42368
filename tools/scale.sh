# scaling run on one box: bench.py at N = 1, 2, 4, 8 the way the driver launches it (torchrun for N > 1)
for n in ${NS:-1 2 4 8}; do
  if [ $n = 1 ]; then timeout 600 python bench.py --gpus 1 --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; fi
  echo "N=$n rc=$?"
  python - $n <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open("gpurun_out/scale_%s.json"%n) if l.startswith("{")][-1])
    print("  value %.0f Mpixel/s  ms/step %.4f  e2e %.0f (sync %.0f)  inverse %.0f  clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["sync_value"], d["inverse"]["value"], d["clocks"]))
except Exception as e:
    print("  FAILED", e, open("gpurun_out/scale_%s.err"%n).read()[-400:])
PY
done
