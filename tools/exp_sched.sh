# job-order experiment: lag sweep of the group-pipelined schedule on the bench workload (resident legs only),
# with and without the L2 eviction policies, then DRAM traffic of the forward kernel
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err
  python - "$label" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/b.json")); print("%-22s"%sys.argv[1], "fwd step %.4f alone %.4f (frac %.3f) inv %.4f"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["inverse"]["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "FAILED", open("gpurun_out/b.err").read()[-300:])
PY
}
NOH="J2K_RING_POL_IN=0 J2K_RING_POL_LLW=0 J2K_RING_POL_BAND=0"
run lag0 J2K_RING_LAG=0
run lag3_nohint J2K_RING_LAG=3 $NOH
run lag2 J2K_RING_LAG=2
run lag3 J2K_RING_LAG=3
run lag4 J2K_RING_LAG=4
run lag3_llw_only J2K_RING_LAG=3 J2K_RING_POL_IN=0 J2K_RING_POL_BAND=0
run lag3_llr_first J2K_RING_LAG=3 J2K_RING_POL_LLR=1
run lag3_llr_last J2K_RING_LAG=3 J2K_RING_POL_LLR=2
run lag2_chunk32 J2K_RING_LAG=2 J2K_RING_CHUNK=32
run lag0_hints J2K_RING_LAG=0 J2K_RING_POL_IN=1 J2K_RING_POL_BAND=1
for lag in 3; do
J2K_RING_LAG=$lag timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:ring_kernel --launch-skip 4 --launch-count 2 --csv --log-file gpurun_out/sched_ncu_lag$lag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --streams 1 > gpurun_out/sched_ncu_lag$lag.log 2>&1
python - $lag <<'PY'
import csv,sys
rows=list(csv.DictReader(l for l in open("gpurun_out/sched_ncu_lag%s.csv"%sys.argv[1]) if l.startswith('"')))
for r in rows: print("lag", sys.argv[1], r["Kernel Name"][:30], r["Metric Name"], r["Metric Value"], r["Metric Unit"])
PY
done
