# final check of the session: full GPU suite, smoke, bench both arms, config table, the two inv3w captures refreshed
tag=r02; out=gpurun_out/prof_$tag; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $out/gputest_$tag.log; cat $out/gputest_$tag.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee $out/smoke_$tag.log
timeout 900 python bench.py > $out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo bench rc=$?
timeout 900 python bench.py --impl reference > $out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err; echo ref rc=$?
timeout 600 python tools/config_bench.py --steps 20 > $out/config_bench_$tag.jsonl 2> gpurun_out/cfg.err; echo cfg rc=$?
summ() { f=gpurun_out/prof_${1}_$tag.ncu-rep; [ -f $f ] || { echo missing $f; return; }
  python tools/ncu_summary.py $f $out/ncu_${tag}_$1.json "$2" > /dev/null
  { echo "# executed warp instructions per code region (tools/ncu_groups.py) and per opcode (tools/ncu_mix.py): $1"; python tools/ncu_groups.py $f 10; python tools/ncu_mix.py $f 24; } > $out/ncu_${tag}_${1}_instructions.txt 2>&1
  rm -f $f; }
cap() { timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip 3 --launch-count 1 \
    -f -o gpurun_out/prof_${3}_$tag python tools/config_bench.py --steps 2 --only "$1" > gpurun_out/ncu_${3}_$tag.log 2>&1; echo $3 rc=$?
  summ $3 "tools/exp_r2aj.sh: ncu --set full --clock-control none --import-source on -k regex:$2, tools/config_bench.py --only '$1', launch 4"; }
cap "DX" fwd_ring_kernel fwd_dx
cap "DX" inv_ring_kernel inv_dx
