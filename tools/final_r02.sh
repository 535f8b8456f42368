# Round-2 evidence run (one gpurun call): parity suite, smoke, the default bench line (both arms), the per-config lines, the
# launch list of the bench command, and one ncu --set full capture of the dominant kernel of every config.  gpurun brings back
# at most 64 MiB, so every .ncu-rep is summarised ON THE BOX (ncu -i needs no GPU) into gpurun_out/prof_$tag/ and deleted;
# tools/summarise_r02.sh copies that directory into profiles/.  Numbers printed under ncu are never bench values.
tag=${1:-r02}; keep=${2:-none}
out=gpurun_out/prof_$tag; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $out/gputest_$tag.log; cat $out/gputest_$tag.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee $out/smoke_$tag.log
timeout 900 python bench.py > $out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo bench rc=$?
timeout 900 python bench.py --impl reference > $out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err; echo ref rc=$?
timeout 600 python tools/config_bench.py --steps 20 > $out/config_bench_$tag.jsonl 2> gpurun_out/cfg.err; echo cfg rc=$?
summ() { # name note
  f=gpurun_out/prof_${1}_$tag.ncu-rep
  [ -f $f ] || { echo missing $f; return; }
  python tools/ncu_summary.py $f $out/ncu_${tag}_$1.json "$2" > /dev/null
  { echo "# executed warp instructions per code region (tools/ncu_groups.py) and per opcode (tools/ncu_mix.py): $1"; python tools/ncu_groups.py $f 10; python tools/ncu_mix.py $f 24; } > $out/ncu_${tag}_${1}_instructions.txt 2>&1
  case " $keep " in *" $1 "*) ;; *) rm -f $f;; esac
}
common="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --streams 1 --no-configs --no-ht"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_ring_kernel --launch-skip 4 --launch-count 1 \
  -f -o gpurun_out/prof_fwd_ring_$tag $common --no-inverse > gpurun_out/ncu_fwd_$tag.log 2>&1; echo fwd rc=$?
summ fwd_ring "tools/final_r02.sh $tag: $common --no-inverse, fwd_ring_kernel, 32 C2 frames, launch 5"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inv_ring_kernel --launch-skip 4 --launch-count 1 \
  -f -o gpurun_out/prof_inv_ring_$tag $common > gpurun_out/ncu_inv_$tag.log 2>&1; echo inv rc=$?
summ inv_ring "tools/final_r02.sh $tag: $common, inv_ring_kernel, 32 C2 frames, launch 5"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv $common > gpurun_out/ncu_list_$tag.log 2>&1; echo list rc=$?
cap() { # config-substring kernel-regex name
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip 3 --launch-count 1 \
    -f -o gpurun_out/prof_${3}_$tag python tools/config_bench.py --steps 2 --only "$1" > gpurun_out/ncu_${3}_$tag.log 2>&1; echo $3 rc=$?
  summ $3 "tools/final_r02.sh $tag: ncu --set full --clock-control none --import-source on -k regex:$2, tools/config_bench.py --only '$1', launch 4"
}
cap "C3(i)" fwd3w_kernel fwd_rgb97
cap "C3(i)" inv3w_kernel inv_rgb97
cap "C5" fwd3w_kernel fwd_c5
cap "C5" inv3w_kernel inv_c5
cap "C1" fwd_ring_kernel fwd_c1
cap "C1" inv_ring_kernel inv_c1
cap "C3(ii)" fwd_ring_kernel fwd_rgb53
cap "C3(ii)" inv_ring_kernel inv_rgb53
cap "C4" fwd_ring_kernel fwd_c4
cap "C4" inv_ring_kernel inv_c4
cap "DX" fwd_ring_kernel fwd_dx
cap "DX" inv_ring_kernel inv_dx
du -sh gpurun_out
