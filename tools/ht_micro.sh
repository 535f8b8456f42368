# HT decode kernels: timing harness (tools/ht_micro.cu, built in the container) + optional ncu capture summarised on the box
B=tools/_ht
$B/ht_micro $B/ht_2048_12.bin 8 20 4 0
$B/ht_micro $B/ht_2048_12.bin 8 20 4 1
$B/ht_micro $B/ht_2048_12.bin 8 20 4 2
$B/ht_micro $B/ht_2048_12.bin 64 10 4 0
$B/ht_micro $B/ht_2048_12.bin 64 10 4 1
$B/ht_micro $B/ht_2048_12.bin 64 10 4 2
if [ "$1" = "ncu" ]; then
for k in ht_vlc ht_magsgn; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 2 --launch-count 1 -f -o gpurun_out/prof_$k $B/ht_micro $B/ht_2048_12.bin 64 2 4 > gpurun_out/ncu_$k.log 2>&1; echo ncu rc=$?
python tools/ncu_summary.py gpurun_out/prof_$k.ncu-rep gpurun_out/ncu_$k.json "ht_micro" > /dev/null
python - $k <<'PY'
import json,sys
k=json.load(open("gpurun_out/ncu_%s.json"%sys.argv[1]))["kernels"][0]
for key,v in k.items():
    if key!="Kernel Name" and ("stalled" not in key or float(v["value"])>0.3): print(key, v["value"], v["unit"])
PY
python tools/ncu_hot.py gpurun_out/prof_$k.ncu-rep "" 16 2>&1 | head -40
rm -f gpurun_out/prof_$k.ncu-rep
done
fi
