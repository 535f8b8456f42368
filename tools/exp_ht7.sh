timeout 300 ncu --set full --clock-control none --import-source on -k regex:ht_enc_pack --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_ht_pack python tools/ht_enc_probe.py 8 1 > gpurun_out/ncu_ht_pack.log 2>&1; echo rc=$?
python tools/ncu_summary.py gpurun_out/prof_ht_pack.ncu-rep gpurun_out/ncu_ht_pack.json "pack" > /dev/null
python - <<'PY'
import json
k=json.load(open("gpurun_out/ncu_ht_pack.json"))["kernels"][0]
for key,v in k.items():
    if key!="Kernel Name": print(key, v["value"], v["unit"])
PY
python tools/ncu_hot.py gpurun_out/prof_ht_pack.ncu-rep "" 24 2>&1 | head -50
rm -f gpurun_out/prof_ht_pack.ncu-rep
