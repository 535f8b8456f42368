timeout 1200 python -m pytest tests/test_ht_gpu.py -m gpu -x -q 2>&1 | tail -4
J2K_HT_SUBBATCH_MSAMPLES=64 python tools/ht_enc_probe.py 32 3
J2K_HT_SUBBATCH_MSAMPLES=128 python tools/ht_enc_probe.py 32 3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ht_enc_launches.csv python tools/ht_enc_probe.py 8 1 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/ht_enc_launches.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ki].split("(")[0][:50]; v=float(r[vi].replace(",","")); 
    if r[ui]=="ns": v/=1e3
    if r[ui]=="ms": v*=1e3
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in agg.items(): print("%-52s n=%3d total %.1f us  avg %.1f us"%(k,n,t,t/n))
PY
