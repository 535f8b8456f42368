# bench.py at N = 2 the way the driver launches it (new HT legs under torchrun) + the reference arm
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/scale2_r02.json 2> gpurun_out/scale2_r02.err; echo rc=$?; tail -3 gpurun_out/scale2_r02.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/scale2_r02.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"n",d["n_gpus"])
print("ht_decode e2e",d["ht_decode"]["e2e"]["value"],d["ht_decode"]["e2e_planes"]["value"],d["ht_decode"]["ht_decode_Mpixel_s"])
print("ht_encode e2e",d["ht_encode"]["e2e"]["value"],d["ht_encode"]["e2e_planes"]["value"],d["ht_encode"]["ht_encode_Mpixel_s"], d["ht_encode"]["decodes_back_to_the_coefficients"])
PY
