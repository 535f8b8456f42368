# round-2 experiment F: is the single-stream forward launch slower than round 1's?  A/B of the two trees on one box, interleaved
one() { # label dir
  (cd $2 && timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-e2e $3 > /tmp/ab.json 2> /tmp/ab.err; python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open("/tmp/ab.json"))
    print("%-6s step %.4f ms  serial kernel %.4f ms (frac %.3f)  single launches %s  inv %.4f  sustained %.4f" % (sys.argv[1], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["kernel_ms_single_launches"][:4], d["inverse"]["ms_per_step"], d["sustained"]["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "FAILED", open("/tmp/ab.err").read()[-400:])
PY
)
}
for rep in 1 2 3; do
one r1 .r1ab ""
one r2 . "--no-configs"
done
