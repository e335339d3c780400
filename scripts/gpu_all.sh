#!/bin/bash
# full GPU visit: parity suite, every bench config (ours + reference arm), sanitizer
TAG=${1:-all}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
for c in 1 2 3 4 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 > $OUT/bench_cfg$c.json 2> $OUT/bench_cfg$c.err; echo "cfg$c exit $?"
  timeout 900 python bench.py --config $c --impl reference --steps 5 --warmup 1 --no-cpu-baseline > $OUT/ref_cfg$c.json 2> $OUT/ref_cfg$c.err; echo "ref cfg$c exit $?"
done
python - <<PY
import json
for c in range(1,6):
    for kind in ("bench","ref"):
        try:
            d=json.loads(open("$OUT/%s_cfg%d.json"%(kind,c)).read().strip().splitlines()[-1])
            r=d.get("roofline",{})
            print(kind, c, "GCUPS", round(d["value"],3), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],2), "roofline", r.get("kernel","")[:28], round(r.get("frac",0),3), "lit", (d.get("literal") or {}).get("ms_per_call"), d.get("reference_on_gpu",{}).get("value"), d.get("check"))
            if kind=="bench": print("     phases", d["phases_ms"])
        except Exception as e:
            print(kind, c, "failed", e)
PY
if [ "$2" = "sanitize" ]; then bash scripts/gpu_sanitize.sh; fi
