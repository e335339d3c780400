#!/bin/bash
TAG=${1:-b5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for c in ${2:-1 2 3 4 5}; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 ${3:-} > $OUT/bench_cfg$c.json 2> $OUT/bench_cfg$c.err; echo "cfg$c exit $?"
done
python - <<PY
import json
for c in range(1,6):
    try:
        d=json.loads(open("$OUT/bench_cfg%d.json"%c).read().strip().splitlines()[-1])
        r=d.get("roofline",{})
        print(c, "GCUPS", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],2), "|", r.get("kernel","")[:28], round(r.get("frac",0),3), "| lit", (d.get("literal") or {}).get("ms_per_call"), "| clocks", d["clocks"].get("sm_mhz"), d["clocks"].get("samples_in_timed_region"), d["clocks"].get("reasons"), d["config"]["l2"][:40])
        print("     phases", {k:v for k,v in d["phases_ms"].items() if v})
    except Exception as e:
        print(c, "failed", e)
PY
