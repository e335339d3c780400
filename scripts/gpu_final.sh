#!/bin/bash
# round-end visit: full GPU suite, smoke, every bench config (both arms), ncu launch lists and the
# `--set full` captures of the kernels that changed this session (fused n-best kernel, split pack kernel)
OUT=gpurun_out/final2
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "default bench exit $?"
for c in 1 2 3 4 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 > $OUT/bench_cfg$c.json 2> $OUT/bench_cfg$c.err; echo "cfg$c exit $?"
  timeout 900 python bench.py --config $c --impl reference --steps 5 --warmup 1 --no-cpu-baseline > $OUT/ref_cfg$c.json 2> $OUT/ref_cfg$c.err; echo "ref cfg$c exit $?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
    --log-file $OUT/launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file $OUT/launches_cfg5.csv python bench.py --config 5 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch5.log 2>&1
cap() {  # cap <name> <kernel regex> <skip> <bench args...>
  local name=$1 k=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 \
      -f -o $OUT/$name python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > $OUT/$name.log 2>&1
  echo "$name exit $?"
}
cap fused lev_bv_fused_kernel 3 --config 2
cap pack_split lev_pack_seqfirst_kernel 4 --config 5
python - <<PY
import json
for c in range(1,6):
    for kind in ("bench","ref"):
        try:
            d=json.loads(open("$OUT/%s_cfg%d.json"%(kind,c)).read().strip().splitlines()[-1])
            r=d.get("roofline",{})
            print(kind, c, "GCUPS", round(d["value"],3), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],2), "roofline", r.get("kernel","")[:28], round(r.get("frac",0),3), "lit", (d.get("literal") or {}).get("ms_per_call"), d.get("reference_on_gpu",{}).get("value"))
        except Exception as e:
            print(kind, c, "failed", e)
PY
