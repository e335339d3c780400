#!/bin/bash
# after the dynamic block claiming of the fused kernel: full GPU suite, default bench, ncu of the fused kernel
OUT=gpurun_out/final3
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "default bench exit $?"
timeout 600 python bench.py --config 2 --steps 20 --warmup 5 > $OUT/bench_cfg2.json 2> $OUT/bench_cfg2.err; echo "cfg2 exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
    --log-file $OUT/launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lev_bv_fused_kernel -s 3 -c 1 \
      -f -o $OUT/fused python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config 2 > $OUT/fused.log 2>&1; echo "ncu exit $?"
python - <<PY
import json
for n in ("bench_default","bench_cfg2"):
    d=json.loads(open("$OUT/%s.json"%n).read().strip().splitlines()[-1])
    print(n, "ms", round(d["ms_per_step"],4), "GCUPS", round(d["value"],1), "kernel", d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"], "lit", d["literal"]["ms_per_call"], d["phases_ms"])
PY
