#!/bin/bash
# pack kernel split along the sequence axis (few, long sequences): parity, cfg5 / cfg3 numbers with and
# without the split, launch list of the cfg5 step
OUT=gpurun_out/split
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
for c in 5 3; do
  timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_cfg$c.json 2> $OUT/bench_cfg$c.err; echo "cfg$c exit $?"
  B200LEV_PACK_SPLIT=0 timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_cfg${c}_nosplit.json 2> $OUT/bench_cfg${c}_nosplit.err; echo "cfg$c nosplit exit $?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_cfg5.csv \
   python bench.py --config 5 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1
python - <<PY
import json
for n in ("bench_cfg5","bench_cfg5_nosplit","bench_cfg3","bench_cfg3_nosplit"):
    try:
        d=json.loads(open("$OUT/%s.json"%n).read().strip().splitlines()[-1])
        print(n, "ms", round(d["ms_per_step"],4), "GCUPS", round(d["value"],1), "lit", (d.get("literal") or {}).get("ms_per_call"), {k:v for k,v in d["phases_ms"].items() if v})
    except Exception as e:
        print(n, "failed", e)
PY
