#!/bin/bash
# fused bit-vector kernel: parity on the GPU, then bench A/B (fused vs two-kernel form)
OUT=gpurun_out/${1:-fused}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bitvec or nbest or cfg2 or graph or routes" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_fused.json 2> $OUT/bench_fused.err; echo "exit $?" >> $OUT/bench_fused.err
B200LEV_BV_FUSED=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_two.json 2> $OUT/bench_two.err
python - <<PY
import json
for n in ("fused","two"):
    try:
        d=json.load(open("$OUT/bench_%s.json"%n))
        print(n, "ms", round(d["ms_per_step"],4), "GCUPS", round(d["value"],1), "phases", d["phases_ms"], "literal", round(d["literal"]["ms_per_call"],4), "e2e", round(d["e2e"]["ms_per_step"],3))
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 $OUT/bench_fused.err
