#!/bin/bash
# N-GPU visit: NCCL test, then bench cfg4 (strong, all-reduce in the step) and cfg2 (weak) on N ranks
N=${1:-2}
OUT=gpurun_out/multi$N
mkdir -p $OUT
timeout 900 python -m pytest tests/test_dist_nccl.py -m gpu -x -q 2>&1 | tail -3
for c in 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 \
      bench.py --gpus $N --config $c --steps 20 --warmup 5 > $OUT/bench_cfg$c.json 2> $OUT/bench_cfg$c.err
  echo "cfg$c exit $?"; tail -2 $OUT/bench_cfg$c.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_cfg$c.json").read().strip().splitlines()[-1])
    print("cfg$c N=$N", "GCUPS", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d.get("check"), d.get("collective"), d["scaling"])
except Exception as e:
    print("failed", e)
PY
done
