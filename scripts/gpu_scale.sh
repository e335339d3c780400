#!/bin/bash
# scaling visit on N GPUs: bench cfg2 (weak) and cfg4 (strong, all-reduce in the step), both arms
N=${1:-2}
OUT=gpurun_out/scale
mkdir -p $OUT
for c in 2 4; do
  if [ "$N" = 1 ]; then
    timeout 900 python bench.py --gpus 1 --config $c --steps 20 --warmup 5 --no-cpu-baseline > $OUT/cfg${c}_n$N.json 2> $OUT/cfg${c}_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 \
      bench.py --gpus $N --config $c --steps 20 --warmup 5 > $OUT/cfg${c}_n$N.json 2> $OUT/cfg${c}_n$N.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/cfg${c}_n$N.json").read().strip().splitlines()[-1])
    print("cfg$c N=$N", "GCUPS", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "e2e16", round((d.get("e2e_int16_tokens") or {}).get("value",0),1), d.get("collective"), d["scaling"])
except Exception as e:
    print("cfg$c N=$N failed", e)
PY
done
