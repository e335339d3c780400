#!/bin/bash
# A/B of two builds of the library on one box: pydrobert-pytorch_b200/build/libb200lev_base.so (the
# committed kernels) against the in-tree libb200lev.so; alternating runs of one bench config.
#   usage: scripts/gpu_ab.sh [cfg] [rounds]
CFG=${1:-2}
N=${2:-3}
OUT=gpurun_out/ab
mkdir -p $OUT
LIB=pydrobert-pytorch_b200/b200lev/libb200lev.so
cp $LIB /tmp/new.so
cp pydrobert-pytorch_b200/build/libb200lev_base.so /tmp/base.so
for i in $(seq 1 $N); do
  for v in new base; do
    cp /tmp/$v.so $LIB
    timeout 300 python bench.py --config $CFG --steps 40 --warmup 5 --no-cpu-baseline > $OUT/${v}_$i.json 2> $OUT/${v}_$i.err
    python - <<PY
import json
d=json.loads(open("$OUT/${v}_$i.json").read().strip().splitlines()[-1])
print("$v", $i, "ms", round(d["ms_per_step"],4), "kernel", d["roofline"].get("kernel_ms"), "GCUPS", round(d["value"],1))
PY
  done
done
cp /tmp/new.so $LIB
