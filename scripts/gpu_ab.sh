#!/bin/bash
# A/B/... of several builds of the library on ONE box: every pydrobert-pytorch_b200/build/variants/*.so
# (scripts/build_variants.sh) takes the in-tree library's place in turn; alternating runs of one config.
#   usage: scripts/gpu_ab.sh [cfg] [rounds]
CFG=${1:-2}
N=${2:-2}
OUT=gpurun_out/ab
mkdir -p $OUT
LIB=pydrobert-pytorch_b200/b200lev/libb200lev.so
cp $LIB /tmp/_intree.so
for i in $(seq 1 $N); do
  for f in pydrobert-pytorch_b200/build/variants/*.so; do
    v=$(basename $f .so)
    cp $f $LIB
    timeout 300 python bench.py --config $CFG --steps 40 --warmup 5 --no-cpu-baseline > $OUT/${v}_$i.json 2> $OUT/${v}_$i.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/${v}_$i.json").read().strip().splitlines()[-1])
    print("$v", $i, "ms", round(d["ms_per_step"],4), "kernel", round(d["roofline"].get("kernel_ms"),4), "GCUPS", round(d["value"],1))
except Exception as e:
    print("$v", $i, "failed", e)
PY
  done
done
cp /tmp/_intree.so $LIB
