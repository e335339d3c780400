#!/bin/bash
# one `ncu --set full` capture of a kernel of the bench step:  gpu_ncu.sh <tag> <kernel regex> [skip] [bench args...]
TAG=$1; KERNEL=$2; SKIP=${3:-5}; shift 3
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s $SKIP -c 1 \
    -f -o $OUT/prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
ls -la $OUT
