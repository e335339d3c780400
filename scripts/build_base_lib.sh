#!/bin/bash
# pydrobert-pytorch_b200/build/libb200lev_base.so = the in-tree objects with the COMMITTED version of one
# source file (default lev_bvfused.cu) in place of the working-tree one: the "base" arm of scripts/gpu_ab.sh
set -e
F=${1:-lev_bvfused.cu}
cd "$(dirname "$0")/.."
python pydrobert-pytorch_b200/build.py > /dev/null
git show HEAD:pydrobert-pytorch_b200/csrc/$F > pydrobert-pytorch_b200/csrc/_base_$F
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
     -c pydrobert-pytorch_b200/csrc/_base_$F -o /tmp/_base_${F%.cu}.o
rm pydrobert-pytorch_b200/csrc/_base_$F
objs=$(ls pydrobert-pytorch_b200/build/*.o | grep -v "/${F%.cu}.o")
nvcc -shared -o pydrobert-pytorch_b200/build/libb200lev_base.so $objs /tmp/_base_${F%.cu}.o -lcudart 2>/dev/null
ls -la pydrobert-pytorch_b200/build/libb200lev_base.so
