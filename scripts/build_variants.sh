#!/bin/bash
# one library per tuning setting of ONE source file, under pydrobert-pytorch_b200/build/variants/:
#   scripts/build_variants.sh lev_bvfused.cu name1 "-DX=1" name2 "-DX=2 -DY=1" ...
set -e
F=$1; shift
cd "$(dirname "$0")/.."
python pydrobert-pytorch_b200/build.py > /dev/null
mkdir -p pydrobert-pytorch_b200/build/variants
rm -f pydrobert-pytorch_b200/build/variants/*.so
objs=$(ls pydrobert-pytorch_b200/build/*.o | grep -v "/${F%.cu}.o")
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
       $flags -Xptxas -v -c pydrobert-pytorch_b200/csrc/$F -o /tmp/_var_$name.o 2>&1 | grep -A2 "${VARIANT_KERNEL:-kernelIlLi4ELi1E}" | grep "Used" || true
  nvcc -shared -o pydrobert-pytorch_b200/build/variants/$name.so $objs /tmp/_var_$name.o -lcudart 2>/dev/null
done
ls -la pydrobert-pytorch_b200/build/variants/
