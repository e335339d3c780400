mkdir -p gpurun_out/bv
for k in ${1:-lev_bv_uid_kernel}; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 5 -c 1 -f -o gpurun_out/bv/prof_$k python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bv/ncu_$k.log 2>&1
done
