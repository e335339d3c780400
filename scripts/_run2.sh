bash scripts/gpu_check.sh r1i "test bench ncu"
NCU_KERNEL=lev_pack_kernel NCU_SKIP=7 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lev_pack_kernel -s 7 -c 1 -f -o gpurun_out/r1i/prof_pack python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1i/ncu_pack.log 2>&1
