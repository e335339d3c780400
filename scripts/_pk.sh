mkdir -p gpurun_out/r1r
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lev_pack_kernel -s 9 -c 1 -f -o gpurun_out/r1r/prof_pack python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1r/ncu_pack.log 2>&1
