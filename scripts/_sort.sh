mkdir -p gpurun_out/sort
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lev_sort_kernel -s 5 -c 1 -f -o gpurun_out/sort/prof_sort python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sort/ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lev_prefix_finalize -s 5 -c 1 -f -o gpurun_out/sort/prof_fin python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sort/ncu2.log 2>&1
