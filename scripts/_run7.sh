mkdir -p gpurun_out/r1q
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r1q/launches.csv python scripts/bench_cfg5.py 1 > gpurun_out/r1q/l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lev_cta_kernel -s 3 -c 1 -f -o gpurun_out/r1q/prof_cta python scripts/bench_cfg5.py 1 > gpurun_out/r1q/ncu.log 2>&1
