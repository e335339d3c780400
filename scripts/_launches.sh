mkdir -p gpurun_out/r2h
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2h/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2h/ncu_launch.log 2>&1
