mkdir -p gpurun_out/r1n
python scripts/bench_cfg5.py 5 > gpurun_out/r1n/cfg5.jsonl 2>&1; cat gpurun_out/r1n/cfg5.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lev_cta_kernel -s 3 -c 1 -f -o gpurun_out/r1n/prof_cta python scripts/bench_cfg5.py 1 > gpurun_out/r1n/ncu.log 2>&1
