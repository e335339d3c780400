bash scripts/gpu_check.sh r1m "test bench"
timeout 900 python scripts/bench_configs.py > gpurun_out/r1m/configs.jsonl 2> gpurun_out/r1m/configs.err; tail -3 gpurun_out/r1m/configs.err; cat gpurun_out/r1m/configs.jsonl
