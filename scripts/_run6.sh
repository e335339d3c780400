bash scripts/gpu_check.sh r1p "test"
python scripts/bench_cfg5.py 5 > gpurun_out/r1p/cfg5.jsonl 2>&1; cat gpurun_out/r1p/cfg5.jsonl
timeout 600 python scripts/bench_configs.py > gpurun_out/r1p/configs.jsonl 2> gpurun_out/r1p/configs.err; tail -3 gpurun_out/r1p/configs.err; cat gpurun_out/r1p/configs.jsonl
