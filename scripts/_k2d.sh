B200LEV_NVCC_EXTRA="-DLEV_CTA_CLOCK" python pydrobert-pytorch_b200/build.py --force > /dev/null 2>&1
python scripts/bench_k2_clock.py
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm --format=csv
python pydrobert-pytorch_b200/build.py --force > /dev/null 2>&1
