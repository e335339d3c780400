mkdir -p gpurun_out/occ
for n in 4 5 6; do
  B200LEV_NVCC_EXTRA="-DLEVG_MIN_CTAS=$n" python pydrobert-pytorch_b200/build.py --force > /dev/null 2>&1
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/occ/bench_$n.json 2>gpurun_out/occ/err_$n.txt
  python -c "
import json
d=json.load(open('gpurun_out/occ/bench_$n.json'))
print('MIN_CTAS=$n step ms', round(d['ms_per_step'],4), 'dp_only ms', round(d['roofline']['kernel_ms'],4), 'pack ms', round(d['roofline_pack']['kernel_ms'],4))"
done
python pydrobert-pytorch_b200/build.py --force > /dev/null 2>&1
