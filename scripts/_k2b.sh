python scripts/bench_k2_micro.py 2>&1 | grep -E '"N": 1,|"N": 592|"N": 148'
python scripts/bench_cfg5.py 5
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
