for w in 4 8 16; do echo "== warps $w"; B200LEV_CTA_WARPS=$w python scripts/bench_k2_micro.py 2>&1 | grep -E '"T": 2000'; B200LEV_CTA_WARPS=$w python scripts/bench_cfg5.py 5 | head -1; done
