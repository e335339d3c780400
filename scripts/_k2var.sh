for v in "" "-DLEV_CTA_NOFENCE" "-DLEV_CTA_PUB=32" "-DLEV_CTA_PUB=32 -DLEV_CTA_NOFENCE"; do
  B200LEV_NVCC_EXTRA="$v" python pydrobert-pytorch_b200/build.py --force > /dev/null 2>&1
  echo "== variant [$v]"; python scripts/bench_k2_micro.py 2>&1 | grep -E '"N": 1,|"N": 592'
done
python pydrobert-pytorch_b200/build.py --force > /dev/null 2>&1
