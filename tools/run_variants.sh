# usage: tools/run_variants.sh <suffix> ...   (tuning builds rvspecfit_b200/librvs_b200<suffix>.so)
for v in "$@"; do
  echo "variant $v"
  RVS_LIB=rvspecfit_b200/librvs_b200$v.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tma or desi_three" 2>&1 | tail -1
  RVS_LIB=rvspecfit_b200/librvs_b200$v.so timeout 200 python bench.py --stage-profile --evals 100 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())['stage_profile']
print({k:round(v['us_per_launch'],1) for k,v in d.items() if isinstance(v,dict)})"
done
