# usage: tools/run_variants.sh "<ENV=..> <lib suffix or -> <bench args>" ...
for v in "$@"; do
  echo "variant $v"
  set -- $v
  e=$1; shift
  l=$1; shift
  [ "$l" = "-" ] && l=""
  env $e RVS_LIB=rvspecfit_b200/librvs_b200$l.so timeout 300 python bench.py --no-cpu --evals 100 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), 'us/call', round(1e3*d['roofline']['ms_per_call'],1), 'scan ms', round(d['kernels']['scan_ms_per_launch'],2))"
done
