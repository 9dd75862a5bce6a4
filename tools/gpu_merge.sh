#!/bin/bash
tag=$1; shift
mkdir -p gpurun_out
timeout 1500 python -W ignore -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/${tag}_pytest.log
for env in "" "RVS_NO_MERGE=1"; do
  echo "== env: $env"
  env $env timeout 900 python -W ignore tools/tune_fit.py 4096 2:256:1 2>&1 | grep -v Warning | head -6
  env $env timeout 600 python -W ignore bench.py --mode proxy --no-cpu --no-e2e --batch 2048 --groups 2 --evals 200 --steps 2 --warmup 1 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); r = d['roofline']
        print('proxy value %.1f frac %.3f ms/call %.3f' % (d['value'], r['frac'], r['ms_per_call']))"
done
