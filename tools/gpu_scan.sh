#!/bin/bash
tag=$1; shift
mkdir -p gpurun_out
timeout 1500 python -W ignore -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/${tag}_pytest.log
for env in "" "RVS_NO_PACK=1"; do
  echo "== env: $env"
  env $env timeout 600 python -W ignore bench.py --mode proxy --no-cpu --no-e2e --batch 2048 --groups 2 --evals 40 --steps 2 --warmup 1 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); r = d['roofline']; k = d['kernels']
        print('proxy value %.1f frac %.3f ms/call %.3f scan %.2f ms/launch (%d items)' % (d['value'], r['frac'], r['ms_per_call'], k['scan_ms_per_launch'], k['scan_items_per_launch']))"
done
timeout 900 python -W ignore tools/tune_fit.py 4096 2:256:1 3:256:1 4:256:1 2>&1 | grep -v Warning | head -12
