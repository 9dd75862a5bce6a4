#!/bin/bash
# kernel-study runs in proxy mode: bash tools/gpu_proxy.sh <tag> "args" "args" ...
tag=$1; shift
mkdir -p gpurun_out
i=0
for run in "$@"; do
  i=$((i+1))
  timeout 600 python -W ignore bench.py --mode proxy --no-cpu --no-e2e $run > gpurun_out/${tag}_p$i.json 2> gpurun_out/${tag}_p$i.err
  echo "== $run exit $?"
  python - <<PY
import json
for line in open('gpurun_out/${tag}_p$i.json'):
    if line.startswith('{'):
        d = json.loads(line)
        if 'stage_profile' in d:
            print({k: (round(v['us_per_launch'], 1) if isinstance(v, dict) else v) for k, v in d['stage_profile'].items()})
        else:
            k = d['kernels']; r = d['roofline']
            print('value %.1f  step %.1f ms  frac %.3f  ms/call %.3f items/call %s scan %.2f ms/launch' % (d['value'], d['ms_per_step'], r['frac'], r['ms_per_call'], r['items_per_call'], k.get('scan_ms_per_launch', 0)))
PY
  tail -2 gpurun_out/${tag}_p$i.err
done
