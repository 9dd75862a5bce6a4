#!/bin/bash
# Compare builds of the library (kernel tuning): bash tools/gpu_variants.sh <tag> lib1.so lib2.so ...
tag=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib"
  RVS_B200_LIB=$PWD/$lib timeout 600 python bench.py --batch 2048 --evals 100 --steps 2 --warmup 1 --no-cpu 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); k = d['kernels']
        print('value %.1f spectra/s  step %.1f ms  fused %.1f us/launch  scan %.2f ms/launch  frac %.3f' % (d['value'], d['ms_per_step'], 1e3*k['fused_ms_per_launch'], k['scan_ms_per_launch'], d['roofline']['frac']))
    elif 'Error' in line or 'error' in line: print(line.rstrip())
" | tee -a gpurun_out/${tag}_variants.txt
done
