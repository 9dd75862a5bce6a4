#!/bin/bash
# A/B of environment switches on the fit: bash tools/gpu_ab.sh <tag> "B settings" ENV1 ENV2 ...
tag=$1; shift; run=$1; shift
mkdir -p gpurun_out
for env in "$@"; do
  echo "== env: $env"
  env $env timeout 900 python -W ignore tools/tune_fit.py $run 2>&1 | grep -v Warning | head -8 | cut -c1-330
done
