#!/bin/bash
# fit-mode tuning sweeps.  usage: bash tools/gpu_fit.sh <tag> "B settings..." "B settings..." ...
tag=${1:-f}; shift
mkdir -p gpurun_out
if [ -n "$RVS_PYTEST" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
  tail -5 gpurun_out/${tag}_pytest.log
fi
i=0
for run in "$@"; do
  i=$((i+1))
  timeout 1200 python -W ignore tools/tune_fit.py $run > gpurun_out/${tag}_tune$i.log 2>&1
  echo "tune $run exit $?"; grep -v Warning gpurun_out/${tag}_tune$i.log | head -40
done
