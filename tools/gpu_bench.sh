#!/bin/bash
# default bench (+ optional reference arm and profiler passes).  usage: bash tools/gpu_bench.sh <tag> [bench args]
tag=${1:-b}; shift
mkdir -p gpurun_out
timeout 1500 python -W ignore bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; tail -c 4500 gpurun_out/${tag}_bench.json; grep -v Warning gpurun_out/${tag}_bench.err | tail -5
if [ -n "$RVS_REF" ]; then
  timeout 900 python -W ignore bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err
  echo "ref exit $?"; tail -c 1500 gpurun_out/${tag}_ref.json
fi
if [ -n "$RVS_NCU" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/${tag}_launches.csv python -W ignore bench.py --batch 512 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/${tag}_ncu_launch.log 2>&1
  echo "ncu launch exit $?"; wc -l gpurun_out/${tag}_launches.csv
fi
