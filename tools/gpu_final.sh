#!/bin/bash
# bench of record + profiler passes.  usage: bash tools/gpu_final.sh <tag> [bench args]
tag=${1:-b}; shift
mkdir -p gpurun_out
timeout 1500 python -W ignore bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; tail -c 5000 gpurun_out/${tag}_bench.json; grep -v Warning gpurun_out/${tag}_bench.err | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/${tag}_launches.csv python -W ignore bench.py --batch 1024 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/${tag}_ncu_launch.log 2>&1
echo "ncu launch exit $?"; wc -l gpurun_out/${tag}_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chunk_kernel|gram_mma_kernel|resid_mma_kernel|gram_solve_kernel|prep_kernel" -s 30 -c 10 \
  -o gpurun_out/${tag}_eval -f python -W ignore bench.py --mode proxy --batch 4096 --groups 2 --evals 8 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/${tag}_eval.ncu-rep
