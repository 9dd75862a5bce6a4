#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu launch list, ncu full capture of the fused kernels.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [kernel regex]
tag=${1:-r1}
rx=${2:-prep_kernel|chunk_kernel|gram_mma_kernel|gram_solve_kernel|resid_mma_kernel|scan_mma}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; tail -c 3500 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/${tag}_launches.csv python bench.py --batch 2048 --evals 12 --steps 1 --warmup 1 --no-cpu > gpurun_out/${tag}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 14 -c 12 \
  -o gpurun_out/${tag}_fused -f python bench.py --batch 2048 --evals 4 --steps 1 --warmup 0 --no-cpu > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
