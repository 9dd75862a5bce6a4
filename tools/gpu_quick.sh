#!/bin/bash
# parity tests + smoke + default bench (no profiler).  usage: bash tools/gpu_quick.sh <tag> [bench args]
tag=${1:-q}; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -30 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; tail -c 3500 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
