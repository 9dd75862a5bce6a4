#!/bin/bash
tag=${1:-c}; shift
mkdir -p gpurun_out
timeout 900 python -W ignore -m pytest tests/test_gpu_ccf.py -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/${tag}_pytest.log
timeout 900 python -W ignore bench.py --workload gaia_rvs --mode ccf "$@" > gpurun_out/${tag}_ccf.json 2> gpurun_out/${tag}_ccf.err
echo "ccf bench exit $?"; tail -c 3000 gpurun_out/${tag}_ccf.json; tail -5 gpurun_out/${tag}_ccf.err
