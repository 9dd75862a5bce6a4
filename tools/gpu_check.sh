#!/bin/bash
tag=$1; shift
mkdir -p gpurun_out
timeout 1500 python -W ignore -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/${tag}_pytest.log
timeout 900 python -W ignore tools/tune_fit.py "$@" 2>&1 | grep -v Warning | head -12
