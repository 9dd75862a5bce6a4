#!/bin/bash
# the whole GPU suite (or selected files): bash tools/gpu_tests.sh <tag> [pytest args]
tag=${1:-t}; shift
mkdir -p gpurun_out
timeout 1500 python -W ignore -m pytest "${@:-tests}" -m gpu -q -x --durations=8 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -60 gpurun_out/${tag}_pytest.log
