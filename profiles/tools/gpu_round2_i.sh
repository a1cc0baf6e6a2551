#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi.py -m gpu -x -q --durations=5 > gpurun_out/r02i_pytest_multi.log 2>&1
echo "rc=$?" >> gpurun_out/r02i_pytest_multi.log
echo done
