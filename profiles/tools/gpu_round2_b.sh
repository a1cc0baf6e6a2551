#!/bin/bash
# second GPU call (1 GPU): emulated multi-rank factorization tests, per-launch profile of the headline config
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi.py -m gpu -x -q --durations=10 > gpurun_out/r02b_pytest_multi.log 2>&1
echo "rc=$?" >> gpurun_out/r02b_pytest_multi.log
timeout 600 python profiles/tools/profile_factor_csv.py p3d100 gpurun_out/r02b_prof_p3d100.csv > gpurun_out/r02b_prof_p3d100.txt 2>&1
timeout 600 python profiles/tools/profile_factor_csv.py p3d64 gpurun_out/r02b_prof_p3d64.csv > gpurun_out/r02b_prof_p3d64.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "factor_entries or not_positive" > gpurun_out/r02b_pytest_parity.log 2>&1
echo done
