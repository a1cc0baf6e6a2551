#!/bin/bash
# first GPU call of round 2: full gpu test suite, headline bench, sanitizer runs, ncu launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
nproc > gpurun_out/r02a_nproc.txt; free -g >> gpurun_out/r02a_nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?" >> gpurun_out/r02a_bench.err
for tool in racecheck memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/tools/sanitize_case.py p3d-12 el3d-6 > gpurun_out/r02a_sanitizer_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/r02a_sanitizer_$tool.log
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02a_launches.csv \
   python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02a_ncu_bench.log 2>&1
echo done
