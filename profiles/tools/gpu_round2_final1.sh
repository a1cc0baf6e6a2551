#!/bin/bash
# final 1-GPU evidence: full gpu test suite, default bench line (with extras and CPU baseline), ncu launch list
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/r02z_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02z_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r02z_bench_n1.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02z_launches.csv \
   python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02z_ncu_bench.log 2>&1
echo done
