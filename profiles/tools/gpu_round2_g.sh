#!/bin/bash
# 1-GPU call: two-level inner updates (parity + N=1 benches), REDG in the pipelined solve, ncu capture of tile launches, solve trace
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multi.py -m gpu -x -q > gpurun_out/r02g_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r02g_pytest.log
timeout 300 python bench.py --workload p3d64 --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02g_bench_n1_p3d64.json 2> gpurun_out/r02g_bench_n1_p3d64.err
timeout 400 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02g_bench_n1_p3d100.json 2> gpurun_out/r02g_bench_n1_p3d100.err
timeout 300 python profiles/tools/profile_factor_csv.py p3d100 gpurun_out/r02g_prof_p3d100.csv > gpurun_out/r02g_prof_p3d100.txt 2>&1
timeout 400 python profiles/tools/solve_trace.py 100 0 0 > gpurun_out/r02g_solve_trace_100.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_tma --launch-skip 145 --launch-count 3 \
   -o gpurun_out/r02g_tile_tma_p3d100 -f python profiles/tools/one_factor_solve.py p3d100 1 0 > gpurun_out/r02g_ncu_tile.log 2>&1
echo done
