#!/bin/bash
# 1-GPU call: DMMA many-rhs solve kernels (tests + timing), ncu --set full captures on the headline config
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "nrhs_sweep or solve_vs_oracle or solve_variants or many_rhs" > gpurun_out/r02e_pytest_solve.log 2>&1
echo "rc=$?" >> gpurun_out/r02e_pytest_solve.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "config5 or config2" > gpurun_out/r02e_pytest_full.log 2>&1
echo "rc=$?" >> gpurun_out/r02e_pytest_full.log
for nrhs in 16 64; do
  timeout 300 python bench.py --workload p3d80 --nrhs $nrhs --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02e_bench_p3d80_nrhs$nrhs.json 2> gpurun_out/r02e_bench_p3d80_nrhs$nrhs.err
done
# ncu --set full: three large tile launches near the end of the first factorization, both solve sweeps
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_tma --launch-skip 145 --launch-count 3 \
   -o gpurun_out/r02e_tile_tma_p3d100 -f python profiles/tools/one_factor_solve.py p3d100 1 0 > gpurun_out/r02e_ncu_tile.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve_pipe --launch-count 2 \
   -o gpurun_out/r02e_solve_pipe_p3d100 -f python profiles/tools/one_factor_solve.py p3d100 1 1 > gpurun_out/r02e_ncu_solve.log 2>&1
echo done
