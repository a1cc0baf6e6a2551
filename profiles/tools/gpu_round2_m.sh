#!/bin/bash
# 1-GPU: parity after the look-ahead changes + N=1 benches
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multi.py -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r02m_pytest.log
timeout 300 python bench.py --workload p3d64 --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02m_bench_n1_p3d64.json 2> gpurun_out/r02m_bench_n1_p3d64.err
timeout 400 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02m_bench_n1_p3d100.json 2> gpurun_out/r02m_bench_n1_p3d100.err
SPLLT_B200_NO_CHAIN_AHEAD=1 timeout 400 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02m_bench_n1_p3d100_noahead.json 2> gpurun_out/r02m_bench_n1_p3d100_noahead.err
echo done
