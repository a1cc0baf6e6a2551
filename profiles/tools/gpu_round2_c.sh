#!/bin/bash
# third GPU call (2 GPUs): real 2-rank factorization (peer memory), N=2 bench, N=1 sanity
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02c_topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi.py -m gpu -x -q --durations=10 > gpurun_out/r02c_pytest_multi.log 2>&1
echo "rc=$?" >> gpurun_out/r02c_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus 2 --workload p3d64 --steps 5 --warmup 3 --no-extra > gpurun_out/r02c_bench_n2_p3d64.json 2> gpurun_out/r02c_bench_n2_p3d64.err
echo "rc=$?" >> gpurun_out/r02c_bench_n2_p3d64.err
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-extra > gpurun_out/r02c_bench_n2_p3d100.json 2> gpurun_out/r02c_bench_n2_p3d100.err
echo "rc=$?" >> gpurun_out/r02c_bench_n2_p3d100.err
timeout 600 python bench.py --workload p3d64 --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02c_bench_n1_p3d64.json 2> gpurun_out/r02c_bench_n1_p3d64.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02c_bench_n1_p3d100.json 2> gpurun_out/r02c_bench_n1_p3d100.err
echo done
