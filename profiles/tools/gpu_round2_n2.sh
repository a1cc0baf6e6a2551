#!/bin/bash
set -x
mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612"
timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-extra > gpurun_out/r02z_bench_n$N.json 2> gpurun_out/r02z_bench_n$N.err
timeout 200 $TR bench.py --gpus $N --workload p3d64 --steps 5 --warmup 3 --no-extra > gpurun_out/r02z_bench_n${N}_p3d64.json 2> gpurun_out/r02z_bench_n${N}_p3d64.err
timeout 300 python -m pytest tests/test_multi.py -m gpu -x -q > gpurun_out/r02z_pytest_multi_2gpu.log 2>&1
echo "rc=$?" >> gpurun_out/r02z_pytest_multi_2gpu.log
echo done
