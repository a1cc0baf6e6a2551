#!/bin/bash
# 8-GPU call: BASELINE configs[3] (elasticity 60^3 x 3, 8 GPUs) and configs[4] (Poisson 80^3 solve, nrhs 1/16/64, 8 GPUs)
set -x
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29618"
timeout 150 $TR bench.py --gpus $N --workload p3d80 --nrhs 1,16,64 --steps 3 --warmup 3 --no-extra > gpurun_out/r02u_bench_n8_p3d80.json 2> gpurun_out/r02u_bench_n8_p3d80.err
echo "rc=$?" >> gpurun_out/r02u_bench_n8_p3d80.err
timeout 240 $TR bench.py --gpus $N --workload el3d60 --steps 3 --warmup 3 --no-extra > gpurun_out/r02u_bench_n8_el3d60.json 2> gpurun_out/r02u_bench_n8_el3d60.err
echo "rc=$?" >> gpurun_out/r02u_bench_n8_el3d60.err
echo done
