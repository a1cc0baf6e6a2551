#!/bin/bash
# 4-GPU call: headline config at N = 4 with per-rank per-launch profiles; then (GPU 0 only) the ncu --set full
# capture of three large tile launches at 100^3 and the solve task trace at 100^3
set -x
mkdir -p gpurun_out
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614"
SPLLT_BENCH_PROFILE_CSV=gpurun_out/r02k_prof_n$N timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-extra > gpurun_out/r02k_bench_n$N.json 2> gpurun_out/r02k_bench_n$N.err
echo "rc=$?" >> gpurun_out/r02k_bench_n$N.err
timeout 200 $TR bench.py --gpus $N --workload p3d64 --steps 5 --warmup 3 --no-extra > gpurun_out/r02k_bench_n${N}_p3d64.json 2> gpurun_out/r02k_bench_n${N}_p3d64.err
echo done
