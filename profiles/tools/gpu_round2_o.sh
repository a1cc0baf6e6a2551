#!/bin/bash
set -x
mkdir -p gpurun_out
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614"
timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-extra > gpurun_out/r02p_bench_n$N.json 2> gpurun_out/r02p_bench_n$N.err
SPLLT_B200_PUSH_INLINE=1 timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-extra > gpurun_out/r02p_bench_n${N}_inline.json 2> gpurun_out/r02p_bench_n${N}_inline.err
echo done
