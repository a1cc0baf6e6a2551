#!/bin/bash
# fourth GPU call (8 GPUs): headline config at N = 8 and N = 4, per-launch profiles of every rank
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02l_topo.txt 2>&1
for N in 8; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N"
  SPLLT_BENCH_PROFILE_CSV=gpurun_out/r02l_prof_n$N timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-extra > gpurun_out/r02l_bench_n$N.json 2> gpurun_out/r02l_bench_n$N.err
  echo "rc=$?" >> gpurun_out/r02l_bench_n$N.err
done
echo done
