#!/bin/bash
# N=4 sweep of the partition depth / look-ahead window (factor time only, quick)
set -x
mkdir -p gpurun_out
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614"
for cfg in "SPLLT_B200_SUBTREES_PER_RANK=2" "SPLLT_B200_SUBTREES_PER_RANK=3" "SPLLT_B200_TOP_WINDOW=4" "SPLLT_B200_BALANCE=1.04"; do
  env $cfg timeout 200 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-extra > gpurun_out/r02q_n4_$cfg.json 2> gpurun_out/r02q_n4_$cfg.err
done
echo done
