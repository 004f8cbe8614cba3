#!/bin/bash
tag=${1:-prof2}; N=${2:-2}
out=gpurun_out; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/profile_dist.py > $out/${tag}_profile_dist.txt 2>&1
head -120 $out/${tag}_profile_dist.txt
