#!/bin/bash
# 8 GPUs of one box: the bench at N = 1, 2, 4, 8 back to back (what the driver does at round end).
tag=${1:-scale}
out=gpurun_out; mkdir -p $out
nvidia-smi -L > $out/${tag}_gpus.txt 2>&1
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $out/${tag}_bench_N1.json 2> $out/${tag}_bench_N1.err
grep -E "device-resident|end-to-end" $out/${tag}_bench_N1.err
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 10 --warmup 3 > $out/${tag}_bench_N$n.json 2> $out/${tag}_bench_N$n.err
  grep -E "per-step|device-resident|end-to-end" $out/${tag}_bench_N$n.err
done
