#!/bin/bash
# N GPUs of one box: the GPU suite (incl. the two-rank parity test) and the bench at N ranks.
tag=${1:-multi}; N=${2:-2}
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/${tag}_gpus.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -m gpu -q > $out/${tag}_pytest_multi.log 2>&1; echo "exit $?" >> $out/${tag}_pytest_multi.log
tail -8 $out/${tag}_pytest_multi.log
for n in $N; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 10 --warmup 3 > $out/${tag}_bench_N$n.json 2> $out/${tag}_bench_N$n.err
  head -c 3000 $out/${tag}_bench_N$n.json; echo
  grep -E "per-step|end-to-end|device-resident" $out/${tag}_bench_N$n.err
done
