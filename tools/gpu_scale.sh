#!/bin/bash
# 8 GPUs of one box: the bench at N = 1, 2, 4, 8 back to back (what the driver does at round end) + the two-rank test.
tag=${1:-scale}
out=gpurun_out; mkdir -p $out
nvidia-smi -L > $out/${tag}_gpus.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "exit $?" >> $out/${tag}_pytest.log
tail -4 $out/${tag}_pytest.log
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $out/${tag}_bench_N1.json 2> $out/${tag}_bench_N1.err
grep -E "device-resident|end-to-end" $out/${tag}_bench_N1.err
for n in 2 4 8; do
  QSFT_K4_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 10 --warmup 3 > $out/${tag}_bench_N$n.json 2> $out/${tag}_bench_N$n.err
  grep -E "per-step|device-resident|end-to-end" $out/${tag}_bench_N$n.err
done
# A/B at N = 8: replicated peel (U gathered through the multicast mapping)
QSFT_K4_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 8 --steps 10 --warmup 3 --peel-mode replicated > $out/${tag}_bench_N8_replicated.json 2> $out/${tag}_bench_N8_replicated.err
grep -E "device-resident|end-to-end" $out/${tag}_bench_N8_replicated.err
grep "qsft_peel_loop" $out/${tag}_bench_N8.err | tail -3
