#!/bin/bash
# 2 GPUs: two-rank parity test, bench at N=2 (scatter + sharded on-device peel) with the peel kernel's own duration.
tag=${1:-multi2}
out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -k "two_ranks or k3" > $out/${tag}_pytest_multi.log 2>&1; echo "exit $?" >> $out/${tag}_pytest_multi.log
tail -5 $out/${tag}_pytest_multi.log
QSFT_K4_TIMING=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $out/${tag}_bench_N2.json 2> $out/${tag}_bench_N2.err
grep -E "per-step|device-resident|end-to-end" $out/${tag}_bench_N2.err
grep "qsft_peel_loop" $out/${tag}_bench_N2.err | tail -6
QSFT_K4_TIMING=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --gpus 2 --steps 10 --warmup 3 --peel-mode replicated > $out/${tag}_bench_N2_repl.json 2> $out/${tag}_bench_N2_repl.err
grep -E "device-resident|end-to-end" $out/${tag}_bench_N2_repl.err
grep "qsft_peel_loop" $out/${tag}_bench_N2_repl.err | tail -3
