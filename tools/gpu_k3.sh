#!/bin/bash
# K3 TMA pipeline: parity tests, stand-alone timing (default vs QSFT_K3_IMPL=1), ncu capture of the kernel.
tag=${1:-k3}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.csv 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "k3" > $out/${tag}_pytest_k3.log 2>&1; echo "exit $?" >> $out/${tag}_pytest_k3.log
tail -5 $out/${tag}_pytest_k3.log
timeout 200 python tools/microbench.py --only k3 > $out/${tag}_microbench.json 2> $out/${tag}_microbench.err
cat $out/${tag}_microbench.json
QSFT_K3_IMPL=1 timeout 200 python tools/microbench.py --only k3 > $out/${tag}_microbench_old.json 2> $out/${tag}_microbench_old.err
cat $out/${tag}_microbench_old.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k3_q4_tma' -s 3 -c 2 -o $out/${tag}_k3tma \
    python tools/microbench.py --only k3 > /dev/null 2>&1
ls -la $out | tail -5
