#!/bin/bash
tag=${1:-r2v}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.csv 2>&1
timeout 300 python tools/k2l_probe.py > $out/${tag}_k2l_probe.json 2> $out/${tag}_k2l_probe.err
cat $out/${tag}_k2l_probe.json; tail -3 $out/${tag}_k2l_probe.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'lt_gemm_spts' -s 2 -c 1 -o $out/${tag}_k2l \
    python tools/k2l_probe.py > /dev/null 2>&1
ls -la $out | tail -3
