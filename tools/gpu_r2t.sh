#!/bin/bash
# Round-2 GPU call: K4 loop with private column copies / early stage release.
tag=${1:-r2t}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.csv 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "exit $?" >> $out/${tag}_pytest.log
tail -6 $out/${tag}_pytest.log
timeout 300 python tools/microbench.py --only k4 > $out/${tag}_microbench.json 2> $out/${tag}_microbench.err
cat $out/${tag}_microbench.json
tail -3 $out/${tag}_microbench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k4_peel_loop' -s 2 -c 1 -o $out/${tag}_k4loop \
    python tools/microbench.py --only k4 > /dev/null 2>&1
ls -la $out | tail -5
