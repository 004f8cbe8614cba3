#!/bin/bash
# Round-2 GPU call: K4 loop v2 (CTA-wide TMA tiles) + tuned K3 (deferred publish, packed f32x2 butterflies).
tag=${1:-r2b}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.csv 2>&1
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "k3" > $out/${tag}_pytest_k3.log 2>&1; echo "exit $?" >> $out/${tag}_pytest_k3.log
tail -4 $out/${tag}_pytest_k3.log
timeout 400 python -m pytest tests/test_gpu_zz_detectors.py -m gpu -q -k "device_loop" > $out/${tag}_pytest_k4loop.log 2>&1; echo "exit $?" >> $out/${tag}_pytest_k4loop.log
tail -15 $out/${tag}_pytest_k4loop.log
timeout 1200 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "exit $?" >> $out/${tag}_pytest.log
tail -15 $out/${tag}_pytest.log
timeout 300 python tools/microbench.py --only k3,k4 > $out/${tag}_microbench.json 2> $out/${tag}_microbench.err
cat $out/${tag}_microbench.json
tail -3 $out/${tag}_microbench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k3_q4_tma' -s 3 -c 1 -o $out/${tag}_k3tma \
    python tools/microbench.py --only k3 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k4_peel_loop' -s 2 -c 1 -o $out/${tag}_k4loop \
    python tools/microbench.py --only k4 > /dev/null 2>&1
ls -la $out | tail -8
