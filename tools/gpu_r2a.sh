#!/bin/bash
# Round-2 first GPU call: validates the TMA K3 pipeline and the on-device K4 loop, then measures.
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.csv 2>&1
# 1. the two new kernels first, each under its own timeout (a hang must not eat the call)
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "k3" > $out/${tag}_pytest_k3.log 2>&1; echo "exit $?" >> $out/${tag}_pytest_k3.log
tail -4 $out/${tag}_pytest_k3.log
timeout 400 python -m pytest tests/test_gpu_zz_detectors.py -m gpu -q -k "device_loop" > $out/${tag}_pytest_k4loop.log 2>&1; echo "exit $?" >> $out/${tag}_pytest_k4loop.log
tail -15 $out/${tag}_pytest_k4loop.log
QSFT_K4_NO_TMA=1 timeout 400 python -m pytest tests/test_gpu_zz_detectors.py -m gpu -q -k "device_loop" > $out/${tag}_pytest_k4loop_notma.log 2>&1; echo "exit $?" >> $out/${tag}_pytest_k4loop_notma.log
tail -4 $out/${tag}_pytest_k4loop_notma.log
# 2. the whole GPU suite (no -x: every failure is wanted)
timeout 1200 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "exit $?" >> $out/${tag}_pytest.log
tail -15 $out/${tag}_pytest.log
# 3. stand-alone kernel timings
timeout 300 python tools/microbench.py > $out/${tag}_microbench.json 2> $out/${tag}_microbench.err
cat $out/${tag}_microbench.json
QSFT_K3_IMPL=1 timeout 200 python tools/microbench.py --only k3 > $out/${tag}_microbench_k3old.json 2> $out/${tag}_microbench_k3old.err
cat $out/${tag}_microbench_k3old.json
# 4. bench line
timeout 600 python bench.py --no-extras > $out/${tag}_bench_N1.json 2> $out/${tag}_bench_N1.err
head -c 4000 $out/${tag}_bench_N1.json
# 5. ncu: full capture of the two new kernels out of the microbench
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k3_q4_tma' -s 3 -c 2 -o $out/${tag}_k3tma \
    python tools/microbench.py --only k3 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k4_peel_loop' -s 2 -c 2 -o $out/${tag}_k4loop \
    python tools/microbench.py --only k4 > /dev/null 2>&1
ls -la $out | tail -8
