#!/bin/bash
tag=${1:-r3e}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lattice" > $out/${tag}_pytest_lattice.log 2>&1; echo "exit $?" >> $out/${tag}_pytest_lattice.log
tail -3 $out/${tag}_pytest_lattice.log
timeout 300 python tools/q3_bench.py > $out/${tag}_q3_bench.json 2> $out/${tag}_q3_bench.err
cat $out/${tag}_q3_bench.json; tail -3 $out/${tag}_q3_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'lt_gemm_kernel' -s 12 -c 1 -o $out/${tag}_k2l_q3 \
    python tools/q3_bench.py > /dev/null 2>&1
ls -la $out | tail -3
