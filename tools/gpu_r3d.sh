#!/bin/bash
tag=${1:-r3d}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lattice" > $out/${tag}_pytest_lattice.log 2>&1; echo "exit $?" >> $out/${tag}_pytest_lattice.log
tail -15 $out/${tag}_pytest_lattice.log
timeout 1200 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "exit $?" >> $out/${tag}_pytest.log
tail -6 $out/${tag}_pytest.log
timeout 300 python tools/bench_configs.py > $out/${tag}_bench_configs.json 2> $out/${tag}_bench_configs.err
cat $out/${tag}_bench_configs.json | head -30
