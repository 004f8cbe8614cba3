#!/bin/bash
# One GPU: the GPU test suite, the default bench line (with extras), the reference arm, launch list + K2L/K3/K4 ncu captures.
tag=${1:-bench}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "exit $?" >> $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
timeout 900 python bench.py > $out/${tag}_bench_N1.json 2> $out/${tag}_bench_N1.err
head -c 6000 $out/${tag}_bench_N1.json; echo
grep -E "per-step|end-to-end|device-resident" $out/${tag}_bench_N1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > $out/${tag}_bench_reference.json 2> /dev/null
if [ "$2" = "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
fi
ls -la $out | tail -5
