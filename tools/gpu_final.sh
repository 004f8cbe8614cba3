#!/bin/bash
# End-of-round check on one GPU: whole GPU suite, smoke(), the default bench (with extras) and the reference arm, and one
# ncu --set full capture of the K3 pipeline at the config-5 block size.
tag=${1:-final}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "exit $?" >> $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log
timeout 900 python bench.py > $out/${tag}_bench_N1.json 2> $out/${tag}_bench_N1.err
tail -5 $out/${tag}_bench_N1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
tail -2 $out/${tag}_bench_reference.err; cut -c1-400 $out/${tag}_bench_reference.json
timeout 200 ncu --set full --clock-control none -k regex:k3_q4_tma -c 3 -o $out/${tag}_k3 python tools/microbench.py --only k3 > /dev/null 2>&1
ls -la $out/${tag}_k3.ncu-rep
