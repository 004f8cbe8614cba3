#!/bin/bash
# One gpurun call that (re)measures everything the round report needs, on ONE B200:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r2a'
# Outputs land in gpurun_out/<tag>_*; copy what you want judged into profiles/.
#   1. GPU parity tests (incl. tests/test_gpu_zz_detectors.py: detectors, nso2, harness, 100-bit pipeline)
#   2. stand-alone kernel timings incl. the K3 ticket-lag sweep (tools/microbench.py)
#   3. the default bench line (N = 1) + the reference arm
#   4. ncu: launch list of a short bench run, and one --set full capture of the K3 / K4 kernels
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
# tests of the code written without GPU time (detectors, nso2, harness, 100-bit pipelines): no -x, every failure is wanted
QSFT_TEST_UNVALIDATED=1 timeout 900 python -m pytest tests/test_gpu_zz_detectors.py -m gpu -q -k "not experimental" > $out/${tag}_pytest_unvalidated.log 2>&1
timeout 300 python tools/microbench.py > $out/${tag}_microbench.json 2> $out/${tag}_microbench.err
# opt-in kernels written without GPU time (K4 classification v2): parity against the default kernel, then its timings
QSFT_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_zz_detectors.py -k experimental -q > $out/${tag}_pytest_experimental.log 2>&1
QSFT_K4_IMPL=2 timeout 300 python tools/microbench.py > $out/${tag}_microbench_k4v2.json 2> $out/${tag}_microbench_k4v2.err
timeout 600 python bench.py > $out/${tag}_bench_N1.json 2> $out/${tag}_bench_N1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > $out/${tag}_bench_reference.json 2> /dev/null
# whole-step A/B of the opt-in variants (same box, back to back)
QSFT_LATTICE_EXPAND=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_bench_N1_expand.json 2> /dev/null
QSFT_K4_IMPL=2 QSFT_K4_FASTDET=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_bench_N1_k4v2.json 2> /dev/null
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_bench_N1_default_again.json 2> /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k3_q4_twopass|k4_classify|k4_apply|k4_reduce' -c 8 \
    -o $out/${tag}_k3k4 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > /dev/null 2>&1
tail -3 $out/${tag}_pytest.log
cat $out/${tag}_bench_N1.json | head -c 3000
