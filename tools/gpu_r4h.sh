#!/bin/bash
# odd primes through the lattice GEMM: parity tests + one-block timings against K1 + K2
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "lattice" --no-header 2>&1 | tail -15
timeout 300 python tools/qodd_bench.py > gpurun_out/r4h_qodd_bench.json 2> gpurun_out/r4h_qodd_bench.err; tail -3 gpurun_out/r4h_qodd_bench.err; cat gpurun_out/r4h_qodd_bench.json
