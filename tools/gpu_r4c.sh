#!/bin/bash
# q = 2 through the lattice GEMM: parity tests + one-block timings against K1 + K2
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "lattice or q2_large" --no-header 2>&1 | tail -15
timeout 300 python tools/q2_bench.py > gpurun_out/r4c_q2_bench.json 2> gpurun_out/r4c_q2_bench.err; tail -3 gpurun_out/r4c_q2_bench.err; cat gpurun_out/r4c_q2_bench.json
