#!/bin/bash
# 2-GPU call: whole GPU suite (incl. the two-rank test), bench at N=1 and N=2 with the asynchronous device loop + balanced shards
tag=${1:-r2w}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.csv 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "exit $?" >> $out/${tag}_pytest.log
tail -6 $out/${tag}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $out/${tag}_bench_N1.json 2> $out/${tag}_bench_N1.err
tail -4 $out/${tag}_bench_N1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $out/${tag}_bench_N2.json 2> $out/${tag}_bench_N2.err
tail -4 $out/${tag}_bench_N2.err
