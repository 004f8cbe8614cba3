#!/bin/bash
# Round-2 GPU call: K4 loop v3 (scanner / candidate warp roles) + K3 ring-depth sweep + e2e host profile.
tag=${1:-r2c}
out=gpurun_out
mkdir -p $out
timeout 400 python -m pytest tests/test_gpu_zz_detectors.py -m gpu -q -k "device_loop" > $out/${tag}_pytest_k4loop.log 2>&1; echo "exit $?" >> $out/${tag}_pytest_k4loop.log
tail -15 $out/${tag}_pytest_k4loop.log
timeout 1200 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "exit $?" >> $out/${tag}_pytest.log
tail -15 $out/${tag}_pytest.log
timeout 300 python tools/microbench.py --only k3,k4 > $out/${tag}_microbench.json 2> $out/${tag}_microbench.err
cat $out/${tag}_microbench.json
tail -3 $out/${tag}_microbench.err
#timeout 200 python tools/profile_e2e.py > $out/${tag}_profile_e2e.txt 2>&1
#OUT=dict timeout 200 python tools/profile_e2e.py > $out/${tag}_profile_e2e_dict.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k4_peel_loop' -s 2 -c 1 -o $out/${tag}_k4loop \
    python tools/microbench.py --only k4 > /dev/null 2>&1
ls -la $out | tail -6
