#!/bin/bash
# K3 TMA pipeline: ring depth sweep per shape
for st in 2 3; do
  QSFT_K3_STAGES=$st timeout 120 python tools/microbench.py --only k3 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('stages $st', {k.replace('k3_gwht ',''):(round(v['ms'],4), round(v.get('frac',0),3)) for k,v in d.items() if 'x 4^' in k})"
done
