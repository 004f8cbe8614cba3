"""BASELINE configs 1 and 2 run IN FULL by the UNMODIFIED reference (no extrapolation): constructor (= query indices + sampling
+ FFT, input_signal_subsampled.py:107-155) and QSFT.transform, timed on this host's cores, with the same seeds, shapes and
noise levels as tools/bench_configs.py.  TEST INFRASTRUCTURE (like gen_golden.py), CPU only -- run where a copy of the reference exists (/root/reference or oracle/_ref).
    python oracle/reference_configs_cpu.py > profiles/r2/reference_cpu_configs_1_2.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_shim  # noqa: E402

root = ref_shim.reference_root()
ref_shim.install(root)
from qsft.qsft import QSFT  # noqa: E402
from synt_exp.synt_src.synthetic_signal import SyntheticSubsampledSignal, generate_signal_w  # noqa: E402

CONFIGS = [
    ("1: q=4 n=10 b=4 S=100 identity/identity", dict(n=10, q=4, S=100, b=4, C=3, R=1, src="identity", chan="identity", snr=None)),
    ("2: q=4 n=20 b=7 S=1000 nso R=3 20dB", dict(n=20, q=4, S=1000, b=7, C=3, R=3, src="identity", chan="nso", snr=20)),
]
only = set(sys.argv[1:])
out = {"host_cores": os.cpu_count(), "reference": root, "note": "unmodified reference, full runs, wall-clock seconds", "configs": []}
for name, c in CONFIGS:
    if only and name[0] not in only:
        continue
    noise_sd = 0.0 if c["snr"] is None else float(np.sqrt(c["S"] / 10 ** (c["snr"] / 10)))
    qa = {"query_method": "complex", "num_subsample": c["C"], "delays_method_source": c["src"], "subsampling_method": "qsft",
          "delays_method_channel": c["chan"], "num_repeat": c["R"], "b": c["b"]}
    np.random.seed(0)
    sw, locq, st = generate_signal_w(c["n"], c["q"], c["S"], 1, 1, 0, full=False)
    t0 = time.time()
    sig = SyntheticSubsampledSignal(signal_w=sw, locq=locq, strengths=st, noise_sd=noise_sd, n=c["n"], q=c["q"], query_args=dict(qa))
    t1 = time.time()
    res = QSFT(num_subsample=c["C"], num_repeat=c["R"], b=c["b"], reconstruct_method_source=c["src"],
               reconstruct_method_channel=c["chan"]).transform(sig, verbosity=0)
    t2 = time.time()
    found = len(set(res) & set(sw))
    diff = dict(sw)
    for k, v in res.items():
        diff[k] = diff.get(k, 0) - v
    nmse = float(np.sum(np.abs(list(diff.values())) ** 2) / np.sum(np.abs(list(sw.values())) ** 2))
    fft_s = float(sum(sum(d.values()) for row_ in sig.transformTimes for d in row_))      # the reference's own transformTimes
    row = {"config": name, "construct_s": t1 - t0, "of_which_fft_s": fft_s, "transform_s": t2 - t1, "found": found, "true": len(sw),
           "spurious": len(res) - found, "nmse": nmse}
    out["configs"].append(row)
    print(json.dumps(row), file=sys.stderr, flush=True)
print(json.dumps(out, indent=1))
