"""Host-side profile of a SMALL end-to-end transform (BASELINE config 2 shape): the GPU work is a fraction of a millisecond,
so what is left is the per-call cost of the host path (Python, ctypes, small uploads, synchronisations)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsft_b200  # noqa: E402

q, n, b, S, C, R = 4, 20, 7, 1000, 3, 3
qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "identity", "subsampling_method": "qsft",
      "delays_method_channel": "nso", "num_repeat": R, "b": b}
OUT = os.environ.get("OUT", "arrays")


def one(seed, timed=True):
    np.random.seed(seed)
    sw, locq, st = qsft_b200.generate_signal_w(n, q, S, 1, 1, 0, full=False)
    Ms, Ds = qsft_b200.get_Ms_and_Ds(n, q, **qa)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sig = qsft_b200.SyntheticSubsampledSignal(signal_w=sw, locq=locq, strengths=st, noise_sd=0.0, n=n, q=q,
                                              query_args=dict(qa), Ms=Ms, Ds=Ds, noise_rng="device")
    t1 = time.perf_counter()
    res = qsft_b200.QSFT(num_subsample=C, num_repeat=R, b=b, reconstruct_method_source="identity",
                         reconstruct_method_channel="nso").transform(sig, output=OUT)
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    return t1 - t0, t2 - t1


for s in range(5):
    one(s)
ts = np.array([one(10 + s) for s in range(50)]) * 1e3
print("construct / transform ms: median", np.median(ts, axis=0).round(3), "min", ts.min(axis=0).round(3))
pr = cProfile.Profile()
pr.enable()
for s in range(20):
    one(100 + s)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(45)
