"""Host-side profile of one end-to-end transform through the public API (where do the non-kernel milliseconds go?)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsft_b200  # noqa: E402

q, n, b, S, C, R = 4, 40, 10, 100_000, 3, int(os.environ.get("R", "1"))
qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "identity", "subsampling_method": "qsft",
      "delays_method_channel": "nso", "num_repeat": R, "b": b}


def one(seed):
    np.random.seed(seed)
    sw, locq, st = qsft_b200.generate_signal_w(n, q, S, 1, 1, 0, full=False)
    Ms, Ds = qsft_b200.get_Ms_and_Ds(n, q, **qa)
    torch.cuda.synchronize()
    t0 = time.time()
    sig = qsft_b200.SyntheticSubsampledSignal(signal_w=sw, locq=locq, strengths=st, noise_sd=0.0, n=n, q=q,
                                              query_args=dict(qa), Ms=Ms, Ds=Ds, noise_rng="device")
    torch.cuda.synchronize()
    t1 = time.time()
    res = qsft_b200.QSFT(num_subsample=C, num_repeat=R, b=b, reconstruct_method_source="identity",
                         reconstruct_method_channel="nso").transform(sig, output=os.environ.get("OUT", "arrays"))
    torch.cuda.synchronize()
    t2 = time.time()
    return t1 - t0, t2 - t1, len(res["values"]) if isinstance(res, dict) and "values" in res else len(res)


for s in range(2):
    print("warm", one(s))
pr = cProfile.Profile()
pr.enable()
out = one(5)
pr.disable()
print("timed", out)
pstats.Stats(pr).sort_stats("cumtime").print_stats(28)
