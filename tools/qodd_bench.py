"""q = 5 / 7: lattice evaluation (dense tcgen05 GEMM over Z[w], d = q - 1) against the plain K1 + K2 path, one (M, D) block."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qsft_b200 import ops, utils  # noqa: E402

dev = torch.device("cuda", 0)
out = {}
for (q, n, b, S, P) in [(5, 20, 6, 5000, 21), (5, 20, 8, 20000, 21), (7, 16, 5, 5000, 17), (7, 16, 7, 20000, 17)]:
    rng = np.random.RandomState(0)
    M = rng.randint(0, q, size=(n, b))
    D = rng.randint(0, q, size=(P, n))
    loc = rng.randint(0, q, size=(S, n))
    a = np.exp(2j * np.pi * rng.uniform(0, 1, S)).astype(np.complex64)
    ld = utils.padded_ld(n)
    loc_d = ops.pad_digits(loc, ld, dev)
    a_d = torch.from_numpy(a).to(dev)

    def lattice():
        return ops.eval_synth_lattice(M, D, loc_d, a_d, q, residual_passes=0)

    def plain():
        _, dig = ops.query_lattice(M, D, q, device=dev, want_idx=False, want_digits=True, ld=ld)
        return ops.eval_synth(dig.view(P * q ** b, ld), loc_d, a_d, q, n)

    res = {}
    for name, fn in (("lattice", lattice), ("plain K1 + K2", plain)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        res[name + " ms"] = round(e0.elapsed_time(e1) / 5, 4)
        del r
    res["pairs"] = float(P) * q ** b * S
    res["speed-up"] = round(res["plain K1 + K2 ms"] / res["lattice ms"], 2)
    d = q - 1
    res[f"int8 TOPS ({12 * d * d} ops / pair)"] = round(12 * d * d * res["pairs"] / (res["lattice ms"] * 1e-3) / 1e12, 1)
    out[f"q={q} n={n} b={b} S={S} P={P}"] = res
print(json.dumps(out, indent=1))
