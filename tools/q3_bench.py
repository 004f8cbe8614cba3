"""q = 3: lattice evaluation (dense tcgen05 GEMM over Z[w]) against the plain K1 + K2 path, one (M, D) block."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qsft_b200 import ops, utils  # noqa: E402

dev = torch.device("cuda", 0)
q = 3
out = {}
for (n, b, S, P) in [(30, 8, 5000, 33), (30, 10, 20000, 33), (30, 12, 50000, 11)]:
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
        if name != "lattice" and P * q ** b * S > 4e13:
            continue
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
    res["pairs"] = float(P) * q ** b * S
    if "plain K1 + K2 ms" in res:
        res["speed-up"] = round(res["plain K1 + K2 ms"] / res["lattice ms"], 2)
    # dense int8 ops executed by the lattice path: 2 launches x 3 limbs x 2 x 2 (matrix) x 2 ops per MAC and pair
    res["int8 TOPS (48 ops / pair)"] = round(48 * res["pairs"] / (res["lattice ms"] * 1e-3) / 1e12, 1)
    out[f"q=3 n={n} b={b} S={S} P={P}"] = res
print(json.dumps(out, indent=1))
