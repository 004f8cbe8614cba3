"""Debug helper: the sparse lattice-evaluation kernel (default: A' generated into tensor memory) against the dense kernel
(QSFT_LATTICE_SPARSE=0) on small problems; both are exact integer arithmetic."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qsft_b200 import ops, utils  # noqa: E402

DEV = torch.device("cuda:0")
q = 4


def run(mode, M, D, loc_d, a_d):
    os.environ["QSFT_LATTICE_SPARSE"] = str(mode)
    out = ops.eval_synth_lattice(M, D, loc_d, a_d, q)
    torch.cuda.synchronize()
    return out


for (n, b, S, P, seed) in [(14, 7, 700, 4, 0), (14, 7, 700, 5, 0), (40, 8, 3000, 3, 1), (40, 10, 257, 2, 2), (33, 7, 1, 1, 4)]:
    rng = np.random.default_rng(seed)
    M, D = rng.integers(0, q, (n, b)), rng.integers(0, q, (P, n))
    loc = rng.integers(0, q, (n, S))
    a = rng.uniform(0.2, 2, S) * np.exp(1j * rng.uniform(0, 2 * np.pi, S))
    ld = utils.padded_ld(n)
    loc_d = ops.pad_digits(loc.T, ld, DEV)
    a_d = torch.from_numpy(a.astype(np.complex64)).to(DEV)
    dense = run(0, M, D, loc_d, a_d)
    for mode in (2,):
        sp = run(mode, M, D, loc_d, a_d)
        diff = (sp - dense).abs()
        B1 = 4 ** (b // 2)
        d3 = diff.view(P, B1, -1)
        print(f"mode={mode} n={n} b={b} S={S} P={P}: equal={torch.equal(sp, dense)} max|diff|={diff.max().item():.3e} "
              f"max|dense|={dense.abs().max().item():.3e}", flush=True)
        if not torch.equal(sp, dense):
            bad = (d3 > 0)
            print("  bad fraction per delay row:", [round(float(bad[p].float().mean()), 3) for p in range(P)])
            print("  bad fraction per l_lo 64-block:",
                  [round(float(bad[:, :, i * 64:(i + 1) * 64].float().mean()), 3) for i in range(min(8, d3.shape[2] // 64))])
            print("  bad fraction per l_hi 16-block:",
                  [round(float(bad[:, i * 16:(i + 1) * 16].float().mean()), 3) for i in range(min(8, B1 // 16))])
            print("  sample sp/dense:", sp[0, 0, :3].tolist() if sp.dim() == 3 else sp[0, :3].tolist(), dense[0, :3].tolist())
print("done")
