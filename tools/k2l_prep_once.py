"""One lattice evaluation of a config-5 block with ONE delay row (the operand generators at full size, a short GEMM): for ncu."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsft_b200  # noqa: E402
from qsft_b200 import ops, utils  # noqa: E402

dev = torch.device("cuda", 0)
q, n, b, S = 4, 40, 10, 100_000
np.random.seed(3)
sw, locq, strengths = qsft_b200.generate_signal_w(n, q, S, 1, 1, 0, full=False)
Ms, Ds = qsft_b200.get_Ms_and_Ds(n, q, query_method="complex", num_subsample=1, delays_method_source="identity",
                                 delays_method_channel="nso", num_repeat=1, b=b)
loc = ops.pad_digits(locq.T, utils.padded_ld(n), dev)
a = torch.from_numpy(strengths.astype(np.complex64)).to(dev)
D = np.vstack(Ds[0])[:int(os.environ.get("ROWS", "1"))]
out_t = torch.empty((D.shape[0], q ** b), dtype=torch.complex64, device=dev)
for _ in range(int(os.environ.get("REPS", "2"))):
    ops.eval_synth_lattice(Ms[0], D, loc, a, q, out=out_t, residual_passes=0)
torch.cuda.synchronize()
print("done")
