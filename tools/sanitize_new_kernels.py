"""Small invocations of the kernels added late in round 2 (odd-prime lattice operands, q = 2 lattice, incremental phase tables,
the result gather) -- meant to run under `compute-sanitizer --tool memcheck`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsft_b200  # noqa: E402
from qsft_b200 import ops, utils  # noqa: E402

dev = torch.device("cuda", 0)
rng = np.random.RandomState(0)
for (q, n, b, S, P) in [(5, 12, 5, 300, 4), (7, 8, 3, 37, 5), (2, 20, 14, 130, 1), (4, 12, 7, 257, 2), (3, 12, 7, 100, 3)]:
    M, D = rng.randint(0, q, size=(n, b)), rng.randint(0, q, size=(P, n))
    loc = rng.randint(0, q, size=(S, n))
    a = np.exp(2j * np.pi * rng.uniform(0, 1, S)).astype(np.complex64)
    ld = utils.padded_ld(n)
    out = ops.eval_synth_lattice(M, D, ops.pad_digits(loc, ld, dev), torch.from_numpy(a).to(dev), q, residual_passes=0)
    torch.cuda.synchronize()
    print("lattice", q, n, b, S, P, tuple(out.shape), float(out.abs().max()))
qa = {"query_method": "complex", "num_subsample": 3, "delays_method_source": "identity", "subsampling_method": "qsft",
      "delays_method_channel": "nso", "num_repeat": 2, "b": 4}
np.random.seed(1)
sig = qsft_b200.get_random_subsampled_signal(n=10, q=4, sparsity=60, a_min=1, a_max=1, noise_sd=0.0, query_args=dict(qa))
res = qsft_b200.QSFT(num_subsample=3, num_repeat=2, b=4, reconstruct_method_source="identity",
                     reconstruct_method_channel="nso").transform(sig, output="arrays")
print("transform", len(res["values"]), len(sig.signal_w))
