"""Measurement aid: time the lattice GEMM (config 5, one (M, D) block of 41 delay rows) with its two operand feeds switched
off in turn (QSFT_LT_PROBE, see lt_gemm_spts_kernel): which feed keeps a stage above the MMA issue floor?"""
import json
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
D = np.vstack(Ds[0])
out_t = torch.empty((D.shape[0], q ** b), dtype=torch.complex64, device=dev)
res = {}
for probe in (0, 1, 2, 3, 0):
    os.environ["QSFT_LT_PROBE"] = str(probe)
    for _ in range(2):
        ops.eval_synth_lattice(Ms[0], D, loc, a, q, out=out_t, residual_passes=0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.eval_synth_lattice(Ms[0], D, loc, a, q, out=out_t, residual_passes=0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    stages = (2 * S + 255) // 256
    ctas_per_sm = (D.shape[0] * 2 * 1024 // 128) * 8 / 148
    key = {0: "default", 1: "no limb loads after the first ring lap", 2: "no A' stores after the first TMEM lap", 3: "neither feed"}[probe]
    res.setdefault(key, []).append({"ms_whole_call": round(ms, 3), "us_per_stage_approx": round(ms * 1e3 / (ctas_per_sm * stages), 4)})
os.environ.pop("QSFT_LT_PROBE")
print(json.dumps(res, indent=1))
