"""Model-quality evaluation on held-out samples (mirror of TestHelper._test_qary, qsft/test_helper.py:235-260).

The reference re-evaluates the recovered sparse model  y_hat[m] = sum_k beta[k] w^<m, k>  on test queries with a dense
NumPy matrix product; here it is one call of the evaluation kernel (K2, arbitrary non-lattice queries)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .utils import index_limbs, ints_to_limbs, padded_ld


def evaluate_model(beta, sample_idx_dec, q, n, device=None):
    """y_hat for decimal query indices (Python ints, up to 128 bits) under the sparse model `beta` {tuple(k): coef}."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    if len(beta) == 0 or len(sample_idx_dec) == 0:
        return np.zeros(len(sample_idx_dec), dtype=complex)
    ld = padded_ld(n)
    keys = np.array(list(beta.keys()), dtype=np.int8).reshape(len(beta), n)
    vals = np.array(list(beta.values())).astype(np.complex64)
    loc = ops.pad_digits(keys, ld, device)
    a = torch.from_numpy(vals).to(device)
    limbs = ints_to_limbs(sample_idx_dec, index_limbs(q, n))
    idx = torch.from_numpy(limbs.view(np.int64)).to(device)
    dig = ops.dec_to_qary(idx, q, n, ld)
    return ops.eval_synth(dig, loc, a, q, n).cpu().numpy().astype(complex)


def test_nmse(beta, sample_idx_dec, samples, q, n, device=None):
    """|| y_hat - y ||^2 / || y ||^2 on the test set; 1 for an empty model (qsft/test_helper.py:235-260)."""
    if len(beta) == 0:
        return 1
    samples = np.asarray(samples)
    y_hat = evaluate_model(beta, list(sample_idx_dec), q, n, device)
    return float(np.linalg.norm(y_hat - samples) ** 2 / np.linalg.norm(samples) ** 2)


test_nmse.__test__ = False   # not a pytest test
