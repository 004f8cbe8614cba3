"""Host-side index codecs and small helpers of the q-SFT path (mirror of the names in the reference's qsft/utils.py
so callers can switch imports).  All arithmetic that matters for throughput runs on the GPU (qsft_b200/ops.py);
these functions exist for Python-int interop and for tiny host-side bookkeeping."""
from __future__ import annotations

import json
import pickle
import zlib

import numpy as np

MASK64 = (1 << 64) - 1


def index_limbs(q: int, n: int) -> int:
    """Number of 64-bit limbs needed for indices in [0, q^n): 1 or 2 (raises beyond 128 bits)."""
    top = q ** n - 1
    if top < (1 << 64):
        return 1
    if top < (1 << 128):
        return 2
    raise ValueError(f"q^n = {q}^{n} needs more than 128 bits")


def padded_ld(n: int) -> int:
    """Digit-row stride used on the device: >= n, multiple of 32 (so it is also a legal int8 MMA K extent)."""
    return max(32, (n + 31) // 32 * 32)


def ints_to_limbs(values, limbs: int) -> np.ndarray:
    """Sequence of Python ints -> uint64 array (N,) [limbs == 1] or (N, 2) = (hi, lo) [limbs == 2]."""
    vals = [int(v) for v in values]
    if limbs == 1:
        return np.array(vals, dtype=np.uint64)
    out = np.empty((len(vals), 2), dtype=np.uint64)
    out[:, 0] = [v >> 64 for v in vals]
    out[:, 1] = [v & MASK64 for v in vals]
    return out


def limbs_to_ints(arr: np.ndarray) -> np.ndarray:
    """Inverse of ints_to_limbs; returns an object array of Python ints (what the reference hands to subsample())."""
    arr = np.asarray(arr)
    if arr.ndim >= 1 and arr.shape[-1] == 2 and arr.dtype == np.uint64 and arr.ndim == 2:
        out = np.empty(arr.shape[0], dtype=object)
        out[:] = [(int(h) << 64) | int(l) for h, l in arr]
        return out
    out = np.empty(arr.shape, dtype=object)
    out[...] = [int(v) for v in arr.ravel()]
    return out


def qary_vec_to_dec(x, q):
    """(n, N) MSB-first digit columns -> object array of Python ints (qsft/utils.py:74-76)."""
    x = np.asarray(x)
    acc = np.zeros(x.shape[1:], dtype=object)
    for row in x:
        acc = acc * q + row.astype(object)
    return acc


def dec_to_qary_vec(x, q, n):
    """Python ints -> (n, N) int digits, MSB first (qsft/utils.py:79-84)."""
    out = np.zeros((n, len(x)), dtype=int)
    for col, v in enumerate(x):
        v = int(v)
        for i in range(n - 1, -1, -1):
            v, out[i, col] = divmod(v, q)
    return out


def qary_ints(m, q, dtype=int):
    """All vectors of Z_q^m as columns in counting order (qsft/utils.py:107-108)."""
    c = np.arange(q ** m)
    return np.stack([(c // q ** (m - 1 - i)) % q for i in range(m)]).astype(dtype)


def sort_qary_vecs(qary_vecs):
    qary_vecs = np.array(qary_vecs)
    return qary_vecs[np.lexsort(qary_vecs.T[::-1, :])]


def calc_hamming_weight(qary_vecs):
    return np.sum(np.array(qary_vecs) != 0, axis=1)


def random_signal_strength_model(sparsity, a, b):
    """qsft/utils.py:161-164 -- consumes np.random in the reference's order (uniform magnitudes, then phases)."""
    magnitude = np.random.uniform(a, b, sparsity)
    phase = np.random.uniform(0, 2 * np.pi, sparsity)
    return magnitude * np.exp(1j * phase)


def save_data(data, filename):
    """zlib(9) + pickle, file-compatible with the reference cache (qsft/utils.py:189-197)."""
    with open(filename, "wb") as f:
        f.write(zlib.compress(pickle.dumps(data, pickle.HIGHEST_PROTOCOL), 9))


def load_data(filename):
    with open(filename, "rb") as f:
        return pickle.loads(zlib.decompress(f.read()))


class NpEncoder(json.JSONEncoder):
    """JSON encoder accepting NumPy scalars / arrays (config.json of the experiment harness, qsft/utils.py:199-207)."""

    def default(self, obj):
        if isinstance(obj, np.integer):
            return int(obj)
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        return super().default(obj)


def _gwht_device(x, q, n, inverse):
    """K3 on one dense vector: numpy / tensor of q^n complex values -> NumPy complex128 (same shape)."""
    import torch
    from . import ops
    if not torch.cuda.is_available():
        raise RuntimeError("qsft_b200 needs a CUDA device (there is no CPU fallback)")
    a = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    shape = a.shape
    if a.size != q ** n:
        raise ValueError(f"expected q^n = {q ** n} values, got {a.size}")
    v = np.ascontiguousarray(a.reshape(1, -1), dtype=np.complex64)
    if inverse:
        v = np.conj(v)
    t = ops.gwht_batch_(torch.from_numpy(v).cuda(), q, n)
    out = t.cpu().numpy().astype(np.complex128).reshape(shape)
    return np.conj(out) * float(q ** n) if inverse else out


def gwht(x, q, n):
    """Dense q-ary Fourier transform with forward scaling 1/q^n (qsft/utils.py:31-36), computed by the K3 kernel."""
    return _gwht_device(x, q, n, False)


def igwht(x, q, n):
    """Inverse of gwht (qsft/utils.py:44-49): ifftn * q^n = conj(gwht(conj(x))) * q^n."""
    return _gwht_device(x, q, n, True)


def gwht_tensored(x, q, n):
    """gwht of an array already shaped [q] * n (qsft/utils.py:38-41)."""
    return _gwht_device(x, q, n, False)


def igwht_tensored(x, q, n):
    return _gwht_device(x, q, n, True)
