"""Subsampling matrices M and delay matrices D (host side, NumPy RNG; mirrors qsft/query.py names and RNG order).

These are tiny (C x n x b and P x n integers) and must consume np.random exactly like the reference so that a seed
reproduces the same experiment; everything derived from them (lattices, samples, transforms) is computed on the GPU."""
from __future__ import annotations

import numpy as np

from .reed_solomon import ReedSolomon


def get_Ms_simple(n, b, q, num_to_get=None):
    """Block-identity Ms; block index runs DOWN so Ms[0] selects the last block (qsft/query.py:13-23)."""
    Ms = []
    for blk in reversed(range(num_to_get)):
        M = np.zeros((n, b), dtype=np.int32)
        M[b * blk:b * (blk + 1)] = np.eye(b, dtype=np.int32)
        Ms.append(M)
    return Ms


def get_Ms_complex(n, b, q, num_to_get=None):
    """Uniform random Ms (qsft/query.py:26-35)."""
    return [np.random.randint(q, size=(n, b)) for _ in range(num_to_get)]


def get_Ms(n, b, q, num_to_get=None, method="simple"):
    """qsft/query.py:38-70."""
    if num_to_get is None:
        num_to_get = max(n // b, 3)
    if method == "simple" and num_to_get > n // b:
        raise ValueError("When query_method is 'simple', the number of M matrices to return cannot be larger than n // b")
    gen = {"simple": get_Ms_simple, "complex": get_Ms_complex}.get(method)
    if gen is None:
        raise ValueError(f"unknown query_method {method!r} (expected 'simple' or 'complex')")
    return gen(n, b, q, num_to_get)


def get_D_identity(n, **kwargs):
    return np.vstack((np.zeros(n), np.eye(n))).astype(int)


def get_D_random(n, **kwargs):
    return np.random.choice(kwargs.get("q"), (kwargs.get("num_delays"), n))


def get_D_source_coded(n, **kwargs):
    return np.array(ReedSolomon(n, kwargs.get("t"), kwargs.get("q")).get_delay_matrix(), dtype=int)


def get_D_nso(n, D_source, **kwargs):
    """R random offsets minus the source delays (qsft/query.py:95-106)."""
    q = kwargs.get("q")
    offsets = get_D_random(n, q=q, num_delays=kwargs.get("num_repeat"))
    return [(row - D_source) % q for row in offsets]


def get_D_channel_coded(n, D, **kwargs):
    raise NotImplementedError("One day this might be implemented")


def get_D_channel_identity(n, D, **kwargs):
    return [D % kwargs.get("q")]


def get_D(n, **kwargs):
    """qsft/query.py:116-143: source delays then channel coding; returns a list of R (P_src, n) arrays."""
    src = {"random": get_D_random, "identity": get_D_identity, "coded": get_D_source_coded}.get(
        kwargs.get("delays_method_source", "random"))
    chan = {"nso": get_D_nso, "coded": get_D_channel_coded, "identity": get_D_channel_identity}.get(
        kwargs.get("delays_method_channel", "identity"))
    if src is None or chan is None:
        raise ValueError("unknown delays_method_source / delays_method_channel")
    return chan(n, src(n, **kwargs), **kwargs)


def get_Ms_and_Ds(n, q, **kwargs):
    """qsft/query.py:182-203: every M shares the same list of delay blocks."""
    Ms = get_Ms(n, kwargs.get("b"), q, method=kwargs.get("query_method"), num_to_get=kwargs.get("num_subsample"))
    D = get_D(n, q=q, **kwargs)
    return Ms, [D for _ in Ms]


def get_reed_solomon_dec(n, t_max, q):
    """Syndrome decoder callable for a t_max-error-correcting RS code (qsft/query.py:231-243).  The returned bound
    method carries its ReedSolomon object (`.__self__`), which QSFT uses to run the decoder on the GPU."""
    if q in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29):
        return ReedSolomon(n, t_max, q).syndrome_decode
    raise NotImplementedError("q is not a prime number under 30!")
