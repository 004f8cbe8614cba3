"""CPU: the DEVICE source of K1 (query lattice, index codecs; k1_lattice.cu) and of the SIMT evaluation kernel
(k2_eval_simt.cu) executed by the SIMT emulation in tests/emu, bit-exact against the index fixtures of the unmodified
reference (up to 100-bit indices) and, for the samples, against the reference's own samples.  Test infrastructure: the
product has no CPU path."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import qsft_oracle as orc
from conftest import FULL_CASES, INDEX_CASES, WIDE_FULL_CASES, case_params, load_golden, u128_to_ints

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import build_emu  # noqa: E402


@pytest.fixture(scope="module")
def k1():
    L = C.CDLL(build_emu.build(which="k1"))
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_longlong
    L.emu_query_lattice.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, vp, i32, i32]
    L.emu_dec_to_qary.argtypes = [vp, i32, i64, i32, i32, vp, i32]
    L.emu_qary_to_dec.argtypes = [vp, i32, i64, i32, i32, vp, i32]
    return L


@pytest.fixture(scope="module")
def k2():
    L = C.CDLL(build_emu.build(which="k2"))
    L.emu_eval_synth.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def padded_ld(n):
    return max(32, (n + 31) // 32 * 32)


def limbs_of(q, n):
    return 1 if q ** n - 1 < (1 << 64) else 2


def lattice(L, M, D, q, num_sms=148):
    M, D = np.ascontiguousarray(M, dtype=np.int8), np.ascontiguousarray(D, dtype=np.int8)
    n, b = M.shape
    P, B, ld, limbs = D.shape[0], q ** b, padded_ld(n), limbs_of(q, n)
    idx = np.zeros((P, B, limbs), dtype=np.uint64)
    dig = np.full((P, B, ld), 77, dtype=np.int8)
    assert L.emu_query_lattice(_p(M), _p(D), q, n, b, P, _p(idx), limbs, _p(dig), ld, num_sms) == 0
    return idx, dig


@pytest.mark.parametrize("name", INDEX_CASES)
def test_emulated_k1_wide_indices_bit_exact(k1, name):
    g = load_golden(name)
    q, n, b, P = (int(v) for v in g["meta"])
    for sms in (148, 1):                                   # with / without splitting the delay rows over blocks
        idx, dig = lattice(k1, g["M"], g["D"], q, sms)
        if idx.shape[-1] == 2:
            assert np.array_equal(idx[..., 0], g["hi"]) and np.array_equal(idx[..., 1], g["lo"])
        else:
            assert np.array_equal(idx[..., 0], g["lo"]) and not g["hi"].any()
        want = orc.query_digits(g["M"], g["D"], q).transpose(0, 2, 1)
        assert np.array_equal(dig[..., :n], want) and not dig[..., n:].any()
    # codecs round trip on the device source
    flat = np.ascontiguousarray(idx.reshape(-1, idx.shape[-1]))
    d2 = np.full((flat.shape[0], dig.shape[-1]), 55, dtype=np.int8)
    assert k1.emu_dec_to_qary(_p(flat), flat.shape[1], flat.shape[0], q, n, _p(d2), d2.shape[1]) == 0
    assert np.array_equal(d2, dig.reshape(d2.shape))
    i2 = np.zeros_like(flat)
    assert k1.emu_qary_to_dec(_p(d2), d2.shape[1], flat.shape[0], q, n, _p(i2), flat.shape[1]) == 0
    assert np.array_equal(i2, flat)
    assert np.array_equal(d2[:64, :n].T, g["digits64"])


@pytest.mark.parametrize("name", FULL_CASES + WIDE_FULL_CASES)
def test_emulated_k1_k2_group00_matches_reference(k1, k2, name):
    """Lattice indices of group (0, 0) and the samples of its delay row 1, as produced by the reference."""
    g = load_golden(name)
    p = case_params(g)
    q, n = p["q"], p["n"]
    idx, dig = lattice(k1, g["Ms"][0], g["Ds"][0], q)
    if idx.shape[-1] == 2:
        assert np.array_equal(idx[..., 0], g["idx00_hi"]) and np.array_equal(idx[..., 1], g["idx00_lo"])
    else:
        assert np.array_equal(idx[..., 0], g["idx00_lo"]) and not g["idx00_hi"].any()
    ld = dig.shape[-1]
    loc = np.zeros((g["locq"].shape[1], ld), dtype=np.int8)
    loc[:, :n] = g["locq"].T
    a = np.ascontiguousarray(g["strengths"].astype(np.complex64))
    qd = np.ascontiguousarray(dig[1])
    out = np.zeros(qd.shape[0], dtype=np.complex64)
    assert k2.emu_eval_synth(_p(qd), qd.shape[0], _p(loc), _p(a), loc.shape[0], q, n, ld, _p(out)) == 0
    want = g["samples00_row1"]
    assert np.max(np.abs(out - want)) <= 1e-5 * max(1.0, np.max(np.abs(want)))


@pytest.mark.parametrize("q,n,S,N", [(3, 30, 257, 601), (7, 22, 300, 512), (5, 6, 1, 10), (4, 20, 513, 1), (2, 100, 64, 300)])
def test_emulated_k2_vs_oracle(k2, q, n, S, N):
    rng = np.random.default_rng(q * 1000 + n)
    qd, locq = rng.integers(0, q, (N, n)), rng.integers(0, q, (n, S))
    a = rng.uniform(0.5, 2, S) * np.exp(1j * rng.uniform(0, 2 * np.pi, S))
    want = orc.synth_eval_digits(qd, locq, a, q)
    ld = padded_ld(n)
    Q = np.zeros((N, ld), dtype=np.int8)
    Q[:, :n] = qd
    Lc = np.zeros((S, ld), dtype=np.int8)
    Lc[:, :n] = locq.T
    a32 = np.ascontiguousarray(a.astype(np.complex64))
    out = np.zeros(N, dtype=np.complex64)
    assert k2.emu_eval_synth(_p(Q), N, _p(Lc), _p(a32), S, q, n, ld, _p(out)) == 0
    scale = np.sqrt(np.sum(np.abs(a) ** 2))
    assert np.max(np.abs(out - want)) <= 2e-6 * scale + 1e-6 * np.max(np.abs(want))
