"""CPU: the DEVICE source of the K3 kernels (qsft_b200/csrc/k3_gwht.cu) executed by the SIMT emulation in tests/emu and
compared with the oracle's gwht (= the reference's scipy.fft.fftn / q^b).  Covers the generic pass kernel, the q = 4
radix-16 kernels and the single-launch two-pass kernel for every ticket lag: the emulation runs the CTAs one after the
other in ticket order, so a ticket order that scheduled a strided tile before its dependencies would be reported
(on a GPU it would spin).  Test infrastructure: the product has no CPU path."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import qsft_oracle as orc

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import build_emu  # noqa: E402


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(build_emu.build(which="k3"))
    L.emu_gwht.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    return L


def run(L, x, q, b, lag=1, peer=False):
    y = np.ascontiguousarray(x.astype(np.complex64))
    pb = np.full_like(y, 7.0) if peer else None
    used = C.c_int(0)
    rc = L.emu_gwht(y.ctypes.data_as(C.c_void_p), y.shape[0], q, b, lag, None if pb is None else pb.ctypes.data_as(C.c_void_p),
                    C.byref(used))
    assert rc == 0, "a strided tile ran before its block's contiguous tiles"
    return y, pb, used.value


@pytest.mark.parametrize("q,b,batch", [(4, 7, 3), (4, 8, 2), (4, 6, 3), (4, 4, 5), (4, 2, 1), (3, 7, 2), (3, 8, 1), (2, 13, 1),
                                       (2, 5, 3), (5, 6, 1), (7, 3, 2), (11, 3, 2), (6, 4, 1)])
def test_emulated_gwht_vs_oracle(emu, q, b, batch):
    rng = np.random.default_rng(q * 100 + b)
    x = rng.normal(size=(batch, q ** b)) + 1j * rng.normal(size=(batch, q ** b))
    want = np.stack([orc.gwht(r, q, b) for r in x])
    got, peer, used = run(emu, x, q, b, lag=1, peer=True)
    tol = 2e-6 * np.max(np.abs(x)) / np.sqrt(q ** b) * np.sqrt(b) + 1e-7
    assert np.max(np.abs(got - want)) <= tol
    assert np.array_equal(peer, got)                       # fused all-gather: the peer buffer holds the same bits
    assert used == (1 if (q == 4 and 7 <= b <= 12) else 0)


@pytest.mark.parametrize("b,batch", [(7, 5), (8, 3)])
def test_emulated_twopass_every_lag(emu, b, batch):
    """Any ticket lag gives bit-identical results and never schedules a strided tile before its dependencies."""
    q = 4
    rng = np.random.default_rng(b)
    x = rng.normal(size=(batch, q ** b)) + 1j * rng.normal(size=(batch, q ** b))
    ref, _, used = run(emu, x, q, b, lag=-1)               # separate launches per pass
    assert used == 0
    for lag in [0, 1, 2, 3, batch, batch + 4]:
        got, _, used = run(emu, x, q, b, lag=lag)
        assert used == 1 and np.array_equal(got, ref), lag
    want = np.stack([orc.gwht(r, q, b) for r in x])
    assert np.max(np.abs(ref - want)) <= 1e-6
