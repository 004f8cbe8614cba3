"""CPU: the DEVICE source of the K4 kernels (qsft_b200/csrc/k4_peel.cu: classification, reduce, apply, the stand-alone
detectors; k4_peel_loop.cu: the persistent on-device round loop as impl 2 / 3 = 128- / 32-bin tiles, 4 = as 2; one block) executed by the SIMT emulation in tests/emu (g++, one OS thread per CUDA thread) and
compared with the fixtures of the unmodified reference.  This checks the kernels' LOGIC without a GPU -- indexing,
reductions, decisions, the round loop; it says nothing about the memory model or speed, and it is test infrastructure:
the product has no CPU path."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import FULL_CASES, NSO2_CASES, WIDE_FULL_CASES, case_params, load_golden

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import build_emu  # noqa: E402


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(build_emu.build())
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_longlong, C.c_float
    desc = [i32] * 11 + [f32, vp, vp, vp, vp]
    L.emu_classify.argtypes = desc + [vp, vp, vp, vp, vp, vp, i64, i32, vp, i32, i64, i64]
    L.emu_peel.argtypes = desc + [vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, i64, i32,
                                  C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    L.emu_detect.argtypes = [vp, i64, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32]
    L.emu_mle.argtypes = [vp, i64, i32, vp, i32, vp, vp]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def padded_ld(n):
    return max(32, (n + 31) // 32 * 32)


class Problem:
    """Host mirror of ops.PeelProblem for the emulated kernels."""

    def __init__(self, q, n, b, Ms, Ds, P_src, channel, cutoff, rs=None):
        self.q, self.n, self.b, self.C = q, n, b, len(Ms)
        self.rs = rs
        if rs is not None:
            e, l = rs.device_tables()
            self.rs_exp, self.rs_log = np.ascontiguousarray(e, dtype=np.int32), np.ascontiguousarray(l, dtype=np.int32)
        Ds = np.asarray(Ds)
        self.P, self.P_src, self.B, self.ld = Ds.shape[1], P_src, q ** b, padded_ld(n)
        self.MT = np.zeros((self.C, b, self.ld), dtype=np.int8)
        for c, M in enumerate(Ms):
            self.MT[c, :, :n] = np.asarray(M).T
        self.D = np.zeros((self.C, self.P, self.ld), dtype=np.int8)
        self.D[:, :, :n] = Ds
        self.channel, self.cutoff = channel, cutoff
        mf = 4 * self.C * self.B
        self.max_finds = mf
        self.find_cj = np.zeros(mf, dtype=np.int64)
        self.find_k = np.zeros((mf, self.ld), dtype=np.int8)
        self.find_rho = np.zeros(mf, dtype=np.complex64)
        self.find_round = np.zeros(mf, dtype=np.int32)
        self.find_id = np.full((self.C, self.B), -7, dtype=np.int32)
        self.counters = np.zeros(8, dtype=np.uint64)
        self.seen0 = np.zeros(self.B, dtype=np.int32)
        self.uk = np.zeros((mf, self.ld), dtype=np.int8)
        self.usum = np.zeros(mf, dtype=np.complex64)
        self.ucnt = np.zeros(mf, dtype=np.int32)
        self.ukey = np.zeros(mf, dtype=np.int64)
        self.unext = np.zeros(mf, dtype=np.int32)

    def desc(self):
        if self.rs is not None:
            return [self.q, self.n, self.b, self.C, self.P, self.P_src, self.channel, 1, self.rs.t, self.rs.s, self.ld,
                    C.c_float(self.cutoff), _p(self.MT), _p(self.D), _p(self.rs_exp), _p(self.rs_log)]
        return [self.q, self.n, self.b, self.C, self.P, self.P_src, self.channel, 0, 0, 0, self.ld, C.c_float(self.cutoff),
                _p(self.MT), _p(self.D), None, None]

    def classify(self, L, U, impl, ranges=None):
        self.counters[:] = 0
        self.find_id[:] = -7
        for (jb, je) in (ranges or [(0, -1)]):               # bin ranges append to the same find list (sharded peel)
            assert L.emu_classify(*self.desc(), _p(U), _p(self.find_cj), _p(self.find_k), _p(self.find_rho),
                                  _p(self.find_round), _p(self.find_id), self.max_finds, 1, _p(self.counters), impl, jb, je) == 0
        nf, nm = int(self.counters[0]), int(self.counters[1])
        order = np.argsort(self.find_cj[:nf])
        return {"nf": nf, "nm": nm, "cj": self.find_cj[:nf][order].copy(), "k": self.find_k[:nf][order].copy(),
                "rho": self.find_rho[:nf][order].copy(), "fid": self.find_id.copy()}

    def peel(self, L, U, impl):
        nf, nu, nr = C.c_longlong(0), C.c_longlong(0), C.c_int(0)
        rc = L.emu_peel(*self.desc(), _p(U), _p(self.find_cj), _p(self.find_k), _p(self.find_rho), _p(self.find_round),
                        _p(self.find_id), self.max_finds, _p(self.seen0), _p(self.uk), _p(self.usum), _p(self.ucnt),
                        _p(self.ukey), _p(self.unext), self.max_finds, impl, C.byref(nf), C.byref(nu), C.byref(nr))
        assert rc == 0
        nu = nu.value
        order = np.argsort(self.ukey[:nu])
        keys = [tuple(int(v) for v in r[:self.n]) for r in self.uk[:nu][order]]
        vals = self.usum[:nu][order].astype(np.complex128) / self.ucnt[:nu][order]
        return keys, vals, nf.value, nr.value


def _problem_from_golden(g, p, nso_subtype="nso1"):
    q, n = p["q"], p["n"]
    U = np.ascontiguousarray(g["mdu_Us"].reshape(p["trC"], -1, q ** p["trb"])).astype(np.complex64)
    D = g["mdu_Ds"].reshape(p["trC"], -1, n)
    channel = 0 if p["chan"] == "identity" else (1 if nso_subtype == "nso1" else 2)
    cutoff = 1e-9 + 1.5 * p["noise_sd"] ** 2 / q ** p["trb"]
    return Problem(q, n, p["trb"], list(g["mdu_Ms"]), D, p["P_src"], channel, cutoff), U


@pytest.mark.parametrize("impl", [1, 2, 3, 4])
@pytest.mark.parametrize("name", FULL_CASES + WIDE_FULL_CASES)
def test_emulated_peel_from_reference_bins(emu, name, impl):
    """The kernels' round loop on the reference's own bins: same distinct coefficients in the same first-seen order."""
    g = load_golden(name)
    p = case_params(g)
    prob, U = _problem_from_golden(g, p)
    keys, vals, nf, nr = prob.peel(emu, U, impl)
    want = [tuple(int(v) for v in k) for k in g["res_keys"]]
    assert keys == want
    assert np.max(np.abs(vals - g["res_vals"])) <= 1e-5 * np.max(np.abs(g["res_vals"]))


@pytest.mark.parametrize("impl", [1, 2, 4])
@pytest.mark.parametrize("name", NSO2_CASES)
def test_emulated_peel_nso2(emu, name, impl):
    g = load_golden(name)
    p = case_params(g)
    prob, U = _problem_from_golden(g, p, "nso2")
    keys, vals, _, _ = prob.peel(emu, U, impl)
    assert keys == [tuple(int(v) for v in k) for k in g["res_keys"]]
    assert np.max(np.abs(vals - g["res_vals"])) <= 1e-5 * np.max(np.abs(g["res_vals"]))


def test_emulated_detectors(emu):
    g = load_golden("detect_units")
    g2 = load_golden("detect_units2")
    for gg, tag, channel in [(g, "nl_q4", 0), (g, "nl_q3", 0), (g, "nso_q4", 1), (g, "nso_q5", 1), (g, "nso_q2", 1),
                             (g2, "nso2_q4", 2), (g2, "nso2_q3", 2), (g2, "nso2_q5", 2), (g2, "nso2_q2", 2), (g2, "nso2_q7", 2)]:
        q, p1, R = (int(v) for v in gg[tag + "_meta"])
        cols = np.ascontiguousarray(gg[tag + "_cols"].astype(np.complex64))
        out = np.full((len(cols), 16), 99, dtype=np.int8)
        assert emu.emu_detect(_p(cols), len(cols), q, 0, p1 * R, p1, channel, 0, 0, 0, None, None, _p(out), 16) == 0
        assert np.array_equal(out[:, :p1 - 1], gg[tag + "_k"]) and not out[:, p1 - 1:].any(), tag
    for tag in ["mle_q2", "mle_q3", "mle_q4"]:
        cols = np.ascontiguousarray(g2[tag + "_cols"].astype(np.complex64))
        S = np.ascontiguousarray(g2[tag + "_S"].astype(np.complex64))
        ksel = np.zeros(len(cols), dtype=np.int32)
        res = np.zeros(len(cols), dtype=np.float32)
        assert emu.emu_mle(_p(cols), len(cols), S.shape[0], _p(S), S.shape[1], _p(ksel), _p(res)) == 0
        assert np.array_equal(g2[tag + "_selection"][ksel], g2[tag + "_ksel"]), tag


def test_emulated_detect_coded(emu):
    """Channel stage + Reed-Solomon decode inside the detection kernel == oracle decode (incl. decoder failures)."""
    import qsft_oracle as orc
    from qsft_b200.reed_solomon import ReedSolomon
    n, t, q, R = 12, 2, 3, 2
    rs = ReedSolomon(n, t, q)
    D = rs.get_delay_matrix()
    odec = orc.get_reed_solomon_dec(n, t, q)
    p1 = D.shape[0]
    e, l = rs.device_tables()
    e, l = np.ascontiguousarray(e, dtype=np.int32), np.ascontiguousarray(l, dtype=np.int32)
    rng = np.random.default_rng(3)
    cols, want = [], []
    for i in range(40):
        k = np.zeros(n, dtype=int)
        w = int(rng.integers(0, t + 2))
        k[rng.choice(n, w, replace=False)] = rng.integers(1, q, w)
        off = rng.integers(0, q, (R, n))
        ph = np.concatenate([((off[r] - D) % q) @ k % q for r in range(R)])
        col = (1.3 - 0.4j) * np.exp(2j * np.pi * ph / q)
        cols.append(col)
        sym = orc.detect_nso1(col[:, None], q, p1)[:, 0]
        want.append(np.array(odec(list(sym))[0][0, :], dtype=int))
    cols = np.ascontiguousarray(np.array(cols).astype(np.complex64))
    out = np.full((len(cols), 16), 99, dtype=np.int8)
    assert emu.emu_detect(_p(cols), len(cols), q, n, p1 * R, p1, 1, 1, t, rs.s, _p(e), _p(l), _p(out), 16) == 0
    assert np.array_equal(out[:, :n], np.array(want)) and not out[:, n:].any()


def test_emulated_quadrant_detection_equals_exact(emu, monkeypatch):
    """Opt-in QSFT_K4_FASTDET=1 (q = 2 / 4 symbols by quadrant comparison) decides exactly like the default path, also on
    columns placed on and next to the decision boundaries."""
    rng = np.random.default_rng(17)
    for q, channel, p1, R in [(4, 0, 9, 1), (4, 1, 9, 1), (4, 1, 7, 3), (2, 0, 11, 1), (2, 1, 11, 2), (4, 2, 9, 2)]:
        P, N = p1 * R, 4000
        ph = rng.integers(0, q, (N, P)) + rng.choice([0.0, 0.0, 0.5, 0.49, 0.51, 0.499999, 0.25], (N, P)) \
            + rng.normal(0, 0.02, (N, P)) * rng.integers(0, 2, (N, 1))
        amp = rng.uniform(0.2, 3, (N, 1)) * np.exp(1j * rng.uniform(0, 2 * np.pi, (N, 1)))
        cols = amp * np.exp(2j * np.pi * ph / q)
        cols[::97] = 0                                       # vanishing columns
        cols[5::131, 0] = 0                                  # vanishing reference delay
        cols = np.ascontiguousarray(cols.astype(np.complex64))
        outs = []
        for fast in ("0", "1"):
            monkeypatch.setenv("QSFT_K4_FASTDET", fast)
            out = np.full((N, 16), 99, dtype=np.int8)
            assert emu.emu_detect(_p(cols), N, q, 0, P, p1, channel, 0, 0, 0, None, None, _p(out), 16) == 0
            outs.append(out)
        assert np.array_equal(outs[0], outs[1]), (q, channel)


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("name", ["cfg1_q4_n10_b4_identity", "cfg2r_q4_n14_b5_nso_noisy", "q2_n12_b4_simple",
                                  "q4_n10_allbs_subselect"])
def test_emulated_peel_with_quadrant_detection(emu, name, impl, monkeypatch):
    monkeypatch.setenv("QSFT_K4_FASTDET", "1")
    g = load_golden(name)
    p = case_params(g)
    prob, U = _problem_from_golden(g, p)
    keys, vals, _, _ = prob.peel(emu, U, impl)
    assert keys == [tuple(int(v) for v in k) for k in g["res_keys"]]
    assert np.max(np.abs(vals - g["res_vals"])) <= 1e-5 * np.max(np.abs(g["res_vals"]))


def test_emulated_peel_coded_source_vs_oracle(emu):
    """Reed-Solomon source decoding inside the classification kernel (config-3 shape, reduced): the emulated peel loop
    on the oracle's bins == the oracle's transform with its own decoder, same order."""
    import qsft_oracle as orc
    from qsft_b200.reed_solomon import ReedSolomon
    np.random.seed(4)
    n, q, S, b, Cn, t, R = 14, 3, 30, 3, 3, 3, 2
    qa = {"query_method": "complex", "num_subsample": Cn, "delays_method_source": "coded", "subsampling_method": "qsft",
          "delays_method_channel": "nso", "num_repeat": R, "b": b, "t": t}
    signal_w, locq, strengths = orc.generate_signal_w(n, q, S, 1, 1, max_weight=t)
    sig = orc.OracleSignal(n, q, qa, locq, strengths, noise_sd=0.0, signal_w=signal_w)
    Ms, Ds, Us = sig.get_MDU(Cn, R, b)

    class Fixed:
        q, n, noise_sd = sig.q, sig.n, 0.0

        def get_source_parity(self):
            return sig.get_source_parity()

        def get_MDU(self, *a, **k):
            return Ms, Ds, [[np.array(u) for u in us] for us in Us], None

    want = orc.transform(Fixed(), Cn, R, b, "coded", "nso", source_decoder=orc.get_reed_solomon_dec(n, t, q))
    U = np.ascontiguousarray(np.array([np.vstack(us) for us in Us]).astype(np.complex64))
    D = np.array([np.vstack(d) for d in Ds])
    prob = Problem(q, n, b, Ms, D, sig.get_source_parity(), 1, 1e-9, rs=ReedSolomon(n, t, q))
    for impl in (1, 2, 4):
        keys, vals, _, _ = prob.peel(emu, U.copy(), impl)
        assert keys == list(want.keys()) and len(keys) >= 0.8 * len(signal_w)
    assert np.max(np.abs(vals - np.array(list(want.values())))) <= 1e-5


@pytest.mark.parametrize("impl", [1])
def test_emulated_classify_bin_ranges(emu, impl):
    """Bin-sharded classification (multi-GPU peel): ragged ranges that do not align with the CTA tiles give the same finds
    and the same find table as one pass over all bins."""
    g = load_golden("cfg2r_q4_n14_b5_nso_noisy")
    p = case_params(g)
    prob, U = _problem_from_golden(g, p)
    full = prob.classify(emu, U, impl)
    B = prob.B
    parts = prob.classify(emu, U, impl, ranges=[(0, 37), (37, 37), (37, 700), (700, B)])
    assert parts["nf"] == full["nf"] and parts["nm"] == full["nm"]
    assert np.array_equal(parts["cj"], full["cj"]) and np.array_equal(parts["k"], full["k"])
    assert np.array_equal(parts["fid"] >= 0, full["fid"] >= 0) and not (parts["fid"] == -7).any()


@pytest.mark.parametrize("q,n,b,Cn", [(2, 3, 3, 4), (2, 4, 2, 3), (3, 2, 2, 2), (2, 5, 4, 4)])
def test_emulated_peel_tiny_alphabet_guard_vs_oracle(emu, q, n, b, Cn):
    """q^n <= 15 C B: the loop of qsft.py:151 can stop on `num_peeling < q^n` before the round limit (a fuzz of
    the emulated host-driven loop had diverged there).  Host-driven rounds (impl 1) and the on-device loop kernel (impl 2)
    on the oracle's bins == the oracle's transform: same keys in the same order, same number of rounds."""
    import qsft_oracle as orc
    for seed in range(8):
        np.random.seed(300 + seed)
        S = max(1, min(q ** n // 2, 2 + seed % 4))
        qa = {"query_method": "complex", "num_subsample": Cn, "delays_method_source": "identity", "subsampling_method": "qsft",
              "delays_method_channel": "identity", "num_repeat": 1, "b": b}
        signal_w, locq, strengths = orc.generate_signal_w(n, q, S, 1, 1)
        sig = orc.OracleSignal(n, q, qa, locq, strengths, noise_sd=0.0, signal_w=signal_w)
        Ms, Ds, Us = sig.get_MDU(Cn, 1, b)

        class Fixed:
            noise_sd = 0.0

            def __init__(self):
                self.q, self.n = q, n

            def get_source_parity(self):
                return sig.get_source_parity()

            def get_MDU(self, *a, **k):
                return Ms, Ds, [[np.array(u) for u in us] for us in Us], None

        want = orc.transform(Fixed(), Cn, 1, b, "identity", "identity", report=True)
        U = np.ascontiguousarray(np.array([np.vstack(us) for us in Us]).astype(np.complex64))
        D = np.array([np.vstack(d) for d in Ds])
        prob = Problem(q, n, b, Ms, D, sig.get_source_parity(), 0, 1e-9)
        for impl in (1, 2):
            keys, vals, _, rounds = prob.peel(emu, U.copy(), impl)
            assert keys == list(want["gwht"].keys()), (seed, impl)
            assert rounds == want["rounds"], (seed, impl, rounds, want["rounds"])
            if keys:
                assert np.max(np.abs(vals - np.array(list(want["gwht"].values())))) <= 1e-5, (seed, impl)
