"""CPU tests of the product's host logic and of the C-ABI surface (no compute calls: there is no GPU here)."""
import os
import re

import numpy as np
import pytest

from conftest import FULL_CASES, WIDE_FULL_CASES, INDEX_CASES, ROOT, case_params, load_golden, u128_to_ints


def test_library_loads_and_exports_every_declared_symbol():
    import qsft_b200
    from qsft_b200 import _lib
    L = qsft_b200.lib()
    header = open(os.path.join(ROOT, "include", "qsft_b200.h")).read()
    declared = set(re.findall(r"\b(qsft_[a-z0-9_]+)\s*\(", header))
    declared.discard("qsft_peel_desc")
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/qsft_b200.h but not exported"
    assert set(_lib.EXPORTS) == declared
    assert L.qsft_version() >= 100


def test_ctypes_bindings_match_the_header_prototypes():
    """Every prototype of include/qsft_b200.h is bound with the same number of arguments in qsft_b200/_lib.py (a binding that
    drifts from the header passes garbage in registers without any error)."""
    import qsft_b200
    L = qsft_b200.lib()
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "qsft_b200.h")).read(), flags=re.S)
    protos = re.findall(r"\b(?:int|int64_t|void|const char\*)\s+(qsft_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)
    assert len(protos) >= 25
    for name, args in protos:
        args = args.strip()
        n_args = 0 if args in ("", "void") else args.count(",") + 1
        bound = getattr(L, name).argtypes
        if n_args == 0:
            assert not bound, name
        else:
            assert bound is not None and len(bound) == n_args, (name, n_args, None if bound is None else len(bound))


def test_shard_rows_cover_the_table_exactly_once():
    """Row blocks of the rank-shared support staging (ops.pad_digits_sharded): equal block size, every row in exactly one block,
    ragged and empty tails."""
    from qsft_b200.ops import shard_rows
    for N in (0, 1, 5, 8, 60, 1000, 100_000, 100_003):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for rank in range(world):
                per, lo, hi = shard_rows(N, world, rank)
                assert 0 <= lo <= hi <= N and hi - lo <= per and per * world >= N
                assert lo == min(N, rank * per)
                seen.extend(range(lo, hi))
            assert seen == list(range(N)), (N, world)


def test_no_cpu_fallback_without_cuda():
    import torch
    import qsft_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    qa = {"query_method": "complex", "num_subsample": 2, "delays_method_source": "identity",
          "subsampling_method": "qsft", "delays_method_channel": "identity", "num_repeat": 1, "b": 2}
    with pytest.raises(RuntimeError):
        qsft_b200.get_random_subsampled_signal(n=4, q=2, noise_sd=0, sparsity=3, a_min=1, a_max=1, query_args=qa)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "qsft_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "qsft_oracle" not in src and "import oracle" not in src and "/root/reference" not in src, f


@pytest.mark.parametrize("name", FULL_CASES + WIDE_FULL_CASES)
def test_host_rng_order_matches_reference(name):
    """generate_signal_w -> get_Ms_and_Ds consume np.random like the reference (same seed, same arrays)."""
    from qsft_b200.synthetic_signal import generate_signal_w
    from qsft_b200.query import get_Ms_and_Ds
    g = load_golden(name)
    p = case_params(g)
    np.random.seed(p["seed"])
    signal_w, locq, strengths = generate_signal_w(p["n"], p["q"], p["S"], 1, 1, p["noise_sd"], full=False,
                                                  max_weight=p["max_weight"])
    Ms, Ds = get_Ms_and_Ds(p["n"], p["q"], **p["query_args"])
    assert np.array_equal(locq, g["locq"])
    assert np.array_equal(strengths, g["strengths"])
    assert np.array_equal(np.array(Ms), g["Ms"])
    assert np.array_equal(np.array(Ds[0]), g["Ds"])
    assert all(d is Ds[0] for d in Ds)


@pytest.mark.parametrize("name", INDEX_CASES)
def test_host_codecs(name):
    from qsft_b200 import utils
    g = load_golden(name)
    q, n, b, P = (int(v) for v in g["meta"])
    ints = u128_to_ints(g["hi"], g["lo"])
    limbs = utils.index_limbs(q, n)
    arr = utils.ints_to_limbs(ints, limbs)
    assert [int(v) for v in utils.limbs_to_ints(arr)] == ints
    dig = utils.dec_to_qary_vec(ints[:64], q, n)
    assert np.array_equal(dig, g["digits64"])
    assert [int(v) for v in utils.qary_vec_to_dec(dig, q)] == ints[:64]


def test_index_limbs_and_ld():
    from qsft_b200 import utils
    assert utils.index_limbs(4, 32) == 1 and utils.index_limbs(4, 33) == 2 and utils.index_limbs(4, 64) == 2
    assert utils.index_limbs(2, 128) == 2
    with pytest.raises(ValueError):
        utils.index_limbs(4, 65)
    assert utils.padded_ld(10) == 32 and utils.padded_ld(40) == 64 and utils.padded_ld(64) == 64 and utils.padded_ld(100) == 128


def test_query_errors_like_reference():
    from qsft_b200 import query
    with pytest.raises(ValueError):
        query.get_Ms(12, 4, 2, num_to_get=4, method="simple")
    with pytest.raises(NotImplementedError):
        query.get_D(6, q=3, delays_method_source="identity", delays_method_channel="coded")
    with pytest.raises(NotImplementedError):
        query.get_reed_solomon_dec(10, 2, 4)
    Ms = query.get_Ms(12, 4, 2, num_to_get=3, method="simple")
    assert np.array_equal(Ms[0][8:12], np.eye(4)) and np.array_equal(Ms[2][0:4], np.eye(4))


def test_reed_solomon_host_matches_oracle():
    import qsft_oracle as orc
    from qsft_b200.reed_solomon import ReedSolomon
    for (n, t, q) in [(30, 4, 3), (20, 3, 5), (12, 2, 2), (100, 2, 2)]:       # (100, 2, 2): GF(2^7)
        a, o = ReedSolomon(n, t, q), orc.RSCode(n, t, q)
        D = a.get_delay_matrix()
        assert np.array_equal(D, o.get_delay_matrix())
        rng = np.random.default_rng(n)
        for it in range(100):
            if it % 2:
                syn = rng.integers(0, q, 2 * t * a.s)
            else:
                k = np.zeros(n, dtype=int)
                w = int(rng.integers(0, t + 1))
                k[rng.choice(n, w, replace=False)] = rng.integers(1, q, w)
                syn = (D[1:] @ k) % q
            r1, r2 = a.syndrome_decode(list(syn)), o.syndrome_decode(list(syn))
            assert np.array_equal(r1[0], r2[0]) and r1[1] == r2[1]


def test_field_polynomials_are_matlab_defaults():
    """galois 0.1.x builds its RS fields from matlab_primitive_poly(p, m).  For p = 2 these are MATLAB's published
    gf() defaults (decimal): the lexicographically-minimal primitive polynomial except m = 7.  Irreducibility of every
    polynomial used is cross-checked with sympy."""
    import qsft_oracle as orc
    from qsft_b200.reed_solomon import GaloisField
    from sympy.polys.domains import ZZ
    from sympy.polys.galoistools import gf_irreducible_p
    matlab = {2: 7, 3: 11, 4: 19, 5: 37, 6: 67, 7: 137, 8: 285, 9: 529, 10: 1033}
    for m, dec in matlab.items():
        assert (1 << m) + GaloisField(2, m).poly_low == dec
        assert (1 << m) + orc.GFext(2, m).poly_low == dec
    for p, m in [(3, 2), (3, 3), (3, 4), (5, 2), (5, 3), (7, 2), (11, 2), (13, 2)]:
        h, o = GaloisField(p, m), orc.GFext(p, m)
        assert h.poly_low == o.poly_low
        coeffs = [1] + [(h.poly_low // p ** i) % p for i in range(m - 1, -1, -1)]       # monic, high degree first
        assert gf_irreducible_p([ZZ(c) for c in coeffs], p, ZZ)
        # primitive: x generates all p^m - 1 non-zero elements
        assert len(set(int(v) for v in h.exp[: p ** m - 1])) == p ** m - 1
        # antilog table == x^e mod f computed independently by sympy's GF(p)[x] arithmetic
        from sympy.polys.galoistools import gf_pow_mod
        f = [ZZ(c) for c in coeffs]
        for e in list(range(0, min(p ** m - 1, 40))) + [p ** m - 2]:
            r = gf_pow_mod([ZZ(1), ZZ(0)], e, f, p, ZZ) or [ZZ(0)]
            val = 0
            for cdig in r:                                  # high degree first -> base-p integer
                val = val * p + int(cdig) % p
            assert int(h.exp[e]) == val == int(o.exp[e]), (p, m, e)


def test_finds_to_dict_matches_reference_averaging():
    from qsft_b200.qsft import QSFT
    # two rounds; k A found twice in round 1 (groups 0 and 2) and once in round 2, k B once
    kA, kB = [1, 0, 2], [0, 3, 3]
    cj = np.array([5, 40, 17, 9])
    k = np.array([kA, kA, kB, kA], dtype=np.int8)
    rho = np.array([1 + 1j, 3 + 1j, 2j, 5 + 4j], dtype=np.complex64)
    rnd = np.array([1, 1, 1, 2], dtype=np.int32)
    perm = np.array([3, 1, 0, 2])     # device order is arbitrary
    gw, keys = QSFT._finds_to_dict(cj[perm], k[perm], rho[perm], rnd[perm])
    assert list(gw.keys()) == [tuple(kA), tuple(kB)]
    assert abs(gw[tuple(kA)] - (1 + 1j + 3 + 1j + 5 + 4j) / 3) < 1e-6 and abs(gw[tuple(kB)] - 2j) < 1e-6
    assert keys.tolist() == [kA, kB]
    gw0, keys0 = QSFT._finds_to_dict(np.zeros(0, np.int64), np.zeros((0, 3), np.int8), np.zeros(0, np.complex64), np.zeros(0, np.int32))
    assert gw0 == {} and len(keys0) == 0


def test_ctypes_struct_layout_matches_header(tmp_path):
    """The ctypes mirrors (PeelDesc, Uniq) must have the C layout of include/qsft_b200.h (checked with gcc)."""
    import ctypes
    import subprocess
    from qsft_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "qsft_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(qsft_peel_desc), offsetof(qsft_peel_desc, ld),'
                   ' offsetof(qsft_peel_desc, cutoff), offsetof(qsft_peel_desc, MT), offsetof(qsft_peel_desc, rs_log),'
                   ' sizeof(qsft_uniq), offsetof(qsft_uniq, max_uniq));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    P, U = _lib.PeelDesc, _lib.Uniq
    want = [ctypes.sizeof(P), P.ld.offset, P.cutoff.offset, P.MT.offset, P.rs_log.offset, ctypes.sizeof(U), U.max_uniq.offset]
    assert got == want


def test_k3_twopass_ticket_order_is_deadlock_free():
    """The single-launch two-pass DFT hands tiles out by ticket; a strided tile spins until every contiguous tile of its
    block is written.  For every lag the order must be a bijection onto the tiles and every dependency must carry a
    LOWER ticket (lower tickets are resident or finished => the spin always ends).  Runs the library's own decode."""
    import ctypes as C
    from qsft_b200 import _lib
    L = _lib.lib()
    blk, tile, strided = C.c_int64(), C.c_int(), C.c_int()
    for nb, t1, t2 in [(1, 4, 4), (5, 4, 4), (7, 3, 5), (41, 16, 16), (3, 256, 256), (2, 1, 1)]:
        for lag in [0, 1, 2, 3, nb - 1, nb, nb + 5]:
            if lag < 0:
                continue
            seen = {}
            c_last = {}                                      # block -> highest ticket among its contiguous tiles
            s_first = {}                                     # block -> lowest ticket among its strided tiles
            for tk in range(nb * (t1 + t2)):
                _lib.check(L.qsft_k3_ticket_decode(tk, nb, t1, t2, lag, C.byref(blk), C.byref(tile), C.byref(strided)))
                key = (blk.value, tile.value, strided.value)
                assert key not in seen and 0 <= blk.value < nb and 0 <= tile.value < (t2 if strided.value else t1)
                seen[key] = tk
                if strided.value:
                    s_first.setdefault(blk.value, tk)
                else:
                    c_last[blk.value] = tk
            assert len(seen) == nb * (t1 + t2)
            for k in range(nb):
                assert c_last[k] < s_first[k], (nb, t1, t2, lag, k)
                if 0 < lag < nb and k + lag < nb:            # tickets between a block's two passes (full distance once k >= lag)
                    assert s_first[k] - c_last[k] - 1 == lag * t1 + min(k, lag) * t2
    assert L.qsft_k3_ticket_decode(8, 1, 4, 4, 0, C.byref(blk), C.byref(tile), C.byref(strided)) != 0   # out of range


def test_host_pack_digits_matches_numpy_cast():
    """The library's host packer (qsft_host_pack_digits, what ops.pad_digits stages the support with): narrowing cast,
    either orientation / any stride, zero padding -- against ndarray.astype(int8); a large table exercises the threads."""
    import ctypes as C
    from qsft_b200 import _lib
    L = _lib.lib()

    def pack(arr, ld, threads):
        N, n = arr.shape
        out = np.full((N, ld), 77, dtype=np.int8)
        isz = arr.dtype.itemsize
        rc = L.qsft_host_pack_digits(C.c_void_p(arr.ctypes.data), isz, N, n, arr.strides[0] // isz, arr.strides[1] // isz,
                                     C.c_void_p(out.ctypes.data), ld, threads)
        assert rc == 0, L.qsft_last_error()
        return out

    rng = np.random.default_rng(0)
    for dt in [np.int64, np.int32, np.int16, np.int8, np.uint8, np.uint16, np.uint32, np.uint64]:
        a = rng.integers(0, 100, (7, 13)).astype(dt)
        for arr in [a, a.T, a[:, ::2], a[::-1], np.asfortranarray(a)]:
            got = pack(arr, 16, 3)
            assert np.array_equal(got[:, :arr.shape[1]], arr.astype(np.int8)) and not got[:, arr.shape[1]:].any(), dt
    big = rng.integers(0, 4, (40, 50_001))                                # locq as the reference keeps it: (n, S) int64
    for threads in (1, 4, 16):
        got = pack(big.T, 64, threads)
        assert np.array_equal(got[:, :40], big.T.astype(np.int8)) and not got[:, 40:].any()
    assert pack(np.zeros((0, 4), dtype=np.int64), 16, 2).shape == (0, 16)
    a = np.zeros((2, 20), dtype=np.int64)
    assert L.qsft_host_pack_digits(C.c_void_p(a.ctypes.data), 8, 2, 20, 20, 1, C.c_void_p(a.ctypes.data), 16, 1) == -1   # ld < n
    assert L.qsft_host_pack_digits(C.c_void_p(a.ctypes.data), 3, 2, 20, 20, 1, C.c_void_p(a.ctypes.data), 32, 1) == -1   # elem_bytes

def test_c_abi_rejects_bad_arguments_without_touching_the_gpu():
    """Every entry point validates its arguments before the first CUDA call: QSFT_EINVAL (-1) + a message, on any host."""
    import ctypes as C
    from qsft_b200 import _lib
    L = _lib.lib()
    null = C.c_void_p(0)
    one = C.c_void_p(16)                                  # never dereferenced: validation fails first
    EINVAL = -1

    def bad(rc, needle):
        assert rc == EINVAL
        assert needle in L.qsft_last_error().decode(), L.qsft_last_error()

    bad(L.qsft_query_lattice(one, one, 1, 10, 4, 3, one, 1, null, 0, null), "q=1")
    bad(L.qsft_query_lattice(one, one, 4, 200, 4, 3, one, 2, null, 0, null), "n=200")
    bad(L.qsft_query_lattice(one, one, 4, 40, 4, 3, one, 1, null, 0, null), "does not fit")      # 80 bits in one limb
    bad(L.qsft_query_lattice(one, one, 4, 10, 4, 3, null, 1, null, 0, null), "null")
    bad(L.qsft_query_lattice(one, one, 4, 10, 4, 3, null, 1, one, 24, null), "ld=24")
    bad(L.qsft_eval_synth(one, 5, one, one, 5, 4, 10, 20, one, 0, null), "ld=20")
    bad(L.qsft_eval_synth(one, 5, one, one, 5, 4, 10, 32, one, 7, null), "impl")
    bad(L.qsft_eval_synth(one, 5, null, one, 5, 4, 10, 32, one, 0, null), "null")
    bad(L.qsft_gwht_batch(one, 3, 1, 4, null), "q=1")
    bad(L.qsft_gwht_batch(one, -1, 4, 4, null), "negative")
    bad(L.qsft_gwht_batch(null, 3, 4, 4, null), "null")
    assert L.qsft_gwht_batch(null, 0, 4, 4, null) == 0    # empty batch: nothing to do, nothing touched
    assert L.qsft_eval_lattice_supported(4, 40, 10, 41, 100000) == 1
    assert L.qsft_eval_lattice_supported(3, 40, 10, 41, 100000) == 1   # q = 3: dense Z[w] variant
    assert L.qsft_eval_lattice_supported(3, 40, 6, 41, 100000) == 0    # ... needs b >= 7
    assert L.qsft_eval_lattice_supported(5, 40, 10, 41, 100000) == 1   # q = 5 / 7: generic Z[w] variant
    assert L.qsft_eval_lattice_supported(7, 40, 3, 41, 100000) == 1 and L.qsft_eval_lattice_supported(5, 40, 14, 41, 100000) == 0
    assert L.qsft_eval_lattice_supported(11, 40, 6, 41, 100000) == 0 and L.qsft_eval_lattice_supported(6, 40, 6, 41, 100000) == 0
    assert L.qsft_eval_lattice_supported(2, 40, 14, 41, 100000) == 1 and L.qsft_eval_lattice_supported(2, 40, 13, 41, 100000) == 0
    assert L.qsft_eval_lattice_supported(4, 40, 6, 41, 100000) == 0
    bad(L.qsft_eval_synth_lattice(one, one, one, one, 100, 11, 10, 8, 3, 32, one, null), "q = 4")
    desc = _lib.PeelDesc(q=4, n=10, b=4, C=3, P=11, P_src=11, channel=0, source=0, rs_t=0, rs_s=0, ld=32, cutoff=1e-9,
                         MT=16, D=16, rs_exp=0, rs_log=0)
    cnt = C.c_int64(0)

    def peel(d):
        return L.qsft_peel(C.byref(d), one, one, one, one, one, one, 10, one, None, C.byref(cnt), C.byref(cnt),
                           C.byref(C.c_int(0)), null)

    for field, value, needle in [("channel", 3, "channel"), ("source", 2, "source"), ("ld", 24, "ld=24"), ("P", 12, "multiple"),
                                 ("P_src", 10, "multiple"), ("b", 0, "b=0"), ("q", 1, "q=1"), ("MT", 0, "null")]:
        d = _lib.PeelDesc.from_buffer_copy(desc)
        setattr(d, field, value)
        bad(peel(d), needle)
    d = _lib.PeelDesc.from_buffer_copy(desc)
    d.P, d.P_src, d.channel = 22, 11, 0                   # identity channel cannot use repeats
    bad(peel(d), "num_repeat")
    d = _lib.PeelDesc.from_buffer_copy(desc)
    d.source, d.rs_t, d.rs_s = 1, 2, 2                    # coded: P_src must be 2ts + 1 = 9
    bad(peel(d), "2ts + 1")
    bad(L.qsft_singleton_detect(one, 4, 4, 10, 10, 3, 1, 0, 0, 0, null, null, one, 16, null), "multiple")
    bad(L.qsft_singleton_detect(one, 4, 4, 10, 10, 5, 0, 0, 0, 0, null, null, one, 16, null), "num_repeat")
    bad(L.qsft_singleton_detect(one, 4, 4, 10, 10, 5, 1, 0, 0, 0, null, null, one, 3, null), "ld_out")
    bad(L.qsft_detect_mle(one, 4, 0, one, 5, one, null, null), "positive")
    assert L.qsft_detect_mle(null, 0, 3, null, 5, null, null, null) == 0


def test_experiment_harness_host_logic(tmp_path):
    """TestHelper / run_tests plumbing with a stand-in signal factory (no GPU): folders, query_args per method,
    config.json, the sweep table's columns and order (qsft/test_helper.py, qsft/parallel_tests.py)."""
    import json
    from qsft_b200.parallel_tests import run_tests
    from qsft_b200.test_helper import TestHelper

    made = []

    class FakeSignal:
        def __init__(self, **kw):
            self.kw, self.noise_sd, self.signal_t = kw, kw.get("noise_sd"), {3: 1.0 + 0j, 5: 2.0 + 0j}
            made.append(self)

    class Helper(TestHelper):
        def generate_signal(self, signal_args):
            return FakeSignal(**signal_args)

        def compute_model(self, method, model_kwargs, report=False, verbosity=0):
            sig = self.train_signal if method == "qsft" else self.train_signal_coded
            sig.noise_sd = model_kwargs["noise_sd"]
            return {"gwht": {(0, 1): 1.0, (1, 1): 2.0}, "runtime": 0.5, "n_samples": model_kwargs["n_samples"],
                    "locations": [], "max_hamming_weight": 2, "avg_hamming_weight": 1.5}

        def test_model(self, method, **kwargs):
            return 0.25 * len(kwargs["beta"])

    sub = {"num_subsample": 3, "num_repeat": 2, "b": 4, "all_bs": [3, 4]}
    h = Helper(signal_args={"n": 6, "q": 3, "t": 2, "locq": None, "strengths": None}, methods=["qsft", "qsft_coded"],
               subsampling_args=sub, test_args={"n_samples": 100}, exp_dir=tmp_path)
    assert json.load(open(tmp_path / "config.json")) == {"query_args": sub}
    train, coded, test = made
    assert train.kw["folder"] == tmp_path / "train" and coded.kw["folder"] == tmp_path / "train_coded"
    assert train.kw["query_args"] == {**sub, "subsampling_method": "qsft", "query_method": "complex",
                                      "delays_method_source": "identity", "delays_method_channel": "nso"}
    assert coded.kw["query_args"]["delays_method_source"] == "coded" and coded.kw["query_args"]["t"] == 2
    assert test.kw["query_args"] == {"subsampling_method": "uniform", "n_samples": 100} and test.kw["noise_sd"] == 0
    assert (tmp_path / "test").is_dir() and sub == {"num_subsample": 3, "num_repeat": 2, "b": 4, "all_bs": [3, 4]}
    df = run_tests("qsft", h, 2, [2, 3], [1], [3, 4], [0.1], parallel=False)
    assert list(df.columns) == ["num_subsample", "num_repeat", "b", "noise_sd", "iter", "n", "q", "runtime",
                                "found_sparsity", "n_samples", "ratio_samples", "max_hamming_weight", "nmse", "method"]
    assert len(df) == 8 and list(df["b"][:4]) == [3, 3, 4, 4] and list(df["iter"][:2]) == [0, 1]
    assert df["n_samples"][0] == 2 * 3 ** 3 * 1 * 7 and df["nmse"][0] == 0.5 and df["found_sparsity"][0] == 2
    assert abs(df["ratio_samples"][0] - 2 * 27 * 7 / 3 ** 6) < 1e-12 and train.noise_sd == 0.1
    with pytest.raises(NotImplementedError):
        Helper(signal_args={"n": 6, "q": 3}, methods=["lasso"], subsampling_args=sub, test_args={}, exp_dir=tmp_path)
    with pytest.raises(NotImplementedError):
        TestHelper.compute_model(h, "gwht", {})


def test_rs_parity_check_matrix_known_answer():
    """Known answer for the Reed-Solomon conventions (field polynomial, H[j, i] = (alpha^(1+j))^(nt-1-i), degree-descending
    vectors): the first rows of galois.ReedSolomon(15, 9).H over GF(2^4) as printed in the galois documentation (quoted
    from memory -- galois itself is not installed here), recovered from the delay matrix of ReedSolomon(n=15, t=3, q=2)."""
    import qsft_oracle as orc
    from qsft_b200.reed_solomon import ReedSolomon
    want = [[9, 13, 15, 14, 7, 10, 5, 11, 12, 6, 3, 8, 4, 2, 1],
            [13, 14, 10, 11, 6, 8, 2, 9, 15, 7, 5, 12, 3, 4, 1]]
    for cls in (ReedSolomon, orc.RSCode):
        rs = cls(15, 3, 2)
        D = np.array(rs.get_delay_matrix())
        assert D.shape == (2 * 3 * 4 + 1, 15) and not D[0].any()
        H = [[int("".join(str(int(v)) for v in D[4 * j + 1:4 * j + 5, i]), 2) for i in range(15)] for j in range(2)]
        assert H == want


def test_reconstruct_host_stage_and_dispatch_errors():
    """reconstruct.singleton_detection_coded is the host-side source stage (reconstruct.py:34-51); dispatch errors are
    raised before any device work."""
    import qsft_oracle as orc
    from qsft_b200 import get_reed_solomon_dec, reconstruct
    n, t, q = 12, 2, 3
    dec, odec = get_reed_solomon_dec(n, t, q), orc.get_reed_solomon_dec(n, t, q)
    D = dec.__self__.get_delay_matrix()
    rng = np.random.default_rng(1)
    for _ in range(20):
        k = np.zeros(n, dtype=int)
        k[rng.choice(n, 2, replace=False)] = rng.integers(1, q, 2)
        sym = (D[1:] @ k) % q
        got = reconstruct.singleton_detection_coded(sym, source_decoder=dec)
        assert got.dtype == np.int32 and np.array_equal(got, k)
        assert np.array_equal(got, np.array(odec(list(sym))[0][0, :], dtype=np.int32))
    with pytest.raises(TypeError):
        reconstruct.singleton_detection(np.ones(5, dtype=complex), method_channel="bogus", q=3)
    with pytest.raises(TypeError):
        reconstruct.singleton_detection(np.ones(5, dtype=complex), method_source="bogus", method_channel="nso", q=3)
    with pytest.raises(ValueError):
        reconstruct.singleton_detection(np.ones(5, dtype=complex), method_source="coded", method_channel="nso", q=3)


def test_sparse_spectrum_behaves_like_the_reference_dict():
    """QSFT.transform's result container (qsft_b200/result.py): a dict {tuple(k): complex} filled on first use."""
    import copy
    import pickle
    from qsft_b200.result import SparseSpectrum
    rng = np.random.default_rng(5)
    loc = rng.integers(0, 4, (50, 7)).astype(np.int8)
    loc = np.unique(loc, axis=0)[::-1].copy()                # distinct keys, some fixed order
    val = rng.normal(size=len(loc)) + 1j * rng.normal(size=len(loc))
    want = {tuple(int(v) for v in k): complex(c) for k, c in zip(loc, val)}

    def fresh():
        return SparseSpectrum(loc, val)

    assert isinstance(fresh(), dict) and len(fresh()) == len(want) and bool(fresh())
    assert fresh() == want and want == fresh() and not (fresh() != want)
    assert list(fresh()) == list(want) and list(fresh().keys()) == list(want.keys())
    assert list(fresh().items()) == list(want.items()) and list(fresh().values()) == list(want.values())
    k0 = next(iter(want))
    assert fresh()[k0] == want[k0] and k0 in fresh() and (9,) * 7 not in fresh() and fresh().get((9,) * 7, 1.5) == 1.5
    assert dict(fresh()) == want and {**fresh()} == want and fresh().copy() == want
    assert pickle.loads(pickle.dumps(fresh())) == want and copy.deepcopy(fresh()) == want
    assert repr(fresh()) == repr(want)
    d = fresh()
    d[(9,) * 7] = 2j
    assert len(d) == len(want) + 1 and d.pop((9,) * 7) == 2j and d == want
    d = fresh()
    del d[k0]
    assert len(d) == len(want) - 1 and k0 not in d
    assert np.array_equal(fresh().locations, loc) and np.array_equal(fresh().coefficients, val)
    empty = SparseSpectrum(np.zeros((0, 7), dtype=np.int8), np.zeros(0, dtype=complex))
    assert len(empty) == 0 and not empty and empty == {} and list(empty.items()) == []


def test_balanced_row_shards_cover_all_rows_and_never_cost_more_than_the_uniform_split():
    """Multi-GPU row placement (input_signal_subsampled.py: _balanced_shard): contiguous, complete, no range longer than
    the uniform share (the row buffer keeps its size), and the most loaded rank never costs more than with the uniform
    split; with a proportional cost the split is the uniform one up to where the remainder goes."""
    from qsft_b200.input_signal_subsampled import SubsampledSignal

    class _Dist:
        def __init__(self, world, rank):
            self.world_size, self.rank = world, rank

    class _Waves(SubsampledSignal):
        def __init__(self, world, rank):
            self.dist = _Dist(world, rank)

        def _block_cost(self, rows):                       # wave-quantised kernel + a set-up per block touched
            return float(-(-rows * 64 // 74)) + 0.6

    def cost(sig, lo, hi, block):
        return sum(sig._block_cost(min(hi, k + block) - max(lo, k)) for k in range(0, 1000, block) if min(hi, k + block) > max(lo, k))

    for total, block, world in [(123, 41, 8), (123, 41, 4), (123, 41, 2), (63, 21, 8), (10, 5, 3), (7, 7, 8)]:
        shards = [_Waves(world, r)._balanced_shard(total, block) for r in range(world)]
        per = -(-total // world)
        assert shards[0][0] == 0 and shards[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(shards, shards[1:]))
        assert all(0 <= hi - lo <= per for lo, hi in shards)
        sig = _Waves(world, 0)
        uniform = [(min(total, r * per), min(total, (r + 1) * per)) for r in range(world)]
        assert max(cost(sig, lo, hi, block) for lo, hi in shards) <= max(cost(sig, lo, hi, block) for lo, hi in uniform)
    assert max(cost(_Waves(8, 0), lo, hi, 41) for lo, hi in [_Waves(8, r)._balanced_shard(123, 41) for r in range(8)]) == 14.6
