"""CPU: pin the oracle restatement (oracle/qsft_oracle.py) to fixtures produced by the unmodified reference
(oracle/gen_golden.py).  Bit-exact for integer work, 1e-12 for complex128 values."""
import numpy as np
import pytest

import qsft_oracle as orc
from conftest import FULL_CASES, WIDE_FULL_CASES, INDEX_CASES, NSO2_CASES, case_params, load_golden, u128_to_ints


def build(g):
    p = case_params(g)
    np.random.seed(p["seed"])
    signal_w, locq, strengths = orc.generate_signal_w(p["n"], p["q"], p["S"], 1, 1, max_weight=p["max_weight"])
    sig = orc.OracleSignal(p["n"], p["q"], p["query_args"], locq, strengths, noise_sd=p["noise_sd"],
                           signal_w=signal_w)
    return p, sig


@pytest.mark.parametrize("name", FULL_CASES + WIDE_FULL_CASES)
def test_full_pipeline_matches_reference(name):
    g = load_golden(name)
    p, sig = build(g)
    # RNG-parity-critical setup
    assert np.array_equal(sig.locq, g["locq"])
    assert np.allclose(sig.strengths, g["strengths"], rtol=0, atol=0)
    assert np.array_equal(np.array(sig.Ms), g["Ms"])
    assert np.array_equal(np.array(sig.Ds[0]), g["Ds"])
    # query indices and samples of group (0, 0)
    idx = orc.query_indices(sig.Ms[0], sig.Ds[0][0], p["q"])
    want = u128_to_ints(g["idx00_hi"], g["idx00_lo"])
    assert [int(v) for row in idx for v in row] == want
    assert np.allclose(sig.subsample(idx[1]), g["samples00_row1"], rtol=1e-12, atol=1e-12)
    # transforms at every b
    for bb in sig.all_bs:
        mine = np.array([[sig.Us[i][j][bb] for j in range(p["R"])] for i in range(p["C"])])
        assert np.allclose(mine, g[f"Us_b{bb}"], rtol=1e-12, atol=1e-13)
    # get_MDU (selection order + noise) and the peeling result
    state = np.random.get_state()
    Ms_r, Ds_r, Us_r = sig.get_MDU(p["trC"], p["trR"], p["trb"])
    assert np.array_equal(np.array(Ms_r), g["mdu_Ms"])
    assert np.array_equal(np.array(Ds_r), g["mdu_Ds"])
    assert np.allclose(np.array(Us_r), g["mdu_Us"], rtol=1e-12, atol=1e-13)
    np.random.set_state(state)
    res = orc.transform(sig, p["trC"], p["trR"], p["trb"], reconstruct_method_source=p["src"],
                        reconstruct_method_channel=p["chan"], report=True, sort=True)
    assert np.random.random() == float(g["rng_probe"])
    keys = [tuple(int(v) for v in k) for k in g["res_keys"]]
    assert list(res["gwht"].keys()) == keys                      # same finds in the same first-seen order
    assert np.allclose(np.array(list(res["gwht"].values())), g["res_vals"], rtol=1e-10, atol=1e-12)
    assert res["n_samples"] == int(g["n_samples"])
    assert np.array_equal(np.array(res["locations"]), g["locations"])
    assert res["max_hamming_weight"] == int(g["max_hw"])
    assert abs(res["avg_hamming_weight"] - float(g["avg_hw"])) < 1e-12


@pytest.mark.parametrize("name", FULL_CASES + WIDE_FULL_CASES)
def test_closed_form_bins_identity(name):
    g = load_golden(name)
    p = case_params(g)
    for c in range(p["C"]):
        for r in range(p["R"]):
            U = orc.closed_form_bins(g["Ms"][c], g["Ds"][r], g["locq"], g["strengths"], p["q"])
            assert np.allclose(U, g[f"Us_b{p['b']}"][c, r], atol=1e-12)


@pytest.mark.parametrize("name", INDEX_CASES)
def test_wide_indices(name):
    g = load_golden(name)
    q, n, b, P = (int(v) for v in g["meta"])
    idx = orc.query_indices(g["M"], g["D"], q)
    got = [int(v) for row in idx for v in row]
    assert got == u128_to_ints(g["hi"], g["lo"])
    assert np.array_equal(orc.dec_to_qary_vec(got[:64], q, n), g["digits64"])
    # codec round trip
    dig = orc.dec_to_qary_vec(got, q, n)
    assert [int(v) for v in orc.qary_vec_to_dec(dig, q)] == got


def test_detection_units():
    g = load_golden("detect_units")
    for tag in ["nl_q4", "nl_q3", "nso_q4", "nso_q5", "nso_q2"]:
        q, p1, R = (int(v) for v in g[tag + "_meta"])
        cols = g[tag + "_cols"].T
        if tag.startswith("nl"):
            k = orc.detect_noiseless(cols, q)
        else:
            k = orc.detect_nso1(cols, q, p1)
        assert np.array_equal(k.T, g[tag + "_k"]), tag


def test_detection_units_nso2_mle():
    """The detectors QSFT.transform never selects, pinned at function level."""
    g = load_golden("detect_units2")
    for tag in ["nso2_q4", "nso2_q3", "nso2_q5", "nso2_q2", "nso2_q7"]:
        q, p1, R = (int(v) for v in g[tag + "_meta"])
        k = orc.detect_nso2(g[tag + "_cols"].T, q, p1)
        assert np.array_equal(k.T, g[tag + "_k"]), tag
    for tag in ["mle_q2", "mle_q3", "mle_q4"]:
        sel, S = g[tag + "_selection"], g[tag + "_S"]
        for col, want_k, want_sig in zip(g[tag + "_cols"], g[tag + "_ksel"], g[tag + "_sig"]):
            k, sig, idx = orc.detect_mle(col, sel, S)
            assert int(k) == int(want_k) and np.array_equal(sig, want_sig) and sel[idx] == k, tag


@pytest.mark.parametrize("name", NSO2_CASES)
def test_nso2_pipeline_matches_reference(name):
    g = load_golden(name)
    assert str(g["nso_subtype"]) == "nso2"
    p, sig = build(g)
    assert np.array_equal(np.array(sig.Ms), g["Ms"]) and np.array_equal(np.array(sig.Ds[0]), g["Ds"])
    mine = np.array([[sig.Us[i][j][p["b"]] for j in range(p["R"])] for i in range(p["C"])])
    assert np.allclose(mine, g[f"Us_b{p['b']}"], rtol=1e-12, atol=1e-13)
    res = orc.transform(sig, p["trC"], p["trR"], p["trb"], reconstruct_method_source=p["src"],
                        reconstruct_method_channel=p["chan"], report=True, sort=True, nso_subtype="nso2")
    assert np.random.random() == float(g["rng_probe"])
    assert list(res["gwht"].keys()) == [tuple(int(v) for v in k) for k in g["res_keys"]]
    assert np.allclose(np.array(list(res["gwht"].values())), g["res_vals"], rtol=1e-10, atol=1e-12)
    assert np.array_equal(np.array(res["locations"]), g["locations"])


def test_gwht_units():
    g = load_golden("gwht_units")
    for key in g.files:
        if key.startswith("x_"):
            _, qs, bs = key.split("_")
            q, b = int(qs[1:]), int(bs[1:])
            assert np.allclose(orc.gwht(g[key], q, b), g["y" + key[1:]], atol=1e-13)


def test_qary_ints_order():
    import itertools
    for q, m in [(2, 5), (3, 3), (4, 3), (5, 2)]:
        want = np.array(list(itertools.product(np.arange(q), repeat=m))).T
        assert np.array_equal(orc.qary_ints(m, q), want)


def test_get_Ms_simple_and_errors():
    Ms = orc.get_Ms(12, 4, 2, num_to_get=3, method="simple")
    assert np.array_equal(Ms[0][8:12], np.eye(4)) and Ms[0][:8].sum() == 0
    assert np.array_equal(Ms[2][0:4], np.eye(4))
    with pytest.raises(ValueError):
        orc.get_Ms(12, 4, 2, num_to_get=4, method="simple")
    with pytest.raises(NotImplementedError):
        orc.get_D(6, 3, "identity", "coded")


# ---- Reed-Solomon restatement: PARITY UNPINNED (galois absent) -> self-consistency only -----------------
@pytest.mark.parametrize("n,t,q", [(30, 4, 3), (10, 2, 3), (20, 3, 5), (12, 2, 2), (7, 1, 7)])
def test_rs_self_consistency(n, t, q):
    rs = orc.RSCode(n, t, q)
    D = rs.get_delay_matrix()
    assert D.shape == (2 * t * rs.s + 1, n) and not D[0].any()
    rng = np.random.default_rng(n * 100 + t)
    for _ in range(100):
        w = int(rng.integers(0, t + 1))
        k = np.zeros(n, dtype=int)
        pos = rng.choice(n, w, replace=False)
        k[pos] = rng.integers(1, q, w)
        dec, ne = rs.syndrome_decode(list((D[1:] @ k) % q))
        assert np.array_equal(dec[0], k) and ne == w


def test_rs_end_to_end_coded_transform():
    """Config-3 shaped (reduced): q=3 coded delays t=3, low-weight support, noiseless -> exact recovery."""
    np.random.seed(4)
    n, q, S, b, C, t = 14, 3, 30, 3, 3, 3
    qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "coded", "subsampling_method": "qsft",
          "delays_method_channel": "identity", "num_repeat": 1, "b": b, "t": t}
    signal_w, locq, strengths = orc.generate_signal_w(n, q, S, 1, 1, max_weight=t)
    sig = orc.OracleSignal(n, q, qa, locq, strengths, noise_sd=0.0, signal_w=signal_w)
    res = orc.transform(sig, C, 1, b, "coded", "identity", source_decoder=orc.get_reed_solomon_dec(n, t, q))
    assert set(res.keys()) == set(signal_w.keys())
    assert orc.nmse(res, signal_w) < 1e-20
