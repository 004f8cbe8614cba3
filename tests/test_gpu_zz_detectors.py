"""GPU parity for the stand-alone detector entry points (qsft_singleton_detect / qsft_detect_mle) and the nso2 channel
of the peel: against fixtures produced by the reference's own reconstruct.py functions (oracle/gen_golden.py
--detectors) and against the oracle.  Symbols are integers: bit-exact."""
import numpy as np
import pytest
import torch

import qsft_oracle as orc
from conftest import NSO2_CASES, WIDE_FULL_CASES, case_params, load_golden, u128_to_ints

import os

pytestmark = [pytest.mark.gpu]

if torch.cuda.is_available():
    import qsft_b200
    from qsft_b200 import ops, reconstruct, utils
    DEV = torch.device("cuda", 0)


def test_detect_units_identity_and_nso1_standalone():
    """reconstruct.singleton_detection for a batch of columns == the reference's per-column results."""
    g = load_golden("detect_units")
    for tag in ["nl_q4", "nl_q3", "nso_q4", "nso_q5", "nso_q2"]:
        q, p1, R = (int(v) for v in g[tag + "_meta"])
        chan = "identity" if tag.startswith("nl") else "nso"
        cols = g[tag + "_cols"]                                   # (N, P) complex128
        got = reconstruct.singleton_detection(cols.T, method_channel=chan, method_source="identity", q=q,
                                              source_parity=p1, nso_subtype="nso1")
        assert np.array_equal(got.T, g[tag + "_k"]), tag
        one = reconstruct.singleton_detection(cols[3], method_channel=chan, method_source="identity", q=q,
                                              source_parity=p1)
        assert one.shape == (p1 - 1,) and np.array_equal(one, g[tag + "_k"][3])


def test_detect_units_nso2():
    g = load_golden("detect_units2")
    for tag in ["nso2_q4", "nso2_q3", "nso2_q5", "nso2_q2", "nso2_q7"]:
        q, p1, R = (int(v) for v in g[tag + "_meta"])
        got = reconstruct.singleton_detection(g[tag + "_cols"].T, method_channel="nso", method_source="identity", q=q,
                                              source_parity=p1, nso_subtype="nso2")
        assert np.array_equal(got.T, g[tag + "_k"]), tag


def test_detect_units_mle():
    g = load_golden("detect_units2")
    for tag in ["mle_q2", "mle_q3", "mle_q4"]:
        sel, S = g[tag + "_selection"], g[tag + "_S"]
        ksel, sig = reconstruct.singleton_detection_mle(g[tag + "_cols"].T, selection=sel, S_slice=S)
        assert np.array_equal(ksel, g[tag + "_ksel"]), tag
        assert np.array_equal(sig.T, g[tag + "_sig"])
        k1, s1 = reconstruct.singleton_detection(g[tag + "_cols"][5], method_channel="mle", selection=sel, S_slice=S)
        assert int(k1) == int(g[tag + "_ksel"][5]) and np.array_equal(s1, g[tag + "_sig"][5])
        # residual = the oracle's minimum residual norm
        cols = torch.from_numpy(g[tag + "_cols"].astype(np.complex64)).to(DEV)
        idx, res = ops.detect_mle(cols, torch.from_numpy(S.astype(np.complex64)).to(DEV))
        for c in range(0, len(cols), 7):
            P = S.shape[0]
            alphas = np.conjugate(S).T @ g[tag + "_cols"][c] / P
            want = np.linalg.norm(g[tag + "_cols"][c] - (alphas * S).T, axis=1)
            assert int(idx[c]) == int(np.argmin(want))
            assert abs(float(res[c]) - want.min()) <= 1e-5 * max(1.0, want.min())


def test_detect_mle_large_random_vs_oracle():
    """More candidates than lanes, several warps per block, ties impossible (random data)."""
    rng = np.random.default_rng(5)
    P, K, N = 23, 1000, 77
    S = np.exp(2j * np.pi * rng.integers(0, 5, (P, K)) / 5)
    true = rng.integers(0, K, N)
    cols = (S[:, true] * rng.uniform(0.5, 2, N) + 0.3 * (rng.normal(size=(P, N)) + 1j * rng.normal(size=(P, N)))).T
    idx, _ = ops.detect_mle(torch.from_numpy(cols.astype(np.complex64)).to(DEV).contiguous(),
                            torch.from_numpy(S.astype(np.complex64)).to(DEV))
    want = [orc.detect_mle(c, np.arange(K), S)[2] for c in cols]
    assert idx.cpu().tolist() == want


def test_detect_coded_standalone_vs_oracle():
    """Channel stage + Reed-Solomon source stage in one call == oracle decode of the same symbols."""
    n, t, q, R = 20, 3, 3, 2
    dec = qsft_b200.get_reed_solomon_dec(n, t, q)
    rs = dec.__self__
    D = rs.get_delay_matrix()
    odec = orc.get_reed_solomon_dec(n, t, q)
    p1 = D.shape[0]
    rng = np.random.default_rng(3)
    cols, want = [], []
    for i in range(60):
        k = np.zeros(n, dtype=int)
        w = int(rng.integers(0, t + 2))                       # includes weight t + 1: decoder failure -> zeros
        pos = rng.choice(n, w, replace=False)
        k[pos] = rng.integers(1, q, w)
        off = rng.integers(0, q, (R, n))
        ph = np.concatenate([((off[r] - D) % q) @ k % q for r in range(R)])
        col = (1.3 - 0.4j) * np.exp(2j * np.pi * ph / q)
        cols.append(col)
        sym = orc.detect_nso1(col[:, None], q, p1)[:, 0]
        want.append(np.array(odec(list(sym))[0][0, :], dtype=int))
    got = reconstruct.singleton_detection(np.array(cols).T, method_channel="nso", method_source="coded", q=q,
                                          source_parity=p1, source_decoder=dec)
    assert np.array_equal(got.T, np.array(want))


def test_detect_argument_errors():
    cols = torch.zeros((4, 10), dtype=torch.complex64, device=DEV)
    with pytest.raises(qsft_b200.QsftError):
        ops.singleton_detect(cols, 4, 3, "nso")                # P not a multiple of P_src
    with pytest.raises(qsft_b200.QsftError):
        ops.singleton_detect(cols, 4, 5, "identity")           # identity channel needs num_repeat == 1
    with pytest.raises(ValueError):
        ops.singleton_detect(cols, 4, 5, "nso", nso_subtype="nso3")
    with pytest.raises(NotImplementedError):
        ops.channel_code("mle")
    assert ops.singleton_detect(cols[:0], 4, 5, "nso").shape == (0, 4)


@pytest.mark.parametrize("name", NSO2_CASES)
def test_k4_peel_nso2_from_reference_bins(name):
    """Peel the reference's own (noisy) bins with channel = nso2: same finds in the same order as the reference run
    whose singleton_detection was switched to nso_subtype="nso2"."""
    g = load_golden(name)
    p = case_params(g)
    q, n = p["q"], p["n"]
    U = torch.from_numpy(np.ascontiguousarray(g["mdu_Us"].reshape(p["trC"], -1, q ** p["trb"])).astype(np.complex64)).to(DEV)
    D = g["mdu_Ds"].reshape(p["trC"], -1, n)
    cutoff = 1e-9 + 1.5 * p["noise_sd"] ** 2 / q ** p["trb"]
    prob = ops.PeelProblem(q, n, p["trb"], list(g["mdu_Ms"]), D, p["P_src"], p["chan"], p["src"], cutoff, DEV,
                           nso_subtype="nso2")
    prob.alloc(4 * U.shape[0] * U.shape[2])
    prob.peel(U)
    dk, dv, _ = prob.distinct()
    want_keys = [tuple(int(v) for v in k) for k in g["res_keys"]]
    assert [tuple(int(v) for v in r) for r in dk] == want_keys
    assert np.max(np.abs(dv - g["res_vals"])) <= 1e-5 * np.max(np.abs(g["res_vals"]))


@pytest.mark.parametrize("name", NSO2_CASES)
def test_end_to_end_nso2_same_seed(name):
    g = load_golden(name)
    p = case_params(g)
    np.random.seed(p["seed"])
    sig = qsft_b200.get_random_subsampled_signal(n=p["n"], q=p["q"], sparsity=p["S"], a_min=1, a_max=1,
                                                  noise_sd=p["noise_sd"], query_args=dict(p["query_args"]),
                                                  max_weight=p["max_weight"])
    sft = qsft_b200.QSFT(num_subsample=p["trC"], num_repeat=p["trR"], b=p["trb"], reconstruct_method_source=p["src"],
                         reconstruct_method_channel=p["chan"], nso_subtype="nso2")
    res = sft.transform(sig, report=True, sort=True)
    assert np.random.random() == float(g["rng_probe"])
    want_keys = [tuple(int(v) for v in k) for k in g["res_keys"]]
    if p["noise_sd"] == 0:
        assert list(res["gwht"].keys()) == want_keys
        assert np.array_equal(np.array(res["locations"]), g["locations"])
    else:
        assert set(res["gwht"].keys()) == set(want_keys)
    got = np.array([res["gwht"][k] for k in want_keys])
    assert np.max(np.abs(got - g["res_vals"])) <= 1e-4


def test_synthetic_helper_sweep_end_to_end(tmp_path):
    """The reference's experiment harness on top of the CUDA path: SyntheticHelper builds train (qsft lattice, cached in
    the reference's folder layout) and test (uniform) signals, run_tests sweeps decoder settings, the test NMSE equals
    the reference's dense formula (qsft/test_helper.py:235-260) evaluated in NumPy."""
    from qsft_b200.parallel_tests import run_tests
    n, q, S = 8, 4, 20
    np.random.seed(31)
    sw, locq, strengths = qsft_b200.generate_signal_w(n, q, S, 1, 1, full=False)
    helper = qsft_b200.SyntheticHelper(signal_args={"n": n, "q": q, "locq": locq, "strengths": strengths},
                                       methods=["qsft"], subsampling=True, exp_dir=tmp_path,
                                       subsampling_args={"num_subsample": 3, "num_repeat": 2, "b": 4, "all_bs": [3, 4]},
                                       test_args={"n_samples": 500})
    for rel in ["config.json", "train/Ms_and_Ds.pickle", "train/samples/M0_D0.pickle", "train/transforms/U2_1.pickle",
                "test/signal_t.pickle"]:
        assert (tmp_path / rel).is_file(), rel
    df = run_tests("qsft", helper, 1, [2, 3], [1, 2], [3, 4], [0.0], parallel=False)
    assert len(df) == 8 and (df["n_samples"] == df["num_subsample"] * q ** df["b"] * df["num_repeat"] * (n + 1)).all()
    best = df[(df["num_subsample"] == 3) & (df["b"] == 4)]
    assert (best["found_sparsity"] == len(sw)).all() and (best["nmse"] < 1e-9).all()
    # NMSE of a deliberately wrong model against the reference's formula
    beta = {k: v * (1.1 if i % 2 else 1.0) for i, (k, v) in enumerate(sw.items())}
    got = helper.test_model("qsft", beta=beta)
    idx = list(helper.test_signal.signal_t.keys())
    y = np.array(list(helper.test_signal.signal_t.values()))
    dig = orc.dec_to_qary_vec(idx, q, n)
    y_hat = np.exp(2j * np.pi * (dig.T @ np.array(list(beta.keys())).T) / q) @ np.array(list(beta.values()))
    want = np.linalg.norm(y_hat - y) ** 2 / np.linalg.norm(y) ** 2
    assert abs(got - want) <= 1e-4 * want
    # a second helper on the same directory reuses the cached samples / transforms and gives the same model
    helper2 = qsft_b200.SyntheticHelper(signal_args={"n": n, "q": q, "locq": locq, "strengths": strengths},
                                        methods=["qsft"], subsampling=True, exp_dir=tmp_path,
                                        subsampling_args={"num_subsample": 3, "num_repeat": 2, "b": 4, "all_bs": [3, 4]},
                                        test_args={"n_samples": 500})
    assert helper2.test_signal.signal_t.keys() == helper.test_signal.signal_t.keys()
    m1 = helper.compute_model("qsft", {"num_subsample": 3, "num_repeat": 2, "b": 4, "noise_sd": 0.0})
    m2 = helper2.compute_model("qsft", {"num_subsample": 3, "num_repeat": 2, "b": 4, "noise_sd": 0.0})
    assert set(m1.keys()) == set(m2.keys()) == set(sw.keys())


@pytest.mark.parametrize("name", WIDE_FULL_CASES)
def test_wide_index_pipeline_matches_reference(name):
    """BASELINE config 4 shape (q=4, n=50: 100-bit indices, weight <= 3 support, nso R=3, 30 dB), reduced to b=4, against
    a full run of the unmodified reference: lattice indices (hi, lo limbs), samples, bins, peel from the reference's
    bins, and the same-seed end-to-end transform."""
    g = load_golden(name)
    p = case_params(g)
    q, n, b = p["q"], p["n"], p["b"]
    assert utils.index_limbs(q, n) == 2
    # K1: indices of group (0, 0)
    idx, _ = ops.query_lattice(g["Ms"][0], g["Ds"][0], q, device=DEV)
    host = idx.cpu().numpy().view(np.uint64)
    assert np.array_equal(host[..., 0], g["idx00_hi"]) and np.array_equal(host[..., 1], g["idx00_lo"])
    # construct (K1 + K2 + K3) with the reference's seed
    np.random.seed(p["seed"])
    sig = qsft_b200.get_random_subsampled_signal(n=n, q=q, sparsity=p["S"], a_min=1, a_max=1, noise_sd=p["noise_sd"],
                                                  query_args=dict(p["query_args"]), max_weight=p["max_weight"])
    assert np.array_equal(np.array(sig.Ms), g["Ms"]) and np.array_equal(np.array(sig.Ds[0]), g["Ds"])
    got = sig.subsample(u128_to_ints(g["idx00_hi"][1], g["idx00_lo"][1]))
    assert np.max(np.abs(got - g["samples00_row1"])) <= 1e-5 * max(1.0, np.max(np.abs(g["samples00_row1"])))
    mine = np.array([[sig.Us[i][j][b].cpu().numpy() for j in range(p["R"])] for i in range(p["C"])])
    assert np.max(np.abs(mine - g[f"Us_b{b}"])) <= 1e-5 * np.max(np.abs(g[f"Us_b{b}"]))
    # end to end, same seed: same support, NMSE within 1 % of the reference's
    sft = qsft_b200.QSFT(num_subsample=p["trC"], num_repeat=p["trR"], b=p["trb"], reconstruct_method_source=p["src"],
                         reconstruct_method_channel=p["chan"])
    res = sft.transform(sig, report=True, sort=True)
    assert np.random.random() == float(g["rng_probe"])
    want_keys = [tuple(int(v) for v in k) for k in g["res_keys"]]
    assert set(res["gwht"].keys()) == set(want_keys)
    vals = np.array([res["gwht"][k] for k in want_keys])
    assert np.max(np.abs(vals - g["res_vals"])) <= 1e-4
    true_w = dict(zip(map(tuple, g["locq"].T.tolist()), g["strengths"]))
    nm_ref, nm_got = orc.nmse(dict(zip(want_keys, g["res_vals"])), true_w), orc.nmse(res["gwht"], true_w)
    assert abs(nm_got - nm_ref) <= 0.01 * nm_ref + 1e-12
    assert res["n_samples"] == int(g["n_samples"]) and res["max_hamming_weight"] == int(g["max_hw"])
    # K4 alone on the reference's own noisy bins: same finds in the same first-seen order
    U = torch.from_numpy(np.ascontiguousarray(g["mdu_Us"].reshape(p["trC"], -1, q ** p["trb"])).astype(np.complex64)).to(DEV)
    D = g["mdu_Ds"].reshape(p["trC"], -1, n)
    prob = ops.PeelProblem(q, n, p["trb"], list(g["mdu_Ms"]), D, p["P_src"], p["chan"], p["src"],
                           1e-9 + 1.5 * p["noise_sd"] ** 2 / q ** p["trb"], DEV)
    prob.alloc(4 * U.shape[0] * U.shape[2])
    prob.peel(U)
    dk, dv, _ = prob.distinct()
    assert [tuple(int(v) for v in r) for r in dk] == want_keys
    assert np.max(np.abs(dv - g["res_vals"])) <= 1e-5 * np.max(np.abs(g["res_vals"]))


@pytest.mark.parametrize("q,n,b,S,R,chan,noise", [(4, 40, 10, 100_000, 1, "nso", 0.0), (4, 20, 7, 1000, 3, "nso", 3.16),
                                                  (4, 10, 4, 100, 1, "identity", 0.0), (3, 12, 5, 60, 2, "nso", 0.0),
                                                  (5, 6, 3, 25, 1, "identity", 0.02), (2, 100, 6, 30, 1, "nso", 0.0),
                                                  (4, 40, 9, 30_000, 1, "nso2", 0.0), (3, 7, 2, 12, 1, "identity", 0.0)])
def test_k4_device_loop_equals_host_driven_rounds(q, n, b, S, R, chan, noise, monkeypatch):
    """The persistent on-device round loop (k4_peel_loop.cu, default) against the host-driven classify / reduce / apply
    rounds (QSFT_K4_IMPL=1, the round-1 path the reference fixtures pinned): same distinct k in the same first-seen
    order, same number of rounds and finds, values equal to fp32 rounding; the device loop leaves U untouched.
    Odd q^b exercises the plain-copy tile variant (no TMA), the others the TMA variant."""
    np.random.seed(q * 100 + n)
    C = 3
    sw, locq, strengths = qsft_b200.generate_signal_w(n, q, S, 1, 1, full=False)
    channel = "nso" if chan.startswith("nso") else "identity"
    Ms, Ds = qsft_b200.get_Ms_and_Ds(n, q, query_method="complex", num_subsample=C, delays_method_source="identity",
                                     delays_method_channel=channel, num_repeat=R, b=b)
    ld = utils.padded_ld(n)
    loc = ops.pad_digits(locq.T, ld, DEV)
    a = torch.from_numpy(strengths.astype(np.complex64)).to(DEV)
    U = torch.stack([torch.cat([ops.closed_form_bins(Ms[c], Ds[c][r], q, loc, a) for r in range(R)]) for c in range(C)]).contiguous()
    if noise:
        U = U + (noise / np.sqrt(2 * q ** b)) * torch.view_as_complex(torch.randn(tuple(U.shape) + (2,), device=DEV))
    Dall = np.stack([np.vstack(Ds[c]) for c in range(C)])
    prob = ops.PeelProblem(q, n, b, Ms, Dall, n + 1, channel, "identity", 1e-9 + 1.5 * noise ** 2 / q ** b, DEV,
                           nso_subtype="nso2" if chan == "nso2" else "nso1")
    prob.alloc(4 * C * q ** b)
    monkeypatch.delenv("QSFT_K4_IMPL", raising=False)
    U0 = U.clone()
    nf_d, nr_d = prob.peel(U)
    assert torch.equal(U, U0)
    nf_d = len(prob.finds(nf_d)[0])                          # valid finds (the loop hands slots out in chunks)
    k_d, v_d, c_d = prob.distinct()
    monkeypatch.setenv("QSFT_K4_IMPL", "1")
    nf_h, nr_h = prob.peel(U0)
    assert len(prob.finds(nf_h)[0]) == nf_h
    k_h, v_h, c_h = prob.distinct()
    assert (nf_d, nr_d) == (nf_h, nr_h) and nf_d > 0
    assert np.array_equal(k_d, k_h) and np.array_equal(c_d, c_h)
    assert np.max(np.abs(v_d - v_h)) <= 1e-5 * np.max(np.abs(v_h))
    if noise == 0.0 and chan != "nso2":
        assert len(k_d) >= 0.9 * len(sw)


def test_dense_gwht_igwht_utils():
    """utils.gwht / igwht (the reference's dense helpers, qsft/utils.py:31-49) on the K3 kernel."""
    g = load_golden("gwht_units")
    for key in g.files:
        if key.startswith("x_"):
            _, qs, bs = key.split("_")
            q, b = int(qs[1:]), int(bs[1:])
            y = utils.gwht(g[key], q, b)
            assert np.max(np.abs(y - g["y" + key[1:]])) <= 1e-6 * np.max(np.abs(g[key]))
            back = utils.igwht(g["y" + key[1:]], q, b)
            assert np.max(np.abs(back - g[key])) <= 1e-5 * np.max(np.abs(g[key]))
            shaped = utils.gwht_tensored(g[key].reshape([q] * b), q, b)
            assert shaped.shape == tuple([q] * b) and np.max(np.abs(shaped.ravel() - y)) == 0


def test_device_noise_statistics_and_determinism():
    """qsft_add_noise (Philox): zero mean, the requested variance on both parts, the same values for the same (seed, offset)
    whatever the buffer split, different values for another seed."""
    n = 1 << 20
    base = torch.zeros(n, dtype=torch.complex64, device=DEV)
    a = ops.add_noise_(base.clone(), 0.5, seed=123)
    b = ops.add_noise_(base.clone(), 0.5, seed=123)
    c = ops.add_noise_(base.clone(), 0.5, seed=124)
    assert torch.equal(a, b) and not torch.equal(a, c)
    re, im = a.real.double(), a.imag.double()
    assert abs(float(re.mean())) < 3e-3 and abs(float(im.mean())) < 3e-3
    assert abs(float(re.var()) - 0.25) < 3e-3 and abs(float(im.var()) - 0.25) < 3e-3
    assert abs(float((re * im).mean())) < 3e-3
    # two halves with the second offset by the elements of the first = one call over the whole buffer
    h = n // 2
    lo = ops.add_noise_(base[:h].clone(), 0.5, seed=123, offset=0)
    hi = ops.add_noise_(base[h:].clone(), 0.5, seed=123, offset=h // 2)
    assert torch.equal(torch.cat([lo, hi]), a)
    odd = ops.add_noise_(torch.zeros(7, dtype=torch.complex64, device=DEV), 1.0, seed=5)
    assert bool((odd != 0).all())


def test_noisy_transform_with_device_noise_nmse():
    """Config-2 shape with noise drawn on the device: statistical parity with the reference's NMSE at the same SNR
    (3.393e-6 for the NumPy stream, seed 0) -- the device stream gives another draw of the same law."""
    n, q, S, b, C, R = 20, 4, 1000, 7, 3, 3
    qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "identity", "subsampling_method": "qsft",
          "delays_method_channel": "nso", "num_repeat": R, "b": b}
    np.random.seed(0)
    sig = qsft_b200.get_random_subsampled_signal(n=n, q=q, sparsity=S, a_min=1, a_max=1, noise_sd=3.1623, query_args=dict(qa),
                                                  noise_rng="device")
    got = qsft_b200.QSFT(num_subsample=C, num_repeat=R, b=b, reconstruct_method_source="identity",
                         reconstruct_method_channel="nso").transform(sig)
    true = sig.signal_w
    assert set(got.keys()) == set(true.keys())
    nmse = sum(abs(got[k] - v) ** 2 for k, v in true.items()) / sum(abs(v) ** 2 for v in true.values())
    assert 2.0e-6 < nmse < 5.5e-6, nmse
