"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the committed reference fixtures.
Bit-exact for integer / index work; complex64 values within rtol 1e-5 of the reference's complex128."""
import numpy as np
import pytest
import torch

import qsft_oracle as orc
from conftest import FULL_CASES, INDEX_CASES, case_params, load_golden, u128_to_ints

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import qsft_b200
    from qsft_b200 import ops, utils
    DEV = torch.device("cuda", 0)


def _build_signal(p, **kw):
    np.random.seed(p["seed"])
    return qsft_b200.get_random_subsampled_signal(n=p["n"], q=p["q"], sparsity=p["S"], a_min=1, a_max=1,
                                                  noise_sd=p["noise_sd"], query_args=dict(p["query_args"]),
                                                  max_weight=p["max_weight"], **kw)


# ---- K1 ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", INDEX_CASES)
def test_k1_lattice_bit_exact_wide(name):
    g = load_golden(name)
    q, n, b, P = (int(v) for v in g["meta"])
    idx, dig = ops.query_lattice(g["M"], g["D"], q, device=DEV, want_idx=True, want_digits=True)
    host = idx.cpu().numpy().view(np.uint64)
    limbs = utils.index_limbs(q, n)
    if limbs == 2:
        assert np.array_equal(host[..., 0], g["hi"]) and np.array_equal(host[..., 1], g["lo"])
    else:
        assert np.array_equal(host, g["lo"]) and not g["hi"].any()
    want_dig = orc.query_digits(g["M"], g["D"], q).transpose(0, 2, 1)          # (P, B, n)
    got = dig.cpu().numpy()
    assert np.array_equal(got[..., :n], want_dig) and not got[..., n:].any()
    # device codecs round trip
    flat = idx.reshape(-1, 2) if limbs == 2 else idx.reshape(-1)
    d2 = ops.dec_to_qary(flat.contiguous(), q, n)
    assert torch.equal(d2, dig.reshape(-1, dig.shape[-1]))
    i2 = ops.qary_to_dec(d2, q, n)
    assert torch.equal(i2, flat)
    assert np.array_equal(d2[:64, :n].cpu().numpy().T, g["digits64"])


@pytest.mark.parametrize("name", FULL_CASES)
def test_k1_lattice_matches_reference_group00(name):
    g = load_golden(name)
    p = case_params(g)
    idx, _ = ops.query_lattice(g["Ms"][0], g["Ds"][0], p["q"], device=DEV)
    host = idx.cpu().numpy().view(np.uint64)
    assert np.array_equal(host, g["idx00_lo"]) and not g["idx00_hi"].any()


def test_k1_large_lattice_properties():
    """Full-size lattice (q=4, n=40, b=10): every index decodes back to (M l + d) mod q -- checked on device."""
    rng = np.random.default_rng(0)
    q, n, b, P = 4, 40, 10, 3
    M, D = rng.integers(0, q, (n, b)), rng.integers(0, q, (P, n))
    idx, dig = ops.query_lattice(M, D, q, device=DEV, want_idx=True, want_digits=True)
    back = ops.dec_to_qary(idx.reshape(-1, 2), q, n)
    assert torch.equal(back, dig.reshape(-1, dig.shape[-1]))
    # spot check 1000 random lattice points against the oracle formula
    ls = rng.integers(0, q ** b, 1000)
    L = np.stack([(ls // q ** (b - 1 - i)) % q for i in range(b)])
    want = ((M @ L) % q + D[1][:, None]) % q
    assert np.array_equal(dig[1][torch.from_numpy(ls).to(DEV)][:, :n].cpu().numpy().T, want)


# ---- K2 ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("q,n,S,N", [(4, 10, 100, 3000), (3, 30, 257, 1001), (4, 40, 1000, 5000), (2, 100, 64, 777),
                                     (7, 22, 300, 512), (5, 6, 1, 10), (4, 20, 513, 1)])
def test_k2_eval_vs_oracle(impl, q, n, S, N):
    rng = np.random.default_rng(q * 1000 + n)
    qd, loc = rng.integers(0, q, (N, n)), rng.integers(0, q, (n, S))
    a = rng.uniform(0.5, 2, S) * np.exp(1j * rng.uniform(0, 2 * np.pi, S))
    want = orc.synth_eval_digits(qd, loc, a, q)
    ld = utils.padded_ld(n)
    got = ops.eval_synth(ops.pad_digits(qd, ld, DEV), ops.pad_digits(loc.T, ld, DEV),
                         torch.from_numpy(a.astype(np.complex64)).to(DEV), q, n, impl=impl).cpu().numpy()
    scale = np.sqrt(np.sum(np.abs(a) ** 2))
    assert np.max(np.abs(got - want)) <= 2e-6 * scale + 1e-6 * np.max(np.abs(want))


def test_k2_edge_cases():
    ld = 32
    z = torch.zeros((0, ld), dtype=torch.int8, device=DEV)
    a0 = torch.zeros(0, dtype=torch.complex64, device=DEV)
    assert ops.eval_synth(z, z, a0, 4, 10).numel() == 0
    qd = torch.ones((5, ld), dtype=torch.int8, device=DEV)
    out = ops.eval_synth(qd, z, a0, 4, 10)              # empty support -> zeros
    assert torch.equal(out, torch.zeros(5, dtype=torch.complex64, device=DEV))
    with pytest.raises(qsft_b200.QsftError):
        ops.eval_synth(qd, z, a0, 1, 10)                # q out of range


@pytest.mark.parametrize("name", FULL_CASES)
def test_k2_subsample_matches_reference_samples(name):
    """SyntheticSubsampledSignal.subsample(python ints) against samples produced by the reference."""
    g = load_golden(name)
    p = case_params(g)
    sig = _build_signal(p)
    ints = u128_to_ints(g["idx00_hi"][1], g["idx00_lo"][1])
    got = sig.subsample(ints)
    want = g["samples00_row1"]
    assert np.max(np.abs(got - want)) <= 1e-5 * max(1.0, np.max(np.abs(want)))
    assert sig.subsample([]).shape == (0,)


# ---- K3 ---------------------------------------------------------------------------------------------------
def test_k3_gwht_units_vs_reference():
    g = load_golden("gwht_units")
    for key in g.files:
        if key.startswith("x_"):
            _, qs, bs = key.split("_")
            q, b = int(qs[1:]), int(bs[1:])
            x = torch.from_numpy(g[key].astype(np.complex64)).to(DEV).reshape(1, -1).contiguous()
            y = ops.gwht_batch_(x, q, b).cpu().numpy()[0]
            want = g["y" + key[1:]]
            assert np.max(np.abs(y - want)) <= 1e-6 * np.max(np.abs(g[key])), key


@pytest.mark.parametrize("q,b,batch", [(4, 7, 5), (3, 8, 3), (4, 8, 2), (2, 13, 3), (5, 6, 2), (11, 3, 4), (4, 10, 2),
                                       (3, 9, 1), (6, 4, 2)])
def test_k3_gwht_multi_pass_vs_oracle(q, b, batch):
    rng = np.random.default_rng(b)
    x = rng.normal(size=(batch, q ** b)) + 1j * rng.normal(size=(batch, q ** b))
    want = np.stack([orc.gwht(r, q, b) for r in x])
    got = ops.gwht_batch_(torch.from_numpy(x.astype(np.complex64)).to(DEV), q, b).cpu().numpy()
    assert np.max(np.abs(got - want)) <= 2e-6 * np.max(np.abs(x)) / np.sqrt(q ** b) * np.sqrt(b) + 1e-7


@pytest.mark.parametrize("b,batch", [(6, 1), (6, 37), (7, 1), (7, 9), (8, 5), (9, 3), (10, 1), (10, 7)])
def test_k3_tma_pipeline_bit_identical_to_register_staged_kernels(b, batch, monkeypatch):
    """The TMA pipeline (k3_gwht_tma.cu, default for q = 4, 6 <= b <= 10) performs the same butterflies in the same order
    as the register-staged kernels (QSFT_K3_IMPL=1): bit-identical output, also through the peer-store entry point."""
    q = 4
    B = q ** b
    x = (torch.randn(batch, B, device=DEV) + 1j * torch.randn(batch, B, device=DEV)).to(torch.complex64)
    got = ops.gwht_batch_(x.clone(), q, b)
    peers = [torch.full((batch + 2, B), 7.0, dtype=torch.complex64, device=DEV) for _ in range(2)]
    got_b = ops.gwht_batch_bcast_(x.clone(), q, b, [p[1:].data_ptr() for p in peers])
    monkeypatch.setenv("QSFT_K3_IMPL", "1")
    want = ops.gwht_batch_(x.clone(), q, b)
    assert torch.equal(got, want)
    assert torch.equal(got_b, want)
    for p in peers:
        assert torch.equal(p[1:1 + batch], want)
        assert bool((p[0] == 7.0).all()) and bool((p[-1] == 7.0).all())


def test_k3_linearity_and_delta_full_size():
    """Size-independent properties at BASELINE size 4^10: transform of a delta is a pure character / q^b;
    linearity."""
    q, b = 4, 10
    B = q ** b
    x = torch.zeros((2, B), dtype=torch.complex64, device=DEV)
    l0 = 123457
    x[0, l0] = 1.0
    x[1] = torch.randn(B, device=DEV) + 1j * torch.randn(B, device=DEV)
    x1 = x[1].clone()
    y = ops.gwht_batch_(x.clone(), q, b)
    js = torch.arange(B, device=DEV)
    dots = torch.zeros(B, dtype=torch.int64, device=DEV)
    for i in range(b):
        dots += ((js // q ** i) % q) * ((l0 // q ** i) % q)
    want = torch.exp(-2j * np.pi * (dots % q).to(torch.float32) / q) / B
    assert torch.max(torch.abs(y[0] - want)) < 1e-9
    z = x.clone()
    z[0] = 2.5 * x[0] - 1j * x1
    yz = ops.gwht_batch_(z, q, b)
    assert torch.max(torch.abs(yz[0] - (2.5 * y[0] - 1j * y[1]))) < 1e-8


@pytest.mark.parametrize("q,b,batch", [(4, 10, 5), (4, 6, 41), (4, 4, 9), (4, 12, 2), (3, 7, 3), (2, 9, 4)])
def test_k3_bcast_matches_plain(q, b, batch):
    """The fused K3 + all-gather entry point writes the same bits to the local rows and to every peer buffer
    (peers emulated by other allocations of this device) as the plain in-place transform."""
    B = q ** b
    x = torch.randn(batch, B, device=DEV) + 1j * torch.randn(batch, B, device=DEV)
    x = x.to(torch.complex64)
    want = ops.gwht_batch_(x.clone(), q, b)
    peers = [torch.full((batch + 2, B), 7.0, dtype=torch.complex64, device=DEV) for _ in range(3)]
    got = ops.gwht_batch_bcast_(x.clone(), q, b, [p[1:].data_ptr() for p in peers])
    assert torch.equal(got, want)
    for p in peers:
        assert torch.equal(p[1:1 + batch], want)
        assert bool((p[0] == 7.0).all()) and bool((p[-1] == 7.0).all())     # neighbours untouched
    assert torch.equal(ops.gwht_batch_bcast_(x.clone(), q, b, []), want)


# ---- construct (K1 + K2 + K3) against the reference's Us ----------------------------------------------------
@pytest.mark.parametrize("name", FULL_CASES)
def test_construct_matches_reference_Us(name):
    g = load_golden(name)
    p = case_params(g)
    sig = _build_signal(p)
    assert np.array_equal(np.array(sig.Ms), g["Ms"]) and np.array_equal(np.array(sig.Ds[0]), g["Ds"])
    for bb in sig.all_bs:
        mine = np.array([[sig.Us[i][j][bb].cpu().numpy() for j in range(p["R"])] for i in range(p["C"])])
        want = g[f"Us_b{bb}"]
        assert np.max(np.abs(mine - want)) <= 1e-5 * np.max(np.abs(want)), bb
    assert sig.get_source_parity() == p["P_src"]
    with pytest.raises(ValueError):
        sig.get_MDU(p["C"] + 1, 1, p["b"])


# ---- K4: peel from the reference's own bins --------------------------------------------------------------
@pytest.mark.parametrize("name", FULL_CASES)
def test_k4_peel_from_reference_bins(name):
    g = load_golden(name)
    p = case_params(g)
    q, n = p["q"], p["n"]
    U = torch.from_numpy(np.ascontiguousarray(g["mdu_Us"].reshape(p["trC"], -1, q ** p["trb"])).astype(np.complex64)).to(DEV)
    D = g["mdu_Ds"].reshape(p["trC"], -1, n)
    cutoff = 1e-9 + 1.5 * p["noise_sd"] ** 2 / q ** p["trb"]
    prob = ops.PeelProblem(q, n, p["trb"], list(g["mdu_Ms"]), D, p["P_src"], p["chan"], p["src"], cutoff, DEV)
    prob.alloc(4 * U.shape[0] * U.shape[2])
    nf, nr = prob.peel(U)
    gw, keys = qsft_b200.QSFT._finds_to_dict(*prob.finds(nf))
    want_keys = [tuple(int(v) for v in k) for k in g["res_keys"]]
    assert list(gw.keys()) == want_keys                   # same support, same first-seen order
    got = np.array([gw[k] for k in want_keys])
    assert np.max(np.abs(got - g["res_vals"])) <= 1e-5 * np.max(np.abs(g["res_vals"]))
    # the device-side distinct-k list (averaging of duplicate finds on the GPU) agrees with the host grouping
    dk, dv, dc = prob.distinct()
    assert [tuple(int(v) for v in r) for r in dk] == want_keys
    assert np.max(np.abs(dv - got)) <= 1e-6 * np.max(np.abs(got))
    assert int(dc.sum()) == len(prob.finds(nf)[0])


# ---- end to end with the same seed ----------------------------------------------------------------------
@pytest.mark.parametrize("name", FULL_CASES)
def test_end_to_end_same_seed(name):
    g = load_golden(name)
    p = case_params(g)
    sig = _build_signal(p)
    sft = qsft_b200.QSFT(num_subsample=p["trC"], num_repeat=p["trR"], b=p["trb"],
                         reconstruct_method_source=p["src"], reconstruct_method_channel=p["chan"])
    res = sft.transform(sig, report=True, sort=True)
    assert np.random.random() == float(g["rng_probe"])           # RNG consumed exactly like the reference
    want_keys = [tuple(int(v) for v in k) for k in g["res_keys"]]
    true_w = dict(zip(map(tuple, g["locq"].T.tolist()), g["strengths"]))
    ref_gw = dict(zip(want_keys, g["res_vals"]))
    if p["noise_sd"] == 0:
        assert list(res["gwht"].keys()) == want_keys
        got = np.array([res["gwht"][k] for k in want_keys])
        assert np.max(np.abs(got - g["res_vals"])) <= 1e-5
        assert np.array_equal(np.array(res["locations"]), g["locations"])
    else:
        assert set(res["gwht"].keys()) == set(want_keys)
        got = np.array([res["gwht"][k] for k in want_keys])
        assert np.max(np.abs(got - g["res_vals"])) <= 1e-4
        nm_ref, nm_got = orc.nmse(ref_gw, true_w), orc.nmse(res["gwht"], true_w)
        assert abs(nm_got - nm_ref) <= 0.01 * nm_ref + 1e-12
    assert res["n_samples"] == int(g["n_samples"])
    assert res["max_hamming_weight"] == int(g["max_hw"])


def test_coded_reed_solomon_end_to_end_vs_oracle():
    """Config-3 shaped (reduced): q=3 coded delays.  Parity for D generation is UNPINNED (galois absent); the test
    checks CUDA == oracle on the same D and exact support recovery."""
    n, q, S, b, C, t = 20, 3, 60, 4, 3, 3
    qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "coded", "subsampling_method": "qsft",
          "delays_method_channel": "nso", "num_repeat": 2, "b": b, "t": t}
    np.random.seed(8)
    sig = qsft_b200.get_random_subsampled_signal(n=n, q=q, sparsity=S, a_min=1, a_max=1, noise_sd=0.0,
                                                  query_args=dict(qa), max_weight=t)
    np.random.seed(8)
    sw, locq, strengths = orc.generate_signal_w(n, q, S, 1, 1, max_weight=t)
    osig = orc.OracleSignal(n, q, dict(qa), locq, strengths, noise_sd=0.0, signal_w=sw)
    assert np.array_equal(np.array(sig.Ds[0]), np.array(osig.Ds[0]))
    st = np.random.get_state()
    want = orc.transform(osig, C, 2, b, "coded", "nso", source_decoder=orc.get_reed_solomon_dec(n, t, q))
    np.random.set_state(st)
    sft = qsft_b200.QSFT(num_subsample=C, num_repeat=2, b=b, reconstruct_method_source="coded",
                         reconstruct_method_channel="nso", source_decoder=qsft_b200.get_reed_solomon_dec(n, t, q))
    got = sft.transform(sig)
    assert list(got.keys()) == list(want.keys())                 # CUDA == oracle, same order
    assert set(got.keys()) <= set(sw.keys()) and len(got) >= 0.9 * len(sw)   # (bins are heavily loaded: S/B = 0.74)
    assert max(abs(got[k] - want[k]) for k in want) < 1e-5
    assert max(abs(got[k] - sw[k]) for k in got) < 1e-5


def test_peel_large_closed_form_exact_recovery():
    """BASELINE config-5 shape (q=4, n=40, b=10, C=3, nso R=1) at reduced sparsity: bins from the closed form,
    device peel must recover the support exactly (size-independent property: result == signal_w)."""
    q, n, b, C, R, S = 4, 40, 10, 3, 1, 20000
    np.random.seed(1)
    sw, locq, strengths = qsft_b200.generate_signal_w(n, q, S, 1, 1, 0, full=False)
    Ms, Ds = qsft_b200.get_Ms_and_Ds(n, q, query_method="complex", num_subsample=C, delays_method_source="identity",
                                     delays_method_channel="nso", num_repeat=R, b=b)
    ld = utils.padded_ld(n)
    loc = ops.pad_digits(locq.T, ld, DEV)
    a = torch.from_numpy(strengths.astype(np.complex64)).to(DEV)
    U = torch.stack([torch.cat([ops.closed_form_bins(Ms[c], Ds[c][r], q, loc, a) for r in range(R)]) for c in range(C)])
    D = np.stack([np.vstack(Ds[c]) for c in range(C)])
    prob = ops.PeelProblem(q, n, b, Ms, D, n + 1, "nso", "identity", 1e-9, DEV)
    prob.alloc(4 * C * q ** b)
    nf, nr = prob.peel(U.contiguous())
    gw, _ = qsft_b200.QSFT._finds_to_dict(*prob.finds(nf))
    assert set(gw.keys()) == set(sw.keys())
    err = max(abs(gw[k] - v) for k, v in sw.items())
    assert err < 1e-5, err


# ---- K2L: lattice-factorised evaluation (q = 4) -----------------------------------------------------------
@pytest.mark.parametrize("mode", [2, 0])
@pytest.mark.parametrize("q,n,b,S,P,seed", [(4, 14, 7, 700, 5, 0), (4, 40, 8, 3000, 3, 1), (4, 40, 10, 257, 2, 2),
                                            (4, 20, 9, 64, 4, 3), (4, 33, 7, 1, 1, 4),
                                            (2, 30, 14, 600, 3, 5), (2, 64, 15, 2000, 2, 6), (2, 20, 16, 513, 1, 7), (2, 100, 14, 1, 2, 8)])
def test_k2_lattice_matches_plain_path_and_oracle(q, n, b, S, P, seed, mode, monkeypatch):
    """mode = QSFT_LATTICE_SPARSE: 2 (default) 2:4-sparse A' generated in tensor memory, 0 dense A' materialised in HBM.
    q = 2 runs through the q = 4 kernels (phases doubled, one bit per lattice digit)."""
    monkeypatch.setenv("QSFT_LATTICE_SPARSE", str(mode))
    rng = np.random.default_rng(seed)
    M, D = rng.integers(0, q, (n, b)), rng.integers(0, q, (P, n))
    loc = rng.integers(0, q, (n, S))
    a = rng.uniform(0.2, 2, S) * np.exp(1j * rng.uniform(0, 2 * np.pi, S))
    ld = utils.padded_ld(n)
    loc_d = ops.pad_digits(loc.T, ld, DEV)
    a_d = torch.from_numpy(a.astype(np.complex64)).to(DEV)
    assert ops.lattice_supported(q, n, b, P, S)
    got = ops.eval_synth_lattice(M, D, loc_d, a_d, q)
    _, dig = ops.query_lattice(M, D, q, device=DEV, want_idx=False, want_digits=True, ld=ld)
    plain = ops.eval_synth(dig.view(-1, ld), loc_d, a_d, q, n, impl=1).view(P, q ** b)
    scale = float(np.sqrt(np.sum(np.abs(a) ** 2)))
    assert torch.max(torch.abs(got - plain)).item() <= 3e-6 * scale + 2e-6 * float(np.max(np.abs(a))) * np.sqrt(S)
    # oracle on a sample of lattice points of delay row P-1
    ls = rng.integers(0, q ** b, 200)
    L = np.stack([(ls // q ** (b - 1 - i)) % q for i in range(b)])
    qd = (((M @ L) % q + D[P - 1][:, None]) % q).T
    want = orc.synth_eval_digits(qd, loc, a, q)
    assert np.max(np.abs(got[P - 1].cpu().numpy()[ls] - want)) <= 3e-6 * scale + 2e-6 * float(np.max(np.abs(a))) * np.sqrt(S)


@pytest.mark.parametrize("q,n,b,S,P,seed,a_lo", [(3, 12, 7, 300, 4, 0, 1.0), (3, 30, 8, 5000, 9, 1, 1.0), (3, 20, 9, 777, 3, 2, 0.01),
                                                 (3, 16, 8, 1, 5, 3, 1.0),
                                                 (5, 12, 5, 300, 4, 4, 1.0), (5, 20, 6, 2000, 7, 5, 1.0), (5, 9, 7, 513, 2, 6, 0.01),
                                                 (7, 10, 4, 400, 6, 7, 1.0), (7, 16, 5, 1500, 3, 8, 1.0), (7, 8, 3, 1, 5, 9, 1.0)])
def test_k2_lattice_odd_q_matches_oracle_and_plain_path(q, n, b, S, P, seed, a_lo):
    """Odd-prime lattice evaluation (dense tcgen05 GEMM over Z[w] coefficients, ragged tiles, two real launches + combination;
    q = 3 specialised kernels, q = 5 / 7 the generic d = q - 1 construction) against the plain K1 + K2 path on the whole lattice
    and against the fp64 oracle on a sample of it; a_lo = 0.01 takes the residual pass."""
    rng = np.random.RandomState(seed)
    M = rng.randint(0, q, size=(n, b))
    D = rng.randint(0, q, size=(P, n))
    loc = rng.randint(0, q, size=(S, n))
    a = (rng.uniform(a_lo, 1.0, S) * np.exp(2j * np.pi * rng.uniform(0, 1, S))).astype(np.complex64)
    ld = utils.padded_ld(n)
    loc_d = ops.pad_digits(loc, ld, DEV)
    a_d = torch.from_numpy(a).to(DEV)
    assert ops.lattice_supported(q, n, b, P, S)
    got = ops.eval_synth_lattice(M, D, loc_d, a_d, q).cpu().numpy()
    _, dig = ops.query_lattice(M, D, q, device=DEV, want_idx=False, want_digits=True, ld=ld)
    plain = ops.eval_synth(dig.view(P * q ** b, ld), loc_d, a_d, q, n).view(P, q ** b).cpu().numpy()
    scale = float(np.sqrt(np.sum(np.abs(a) ** 2)))
    assert np.max(np.abs(got - plain)) <= 3e-6 * scale + 2e-6 * float(np.max(np.abs(a))) * np.sqrt(S)
    ls = rng.randint(0, q ** b, size=400)
    L = np.array([[(l // q ** (b - 1 - i)) % q for i in range(b)] for l in ls]).T
    for p in (0, P - 1):
        qd = (((M @ L) % q + D[p][:, None]) % q).T
        want = orc.synth_eval_digits(qd, loc.T, a, q)
        assert np.max(np.abs(got[p][ls] - want)) <= 3e-6 * scale + 2e-6 * float(np.max(np.abs(a))) * np.sqrt(S)


def test_k2_lattice_unsupported_shapes():
    assert not ops.lattice_supported(11, 10, 4, 3, 100) and not ops.lattice_supported(6, 10, 4, 3, 100)   # q = 2, 3, 4, 5, 7 only
    assert ops.lattice_supported(5, 10, 6, 3, 100) and not ops.lattice_supported(5, 10, 2, 3, 100)          # q = 5: 5^b2 >= 43 columns
    assert not ops.lattice_supported(2, 20, 13, 3, 100) and ops.lattice_supported(2, 20, 14, 3, 100)   # q = 2: 14 <= b <= 28
    assert not ops.lattice_supported(3, 10, 6, 3, 100)       # q = 3: b too small
    assert ops.lattice_supported(3, 10, 8, 3, 100)
    assert not ops.lattice_supported(4, 10, 4, 3, 100)       # b too small
    loc = torch.zeros((4, 32), dtype=torch.int8, device=DEV)
    a = torch.ones(4, dtype=torch.complex64, device=DEV)
    with pytest.raises(qsft_b200.QsftError):
        ops.eval_synth_lattice(np.zeros((10, 4), int), np.zeros((2, 10), int), loc, a, 4)


def test_q2_large_lattice_transform_uses_the_lattice_gemm_and_recovers_the_support():
    """q = 2, b = 14 (2^14 bins): sampled by the lattice GEMM (eval_impl 3 raises if the shape were unsupported), transformed
    and peeled; the result equals the signal and the run that samples with K1 + K2 (eval_impl 2)."""
    n, q, b, S, C = 40, 2, 14, 2000, 3
    qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "identity", "subsampling_method": "qsft",
          "delays_method_channel": "identity", "num_repeat": 1, "b": b}
    res = {}
    for impl in (3, 2):
        np.random.seed(21)
        sig = qsft_b200.get_random_subsampled_signal(n=n, q=q, sparsity=S, a_min=0.5, a_max=1, noise_sd=0.0,
                                                      query_args=dict(qa), eval_impl=impl)
        res[impl] = (qsft_b200.QSFT(num_subsample=C, num_repeat=1, b=b, reconstruct_method_source="identity",
                                    reconstruct_method_channel="identity").transform(sig), sig.signal_w)
    got, sw = res[3]
    assert list(got.keys()) == list(res[2][0].keys()) and set(got.keys()) == set(sw.keys())
    assert max(abs(got[k] - sw[k]) for k in sw) <= 1e-5
    assert max(abs(got[k] - res[2][0][k]) for k in sw) <= 1e-5


@pytest.mark.parametrize("impl", [2, 3, 0])
def test_construct_and_transform_all_eval_impls_agree(impl):
    """q = 4, b = 7 (lattice path eligible): the three evaluation kernels give the same bins and the same transform."""
    n, q, S, b, C, R = 16, 4, 900, 7, 3, 2
    qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "identity", "subsampling_method": "qsft",
          "delays_method_channel": "nso", "num_repeat": R, "b": b}
    np.random.seed(12)
    sig = qsft_b200.get_random_subsampled_signal(n=n, q=q, sparsity=S, a_min=0.5, a_max=1.5, noise_sd=0.0,
                                                  query_args=dict(qa), eval_impl=impl)
    for c in range(C):
        for r in range(R):
            want = orc.closed_form_bins(sig.Ms[c], sig.Ds[c][r], np.asarray(sig.locq), sig.strengths, q)
            got = sig.Us[c][r][b].cpu().numpy()
            assert np.max(np.abs(got - want)) <= 2e-6 * np.max(np.abs(want)) + 1e-7
    res = qsft_b200.QSFT(num_subsample=C, num_repeat=R, b=b, reconstruct_method_source="identity",
                         reconstruct_method_channel="nso").transform(sig)
    assert set(res.keys()) == set(sig.signal_w.keys())
    assert max(abs(res[k] - v) for k, v in sig.signal_w.items()) < 1e-5
    arr = qsft_b200.QSFT(num_subsample=C, num_repeat=R, b=b, reconstruct_method_source="identity",
                         reconstruct_method_channel="nso").transform(sig, output="arrays")
    assert [tuple(int(v) for v in r) for r in arr["locations"]] == list(res.keys())
    assert np.allclose(arr["values"], np.array(list(res.values())), rtol=0, atol=1e-6)


def test_transform_repeats_the_peel_when_the_find_list_is_too_small():
    """The reference has no limit on the number of singletons; a find list that runs out of slots is grown to the true bound
    (15 C B) and the peel repeated -- the on-device loop leaves the bins untouched, so the second run sees the same data."""
    n, q, b, C, R, S = 14, 4, 5, 3, 1, 400
    qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "identity", "subsampling_method": "qsft",
          "delays_method_channel": "nso", "num_repeat": R, "b": b}
    np.random.seed(31)
    sig = qsft_b200.get_random_subsampled_signal(n=n, q=q, sparsity=S, a_min=1, a_max=1, noise_sd=0.0, query_args=dict(qa))
    sft = qsft_b200.QSFT(num_subsample=C, num_repeat=R, b=b, reconstruct_method_source="identity", reconstruct_method_channel="nso")
    tight = sft.transform(sig, output="arrays", max_finds=64)
    assert sft.last_stats["finds"] > 64
    np.random.seed(31)
    sig2 = qsft_b200.get_random_subsampled_signal(n=n, q=q, sparsity=S, a_min=1, a_max=1, noise_sd=0.0, query_args=dict(qa))
    roomy = sft.transform(sig2, output="arrays")
    assert {tuple(k) for k in tight["locations"].tolist()} == {tuple(k) for k in roomy["locations"].tolist()} == set(sig.signal_w.keys())


def test_transform_device_async_pipelined_equals_synchronous():
    """output="device_async": three transforms queued back to back without any read-back; each PendingSpectrum, waited for
    afterwards, reports the statistics of ITS transform, and the last one's device list equals the synchronous result."""
    n, q, b, C, R = 16, 4, 7, 3, 1
    qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "identity", "subsampling_method": "qsft",
          "delays_method_channel": "nso", "num_repeat": R, "b": b}
    sigs = []
    for seed, S in ((21, 400), (22, 900), (23, 650)):
        np.random.seed(seed)
        sigs.append(qsft_b200.get_random_subsampled_signal(n=n, q=q, sparsity=S, a_min=1, a_max=1, noise_sd=0.0, query_args=dict(qa)))

    def run(sig, output):
        return qsft_b200.QSFT(num_subsample=C, num_repeat=R, b=b, reconstruct_method_source="identity",
                              reconstruct_method_channel="nso").transform(sig, output=output)

    pend = [run(sig, "device_async") for sig in sigs]
    stats = [p.wait() for p in pend]
    assert [st["distinct"] for st in stats] == [len(sig.signal_w) for sig in sigs]
    assert all(st["rounds"] >= 1 for st in stats)
    dev = pend[-1].device()                                 # the workspace still holds the last transform
    got_k, got_sum, got_cnt = dev["k"].cpu().numpy(), dev["sum"].cpu().numpy(), dev["count"].cpu().numpy()
    want = run(sigs[-1], "arrays")          # (get_MDU draws a new group order from the host RNG: compare as mappings)
    got = {tuple(int(v) for v in k): val for k, val in zip(got_k, got_sum / got_cnt)}
    ref = {tuple(int(v) for v in k): val for k, val in zip(want["locations"], want["values"])}
    assert got.keys() == ref.keys() == sigs[-1].signal_w.keys()
    assert max(abs(got[k] - ref[k]) for k in ref) < 1e-6


# ---- BASELINE configs at full size ----------------------------------------------------------------------------
def test_config2_full_size_noisy_nmse_matches_reference_run():
    """BASELINE config 2 at full size (q=4 n=20 b=7 S=1000 C=3 nso R=3, 20 dB, seed 0).  The unmodified reference
    recovered 1000/1000 coefficients with NMSE 3.3930553e-06 on this seed (BASELINE.md section 2, 523 s on 8 cores);
    same seed => same Ms, Ds, support, get_MDU order and noise, so the NMSE must agree within 1 %."""
    n, q, S, b, C, R = 20, 4, 1000, 7, 3, 3
    noise_sd = float(np.sqrt(S / 10 ** (20 / 10)))
    qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "identity", "subsampling_method": "qsft",
          "delays_method_channel": "nso", "num_repeat": R, "b": b}
    np.random.seed(0)
    sig = qsft_b200.get_random_subsampled_signal(n=n, q=q, sparsity=S, a_min=1, a_max=1, noise_sd=noise_sd,
                                                  query_args=dict(qa))
    res = qsft_b200.QSFT(num_subsample=C, num_repeat=R, b=b, reconstruct_method_source="identity",
                         reconstruct_method_channel="nso").transform(sig, report=True)
    assert res["n_samples"] == 3096576
    assert set(res["gwht"].keys()) == set(sig.signal_w.keys())
    nm = orc.nmse(res["gwht"], sig.signal_w)
    assert abs(nm - 3.3930553e-06) <= 0.01 * 3.3930553e-06, nm


def test_config3_shape_coded_q3_vs_oracle_closed_form():
    """BASELINE config 3 (q=3 n=30 b=8 S=5000, coded delays t=4, C=3): construct on the GPU (plain tcgen05 path, q=3),
    compare the bins with the closed form and the transform with the oracle fed the same bins."""
    n, q, S, b, C, t = 30, 3, 5000, 8, 3, 4
    qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "coded", "subsampling_method": "qsft",
          "delays_method_channel": "identity", "num_repeat": 1, "b": b, "t": t}
    np.random.seed(3)
    sig = qsft_b200.get_random_subsampled_signal(n=n, q=q, sparsity=S, a_min=1, a_max=1, noise_sd=0.0,
                                                  query_args=dict(qa), max_weight=t)
    assert sig.get_source_parity() == 33
    np.random.seed(3)
    sw, locq, strengths = orc.generate_signal_w(n, q, S, 1, 1, max_weight=t)
    osig = orc.OracleSignal(n, q, dict(qa), locq, strengths, noise_sd=0.0, signal_w=sw, use_closed_form=True)
    for c in range(C):
        want = osig.Us[c][0][b]
        got = sig.Us[c][0][b].cpu().numpy()
        assert np.max(np.abs(got - want)) <= 1e-5 * np.max(np.abs(want))
    st = np.random.get_state()
    want = orc.transform(osig, C, 1, b, "coded", "identity", source_decoder=orc.get_reed_solomon_dec(n, t, q))
    np.random.set_state(st)
    got = qsft_b200.QSFT(num_subsample=C, num_repeat=1, b=b, reconstruct_method_source="coded",
                         reconstruct_method_channel="identity",
                         source_decoder=qsft_b200.get_reed_solomon_dec(n, t, q)).transform(sig)
    assert list(got.keys()) == list(want.keys())
    assert max(abs(got[k] - want[k]) for k in want) < 1e-5
    assert len(got) >= 0.99 * len(sw) and set(got.keys()) <= set(sw.keys())


def test_config4_shape_wide_index_low_degree():
    """BASELINE config 4 shape with the reference-runnable delays (identity source + nso; 'coded' needs prime q):
    q=4 n=50 (100-bit indices) b=8, weight <= 3 support, 30 dB.  Exact support recovery and NMSE below the noise."""
    n, q, S, b, C, R = 50, 4, 1000, 8, 3, 3
    noise_sd = float(np.sqrt(S / 10 ** (30 / 10)))
    qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "identity", "subsampling_method": "qsft",
          "delays_method_channel": "nso", "num_repeat": R, "b": b}
    np.random.seed(4)
    sig = qsft_b200.get_random_subsampled_signal(n=n, q=q, sparsity=S, a_min=1, a_max=1, noise_sd=noise_sd,
                                                  query_args=dict(qa), max_weight=3)
    res = qsft_b200.QSFT(num_subsample=C, num_repeat=R, b=b, reconstruct_method_source="identity",
                         reconstruct_method_channel="nso").transform(sig, report=True, sort=True)
    assert set(res["gwht"].keys()) == set(sig.signal_w.keys())
    assert res["max_hamming_weight"] <= 3
    # the low-weight generator draws duplicate locations (reference quirk): the sampled signal carries the
    # SUM of their strengths while signal_w keeps the last one -> compare against the summed spectrum
    true_w = {}
    for k, a in zip(map(tuple, np.asarray(sig.locq).T.tolist()), sig.strengths):
        true_w[k] = true_w.get(k, 0) + a
    assert orc.nmse(res["gwht"], true_w) < 1e-4
    # generic-signal route on the same object: K1 indices as Python ints (100 bits) -> subsample() -> same samples
    idx = sig._get_qsft_query_indices(sig.Ms[0], sig.Ds[0][0][:2])
    assert max(int(v).bit_length() for v in idx[1][:1000]) > 64
    vals = sig.subsample(list(idx[1][:500]))
    dig = orc.dec_to_qary_vec(list(idx[1][:500]), q, n).T
    want = orc.synth_eval_digits(dig, np.asarray(sig.locq), sig.strengths, q)
    assert np.max(np.abs(vals - want)) <= 1e-4


# ---- "next" rows of SURVEY 8f: generic black-box signals, on-disk cache, test-NMSE evaluator -----------------
def test_generic_blackbox_signal_host_subsample_route():
    """A user signal that only implements subsample(query_indices) like the reference's RNA signal: indices come from
    K1 as Python ints, samples go back to the GPU for K3/K4."""
    n, q, b, C = 12, 3, 4, 3
    rng = np.random.default_rng(5)
    locq = rng.integers(0, q, (n, 25))
    a = np.exp(1j * rng.uniform(0, 2 * np.pi, 25))

    class BlackBox(qsft_b200.SubsampledSignal):
        calls = 0

        def subsample(self, query_indices):
            BlackBox.calls += 1
            assert all(isinstance(v, int) for v in list(query_indices)[:3])
            dig = orc.dec_to_qary_vec(list(query_indices), q, n).T
            return orc.synth_eval_digits(dig, locq, a, q)

    qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "identity", "subsampling_method": "qsft",
          "delays_method_channel": "identity", "num_repeat": 1, "b": b}
    np.random.seed(2)
    sig = BlackBox(n=n, q=q, query_args=qa, noise_sd=0)
    sig.noise_sd = 0
    assert BlackBox.calls == C                       # q^b = 81 <= 10000 -> one call per (M, D) block like the reference
    res = qsft_b200.QSFT(num_subsample=C, num_repeat=1, b=b, reconstruct_method_source="identity",
                         reconstruct_method_channel="identity").transform(sig)
    want = dict(zip(map(tuple, locq.T.tolist()), a))
    assert set(res.keys()) == set(want.keys())
    assert max(abs(res[k] - v) for k, v in want.items()) < 1e-5


def test_folder_cache_reference_layout(tmp_path):
    g = load_golden("q4_n10_allbs_subselect")
    p = case_params(g)
    folder = str(tmp_path / "sig")
    sig = _build_signal(p, folder=folder)
    import os
    assert os.path.isfile(f"{folder}/Ms_and_Ds.pickle") and os.path.isfile(f"{folder}/samples/M0_D0.pickle")
    Us_ij, Ts_ij = utils.load_data(f"{folder}/transforms/U2_1.pickle")
    assert sorted(Us_ij.keys()) == [2, 3, 4] and np.asarray(Us_ij[4]).shape == (p["P_src"], 4 ** 4)
    assert np.max(np.abs(np.asarray(Us_ij[3]) - g["Us_b3"][2, 1])) < 1e-6
    # second construction loads everything from disk (RNG untouched, no kernels needed for sampling)
    state = np.random.get_state()
    sig2 = qsft_b200.SyntheticSubsampledSignal(signal_w=sig.signal_w, locq=sig.locq, strengths=sig.strengths,
                                               noise_sd=sig.noise_sd, n=p["n"], q=p["q"], query_args=dict(p["query_args"]),
                                               folder=folder)
    assert np.array_equal(np.random.get_state()[1], state[1])
    assert np.array_equal(np.array(sig2.Ms), np.array(sig.Ms))
    for bb in sig.all_bs:
        assert torch.allclose(sig2.Us[1][0][bb], sig.Us[1][0][bb], atol=1e-7)


def test_test_nmse_evaluator():
    from qsft_b200.test_helper import evaluate_model, test_nmse
    q, n = 4, 40
    rng = np.random.default_rng(8)
    K = rng.integers(0, q, (300, n))
    beta = {tuple(int(v) for v in k): complex(np.exp(1j * rng.uniform(0, 6.28))) for k in K}
    idx = [int(rng.integers(0, 2 ** 62)) * (2 ** 18) + int(rng.integers(0, 2 ** 18)) for _ in range(2000)]   # 80-bit
    dig = orc.dec_to_qary_vec(idx, q, n).T
    y = orc.synth_eval_digits(dig, np.array(list(beta.keys())).T, np.array(list(beta.values())), q)
    assert np.max(np.abs(evaluate_model(beta, idx, q, n) - y)) < 1e-4
    assert test_nmse(beta, idx, y, q, n) < 1e-10
    half = dict(list(beta.items())[:150])
    y_half = orc.synth_eval_digits(dig, np.array(list(half.keys())).T, np.array(list(half.values())), q)
    ref = np.linalg.norm(y_half - y) ** 2 / np.linalg.norm(y) ** 2
    assert abs(test_nmse(half, idx, y, q, n) - ref) < 1e-6
    assert test_nmse({}, idx, y, q, n) == 1


def test_k2_lattice_row_chunking_gives_identical_samples(monkeypatch):
    """The limb operand is generated per chunk of delay rows when it would exceed the scratch budget."""
    q, n, b, S, P = 4, 24, 8, 500, 7
    rng = np.random.default_rng(21)
    M, D = rng.integers(0, q, (n, b)), rng.integers(0, q, (P, n))
    ld = utils.padded_ld(n)
    loc_d = ops.pad_digits(rng.integers(0, q, (S, n)), ld, DEV)
    a_d = torch.from_numpy(np.exp(1j * rng.uniform(0, 6.28, S)).astype(np.complex64)).to(DEV)
    whole = ops.eval_synth_lattice(M, D, loc_d, a_d, q)          # default: sparse A' generated in tensor memory
    outs = {}
    monkeypatch.setenv("QSFT_LATTICE_SPARSE", "0")               # dense A' materialised in HBM ...
    monkeypatch.delenv("QSFT_LATTICE_SCRATCH_GB", raising=False)
    outs["whole"] = ops.eval_synth_lattice(M, D, loc_d, a_d, q)
    monkeypatch.setenv("QSFT_LATTICE_SCRATCH_GB", "0.0003")      # ... <= 2 * 256 * 1024 B per delay row -> rows in several chunks
    outs["chunked"] = ops.eval_synth_lattice(M, D, loc_d, a_d, q)
    for key, val in outs.items():                                # same integer arithmetic -> bit identical
        assert torch.equal(whole, val), key


@pytest.mark.parametrize("mode", [2, 0])
def test_k2_lattice_residual_pass_wide_dynamic_range(mode, monkeypatch):
    """Strengths spanning 1000 : 1.  One GEMM pass quantises every strength with ONE scale (20 bits below max|a|): the
    absolute error ~7e-7 max|a| is a relative error of 7e-4 for the smallest coefficient.  With the residual pass (chosen
    automatically from the data, or by residual_passes=1) the samples agree with the fp64 oracle to fp32 rounding."""
    monkeypatch.setenv("QSFT_LATTICE_SPARSE", str(mode))
    q, n, b, S, P = 4, 20, 8, 600, 3
    rng = np.random.default_rng(8)
    M, D = rng.integers(0, q, (n, b)), rng.integers(0, q, (P, n))
    loc = rng.integers(0, q, (n, S))
    a = np.exp(1j * rng.uniform(0, 2 * np.pi, S)) * 10.0 ** rng.uniform(-3, 0, S)
    ld = utils.padded_ld(n)
    loc_d = ops.pad_digits(loc.T, ld, DEV)
    a_d = torch.from_numpy(a.astype(np.complex64)).to(DEV)
    one = ops.eval_synth_lattice(M, D, loc_d, a_d, q, residual_passes=0)
    two = ops.eval_synth_lattice(M, D, loc_d, a_d, q, residual_passes=1)
    auto = ops.eval_synth_lattice(M, D, loc_d, a_d, q)
    assert torch.equal(two, auto)
    ls = rng.integers(0, q ** b, 400)
    L = np.stack([(ls // q ** (b - 1 - i)) % q for i in range(b)])
    qd = (((M @ L) % q + D[1][:, None]) % q).T
    want = orc.synth_eval_digits(qd, loc, a, q)
    a32 = a.astype(np.complex64).astype(np.complex128)           # what the device was given
    want32 = orc.synth_eval_digits(qd, loc, a32, q)
    err1 = np.max(np.abs(one[1].cpu().numpy()[ls] - want32))
    err2 = np.max(np.abs(two[1].cpu().numpy()[ls] - want32))
    rms = float(np.sqrt(np.sum(np.abs(a) ** 2)))
    assert err2 <= 4e-7 * rms + 1e-9 and err2 < 0.2 * err1, (err1, err2)     # fp32 rounding of the sample itself
    assert np.max(np.abs(want - want32)) <= 1e-6 * rms


def test_wide_dynamic_range_coefficients_relative_error():
    """Full pipeline with a_min = 0.01, a_max = 1 (reference strength model, qsft/utils.py:161-164): every recovered
    coefficient within 1e-5 RELATIVE of its true value (north-star tolerance), noiseless."""
    np.random.seed(11)
    n, q, S, b, C = 20, 4, 1000, 7, 3
    qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "identity", "subsampling_method": "qsft",
          "delays_method_channel": "nso", "num_repeat": 1, "b": b}
    sig = qsft_b200.get_random_subsampled_signal(n=n, q=q, sparsity=S, a_min=0.01, a_max=1, noise_sd=0, query_args=dict(qa))
    assert sig._residual_passes == 1
    got = qsft_b200.QSFT(num_subsample=C, num_repeat=1, b=b, reconstruct_method_source="identity",
                         reconstruct_method_channel="nso").transform(sig)
    true = sig.signal_w
    assert set(got.keys()) == set(true.keys())
    rel = max(abs(got[k] - v) / abs(v) for k, v in true.items())
    assert rel <= 1e-5, rel
