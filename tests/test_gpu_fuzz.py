"""Differential fuzz of the whole transform path (K1 + K2 / K2L + K3 + K4 through the public API) against the oracle on
seeded random shapes: every q the detectors take, ragged n / b / C / R, both channel detectors, coded source delays,
overloaded bins (the peel stalls: the partial result must be the same partial result), supports of one element, and
alphabets so small that q^n <= 15 C B (the `num_peeling < q^n` guard of qsft.py:151 can bind).
Noiseless: same keys in the same first-seen order, values within 1e-5 (the north star's tolerance)."""
import numpy as np
import pytest
import torch

import qsft_oracle as orc

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import qsft_b200


def _draw(seed):
    """One random problem; small enough that the oracle's transform takes well under a second."""
    r = np.random.RandomState(1000 + seed)
    q = int(r.choice([2, 3, 4, 4, 5, 7]))
    bmax = {2: 8, 3: 5, 4: 5, 5: 3, 7: 3}[q]
    b = int(r.randint(1 if q > 2 else 2, bmax + 1))
    n = int(r.randint(b, min(b + 12, 40) + 1))
    C = int(r.randint(1, 5))
    chan = str(r.choice(["identity", "nso"]))
    R = int(r.randint(1, 4)) if chan == "nso" else 1
    B = q ** b
    load = float(r.choice([0.05, 0.2, 0.5, 1.0, 2.5]))                      # S / B: from empty bins to a stalled peel
    S = int(max(1, min(round(load * B), 0.5 * float(q) ** n, 400)))
    src = "identity"
    t = None
    if q in (3, 5, 7) and n >= 6 and r.rand() < 0.35:
        src, t = "coded", int(r.randint(1, 4))
    qa = {"query_method": str(r.choice(["simple", "complex"])), "num_subsample": C, "delays_method_source": src,
          "subsampling_method": "qsft", "delays_method_channel": chan, "num_repeat": R, "b": b}
    if qa["query_method"] == "simple" and C * b > n:
        qa["query_method"] = "complex"                                     # `simple` needs C b <= n (query.py:12-22)
    if t is not None:
        qa["t"] = t
    a_min = float(r.choice([1.0, 0.3]))
    return dict(q=q, n=n, b=b, C=C, R=R, S=S, src=src, chan=chan, t=t, qa=qa, a_min=a_min)


def _both(p, seed, noise_sd=0.0):
    q, n, b, C, R, t = p["q"], p["n"], p["b"], p["C"], p["R"], p["t"]
    np.random.seed(seed)
    sig = qsft_b200.get_random_subsampled_signal(n=n, q=q, sparsity=p["S"], a_min=p["a_min"], a_max=1, noise_sd=noise_sd,
                                                  query_args=dict(p["qa"]), max_weight=t)
    np.random.seed(seed)
    sw, locq, strengths = orc.generate_signal_w(n, q, p["S"], p["a_min"], 1, max_weight=t)
    osig = orc.OracleSignal(n, q, dict(p["qa"]), locq, strengths, noise_sd=noise_sd, signal_w=sw)
    for c in range(C):
        assert np.array_equal(np.asarray(sig.Ms[c]), np.asarray(osig.Ms[c]))
        for r_ in range(R):
            assert np.array_equal(np.asarray(sig.Ds[c][r_]), np.asarray(osig.Ds[c][r_]))
    st = np.random.get_state()
    dec_o = orc.get_reed_solomon_dec(n, t, q) if t is not None else None
    want = orc.transform(osig, C, R, b, p["src"], p["chan"], source_decoder=dec_o)
    probe_o = np.random.random()
    np.random.set_state(st)
    dec = qsft_b200.get_reed_solomon_dec(n, t, q) if t is not None else None
    got = qsft_b200.QSFT(num_subsample=C, num_repeat=R, b=b, reconstruct_method_source=p["src"],
                         reconstruct_method_channel=p["chan"], source_decoder=dec).transform(sig)
    assert np.random.random() == probe_o                                   # the RNG stream is consumed identically
    return sw, want, got


@pytest.mark.parametrize("seed", range(48))
def test_fuzz_noiseless_transform_equals_oracle(seed):
    p = _draw(seed)
    sw, want, got = _both(p, seed)
    assert list(got.keys()) == list(want.keys()), p                        # same support, same first-seen order
    if want:
        scale = max(1.0, max(abs(v) for v in want.values()))
        assert max(abs(got[k] - want[k]) for k in want) <= 1e-5 * scale, p


@pytest.mark.parametrize("q,n,b,C", [(2, 3, 3, 4), (2, 4, 2, 3), (3, 2, 2, 2), (2, 5, 4, 4)])
def test_tiny_alphabet_peeling_guard(q, n, b, C):
    """q^n <= 15 C B: the reference's loop can stop on `num_peeling < q^n` (qsft.py:151) before the 15-round limit."""
    for seed in range(6):
        qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "identity",
              "subsampling_method": "qsft", "delays_method_channel": "identity", "num_repeat": 1, "b": b}
        S = max(1, min(q ** n // 2, 3 + seed))
        p = dict(q=q, n=n, b=b, C=C, R=1, S=S, src="identity", chan="identity", t=None, qa=qa, a_min=1.0)
        sw, want, got = _both(p, 77 + seed)
        assert list(got.keys()) == list(want.keys()), (p, seed)
        if want:
            assert max(abs(got[k] - want[k]) for k in want) <= 1e-5, (p, seed)


@pytest.mark.parametrize("seed,q,n,b,snr_db", [(0, 4, 12, 4, 25.0), (1, 3, 10, 4, 30.0), (2, 2, 16, 6, 25.0)])
def test_fuzz_noisy_nso_nmse_close_to_oracle(seed, q, n, b, snr_db):
    """Noisy NSO runs on the same seed: same support as the oracle's complex128 run and NMSE within 1 %."""
    C, R, S = 3, 3, max(4, (q ** b) // 8)
    qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "identity", "subsampling_method": "qsft",
          "delays_method_channel": "nso", "num_repeat": R, "b": b}
    p = dict(q=q, n=n, b=b, C=C, R=R, S=S, src="identity", chan="nso", t=None, qa=qa, a_min=1.0)
    noise_sd = float(np.sqrt(S * 10 ** (-snr_db / 10)))
    sw, want, got = _both(p, 500 + seed, noise_sd=noise_sd)
    assert set(got.keys()) == set(want.keys())
    nm_o, nm_g = orc.nmse(want, sw), orc.nmse(got, sw)
    assert abs(nm_g - nm_o) <= 0.01 * nm_o + 1e-12, (nm_g, nm_o)
