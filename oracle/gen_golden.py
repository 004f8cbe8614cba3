"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) through oracle/ref_shim.py.

Run in the build container only (the reference is not present on the GPU box):
    python oracle/gen_golden.py
The resulting small fixtures are committed; tests/ compare both the oracle restatement (CPU tests) and the
CUDA path (gpu tests) against them.  Nothing here is product code.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()

from qsft.qsft import QSFT  # noqa: E402
from qsft.input_signal_subsampled import SubsampledSignal  # noqa: E402
from qsft.reconstruct import singleton_detection, singleton_detection_mle  # noqa: E402
import qsft.qsft as ref_qsft_module  # noqa: E402
from qsft.utils import qary_vec_to_dec, dec_to_qary_vec, gwht  # noqa: E402
from synt_exp.synt_src.synthetic_signal import get_random_subsampled_signal, generate_signal_w  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
MASK64 = (1 << 64) - 1


def split_u128(vals):
    vals = [int(v) for v in vals]
    hi = np.array([v >> 64 for v in vals], dtype=np.uint64)
    lo = np.array([v & MASK64 for v in vals], dtype=np.uint64)
    return hi, lo


def pack_result(gw):
    keys = np.array(list(gw.keys()), dtype=np.int8).reshape(len(gw), -1)
    vals = np.array(list(gw.values()), dtype=np.complex128)
    return keys, vals


def run_case(name, seed, n, q, S, b, C, R, src, chan, noise_sd, query_method="complex", max_weight=None,
             all_bs=None, tr=None, store_samples=True, nso_subtype=None):
    """Full reference run: construct (sample + FFT) then QSFT.transform; stores every stage.
    nso_subtype="nso2": QSFT.transform hard-codes nso_subtype="nso1" (qsft.py:171); the reference's own
    singleton_detection is wrapped so that this one keyword is replaced -- every line that runs is still the
    reference's (reconstruct.py:116-129 instead of :100-113)."""
    # the signal object does not keep `strengths`: draw them once with the same seed (generate_signal_w is the
    # first RNG consumer inside get_random_subsampled_signal), then re-seed and build the real thing
    np.random.seed(seed)
    _, locq0, strengths0 = generate_signal_w(n, q, S, 1, 1, noise_sd, full=False, max_weight=max_weight)
    np.random.seed(seed)
    qa = {"query_method": query_method, "num_subsample": C, "delays_method_source": src,
          "subsampling_method": "qsft", "delays_method_channel": chan, "num_repeat": R, "b": b}
    if all_bs is not None:
        qa["all_bs"] = all_bs
    sig = get_random_subsampled_signal(n=n, q=q, sparsity=S, a_min=1, a_max=1, noise_sd=noise_sd,
                                       query_args=qa, max_weight=max_weight)
    assert np.array_equal(locq0, sig.locq)
    tr = tr or {"num_subsample": C, "num_repeat": R, "b": b}
    state_before = np.random.get_state()
    # what get_MDU hands to the decoder (consumes RNG: choice, choice, normals) -- replay from the saved state
    Ms_r, Ds_r, Us_r = sig.get_MDU(tr["num_subsample"], tr["num_repeat"], tr["b"])
    np.random.set_state(state_before)
    sft = QSFT(num_subsample=tr["num_subsample"], num_repeat=tr["num_repeat"], b=tr["b"],
               reconstruct_method_source=src, reconstruct_method_channel=chan)
    if nso_subtype is not None:
        orig = ref_qsft_module.singleton_detection
        ref_qsft_module.singleton_detection = lambda U, **kw: orig(U, **{**kw, "nso_subtype": nso_subtype})
    try:
        res = sft.transform(sig, report=True, sort=True)
    finally:
        if nso_subtype is not None:
            ref_qsft_module.singleton_detection = orig
    rng_probe = np.random.random()          # pins RNG consumption order
    keys, vals = pack_result(res["gwht"])
    P_src = sig.Ds[0][0].shape[0]
    data = {
        "meta": np.array([seed, n, q, S, b, C, R, P_src], dtype=np.int64),
        "noise_sd": np.float64(noise_sd), "src": src, "chan": chan, "query_method": query_method,
        "max_weight": np.int64(-1 if max_weight is None else max_weight),
        "tr": np.array([tr["num_subsample"], tr["num_repeat"], tr["b"]], dtype=np.int64),
        "all_bs": np.array(sig.all_bs, dtype=np.int64),
        "Ms": np.array(sig.Ms, dtype=np.int8), "Ds": np.array(sig.Ds[0], dtype=np.int8),
        "locq": np.array(sig.locq, dtype=np.int8), "strengths": np.array(strengths0),
        "mdu_Ms": np.array(Ms_r, dtype=np.int8), "mdu_Ds": np.array(Ds_r, dtype=np.int8),
        "mdu_Us": np.array(Us_r),
        "res_keys": keys, "res_vals": vals, "n_samples": np.int64(res["n_samples"]),
        "locations": np.array(res["locations"], dtype=np.int8),
        "avg_hw": np.float64(res["avg_hamming_weight"]), "max_hw": np.int64(res["max_hamming_weight"]),
        "rng_probe": np.float64(rng_probe),
    }
    if nso_subtype is not None:
        data["nso_subtype"] = nso_subtype
    for bb in sig.all_bs:
        data[f"Us_b{bb}"] = np.array([[sig.Us[i][j][bb] for j in range(R)] for i in range(C)])
    if store_samples:
        # re-derive indices/samples of group (c=0, r=0) through the reference's own methods
        idx = sig._get_qsft_query_indices(sig.Ms[0], sig.Ds[0][0])
        hi, lo = split_u128(np.concatenate(idx))
        data["idx00_hi"], data["idx00_lo"] = hi.reshape(P_src, -1), lo.reshape(P_src, -1)
        data["samples00_row1"] = np.array(sig.subsample(idx[1]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
    nm = np.sum(np.abs(np.array([sig.signal_w.get(tuple(k), 0) for k in keys]) - vals) ** 2)
    print(f"{name}: found {len(keys)}/{len(sig.signal_w)} coeffs, sq.err {nm:.3e}, n_samples {res['n_samples']}")


def index_case(name, seed, q, n, b, P):
    """Wide-index fixture: _get_qsft_query_indices for random M, D (80/100-bit python ints)."""
    np.random.seed(seed)
    M = np.random.randint(q, size=(n, b))
    D = np.random.randint(q, size=(P, n))
    fake = types.SimpleNamespace(q=q, b=b, L=None)
    fake.get_all_qary_vectors = lambda: SubsampledSignal.get_all_qary_vectors(fake)
    idx = SubsampledSignal._get_qsft_query_indices(fake, M, D)
    flat = np.concatenate(idx)
    hi, lo = split_u128(flat)
    back = dec_to_qary_vec(list(flat[:64]), q, n)     # reference decode of the first 64 indices
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=np.array([q, n, b, P], dtype=np.int64),
                        M=M.astype(np.int8), D=D.astype(np.int8), hi=hi.reshape(P, -1), lo=lo.reshape(P, -1),
                        digits64=back.astype(np.int8))
    print(f"{name}: {len(flat)} indices, max bits {max(int(v).bit_length() for v in flat)}")


def detect_case(name, seed):
    """singleton_detection unit vectors: noiseless and nso1, incl. noisy columns."""
    np.random.seed(seed)
    out = {}
    for tag, q, p1, R, chan in [("nl_q4", 4, 11, 1, "identity"), ("nl_q3", 3, 9, 1, "identity"),
                                ("nso_q4", 4, 9, 3, "nso"), ("nso_q5", 5, 7, 2, "nso"), ("nso_q2", 2, 13, 4, "nso")]:
        P = p1 * R
        cols = []
        ks = []
        for t in range(64):
            k = np.random.randint(q, size=p1 - 1)
            base = np.concatenate([[0], k])
            if chan == "nso":
                ph = np.concatenate([(np.random.randint(q) - base) % q for _ in range(R)])
            else:
                ph = base
            amp = np.random.uniform(0.5, 2) * np.exp(1j * np.random.uniform(0, 2 * np.pi))
            col = amp * np.exp(2j * np.pi * ph / q) + (t % 4) * 0.15 * (np.random.normal(size=P) + 1j * np.random.normal(size=P))
            cols.append(col)
            ks.append(singleton_detection(col, method_channel=chan, method_source="identity", q=q,
                                          source_parity=p1, nso_subtype="nso1"))
        out[tag + "_cols"] = np.array(cols)
        out[tag + "_k"] = np.array(ks, dtype=np.int8)
        out[tag + "_meta"] = np.array([q, p1, R], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: ok")


def detect_case2(name, seed):
    """Unit vectors for the two detectors QSFT.transform never selects: nso2 (hard decision, reconstruct.py:116-129)
    and mle (reconstruct.py:54-84, called directly with selection / S_slice)."""
    np.random.seed(seed)
    out = {}
    for tag, q, p1, R in [("nso2_q4", 4, 9, 3), ("nso2_q3", 3, 8, 5), ("nso2_q5", 5, 7, 2), ("nso2_q2", 2, 13, 4),
                          ("nso2_q7", 7, 5, 1)]:
        P = p1 * R
        cols, ks = [], []
        for t in range(96):
            k = np.random.randint(q, size=p1 - 1)
            base = np.concatenate([[0], k])
            ph = np.concatenate([(np.random.randint(q) - base) % q for _ in range(R)])
            amp = np.random.uniform(0.5, 2) * np.exp(1j * np.random.uniform(0, 2 * np.pi))
            col = amp * np.exp(2j * np.pi * ph / q) + (t % 4) * 0.15 * (np.random.normal(size=P) + 1j * np.random.normal(size=P))
            cols.append(col)
            ks.append(singleton_detection(col, method_channel="nso", method_source="identity", q=q,
                                          source_parity=p1, nso_subtype="nso2"))
        out[tag + "_cols"] = np.array(cols)
        out[tag + "_k"] = np.array(ks, dtype=np.int8)
        out[tag + "_meta"] = np.array([q, p1, R], dtype=np.int64)
    for tag, q, n, b, P in [("mle_q2", 2, 8, 3, 12), ("mle_q3", 3, 6, 2, 9), ("mle_q4", 4, 5, 2, 10)]:
        M = np.random.randint(q, size=(n, b))
        D = np.random.randint(q, size=(P, n))
        allk = np.array(list(np.ndindex(*([q] * n))), dtype=np.int64).T          # (n, q^n), MSB first
        hashes = (M.T @ allk) % q
        j = hashes[:, np.random.randint(allk.shape[1])]                           # a non-empty bin
        pre = allk[:, np.all(hashes == j[:, None], axis=0)]                       # candidates k with M^T k = j
        selection = np.array([int(v) for v in qary_vec_to_dec(pre, q)], dtype=np.int64)
        S_slice = np.exp(2j * np.pi * ((D @ pre) % q) / q)                        # (P, K) signatures under D
        cols, sel_out, sig_out = [], [], []
        for t in range(48):
            true = np.random.randint(pre.shape[1])
            amp = np.random.uniform(0.5, 2) * np.exp(1j * np.random.uniform(0, 2 * np.pi))
            col = amp * S_slice[:, true] + (t % 4) * 0.2 * (np.random.normal(size=P) + 1j * np.random.normal(size=P))
            ksel, sig = singleton_detection_mle(col, selection=selection, S_slice=S_slice, q=q, source_parity=P)
            cols.append(col)
            sel_out.append(int(ksel))
            sig_out.append(sig)
        out[tag + "_cols"] = np.array(cols)
        out[tag + "_selection"] = selection
        out[tag + "_S"] = S_slice
        out[tag + "_ksel"] = np.array(sel_out, dtype=np.int64)
        out[tag + "_sig"] = np.array(sig_out)
        out[tag + "_meta"] = np.array([q, n, b, P], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: ok")


def gwht_case(name, seed):
    np.random.seed(seed)
    out = {}
    for q, b in [(2, 6), (3, 4), (4, 4), (5, 3), (7, 2), (4, 6)]:
        x = np.random.normal(size=q ** b) + 1j * np.random.normal(size=q ** b)
        out[f"x_q{q}_b{b}"] = x
        out[f"y_q{q}_b{b}"] = gwht(x, q, b)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: ok")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--detectors" in sys.argv:      # only the fixtures added for nso2 / mle (the others stay byte-identical)
        detect_case2("detect_units2", 12)
        run_case("q4_n10_b4_nso2_noisy", 13, n=10, q=4, S=40, b=4, C=3, R=3, src="identity", chan="nso", noise_sd=0.3,
                 nso_subtype="nso2", store_samples=False)
        run_case("q3_n9_b3_nso2", 14, n=9, q=3, S=20, b=3, C=3, R=2, src="identity", chan="nso", noise_sd=0.0,
                 nso_subtype="nso2", store_samples=False)
        sys.exit(0)
    if "--wide" in sys.argv:           # BASELINE config 4 shape, reduced: 100-bit indices through the whole reference pipeline
        run_case("cfg4r_q4_n50_b4_lowweight_nso_noisy", 21, n=50, q=4, S=30, b=4, C=3, R=3, src="identity", chan="nso",
                 noise_sd=float(np.sqrt(30 / 1000.0)), max_weight=3)
        sys.exit(0)
    if "--wide2" in sys.argv:          # q = 2 with 100-bit indices, noiseless identity delays (P_src = 101), loaded bins
        run_case("q2_n100_b5_identity_wide", 22, n=100, q=2, S=14, b=5, C=3, R=1, src="identity", chan="identity",
                 noise_sd=0.0)
        sys.exit(0)
    # BASELINE config 1 (seed 20 = quick_example convention)
    run_case("cfg1_q4_n10_b4_identity", 20, n=10, q=4, S=100, b=4, C=3, R=1, src="identity", chan="identity", noise_sd=0.0)
    # config-2 shaped, reduced: nso R=3, 20 dB  (noise_sd = sqrt(S / 10^(SNR/10)))
    run_case("cfg2r_q4_n14_b5_nso_noisy", 0, n=14, q=4, S=150, b=5, C=3, R=3, src="identity", chan="nso",
             noise_sd=float(np.sqrt(150 / 100.0)))
    # q = 3 low-weight support (generate_signal_w max_weight branch), nso R=2, noiseless
    run_case("q3_n12_b4_lowweight_nso", 3, n=12, q=3, S=40, b=4, C=3, R=2, src="identity", chan="nso", noise_sd=0.0,
             max_weight=3)
    # 'simple' block-identity Ms, q = 2
    run_case("q2_n12_b4_simple", 5, n=12, q=2, S=12, b=4, C=3, R=1, src="identity", chan="identity", noise_sd=0.0,
             query_method="simple")
    # all_bs + get_MDU sub-selection (C'=2 of 3, R'=1 of 2, b'=3 of 4), noisy
    run_case("q4_n10_allbs_subselect", 7, n=10, q=4, S=30, b=4, C=3, R=2, src="identity", chan="nso", noise_sd=0.05,
             all_bs=[2, 3, 4], tr={"num_subsample": 2, "num_repeat": 1, "b": 3})
    # q = 5, identity channel with noise (quick_example style: noisy + identity/identity)
    run_case("q5_n6_b3_identity_noisy", 11, n=6, q=5, S=25, b=3, C=3, R=1, src="identity", chan="identity", noise_sd=0.02)
    # wide indices
    index_case("idx_q4_n40_b3", 1, q=4, n=40, b=3, P=5)        # 80 bits
    index_case("idx_q4_n50_b2", 2, q=4, n=50, b=2, P=4)        # 100 bits
    index_case("idx_q3_n45_b3", 3, q=3, n=45, b=3, P=4)        # 72 bits
    index_case("idx_q7_n22_b2", 4, q=7, n=22, b=2, P=3)        # 62 bits
    index_case("idx_q2_n100_b5", 6, q=2, n=100, b=5, P=3)      # 100 bits
    detect_case("detect_units", 9)
    gwht_case("gwht_units", 10)
