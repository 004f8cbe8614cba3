"""CPU: the whole hot path through the DEVICE sources -- K1 lattice -> K2 (dp4a) evaluation -> K3 q-ary DFT -> K4 peeling
loop -- executed by the SIMT emulation (tests/emu) on the reference's fixtures: the bins equal the reference's Us and the
recovered coefficients equal the reference's result, in the same order.  (The tensor-core kernels cannot be emulated; on
the GPU they are tested bit-for-bit against this SIMT path.)"""
import numpy as np
import pytest

from conftest import case_params, load_golden
from test_emu_k1k2 import _p, k1, k2, lattice  # noqa: F401  (fixtures)
from test_emu_k3 import emu as k3emu, run as k3run  # noqa: F401
from test_emu_k4 import Problem, emu as k4emu  # noqa: F401


@pytest.mark.parametrize("name", ["cfg1_q4_n10_b4_identity", "q3_n12_b4_lowweight_nso", "q2_n12_b4_simple"])
def test_emulated_hot_path_end_to_end(k1, k2, k3emu, k4emu, name):
    g = load_golden(name)
    p = case_params(g)
    assert p["noise_sd"] == 0 and (p["trC"], p["trR"], p["trb"]) == (p["C"], p["R"], p["b"])
    q, n, b, C, R = p["q"], p["n"], p["b"], p["C"], p["R"]
    S = g["locq"].shape[1]
    a = np.ascontiguousarray(g["strengths"].astype(np.complex64))
    U = np.zeros((C, R * p["P_src"], q ** b), dtype=np.complex64)
    for c in range(C):
        for r in range(R):
            idx, dig = lattice(k1, g["Ms"][c], g["Ds"][r], q)                      # K1
            ld = dig.shape[-1]
            loc = np.zeros((S, ld), dtype=np.int8)
            loc[:, :n] = g["locq"].T
            qd = np.ascontiguousarray(dig.reshape(-1, ld))
            x = np.zeros(qd.shape[0], dtype=np.complex64)
            assert k2.emu_eval_synth(_p(qd), qd.shape[0], _p(loc), _p(a), S, q, n, ld, _p(x)) == 0   # K2
            y, _, _ = k3run(k3emu, x.reshape(p["P_src"], -1), q, b)                  # K3
            assert np.max(np.abs(y - g[f"Us_b{b}"][c, r])) <= 1e-5 * np.max(np.abs(g[f"Us_b{b}"][c, r]))
            U[c, r * p["P_src"]:(r + 1) * p["P_src"]] = y
    # the fixture's get_MDU may have permuted groups / repeats: rebuild the decoder's view from its Ms / Ds
    order_c = [int(np.nonzero([np.array_equal(g["mdu_Ms"][i], g["Ms"][c]) for c in range(C)])[0][0]) for i in range(C)]
    Uv = np.zeros_like(U)
    Dv = np.zeros((C, R * p["P_src"], n), dtype=np.int64)
    for i, c in enumerate(order_c):
        for j in range(R):
            r = int(np.nonzero([np.array_equal(g["mdu_Ds"][i][j], g["Ds"][rr]) for rr in range(R)])[0][0])
            Uv[i, j * p["P_src"]:(j + 1) * p["P_src"]] = U[c, r * p["P_src"]:(r + 1) * p["P_src"]]
            Dv[i, j * p["P_src"]:(j + 1) * p["P_src"]] = g["Ds"][r]
    prob = Problem(q, n, b, [g["Ms"][c] for c in order_c], Dv, p["P_src"], 0 if p["chan"] == "identity" else 1, 1e-9)
    keys, vals, _, _ = prob.peel(k4emu, np.ascontiguousarray(Uv), 1)                 # K4
    assert keys == [tuple(int(v) for v in k) for k in g["res_keys"]]
    assert np.max(np.abs(vals - g["res_vals"])) <= 1e-5
