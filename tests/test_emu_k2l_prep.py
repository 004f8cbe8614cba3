"""CPU: the operand-generation kernels of the lattice-factorised evaluation (lt_prep / lt_amax / lt_quant / lt_agen /
lt_bgen / lt_ttab / lt_etab in k2_eval_lattice.cu) executed by the SIMT emulation (tests/emu).  With the tensor-core GEMM
replaced by an exact integer matrix product in NumPy, the operands they produce must reproduce the samples of the lattice
{M l + d_p}: the whole factorisation  x_p[l_hi, l_lo] = sum_s i^(<h_hi, l_hi> + e_ps) * (a_s i^<h_lo, l_lo>)  is checked
against the oracle's direct evaluation, and the packed phase tables against the materialised operand."""
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
    L = C.CDLL(build_emu.build(which="k2l"))
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_longlong
    L.emu_lt_prep.argtypes = [vp, vp, vp, i64, i64, i32, i32, i32, i32, i32, vp, vp, vp, i32]
    L.emu_lt_agen3.argtypes = [vp, vp, i64, i64, i32, i32, i64, i64, vp]
    L.emu_lt_bgen3.argtypes = [vp, vp, i64, i32, i64, i64, i32, vp]
    L.emu_lt_combine3.argtypes = [vp, vp, i64, vp]
    L.emu_lt_agenq.argtypes = [i32, vp, vp, i64, i64, i32, i32, i64, i64, vp]
    L.emu_lt_bgenq.argtypes = [i32, vp, vp, i64, i32, i64, i64, i32, vp]
    L.emu_lt_combineq.argtypes = [i32, vp, vp, i64, i64, vp]
    L.emu_lt_quant.argtypes = [vp, i64, vp, vp, C.c_int]
    L.emu_lt_bgen.argtypes = [vp, vp, i64, i32, i64, i64, vp, i32]
    L.emu_lt_ttab.argtypes = [vp, i64, i32, i64, i64, vp, i32]
    L.emu_lt_etab.argtypes = [vp, i64, i64, i32, i64, vp]
    L.emu_lt_agen.argtypes = [vp, vp, i64, i64, i32, i32, i64, i64, vp, i32]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("q,n,b,S,P,seed", [(4, 12, 7, 37, 5, 0), (4, 40, 8, 90, 3, 1), (4, 9, 7, 64, 2, 2),
                                            (2, 20, 9, 50, 4, 3), (2, 64, 12, 33, 3, 4), (2, 24, 15, 21, 2, 5)])
def test_emulated_lattice_operands_reproduce_the_samples(emu, q, n, b, S, P, seed):
    """q = 4, and q = 2 through the same kernels (digits doubled, lattice indices spread to one field per bit)."""
    spread, mul = (1, 2) if q == 2 else (0, 1)
    rng = np.random.default_rng(seed)
    M, D = rng.integers(0, q, (n, b)), rng.integers(0, q, (P, n))
    locq = rng.integers(0, q, (n, S))
    a = rng.uniform(0.3, 2, S) * np.exp(1j * rng.uniform(0, 2 * np.pi, S))
    b1 = (b - 1) // 2 if q == 2 else b // 2              # lt_split_b1
    b2 = b - b1
    Mhi, Nlo = q ** b1, q ** b2
    ld = max(32, (n + 31) // 32 * 32)
    Kp = (2 * S + 127) // 128 * 128
    Tw = Kp // 32
    Se = (S + 3) & ~3
    M8, D8 = np.ascontiguousarray(M, dtype=np.int8), np.ascontiguousarray(D, dtype=np.int8)
    loc = np.zeros((S, ld), dtype=np.int8)
    loc[:, :n] = locq.T
    hhi, hlo = np.zeros(S, dtype=np.uint32), np.zeros(S, dtype=np.uint32)
    e = np.zeros((P, Se), dtype=np.uint8)
    assert emu.emu_lt_prep(_p(M8), _p(D8), _p(loc), S, Se, n, b, b1, P, ld, _p(hhi), _p(hlo), _p(e), q) == 0
    # bin-hash halves and delay phases against their definitions
    H = (M.T @ locq) % q * mul                                             # (b, S) digits of M^T k (quarter turns), MSB first
    want_hi = sum(H[i].astype(np.int64) << (2 * (b1 - 1 - i)) for i in range(b1))
    want_lo = sum(H[b1 + i].astype(np.int64) << (2 * (b2 - 1 - i)) for i in range(b2))
    assert np.array_equal(hhi, want_hi) and np.array_equal(hlo, want_lo)
    assert np.array_equal(e[:, :S], (D @ locq) % q * mul)
    # quantisation: three balanced base-128 limbs
    a32 = np.ascontiguousarray(a.astype(np.complex64))
    inv_scale = np.zeros(2, dtype=np.float32)
    alimb = np.zeros((S, 2), dtype=np.int32)
    # residual pass: limbs of (a * scale - round(a * scale)) * 2^21; both passes together reproduce a to ~2^-41 max|a|
    alimb2 = np.zeros((S, 2), dtype=np.int32)
    assert emu.emu_lt_quant(_p(a32), S, _p(inv_scale), _p(alimb2), 1) == 0
    assert emu.emu_lt_quant(_p(a32), S, _p(inv_scale), _p(alimb), 0) == 0
    l2 = alimb2.view(np.int8).reshape(S, 2, 4)[:, :, :3].astype(np.int64)
    l1 = alimb.view(np.int8).reshape(S, 2, 4)[:, :, :3].astype(np.int64)
    v1 = (l1[:, :, 0] * 128 + l1[:, :, 1]) * 128 + l1[:, :, 2]
    v2 = (l2[:, :, 0] * 128 + l2[:, :, 1]) * 128 + l2[:, :, 2]
    both = (v1[:, 0] + 1j * v1[:, 1]) * float(inv_scale[0]) + (v2[:, 0] + 1j * v2[:, 1]) * float(inv_scale[1])
    assert np.abs(v2).max() <= 2 ** 20 and abs(float(inv_scale[1]) * 2 ** 21 / float(inv_scale[0]) - 1) < 1e-6
    assert np.max(np.abs(both - a32)) <= 1e-6 * float(inv_scale[0])
    limbs = alimb.view(np.int8).reshape(S, 2, 4)[:, :, :3].astype(np.int64)     # (S, re/im, limb)
    assert np.abs(limbs[:, :, 1:]).max() <= 64
    vq = (limbs[:, :, 0] * 128 + limbs[:, :, 1]) * 128 + limbs[:, :, 2]
    aq = (vq[:, 0] + 1j * vq[:, 1]) * float(inv_scale[0])
    assert np.max(np.abs(aq - a32)) <= 0.75 * float(inv_scale[0]) * np.sqrt(2)
    # materialised operands
    A = np.zeros((P, Mhi, 2, Kp), dtype=np.int8)
    assert emu.emu_lt_agen(_p(hhi), _p(e), S, Se, b1, P, Mhi, Kp, _p(A), spread) == 0
    Bq = np.zeros((3, Nlo, Kp), dtype=np.int8)
    assert emu.emu_lt_bgen(_p(hlo), _p(alimb), S, b2, Nlo, Kp, _p(Bq), spread) == 0
    assert not A[..., 2 * S:].any() and not Bq[..., 2 * S:].any()          # K padding is zero
    assert set(np.unique(A)) <= {-1, 0, 1}
    pairs = A[..., :2 * S].reshape(P, Mhi, 2, S, 2)
    assert (np.count_nonzero(pairs, axis=-1) == 1).all()                   # exactly one zero per (Re, Im) byte pair: 2:4 sparse
    # exact integer GEMM + limb recombination == the samples of the lattice, straight from the definition
    acc = np.einsum("pmrk,lnk->lpmrn", A.astype(np.int64), Bq.astype(np.int64))
    val = ((acc[0] * 128 + acc[1]) * 128 + acc[2]) * float(inv_scale[0])   # (P, Mhi, 2, Nlo)
    got = (val[:, :, 0, :] + 1j * val[:, :, 1, :]).reshape(P, Mhi * Nlo)
    dig = orc.query_digits(M, D, q)                                        # (P, n, B)
    want = np.stack([orc.synth_eval_digits(dig[p].T, locq, aq, q) for p in range(P)])
    assert np.max(np.abs(got - want)) <= 1e-9 * S
    exact = np.stack([orc.synth_eval_digits(dig[p].T, locq, a, q) for p in range(P)])
    assert np.max(np.abs(got - exact)) <= 2 * S * float(inv_scale[0])     # quantisation of the strengths only
    # packed phase tables == the rotations the materialised operand was built from
    Ttab = np.zeros((Mhi, Tw), dtype=np.uint32)
    Etab = np.zeros((P, Tw), dtype=np.uint32)
    assert emu.emu_lt_ttab(_p(hhi), S, b1, Tw, Mhi, _p(Ttab), spread) == 0
    assert emu.emu_lt_etab(_p(e), S, Se, P, Tw, _p(Etab)) == 0
    f = np.arange(16)
    Tf = ((Ttab[:, :, None] >> (2 * f)) & 3).reshape(Mhi, -1)[:, :S]       # (Mhi, S)
    Ef = ((Etab[:, :, None] >> (2 * f)) & 3).reshape(P, -1)[:, :S]
    rot = (Tf[None, :, :] + Ef[:, None, :]) % 4                            # (P, Mhi, S)
    er, ei = np.array([1, 0, -1, 0])[rot], np.array([0, 1, 0, -1])[rot]
    assert np.array_equal(pairs[:, :, 0, :, 0], er) and np.array_equal(pairs[:, :, 0, :, 1], -ei)
    assert np.array_equal(pairs[:, :, 1, :, 0], ei) and np.array_equal(pairs[:, :, 1, :, 1], er)
    w = 1 if q == 2 else 2                                                  # bits per lattice digit in the row index
    lhi_d = np.stack([(np.arange(Mhi) >> (w * (b1 - 1 - i))) & (q - 1) for i in range(b1)])   # (b1, Mhi) MSB first
    assert np.array_equal(Tf, (lhi_d.T @ H[:b1]) % 4)


@pytest.mark.parametrize("n,b,S,P,seed", [(12, 5, 37, 4, 0), (30, 6, 61, 3, 1), (9, 4, 8, 2, 2)])
def test_emulated_q3_lattice_operands_reproduce_the_samples(emu, n, b, S, P, seed):
    """q = 3: the Z[w] operands (lt_agen3 / lt_bgen3) with an exact integer matrix product in place of the tensor-core GEMM,
    the two real launches and lt_combine3 reproduce the samples of the lattice {M l + d_p} from the definition."""
    q = 3
    rng = np.random.default_rng(seed)
    M, D = rng.integers(0, q, (n, b)), rng.integers(0, q, (P, n))
    locq = rng.integers(0, q, (n, S))
    a = rng.uniform(0.3, 2, S) * np.exp(1j * rng.uniform(0, 2 * np.pi, S))
    b1, b2 = b // 2, b - b // 2
    Mhi, Nlo = q ** b1, q ** b2
    ld = max(32, (n + 31) // 32 * 32)
    Kp = (2 * S + 127) // 128 * 128
    Se = (S + 3) & ~3
    M8, D8 = np.ascontiguousarray(M, dtype=np.int8), np.ascontiguousarray(D, dtype=np.int8)
    loc = np.zeros((S, ld), dtype=np.int8)
    loc[:, :n] = locq.T
    hhi, hlo = np.zeros(S, dtype=np.uint32), np.zeros(S, dtype=np.uint32)
    e = np.zeros((P, Se), dtype=np.uint8)
    assert emu.emu_lt_prep(_p(M8), _p(D8), _p(loc), S, Se, n, b, b1, P, ld, _p(hhi), _p(hlo), _p(e), q) == 0
    H = (M.T @ locq) % q
    assert np.array_equal(hhi, sum(H[i].astype(np.int64) << (2 * (b1 - 1 - i)) for i in range(b1)))
    assert np.array_equal(hlo, sum(H[b1 + i].astype(np.int64) << (2 * (b2 - 1 - i)) for i in range(b2)))
    assert np.array_equal(e[:, :S], (D @ locq) % q)
    a32 = np.ascontiguousarray(a.astype(np.complex64))
    inv_scale = np.zeros(2, dtype=np.float32)
    alimb = np.zeros((S, 2), dtype=np.int32)
    assert emu.emu_lt_quant(_p(a32), S, _p(inv_scale), _p(alimb), 0) == 0
    limbs = alimb.view(np.int8).reshape(S, 2, 4)[:, :, :3].astype(np.int64)
    vq = (limbs[:, :, 0] * 128 + limbs[:, :, 1]) * 128 + limbs[:, :, 2]
    aq = (vq[:, 0] + 1j * vq[:, 1]) * float(inv_scale[0])
    A = np.zeros((P, Mhi, 2, Kp), dtype=np.int8)
    assert emu.emu_lt_agen3(_p(hhi), _p(e), S, Se, b1, P, Mhi, Kp, _p(A)) == 0
    assert set(np.unique(A)) <= {-1, 0, 1} and not A[..., 2 * S:].any()
    planes = []
    for part in (0, 1):
        Bq = np.zeros((3, Nlo, Kp), dtype=np.int8)
        assert emu.emu_lt_bgen3(_p(hlo), _p(alimb), S, b2, Nlo, Kp, part, _p(Bq)) == 0
        assert not Bq[..., 2 * S:].any() and np.abs(Bq).max() <= 64
        acc = np.einsum("pmrk,lnk->lpmrn", A.astype(np.int64), Bq.astype(np.int64))
        val = ((acc[0] * 128 + acc[1]) * 128 + acc[2]) * float(inv_scale[0])        # (P, Mhi, 2 = (1, w) parts, Nlo)
        planes.append(np.ascontiguousarray(np.moveaxis(val, 2, 3).reshape(-1, 2).astype(np.float32)))   # float2 = (c_1, c_w)
    N = P * Mhi * Nlo
    out = np.zeros((N, 2), dtype=np.float32)
    assert emu.emu_lt_combine3(_p(planes[0]), _p(planes[1]), N, _p(out)) == 0
    got = (out[:, 0] + 1j * out[:, 1]).reshape(P, Mhi * Nlo)
    dig = orc.query_digits(M, D, q)
    want = np.stack([orc.synth_eval_digits(dig[p].T, locq, aq, q) for p in range(P)])
    assert np.max(np.abs(got - want)) <= 2e-6 * np.sqrt(S) * float(np.max(np.abs(a)))        # fp32 planes + combination


@pytest.mark.parametrize("q,n,b,S,P,seed", [(5, 10, 4, 37, 3, 0), (7, 8, 3, 29, 4, 1), (5, 30, 5, 16, 2, 2), (3, 12, 5, 37, 4, 3),
                                            (7, 20, 4, 11, 2, 4)])
def test_emulated_odd_prime_lattice_operands_reproduce_the_samples(emu, q, n, b, S, P, seed):
    """Odd primes (q = 5, 7; q = 3 through the same generic kernels): Z[w] operands with d = q - 1 rows per l_hi and d bytes
    per support element (lt_agenq / lt_bgenq), an exact integer matrix product in place of the tensor-core GEMM, the GEMM's
    output layout (row pairs = float2, planes [p][l_hi][j][l_lo]) and lt_combineq reproduce the samples of the lattice."""
    d, fw = q - 1, (2 if q <= 4 else 3)
    rng = np.random.default_rng(seed)
    M, D = rng.integers(0, q, (n, b)), rng.integers(0, q, (P, n))
    locq = rng.integers(0, q, (n, S))
    a = rng.uniform(0.3, 2, S) * np.exp(1j * rng.uniform(0, 2 * np.pi, S))
    b1, b2 = b // 2, b - b // 2
    Mhi, Nlo = q ** b1, q ** b2
    ld = max(32, (n + 31) // 32 * 32)
    Kp = (d * S + 127) // 128 * 128
    Se = (S + 3) & ~3
    M8, D8 = np.ascontiguousarray(M, dtype=np.int8), np.ascontiguousarray(D, dtype=np.int8)
    loc = np.zeros((S, ld), dtype=np.int8)
    loc[:, :n] = locq.T
    hhi, hlo = np.zeros(S, dtype=np.uint32), np.zeros(S, dtype=np.uint32)
    e = np.zeros((P, Se), dtype=np.uint8)
    assert emu.emu_lt_prep(_p(M8), _p(D8), _p(loc), S, Se, n, b, b1, P, ld, _p(hhi), _p(hlo), _p(e), q) == 0
    H = (M.T @ locq) % q
    assert np.array_equal(hhi, sum(H[i].astype(np.int64) << (fw * (b1 - 1 - i)) for i in range(b1)))
    assert np.array_equal(hlo, sum(H[b1 + i].astype(np.int64) << (fw * (b2 - 1 - i)) for i in range(b2)))
    a32 = np.ascontiguousarray(a.astype(np.complex64))
    inv_scale = np.zeros(2, dtype=np.float32)
    alimb = np.zeros((S, 2), dtype=np.int32)
    assert emu.emu_lt_quant(_p(a32), S, _p(inv_scale), _p(alimb), 0) == 0
    limbs = alimb.view(np.int8).reshape(S, 2, 4)[:, :, :3].astype(np.int64)
    vq = (limbs[:, :, 0] * 128 + limbs[:, :, 1]) * 128 + limbs[:, :, 2]
    aq = (vq[:, 0] + 1j * vq[:, 1]) * float(inv_scale[0])
    A = np.zeros((P, Mhi, d, Kp), dtype=np.int8)
    assert emu.emu_lt_agenq(q, _p(hhi), _p(e), S, Se, b1, P, Mhi, Kp, _p(A)) == 0
    assert set(np.unique(A)) <= {-1, 0, 1} and not A[..., d * S:].any()
    # the d x d blocks are the multiplication matrices of w^t: column c = coordinates of w^(t + c)
    lhi_d = np.stack([(np.arange(Mhi) // q ** (b1 - 1 - i)) % q for i in range(b1)])          # (b1, Mhi) MSB first
    t = ((lhi_d.T @ H[:b1])[None, :, :] + ((D @ locq) % q)[:, None, :]) % q                   # (P, Mhi, S)
    blocks = A[..., :d * S].reshape(P, Mhi, d, S, d)                                          # [p][l_hi][r][s][c]
    m = (t[:, :, None, :, None] + np.arange(d)[None, None, None, None, :]) % q
    want_blocks = (m == np.arange(d)[None, None, :, None, None]).astype(np.int8) - (m == d).astype(np.int8)
    assert np.array_equal(blocks, want_blocks)
    planes = []
    for part in (0, 1):
        Bq = np.zeros((3, Nlo, Kp), dtype=np.int8)
        assert emu.emu_lt_bgenq(q, _p(hlo), _p(alimb), S, b2, Nlo, Kp, part, _p(Bq)) == 0
        assert not Bq[..., d * S:].any() and np.abs(Bq[1:]).max() <= 64
        acc = np.einsum("pmrk,lnk->lpmrn", A.astype(np.int64), Bq.astype(np.int64))
        val = ((acc[0] * 128 + acc[1]) * 128 + acc[2]) * float(inv_scale[0])        # (P, Mhi, d coordinates, Nlo)
        # what the GEMM kernel writes: rows (2 j, 2 j + 1) of one l_hi -> one float2 at [(p, l_hi)][j][l_lo]
        pl = val.reshape(P * Mhi, d // 2, 2, Nlo).transpose(0, 1, 3, 2)
        planes.append(np.ascontiguousarray(pl.astype(np.float32)))
    out = np.zeros((P * Mhi * Nlo, 2), dtype=np.float32)
    assert emu.emu_lt_combineq(q, _p(planes[0]), _p(planes[1]), P * Mhi, Nlo, _p(out)) == 0
    got = (out[:, 0] + 1j * out[:, 1]).reshape(P, Mhi * Nlo)
    dig = orc.query_digits(M, D, q)
    want = np.stack([orc.synth_eval_digits(dig[p].T, locq, aq, q) for p in range(P)])
    assert np.max(np.abs(got - want)) <= 4e-6 * np.sqrt(S) * float(np.max(np.abs(a)))        # fp32 planes + combination


def test_ts_expand_variants_match_the_definition(emu):
    """The A' expansion inside the tensor-memory GEMM kernel (16 two-bit rotations -> 16 bytes of +-1 and 8 metadata
    nibbles): the default and the opt-in PRMT variant (QSFT_LATTICE_EXPAND=1) agree with each other and with the
    definition -- Re row (er, -ei), Im row (ei, er), (er, ei) = i^r, 2:4 metadata nibble = idx0 | idx1 << 2 -- for every
    16-bit pattern in both halves of the word and for random words."""
    emu.emu_ts_expand.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(0)
    low = np.arange(1 << 16, dtype=np.uint32)
    words = np.ascontiguousarray(np.concatenate([low, low << 16, low | (low << 16), rng.integers(0, 1 << 32, 200_000, dtype=np.uint64)
                                                 .astype(np.uint32)]))
    f = np.arange(16)
    rot = (words[:, None] >> (2 * f)) & 3                                  # (N, 16)
    er, ei = np.array([1, 0, -1, 0])[rot], np.array([0, 1, 0, -1])[rot]
    for im in (0, 1):
        dense = np.stack([ei, er], axis=-1) if im else np.stack([er, -ei], axis=-1)      # (N, 16, 2) logical byte pairs
        outs = []
        for px in (0, 1):
            a4 = np.zeros((len(words), 4), dtype=np.uint32)
            e1 = np.zeros(len(words), dtype=np.uint32)
            assert emu.emu_ts_expand(_p(words), len(words), im, px, _p(a4), _p(e1)) == 0
            outs.append((a4, e1))
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
        comp = outs[0][0].view(np.int8).reshape(len(words), 16)            # one compressed byte per support element
        nib = (outs[0][1][:, None] >> (4 * np.arange(8))) & 15             # one nibble per element pair
        assert (nib & 8).all() and not (nib & 2).any()
        pos = np.stack([nib & 1, (nib >> 2) & 1], axis=-1).reshape(len(words), 16)       # index of the non-zero in each pair
        rebuilt = np.zeros_like(dense)
        np.put_along_axis(rebuilt, pos[..., None], comp[..., None].astype(dense.dtype), axis=-1)
        assert np.array_equal(rebuilt, dense)


def test_ts_phase_expand_variants(emu):
    """Table words (t, e) -> expansion of (t + e) mod 4: the opt-in variant (bit planes computed directly) == the default
    (SWAR add + expansion) == expansion of the field-wise sum, for all pairs of 8-bit patterns in every byte and random words."""
    emu.emu_ts_expand.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    emu.emu_ts_phase_expand.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(1)
    a8, b8 = np.meshgrid(np.arange(256, dtype=np.uint32), np.arange(256, dtype=np.uint32), indexing="ij")
    a8, b8 = a8.ravel(), b8.ravel()
    tw = np.concatenate([a8 * 0x01010101, a8 << 24, rng.integers(0, 1 << 32, 300_000, dtype=np.uint64).astype(np.uint32)])
    ew = np.concatenate([b8 * 0x01010101, b8 << 24, rng.integers(0, 1 << 32, 300_000, dtype=np.uint64).astype(np.uint32)])
    tw, ew = np.ascontiguousarray(tw.astype(np.uint32)), np.ascontiguousarray(ew.astype(np.uint32))
    f = np.arange(16, dtype=np.uint32)
    summed = ((((tw[:, None] >> (2 * f)) & 3) + ((ew[:, None] >> (2 * f)) & 3)) & 3)
    r16 = np.ascontiguousarray(np.bitwise_or.reduce(summed.astype(np.uint32) << (2 * f), axis=1).astype(np.uint32))
    for im in (0, 1):
        want_a, want_e = np.zeros((len(tw), 4), dtype=np.uint32), np.zeros(len(tw), dtype=np.uint32)
        assert emu.emu_ts_expand(_p(r16), len(r16), im, 0, _p(want_a), _p(want_e)) == 0
        for px in (0, 1):
            a4, e1 = np.zeros((len(tw), 4), dtype=np.uint32), np.zeros(len(tw), dtype=np.uint32)
            assert emu.emu_ts_phase_expand(_p(tw), _p(ew), len(tw), im, px, _p(a4), _p(e1)) == 0
            assert np.array_equal(a4, want_a) and np.array_equal(e1, want_e), (im, px)
