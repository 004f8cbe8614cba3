"""CPU oracle for the q-SFT transform path (query lattice -> synthetic evaluation -> q-ary DFT -> peeling).

TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg may import this module.  The product package (qsft_b200/) never does and fails
loudly when its CUDA library is missing.

This is a NumPy *restatement* of the reference algorithm (basics-lab/qsft, pure Python).  Every function cites
the reference file:line (paths relative to /root/reference) whose behaviour it reproduces.  It is written
array-at-a-time where the reference loops in Python, but makes the same decisions in the same order and
consumes the global NumPy RNG in the same order, so that the same seed yields the same Ms, Ds, support,
noise and therefore the same recovered transform.

Pinning: oracle/gen_golden.py runs the UNMODIFIED reference (through oracle/ref_shim.py) and stores its
inputs/outputs under tests/golden/*.npz; tests/test_oracle_golden.py checks this file against every one of
them.  Exception: the Reed-Solomon ("coded") delay path depends on galois==0.1.1, which is neither vendored in
the reference nor installed here -> for `RSCode` below PARITY IS UNPINNED (self-consistency tests only).
"""
from __future__ import annotations

import math
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import scipy.fft

TWO_PI = 2.0 * np.pi


# ----------------------------------------------------------------------------------------------------------
# index codecs  (qsft/utils.py:74-84, 107-108, 180-187)
# ----------------------------------------------------------------------------------------------------------
def qary_vec_to_dec(x, q):
    """Columns of x (n, N) are MSB-first base-q digit vectors -> arbitrary-precision ints (object array).
    qsft/utils.py:74-76."""
    x = np.asarray(x)
    n = x.shape[0]
    acc = np.zeros(x.shape[1:], dtype=object)
    for i in range(n):  # Horner: same value as sum_i q^(n-1-i) x_i
        acc = acc * q + x[i].astype(object)
    return acc


def dec_to_qary_vec(x, q, n):
    """Decimal indices (sequence of python ints) -> (n, N) int digit array, MSB first.  qsft/utils.py:79-84
    (object-dtype arithmetic so that indices wider than 64 bits stay exact)."""
    v = np.array([int(t) for t in x], dtype=object)
    out = np.zeros((n, len(v)), dtype=np.int64)
    for i in range(n - 1, -1, -1):
        if len(v):
            out[i] = (v % q).astype(np.int64)
            v = v // q
    return out


def qary_ints(m, q):
    """All q^m vectors of Z_q^m as columns, itertools.product order (column c <-> base-q digits of c, MSB first).
    qsft/utils.py:107-108."""
    c = np.arange(q ** m, dtype=np.int64)
    rows = [(c // (q ** (m - 1 - i))) % q for i in range(m)]
    return np.stack(rows, axis=0) if m > 0 else np.zeros((0, 1), dtype=np.int64)


def sort_qary_vecs(vecs):
    """Lexicographic row sort.  qsft/utils.py:180-183."""
    vecs = np.asarray(vecs)
    return vecs[np.lexsort(vecs.T[::-1, :])]


def calc_hamming_weight(vecs):
    """qsft/utils.py:185-187."""
    return np.sum(np.asarray(vecs) != 0, axis=1)


# ----------------------------------------------------------------------------------------------------------
# subsampling matrices and delays  (qsft/query.py)
# ----------------------------------------------------------------------------------------------------------
def get_Ms(n, b, q, num_to_get=None, method="simple"):
    """qsft/query.py:13-70.  'simple': identity block at rows b*i.., i DESCENDING; 'complex': uniform randint."""
    if num_to_get is None:
        num_to_get = max(n // b, 3)
    if method == "simple":
        if num_to_get > n // b:
            raise ValueError("When query_method is 'simple', the number of M matrices to return cannot be larger than n // b")
        out = []
        for i in reversed(range(num_to_get)):
            M = np.zeros((n, b), dtype=np.int32)
            M[b * i:b * (i + 1), :] = np.eye(b)
            out.append(M)
        return out
    if method == "complex":
        return [np.random.randint(q, size=(n, b)) for _ in range(num_to_get)]
    raise TypeError(f"unknown query_method {method!r}")  # reference: 'NoneType' object is not callable


def get_D_source(n, q, method, **kw):
    """qsft/query.py:73-92."""
    if method == "identity":
        return np.vstack((np.zeros(n), np.eye(n))).astype(int)
    if method == "random":
        return np.random.choice(q, (kw.get("num_delays"), n))
    if method == "coded":
        return np.array(RSCode(n, kw.get("t"), q).get_delay_matrix(), dtype=int)
    raise TypeError(f"unknown delays_method_source {method!r}")


def get_D(n, q, delays_method_source="random", delays_method_channel="identity", **kw):
    """qsft/query.py:116-143 (+ :95-114).  Returns a list of R arrays (P_src, n)."""
    D_src = get_D_source(n, q, delays_method_source, **kw)
    if delays_method_channel == "identity":
        return [D_src % q]
    if delays_method_channel == "nso":
        offsets = np.random.choice(q, (kw.get("num_repeat"), n))
        return [(row - D_src) % q for row in offsets]
    if delays_method_channel == "coded":
        raise NotImplementedError("One day this might be implemented")
    raise TypeError(f"unknown delays_method_channel {delays_method_channel!r}")


def get_Ms_and_Ds(n, q, **kw):
    """qsft/query.py:182-203.  The SAME D list object is reused for every M."""
    Ms = get_Ms(n, kw.get("b"), q, method=kw.get("query_method"), num_to_get=kw.get("num_subsample"))
    rest = {k: v for k, v in kw.items() if k not in ("delays_method_source", "delays_method_channel")}
    D = get_D(n, q, kw.get("delays_method_source", "random"), kw.get("delays_method_channel", "identity"), **rest)
    return Ms, [D for _ in Ms]


# ----------------------------------------------------------------------------------------------------------
# query lattice  (qsft/input_signal_subsampled.py:183-206)
# ----------------------------------------------------------------------------------------------------------
def query_digits(M, D_sub, q):
    """(P_src, n, B) digit tensor of (M l + d_p) mod q over all l in Z_q^b (itertools order)."""
    M = np.asarray(M)
    L = qary_ints(M.shape[1], q)
    ML = (M.astype(np.int64) @ L) % q
    return (ML[None, :, :] + np.asarray(D_sub, dtype=np.int64)[:, :, None]) % q


def query_indices(M, D_sub, q):
    """List of P_src object arrays (B,) of python-int decimal indices.  input_signal_subsampled.py:183-206."""
    return [qary_vec_to_dec(blk, q) for blk in query_digits(M, D_sub, q)]


# ----------------------------------------------------------------------------------------------------------
# synthetic sparse signal  (synt_exp/synt_src/synthetic_signal.py)
# ----------------------------------------------------------------------------------------------------------
def random_signal_strength_model(sparsity, a, b):
    """qsft/utils.py:161-164."""
    magnitude = np.random.uniform(a, b, sparsity)
    phase = np.random.uniform(0, TWO_PI, sparsity)
    return magnitude * np.exp(1j * phase)


def generate_signal_w(n, q, sparsity, a_min, a_max, max_weight=None):
    """Sparse support + strengths, full=False branch.  synthetic_signal.py:10-38."""
    max_weight = n if max_weight is None else max_weight
    if max_weight == n:
        locq = sort_qary_vecs(np.random.randint(q, size=(n, sparsity)).T).T
    else:
        vals = np.random.randint(q - 1, size=(max_weight, sparsity)) + 1
        pos = np.random.choice(a=n, size=(sparsity, max_weight))
        locq = np.zeros((n, sparsity), dtype=int)
        for i in range(sparsity):
            locq[pos[i, :], i] = vals[:, i]
        locq = sort_qary_vecs(locq.T).T
    strengths = random_signal_strength_model(sparsity, a_min, a_max)
    signal_w = dict(zip(list(map(tuple, locq.T)), strengths))
    return signal_w, locq, strengths


def synth_eval_digits(digits, locq, strengths, q):
    """x[m] = sum_s a_s exp(2 pi i <m, k_s> / q) for query digit rows (N, n): the reference's sampling_function
    (synthetic_signal.py:95,100-102) -- integer digits times the complex frequency matrix, np.exp, times strengths."""
    freq = (2j * np.pi / q) * np.asarray(locq)
    return np.exp(np.asarray(digits) @ freq) @ strengths


def synth_subsample(query_idx, locq, strengths, q, n, batch=10000, threads=None):
    """SyntheticSubsampledSignal.subsample: decimal indices -> complex128 values, batches of <= 10000 queries
    (synthetic_signal.py:108-118; the reference forks a process pool per call, here a thread pool)."""
    query_idx = list(query_idx)
    if not query_idx:
        return np.zeros(0, dtype=complex)
    nb = len(query_idx) // batch + 1          # np.array_split(query_indices, len // batch + 1)
    bounds = np.linspace(0, len(query_idx), nb + 1).astype(int)
    chunks = [query_idx[bounds[i]:bounds[i + 1]] for i in range(nb)]

    def work(ch):
        return synth_eval_digits(dec_to_qary_vec(ch, q, n).T, locq, strengths, q) if len(ch) else np.zeros(0, complex)

    threads = threads or os.cpu_count() or 1
    if threads == 1 or nb == 1:
        parts = [work(c) for c in chunks]
    else:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(work, chunks))
    return np.concatenate(parts)


def gwht(x, q, n):
    """n-dimensional length-q DFT with forward 1/q^n scaling.  qsft/utils.py:31-36."""
    return (scipy.fft.fftn(np.reshape(x, [q] * n)) / (q ** n)).reshape(q ** n)


def closed_form_bins(M, D_sub, locq, strengths, q):
    """Identity: U_p[j] = sum_{s: M^T k_s = digits(j)} a_s w^{<d_p, k_s>}.  Verification only."""
    M = np.asarray(M, dtype=np.int64)
    b = M.shape[1]
    h = (M.T @ locq) % q
    j = np.zeros(h.shape[1], dtype=np.int64)
    for i in range(b):
        j = j * q + h[i]
    ph = (np.asarray(D_sub, dtype=np.int64) @ locq) % q
    U = np.zeros((len(D_sub), q ** b), dtype=complex)
    for p in range(len(D_sub)):
        np.add.at(U[p], j, strengths * np.exp(1j * TWO_PI / q * ph[p]))
    return U


class OracleSignal:
    """SubsampledSignal + SyntheticSubsampledSignal restated (input_signal_subsampled.py:47-155, 225-269;
    synthetic_signal.py:88-130).  Us[c][r][b'] is an array (P_src, q^b')."""

    def __init__(self, n, q, query_args, locq, strengths, noise_sd=0.0, signal_w=None, Ms=None, Ds=None,
                 use_closed_form=False, threads=None):
        self.n, self.q, self.N = n, q, q ** n
        self.query_args = query_args
        self.b = query_args.get("b")
        self.all_bs = query_args.get("all_bs", [self.b])
        self.num_subsample = query_args.get("num_subsample")
        query_args.setdefault("num_repeat", 1)
        self.num_repeat = query_args["num_repeat"]
        self.locq, self.strengths, self.noise_sd = np.asarray(locq), np.asarray(strengths), noise_sd
        self.signal_w = signal_w
        if Ms is None:
            Ms, Ds = get_Ms_and_Ds(n, q, **query_args)
        self.Ms, self.Ds = Ms, Ds
        self.Us = [[{} for _ in self.Ds[i]] for i in range(len(self.Ms))]
        self.samples = [[None for _ in self.Ds[i]] for i in range(len(self.Ms))]
        for i, M in enumerate(self.Ms):
            for j, D_sub in enumerate(self.Ds[i]):
                if use_closed_form:  # verification shortcut, full b only
                    self.Us[i][j][self.b] = closed_form_bins(M, D_sub, self.locq, self.strengths, q)
                    continue
                idx = query_indices(M, D_sub, q)
                smp = np.stack([self.subsample(row, threads) for row in idx])
                self.samples[i][j] = smp
                for bb in self.all_bs:
                    stride = q ** (self.b - bb)
                    self.Us[i][j][bb] = np.stack([gwht(row[::stride], q, bb) for row in smp])

    def subsample(self, query_indices_, threads=None):
        return synth_subsample(query_indices_, self.locq, self.strengths, self.q, self.n, threads=threads)

    def get_source_parity(self):
        return self.Ds[0][0].shape[0]

    def get_MDU(self, ret_num_subsample, ret_num_repeat, b, trans_times=False):
        """input_signal_subsampled.py:225-262 then the synthetic noise wrapper synthetic_signal.py:120-130."""
        if not (ret_num_subsample <= self.num_subsample and ret_num_repeat <= self.num_repeat and b <= self.b):
            raise ValueError("There are not enough Ms or Ds.")
        sub_idx = np.random.choice(self.num_subsample, ret_num_subsample, replace=False)
        del_idx = np.random.choice(self.num_repeat, ret_num_repeat, replace=False)
        Ms_r, Ds_r, Us_r = [], [], []
        for i in sub_idx:
            Ms_r.append(self.Ms[i][:, :b])
            Ds_r.append([self.Ds[i][j] for j in del_idx])
            Us_r.append([self.Us[i][j][b] for j in del_idx])
        nu = self.noise_sd / np.sqrt(2 * self.q ** b)
        for i in range(len(Us_r)):
            for j in range(len(Us_r[i])):
                size = Us_r[i][j].shape
                noise = np.random.normal(0, nu, size=size + (2,))
                Us_r[i][j] = Us_r[i][j] + (noise[..., 0] + 1j * noise[..., 1])
        if trans_times:
            return Ms_r, Ds_r, Us_r, [[0.0 for _ in row] for row in Us_r]
        return Ms_r, Ds_r, Us_r


# ----------------------------------------------------------------------------------------------------------
# singleton detection  (qsft/reconstruct.py)
# ----------------------------------------------------------------------------------------------------------
def detect_noiseless(cols, q):
    """cols (P, nb) -> symbols (P-1, nb).  reconstruct.py:12-31 (np.round = half-to-even)."""
    ang = np.angle(cols)
    return (np.round(q * (ang[1:] - ang[0]) / TWO_PI).astype(int)) % q


def detect_nso1(cols, q, p1):
    """cols (R*p1, nb) -> symbols (p1-1, nb).  reconstruct.py:100-113 (argmin over q+1 roots, first minimum)."""
    roots = TWO_PI / q * np.arange(q + 1)
    zero = cols[0::p1]
    out = np.zeros((p1 - 1, cols.shape[1]), dtype=int)
    for i in range(1, p1):
        ang = np.angle(np.mean(zero * np.conjugate(cols[i::p1]), axis=0)) % TWO_PI
        out[i - 1] = np.abs(roots[:, None] - ang[None, :]).argmin(axis=0) % q
    return out


def angle_q(x, q):
    """Quantised angle: index of the q-th root of unity nearest to x.  qsft/utils.py:104-105 (float floor divisions)."""
    return (((np.angle(x) % TWO_PI // (np.pi / q)) + 1) // 2) % q


def detect_nso2(cols, q, p1):
    """Hard-decision NSO: cols (R*p1, nb) -> symbols (p1-1, nb).  reconstruct.py:116-129: every repeat votes with its
    quantised phase difference, the votes are averaged AS NUMBERS (not circularly), np.round is half-to-even, and the
    float result is truncated into the int array."""
    a0 = angle_q(cols[0::p1], q)
    out = np.zeros((p1 - 1, cols.shape[1]), dtype=int)
    for i in range(1, p1):
        a = angle_q(cols[i::p1], q)
        out[i - 1] = (np.round(np.mean((a0 - a) % q, axis=0)) % q).astype(int)
    return out


def detect_mle(col, selection, S_slice):
    """reconstruct.py:54-84 for ONE column (P,): least-squares amplitude for every candidate signature (columns of
    S_slice, (P, K)), residual 2-norms, first minimum.  Returns (selection[k_sel], S_slice[:, k_sel], k_sel).
    The reference can only be called directly (QSFT.transform never passes selection / S_slice)."""
    P = S_slice.shape[0]
    alphas = 1 / P * np.dot(np.conjugate(S_slice).T, col)
    residuals = np.linalg.norm(col - (alphas * S_slice).T, ord=2, axis=1)
    k_sel = int(np.argmin(residuals))
    return selection[k_sel], S_slice[:, k_sel], k_sel


def singleton_detection(cols, q, method_channel, method_source, source_parity, source_decoder=None,
                        nso_subtype="nso1"):
    """reconstruct.py:132-168 for a batch of columns.  Returns k (n, nb).  `mle` needs per-column candidate lists and
    is exposed separately (detect_mle)."""
    if method_channel == "identity":
        sym = detect_noiseless(cols, q)
    elif method_channel == "nso":
        sym = detect_nso1(cols, q, source_parity) if nso_subtype == "nso1" else detect_nso2(cols, q, source_parity)
    else:
        raise NotImplementedError("mle is unreachable from QSFT.transform in the reference (SURVEY a15): use detect_mle")
    if method_source == "identity":
        return sym
    if method_source == "coded":
        ks = [np.array(source_decoder(list(sym[:, c]))[0][0, :], dtype=np.int32) for c in range(sym.shape[1])]
        return np.stack(ks, axis=1) if ks else np.zeros((0, 0), dtype=np.int32)
    raise TypeError(method_source)


# ----------------------------------------------------------------------------------------------------------
# peeling decoder  (qsft/qsft.py:56-281)
# ----------------------------------------------------------------------------------------------------------
def transform(signal, num_subsample, num_repeat, b, reconstruct_method_source="identity",
              reconstruct_method_channel="identity", source_decoder=None, report=False, sort=False,
              cutoff=None, trace=None, nso_subtype="nso1"):
    """QSFT.transform.  `trace`, if a list, receives per-round dicts (singletons / multitons / peeled) for tests.
    `nso_subtype` is hard-coded to "nso1" in the reference (qsft.py:171); "nso2" selects reconstruct.py:116-129."""
    q, n = signal.q, signal.n
    omega = np.exp(2j * np.pi / q)
    Ms, Ds, Us, _ = signal.get_MDU(num_subsample, num_repeat, b, trans_times=True)   # qsft.py:111
    Us = np.array([np.vstack(u) for u in Us])                                          # (C, P, B)   :114-120
    Ds = [np.vstack(d) for d in Ds]
    C, P, B = Us.shape
    if cutoff is None:
        cutoff = 1e-9 + 1.5 * (signal.noise_sd ** 2) / (q ** b)                       # :124-126
    p1 = signal.get_source_parity()
    Jdig = qary_ints(b, q)                                                             # (b, B)
    weights = np.array([q ** (b - 1 - i) for i in range(b)], dtype=np.int64)
    result = []
    num_peeling, it, cont = 0, 0, True
    while cont and num_peeling < q ** n and it < 15:                                    # :151
        it += 1
        singles = []          # (i, j, k, rho) in (i, j) order
        n_multi = 0
        for i in range(C):
            U, M, D = Us[i], Ms[i], Ds[i]
            energy = np.sum(np.abs(U) ** 2, axis=0)                                    # ||col||^2   :164
            js = np.nonzero(energy > cutoff * P)[0]
            if len(js) == 0:
                continue
            cols = U[:, js]
            K = singleton_detection(cols, q, reconstruct_method_channel, reconstruct_method_source, p1,
                                    source_decoder, nso_subtype)                                   # :165-173
            sig = omega ** (D @ K)                                                     # :174
            rho = np.sum(np.conjugate(sig) * cols, axis=0) / P                         # :175
            res = np.sum(np.abs(cols - rho * sig) ** 2, axis=0)                        # :176
            match = np.all((M.T @ K) % q == Jdig[:, js], axis=0)                       # :178-179
            ok = match & ~(res > cutoff * P)                                           # :183
            n_multi += int(np.sum(~ok))
            for c in np.nonzero(ok)[0]:
                singles.append((i, int(js[c]), K[:, c].copy(), rho[c]))
        if n_multi == 0 or len(singles) == 0:                                          # :204-205
            cont = False
        ball_values = {}
        for (i, j, k, rho) in singles:                                                 # :209-217
            ball_values[tuple(int(v) for v in k)] = rho      # last (i, j) wins
            result.append((tuple(int(v) for v in k), rho))
        if trace is not None:
            trace.append({"singletons": [(i, j, tuple(int(v) for v in k), rho) for (i, j, k, rho) in singles],
                          "n_multitons": n_multi, "balls": dict(ball_values)})
        for ball, val in ball_values.items():                                          # :223-241
            num_peeling += 1
            k = np.array(ball, dtype=np.int64)
            for l in range(C):
                jl = int(((Ms[l].T @ k) % q) @ weights)
                Us[l][:, jl] -= val * omega ** (Ds[l] @ k)
    gw, cnt = {}, {}
    for k, v in result:                                                                # :247-255
        if k in cnt:
            gw[k] = (gw[k] * cnt[k] + v) / (cnt[k] + 1)
            cnt[k] += 1
        else:
            gw[k], cnt[k] = v, 1
    if not report:
        return gw
    loc = list(gw.keys())
    if loc:
        if sort:
            loc = sort_qary_vecs(loc)
        hw = calc_hamming_weight(loc)
        avg_w, max_w = np.mean(hw), np.max(hw)
    else:
        loc, avg_w, max_w = [], 0, 0
    return {"gwht": gw, "runtime": 0.0, "n_samples": C * P * B, "locations": loc,
            "avg_hamming_weight": avg_w, "max_hamming_weight": max_w, "rounds": it}


def nmse(gwht_est, signal_w):
    """quick_example.py:75-80."""
    diff = dict(signal_w)
    for k, v in gwht_est.items():
        diff[k] = diff.get(k, 0) - v
    return float(np.sum(np.abs(list(diff.values())) ** 2) / np.sum(np.abs(list(signal_w.values())) ** 2))


# ----------------------------------------------------------------------------------------------------------
# Reed-Solomon delays / syndrome decoder  (qsft/ReedSolomon.py, on galois==0.1.1 -- NOT available here)
# PARITY UNPINNED: field construction (primitive polynomial choice) follows what galois 0.1.x is believed to
# do (smallest primitive polynomial in integer order, primitive element x); decoded k for weight <= t error
# patterns is independent of that choice (unique decoding).
# ----------------------------------------------------------------------------------------------------------
def _is_prime(p):
    return p >= 2 and all(p % d for d in range(2, int(math.isqrt(p)) + 1))


class GFext:
    """GF(p^s), elements = ints whose base-p digits are polynomial coefficients (degree-descending when
    written MSB first), built from the smallest primitive polynomial; exp/log tables."""

    def __init__(self, p, s):
        if not _is_prime(p):
            raise NotImplementedError("q is not a prime number")
        self.p, self.s, self.order = p, s, p ** s
        # galois 0.1.x RS default = matlab_primitive_poly(p, s): the lexicographically-minimal primitive polynomial,
        # except GF(2^7) -> x^7 + x^3 + 1 (also GF(2^14), GF(2^16): beyond n <= 128).  [galois docs, from memory]
        first = [9] if (p, s) == (2, 7) else []
        for low in first + list(range(self.order)):  # monic x^s + low
            tab = self._try_poly(low)
            if tab is not None:
                self.poly_low = low
                self.exp = tab
                break
        else:
            raise RuntimeError("no primitive polynomial found")
        self.log = np.zeros(self.order, dtype=np.int64)
        for e, v in enumerate(self.exp[: self.order - 1]):
            self.log[v] = e

    def _mulx(self, v, low):
        """v * x mod (x^s + low) over GF(p)."""
        p, s = self.p, self.s
        top = v // (p ** (s - 1))
        v = (v % (p ** (s - 1))) * p
        if top:
            out, w = 0, 1
            for _ in range(s):  # v - top*low coefficientwise
                out += ((v // w % p - top * (low // w % p)) % p) * w
                w *= p
            v = out
        return v

    def _try_poly(self, low):
        if low % self.p == 0:
            return None
        seen, v, tab = set(), 1, []
        for _ in range(self.order - 1):
            if v in seen:
                return None
            seen.add(v)
            tab.append(v)
            v = self._mulx(v, low)
        if v != 1:
            return None
        return np.array(tab + tab, dtype=np.int64)

    def add(self, a, b):
        p, out, w = self.p, 0, 1
        for _ in range(self.s):
            out += ((a // w + b // w) % p) * w
            w *= p
        return out

    def neg(self, a):
        p, out, w = self.p, 0, 1
        for _ in range(self.s):
            out += ((-(a // w)) % p) * w
            w *= p
        return out

    def sub(self, a, b):
        return self.add(a, self.neg(b))

    def mul(self, a, b):
        if a == 0 or b == 0:
            return 0
        return int(self.exp[self.log[a] + self.log[b]])

    def inv(self, a):
        return int(self.exp[(self.order - 1 - self.log[a]) % (self.order - 1)])

    def pow_alpha(self, e):
        return int(self.exp[e % (self.order - 1)])

    def to_vec(self, a):
        """degree-descending coefficient vector (galois .vector())."""
        return [(a // (self.p ** (self.s - 1 - i))) % self.p for i in range(self.s)]

    def from_vec(self, v):
        out = 0
        for c in v:
            out = out * self.p + int(c) % self.p
        return out


class RSCode:
    """qsft/ReedSolomon.py:7-74 with our own field arithmetic (narrow-sense RS, c = 1, roots alpha^1..alpha^2t)."""

    def __init__(self, n, t, q):
        s = math.ceil(math.log(n) / math.log(q))                                       # :20
        if n > q ** s - 1:                                                              # :21-22
            s += 1
        self.s, self.ns, self.t, self.q = s, n, t, q
        self.nt = q ** s - 1
        self.c = 1
        self.F = GFext(q, s)

    def get_parity_length(self):
        return 2 * self.t * self.s

    def H_entry(self, j, i_full):
        """H[j, i] = (alpha^(c+j))^(nt-1-i)."""
        return self.F.pow_alpha((self.c + j) * (self.nt - 1 - i_full))

    def get_delay_matrix(self):
        """:50-65.  D (2ts+1, n) over Z_q, row 0 zero."""
        D = np.zeros((self.get_parity_length() + 1, self.ns), dtype=np.int64)
        for i in range(self.ns):
            i_full = self.nt - self.ns + i
            for j in range(2 * self.t):
                D[self.s * j + 1:self.s * (j + 1) + 1, i] = self.F.to_vec(self.H_entry(j, i_full))
        return D

    def syndrome_decode(self, syndrome):
        """:26-48.  2ts symbols of Z_q -> (k as (1, n) array, n_errors); failure -> zeros, -1."""
        F, t = self.F, self.t
        S = [F.from_vec(syndrome[self.s * i:self.s * (i + 1)]) for i in range(2 * t)]
        zero = np.zeros((1, self.ns), dtype=np.int64)
        if not any(S):
            return zero, 0
        # Berlekamp-Massey: Lambda(x) = 1 + L1 x + ... with sum_i Lambda_i S_{r-i} = 0
        Lam, Bp, L, m, bb = [1], [1], 0, 1, 1
        for r in range(2 * t):
            d = S[r]
            for i in range(1, L + 1):
                if i < len(Lam):
                    d = F.add(d, F.mul(Lam[i], S[r - i]))
            if d == 0:
                m += 1
                continue
            coef = F.mul(d, F.inv(bb))
            new = list(Lam) + [0] * max(0, len(Bp) + m - len(Lam))
            for i, bv in enumerate(Bp):
                new[i + m] = F.sub(new[i + m], F.mul(coef, bv))
            if 2 * L <= r:
                Bp, bb, L, m = list(Lam), d, r + 1 - L, 1
            else:
                m += 1
            Lam = new
        while len(Lam) > 1 and Lam[-1] == 0:
            Lam.pop()
        deg = len(Lam) - 1
        if deg != L or deg > t or deg == 0:
            return zero, -1
        # Chien search over the n retained positions: locator X = alpha^(nt-1-i_full), root X^-1
        pos = []
        for i in range(self.ns):
            e = self.nt - 1 - (self.nt - self.ns + i)     # = ns-1-i
            xinv = F.pow_alpha(-e)
            acc, pw = 0, 1
            for cf in Lam:
                acc = F.add(acc, F.mul(cf, pw))
                pw = F.mul(pw, xinv)
            if acc == 0:
                pos.append((i, e))
        if len(pos) != deg:
            return zero, -1
        # Forney: Omega = S(x) Lambda(x) mod x^2t, e_l = -X^(1-c) Omega(X^-1)/Lambda'(X^-1), c = 1
        Om = [0] * (2 * t)
        for a in range(2 * t):
            for i2, cf in enumerate(Lam):
                if a - i2 >= 0:
                    Om[a] = F.add(Om[a], F.mul(cf, S[a - i2]))
        k = np.zeros((1, self.ns), dtype=np.int64)
        for (i, e) in pos:
            xinv = F.pow_alpha(-e)
            num, pw = 0, 1
            for cf in Om:
                num = F.add(num, F.mul(cf, pw))
                pw = F.mul(pw, xinv)
            den, pw = 0, 1
            for d1 in range(1, len(Lam)):          # formal derivative: d1 * Lam[d1] x^(d1-1)
                term = 0
                for _ in range(d1 % self.q):
                    term = F.add(term, Lam[d1])
                den = F.add(den, F.mul(term, pw))
                pw = F.mul(pw, xinv)
            if den == 0:
                return zero, -1
            # S_j = sum_l e_l X_l^(j+1)  =>  e_l = - Omega(X^-1) / (X * Lambda'(X^-1)) * X ... derived below
            val = F.mul(num, F.inv(den))
            val = F.neg(val)
            # with S(x) = sum_j S_j x^j and S_j = sum e_l X_l^(j+1):  Omega(X_l^-1) = e_l X_l prod_{m!=l}(1 - X_m/X_l)
            # and Lambda'(X_l^-1) = -X_l prod_{m!=l}(1 - X_m/X_l)  =>  e_l = -Omega/Lambda'
            if val >= self.q:       # error value must lie in the prime subfield (k_i in Z_q)
                return zero, -1
            k[0, i] = val
        return k, deg


def get_reed_solomon_dec(n, t_max, q):
    """qsft/query.py:231-243 (the reference's prime list also contains 15; we require an actual prime)."""
    if q in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29):
        return RSCode(n, t_max, q).syndrome_decode
    raise NotImplementedError("q is not a prime number under 30!")
