"""Reed-Solomon delay matrices and syndrome decoding over GF(q^s), q prime.

Host-side mirror of the reference's qsft/ReedSolomon.py (which subclasses galois.ReedSolomon; galois is not a
dependency here, the field arithmetic is implemented below).  The host class produces the delay matrix D and the
GF log/antilog tables; singleton decoding during peeling runs on the GPU (csrc/k4_peel.cu: rs_decode), the
`syndrome_decode` method here is the API-compatible host entry (get_reed_solomon_dec) for callers that want it.

Field construction: lexicographically-minimal primitive polynomial x^s + ... (MATLAB's default; GF(2^7) uses
x^7 + x^3 + 1), primitive element x (what galois 0.1.x is believed to use through matlab_primitive_poly; decoded
supports do not depend on this choice, the D matrix does -- see DESIGN.md).
"""
from __future__ import annotations

import math

import numpy as np


def _is_prime(p):
    return p >= 2 and all(p % d for d in range(2, math.isqrt(p) + 1))


class GaloisField:
    """GF(p^s); an element is the integer whose base-p digits (MSB = highest degree) are its coefficients."""

    def __init__(self, p: int, s: int):
        if not _is_prime(p):
            raise NotImplementedError("q is not a prime number under 30!")
        self.p, self.s, self.order = p, s, p ** s
        self.exp = None
        # galois 0.1.x builds RS fields from matlab_primitive_poly(p, s) = the lexicographically-minimal primitive
        # polynomial, with the exception GF(2^7) -> x^7 + x^3 + 1 (the other exceptions, 2^14 and 2^16, exceed n <= 128)
        first = [0b0001001] if (p, s) == (2, 7) else []
        for low in first + list(range(1, self.order)):
            tab = self._powers_of_x(low)
            if tab is not None:
                self.poly_low, self.exp = low, tab
                break
        if self.exp is None:
            raise RuntimeError("no primitive polynomial")
        self.log = np.zeros(self.order, dtype=np.int32)
        self.log[self.exp[: self.order - 1]] = np.arange(self.order - 1, dtype=np.int32)

    def _digits(self, v):
        return [(v // self.p ** i) % self.p for i in range(self.s)]  # little endian

    def _undigits(self, d):
        return sum(int(c) % self.p * self.p ** i for i, c in enumerate(d))

    def _powers_of_x(self, low):
        """antilog table (doubled) if x^s + low is primitive, else None."""
        if low % self.p == 0:
            return None
        lowd = self._digits(low)
        v, seen, tab = 1, set(), []
        for _ in range(self.order - 1):
            if v in seen:
                return None
            seen.add(v)
            tab.append(v)
            d = self._digits(v)
            top = d[-1]
            d = [0] + d[:-1]
            if top:
                d = [(a - top * b) % self.p for a, b in zip(d, lowd)]
            v = self._undigits(d)
        if v != 1:
            return None
        return np.array(tab + tab, dtype=np.int32)

    def add(self, a, b):
        return self._undigits([x + y for x, y in zip(self._digits(a), self._digits(b))])

    def neg(self, a):
        return self._undigits([-x for x in self._digits(a)])

    def sub(self, a, b):
        return self.add(a, self.neg(b))

    def mul(self, a, b):
        if a == 0 or b == 0:
            return 0
        return int(self.exp[self.log[a] + self.log[b]])

    def inv(self, a):
        return int(self.exp[(self.order - 1 - self.log[a]) % (self.order - 1)])

    def alpha_pow(self, e):
        return int(self.exp[e % (self.order - 1)])

    def times_int(self, a, m):
        out = 0
        for _ in range(m % self.p):
            out = self.add(out, a)
        return out

    def vector(self, a):
        """degree-descending coefficient list."""
        return self._digits(a)[::-1]

    def from_vector(self, v):
        out = 0
        for c in v:
            out = out * self.p + int(c) % self.p
        return out

    def poly_eval(self, coeffs_low_first, x):
        acc, pw = 0, 1
        for c in coeffs_low_first:
            acc = self.add(acc, self.mul(c, pw))
            pw = self.mul(pw, x)
        return acc


class ReedSolomon:
    """Narrow-sense RS[nt, nt-2t] over GF(q^s) shortened to the last n coordinates (qsft/ReedSolomon.py:7-74)."""

    def __init__(self, n: int, t: int, q: int):
        s = math.ceil(math.log(n) / math.log(q))
        if n > q ** s - 1:
            s += 1
        self.s, self.ns, self.t, self.q = s, n, t, q
        self.n = q ** s - 1          # full code length (galois' .n)
        self.c = 1
        self.field = GaloisField(q, s)

    def get_parity_length(self):
        return 2 * self.t * self.s

    def get_delay_matrix(self):
        """(2ts + 1, n) integer matrix, first row zero; column i holds the GF(q) vectors of H[j, nt-n+i]."""
        F = self.field
        D = np.zeros((self.get_parity_length() + 1, self.ns), dtype=int)
        for i in range(self.ns):
            e = self.ns - 1 - i                      # H[j, i_full] = alpha^((c+j) * (nt-1-i_full)), nt-1-i_full = e
            for j in range(2 * self.t):
                D[self.s * j + 1:self.s * (j + 1) + 1, i] = F.vector(F.alpha_pow((self.c + j) * e))
        return D

    def syndrome_decode(self, syndrome):
        """2ts symbols of Z_q -> (k of shape (1, n), n_errors); (zeros, -1) when decoding fails."""
        F, t, s = self.field, self.t, self.s
        S = [F.from_vector(syndrome[s * i:s * (i + 1)]) for i in range(2 * t)]
        zero = np.zeros((1, self.ns), dtype=int)
        if not any(S):
            return zero, 0
        lam, prev, L, m, bb = [1], [1], 0, 1, 1
        for r in range(2 * t):
            d = S[r]
            for i in range(1, min(L, len(lam) - 1) + 1):
                d = F.add(d, F.mul(lam[i], S[r - i]))
            if d == 0:
                m += 1
                continue
            coef = F.mul(d, F.inv(bb))
            new = lam + [0] * max(0, len(prev) + m - len(lam))
            for i, pv in enumerate(prev):
                new[i + m] = F.sub(new[i + m], F.mul(coef, pv))
            if 2 * L <= r:
                prev, bb, L, m = lam, d, r + 1 - L, 1
            else:
                m += 1
            lam = new
        while len(lam) > 1 and lam[-1] == 0:
            lam = lam[:-1]
        deg = len(lam) - 1
        if deg != L or deg > t or deg == 0:
            return zero, -1
        omega = [0] * (2 * t)
        for a in range(2 * t):
            for i in range(min(deg, a) + 1):
                omega[a] = F.add(omega[a], F.mul(lam[i], S[a - i]))
        dlam = [F.times_int(lam[i], i) for i in range(1, deg + 1)]
        k = zero.copy()
        found = 0
        for i in range(self.ns):
            xinv = F.alpha_pow(-(self.ns - 1 - i))
            if F.poly_eval(lam, xinv) != 0:
                continue
            found += 1
            den = F.poly_eval(dlam, xinv)
            if den == 0:
                return zero, -1
            val = F.neg(F.mul(F.poly_eval(omega, xinv), F.inv(den)))
            if val >= self.q:
                return zero, -1
            k[0, i] = val
        if found != deg:
            return zero, -1
        return k, deg

    def device_tables(self):
        """(exp, log) int32 arrays for csrc/k4_peel.cu."""
        return np.ascontiguousarray(self.field.exp, dtype=np.int32), np.ascontiguousarray(self.field.log, dtype=np.int32)
