// Device code shared by the K4 kernels (k4_peel.cu: stand-alone classification / reduce / apply kernels and detectors;
// k4_peel_loop.cu: the persistent on-device round loop).  Also compiled by the CPU emulation (tests/emu): no __shared__
// declarations and no sm_100a-only instructions in this file.
#pragma once

// plain data shared across translation units (k4_peel.cu fills it from qsft_peel_desc)
struct PeelDev {
    int q, n, b, C, P, P_src, R, channel, source, rs_t, rs_s, ld;
    unsigned int qmagic;      // ceil(2^32 / q): x mod q = x - mulhi(x, qmagic) * q for x < 2^32 / q
    long long B;
    double thresh;            // cutoff * P
    double invP;              // 1 / P
    const int8_t* MT;         // (C, b, ld)   rows = columns of M, zero padded
    const int8_t* D;          // (C, P, ld)
    const int32_t* rs_exp;
    const int32_t* rs_log;
    int rs_order;             // q^s
    int fastdet;              // q = 2 / 4 symbols by quadrant comparison instead of atan2f (default; QSFT_K4_FASTDET=0 disables)
};

// distinct-k output of the peel (see k4_reduce_kernel / qsft_uniq)
struct UniqOut {
    int32_t* seen0;       // (B) chain heads, zero initialised by the caller / qsft_peel
    int8_t* uk;           // (max_uniq, ld)
    float* usum;          // (max_uniq) complex64: sum of rho over all finds of the k
    int32_t* ucnt;        // (max_uniq) number of finds
    long long* ukey;      // (max_uniq) (round << 48) | (c * B + j) of the first find: reference's first-seen order
    int32_t* unext;       // (max_uniq) workspace
    long long max_uniq;
};

namespace {

constexpr int K4_THREADS = 128;
constexpr int RS_MAX_2T = 32;
constexpr double kTwoPi = 6.283185307179586476925286766559;


__device__ __forceinline__ int dp4a_u(uint32_t a, uint32_t b, int c) {
#ifdef QSFT_EMU   // CPU execution of this kernel source by tests/emu (test infrastructure; never defined in the product build)
    for (int i = 0; i < 4; ++i) c += (int)((a >> (8 * i)) & 255u) * (int)((b >> (8 * i)) & 255u);
    return c;
#else
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#endif
}

// <row, k> (not reduced; < 128 * 127^2 < 2^32 / q); `row` has ld bytes (ld % 16 == 0), kw holds the digits of k four per word (zero padded)
template <int NW>
__device__ __forceinline__ int dot_raw(const int8_t* row, int ld, const uint32_t (&kw)[NW]) {
    const uint4* r4 = reinterpret_cast<const uint4*>(row);
    const int nv = ld >> 4;
    int acc = 0;
#pragma unroll
    for (int w = 0; w < NW / 4; ++w) {
        if (w >= nv) break;
        uint4 v = __ldg(r4 + w);
        acc = dp4a_u(v.x, kw[4 * w + 0], acc);
        acc = dp4a_u(v.y, kw[4 * w + 1], acc);
        acc = dp4a_u(v.z, kw[4 * w + 2], acc);
        acc = dp4a_u(v.w, kw[4 * w + 3], acc);
    }
    return acc;
}

__device__ __forceinline__ int fast_mod(int x, int q, unsigned int qmagic) {   // 0 <= x < 2^32 / q
    return x - (int)(__umulhi((unsigned int)x, qmagic) * (unsigned int)q);
}

// bin hash j = dec(M_c^T k mod q), b digits MSB first (qsft.py:178, :227)
template <int NW>
__device__ __forceinline__ long long hash_bin(const PeelDev& d, int c, const uint32_t (&kw)[NW]) {
    long long j = 0;
    const int8_t* mt = d.MT + (size_t)c * d.b * d.ld;
    for (int i = 0; i < d.b; ++i) j = j * d.q + fast_mod(dot_raw<NW>(mt + (size_t)i * d.ld, d.ld, kw), d.q, d.qmagic);
    return j;
}

// weight q^(b-1-i) of hash digit i (0 for i >= b): loop invariant of the per-bin / per-find work, computed once per thread
__device__ __forceinline__ long long hash_weight(const PeelDev& d, int i) {
    if (i >= d.b) return 0;
    long long wgt = 1;
    for (int u = i + 1; u < d.b; ++u) wgt *= d.q;
    return wgt;
}

// the same hash computed by a whole warp: one hash digit per lane (b <= 32), all lanes get the result;
// wgt = hash_weight(d, lane)
template <int NW>
__device__ __forceinline__ long long hash_bin_warp(const PeelDev& d, int c, const uint32_t (&kw)[NW], int lane, long long wgt) {
    long long part = 0;
    if (lane < d.b) part = wgt * fast_mod(dot_raw<NW>(d.MT + ((size_t)c * d.b + lane) * d.ld, d.ld, kw), d.q, d.qmagic);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    return part;
}

// ---------------------------------------------------------------------------------------------------------
// GF(p^s) helpers for the coded path; elements are ints whose base-p digits are polynomial coefficients.
// ---------------------------------------------------------------------------------------------------------
struct GF {
    int p, s, order;
    const int32_t* ex;
    const int32_t* lg;
    __device__ __forceinline__ int add(int a, int b) const {
        int out = 0, w = 1;
        for (int i = 0; i < s; ++i) {
            int da = a % p, db = b % p;
            a /= p; b /= p;
            int v = da + db;
            v = v >= p ? v - p : v;
            out += v * w;
            w *= p;
        }
        return out;
    }
    __device__ __forceinline__ int neg(int a) const {
        int out = 0, w = 1;
        for (int i = 0; i < s; ++i) {
            int da = a % p;
            a /= p;
            out += (da ? p - da : 0) * w;
            w *= p;
        }
        return out;
    }
    __device__ __forceinline__ int sub(int a, int b) const { return add(a, neg(b)); }
    __device__ __forceinline__ int mul(int a, int b) const {
        if (a == 0 || b == 0) return 0;
        return __ldg(ex + __ldg(lg + a) + __ldg(lg + b));
    }
    __device__ __forceinline__ int inv(int a) const { return __ldg(ex + (order - 1 - __ldg(lg + a)) % (order - 1)); }
    __device__ __forceinline__ int alpha_pow(int e) const {
        e %= (order - 1);
        if (e < 0) e += order - 1;
        return __ldg(ex + e);
    }
};

// Syndrome decode: sym (2ts symbols of Z_q) -> k digits (n), returns false on decoder failure (k left all zero,
// like galois returning the unchanged zero codeword with n_errors = -1).
// Out of line, with its parameters BY VALUE: a reference to the kernel's PeelDev parameter would force a local-memory copy of
// the whole struct (and every d.field access through it); the decoder is long and rare, so it is one shared copy of code.
struct RsParams {
    int q, n, rs_t, rs_s, rs_order;
    const int32_t* rs_exp;
    const int32_t* rs_log;
};
__device__ __forceinline__ RsParams rs_params(const PeelDev& d) { return RsParams{d.q, d.n, d.rs_t, d.rs_s, d.rs_order, d.rs_exp, d.rs_log}; }

__device__ __noinline__ bool rs_decode(RsParams d, const uint8_t* sym, uint8_t* kout) {
    GF F{d.q, d.rs_s, d.rs_order, d.rs_exp, d.rs_log};
    const int t = d.rs_t, s = d.rs_s, n = d.n, nt = d.rs_order - 1;
    const int T2 = 2 * t;
    int S[RS_MAX_2T];
    bool any = false;
    for (int i = 0; i < T2; ++i) {
        int v = 0;
        for (int u = 0; u < s; ++u) v = v * d.q + sym[s * i + u];
        S[i] = v;
        any |= (v != 0);
    }
    for (int i = 0; i < n; ++i) kout[i] = 0;
    if (!any) return true;
    int Lam[RS_MAX_2T + 2], Bp[RS_MAX_2T + 2], Nw[RS_MAX_2T + 2];
    for (int i = 0; i < T2 + 2; ++i) Lam[i] = Bp[i] = 0;
    Lam[0] = Bp[0] = 1;
    int L = 0, m = 1, bb = 1, lenL = 1, lenB = 1;
    for (int r = 0; r < T2; ++r) {
        int dd = S[r];
        for (int i = 1; i <= L; ++i)
            if (i < lenL) dd = F.add(dd, F.mul(Lam[i], S[r - i]));
        if (dd == 0) {
            ++m;
            continue;
        }
        int coef = F.mul(dd, F.inv(bb));
        int lenN = max(lenL, lenB + m);
        if (lenN > T2 + 2) return false;
        for (int i = 0; i < lenN; ++i) Nw[i] = i < lenL ? Lam[i] : 0;
        for (int i = 0; i < lenB; ++i) Nw[i + m] = F.sub(Nw[i + m], F.mul(coef, Bp[i]));
        if (2 * L <= r) {
            for (int i = 0; i < lenL; ++i) Bp[i] = Lam[i];
            lenB = lenL;
            bb = dd;
            L = r + 1 - L;
            m = 1;
        } else {
            ++m;
        }
        for (int i = 0; i < lenN; ++i) Lam[i] = Nw[i];
        lenL = lenN;
    }
    while (lenL > 1 && Lam[lenL - 1] == 0) --lenL;
    const int deg = lenL - 1;
    if (deg != L || deg > t || deg == 0) return false;
    // Omega = S(x) Lambda(x) mod x^2t
    int Om[RS_MAX_2T];
    for (int a = 0; a < T2; ++a) {
        int v = 0;
        for (int i = 0; i <= deg && i <= a; ++i) v = F.add(v, F.mul(Lam[i], S[a - i]));
        Om[a] = v;
    }
    int found = 0;
    bool ok = true;
    for (int i = 0; i < n && ok; ++i) {
        const int e = n - 1 - i;              // locator X = alpha^e for retained coordinate i
        const int xinv = F.alpha_pow(-e);
        int acc = 0, pw = 1;
        for (int c = 0; c <= deg; ++c) {
            acc = F.add(acc, F.mul(Lam[c], pw));
            pw = F.mul(pw, xinv);
        }
        if (acc != 0) continue;
        ++found;
        int num = 0;
        pw = 1;
        for (int c = 0; c < T2; ++c) {
            num = F.add(num, F.mul(Om[c], pw));
            pw = F.mul(pw, xinv);
        }
        int den = 0;
        pw = 1;
        for (int c = 1; c <= deg; ++c) {
            int term = 0;
            for (int u = 0; u < c % d.q; ++u) term = F.add(term, Lam[c]);
            den = F.add(den, F.mul(term, pw));
            pw = F.mul(pw, xinv);
        }
        if (den == 0) {
            ok = false;
            break;
        }
        const int val = F.neg(F.mul(num, F.inv(den)));
        if (val >= d.q) ok = false;            // error value must be in the prime subfield
        kout[i] = (uint8_t)val;
    }
    (void)nt;
    if (!ok || found != deg) {
        for (int i = 0; i < n; ++i) kout[i] = 0;
        return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// singleton detection: symbol i (1 <= i < P_src) of one column; element p of the column is col[p * stride].
//   channel 0: noiseless angles (reconstruct.py:12-31), channel 1: nso1 soft decision (reconstruct.py:100-113),
//   channel 2: nso2 hard decision (reconstruct.py:116-129 + angle_q, utils.py:104-105).
// ---------------------------------------------------------------------------------------------------------
// angle_q: (((angle mod 2 pi) // (pi / q)) + 1) // 2 mod q in fp64 like NumPy (np.angle of a complex128 holding the
// fp32 value; the float floor divisions are exact small integers)
__device__ __noinline__ int angle_q_dev(float2 v, int q) {
    double a = atan2((double)v.y, (double)v.x);
    if (a < 0.0) a += kTwoPi;                      // numpy: angle % (2 pi)
    if (a >= kTwoPi) a -= kTwoPi;
    const long long sector = (long long)floor(a / (3.14159265358979323846 / (double)q));
    return (int)(((sector + 1) >> 1) % q);
}

// Index of the q-th root of unity nearest to the direction of (re, im) for q = 2 / 4 by comparisons (opt-in fast path):
// the quadrant boundaries are the diagonals (q = 4) / the imaginary axis (q = 2); a value within ~0.03 rad of a boundary
// (or a vanishing one) returns -1 and takes the exact path, so the decision always equals the exact one.
__device__ __forceinline__ int quadrant_symbol(int q, float re, float im) {
    const float ax = fabsf(re), ay = fabsf(im);
    if (!(ax + ay > 1e-30f)) return -1;
    if (q == 4) {
        if (!(fabsf(ax - ay) > 0.03f * (ax + ay))) return -1;
        return ax > ay ? (re > 0.f ? 0 : 2) : (im > 0.f ? 1 : 3);
    }
    if (!(ax > 0.03f * (ax + ay))) return -1;
    return re > 0.f ? 0 : 1;
}

// column accessors: element p of the column under test
struct StridedCol {
    const float2* col;
    size_t stride;
    int P_src;
    // element i of repeat block r (delay row r * P_src + i)
    __device__ __forceinline__ float2 ri(int r, int i) const { return col[(size_t)(r * P_src + i) * stride]; }
};

__device__ __noinline__ int symbol_noiseless_exact(int q, float2 v0, float2 v) {
    const double a0 = atan2((double)v0.y, (double)v0.x);
    const double a = atan2((double)v.y, (double)v.x);
    const long long r = (long long)rint((double)q * (a - a0) / kTwoPi);          // half-to-even like np.round
    const int m = (int)(r % q);
    return m < 0 ? m + q : m;
}

// ---- decisions on values (shared by the column-accessor form below and the register-resident form of k4_peel_loop.cu) ----
// noiseless (reconstruct.py:12-31): round(q (angle v - angle v0) / 2 pi) mod q = root nearest to the direction of v conj(v0)
// fp32 paths for general q (one shared copy of the atan2f code; scalar parameters, see rs_decode); -1: too close to a
// decision boundary for fp32, redo in fp64
__device__ __noinline__ int symbol_noiseless_angle(int q, float2 v0, float2 v) {
    if (!(fabsf(v0.x) + fabsf(v0.y) > 1e-30f && fabsf(v.x) + fabsf(v.y) > 1e-30f)) return -1;
    const float u = (float)q * (atan2f(v.y, v.x) - atan2f(v0.y, v0.x)) * 0.15915494309189535f;
    const float m = rintf(u);
    if (!(fabsf(u - m) < 0.49f)) return -1;
    int mi = (int)m;                            // |u| < q  =>  m in [-q, q]
    mi = mi < 0 ? mi + q : mi;
    return mi >= q ? mi - q : mi;
}
__device__ __noinline__ int symbol_nso1_angle(int q, float arf, float aif) {
    if (!(fabsf(arf) + fabsf(aif) > 1e-30f)) return -1;
    float thf = atan2f(aif, arf);
    if (thf < 0.f) thf += 6.283185307179586f;
    const float u = thf * (float)q * 0.15915494309189535f;              // in [0, q]
    const float m = rintf(u);
    if (!(fabsf(u - m) < 0.49f)) return -1;
    return ((int)m >= q) ? (int)m - q : (int)m;                         // nearest of the q+1 roots, mod q
}

__device__ __forceinline__ int symbol_noiseless(const PeelDev& d, float2 v0, float2 v) {
    int symv = -1;
    if (d.fastdet && (d.q == 4 || d.q == 2))
        symv = quadrant_symbol(d.q, fmaf(v.x, v0.x, v.y * v0.y), fmaf(v.y, v0.x, -(v.x * v0.y)));
    // fast path in fp32; anything within 0.01 of a rounding boundary is redone in fp64 so the decision
    // always equals the fp64 one (np.angle / np.round in the reference)
    if (symv < 0) symv = symbol_noiseless_angle(d.q, v0, v);
    if (symv < 0) symv = symbol_noiseless_exact(d.q, v0, v);
    return symv;
}

// nso1 (reconstruct.py:100-113) from the fp32 sum over the repeats of z conj(v): -1 when the value sits too close to a
// decision boundary for fp32 (both fast decisions keep a margin far above the rounding of the sum)
__device__ __forceinline__ int symbol_nso1_fast(const PeelDev& d, float arf, float aif) {
    int symv = -1;
    if (d.fastdet && (d.q == 4 || d.q == 2)) symv = quadrant_symbol(d.q, arf, aif);
    if (symv < 0) symv = symbol_nso1_angle(d.q, arf, aif);
    return symv;
}

// ... and from the fp64 sum, exactly like NumPy: argmin over the q + 1 roots, first minimum (np.mean divides by R > 0: the
// angle does not depend on it)
__device__ __noinline__ int symbol_nso1_exact(int q, double ar, double ai) {
    double th = atan2(ai, ar);
    if (th < 0.0) th += kTwoPi;                                          // numpy: angle % (2 pi)
    if (th >= kTwoPi) th -= kTwoPi;
    const double step = kTwoPi / (double)q;
    int best = 0;
    double bd = fabs(0.0 - th);
    for (int m = 1; m <= q; ++m) {
        const double dist = fabs(step * (double)m - th);
        if (dist < bd) {
            bd = dist;
            best = m;
        }
    }
    return best % q;
}

// nso2 (reconstruct.py:116-129): every repeat votes with its quantised phase difference; the votes are averaged as numbers,
// np.round is half-to-even; the sum of R small integers and the division by R are exact / correctly rounded like np.mean
__device__ __forceinline__ int nso2_vote(const PeelDev& d, float2 z, float2 v) {
    const int df = angle_q_dev(z, d.q) - angle_q_dev(v, d.q);
    return df < 0 ? df + d.q : df;
}
__device__ __forceinline__ int symbol_nso2(const PeelDev& d, long long votes) {
    return (int)((long long)rint((double)votes / (double)d.R) % d.q);
}

template <class Col>
__device__ __forceinline__ int detect_symbol(const PeelDev& d, const Col& col, int i) {
    if (d.channel == 0) return symbol_noiseless(d, col.ri(0, 0), col.ri(0, i));
    if (d.channel == 1) {
        float arf = 0.f, aif = 0.f;
        for (int r = 0; r < d.R; ++r) {
            const float2 z = col.ri(r, 0);
            const float2 v = col.ri(r, i);
            arf = fmaf(z.x, v.x, fmaf(z.y, v.y, arf));                       // z * conj(v)
            aif = fmaf(z.y, v.x, fmaf(-z.x, v.y, aif));
        }
        const int symv = symbol_nso1_fast(d, arf, aif);
        if (symv >= 0) return symv;
        double ar = 0.0, ai = 0.0;
        for (int r = 0; r < d.R; ++r) {
            const float2 z = col.ri(r, 0);
            const float2 v = col.ri(r, i);
            ar += (double)z.x * v.x + (double)z.y * v.y;
            ai += (double)z.y * v.x - (double)z.x * v.y;
        }
        return symbol_nso1_exact(d.q, ar, ai);
    }
    long long votes = 0;
    for (int r = 0; r < d.R; ++r) votes += nso2_vote(d, col.ri(r, 0), col.ri(r, i));
    return symbol_nso2(d, votes);
}


// Record `cnt` finds of k (digits kw, sum of their rho = `sum`) of round `round`, first found at cj = c * B + j; j0 = k's
// bin in group 0 (chain key).  Re-finds of a k recorded in an EARLIER round are merged into that entry, otherwise a new
// entry is appended and linked into the chain of j0.  Called once per distinct k and round (by its first find).
template <int NW>
__device__ __forceinline__ void k4_uniq_commit(const PeelDev& d, const uint32_t (&kw)[NW], float2 sum, int cnt, long long cj,
                                               long long j0, int round, int32_t* __restrict__ seen0, int8_t* __restrict__ uk,
                                               float* __restrict__ usum, int32_t* __restrict__ ucnt, long long* __restrict__ ukey,
                                               int32_t* __restrict__ unext, long long max_uniq,
                                               unsigned long long* __restrict__ counters) {
    const int nw = d.ld / 4;
    // merge with an entry of an earlier round, if any
    int32_t head = *reinterpret_cast<volatile int32_t*>(seen0 + j0);
    // entries of this round may have been published by other SMs a moment ago: read the links through L2 (__ldcg)
    for (int32_t e = head; e != 0; e = __ldcg(unext + (e - 1))) {
        if ((int)(__ldcg(ukey + (e - 1)) >> 48) == round) continue;
        const uint32_t* k2 = reinterpret_cast<const uint32_t*>(uk + (size_t)(e - 1) * d.ld);
        bool same = true;
#pragma unroll
        for (int w = 0; w < NW; ++w) same &= ((w < nw) ? k2[w] : 0u) == kw[w];
        if (same) {
            atomicAdd(usum + 2 * (size_t)(e - 1), sum.x);
            atomicAdd(usum + 2 * (size_t)(e - 1) + 1, sum.y);
            atomicAdd(ucnt + (e - 1), cnt);
            return;
        }
    }
    const unsigned long long u = atomicAdd(&counters[4], 1ull);
    if ((long long)u >= max_uniq) return;
    uint32_t* ko = reinterpret_cast<uint32_t*>(uk + (size_t)u * d.ld);
    for (int w = 0; w < nw; ++w) ko[w] = (w < NW) ? kw[w] : 0u;
    usum[2 * u] = sum.x;
    usum[2 * u + 1] = sum.y;
    ucnt[u] = cnt;
    ukey[u] = ((long long)round << 48) | cj;
    unext[u] = head;
    __threadfence();
    for (;;) {
        const int32_t old = atomicCAS(seen0 + j0, head, (int32_t)(u + 1));
        if (old == head) break;
        head = old;                          // another k of this round was linked first: chain behind it
        unext[u] = head;
        __threadfence();
    }
}

}  // namespace
