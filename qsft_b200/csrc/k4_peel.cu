// K4: peeling decoder.  Replaces the round loop of QSFT.transform (qsft/qsft.py:151-241), the singleton detectors
// (qsft/reconstruct.py:12-31 noiseless, :100-113 nso1, :34-51 coded) and the Reed-Solomon syndrome decoder
// (qsft/ReedSolomon.py:26-48, arithmetic of galois' decode_jit restated: Berlekamp-Massey + Chien + Forney).
//
// Layout: U (C, P, B) complex64 with the bin index j contiguous.  Classification maps one THREAD to one bin with lanes
// over consecutive j, so every load of a delay row is a fully coalesced 256-byte warp request (HBM bound: one read of
// U per round).  Rounds are synchronous like the reference: classify on a frozen U, then subtract.
#include "common.cuh"

#include <stdlib.h>

namespace {

constexpr int K4_THREADS = 128;
constexpr int RS_MAX_2T = 32;
constexpr double kTwoPi = 6.283185307179586476925286766559;

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
    int fastdet;              // opt-in (QSFT_K4_FASTDET=1): q = 2 / 4 symbols by quadrant comparison instead of atan2f
};

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
__device__ bool rs_decode(const PeelDev& d, const uint8_t* sym, uint8_t* kout) {
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
__device__ __forceinline__ int angle_q_dev(float2 v, int q) {
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

__device__ __forceinline__ int detect_symbol(const PeelDev& d, const float2* __restrict__ col, size_t stride, int i) {
    const double qd = (double)d.q;
    const bool quad = d.fastdet && (d.q == 4 || d.q == 2);
    int symv;
    if (d.channel == 0) {
        const float2 v0 = col[0];
        const float2 v = col[(size_t)i * stride];
        symv = -1;
        // round(q (angle v - angle v0) / 2 pi) mod q = root nearest to the direction of v conj(v0)
        if (quad) symv = quadrant_symbol(d.q, fmaf(v.x, v0.x, v.y * v0.y), fmaf(v.y, v0.x, -(v.x * v0.y)));
        // fast path in fp32; anything within 0.01 of a rounding boundary is redone in fp64 so the decision
        // always equals the fp64 one (np.angle / np.round in the reference)
        if (symv < 0 && fabsf(v0.x) + fabsf(v0.y) > 1e-30f && fabsf(v.x) + fabsf(v.y) > 1e-30f) {
            const float u = (float)d.q * (atan2f(v.y, v.x) - atan2f(v0.y, v0.x)) * 0.15915494309189535f;
            const float m = rintf(u);
            if (fabsf(u - m) < 0.49f) {
                int mi = (int)m;                    // |u| < q  =>  m in [-q, q]
                mi = mi < 0 ? mi + d.q : mi;
                symv = mi >= d.q ? mi - d.q : mi;
            }
        }
        if (symv < 0) {
            const double a0 = atan2((double)v0.y, (double)v0.x);
            const double a = atan2((double)v.y, (double)v.x);
            const long long r = (long long)rint(qd * (a - a0) / kTwoPi);    // half-to-even like np.round
            const int m = (int)(r % d.q);
            symv = m < 0 ? m + d.q : m;
        }
    } else if (d.channel == 1) {
        double ar = 0.0, ai = 0.0;
        for (int r = 0; r < d.R; ++r) {
            const float2 z = col[(size_t)(r * d.P_src) * stride];
            const float2 v = col[(size_t)(r * d.P_src + i) * stride];
            ar += (double)z.x * v.x + (double)z.y * v.y;                     // z * conj(v)
            ai += (double)z.y * v.x - (double)z.x * v.y;
        }
        // np.mean divides by R > 0: the angle does not depend on it
        symv = -1;
        const float arf = (float)ar, aif = (float)ai;
        if (quad) symv = quadrant_symbol(d.q, arf, aif);
        if (symv < 0 && fabsf(arf) + fabsf(aif) > 1e-30f) {
            float thf = atan2f(aif, arf);
            if (thf < 0.f) thf += 6.283185307179586f;
            const float u = thf * (float)d.q * 0.15915494309189535f;        // in [0, q]
            const float m = rintf(u);
            if (fabsf(u - m) < 0.49f) symv = ((int)m >= d.q) ? (int)m - d.q : (int)m;   // nearest of the q+1 roots, mod q
        }
        if (symv < 0) {
            double th = atan2(ai, ar);
            if (th < 0.0) th += kTwoPi;                                      // numpy: angle % (2 pi)
            if (th >= kTwoPi) th -= kTwoPi;
            const double step = kTwoPi / qd;
            int best = 0;
            double bd = fabs(0.0 - th);
            for (int m = 1; m <= d.q; ++m) {                                 // argmin over q+1 roots, first minimum
                const double dist = fabs(step * (double)m - th);
                if (dist < bd) {
                    bd = dist;
                    best = m;
                }
            }
            symv = best % d.q;
        }
    } else {
        // nso2: every repeat votes with its quantised phase difference; the votes are averaged as numbers, np.round is
        // half-to-even; the sum of R small integers and the division by R are exact / correctly rounded like np.mean
        long long votes = 0;
        for (int r = 0; r < d.R; ++r) {
            const int a0 = angle_q_dev(col[(size_t)(r * d.P_src) * stride], d.q);
            const int a = angle_q_dev(col[(size_t)(r * d.P_src + i) * stride], d.q);
            int df = a0 - a;
            votes += df < 0 ? df + d.q : df;
        }
        symv = (int)((long long)rint((double)votes / (double)d.R) % d.q);
    }
    return symv;
}

// ---------------------------------------------------------------------------------------------------------
// classification.  Phase 1: one THREAD per bin computes the energy (lanes over consecutive bins -> coalesced row
// reads; this is the HBM-bound part).  Phase 2: the WARP walks over its non-zeroton bins one at a time with lanes
// over the delay rows (the rows were just read, so these loads hit L1/L2): all P_src-1 symbols are detected in
// parallel, then rho / residual / bin hash are warp reductions.  No divergent per-thread detection.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NW>
__global__ void __launch_bounds__(K4_THREADS)
k4_classify_kernel(PeelDev d, const float2* __restrict__ U, long long j_begin, long long j_end,
                   long long* __restrict__ find_cj, int8_t* __restrict__ find_k, float2* __restrict__ find_rho,
                   int32_t* __restrict__ find_round, int32_t* __restrict__ find_id, long long max_finds, int round,
                   unsigned long long* __restrict__ counters) {
    __shared__ double2 s_tw[QSFT_MAX_Q + 1];                       // w^t = (cos, sin)(2 pi t / q)
    __shared__ __align__(16) uint8_t s_sym[K4_THREADS / 32][QSFT_MAX_N];   // detected symbols of the warp's current bin
    __shared__ __align__(16) uint8_t s_k[K4_THREADS / 32][QSFT_MAX_N];     // decoded k digits (zero padded)
    if (threadIdx.x < d.q) {
        double sn, cs;
        sincospi(2.0 * (double)threadIdx.x / (double)d.q, &sn, &cs);
        s_tw[threadIdx.x] = make_double2(cs, sn);
    }
    __syncthreads();
    const int c = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long jw = j_begin + (long long)blockIdx.x * K4_THREADS + warp * 32;   // first bin of this warp
    const long long j = jw + lane;
    const long long B = d.B;
    const float2* Uc = U + (size_t)c * d.P * B;

    // phase 1: energy test (qsft.py:164)
    double energy = 0.0;
    if (j < j_end) {
        const float2* col = Uc + j;
        for (int p = 0; p < d.P; ++p) {
            float2 v = col[(size_t)p * B];
            energy += (double)v.x * v.x + (double)v.y * v.y;
        }
        if (!(energy > d.thresh)) find_id[(size_t)c * B + j] = -1;
    }
    unsigned mask = __ballot_sync(0xffffffffu, j < j_end && energy > d.thresh);
    if (mask == 0) return;

    const int nsym = d.P_src - 1;
    const int8_t* Dc = d.D + (size_t)c * d.P * d.ld;
    const long long wgt = hash_weight(d, lane);                   // loop invariants of the per-bin work
    for (int i = d.n + lane; i < 4 * NW && i < QSFT_MAX_N; i += 32) s_k[warp][i] = 0;   // zero padding of k: written once
    unsigned n_multi = 0;
    while (mask) {
        const int bi = __ffs(mask) - 1;
        mask &= mask - 1;
        const long long jb = jw + bi;
        const double e_b = __shfl_sync(0xffffffffu, energy, bi);
        const float2* col = Uc + jb;

        // singleton detection -> symbols, one delay row per lane (reconstruct.py)
        for (int i0 = 1; i0 <= nsym; i0 += 32) {
            const int i = i0 + lane;
            if (i <= nsym) {
                const int symv = detect_symbol(d, col, (size_t)B, i);
                s_sym[warp][i - 1] = (uint8_t)symv;
            }
        }
        __syncwarp();
        // k digits (zero padded to 4 * NW bytes)
        if (d.source == 1) {
            if (lane == 0) rs_decode(d, s_sym[warp], s_k[warp]);
        } else {
            for (int i = lane; i < d.n; i += 32) s_k[warp][i] = s_sym[warp][i];
        }
        __syncwarp();
        uint32_t kw[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) kw[w] = reinterpret_cast<const uint32_t*>(s_k[warp])[w];

        // rho = <signature, col> / P with signature_p = w^(D_p . k)  (qsft.py:174-175), lanes over delay rows
        double rr = 0.0, ri = 0.0;
        for (int p = lane; p < d.P; p += 32) {
            const int t = fast_mod(dot_raw<NW>(Dc + (size_t)p * d.ld, d.ld, kw), d.q, d.qmagic);
            const double cs = s_tw[t].x, sn = s_tw[t].y;
            const float2 v = col[(size_t)p * B];
            rr += cs * v.x + sn * v.y;                                                   // conj(sig) * v
            ri += cs * v.y - sn * v.x;
        }
        rr = warp_sum(rr) * d.invP;
        ri = warp_sum(ri) * d.invP;
        // ||col - rho sig||^2 = ||col||^2 - P |rho|^2 exactly (rho is the projection, |sig_p| = 1); fp64 keeps it accurate
        const double res = e_b - (double)d.P * (rr * rr + ri * ri);
        // bin hash j = dec(M_c^T k mod q) (qsft.py:178-179), one hash digit per lane
        long long part = 0;
        if (lane < d.b) part = wgt * fast_mod(dot_raw<NW>(d.MT + ((size_t)c * d.b + lane) * d.ld, d.ld, kw), d.q, d.qmagic);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        const bool single = (part == jb) && !(res > d.thresh);                          // qsft.py:183
        if (!single) {
            if (lane == 0) {
                find_id[(size_t)c * B + jb] = -1;
                ++n_multi;
            }
        } else {
            unsigned long long f = 0;
            if (lane == 0) f = atomicAdd(&counters[0], 1ull);
            f = __shfl_sync(0xffffffffu, f, 0);
            if ((long long)f < max_finds) {
                uint32_t* ko = reinterpret_cast<uint32_t*>(find_k + (size_t)f * d.ld);
                for (int w = lane; w < d.ld / 4; w += 32) ko[w] = (w < NW) ? reinterpret_cast<const uint32_t*>(s_k[warp])[w] : 0u;
                if (lane == 0) {
                    find_cj[f] = (long long)c * B + jb;
                    find_rho[f] = make_float2((float)rr, (float)ri);
                    if (find_round) find_round[f] = round;
                    find_id[(size_t)c * B + jb] = (int32_t)f;
                }
            } else if (lane == 0) {
                find_id[(size_t)c * B + jb] = -1;
            }
        }
        __syncwarp();
    }
    if (lane == 0 && n_multi) atomicAdd(&counters[1], (unsigned long long)n_multi);
}

// ---------------------------------------------------------------------------------------------------------
// classification, version 2 (opt-in: QSFT_K4_IMPL=2; written at the end of round 1 without GPU time -- NOT YET
// MEASURED, the default stays version 1 above).  Same decisions and outputs as k4_classify_kernel.
//   CTA = 256 consecutive bins of one group.  Phase 1 (thread = bin) streams the tile's columns through the energy sum
//   AND into shared memory (every global byte is read once, coalesced); the CTA compacts its candidate bins; phase 2
//   gives every candidate a group of 8 lanes (lanes over delay rows / symbols / hash digits: P = 41 -> 6 iterations at
//   85 % lane utilisation instead of 2 at 64 %, 3-step group reductions, columns / D / M^T rows from shared memory).
//   Shared layout: columns [P][257] float2 (odd stride: the 8 lanes of a group hit 8 different banks), D rows and M^T rows
//   with stride ld + 16 bytes (conflict-free 16-byte loads), symbols [32 groups][4 NW].
// Identity source decoding only (coded delays fall back to version 1).
// ---------------------------------------------------------------------------------------------------------
constexpr int K4V2_THREADS = 256;
constexpr int K4V2_COLS = K4V2_THREADS + 1;     // padded column stride (float2 elements)
constexpr int K4V2_G = 8;                       // lanes per candidate bin

template <int NW>
__device__ __forceinline__ int dot_raw_s(const int8_t* row, int ld, const uint32_t (&kw)[NW]) {   // row in shared memory
    const uint4* r4 = reinterpret_cast<const uint4*>(row);
    const int nv = ld >> 4;
    int acc = 0;
#pragma unroll
    for (int w = 0; w < NW / 4; ++w) {
        if (w >= nv) break;
        const uint4 v = r4[w];
        acc = dp4a_u(v.x, kw[4 * w + 0], acc);
        acc = dp4a_u(v.y, kw[4 * w + 1], acc);
        acc = dp4a_u(v.z, kw[4 * w + 2], acc);
        acc = dp4a_u(v.w, kw[4 * w + 3], acc);
    }
    return acc;
}

static size_t k4v2_smem_bytes(const PeelDev& d, int nw) {
    return (size_t)d.P * K4V2_COLS * sizeof(float2) + 16 +             // columns (+ alignment slack)
           (size_t)(d.P + d.b) * (d.ld + 16) +                           // D rows, M^T rows
           (size_t)(K4V2_THREADS / K4V2_G) * 4 * nw;                     // symbols
}

template <int NW>
__global__ void __launch_bounds__(K4V2_THREADS, 2)
k4_classify_v2_kernel(PeelDev d, const float2* __restrict__ U, long long j_begin, long long j_end,
                      long long* __restrict__ find_cj, int8_t* __restrict__ find_k, float2* __restrict__ find_rho,
                      int32_t* __restrict__ find_round, int32_t* __restrict__ find_id, long long max_finds, int round,
                      unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char k4v2_smem[];
    __shared__ double2 s_tw[QSFT_MAX_Q + 1];
    __shared__ double s_energy[K4V2_THREADS];
    __shared__ int s_cand[K4V2_THREADS];
    __shared__ int s_wcnt[K4V2_THREADS / 32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c = blockIdx.y;
    const long long B = d.B;
    const int rs = d.ld + 16;                                            // padded row stride of the D / M^T copies
    float2* s_col = reinterpret_cast<float2*>(k4v2_smem);
    size_t off = ((size_t)d.P * K4V2_COLS * sizeof(float2) + 15) & ~(size_t)15;
    int8_t* s_D = reinterpret_cast<int8_t*>(k4v2_smem + off);
    int8_t* s_MT = s_D + (size_t)d.P * rs;
    uint8_t* s_sym = reinterpret_cast<uint8_t*>(s_MT + (size_t)d.b * rs);

    // phase 0: tables
    if (tid < d.q) {
        double sn, cs;
        sincospi(2.0 * (double)tid / (double)d.q, &sn, &cs);
        s_tw[tid] = make_double2(cs, sn);
    }
    {
        const int vpr = d.ld >> 4;                                       // 16-byte vectors per row
        const uint4* gD = reinterpret_cast<const uint4*>(d.D + (size_t)c * d.P * d.ld);
        for (int e = tid; e < d.P * vpr; e += K4V2_THREADS) {
            const int r = e / vpr, v = e - r * vpr;
            *reinterpret_cast<uint4*>(s_D + (size_t)r * rs + 16 * v) = __ldg(gD + e);
        }
        const uint4* gM = reinterpret_cast<const uint4*>(d.MT + (size_t)c * d.b * d.ld);
        for (int e = tid; e < d.b * vpr; e += K4V2_THREADS) {
            const int r = e / vpr, v = e - r * vpr;
            *reinterpret_cast<uint4*>(s_MT + (size_t)r * rs + 16 * v) = __ldg(gM + e);
        }
    }

    // phase 1: energy (same summation order as version 1) + stash of the columns
    const long long j0 = j_begin + (long long)blockIdx.x * K4V2_THREADS;
    const long long j = j0 + tid;
    const float2* Uc = U + (size_t)c * d.P * B;
    double energy = 0.0;
    if (j < j_end) {
        const float2* gcol = Uc + j;
#pragma unroll 8
        for (int p = 0; p < d.P; ++p) {
            const float2 v = gcol[(size_t)p * B];
            s_col[(size_t)p * K4V2_COLS + tid] = v;
            energy += (double)v.x * v.x + (double)v.y * v.y;
        }
        if (!(energy > d.thresh)) find_id[(size_t)c * B + j] = -1;
    }
    s_energy[tid] = energy;
    const bool cand = (j < j_end) && (energy > d.thresh);
    const unsigned cbal = __ballot_sync(0xffffffffu, cand);
    if (lane == 0) s_wcnt[warp] = __popc(cbal);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < K4V2_THREADS / 32; ++w) {
        const int n = s_wcnt[w];
        base += (w < warp) ? n : 0;
        total += n;
    }
    if (cand) s_cand[base + __popc(cbal & ((1u << lane) - 1u))] = tid;
    __syncthreads();
    if (total == 0) return;

    // phase 2: one 8-lane group per candidate bin
    const int grp = tid / K4V2_G, gl = tid % K4V2_G;
    const int nsym = d.P_src - 1;
    uint8_t* sym = s_sym + (size_t)grp * (4 * NW);
    const int iters = (total + K4V2_THREADS / K4V2_G - 1) / (K4V2_THREADS / K4V2_G);
    unsigned n_multi = 0;
    long long wgt[32 / K4V2_G];                                          // hash digit weights of this lane (b <= 32)
#pragma unroll
    for (int u = 0; u < 32 / K4V2_G; ++u) wgt[u] = hash_weight(d, gl + u * K4V2_G);
    for (int it = 0; it < iters; ++it) {                                 // uniform trip count: the shuffles below use full masks
        const int ci = it * (K4V2_THREADS / K4V2_G) + grp;
        const bool valid = ci < total;
        const int my = s_cand[valid ? ci : 0];
        const long long jb = j0 + my;
        const float2* col = s_col + my;                                  // element p of the column = col[p * K4V2_COLS]

        for (int i = 1 + gl; i <= nsym; i += K4V2_G) sym[i - 1] = (uint8_t)detect_symbol(d, col, (size_t)K4V2_COLS, i);
        for (int i = nsym + gl; i < 4 * NW; i += K4V2_G) sym[i] = 0;     // identity source: k = the n = P_src - 1 symbols
        __syncwarp();
        uint32_t kw[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) kw[w] = reinterpret_cast<const uint32_t*>(sym)[w];

        double rr = 0.0, ri = 0.0;
        for (int p = gl; p < d.P; p += K4V2_G) {
            const int t = fast_mod(dot_raw_s<NW>(s_D + (size_t)p * rs, d.ld, kw), d.q, d.qmagic);
            const double cs = s_tw[t].x, sn = s_tw[t].y;
            const float2 v = col[(size_t)p * K4V2_COLS];
            rr += cs * v.x + sn * v.y;
            ri += cs * v.y - sn * v.x;
        }
        long long part = 0;
#pragma unroll
        for (int u = 0; u < 32 / K4V2_G; ++u) {
            const int i = gl + u * K4V2_G;
            if (i < d.b) part += wgt[u] * fast_mod(dot_raw_s<NW>(s_MT + (size_t)i * rs, d.ld, kw), d.q, d.qmagic);
        }
#pragma unroll
        for (int o = K4V2_G / 2; o > 0; o >>= 1) {                       // xor partners stay inside the aligned 8-lane group
            rr += __shfl_xor_sync(0xffffffffu, rr, o);
            ri += __shfl_xor_sync(0xffffffffu, ri, o);
            part += __shfl_xor_sync(0xffffffffu, part, o);
        }
        rr *= d.invP;
        ri *= d.invP;
        const double res = s_energy[my] - (double)d.P * (rr * rr + ri * ri);
        const bool single = valid && (part == jb) && !(res > d.thresh);
        const bool lead = gl == 0;

        const unsigned sb = __ballot_sync(0xffffffffu, lead && single);
        unsigned long long fbase = 0;
        if (lane == 0 && sb) fbase = atomicAdd(&counters[0], (unsigned long long)__popc(sb));
        fbase = __shfl_sync(0xffffffffu, fbase, 0);
        unsigned long long f = fbase + (unsigned long long)__popc(sb & ((1u << lane) - 1u));
        f = __shfl_sync(0xffffffffu, f, lane & ~(K4V2_G - 1));           // the leader's slot for its whole group
        if (single) {
            if ((long long)f < max_finds) {
                uint32_t* ko = reinterpret_cast<uint32_t*>(find_k + (size_t)f * d.ld);
                for (int w = gl; w < d.ld / 4; w += K4V2_G) ko[w] = (w < NW) ? reinterpret_cast<const uint32_t*>(sym)[w] : 0u;
                if (lead) {
                    find_cj[f] = (long long)c * B + jb;
                    find_rho[f] = make_float2((float)rr, (float)ri);
                    if (find_round) find_round[f] = round;
                    find_id[(size_t)c * B + jb] = (int32_t)f;
                }
            } else if (lead) {
                find_id[(size_t)c * B + jb] = -1;
            }
        } else if (valid && lead) {
            find_id[(size_t)c * B + jb] = -1;
            ++n_multi;
        }
        __syncwarp();                                                    // sym is rewritten in the next iteration
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_multi += __shfl_xor_sync(0xffffffffu, n_multi, o);
    if (lane == 0 && n_multi) atomicAdd(&counters[1], (unsigned long long)n_multi);
}

// ---------------------------------------------------------------------------------------------------------
// apply: one warp per find
// ---------------------------------------------------------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(K4_THREADS)
k4_apply_kernel(PeelDev d, float2* __restrict__ U, long long j_begin, long long j_end,
                const long long* __restrict__ find_cj, const int8_t* __restrict__ find_k,
                const float2* __restrict__ find_rho, const int32_t* __restrict__ find_id, long long f_begin,
                long long n_finds, long long id_limit, int dedupe, unsigned long long* __restrict__ owner_count) {
    const int lane = threadIdx.x & 31;
    const long long f = f_begin + (long long)blockIdx.x * (K4_THREADS / 32) + (threadIdx.x >> 5);
    if (f >= f_begin + n_finds) return;
    const long long cj = find_cj[f];
    const int c = (int)(cj / d.B);
    const long long wgt = hash_weight(d, lane);
    uint32_t kw[NW];
    const uint32_t* kin = reinterpret_cast<const uint32_t*>(find_k + (size_t)f * d.ld);
#pragma unroll
    for (int w = 0; w < NW; ++w) kw[w] = (w < d.ld / 4) ? kin[w] : 0u;

    if (dedupe) {
        // ball_values "last (i, j) wins" (qsft.py:215): skip when a higher group found the same k this round
        for (int c2 = c + 1; c2 < d.C; ++c2) {
            const long long j2 = hash_bin_warp<NW>(d, c2, kw, lane, wgt);
            const int32_t f2 = find_id[(size_t)c2 * d.B + j2];
            if (f2 >= 0 && (long long)f2 < id_limit) {
                const uint32_t* k2 = reinterpret_cast<const uint32_t*>(find_k + (size_t)f2 * d.ld);
                bool same = true;
#pragma unroll
                for (int w = 0; w < NW; ++w) same &= ((w < d.ld / 4) ? k2[w] : 0u) == kw[w];
                if (same) return;
            }
        }
    }
    if (owner_count && lane == 0) atomicAdd(owner_count, 1ull);   // distinct balls peeled (num_peeling, qsft.py:224)
    const float2 rho = find_rho[f];
    const float inv_q = 1.0f / (float)d.q;
    for (int l = 0; l < d.C; ++l) {
        const long long j = hash_bin_warp<NW>(d, l, kw, lane, wgt);
        if (j < j_begin || j >= j_end) continue;
        float2* Ul = U + (size_t)l * d.P * d.B + j;
        const int8_t* Dl = d.D + (size_t)l * d.P * d.ld;
        for (int p = lane; p < d.P; p += 32) {
            const int t = fast_mod(dot_raw<NW>(Dl + (size_t)p * d.ld, d.ld, kw), d.q, d.qmagic);
            float sn, cs;
            sincospif(2.0f * (float)t * inv_q, &sn, &cs);
            // rho * w^t
            const float vr = rho.x * cs - rho.y * sn;
            const float vi = rho.x * sn + rho.y * cs;
            float* dst = reinterpret_cast<float*>(Ul + (size_t)p * d.B);
            atomicAdd(dst, -vr);
            atomicAdd(dst + 1, -vi);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// reduce: collapse the finds of one round into the list of distinct k (one thread per find).
// The reference records every singleton (k, rho) and finally averages all rho recorded for the same k
// (qsft.py:209-217, 247-255).  Duplicates of a k inside a round sit in the bins the other groups hash k to, so the
// round's FIRST find of k (lowest group = first in (i, j) order) gathers them through find_id; re-finds of k in a
// later round are merged through a chained hash table keyed by k's group-0 bin (seen0 = chain heads, unext = links).
// ---------------------------------------------------------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(K4_THREADS)
k4_reduce_kernel(PeelDev d, const long long* __restrict__ find_cj, const int8_t* __restrict__ find_k,
                 const float2* __restrict__ find_rho, const int32_t* __restrict__ find_id, long long f_begin,
                 long long n_finds, long long id_limit, int round, int32_t* __restrict__ seen0,
                 int8_t* __restrict__ uk, float* __restrict__ usum, int32_t* __restrict__ ucnt,
                 long long* __restrict__ ukey, int32_t* __restrict__ unext, long long max_uniq,
                 unsigned long long* __restrict__ counters) {
    const long long f = f_begin + (long long)blockIdx.x * K4_THREADS + threadIdx.x;
    if (f >= f_begin + n_finds) return;
    const long long cj = find_cj[f];
    const int c = (int)(cj / d.B);
    const int nw = d.ld / 4;
    uint32_t kw[NW];
    const uint32_t* kin = reinterpret_cast<const uint32_t*>(find_k + (size_t)f * d.ld);
#pragma unroll
    for (int w = 0; w < NW; ++w) kw[w] = (w < nw) ? kin[w] : 0u;
    float2 sum = find_rho[f];
    int cnt = 1;
    long long j0 = cj - (long long)c * d.B;
    for (int c2 = 0; c2 < d.C; ++c2) {
        if (c2 == c) continue;
        const long long j2 = hash_bin<NW>(d, c2, kw);
        if (c2 == 0) j0 = j2;
        const int32_t f2 = find_id[(size_t)c2 * d.B + j2];
        if (f2 >= 0 && (long long)f2 < id_limit) {
            const uint32_t* k2 = reinterpret_cast<const uint32_t*>(find_k + (size_t)f2 * d.ld);
            bool same = true;
#pragma unroll
            for (int w = 0; w < NW; ++w) same &= ((w < nw) ? k2[w] : 0u) == kw[w];
            if (same) {
                if (c2 < c) return;              // an earlier group holds the round's first find of this k
                const float2 r2 = find_rho[f2];
                sum.x += r2.x;
                sum.y += r2.y;
                ++cnt;
            }
        }
    }
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

// ---------------------------------------------------------------------------------------------------------
// stand-alone detectors (the reference's public reconstruct.singleton_detection for a batch of columns): one warp
// per column, lanes over the symbols exactly as in phase 2 of the classification.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(K4_THREADS)
k4_detect_kernel(PeelDev d, const float2* __restrict__ cols, long long N, int8_t* __restrict__ k_out, int ld_out) {
    __shared__ __align__(16) uint8_t s_sym[K4_THREADS / 32][QSFT_MAX_N];
    __shared__ __align__(16) uint8_t s_k[K4_THREADS / 32][QSFT_MAX_N];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long c = (long long)blockIdx.x * (K4_THREADS / 32) + warp;
    if (c >= N) return;                                     // whole warps leave together
    const float2* col = cols + (size_t)c * d.P;             // element p of the column = col[p]
    const int nsym = d.P_src - 1;
    for (int i = 1 + lane; i <= nsym; i += 32) s_sym[warp][i - 1] = (uint8_t)detect_symbol(d, col, 1, i);
    __syncwarp();
    int nout = nsym;
    const uint8_t* src = s_sym[warp];
    if (d.source == 1) {
        if (lane == 0) rs_decode(d, s_sym[warp], s_k[warp]);
        __syncwarp();
        nout = d.n;
        src = s_k[warp];
    }
    int8_t* ko = k_out + (size_t)c * ld_out;
    for (int i = lane; i < ld_out; i += 32) ko[i] = i < nout ? (int8_t)src[i] : (int8_t)0;
}

// singleton_detection_mle (reconstruct.py:54-84): one warp per column, lanes over the K candidate signatures.
//   alpha_k = <S_k, col> / P,  residual_k = || col - alpha_k S_k ||_2,  k_sel = first minimum.  fp64 accumulation.
__global__ void __launch_bounds__(K4_THREADS)
k4_mle_kernel(const float2* __restrict__ cols, long long N, int P, const float2* __restrict__ S, int K,
              int32_t* __restrict__ k_sel, float* __restrict__ residual) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long c = (long long)blockIdx.x * (K4_THREADS / 32) + warp;
    if (c >= N) return;
    const float2* col = cols + (size_t)c * P;
    double best = 1.0e300;                                   // > any residual of finite fp32 data
    int best_k = 0x7fffffff;
    const double invP = 1.0 / (double)P;
    for (int k = lane; k < K; k += 32) {
        double ar = 0.0, ai = 0.0;
        for (int p = 0; p < P; ++p) {
            const float2 sg = S[(size_t)p * K + k];
            const float2 v = col[p];
            ar += (double)sg.x * v.x + (double)sg.y * v.y;  // conj(sg) * v
            ai += (double)sg.x * v.y - (double)sg.y * v.x;
        }
        ar *= invP;
        ai *= invP;
        double r2 = 0.0;
        for (int p = 0; p < P; ++p) {
            const float2 sg = S[(size_t)p * K + k];
            const float2 v = col[p];
            const double er = (double)v.x - (ar * sg.x - ai * sg.y);
            const double ei = (double)v.y - (ar * sg.y + ai * sg.x);
            r2 += er * er + ei * ei;
        }
        const double r = sqrt(r2);
        if (r < best) {                                      // ascending k per lane: strict < keeps the first minimum
            best = r;
            best_k = k;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
        if (ob < best || (ob == best && ok < best_k)) {
            best = ob;
            best_k = ok;
        }
    }
    if (lane == 0) {
        k_sel[c] = best_k;
        if (residual) residual[c] = (float)best;
    }
}

__global__ void k4_closed_form_kernel(const int8_t* __restrict__ MT, const int8_t* __restrict__ D, int q, int n, int b,
                                      int P, long long B, int ld, const int8_t* __restrict__ loc,
                                      const float2* __restrict__ a, long long S, float2* __restrict__ U) {
    const int lane = threadIdx.x & 31;
    const long long s = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (s >= S) return;
    const int8_t* k = loc + (size_t)s * ld;
    long long j = 0;
    for (int i = 0; i < b; ++i) {
        int acc = 0;
        for (int u = 0; u < n; ++u) acc += (int)MT[(size_t)i * ld + u] * (int)k[u];
        j = j * q + acc % q;
    }
    const float2 as = a[s];
    for (int p = lane; p < P; p += 32) {
        int acc = 0;
        for (int u = 0; u < n; ++u) acc += (int)D[(size_t)p * ld + u] * (int)k[u];
        float sn, cs;
        sincospif(2.0f * (float)(acc % q) / (float)q, &sn, &cs);
        float* dst = reinterpret_cast<float*>(U + (size_t)p * B + j);
        atomicAdd(dst, as.x * cs - as.y * sn);
        atomicAdd(dst + 1, as.x * sn + as.y * cs);
    }
}

int make_dev(const qsft_peel_desc* h, PeelDev* d) {
    QSFT_CHECK_ARG(h != nullptr, "null descriptor");
    QSFT_CHECK_ARG(h->q >= 2 && h->q <= QSFT_MAX_Q, "q=%d out of range", h->q);
    QSFT_CHECK_ARG(h->n >= 1 && h->n <= QSFT_MAX_N, "n=%d out of range", h->n);
    QSFT_CHECK_ARG(h->b >= 1 && h->b <= QSFT_MAX_B, "b=%d out of range", h->b);
    QSFT_CHECK_ARG(h->C >= 1 && h->C <= 65535, "C=%d out of range", h->C);
    QSFT_CHECK_ARG(h->P_src >= 2 && h->P >= h->P_src && h->P % h->P_src == 0, "P=%d must be a positive multiple of P_src=%d", h->P, h->P_src);
    QSFT_CHECK_ARG(h->channel >= 0 && h->channel <= 2, "channel must be 0 (identity), 1 (nso1) or 2 (nso2)");
    QSFT_CHECK_ARG(h->source == 0 || h->source == 1, "source must be 0 (identity) or 1 (coded)");
    QSFT_CHECK_ARG(h->ld >= h->n && h->ld % 16 == 0 && h->ld <= 128, "ld=%d must be >= n, a multiple of 16 and <= 128", h->ld);
    QSFT_CHECK_ARG(h->MT && h->D, "null M/D");
    if (h->channel == 0) QSFT_CHECK_ARG(h->P == h->P_src, "identity channel decoding needs num_repeat == 1");
    if (h->source == 0) {
        QSFT_CHECK_ARG(h->P_src - 1 == h->n, "identity source decoding needs P_src = n + 1 (got P_src=%d, n=%d)", h->P_src, h->n);
    } else {
        QSFT_CHECK_ARG(h->rs_t >= 1 && 2 * h->rs_t <= RS_MAX_2T && h->rs_s >= 1, "bad Reed-Solomon parameters");
        QSFT_CHECK_ARG(h->P_src - 1 == 2 * h->rs_t * h->rs_s, "coded source needs P_src = 2ts + 1");
        QSFT_CHECK_ARG(h->P_src - 1 <= QSFT_MAX_N, "too many syndrome symbols");
        QSFT_CHECK_ARG(h->rs_exp && h->rs_log, "null GF tables");
    }
    d->q = h->q; d->n = h->n; d->b = h->b; d->C = h->C; d->P = h->P; d->P_src = h->P_src; d->R = h->P / h->P_src;
    d->channel = h->channel; d->source = h->source; d->rs_t = h->rs_t; d->rs_s = h->rs_s; d->ld = h->ld;
    d->B = ipow64(h->q, h->b);
    d->thresh = (double)h->cutoff * (double)h->P;
    d->invP = 1.0 / (double)h->P;
    d->qmagic = (unsigned int)(((1ull << 32) + h->q - 1) / h->q);
    d->MT = h->MT; d->D = h->D; d->rs_exp = h->rs_exp; d->rs_log = h->rs_log;
    d->rs_order = h->source ? (int)ipow64(h->q, h->rs_s) : 0;
    const char* fd = getenv("QSFT_K4_FASTDET");          // opt-in, unmeasured: see quadrant_symbol
    d->fastdet = (fd && atoi(fd) != 0) ? 1 : 0;
    return QSFT_OK;
}

template <int NW>
int classify_nw(const PeelDev& d, const float2* U, long long jb, long long je, long long* cj, int8_t* fk, float2* rho,
                int32_t* frd, int32_t* fid, long long maxf, int round, unsigned long long* counters, cudaStream_t st) {
    dim3 grid((unsigned)((je - jb + K4_THREADS - 1) / K4_THREADS), (unsigned)d.C);
    k4_classify_kernel<NW><<<grid, K4_THREADS, 0, st>>>(d, U, jb, je, cj, fk, rho, frd, fid, maxf, round, counters);
    QSFT_LAUNCHED();
    return QSFT_OK;
}

template <int NW>
int apply_nw(const PeelDev& d, float2* U, long long jb, long long je, const long long* cj, const int8_t* fk,
             const float2* rho, const int32_t* fid, long long f_begin, long long nf, long long id_limit, int dedupe,
             unsigned long long* owners, cudaStream_t st) {
    const int wpb = K4_THREADS / 32;
    k4_apply_kernel<NW><<<(unsigned)((nf + wpb - 1) / wpb), K4_THREADS, 0, st>>>(d, U, jb, je, cj, fk, rho, fid, f_begin,
                                                                               nf, id_limit, dedupe, owners);
    QSFT_LAUNCHED();
    return QSFT_OK;
}

struct UniqOut {
    int32_t* seen0;       // (B) chain heads, zero initialised by the caller / qsft_peel
    int8_t* uk;           // (max_uniq, ld)
    float* usum;          // (max_uniq) complex64: sum of rho over all finds of the k
    int32_t* ucnt;        // (max_uniq) number of finds
    long long* ukey;      // (max_uniq) (round << 48) | (c * B + j) of the first find: reference's first-seen order
    int32_t* unext;       // (max_uniq) workspace
    long long max_uniq;
};

template <int NW>
int reduce_nw(const PeelDev& d, const long long* cj, const int8_t* fk, const float2* rho, const int32_t* fid,
              long long f_begin, long long nf, long long id_limit, int round, const UniqOut& o,
              unsigned long long* counters, cudaStream_t st) {
    k4_reduce_kernel<NW><<<(unsigned)((nf + K4_THREADS - 1) / K4_THREADS), K4_THREADS, 0, st>>>(
        d, cj, fk, rho, fid, f_begin, nf, id_limit, round, o.seen0, o.uk, o.usum, o.ucnt, o.ukey, o.unext, o.max_uniq, counters);
    QSFT_LAUNCHED();
    return QSFT_OK;
}

#define QSFT_NW_DISPATCH(fn, ...)                          \
    do {                                                   \
        const int nw__ = d.ld / 4;                         \
        if (nw__ <= 4) return fn<4>(__VA_ARGS__);          \
        if (nw__ <= 8) return fn<8>(__VA_ARGS__);          \
        if (nw__ <= 16) return fn<16>(__VA_ARGS__);        \
        return fn<32>(__VA_ARGS__);                        \
    } while (0)

template <int NW>
int classify_v2_nw(const PeelDev& d, const float2* U, long long jb, long long je, long long* cj, int8_t* fk, float2* rho,
                   int32_t* frd, int32_t* fid, long long maxf, int round, unsigned long long* counters, cudaStream_t st) {
    const size_t smem = k4v2_smem_bytes(d, NW);
    static size_t configured = 0;                            // per template instance
    if (smem > configured) {
        QSFT_CUDA(cudaFuncSetAttribute(k4_classify_v2_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((unsigned)((je - jb + K4V2_THREADS - 1) / K4V2_THREADS), (unsigned)d.C);
    k4_classify_v2_kernel<NW><<<grid, K4V2_THREADS, smem, st>>>(d, U, jb, je, cj, fk, rho, frd, fid, maxf, round, counters);
    QSFT_LAUNCHED();
    return QSFT_OK;
}

// QSFT_K4_IMPL=2 selects the shared-memory / 8-lane-group classification when the problem fits it (identity source
// decoding, tile <= 100 KB of shared memory so that two CTAs stay resident); read on every call.
static bool k4_use_v2(const PeelDev& d) {
    const char* e = getenv("QSFT_K4_IMPL");
    if (!e || atoi(e) != 2 || d.source != 0) return false;
    const int nw = d.ld / 4 <= 4 ? 4 : d.ld / 4 <= 8 ? 8 : d.ld / 4 <= 16 ? 16 : 32;
    return k4v2_smem_bytes(d, nw) <= 100 * 1024;
}

int do_classify(const PeelDev& d, const float2* U, long long jb, long long je, long long* cj, int8_t* fk, float2* rho,
                int32_t* frd, int32_t* fid, long long maxf, int round, unsigned long long* counters, cudaStream_t st) {
    if (k4_use_v2(d)) QSFT_NW_DISPATCH(classify_v2_nw, d, U, jb, je, cj, fk, rho, frd, fid, maxf, round, counters, st);
    QSFT_NW_DISPATCH(classify_nw, d, U, jb, je, cj, fk, rho, frd, fid, maxf, round, counters, st);
}

int do_reduce(const PeelDev& d, const long long* cj, const int8_t* fk, const float2* rho, const int32_t* fid,
              long long f_begin, long long nf, long long id_limit, int round, const UniqOut& o,
              unsigned long long* counters, cudaStream_t st) {
    QSFT_NW_DISPATCH(reduce_nw, d, cj, fk, rho, fid, f_begin, nf, id_limit, round, o, counters, st);
}

int do_apply(const PeelDev& d, float2* U, long long jb, long long je, const long long* cj, const int8_t* fk,
             const float2* rho, const int32_t* fid, long long f_begin, long long nf, long long id_limit, int dedupe,
             unsigned long long* owners, cudaStream_t st) {
    QSFT_NW_DISPATCH(apply_nw, d, U, jb, je, cj, fk, rho, fid, f_begin, nf, id_limit, dedupe, owners, st);
}

}  // namespace

extern "C" int qsft_peel_classify(const qsft_peel_desc* h, const float* U, int64_t j_begin, int64_t j_end,
                                  int64_t* find_cj, int8_t* find_k, float* find_rho, int32_t* find_round,
                                  int32_t* find_id, int64_t max_finds, int round, unsigned long long* counters,
                                  void* stream) {
    PeelDev d;
    if (int rc = make_dev(h, &d)) return rc;
    QSFT_CHECK_ARG(U && find_cj && find_k && find_rho && find_id && counters, "null pointer");
    QSFT_CHECK_ARG(0 <= j_begin && j_begin <= j_end && j_end <= d.B, "bad bin range");
    if (j_begin == j_end) return QSFT_OK;
    return do_classify(d, reinterpret_cast<const float2*>(U), j_begin, j_end, (long long*)find_cj, find_k,
                       reinterpret_cast<float2*>(find_rho), find_round, find_id, max_finds, round, counters,
                       (cudaStream_t)stream);
}

extern "C" int qsft_peel_apply(const qsft_peel_desc* h, float* U, int64_t j_begin, int64_t j_end,
                               const int64_t* find_cj, const int8_t* find_k, const float* find_rho,
                               const int32_t* find_id, int64_t f_begin, int64_t n_finds, int dedupe,
                               unsigned long long* owner_count, void* stream) {
    PeelDev d;
    if (int rc = make_dev(h, &d)) return rc;
    QSFT_CHECK_ARG(U && find_cj && find_k && find_rho && (find_id || !dedupe), "null pointer");
    QSFT_CHECK_ARG(0 <= j_begin && j_begin <= j_end && j_end <= d.B, "bad bin range");
    QSFT_CHECK_ARG(f_begin >= 0, "negative f_begin");
    if (n_finds <= 0) return QSFT_OK;
    return do_apply(d, reinterpret_cast<float2*>(U), j_begin, j_end, (const long long*)find_cj, find_k,
                    reinterpret_cast<const float2*>(find_rho), find_id, f_begin, n_finds, f_begin + n_finds, dedupe,
                    owner_count, (cudaStream_t)stream);
}

extern "C" int qsft_peel_reduce(const qsft_peel_desc* h, const int64_t* find_cj, const int8_t* find_k,
                                const float* find_rho, const int32_t* find_id, int64_t f_begin, int64_t n_finds,
                                int round, const qsft_uniq* uq, unsigned long long* counters, void* stream) {
    PeelDev d;
    if (int rc = make_dev(h, &d)) return rc;
    QSFT_CHECK_ARG(find_cj && find_k && find_rho && find_id && uq && counters, "null pointer");
    QSFT_CHECK_ARG(uq->seen0 && uq->uniq_k && uq->uniq_sum && uq->uniq_cnt && uq->uniq_key && uq->uniq_next && uq->max_uniq > 0,
                   "incomplete qsft_uniq");
    QSFT_CHECK_ARG(round >= 1 && round < 32768 && f_begin >= 0, "bad round / f_begin");
    if (n_finds <= 0) return QSFT_OK;
    UniqOut o{uq->seen0, uq->uniq_k, uq->uniq_sum, uq->uniq_cnt, (long long*)uq->uniq_key, uq->uniq_next, uq->max_uniq};
    return do_reduce(d, (const long long*)find_cj, find_k, reinterpret_cast<const float2*>(find_rho), find_id, f_begin,
                     n_finds, f_begin + n_finds, round, o, counters, (cudaStream_t)stream);
}

extern "C" int qsft_peel(const qsft_peel_desc* h, float* U, int64_t* find_cj, int8_t* find_k, float* find_rho,
                         int32_t* find_round, int32_t* find_id, int64_t max_finds, unsigned long long* counters,
                         const qsft_uniq* uq, int64_t* n_finds_out, int64_t* n_uniq_out, int* n_rounds_out, void* stream) {
    PeelDev d;
    if (int rc = make_dev(h, &d)) return rc;
    QSFT_CHECK_ARG(U && find_cj && find_k && find_rho && find_round && find_id && counters && n_finds_out && n_rounds_out,
                   "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    UniqOut uo{};
    if (uq) {
        QSFT_CHECK_ARG(uq->seen0 && uq->uniq_k && uq->uniq_sum && uq->uniq_cnt && uq->uniq_key && uq->uniq_next && uq->max_uniq > 0,
                       "incomplete qsft_uniq");
        uo = UniqOut{uq->seen0, uq->uniq_k, uq->uniq_sum, uq->uniq_cnt, (long long*)uq->uniq_key, uq->uniq_next, uq->max_uniq};
    }
    // `num_peeling < q ** n` (qsft.py:151) can only bind when q^n is tiny: at most C*B balls are peeled per round
    const double peeling_max = pow((double)d.q, (double)d.n);
    const bool guard_can_bind = peeling_max <= 15.0 * (double)d.C * (double)d.B;
    // pinned read-back slot for the round counters: allocated once per host thread (cudaMallocHost costs milliseconds)
    static thread_local unsigned long long* host = nullptr;
    if (!host) QSFT_CUDA(cudaMallocHost(&host, 4 * sizeof(unsigned long long)));
    long long total = 0;
    double num_peeling = 0;
    int round = 0;
    int rc = QSFT_OK;
    cudaError_t ce = cudaSuccess;
    bool cont = true;
    // counters: [0] finds (running), [1] multitons of the round, [2] distinct balls peeled (running), [4] distinct k
    ce = cudaMemsetAsync(counters, 0, 6 * sizeof(unsigned long long), st);
    if (ce == cudaSuccess && uq) ce = cudaMemsetAsync(uo.seen0, 0, (size_t)d.B * sizeof(int32_t), st);
    while (ce == cudaSuccess && rc == QSFT_OK && cont && num_peeling < peeling_max && round < 15) {
        ++round;
        if ((ce = cudaMemsetAsync(counters + 1, 0, sizeof(unsigned long long), st)) != cudaSuccess) break;
        if ((rc = do_classify(d, reinterpret_cast<const float2*>(U), 0, d.B, (long long*)find_cj, find_k,
                              reinterpret_cast<float2*>(find_rho), find_round, find_id, max_finds, round, counters, st)) != 0) break;
        if ((ce = cudaMemcpyAsync(host, counters, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
        if ((ce = cudaStreamSynchronize(st)) != cudaSuccess) break;
        const long long now = (long long)host[0];
        const long long multis = (long long)host[1];
        if (now > max_finds) {
            qsft_set_error("find buffer too small: %lld finds > max_finds=%lld", now, (long long)max_finds);
            return QSFT_EINVAL;
        }
        const long long nf = now - total;
        if (multis == 0 || nf == 0) cont = false;          // qsft.py:204-205
        if (nf > 0 && uq) {
            if ((rc = do_reduce(d, (const long long*)find_cj, find_k, reinterpret_cast<const float2*>(find_rho), find_id,
                                total, nf, now, round, uo, counters, st)) != 0) break;
        }
        // the reference also subtracts after its last round, but nothing reads U afterwards: skip unless the q^n guard needs the count
        if (nf > 0 && (cont || guard_can_bind)) {
            if ((rc = do_apply(d, reinterpret_cast<float2*>(U), 0, d.B, (const long long*)find_cj, find_k,
                               reinterpret_cast<const float2*>(find_rho), find_id, total, nf, now, 1, counters + 2, st)) != 0) break;
            if (guard_can_bind) {
                if ((ce = cudaMemcpyAsync(host + 2, counters + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
                if ((ce = cudaStreamSynchronize(st)) != cudaSuccess) break;
                num_peeling = (double)host[2];
            }
        }
        total = now;
    }
    if (ce != cudaSuccess) {
        qsft_set_error("CUDA error in peel loop: %s", cudaGetErrorString(ce));
        return QSFT_ECUDA;
    }
    if (rc != QSFT_OK) return rc;
    if (uq && n_uniq_out) {
        if ((ce = cudaMemcpyAsync(host, counters + 4, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st)) == cudaSuccess)
            ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) {
            qsft_set_error("CUDA error reading the unique count: %s", cudaGetErrorString(ce));
            return QSFT_ECUDA;
        }
        if ((long long)host[0] > uo.max_uniq) {
            qsft_set_error("unique buffer too small: %llu > %lld", host[0], uo.max_uniq);
            return QSFT_EINVAL;
        }
        *n_uniq_out = (int64_t)host[0];
    }
    *n_finds_out = total;
    *n_rounds_out = round;
    return QSFT_OK;
}

extern "C" int qsft_singleton_detect(const float* cols, int64_t N, int q, int n, int P, int P_src, int channel, int source,
                                     int rs_t, int rs_s, const int32_t* rs_exp, const int32_t* rs_log, int8_t* k_out,
                                     int ld_out, void* stream) {
    QSFT_CHECK_ARG(N >= 0 && N <= 0x7fffffffll * (K4_THREADS / 32), "N=%lld out of range", (long long)N);
    QSFT_CHECK_ARG(N == 0 || (cols && k_out), "null pointer");
    QSFT_CHECK_ARG(q >= 2 && q <= QSFT_MAX_Q, "q=%d out of range", q);
    QSFT_CHECK_ARG(P_src >= 2 && P_src - 1 <= QSFT_MAX_N && P >= P_src && P % P_src == 0,
                   "P=%d must be a positive multiple of P_src=%d (2 <= P_src <= %d)", P, P_src, QSFT_MAX_N + 1);
    QSFT_CHECK_ARG(channel >= 0 && channel <= 2, "channel must be 0 (identity), 1 (nso1) or 2 (nso2)");
    QSFT_CHECK_ARG(source == 0 || source == 1, "source must be 0 (identity) or 1 (coded)");
    if (channel == 0) QSFT_CHECK_ARG(P == P_src, "identity channel decoding needs num_repeat == 1");
    PeelDev d{};
    d.q = q; d.n = source ? n : P_src - 1; d.P = P; d.P_src = P_src; d.R = P / P_src; d.channel = channel; d.source = source;
    if (const char* fd = getenv("QSFT_K4_FASTDET")) d.fastdet = atoi(fd) != 0 ? 1 : 0;
    if (source == 1) {
        QSFT_CHECK_ARG(n >= 1 && n <= QSFT_MAX_N, "n=%d out of range", n);
        QSFT_CHECK_ARG(rs_t >= 1 && 2 * rs_t <= RS_MAX_2T && rs_s >= 1 && P_src - 1 == 2 * rs_t * rs_s, "coded source needs P_src = 2ts + 1");
        QSFT_CHECK_ARG(rs_exp && rs_log, "null GF tables");
        d.rs_t = rs_t; d.rs_s = rs_s; d.rs_exp = rs_exp; d.rs_log = rs_log; d.rs_order = (int)ipow64(q, rs_s);
    }
    QSFT_CHECK_ARG(ld_out >= d.n, "ld_out=%d smaller than the %d output digits", ld_out, d.n);
    if (N == 0) return QSFT_OK;
    const int wpb = K4_THREADS / 32;
    k4_detect_kernel<<<(unsigned)((N + wpb - 1) / wpb), K4_THREADS, 0, (cudaStream_t)stream>>>(
        d, reinterpret_cast<const float2*>(cols), N, k_out, ld_out);
    QSFT_LAUNCHED();
    return QSFT_OK;
}

extern "C" int qsft_detect_mle(const float* cols, int64_t N, int P, const float* S, int K, int32_t* k_sel, float* residual,
                               void* stream) {
    QSFT_CHECK_ARG(N >= 0 && N <= 0x7fffffffll * (K4_THREADS / 32), "N=%lld out of range", (long long)N);
    QSFT_CHECK_ARG(N == 0 || (cols && S && k_sel), "null pointer");
    QSFT_CHECK_ARG(P >= 1 && K >= 1, "P=%d and K=%d must be positive", P, K);
    if (N == 0) return QSFT_OK;
    const int wpb = K4_THREADS / 32;
    k4_mle_kernel<<<(unsigned)((N + wpb - 1) / wpb), K4_THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(cols), N, P, reinterpret_cast<const float2*>(S), K, k_sel, residual);
    QSFT_LAUNCHED();
    return QSFT_OK;
}

extern "C" int qsft_closed_form_bins(const int8_t* MT, const int8_t* D, int q, int n, int b, int P, const int8_t* loc,
                                     int ld, const float* strengths, int64_t S, float* U, void* stream) {
    QSFT_CHECK_ARG(MT && D && loc && strengths && U, "null pointer");
    QSFT_CHECK_ARG(q >= 2 && q <= QSFT_MAX_Q && n >= 1 && n <= QSFT_MAX_N && b >= 1 && b <= QSFT_MAX_B && P >= 1, "bad shape");
    QSFT_CHECK_ARG(ld >= n && ld % 16 == 0, "bad ld");
    if (S <= 0) return QSFT_OK;
    const int wpb = 4;
    k4_closed_form_kernel<<<(unsigned)((S + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        MT, D, q, n, b, P, ipow64(q, b), ld, loc, reinterpret_cast<const float2*>(strengths), S,
        reinterpret_cast<float2*>(U));
    QSFT_LAUNCHED();
    return QSFT_OK;
}
