// K4: peeling decoder.  Replaces the round loop of QSFT.transform (qsft/qsft.py:151-241), the singleton detectors
// (qsft/reconstruct.py:12-31 noiseless, :100-113 nso1, :34-51 coded) and the Reed-Solomon syndrome decoder
// (qsft/ReedSolomon.py:26-48, arithmetic of galois' decode_jit restated: Berlekamp-Massey + Chien + Forney).
//
// Layout: U (C, P, B) complex64 with the bin index j contiguous.  Classification maps one THREAD to one bin with lanes
// over consecutive j, so every load of a delay row is a fully coalesced 256-byte warp request (HBM bound: one read of
// U per round).  Rounds are synchronous like the reference: classify on a frozen U, then subtract.
#include "common.cuh"

#include <stdlib.h>

#include "k4_shared.cuh"

namespace {

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
                const int symv = detect_symbol(d, StridedCol{col, (size_t)B, d.P_src}, i);
                s_sym[warp][i - 1] = (uint8_t)symv;
            }
        }
        __syncwarp();
        // k digits (zero padded to 4 * NW bytes)
        if (d.source == 1) {
            if (lane == 0) rs_decode(rs_params(d), s_sym[warp], s_k[warp]);
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
    k4_uniq_commit<NW>(d, kw, sum, cnt, cj, j0, round, seen0, uk, usum, ucnt, ukey, unext, max_uniq, counters);
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
    for (int i = 1 + lane; i <= nsym; i += 32) s_sym[warp][i - 1] = (uint8_t)detect_symbol(d, StridedCol{col, 1, d.P_src}, i);
    __syncwarp();
    int nout = nsym;
    const uint8_t* src = s_sym[warp];
    if (d.source == 1) {
        if (lane == 0) rs_decode(rs_params(d), s_sym[warp], s_k[warp]);
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
    const char* fd = getenv("QSFT_K4_FASTDET");          // quadrant detection (see quadrant_symbol) unless disabled
    d->fastdet = (fd && atoi(fd) == 0) ? 0 : 1;
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

int do_classify(const PeelDev& d, const float2* U, long long jb, long long je, long long* cj, int8_t* fk, float2* rho,
                int32_t* frd, int32_t* fid, long long maxf, int round, unsigned long long* counters, cudaStream_t st) {
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

namespace {
// distinct-k list -> the host's result layout, in the order `order` gives (sorted first-seen keys): compact digit rows
// (n bytes each, no padding), mean = sum / count in double (what the host did after the copy), counts
__global__ void __launch_bounds__(256)
k4_distinct_gather_kernel(const long long* __restrict__ order, long long nu, int n, int ld, const int8_t* __restrict__ uniq_k,
                          const float2* __restrict__ uniq_sum, const int* __restrict__ uniq_cnt, int8_t* __restrict__ k_out,
                          double2* __restrict__ mean_out, int* __restrict__ cnt_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one output digit
    if (i >= nu * n) return;
    const long long row = i / n;
    const int u = (int)(i - row * n);
    const long long src = order[row];
    k_out[i] = uniq_k[(size_t)src * ld + u];
    if (u == 0) {
        const float2 s = uniq_sum[src];
        const int c = uniq_cnt[src];
        mean_out[row] = make_double2((double)s.x / (double)c, (double)s.y / (double)c);
        cnt_out[row] = c;
    }
}
}  // namespace

extern "C" int qsft_peel_distinct(const qsft_uniq* uq, const int64_t* order, int64_t n_uniq, int n, int ld, int8_t* k_out,
                                  double* mean_out, int32_t* cnt_out, void* stream) {
    QSFT_CHECK_ARG(uq && uq->uniq_k && uq->uniq_sum && uq->uniq_cnt, "incomplete qsft_uniq");
    QSFT_CHECK_ARG(n_uniq >= 0 && n_uniq <= uq->max_uniq && n >= 1 && ld >= n, "bad shape");
    if (n_uniq == 0) return QSFT_OK;
    QSFT_CHECK_ARG(order && k_out && mean_out && cnt_out, "null pointer");
    QSFT_CHECK_ARG(((uintptr_t)mean_out & 15) == 0, "mean_out must be 16-byte aligned");
    const long long total = (long long)n_uniq * n;
    k4_distinct_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const long long*)order, (long long)n_uniq, n, ld, uq->uniq_k, reinterpret_cast<const float2*>(uq->uniq_sum), uq->uniq_cnt,
        k_out, reinterpret_cast<double2*>(mean_out), cnt_out);
    QSFT_LAUNCHED();
    return QSFT_OK;
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
    // default: the whole loop on the device (k4_peel_loop.cu; U is left untouched).  QSFT_K4_IMPL=1 keeps the host-driven
    // classify / reduce / apply rounds below as a cross-check (they modify U in place like the reference).
    {
        const char* impl = getenv("QSFT_K4_IMPL");
        if (!(impl && atoi(impl) == 1) && d.C * d.R <= 16) {
            const float* blocks[16];
            for (int c = 0; c < d.C; ++c)
                for (int r = 0; r < d.R; ++r) blocks[c * d.R + r] = U + 2 * ((size_t)c * d.P + (size_t)r * d.P_src) * (size_t)d.B;
            int64_t nu = 0;
            const int rc = qsft_peel_loop(d, blocks, d.B, find_cj, find_k, find_rho, find_round, find_id, max_finds, counters,
                                          uq ? &uo : nullptr, n_finds_out, &nu, n_rounds_out, st);
            if (rc != QSFT_EUNSUPPORTED) {
                if (rc == QSFT_OK && n_uniq_out) *n_uniq_out = nu;
                return rc;
            }
        }
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

extern "C" int qsft_peel_blocks(const qsft_peel_desc* h, const float* const* blocks, int64_t ldU, int64_t* find_cj, int8_t* find_k,
                                float* find_rho, int32_t* find_round, int32_t* find_id, int64_t max_finds,
                                unsigned long long* counters, const qsft_uniq* uq, int64_t* n_finds_out, int64_t* n_uniq_out,
                                int* n_rounds_out, void* stream) {
    PeelDev d;
    if (int rc = make_dev(h, &d)) return rc;
    QSFT_CHECK_ARG(blocks && find_cj && find_k && find_rho && find_round && find_id && counters, "null pointer");
    QSFT_CHECK_ARG((n_finds_out != nullptr) == (n_rounds_out != nullptr), "n_finds_out and n_rounds_out: both or neither (asynchronous)");
    QSFT_CHECK_ARG(ldU >= d.B, "ldU=%lld smaller than q^b=%lld", (long long)ldU, (long long)d.B);
    if (d.C * d.R > 16) return QSFT_EUNSUPPORTED;
    for (int i = 0; i < d.C * d.R; ++i) QSFT_CHECK_ARG(blocks[i] != nullptr, "null block pointer");
    UniqOut uo{};
    if (uq) {
        QSFT_CHECK_ARG(uq->seen0 && uq->uniq_k && uq->uniq_sum && uq->uniq_cnt && uq->uniq_key && uq->uniq_next && uq->max_uniq > 0,
                       "incomplete qsft_uniq");
        uo = UniqOut{uq->seen0, uq->uniq_k, uq->uniq_sum, uq->uniq_cnt, (long long*)uq->uniq_key, uq->uniq_next, uq->max_uniq};
    }
    int64_t nu = 0, nf_dummy = 0;
    const int rc = qsft_peel_loop(d, blocks, ldU, find_cj, find_k, find_rho, find_round, find_id, max_finds, counters,
                                  uq ? &uo : nullptr, n_finds_out ? n_finds_out : &nf_dummy, &nu, n_rounds_out, (cudaStream_t)stream);
    if (rc == QSFT_OK && n_uniq_out && n_rounds_out) *n_uniq_out = nu;
    return rc;
}

extern "C" int64_t qsft_peel_sharded_workspace_bytes(const qsft_peel_desc* h, int64_t max_finds) {
    PeelDev d;
    if (make_dev(h, &d) != QSFT_OK || max_finds <= 0) return -1;
    return qsft_peel_loop_workspace_bytes(d, max_finds);
}

extern "C" int qsft_peel_blocks_sharded(const qsft_peel_desc* h, const float* const* blocks, int64_t ldU, const qsft_shard* shard,
                                        int64_t max_finds, unsigned long long* counters, const qsft_uniq* uq, int64_t* n_uniq_out,
                                        int* n_rounds_out, void* stream) {
    PeelDev d;
    if (int rc = make_dev(h, &d)) return rc;
    QSFT_CHECK_ARG(blocks && shard && shard->peers && counters && uq, "null pointer");
    QSFT_CHECK_ARG((n_uniq_out != nullptr) == (n_rounds_out != nullptr), "n_uniq_out and n_rounds_out: both or neither (asynchronous)");
    QSFT_CHECK_ARG(shard->world >= 2 && shard->world <= 8 && shard->rank >= 0 && shard->rank < shard->world, "bad rank / world");
    QSFT_CHECK_ARG(ldU >= d.B && max_finds >= shard->world, "bad ldU / max_finds");
    if (d.C * d.R > 16) return QSFT_EUNSUPPORTED;
    for (int i = 0; i < d.C * d.R; ++i) QSFT_CHECK_ARG(blocks[i] != nullptr, "null block pointer");
    for (int p = 0; p < shard->world; ++p) QSFT_CHECK_ARG(shard->peers[p] != nullptr, "null workspace pointer");
    QSFT_CHECK_ARG(uq->seen0 && uq->uniq_k && uq->uniq_sum && uq->uniq_cnt && uq->uniq_key && uq->uniq_next && uq->max_uniq > 0,
                   "incomplete qsft_uniq");
    UniqOut uo{uq->seen0, uq->uniq_k, uq->uniq_sum, uq->uniq_cnt, (long long*)uq->uniq_key, uq->uniq_next, uq->max_uniq};
    KlShardHost sh{shard->rank, shard->world, shard->peers, shard->epoch};
    int64_t nf = 0, nu = 0;
    const int rc = qsft_peel_loop(d, blocks, ldU, nullptr, nullptr, nullptr, nullptr, nullptr, max_finds, counters, &uo, &nf, &nu,
                                  n_rounds_out, (cudaStream_t)stream, &sh);
    if (rc == QSFT_OK && n_uniq_out) *n_uniq_out = nu;
    return rc;
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
    d.fastdet = 1;
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
