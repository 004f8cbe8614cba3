// K4, default path: the whole peeling loop of QSFT.transform (qsft/qsft.py:151-255) in ONE persistent cooperative kernel --
// classification rounds, the reference's stop rule, duplicate averaging and "peeling" all run on the device, with two grid
// barriers per round and no host round trip.
//
// U is NEVER modified.  The reference subtracts every peeled ball from the bins it hashes to (qsft.py:223-241); here a peeled
// ball is LINKED into the (short) list of each of those bins instead (one atomic exchange per group), and a classification
// round subtracts the listed balls from its shared-memory copy of the bin before it looks at it.  That replaces
// 2 * C * P float atomics per ball on DRAM-resident data by C pointer swaps, needs no private copy of U, and every round
// starts from the original samples (no accumulated rounding).
//
// Classification of one round: every warp owns 16- or 32-bin tiles (all P delay rows of those bins), double buffered:
//   * TMA variant (row stride a multiple of 16 bytes): the tile is fetched with cp.async.bulk.tensor boxes of
//     {16 bins = 128 B, P_src rows} per (c, r) block with the 128-byte swizzle, so that both access patterns are free of
//     bank conflicts: lanes over bins (energy scan) and lanes over delay rows (detection, rho, residual);
//   * plain variant (odd q^b): the warp copies the tile with coalesced loads into the same layout.
//   Phase 1: energies (one lane per bin) and the bins' ball lists; bins with listed balls are updated in place by 8-lane
//   groups.  Phase 2: non-zeroton bins are handled four at a time by 8-lane groups: symbols (reconstruct.py:12-31,100-129),
//   optional Reed-Solomon decode, rho and residual (qsft.py:174-183), bin hash check (qsft.py:178-179).
// Link phase of a round: one thread per find: duplicate gathering / averaging exactly like k4_reduce_kernel, and the
// "last (i, j) wins" find of every k (qsft.py:215) links the ball into its C bins.
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

#include "k4_shared.cuh"
#ifndef QSFT_EMU
#include "tma.cuh"
#endif

namespace {

constexpr int KL_THREADS = 256;
constexpr int KL_MAX_BLOCKS = 16;            // (c, r) blocks of U addressed separately (C * R <= 16)
constexpr int KL_G = 8;                      // lanes per bin in the group phases

struct KlBlocks {
    const float2* p[KL_MAX_BLOCKS];          // block c * R + r: (P_src, ldU) complex64, bin index contiguous
};
#ifndef QSFT_EMU
struct KlMaps {
    CUtensorMap m[KL_MAX_BLOCKS];            // the same blocks as 2-D tensors {2 B floats, P_src rows}, box {32, P_src}
};
#endif

struct KlArgs {
    PeelDev d;
    long long ldU;                           // row stride of every block (elements)
    long long* find_cj;
    int8_t* find_k;
    float2* find_rho;
    int32_t* find_round;
    int32_t* find_id;                        // (C, B): written for singletons only; validated through find_cj
    long long max_finds;
    int32_t* head;                           // (C, B): last ball linked into the bin + 1 (0 = none); zeroed by the host
    int32_t* next;                           // (max_finds, C): previous ball of the same bin + 1
    UniqOut uo;
    int has_uniq;
    unsigned long long* counters;            // [0] finds, [2] balls peeled, [4] distinct k, [5] rounds, [6] error flags,
                                             // [7] finds kept
    unsigned long long* multi;               // [r] multitons of round r (1 <= r <= 15; workspace, zeroed by the host)
    unsigned int* gbar;                      // grid barrier counter (zeroed by the host)
    const int* dstruct;                      // device flag: D[c][r][i] = D[c][r][0] - e_{i-1} (identity / nso delays)
    int wpc;                                 // warps per CTA that classify (shared-memory budget)
    int bw;                                  // bins per warp tile: 16 or 32
    int sbox;                                // bytes per (half, repeat) sub-box of a tile: P_src * 128 rounded up to 1024
    int max_rounds;
    int guard_can_bind;
    double peeling_max;
    float rel_floor;                         // residual floor relative to the bin energy (fp32 resolution of U)
};

// ---- tile access -------------------------------------------------------------------------------------------------
// tile = [half h (16 bins)][repeat r][row i (128 B: 16 bins, 16-byte chunks xor-swizzled with i & 7)]
struct TileCol {
    float2* t;
    int R, sbox8;                            // sbox in float2 units
    int lb;                                  // local bin
    __device__ __forceinline__ int off(int r, int i) const {
        return ((lb >> 4) * R + r) * sbox8 + (i << 4) + (((((lb & 15) >> 1) ^ (i & 7)) << 1) | (lb & 1));
    }
    __device__ __forceinline__ float2 ri(int r, int i) const { return t[off(r, i)]; }
    __device__ __forceinline__ float2& ref(int r, int i) const { return t[off(r, i)]; }
};

__device__ __forceinline__ float kl_group_sum(float v) {
#pragma unroll
    for (int o = KL_G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void kl_grid_barrier(unsigned int* gbar, unsigned int& epoch) {
    __syncthreads();
    if (gridDim.x > 1) {
        if (threadIdx.x == 0) {
            ++epoch;
            __threadfence();
            atomicAdd(gbar, 1u);
            const unsigned int target = epoch * gridDim.x;
#ifndef QSFT_EMU
            unsigned int seen;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(gbar) : "memory");
                if (seen >= target) break;
                __nanosleep(32);
            }
#else
            (void)target;
#endif
            __threadfence();
        }
        __syncthreads();
    }
}

// phase t = <D[c][r * P_src + i], k> mod q for the k whose digits sit in kb (bytes) / kw (words)
template <int NW>
struct KlPhase {
    const PeelDev& d;
    const int8_t* Dc;                         // D rows of group c
    const uint8_t* kb;
    const uint32_t (&kw)[NW];
    bool structured;
    __device__ __forceinline__ int base(int r) const {            // phase of row (r, 0)
        return fast_mod(dot_raw<NW>(Dc + (size_t)(r * d.P_src) * d.ld, d.ld, kw), d.q, d.qmagic);
    }
    __device__ __forceinline__ int row(int r, int i, int tbase) const {
        if (i == 0) return tbase;
        if (structured) {
            const int t = tbase - (int)kb[i - 1];
            return t < 0 ? t + d.q : t;
        }
        return fast_mod(dot_raw<NW>(Dc + (size_t)(r * d.P_src + i) * d.ld, d.ld, kw), d.q, d.qmagic);
    }
};

// ---- one classification round ---------------------------------------------------------------------------------------
template <int NW, bool TMA>
__device__ __forceinline__ void kl_classify(const KlArgs& a, const KlBlocks& blk,
#ifndef QSFT_EMU
                                            const CUtensorMap* maps,
#endif
                                            int round, uint8_t* wsm, uint64_t* bars, unsigned int& tiles_done,
                                            const float2* s_tw) {
    const PeelDev& d = a.d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / KL_G, gl = lane % KL_G;
    const int R = d.R, P_src = d.P_src, bw = a.bw, nh = bw >> 4;
    const int sbox8 = a.sbox >> 3;
    const int tile_bytes = nh * R * a.sbox;
    float2* tiles[2] = {reinterpret_cast<float2*>(wsm), reinterpret_cast<float2*>(wsm + tile_bytes)};
    uint8_t* s_sym = wsm + 2 * tile_bytes;                         // [4 groups][2][QSFT_MAX_N]: symbols, decoded k
    float* s_e = reinterpret_cast<float*>(s_sym + 4 * 2 * QSFT_MAX_N);
    const long long B = d.B;
    const long long tpg = (B + bw - 1) / bw;                       // tiles per group
    const long long n_tiles = tpg * d.C;
    const long long gw = (long long)warp * gridDim.x + blockIdx.x, GW = (long long)a.wpc * gridDim.x;
    const float thresh = (float)d.thresh;
    const int nsym = P_src - 1;
    const bool structured = (*a.dstruct != 0);
    unsigned n_multi = 0;
    long long wgt[32 / KL_G];
#pragma unroll
    for (int u = 0; u < 32 / KL_G; ++u) wgt[u] = hash_weight(d, gl + u * KL_G);

    auto issue = [&](long long tt, int stage) {
#ifndef QSFT_EMU
        if constexpr (TMA) {
            if (lane == 0) {
                const int c = (int)(tt / tpg);
                const long long j0 = (tt - (long long)c * tpg) * bw;
                int halves = 0;
                for (int h = 0; h < nh; ++h) halves += (j0 + 16 * h < B) ? 1 : 0;
                tma::mbar_expect_tx(&bars[stage], (uint32_t)(halves * R * P_src * 128));
                for (int h = 0; h < nh; ++h) {
                    if (j0 + 16 * h >= B) continue;
                    for (int r = 0; r < R; ++r)
                        tma::load_2d(reinterpret_cast<uint8_t*>(tiles[stage]) + (size_t)(h * R + r) * a.sbox, &maps[c * R + r],
                                     (int)(2 * (j0 + 16 * h)), 0, &bars[stage]);
                }
            }
        }
#endif
    };

    if (warp < a.wpc) {
#ifndef QSFT_EMU
        if constexpr (TMA) {
            // U is written by other kernels / never by this one, the lists by generic stores: nothing to order for the TMA
            // reads of U beyond the kernel boundary.  First tile of this round:
            if (gw < n_tiles) issue(gw, (int)(tiles_done & 1u));
        }
#endif
        for (long long tt = gw; tt < n_tiles; tt += GW) {
            const int stage = (int)(tiles_done & 1u);
            const uint32_t parity = (tiles_done >> 1) & 1u;
            const int c = (int)(tt / tpg);
            const long long j0 = (tt - (long long)c * tpg) * bw;
            float2* tile = tiles[stage];
            if constexpr (TMA) {
#ifndef QSFT_EMU
                // the other stage was consumed (and partly rewritten in place) one iteration ago: order those generic-proxy
                // accesses before the bulk copy that overwrites it
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (tt + GW < n_tiles) issue(tt + GW, stage ^ 1);
                tma::mbar_wait(&bars[stage], parity);
#endif
            } else {
                // coalesced copy into the tile layout: lanes over bins (bw = 32) or over bins x two rows (bw = 16)
                const int lbf = lane & (bw - 1), sub = lane / bw, nsub = 32 / bw;
                const long long jf = j0 + lbf;
                TileCol tc{tile, R, sbox8, lbf};
                for (int r = 0; r < R; ++r) {
                    const float2* src = blk.p[c * R + r] + jf;
                    for (int i = sub; i < P_src; i += nsub)
                        tc.ref(r, i) = (jf < B) ? src[(size_t)i * a.ldU] : make_float2(0.f, 0.f);
                }
                __syncwarp();
            }
            ++tiles_done;

            // ---- phase 1: energy per bin, ball lists ------------------------------------------------------------------
            const int lb = lane & (bw - 1), sub = lane / bw, nsub = 32 / bw;
            const long long j = j0 + lb;
            const bool valid = (j < B);
            float e = 0.f;
            {
                TileCol tc{tile, R, sbox8, lb};
                for (int r = 0; r < R; ++r)
                    for (int i = sub; i < P_src; i += nsub) {
                        const float2 v = tc.ri(r, i);
                        e = fmaf(v.x, v.x, fmaf(v.y, v.y, e));
                    }
                if (nsub == 2) e += __shfl_xor_sync(0xffffffffu, e, 16);
            }
            int hd = 0;
            if (round > 1 && valid && sub == 0) hd = __ldcg(a.head + (size_t)c * B + j);
            unsigned touched = __ballot_sync(0xffffffffu, hd != 0);
            if (touched) {
                // subtract the balls peeled off these bins in earlier rounds (qsft.py:223-241), 4 bins at a time
                while (touched) {
                    int my = -1;
#pragma unroll
                    for (int g = 0; g < 32 / KL_G; ++g) {
                        const int bit = touched ? __ffs(touched) - 1 : -1;
                        if (touched) touched &= touched - 1;
                        if (g == grp) my = bit;
                    }
                    int f = __shfl_sync(0xffffffffu, hd, my >= 0 ? my : 0) - 1;
                    if (my < 0) f = -1;
                    TileCol tc{tile, R, sbox8, my >= 0 ? my : 0};
                    uint8_t* kb = s_sym + grp * (2 * QSFT_MAX_N);
                    while (__ballot_sync(0xffffffffu, f >= 0)) {
                        if (f >= 0) {
                            const uint4* src = reinterpret_cast<const uint4*>(a.find_k + (size_t)f * d.ld);
                            for (int w = gl; w < d.ld / 16; w += KL_G) reinterpret_cast<uint4*>(kb)[w] = __ldcg(src + w);
                        }
                        __syncwarp();
                        if (f >= 0) {
                            uint32_t kw[NW];
#pragma unroll
                            for (int w = 0; w < NW; ++w) kw[w] = (4 * w < d.ld) ? reinterpret_cast<const uint32_t*>(kb)[w] : 0u;
                            const float2 rho = __ldcg(a.find_rho + f);
                            const KlPhase<NW> ph{d, d.D + (size_t)c * d.P * d.ld, kb, kw, structured};
                            for (int r = 0; r < R; ++r) {
                                const int tb = ph.base(r);
                                for (int i = gl; i < P_src; i += KL_G) {
                                    const float2 w = s_tw[ph.row(r, i, tb)];
                                    float2& v = tc.ref(r, i);
                                    v.x -= rho.x * w.x - rho.y * w.y;
                                    v.y -= rho.x * w.y + rho.y * w.x;
                                }
                            }
                            f = __ldcg(a.next + (size_t)f * d.C + c) - 1;
                        }
                        __syncwarp();
                    }
                    // energy of the updated bin
                    float e2 = 0.f;
                    if (my >= 0)
                        for (int r = 0; r < R; ++r)
                            for (int i = gl; i < P_src; i += KL_G) {
                                const float2 v = tc.ri(r, i);
                                e2 = fmaf(v.x, v.x, fmaf(v.y, v.y, e2));
                            }
                    e2 = kl_group_sum(e2);
                    if (my >= 0 && gl == 0) s_e[my] = e2;
                }
                __syncwarp();
                if (hd != 0) e = s_e[lb];
            }

            // ---- phase 2: non-zeroton bins, four at a time ----------------------------------------------------------
            unsigned cand = __ballot_sync(0xffffffffu, valid && sub == 0 && e > thresh);
            while (cand) {
                int my = -1;
#pragma unroll
                for (int g = 0; g < 32 / KL_G; ++g) {
                    const int bit = cand ? __ffs(cand) - 1 : -1;
                    if (cand) cand &= cand - 1;
                    if (g == grp) my = bit;
                }
                const bool act = my >= 0;
                const int lbm = act ? my : 0;
                const long long jb = j0 + lbm;
                const float e_b = __shfl_sync(0xffffffffu, e, lbm);
                TileCol tc{tile, R, sbox8, lbm};
                uint8_t* sym = s_sym + grp * (2 * QSFT_MAX_N);
                uint8_t* kb = sym;
                if (act) {
                    for (int i = 1 + gl; i <= nsym; i += KL_G) sym[i - 1] = (uint8_t)detect_symbol(d, tc, i);
                    for (int i = nsym + gl; i < 4 * NW && i < QSFT_MAX_N; i += KL_G) sym[i] = 0;
                }
                __syncwarp();
                if (d.source == 1) {
                    kb = sym + QSFT_MAX_N;
                    if (act) {
                        for (int i = d.n + gl; i < 4 * NW && i < QSFT_MAX_N; i += KL_G) kb[i] = 0;
                        if (gl == 0) rs_decode(d, sym, kb);
                    }
                    __syncwarp();
                }
                uint32_t kw[NW];
#pragma unroll
                for (int w = 0; w < NW; ++w) kw[w] = reinterpret_cast<const uint32_t*>(kb)[w];
                const KlPhase<NW> ph{d, d.D + (size_t)c * d.P * d.ld, kb, kw, structured};
                // rho = <signature, col> / P (qsft.py:174-175)
                float rr = 0.f, ri = 0.f;
                if (act)
                    for (int r = 0; r < R; ++r) {
                        const int tb = ph.base(r);
                        for (int i = gl; i < P_src; i += KL_G) {
                            const float2 w = s_tw[ph.row(r, i, tb)];
                            const float2 v = tc.ri(r, i);
                            rr += w.x * v.x + w.y * v.y;                       // conj(sig) * v
                            ri += w.x * v.y - w.y * v.x;
                        }
                    }
                rr = kl_group_sum(rr) * (float)d.invP;
                ri = kl_group_sum(ri) * (float)d.invP;
                // residual ||col - rho sig||^2 (qsft.py:176,183), summed directly: every term is small for a singleton
                float res = 0.f;
                if (act)
                    for (int r = 0; r < R; ++r) {
                        const int tb = ph.base(r);
                        for (int i = gl; i < P_src; i += KL_G) {
                            const float2 w = s_tw[ph.row(r, i, tb)];
                            const float2 v = tc.ri(r, i);
                            const float dx = v.x - (rr * w.x - ri * w.y), dy = v.y - (rr * w.y + ri * w.x);
                            res = fmaf(dx, dx, fmaf(dy, dy, res));
                        }
                    }
                res = kl_group_sum(res);
                // bin hash j = dec(M_c^T k mod q) (qsft.py:178-179)
                long long part = 0;
                if (act) {
#pragma unroll
                    for (int u = 0; u < 32 / KL_G; ++u) {
                        const int i = gl + u * KL_G;
                        if (i < d.b)
                            part += wgt[u] * fast_mod(dot_raw<NW>(d.MT + ((size_t)c * d.b + i) * d.ld, d.ld, kw), d.q, d.qmagic);
                    }
                }
#pragma unroll
                for (int o = KL_G / 2; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                const float lim = fmaxf(thresh, a.rel_floor * e_b);
                const bool single = act && (part == jb) && !(res > lim);
                const bool lead = (gl == 0);
                const unsigned sb = __ballot_sync(0xffffffffu, lead && single);
                unsigned long long fbase = 0;
                if (lane == 0 && sb) fbase = atomicAdd(&a.counters[0], (unsigned long long)__popc(sb));
                fbase = __shfl_sync(0xffffffffu, fbase, 0);
                unsigned long long f = fbase + (unsigned long long)__popc(sb & ((1u << lane) - 1u));
                f = __shfl_sync(0xffffffffu, f, lane & ~(KL_G - 1));
                if (single) {
                    if ((long long)f < a.max_finds) {
                        uint32_t* ko = reinterpret_cast<uint32_t*>(a.find_k + (size_t)f * d.ld);
                        for (int w = gl; w < d.ld / 4; w += KL_G) ko[w] = (w < NW) ? kw[w] : 0u;
                        if (lead) {
                            a.find_cj[f] = (long long)c * B + jb;
                            a.find_rho[f] = make_float2(rr, ri);
                            a.find_round[f] = round;
                            a.find_id[(size_t)c * B + jb] = (int32_t)f;
                        }
                    }
                } else if (act && lead) {
                    ++n_multi;
                }
                __syncwarp();
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_multi += __shfl_xor_sync(0xffffffffu, n_multi, o);
    if (lane == 0 && n_multi) atomicAdd(&a.multi[round], (unsigned long long)n_multi);
}

// ---- link phase: one thread per find of the round ---------------------------------------------------------------------
template <int NW>
__device__ __forceinline__ void kl_link(const KlArgs& a, long long f0, long long f1, int round, bool do_link) {
    const PeelDev& d = a.d;
    const int nw = d.ld / 4;
    const long long B = d.B;
    for (long long f = f0 + (long long)blockIdx.x * KL_THREADS + threadIdx.x; f < f1; f += (long long)gridDim.x * KL_THREADS) {
        const long long cj = a.find_cj[f];
        const int c = (int)(cj / B);
        uint32_t kw[NW];
        const uint32_t* kin = reinterpret_cast<const uint32_t*>(a.find_k + (size_t)f * d.ld);
#pragma unroll
        for (int w = 0; w < NW; ++w) kw[w] = (w < nw) ? kin[w] : 0u;
        float2 sum = a.find_rho[f];
        int cnt = 1;
        bool first = true, last = true;
        long long jl[KL_MAX_BLOCKS];                       // C <= 16 on this path
        for (int c2 = 0; c2 < d.C; ++c2) {
            if (c2 == c) {
                jl[c2] = cj - (long long)c * B;
                continue;
            }
            const long long j2 = hash_bin<NW>(d, c2, kw);
            jl[c2] = j2;
            const int32_t f2 = a.find_id[(size_t)c2 * B + j2];
            // find_id is only written for singletons: an entry is a find of THIS round iff it lies in the round's range and
            // that find really sits in bin (c2, j2)
            if ((long long)f2 >= f0 && (long long)f2 < f1 && a.find_cj[f2] == (long long)c2 * B + j2) {
                const uint32_t* k2 = reinterpret_cast<const uint32_t*>(a.find_k + (size_t)f2 * d.ld);
                bool same = true;
#pragma unroll
                for (int w = 0; w < NW; ++w) same &= ((w < nw) ? k2[w] : 0u) == kw[w];
                if (same) {
                    if (c2 < c) {
                        first = false;                      // an earlier group holds the round's first find of this k
                    } else {
                        last = false;                       // ball_values: a later (i, j) wins (qsft.py:215)
                        const float2 r2 = a.find_rho[f2];
                        sum.x += r2.x;
                        sum.y += r2.y;
                        ++cnt;
                    }
                }
            }
        }
        if (first && a.has_uniq)
            k4_uniq_commit<NW>(d, kw, sum, cnt, cj, jl[0], round, a.uo.seen0, a.uo.uk, a.uo.usum, a.uo.ucnt, a.uo.ukey, a.uo.unext,
                               a.uo.max_uniq, a.counters);
        if (last && do_link) {
            // peel: the ball joins the list of every bin it hashes to (qsft.py:223-241)
            for (int l = 0; l < d.C; ++l) {
                const int32_t prev = atomicExch(a.head + (size_t)l * B + jl[l], (int32_t)(f + 1));
                a.next[(size_t)f * d.C + l] = prev;
            }
            atomicAdd(&a.counters[2], 1ull);                // num_peeling (qsft.py:224)
        }
    }
}

template <int NW, bool TMA>
__global__ void __launch_bounds__(KL_THREADS, 1)
k4_peel_loop_kernel(const KlArgs a, const KlBlocks blk
#ifndef QSFT_EMU
                    , const __grid_constant__ KlMaps maps
#endif
) {
    extern __shared__ __align__(1024) uint8_t kl_smem[];
    __shared__ float2 s_tw[QSFT_MAX_Q + 1];
    __shared__ __align__(8) uint64_t s_bars[2 * (KL_THREADS / 32)];
    const PeelDev& d = a.d;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < d.q) {
        float sn, cs;
        sincospif(2.0f * (float)threadIdx.x / (float)d.q, &sn, &cs);
        if (d.q == 4) {                                     // exact quarter turns
            cs = (threadIdx.x == 0) ? 1.f : (threadIdx.x == 2) ? -1.f : 0.f;
            sn = (threadIdx.x == 1) ? 1.f : (threadIdx.x == 3) ? -1.f : 0.f;
        } else if (d.q == 2) {
            cs = threadIdx.x == 0 ? 1.f : -1.f;
            sn = 0.f;
        }
        s_tw[threadIdx.x] = make_float2(cs, sn);
    }
#ifndef QSFT_EMU
    uint8_t* base = kl_smem + ((1024u - (tma::smem_u32(kl_smem) & 1023u)) & 1023u);
    if (TMA && threadIdx.x == 0) {
        for (int i = 0; i < 2 * (KL_THREADS / 32); ++i) tma::mbar_init(&s_bars[i], 1);
        tma::mbar_fence_init();
    }
#else
    uint8_t* base = kl_smem;
#endif
    __syncthreads();
    const int per_warp = 2 * (a.bw >> 4) * d.R * a.sbox + 4 * 2 * QSFT_MAX_N + 32 * 4;
    uint8_t* wsm = base + (size_t)warp * ((per_warp + 1023) & ~1023);
    unsigned int epoch = 0, tiles_done = 0;
    long long total = 0;
    double num_peeling = 0;
    int round = 0;
    bool cont = true;
    while (cont && num_peeling < a.peeling_max && round < a.max_rounds) {
        ++round;
        kl_classify<NW, TMA>(a, blk,
#ifndef QSFT_EMU
                             maps.m,
#endif
                             round, wsm, &s_bars[2 * warp], tiles_done, s_tw);
        kl_grid_barrier(a.gbar, epoch);
        const long long now = (long long)__ldcg(a.counters + 0);
        const long long multis = (long long)__ldcg(a.multi + round);
        if (now > a.max_finds) {                            // find buffer too small: report, stop (uniform over the grid)
            if (blockIdx.x == 0 && threadIdx.x == 0) a.counters[6] = 1ull;
            total = a.max_finds;
            break;
        }
        const long long nf = now - total;
        if (multis == 0 || nf == 0) cont = false;           // qsft.py:204-205
        // the reference also subtracts after its last round, but nothing reads the bins afterwards: skip unless the q^n
        // guard needs the count
        const bool do_link = cont || a.guard_can_bind;
        if (nf > 0) kl_link<NW>(a, total, now, round, do_link);
        total = now;
        if (cont || a.guard_can_bind) {
            kl_grid_barrier(a.gbar, epoch);
            if (a.guard_can_bind) num_peeling = (double)__ldcg(a.counters + 2);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.counters[5] = (unsigned long long)round;
        a.counters[7] = (unsigned long long)total;
    }
}

// D[c][r * P_src + i] == (D[c][r * P_src] - e_{i-1}) mod q for all c, r, i >= 1 (and P_src == n + 1)?
__global__ void kl_dstruct_kernel(PeelDev d, int* flag) {
    __shared__ int bad;
    if (threadIdx.x == 0) bad = (d.P_src != d.n + 1) ? 1 : 0;
    __syncthreads();
    const long long total = (long long)d.C * d.P * d.n;
    for (long long e = threadIdx.x; e < total && !bad; e += blockDim.x) {
        const int u = (int)(e % d.n);
        const long long cp = e / d.n;
        const int p = (int)(cp % d.P), c = (int)(cp / d.P);
        const int r = p / d.P_src, i = p - r * d.P_src;
        if (i == 0) continue;
        const int d0 = d.D[((size_t)c * d.P + r * d.P_src) * d.ld + u];
        int want = d0 - (u == i - 1 ? 1 : 0);
        if (want < 0) want += d.q;
        if ((int)d.D[((size_t)c * d.P + p) * d.ld + u] != want) bad = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) *flag = bad ? 0 : 1;
}

}  // namespace (device part; the CPU emulation cuts here)

// ---- host side ---------------------------------------------------------------------------------------------------------
namespace {

template <int NW>
int kl_launch(const KlArgs& a, const KlBlocks& blk, const KlMaps& maps, bool use_tma, size_t smem, int grid, cudaStream_t st) {
    void* params[] = {(void*)&a, (void*)&blk, (void*)&maps};
    const void* fn = use_tma ? (const void*)k4_peel_loop_kernel<NW, true> : (const void*)k4_peel_loop_kernel<NW, false>;
    QSFT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QSFT_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(KL_THREADS), params, smem, st));
    g_qsft_launches.fetch_add(1, std::memory_order_relaxed);
    return QSFT_OK;
}

}  // namespace

// Whole peel loop on the device.  blocks[c * R + r] -> (P_src, ldU) complex64 rows of group c, repeat r (device pointers,
// host array).  Outputs as qsft_peel.  Returns QSFT_EUNSUPPORTED when the shape does not fit this kernel (C * R > 16,
// P_src > 256, tile larger than the shared memory).
int qsft_peel_loop(const PeelDev& d, const float* const* blocks, int64_t ldU, int64_t* find_cj, int8_t* find_k, float* find_rho,
                   int32_t* find_round, int32_t* find_id, int64_t max_finds, unsigned long long* counters, const UniqOut* uo,
                   int64_t* n_finds_out, int64_t* n_uniq_out, int* n_rounds_out, cudaStream_t st) {
    const int nblk = d.C * d.R;
    if (nblk > KL_MAX_BLOCKS || d.P_src > 256 || d.C > KL_MAX_BLOCKS) return QSFT_EUNSUPPORTED;
    static int sms = 0, smem_max = 0;
    if (!sms) {
        int dev = 0;
        QSFT_CUDA(cudaGetDevice(&dev));
        int coop = 0;
        QSFT_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        QSFT_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        sms = coop ? qsft_num_sms() : -1;
    }
    if (sms < 0) return QSFT_EUNSUPPORTED;
    KlArgs a{};
    a.d = d;
    a.ldU = ldU;
    a.sbox = (d.P_src * 128 + 1023) & ~1023;
    // tile width: 32 bins unless that leaves fewer than 4 warps per SM
    const int fixed = 4 * 2 * QSFT_MAX_N + 32 * 4;
    const int budget = smem_max - 2048 - 1024;
    a.bw = 32;
    auto per_warp = [&](int bw) { return ((2 * (bw >> 4) * d.R * a.sbox + fixed) + 1023) & ~1023; };
    if (budget / per_warp(32) < 4) a.bw = 16;
    a.wpc = budget / per_warp(a.bw);
    if (a.wpc > KL_THREADS / 32) a.wpc = KL_THREADS / 32;
    if (a.wpc < 1) return QSFT_EUNSUPPORTED;
    const size_t smem = (size_t)a.wpc * per_warp(a.bw) + 1024;
    const bool use_tma = (ldU % 2 == 0) && (getenv("QSFT_K4_NO_TMA") == nullptr);
    KlBlocks blk{};
    for (int i = 0; i < nblk; ++i) {
        blk.p[i] = reinterpret_cast<const float2*>(blocks[i]);
        if (use_tma && ((uintptr_t)blocks[i] & 15)) return QSFT_EUNSUPPORTED;
    }
    // workspace: ball lists, grid barrier, delay-structure flag
    const size_t head_b = (size_t)d.C * d.B * 4, next_b = (size_t)max_finds * d.C * 4;
    uint8_t* ws = nullptr;
    const size_t head_off = 256, next_off = head_off + ((head_b + 255) & ~(size_t)255);
    QSFT_CUDA(qsft_scratch_alloc((void**)&ws, next_off + next_b, st));
    QSFT_CUDA(cudaMemsetAsync(ws, 0, head_off + head_b, st));                  // barrier, flag, list heads
    a.gbar = reinterpret_cast<unsigned int*>(ws);
    int* dflag = reinterpret_cast<int*>(ws + 64);
    a.multi = reinterpret_cast<unsigned long long*>(ws + 128);                 // 16 slots
    a.dstruct = dflag;
    a.head = reinterpret_cast<int32_t*>(ws + head_off);
    a.next = reinterpret_cast<int32_t*>(ws + next_off);
    KlMaps hm;
    memset(&hm, 0, sizeof(hm));
    if (use_tma) {
        for (int i = 0; i < nblk; ++i) {
            const cuuint64_t dims[2] = {(cuuint64_t)(2 * d.B), (cuuint64_t)d.P_src};
            const cuuint64_t strides[1] = {(cuuint64_t)ldU * 8};
            const cuuint32_t box[2] = {32, (cuuint32_t)d.P_src};
            if (int rc = tma::make_map(&hm.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, blocks[i], dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) {
                cudaFreeAsync(ws, st);
                return rc;
            }
        }
    }
    a.find_cj = (long long*)find_cj;
    a.find_k = find_k;
    a.find_rho = reinterpret_cast<float2*>(find_rho);
    a.find_round = find_round;
    a.find_id = find_id;
    a.max_finds = max_finds;
    if (uo) a.uo = *uo;
    a.has_uniq = uo ? 1 : 0;
    a.counters = counters;
    a.max_rounds = 15;
    if (const char* mr = getenv("QSFT_K4_MAX_ROUNDS"))       // measurement aid (tools/microbench.py): cost of the first rounds alone
        if (atoi(mr) >= 1 && atoi(mr) < 15) a.max_rounds = atoi(mr);
    a.peeling_max = pow((double)d.q, (double)d.n);
    a.guard_can_bind = a.peeling_max <= 15.0 * (double)d.C * (double)d.B ? 1 : 0;
    a.rel_floor = 1e-10f;
    QSFT_CUDA(cudaMemsetAsync(counters, 0, 8 * sizeof(unsigned long long), st));
    if (uo) QSFT_CUDA(cudaMemsetAsync(uo->seen0, 0, (size_t)d.B * sizeof(int32_t), st));
    kl_dstruct_kernel<<<1, 256, 0, st>>>(d, dflag);
    QSFT_LAUNCHED();
    int rc;
    const int nw = d.ld / 4;
    if (nw <= 4) rc = kl_launch<4>(a, blk, hm, use_tma, smem, sms, st);
    else if (nw <= 8) rc = kl_launch<8>(a, blk, hm, use_tma, smem, sms, st);
    else if (nw <= 16) rc = kl_launch<16>(a, blk, hm, use_tma, smem, sms, st);
    else rc = kl_launch<32>(a, blk, hm, use_tma, smem, sms, st);
    if (rc != QSFT_OK) {
        cudaFreeAsync(ws, st);
        return rc;
    }
    static thread_local unsigned long long* host = nullptr;
    if (!host) QSFT_CUDA(cudaMallocHost(&host, 8 * sizeof(unsigned long long)));
    QSFT_CUDA(cudaMemcpyAsync(host, counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    QSFT_CUDA(cudaFreeAsync(ws, st));
    QSFT_CUDA(cudaStreamSynchronize(st));                   // the only synchronisation of the peel: result sizes
    if (host[6]) {
        qsft_set_error("find buffer too small: %llu finds > max_finds=%lld", host[0], (long long)max_finds);
        return QSFT_EINVAL;
    }
    if (uo && (long long)host[4] > uo->max_uniq) {
        qsft_set_error("unique buffer too small: %llu > %lld", host[4], uo->max_uniq);
        return QSFT_EINVAL;
    }
    *n_finds_out = (int64_t)host[7];
    if (n_uniq_out) *n_uniq_out = (int64_t)host[4];
    *n_rounds_out = (int)host[5];
    return QSFT_OK;
}
