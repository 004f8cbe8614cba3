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
// Classification of one round: the CTA walks over tiles of W bins (W = 128 .. 16) x all P delay rows of one group:
//   * TMA variant (q^b a multiple of 16): 1 producer warp + 16 consumer warps, ring of 2 .. 6 stages.  A tile is ONE
//     cp.async.bulk.tensor box {16 bins = 128 B, W / 16 chunks, P_src rows} per repeat block (every delay row contributes
//     W * 8 contiguous bytes) and lands as [row][chunk][128 B] with the 128-byte swizzle.  The bins' ball-list heads travel
//     with the tile (cp.async.bulk).
//   * plain variant (odd q^b, tiny q^b; also what the CPU emulation of tests/emu runs): the 16 consumer warps copy the tile
//     with coalesced loads into the same layout, single stage.
//   Step 0 (rounds > 1): bins with listed balls are updated in place by 8-lane groups.  Step 1: energies, 512 threads over
//   W bins x row slices.  One CTA barrier.  Step 2: every warp rebuilds the tile's candidate mask and takes its share of the
//   non-zeroton bins, four at a time by 8-lane groups: symbols (reconstruct.py:12-31,100-129), optional Reed-Solomon decode,
//   rho and residual in ONE pass over the rows (qsft.py:174-183), bin hash check (qsft.py:178-179).
// Link phase of a round: one thread per find: duplicate gathering / averaging exactly like k4_reduce_kernel, and the
// "last (i, j) wins" find of every k (qsft.py:215) links the ball into its C bins.
//
// Rounds after the first do NOT scan U again.  A bin whose ball list did not change in the previous link phase holds what it
// held when it was last classified -- a zeroton or a multiton (a singleton's own ball is always linked into it) -- and would
// be classified the same way, so only the bins the link phase touched ("dirty" bins, listed as they are touched) are looked
// at, and the multiton count of the stop rule (qsft.py:204-205) is carried from round to round (per-bin class byte: the
// count changes by the dirty bins that stop / start being multitons).  Most dirty bins are the singletons themselves: the bin
// where k was found with value rho and residual res, after the ball (k, rho_last) has been subtracted, has the energy
// res + P |rho - rho_last|^2 EXACTLY (the residual of a least-squares fit is orthogonal to the signature), which the link
// phase knows without touching the bin (`zres`).  Only dirty bins that received a ball of a k they were not themselves found
// with (multitons of the previous round, or several balls at once) are gathered from U (column reads into a warp's private
// buffer), get all their listed balls subtracted and are classified like a bin of the first round.
#pragma once
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

#include "k4_shared.cuh"
#ifndef QSFT_EMU
#include "tma.cuh"
#endif

constexpr int KL_MAX_BLOCKS = 16;            // (c, r) blocks of U addressed separately (C * R <= 16)

struct KlBlocks {
    const float2* p[KL_MAX_BLOCKS];          // block c * R + r: (P_src, ldU) complex64, bin index contiguous
};
#ifndef QSFT_EMU
struct KlMaps {
    CUtensorMap m[KL_MAX_BLOCKS];            // the same blocks as 3-D tensors {32 floats, B / 16 chunks, P_src rows}
};
#endif

// Bin-sharded loop (one process per GPU, every rank holds all of U): each rank classifies the bins [jb, je) of every group
// and appends its finds to ITS segment of the find list; the list, the (C, B) find table and a few control words live in a
// symmetric (peer-mapped) workspace of identical layout on every rank.  After a round's classification every rank PUSHES its
// new slots into the same places of every peer's workspace with plain stores over NVLink, publishes its counters and a flag
// (release, system scope), waits for the peers' flags, and runs the link phase on the round's finds of ALL ranks -- so the
// ball lists, the distinct-k list and the stop rule evolve identically everywhere.  One exchange per round, inside the
// kernel: the all-gather of (k, rho) lists of the reference's round structure (qsft.py:209-241), fused into the peel.
struct KlShard {
    uint8_t* peer[8];                        // base of every rank's workspace (own one included)
    long long off_cj, off_k, off_rho, off_round, off_id, off_res, off_ctl;      // byte offsets of the arrays in a workspace
    unsigned int epoch;                      // flags hold (epoch << 8) | round: no clearing between peels
};
// control block of a workspace: [round 0 .. 15][source rank 0 .. 7]
struct KlCtl {
    unsigned long long now[16][8];           // slots used in the source rank's segment after the round
    unsigned long long multi[16][8];         // multitons the source rank saw in the round
    unsigned int flag[16][8];
};

struct KlArgs {
    PeelDev d;
    long long ldU;                           // row stride of every block (elements)
    long long* find_cj;
    int8_t* find_k;
    float2* find_rho;
    int32_t* find_round;
    int32_t* find_id;                        // (C, B): written for singletons only; validated through find_cj
    float* find_res;                         // residual of the singleton test of every find (see `zres`)
    long long max_finds;
    int32_t* head;                           // (C, B): last ball linked into the bin + 1 (0 = none); zeroed by the host
    int32_t* next;                           // (max_finds, C): previous ball of the same bin + 1
    unsigned int* zres;                      // (C, B) float bits, 0 = untouched since the bin was last looked at: energy left in the
                                             // bin once the round's ball is subtracted, when that is known without reading the
                                             // bin (the bin's own find carries the same k), +inf when it is not; atomic max
    uint8_t* cls;                            // (C, B): 2 = the bin was a multiton when it was last classified
    long long* dirty;                        // (max_dirty) c * B + j of the bins (of this rank's range) touched by the last link phase
    long long max_dirty;
    unsigned long long* dcount;              // [2]: entries of `dirty` written by the link phase of round r -> dcount[r & 1]
    UniqOut uo;
    int has_uniq;
    unsigned long long* counters;            // [0] finds, [2] balls peeled, [4] distinct k, [5] rounds, [6] error flags,
                                             // [7] finds kept
    unsigned long long* multi;               // [r] multitons of round r (1 <= r <= 15; workspace, zeroed by the host)
    unsigned int* gbar;                      // grid barrier counter (zeroed by the host)
    const int* dstruct;                      // device flag: D[c][r][i] = D[c][r][0] - e_{i-1} (identity / nso delays)
    int W, lgW;                              // bins per tile (power of two, 16 .. 128)
    int box;                                 // bytes per repeat block of a tile: (W / 16) * P_src * 128 rounded up to 1024
    int stage_bytes;                         // R * box + list heads, rounded up to 1024
    int nstages;
    int priv_bytes;                          // per candidate warp: private copy of its four bins' columns
    int chunk;                               // find slots a warp reserves at a time
    int max_rounds;
    int guard_can_bind;
    double peeling_max;
    float rel_floor;                         // residual floor relative to the bin energy (fp32 resolution of U)
    // ---- bin-sharded loop over several GPUs (world > 1; see KlShard below) ----
    int rank, world;
    long long jb, je;                        // this rank classifies bins [jb, je) of every group (jb a multiple of 128)
    long long seg;                           // find slots per rank: rank r owns slots [r * seg, (r + 1) * seg)
    KlShard sh;
};

namespace {

constexpr int KL_NS = 4;                     // scanner warps: two teams of two (64 threads x two bins = a tile of 128 bins)
constexpr int KL_NC = 11;                    // candidate warps (16 warps with the producer: 128 registers per thread).  Measured slower: 24 warps at 80 registers (scan 0.38 instead of 0.27 ms, whole peel 2.0 instead of 1.7 ms); one scanner team + 13 candidate warps (scan 0.47 instead of 0.31 ms, whole peel 1.12 instead of 1.05 ms)
constexpr int KL_CT = (KL_NS + KL_NC) * 32;  // threads without the TMA producer warp
constexpr int KL_G = 8;                      // lanes per bin in the group phases
constexpr int KL_MAXW = 128;                 // bins per tile, at most
constexpr int KL_MAX_STAGES = 6;
constexpr int KL_SYM = 2 * QSFT_MAX_N;       // per group: detected symbols + decoded k
constexpr int KL_QN = 256;                   // entries of the CTA's work queue (at most 6 stages x 32 item groups + KL_NC sentinels pending)
constexpr int KL_CTRL_BYTES = 256 + KL_MAX_STAGES * 576 + KL_QN * 4 + 128 + KL_NC * 4 * KL_SYM;

// ---- tile access -------------------------------------------------------------------------------------------------
// stage = [repeat r][row i][chunk ch (16 bins)][128 B: 16 bins, 16-byte chunks xor-swizzled with the line index & 7]: every
// delay row contributes W * 8 CONTIGUOUS bytes, and the bulk copy walks them in that order (128-byte pieces of a row one
// after the other, then the next row): long DRAM bursts.  (With the rows as the faster box dimension the copy engine jumps a
// whole row of U -- megabytes -- between consecutive 128-byte requests and DRAM delivers about half its bandwidth.)
// Lanes over bins of one row (scan) are conflict free; lanes over rows of one bin (the one-time copy of a candidate bin into
// registers) share banks, which is the cheaper side to pay on.
struct TileCol {
    uint8_t* t;
    int P_src, box;
    int lb;                                  // local bin
    int lgc;                                 // log2(chunks per row) = lgW - 4
    // With eight or more chunks per row (W = 128) the swizzle term does not depend on the row: line & 7 = (lb >> 4) & 7, and
    // the offset is r * box + i * (W * 8) + a per-bin constant -- one multiply-add per element instead of eight instructions
    // (the scan's address arithmetic was 10 % of all instructions of the kernel).
    __device__ __forceinline__ int off(int r, int i) const {
        if (lgc >= 3) {
            const int c0 = lb >> 4;
            const int k = c0 * 128 + (((((lb & 15) >> 1) ^ (c0 & 7)) << 4) | ((lb & 1) << 3));
            return r * box + (i << (lgc + 7)) + k;
        }
        const int line = (i << lgc) + (lb >> 4);
        return r * box + line * 128 + (((((lb & 15) >> 1) ^ (line & 7)) << 4) | ((lb & 1) << 3));
    }
    __device__ __forceinline__ float2 ri(int r, int i) const { return *reinterpret_cast<const float2*>(t + off(r, i)); }
    __device__ __forceinline__ float2& ref(int r, int i) const { return *reinterpret_cast<float2*>(t + off(r, i)); }
};

// A candidate warp's private copy of the columns of its four work items: [row r * P_src + i][item 0 .. 3] complex64.  The
// eight lanes of an item walk rows i = gl, gl + 8, ...: a warp access covers 256 contiguous bytes (conflict free), and the
// address is a shift and an add instead of the swizzle arithmetic of the stage.
struct PrivCol {
    float2* p;                               // + item
    int P_src;
    __device__ __forceinline__ float2 ri(int r, int i) const { return p[(r * P_src + i) << 2]; }
    __device__ __forceinline__ float2& ref(int r, int i) const { return p[(r * P_src + i) << 2]; }
};

__device__ __forceinline__ float kl_group_sum(float v) {
#pragma unroll
    for (int o = KL_G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// position of the n-th (0-based) set bit of m; n < popc(m)
__device__ __forceinline__ int kl_nth_bit(unsigned m, int n) {
    int pos = 0;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const unsigned low = m & ((1u << s) - 1u);
        const int c = __popc(low);
        if (n >= c) {
            n -= c;
            m >>= s;
            pos += s;
        } else {
            m = low;
        }
    }
    return pos;
}

// local bin of rank `rank` in the tile's masks (4 words of 32 bins), -1 when rank >= total
__device__ __forceinline__ int kl_pick(const unsigned (&mask)[4], int rank) {
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const int c = __popc(mask[w]);
        if (rank >= 0 && rank < c) return w * 32 + kl_nth_bit(mask[w], rank);
        rank -= c;
    }
    return -1;
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

// Find slots are handed out to the warps in chunks of KL_CHUNK (one global atomic per chunk: a single counter bumped by every
// warp of every SM for every group of candidates serialises in L2).  Slots of a chunk that stay unused at the end of a round
// are marked with find_cj = -1; the link phase and every reader of the find list skip them.
constexpr int KL_CHUNK = 64;                 // at most; KlArgs.chunk = 4 .. 64 by problem size (kl_chunk)
struct KlSlots {
    long long next, end;                     // this warp's chunk: slots [next, end) are free
    long long lim;                           // end of this rank's segment (slots beyond it are counted, not stored)
};

// reserves `cnt` (<= 4, warp-uniform) consecutive slots; all lanes get the first one
__device__ __forceinline__ long long kl_take(const KlArgs& a, KlSlots& sl, int cnt) {
    const int lane = threadIdx.x & 31;
    if (sl.next + cnt > sl.end) {
        for (long long f = sl.next + lane; f < sl.end; f += 32)
            if (f < sl.lim) a.find_cj[f] = -1;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&a.counters[0], (unsigned long long)a.chunk);      // cursor inside this rank's segment
        base = __shfl_sync(0xffffffffu, base, 0);
        sl.next = a.seg * a.rank + (long long)base;
        sl.end = sl.next + a.chunk;
        sl.lim = a.seg * (a.rank + 1);
    }
    const long long f = sl.next;
    sl.next += cnt;
    return f;
}

// end of a round: the rest of the warp's chunk is given up
__device__ __forceinline__ void kl_close(const KlArgs& a, KlSlots& sl) {
    const int lane = threadIdx.x & 31;
    for (long long f = sl.next + lane; f < sl.end; f += 32)
        if (f < sl.lim) a.find_cj[f] = -1;
    sl.next = sl.end = 0;
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

// Per-stage hand-over from the scanner warps to the candidate warps.
struct KlTileInfo {
    unsigned mask[4];                        // work items of the tile: bins that are not zerotons or carry peeled balls
    unsigned tmask[4];                       // ... of which: bins with peeled balls (their energy is not known yet)
    float e[KL_MAXW];                        // bin energies (bins without peeled balls)
    long long j0;                            // first bin of the tile
    int c;                                   // group
    int left;                                // item groups (of four bins) not yet copied out of the stage
};

// ---- scanner warps: energies and the tile's work list ----------------------------------------------------------------------
// Two teams of two warps take alternate tiles; thread t (0 .. 63) of a team owns bins t and t + 64 of the tile.
__device__ __forceinline__ void kl_scan(const KlArgs& a, uint8_t* stage, long long j0, int round, KlTileInfo* info, int t) {
    const PeelDev& d = a.d;
    const int R = d.R, P_src = d.P_src;
    const int32_t* s_head = reinterpret_cast<const int32_t*>(stage + (size_t)R * a.box);
    const bool valid0 = t < a.W && j0 + t < a.je, valid1 = t + 64 < a.W && j0 + t + 64 < a.je;
    int hd0 = 0, hd1 = 0;
    if (round > 1) {
        if (valid0) hd0 = s_head[t];
        if (valid1) hd1 = s_head[t + 64];
    }
    // rows in blocks of eight, the two bins side by side: sixteen independent shared-memory loads are in flight before the
    // first FMA needs one (a row-at-a-time loop spends its time waiting for each load: the scan, not DRAM, then sets the pace)
    float e0 = 0.f, e1 = 0.f;
    {
        const bool on0 = valid0 && hd0 == 0, on1 = valid1 && hd1 == 0;
        TileCol tc0{stage, P_src, a.box, on0 ? t : 0, a.lgW - 4};
        TileCol tc1{stage, P_src, a.box, on1 ? t + 64 : 0, a.lgW - 4};
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (on0 || on1) {
            for (int r = 0; r < R; ++r) {
                int i = 0;
                for (; i + 8 <= P_src; i += 8) {
                    float2 v0[8], v1[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        v0[u] = tc0.ri(r, i + u);
                        v1[u] = tc1.ri(r, i + u);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        acc[u & 1] = fmaf(v0[u].x, v0[u].x, fmaf(v0[u].y, v0[u].y, acc[u & 1]));
                        acc[2 + (u & 1)] = fmaf(v1[u].x, v1[u].x, fmaf(v1[u].y, v1[u].y, acc[2 + (u & 1)]));
                    }
                }
                for (; i < P_src; ++i) {
                    const float2 v0 = tc0.ri(r, i), v1 = tc1.ri(r, i);
                    acc[0] = fmaf(v0.x, v0.x, fmaf(v0.y, v0.y, acc[0]));
                    acc[2] = fmaf(v1.x, v1.x, fmaf(v1.y, v1.y, acc[2]));
                }
            }
        }
        if (on0) e0 = acc[0] + acc[1];
        if (on1) e1 = acc[2] + acc[3];
    }
    info->e[t] = e0;
    info->e[t + 64] = e1;
    const float thresh = (float)d.thresh;
    const unsigned m0 = __ballot_sync(0xffffffffu, valid0 && (hd0 != 0 || e0 > thresh));
    const unsigned m1 = __ballot_sync(0xffffffffu, valid1 && (hd1 != 0 || e1 > thresh));
    const unsigned t0 = __ballot_sync(0xffffffffu, hd0 != 0);
    const unsigned t1 = __ballot_sync(0xffffffffu, hd1 != 0);
    if ((t & 31) == 0) {
        const int w = t >> 5;                                   // bins 32 w .. and 64 + 32 w ..
        info->mask[w] = m0;
        info->mask[w + 2] = m1;
        info->tmask[w] = t0;
        info->tmask[w + 2] = t1;
    }
}

// ---- candidate warps (cw = 0 .. KL_NC - 1): this warp's share of the tile's work items, four at a time ------------------
// kl_cand_body works on the columns of the warp's four items through `col` (the warp's private copy, PrivCol); `f` = the bin's
// first listed ball (-1: none).  Returns the bin's class to all lanes of its group: 0 zeroton, 1 singleton (find recorded),
// 2 multiton.
template <int NW, class Col>
__device__ __forceinline__ int kl_cand_body(const KlArgs& a, const Col& col, int c, long long jb, int round, bool act, bool touched,
                                            int f, float e_b, uint8_t* sym, const float2* s_tw, bool structured,
                                            const long long (&wgt)[32 / KL_G], KlSlots& slots) {
    const PeelDev& d = a.d;
    const int lane = threadIdx.x & 31;
    const int gl = lane % KL_G;
    const int R = d.R, P_src = d.P_src;
    const long long B = d.B;
    const float thresh = (float)d.thresh;
    const int nsym = P_src - 1;
    // bins with peeled balls (qsft.py:223-241 applied to the copy in shared memory), then their energy
    if (round > 1 && __ballot_sync(0xffffffffu, touched)) {
        while (__ballot_sync(0xffffffffu, f >= 0)) {
            // the ball's k, rho and list link are independent loads (L2): all in flight before the first is needed
            float2 rho = make_float2(0.f, 0.f);
            int fn = -1;
            if (f >= 0) {
                const uint4* src = reinterpret_cast<const uint4*>(a.find_k + (size_t)f * d.ld);
                for (int w = gl; w < d.ld / 16; w += KL_G) reinterpret_cast<uint4*>(sym)[w] = __ldcg(src + w);
                rho = __ldcg(a.find_rho + f);
                fn = __ldcg(a.next + (size_t)f * d.C + c) - 1;
            }
            __syncwarp();
            if (f >= 0) {
                uint32_t kw[NW];
#pragma unroll
                for (int w = 0; w < NW; ++w) kw[w] = (4 * w < d.ld) ? reinterpret_cast<const uint32_t*>(sym)[w] : 0u;
                const KlPhase<NW> ph{d, d.D + (size_t)c * d.P * d.ld, sym, kw, structured};
                for (int r = 0; r < R; ++r) {
                    const int tb = ph.base(r);
                    for (int i = gl; i < P_src; i += KL_G) {
                        const float2 w = s_tw[ph.row(r, i, tb)];
                        float2& v = col.ref(r, i);
                        v.x -= rho.x * w.x - rho.y * w.y;
                        v.y -= rho.x * w.y + rho.y * w.x;
                    }
                }
            }
            f = fn;
            __syncwarp();
        }
        float e2 = 0.f;
        if (touched)
            for (int r = 0; r < R; ++r)
                for (int i = gl; i < P_src; i += KL_G) {
                    const float2 v = col.ri(r, i);
                    e2 = fmaf(v.x, v.x, fmaf(v.y, v.y, e2));
                }
        e2 = kl_group_sum(e2);
        if (touched) {
            e_b = e2;
            act = e2 > thresh;                                  // energy test (qsft.py:164)
        }
    }
    uint8_t* kb = sym;
    if (act) {
        for (int i = 1 + gl; i <= nsym; i += KL_G) sym[i - 1] = (uint8_t)detect_symbol(d, col, i);
        for (int i = nsym + gl; i < 4 * NW && i < QSFT_MAX_N; i += KL_G) sym[i] = 0;
    }
    __syncwarp();
    if (d.source == 1) {
        kb = sym + QSFT_MAX_N;
        if (act) {
            for (int i = d.n + gl; i < 4 * NW && i < QSFT_MAX_N; i += KL_G) kb[i] = 0;
            if (gl == 0) rs_decode(rs_params(d), sym, kb);
        }
        __syncwarp();
    }
    uint32_t kw[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) kw[w] = reinterpret_cast<const uint32_t*>(kb)[w];
    const KlPhase<NW> ph{d, d.D + (size_t)c * d.P * d.ld, kb, kw, structured};
    // rho = <signature, col> / P (qsft.py:174-175) and the residual ||col - rho sig||^2 (qsft.py:176,183) in one pass:
    // with z_i = conj(sig_i) col_i and the shift z0 = z of row (0, 0),  rho = z0 + mean(z_i - z0)  and
    // residual = sum |z_i - z0|^2 - |sum (z_i - z0)|^2 / P.  For a singleton every z_i - z0 is at rounding level, so the
    // subtraction cancels nothing that matters; for a multiton the residual is large either way.
    float sx = 0.f, sy = 0.f, s2 = 0.f;
    float2 z0 = make_float2(0.f, 0.f);
    if (act) {
        const int tb0 = ph.base(0);
        {
            const float2 w = s_tw[tb0];
            const float2 v = col.ri(0, 0);
            z0 = make_float2(w.x * v.x + w.y * v.y, w.x * v.y - w.y * v.x);
        }
        for (int r = 0; r < R; ++r) {
            const int tb = r == 0 ? tb0 : ph.base(r);
            for (int i = gl; i < P_src; i += KL_G) {
                const float2 w = s_tw[ph.row(r, i, tb)];
                const float2 v = col.ri(r, i);
                const float dx = (w.x * v.x + w.y * v.y) - z0.x;            // conj(sig) * v - z0
                const float dy = (w.x * v.y - w.y * v.x) - z0.y;
                sx += dx;
                sy += dy;
                s2 = fmaf(dx, dx, fmaf(dy, dy, s2));
            }
        }
    }
    sx = kl_group_sum(sx);
    sy = kl_group_sum(sy);
    s2 = kl_group_sum(s2);
    const float invP = (float)d.invP;
    const float rr = z0.x + sx * invP, ri = z0.y + sy * invP;
    const float res = s2 - (sx * sx + sy * sy) * invP;
    // bin hash j = dec(M_c^T k mod q) (qsft.py:178-179)
    long long hsum = 0;
    if (act) {
#pragma unroll
        for (int u = 0; u < 32 / KL_G; ++u) {
            const int i = gl + u * KL_G;
            if (i < d.b)
                hsum += wgt[u] * fast_mod(dot_raw<NW>(d.MT + ((size_t)c * d.b + i) * d.ld, d.ld, kw), d.q, d.qmagic);
        }
    }
#pragma unroll
    for (int o = KL_G / 2; o > 0; o >>= 1) hsum += __shfl_xor_sync(0xffffffffu, hsum, o);
    const float lim = fmaxf(thresh, a.rel_floor * e_b);
    const bool single = act && (hsum == jb) && !(res > lim);
    const bool lead = (gl == 0);
    const unsigned sb = __ballot_sync(0xffffffffu, lead && single);
    unsigned long long fbase = 0;
    if (sb) fbase = (unsigned long long)kl_take(a, slots, __popc(sb));
    unsigned long long fs = fbase + (unsigned long long)__popc(sb & ((1u << lane) - 1u));
    fs = __shfl_sync(0xffffffffu, fs, lane & ~(KL_G - 1));
    if (single) {
        if ((long long)fs < slots.lim) {
            uint32_t* ko = reinterpret_cast<uint32_t*>(a.find_k + (size_t)fs * d.ld);
            for (int w = gl; w < d.ld / 4; w += KL_G) ko[w] = (w < NW) ? kw[w] : 0u;
            if (lead) {
                a.find_cj[fs] = (long long)c * B + jb;
                a.find_rho[fs] = make_float2(rr, ri);
                a.find_round[fs] = round;
                a.find_res[fs] = res;
                a.find_id[(size_t)c * B + jb] = (int32_t)fs;
            }
        }
    }
    __syncwarp();
    return single ? 1 : act ? 2 : 0;
}

// group g of a tile (first round: no bin carries a ball yet).  The warp copies the columns of its four items out of the stage
// into its private buffer and hands the stage back at once -- the ring slot is then held for the copy, not for the
// microseconds of latency-bound work per item that follow (with in-stage work the ring, 4 .. 6 tiles deep, stalled behind
// its slowest tile: ncu showed the candidate warps idle at their mailboxes for a quarter of all samples and the scanners
// waiting for data).
template <int NW>
__device__ __forceinline__ void kl_cand_any(const KlArgs& a, uint8_t* stage, KlTileInfo* info, int c, long long j0, int round, int g,
                                            uint8_t* s_symw, float2* priv, const float2* s_tw, bool structured,
                                            const long long (&wgt)[32 / KL_G], unsigned& n_multi, uint64_t* empty_bar, KlSlots& slots) {
    const PeelDev& d = a.d;
    const int lane = threadIdx.x & 31;
    const int grp = lane / KL_G, gl = lane % KL_G;
    const int R = d.R, P_src = d.P_src;
    uint8_t* sym = s_symw + grp * KL_SYM;
    const unsigned mask[4] = {info->mask[0], info->mask[1], info->mask[2], info->mask[3]};
    const int total = __popc(mask[0]) + __popc(mask[1]) + __popc(mask[2]) + __popc(mask[3]);
    const int rank = 4 * g + grp;
    const int my = rank < total ? kl_pick(mask, rank) : -1;
    const bool act = my >= 0;
    const int lbm = act ? my : 0;
    const long long jb = j0 + lbm;
    const TileCol tc{stage, P_src, a.box, lbm, a.lgW - 4};
    const float e_b = info->e[lbm];
    const PrivCol pc{priv + grp, P_src};
    if (act)
        for (int r = 0; r < R; ++r)
            for (int i = gl; i < P_src; i += KL_G) pc.ref(r, i) = tc.ri(r, i);
    __syncwarp();
#ifndef QSFT_EMU
    if (empty_bar != nullptr && lane == 0 && atomicSub(&info->left, 1) == 1) tma::mbar_arrive(empty_bar);
#else
    (void)empty_bar;
#endif
    const int cls = kl_cand_body<NW>(a, pc, c, jb, round, act, false, -1, e_b, sym, s_tw, structured, wgt, slots);
    if (gl == 0 && cls == 2) {
        ++n_multi;
        a.cls[(size_t)c * d.B + jb] = 2;
    }
}

// ---- rounds after the first: the bins the last link phase touched ------------------------------------------------------------
// Candidate warps take the dirty list four entries at a time.  An entry whose `zres` proves it a zeroton costs two loads; the
// others are gathered from U (one 8-byte read per delay row: these are few), get their listed balls subtracted and are
// classified.  Returns nothing; the round's change of the multiton count is added to a.multi[round] (two's complement).
template <int NW>
__device__ __forceinline__ void kl_dirty_round(const KlArgs& a, const KlBlocks& blk, int round, long long nd, uint8_t* s_sym,
                                               uint8_t* s_priv, const float2* s_tw) {
    const PeelDev& d = a.d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < KL_NS || warp >= KL_NS + KL_NC) return;
    const int cw = warp - KL_NS;
    const int grp = lane / KL_G, gl = lane % KL_G;
    const int R = d.R, P_src = d.P_src;
    const long long B = d.B;
    const float thresh = (float)d.thresh;
    const bool structured = (*a.dstruct != 0);
    long long wgt[32 / KL_G];
#pragma unroll
    for (int u = 0; u < 32 / KL_G; ++u) wgt[u] = hash_weight(d, gl + u * KL_G);
    uint8_t* sym = s_sym + (size_t)cw * 4 * KL_SYM + grp * KL_SYM;
    const PrivCol pc{reinterpret_cast<float2*>(s_priv + (size_t)cw * a.priv_bytes) + grp, P_src};
    KlSlots slots{0, 0, 0};
    int delta = 0;
    const long long ngroups = (nd + 3) >> 2;
    for (long long g = (long long)blockIdx.x * KL_NC + cw; g < ngroups; g += (long long)gridDim.x * KL_NC) {
        const long long e = 4 * g + grp;
        const bool have = e < nd;
        const long long cj = have ? __ldcg(a.dirty + e) : 0;
        float z = 0.f;
        if (have) z = __uint_as_float(__ldcg(a.zres + cj));
        const bool heavy = have && z > thresh;                  // energy test (qsft.py:164) on the known residual
        const unsigned any_heavy = __ballot_sync(0xffffffffu, heavy);
        if (have && gl == 0) a.zres[cj] = 0u;                   // looked at (by all lanes: after the ballot) -> ready for the next link phase
        if (!any_heavy) continue;
        const int c = (int)(cj / B);
        const long long j = cj - (long long)c * B;
        int hd = 0, was = 0;
        if (heavy) {
            hd = __ldcg(a.head + cj);
            was = a.cls[cj];
            for (int r = 0; r < R; ++r) {
                const float2* src = blk.p[c * R + r] + j;
                for (int i = gl; i < P_src; i += KL_G) pc.ref(r, i) = __ldcg(src + (size_t)i * a.ldU);
            }
        }
        __syncwarp();
        const int cls = kl_cand_body<NW>(a, pc, c, j, round, heavy, heavy, hd - 1, 0.f, sym, s_tw, structured, wgt, slots);
        if (heavy && gl == 0) {
            delta += (cls == 2 ? 1 : 0) - (was == 2 ? 1 : 0);
            if ((cls == 2) != (was == 2)) a.cls[cj] = (uint8_t)(cls == 2 ? 2 : 0);
        }
    }
    kl_close(a, slots);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) delta += __shfl_xor_sync(0xffffffffu, delta, o);
    if (lane == 0 && delta != 0) atomicAdd(&a.multi[round], (unsigned long long)(long long)delta);
}

// ---- one classification round ---------------------------------------------------------------------------------------
// TMA variant, warp roles: 1 producer warp; 2 scanner teams of 2 warps (alternate tiles); KL_NC candidate warps.
//   producer -> scanners : the stage's `full` mbarrier (bulk copies landed)
//   scanners -> candidates: ONE work queue per CTA in shared memory.  After its scan the team's leader appends the tile's groups
//                          of four work items, in tile order (the two teams take turns through `s_pub`), and after the CTA's
//                          last tile of the round one sentinel per candidate warp.  A candidate warp draws a ticket (atomic
//                          on the queue head) and waits for that entry: whichever warp is free takes the next group, so no
//                          group -- and with it no ring slot -- waits behind a warp that is busy with another tile's items
//                          (with a mailbox per warp, filled round-robin, ncu still showed the candidate warps idle for 23 %
//                          of their time while the scanners waited for ring slots).  Entries carry the lap number of the
//                          ring in their generation field, so a slot is never cleared.
//   candidates -> producer: the tile's `left` counter; whoever copies the last group out of the stage (or the scanner, when
//                          the tile has no work) arrives on the stage's `empty` mbarrier.
// Plain variant: scan and candidate work separated by CTA barriers, groups assigned statically, single stage.
template <int NW, bool TMA>
__device__ __forceinline__ void kl_classify(const KlArgs& a, const KlBlocks& blk,
#ifndef QSFT_EMU
                                            const CUtensorMap* maps,
#endif
                                            int round, uint8_t* stages, uint64_t* bars, unsigned int& tiles_done,
                                            KlTileInfo* infos, unsigned int* mbox, uint8_t* s_sym, uint8_t* s_priv, const float2* s_tw) {
    const PeelDev& d = a.d;
    const int W = a.W, R = d.R, P_src = d.P_src;
    const long long B = d.B;
    const long long tpg = (a.je - a.jb + W - 1) >> a.lgW;          // tiles per group (of this rank's bin range)
    const long long n_tiles = tpg * d.C;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long mine = n_tiles > (long long)blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    unsigned n_multi = 0;
    KlSlots slots{0, 0, 0};
    const bool is_cand = warp >= KL_NS && warp < KL_NS + KL_NC;
    bool structured = false;
    long long wgt[32 / KL_G] = {0, 0, 0, 0};
    if (is_cand) {
        structured = (*a.dstruct != 0);
#pragma unroll
        for (int u = 0; u < 32 / KL_G; ++u) wgt[u] = hash_weight(d, (lane % KL_G) + u * KL_G);
    }
    uint8_t* s_symw = s_sym + (size_t)(is_cand ? warp - KL_NS : 0) * 4 * KL_SYM;
    float2* s_privw = reinterpret_cast<float2*>(s_priv + (size_t)(is_cand ? warp - KL_NS : 0) * a.priv_bytes);
#ifndef QSFT_EMU
    if (TMA) {
        uint64_t* full = bars;
        uint64_t* empty = bars + KL_MAX_STAGES;
        volatile unsigned int* q_tail = mbox + KL_QN;              // entries appended so far (only the publishing leader writes)
        unsigned int* q_head = mbox + KL_QN + 1;                   // tickets drawn so far
        volatile unsigned int* s_pub = mbox + KL_QN + 2;           // tiles published so far (CTA-wide sequence number)
        unsigned int it = tiles_done;
        if (warp == KL_NS + KL_NC) {
            // ---- producer warp --------------------------------------------------------------------------------------
            if (lane == 0) {
                const uint32_t box_bytes = (uint32_t)((W >> 4) * P_src * 128);
                for (long long tt = blockIdx.x; tt < n_tiles; tt += gridDim.x, ++it) {
                    const int st = (int)(it % (unsigned)a.nstages);
                    const uint32_t ph = (it / (unsigned)a.nstages) & 1u;
                    tma::mbar_wait(&empty[st], ph ^ 1u);
                    const int c = (int)(tt / tpg);
                    const long long j0 = a.jb + ((tt - (long long)c * tpg) << a.lgW);
                    uint8_t* dst = stages + (size_t)st * a.stage_bytes;
                    const long long left = a.je - j0;
                    const uint32_t head_bytes = round > 1 ? (uint32_t)((left < W ? left : W) * 4) : 0u;
                    tma::mbar_expect_tx(&full[st], (uint32_t)R * box_bytes + head_bytes);
                    for (int r = 0; r < R; ++r)
                        tma::load_3d(dst + (size_t)r * a.box, &maps[c * R + r], 0, (int)(j0 >> 4), 0, &full[st]);
                    if (head_bytes) tma::bulk_g2s(dst + (size_t)R * a.box, a.head + (size_t)c * B + j0, head_bytes, &full[st]);
                }
            }
        } else if (warp < KL_NS) {
            // ---- scanner warps: team = warp / 2 takes the tiles of its parity ----------------------------------------
            const int team = warp >> 1, t = threadIdx.x & 63;
            for (long long tt = blockIdx.x; tt < n_tiles; tt += gridDim.x, ++it) {
                if ((int)(it & 1u) != team) continue;
                const int st = (int)(it % (unsigned)a.nstages);
                const uint32_t ph = (it / (unsigned)a.nstages) & 1u;
                const int c = (int)(tt / tpg);
                const long long j0 = a.jb + ((tt - (long long)c * tpg) << a.lgW);
                KlTileInfo* info = &infos[st];
                tma::mbar_wait(&full[st], ph);
                kl_scan(a, stages + (size_t)st * a.stage_bytes, j0, round, info, t);
                if (team == 0) asm volatile("bar.sync 2, 64;" ::: "memory");
                else asm volatile("bar.sync 3, 64;" ::: "memory");
                if (t == 0) {
                    while (*s_pub != it) __nanosleep(20);           // tiles are published in order
                    __threadfence_block();
                    const int total = __popc(info->mask[0]) + __popc(info->mask[1]) + __popc(info->mask[2]) + __popc(info->mask[3]);
                    const int ng = (total + 3) >> 2;
                    info->j0 = j0;
                    info->c = c;
                    info->left = ng;
                    __threadfence_block();
                    if (ng == 0) tma::mbar_arrive(&empty[st]);
                    unsigned int tail = *q_tail;
                    for (int g = 0; g < ng; ++g, ++tail)
                        mbox[tail & (KL_QN - 1)] = 0x80000000u | (((tail / KL_QN) & 0x7fu) << 16) | ((unsigned)st << 8) | (unsigned)g;
                    if (tt + gridDim.x >= n_tiles)                   // the CTA's last tile of the round: everybody goes home
                        for (int cw = 0; cw < KL_NC; ++cw, ++tail)
                            mbox[tail & (KL_QN - 1)] = 0x80000000u | (((tail / KL_QN) & 0x7fu) << 16) | 0xffffu;
                    *q_tail = tail;
                    __threadfence_block();
                    *s_pub = it + 1;
                }
            }
        } else if (mine > 0) {
            // ---- candidate warps ------------------------------------------------------------------------------------
            for (;;) {
                unsigned int ticket = 0;
                if (lane == 0) ticket = atomicAdd(q_head, 1u);
                ticket = __shfl_sync(0xffffffffu, ticket, 0);
                volatile unsigned int* slot = &mbox[ticket & (KL_QN - 1)];
                const unsigned int want = 0x8000u | ((ticket / KL_QN) & 0x7fu);
                unsigned int e;
                while (((e = *slot) >> 16) != want) __nanosleep(32);
                __syncwarp();
                __threadfence_block();
                if ((e & 0xffffu) == 0xffffu) break;
                const int st = (int)((e >> 8) & 7u), g = (int)(e & 0xffu);
                KlTileInfo* info = &infos[st];
                kl_cand_any<NW>(a, stages + (size_t)st * a.stage_bytes, info, info->c, info->j0, round, g, s_symw, s_privw, s_tw, structured,
                                    wgt, n_multi, &empty[st], slots);
            }
        }
        tiles_done += (unsigned int)mine;
    }
#endif
    if (!TMA) {
        unsigned int it = tiles_done;
        for (long long tt = blockIdx.x; tt < n_tiles; tt += gridDim.x, ++it) {
            const int c = (int)(tt / tpg);
            const long long j0 = a.jb + ((tt - (long long)c * tpg) << a.lgW);
            // coalesced copy into the tile layout (single stage): the previous tile's readers are done first
            __syncthreads();
            {
                const int lb = threadIdx.x & (W - 1), prt = threadIdx.x >> a.lgW, nparts = (int)blockDim.x >> a.lgW;
                const long long jf = j0 + lb;
                TileCol tc{stages, P_src, a.box, lb, a.lgW - 4};
                for (int r = 0; r < R; ++r) {
                    const float2* src = blk.p[c * R + r] + jf;
                    for (int i = prt; i < P_src; i += nparts) tc.ref(r, i) = (jf < a.je) ? src[(size_t)i * a.ldU] : make_float2(0.f, 0.f);
                }
                if (round > 1 && threadIdx.x < W)
                    reinterpret_cast<int32_t*>(stages + (size_t)R * a.box)[threadIdx.x] =
                        (j0 + threadIdx.x < a.je) ? __ldcg(a.head + (size_t)c * B + j0 + threadIdx.x) : 0;
            }
            __syncthreads();
            if (warp < 2) kl_scan(a, stages, j0, round, &infos[0], (int)threadIdx.x);
            __syncthreads();
            if (is_cand) {
                const int total = __popc(infos[0].mask[0]) + __popc(infos[0].mask[1]) + __popc(infos[0].mask[2]) + __popc(infos[0].mask[3]);
                for (int g = (int)((warp - KL_NS + 5u * it) % (unsigned)KL_NC); 4 * g < total; g += KL_NC)
                    kl_cand_any<NW>(a, stages, &infos[0], c, j0, round, g, s_symw, s_privw, s_tw, structured, wgt, n_multi, nullptr, slots);
            }
        }
        tiles_done += (unsigned int)mine;
    }
    if (is_cand) kl_close(a, slots);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_multi += __shfl_xor_sync(0xffffffffu, n_multi, o);
    if (lane == 0 && n_multi) atomicAdd(&a.multi[round], (unsigned long long)n_multi);
}

// ---- link phase: one thread per find slot of the round ---------------------------------------------------------------
// lo[p] / hi[p]: the round's slots in rank p's segment.  All reads of the find list go through L2 (__ldcg): with several
// ranks the peers' entries were written over NVLink during this kernel.
struct KlRound {
    long long lo[8], hi[8];
};

template <int NW>
__device__ __forceinline__ void kl_link(const KlArgs& a, const KlRound& rd, int p, int round, bool do_link) {
    const PeelDev& d = a.d;
    const int nw = d.ld / 4;
    const long long B = d.B;
    const long long f0 = rd.lo[p], f1 = rd.hi[p];
    const float Pf = (float)d.P;
    const int lane = threadIdx.x & 31;
    // warp-uniform trip count: the dirty-list appends below are aggregated per warp (one atomic per warp and group instead of
    // one per bin: hundreds of thousands of atomics on ONE counter cost more than the rest of the link phase)
    for (long long fb = f0 + (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); fb < f1; fb += (long long)gridDim.x * blockDim.x) {
        const long long f = fb + lane;
        const long long cj = f < f1 ? __ldcg(a.find_cj + f) : -1;
        unsigned fresh = 0;                                 // bit l: this thread touched bin (l, jl[l]) first in this round
        long long jl[KL_MAX_BLOCKS];                       // C <= 16 on this path
        if (cj >= 0) {                                      // (unused slots of a warp's chunk carry -1)
        const int c = (int)(cj / B);
        uint32_t kw[NW];
        const uint32_t* kin = reinterpret_cast<const uint32_t*>(a.find_k + (size_t)f * d.ld);
#pragma unroll
        for (int w = 0; w < NW; ++w) kw[w] = (w < nw) ? __ldcg(kin + w) : 0u;
        const float2 rho = __ldcg(a.find_rho + f);
        float2 sum = rho;
        int cnt = 1;
        bool first = true, last = true;
        float zl[KL_MAX_BLOCKS];                           // energy left in bin (c2, jl[c2]) once (k, rho) is subtracted, if known
        for (int c2 = 0; c2 < d.C; ++c2) {
            zl[c2] = __uint_as_float(0x7f800000u);           // +inf: not known without reading the bin
            if (c2 == c) {
                jl[c2] = cj - (long long)c * B;
                zl[c2] = __ldcg(a.find_res + f);            // this find's own bin: its residual
                continue;
            }
            const long long j2 = hash_bin<NW>(d, c2, kw);
            jl[c2] = j2;
            const long long f2 = (long long)__ldcg(a.find_id + (size_t)c2 * B + j2);
            // find_id is only written for singletons: an entry is a find of THIS round iff it lies in the round's range of its
            // segment and that find really sits in bin (c2, j2)
            bool cur = false;
            if (f2 >= 0) {
                const int own = (int)(f2 / a.seg);
                cur = own < a.world && f2 >= rd.lo[own] && f2 < rd.hi[own] && __ldcg(a.find_cj + f2) == (long long)c2 * B + j2;
            }
            if (cur) {
                const uint32_t* k2 = reinterpret_cast<const uint32_t*>(a.find_k + (size_t)f2 * d.ld);
                bool same = true;
#pragma unroll
                for (int w = 0; w < NW; ++w) same &= ((w < nw) ? __ldcg(k2 + w) : 0u) == kw[w];
                if (same) {
                    const float2 r2 = __ldcg(a.find_rho + f2);
                    // bin (c2, j2) was found with the same k, value r2, residual res2: minus (k, rho) it holds res2 + P |r2 - rho|^2
                    const float dx = r2.x - rho.x, dy = r2.y - rho.y;
                    zl[c2] = fmaf(Pf, fmaf(dx, dx, dy * dy), __ldcg(a.find_res + f2));
                    if (c2 < c) {
                        first = false;                      // an earlier group holds the round's first find of this k
                    } else {
                        last = false;                       // ball_values: a later (i, j) wins (qsft.py:215)
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
            // peel: the ball joins the list of every bin it hashes to (qsft.py:223-241); bins of this rank's range are entered
            // into the dirty list by whoever touches them first in this round
            for (int l = 0; l < d.C; ++l) {
                const size_t bin = (size_t)l * B + jl[l];
                const int32_t prev = atomicExch(a.head + bin, (int32_t)(f + 1));
                a.next[(size_t)f * d.C + l] = prev;
                if (jl[l] >= a.jb && jl[l] < a.je) {
                    const float z = fmaxf(zl[l], 1e-37f);   // never 0 (0 = untouched); negative rounding noise and NaN -> tiny
                    if (atomicMax(a.zres + bin, __float_as_uint(z)) == 0u) fresh |= 1u << l;
                }
            }
            fresh |= 0x80000000u;                           // a ball was peeled
        }
        }
        {
            const unsigned m = __ballot_sync(0xffffffffu, (fresh & 0x80000000u) != 0u);
            if (lane == 0 && m) atomicAdd(&a.counters[2], (unsigned long long)__popc(m));   // num_peeling (qsft.py:224)
        }
        for (int l = 0; l < d.C; ++l) {
            const bool mine = (fresh >> l) & 1u;
            const unsigned m = __ballot_sync(0xffffffffu, mine);
            if (m == 0u) continue;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(a.dcount + (round & 1), (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            const long long slot = (long long)base + __popc(m & ((1u << lane) - 1u));
            if (mine && slot < a.max_dirty) a.dirty[slot] = (long long)l * B + jl[l];
        }
    }
}

#ifndef QSFT_EMU
// ---- exchange of a round's finds between the ranks (world > 1) -----------------------------------------------------------
// Every warp copies slots of this rank's segment into the same slots of every peer's workspace (k row, bin, rho, round) and
// enters valid finds into the peers' (C, B) find tables.
__device__ __forceinline__ void kl_push(const KlArgs& a, long long s0, long long s1) {
    const PeelDev& d = a.d;
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5), w0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int kvec = d.ld / 16;                             // uint4 per k row (2 .. 8)
    for (long long f = s0 + w0; f < s1; f += warps) {
        const long long cj = a.find_cj[f];
        uint4 kv = make_uint4(0u, 0u, 0u, 0u);
        if (lane < kvec) kv = reinterpret_cast<const uint4*>(a.find_k + (size_t)f * d.ld)[lane];
        const float2 rho = a.find_rho[f];
        const int32_t rnd = a.find_round[f];
        const float res = a.find_res[f];
        for (int p = 0; p < a.world; ++p) {
            if (p == a.rank) continue;
            uint8_t* ws = a.sh.peer[p];
            if (lane < kvec) reinterpret_cast<uint4*>(ws + a.sh.off_k + (size_t)f * d.ld)[lane] = kv;
            if (lane == 8) reinterpret_cast<long long*>(ws + a.sh.off_cj)[f] = cj;
            if (lane == 9) reinterpret_cast<float2*>(ws + a.sh.off_rho)[f] = rho;
            if (lane == 10) reinterpret_cast<int32_t*>(ws + a.sh.off_round)[f] = rnd;
            if (lane == 11 && cj >= 0) reinterpret_cast<int32_t*>(ws + a.sh.off_id)[cj] = (int32_t)f;
            if (lane == 12) reinterpret_cast<float*>(ws + a.sh.off_res)[f] = res;
        }
    }
    __threadfence_system();
}
#endif

template <int NW, bool TMA>
__global__ void __launch_bounds__(TMA ? KL_CT + 32 : KL_CT, 1)
k4_peel_loop_kernel(const KlArgs a, const KlBlocks blk
#ifndef QSFT_EMU
                    , const __grid_constant__ KlMaps maps
#endif
) {
    extern __shared__ __align__(1024) uint8_t kl_smem[];
    __shared__ float2 s_tw[QSFT_MAX_Q + 1];
    const PeelDev& d = a.d;
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
#else
    uint8_t* base = kl_smem;
#endif
    uint8_t* ctrl = base + (size_t)a.nstages * a.stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ctrl);                           // full[6], empty[6]
    KlTileInfo* infos = reinterpret_cast<KlTileInfo*>(ctrl + 256);                // [KL_MAX_STAGES], at most 576 bytes each
    unsigned int* mbox = reinterpret_cast<unsigned int*>(ctrl + 256 + KL_MAX_STAGES * 576);   // work queue, its tail / head, s_pub
    uint8_t* s_sym = ctrl + 256 + KL_MAX_STAGES * 576 + KL_QN * 4 + 128;          // [KL_NC][4][KL_SYM]
    uint8_t* s_priv = ctrl + KL_CTRL_BYTES;                                       // [KL_NC][priv_bytes]
    static_assert(sizeof(KlTileInfo) <= 576 && KL_CTRL_BYTES % 128 == 0, "control block layout");
#ifndef QSFT_EMU
    if (TMA) {
        for (int i = threadIdx.x; i < KL_QN + 32; i += blockDim.x) mbox[i] = 0u;
        if (threadIdx.x == 0) {
            for (int i = 0; i < KL_MAX_STAGES; ++i) {
                tma::mbar_init(&bars[i], 1);
                tma::mbar_init(&bars[KL_MAX_STAGES + i], 1);
            }
            tma::mbar_fence_init();
        }
    }
#endif
    __syncthreads();
    unsigned int epoch = 0, tiles_done = 0;
    KlRound rd;
    long long used[8];                                      // slots used so far in every rank's segment
    for (int p = 0; p < 8; ++p) used[p] = 0;
    double num_peeling = 0;
    int round = 0;
    bool cont = true, overflow = false;
    long long my_multi = 0;                                 // multiton bins of this rank's range (carried from round to round)
    long long nd = 0;                                       // dirty bins left by the last link phase
    while (cont && num_peeling < a.peeling_max && round < a.max_rounds) {
        ++round;
        if (round == 1)
            kl_classify<NW, TMA>(a, blk,
#ifndef QSFT_EMU
                                 maps.m,
#endif
                                 1, base, bars, tiles_done, infos, mbox, s_sym, s_priv, s_tw);
        else
            kl_dirty_round<NW>(a, blk, round, nd, s_sym, s_priv, s_tw);
        kl_grid_barrier(a.gbar, epoch);
        // everybody has read `nd`: its counter is free for the link phase of the next round
        if (blockIdx.x == 0 && threadIdx.x == 0) a.dcount[(round + 1) & 1] = 0ull;
        long long now[8], multis = 0, nf = 0;
        now[a.rank] = (long long)__ldcg(a.counters + 0);
        my_multi += (long long)__ldcg(a.multi + round);      // round 1: the count; later rounds: the change
        multis = my_multi;
#ifndef QSFT_EMU
        if (a.world > 1) {
            // this rank's new slots (unused ones included: they carry find_cj = -1) -> every peer, then counters + flag
            const long long segb = a.seg * a.rank;
            const long long top = now[a.rank] < a.seg ? now[a.rank] : a.seg;
            kl_push(a, segb + used[a.rank], segb + top);
            kl_grid_barrier(a.gbar, epoch);
            const unsigned int tag = (a.sh.epoch << 8) | (unsigned int)round;
            if (blockIdx.x == 0 && threadIdx.x < a.world) {
                KlCtl* ctl = reinterpret_cast<KlCtl*>(a.sh.peer[threadIdx.x] + a.sh.off_ctl);
                ctl->now[round][a.rank] = (unsigned long long)now[a.rank];
                ctl->multi[round][a.rank] = (unsigned long long)multis;
                __threadfence_system();
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&ctl->flag[round][a.rank]), "r"(tag) : "memory");
            }
            KlCtl* mine = reinterpret_cast<KlCtl*>(a.sh.peer[a.rank] + a.sh.off_ctl);
            if (threadIdx.x < a.world) {
                unsigned int seen;
                for (;;) {
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(&mine->flag[round][threadIdx.x]) : "memory");
                    if (seen == tag) break;
                    __nanosleep(100);
                }
            }
            __syncthreads();
            multis = 0;
            for (int p = 0; p < a.world; ++p) {
                now[p] = (long long)__ldcg(&mine->now[round][p]);
                multis += (long long)__ldcg(&mine->multi[round][p]);
            }
        }
#endif
        for (int p = 0; p < a.world; ++p) {
            if (now[p] > a.seg) overflow = true;            // find buffer too small (uniform over the grid and the ranks)
            rd.lo[p] = a.seg * p + used[p];
            rd.hi[p] = a.seg * p + (now[p] < a.seg ? now[p] : a.seg);
            nf += now[p] - used[p];
        }
        if (overflow) {
            if (blockIdx.x == 0 && threadIdx.x == 0) a.counters[6] = 1ull;
            break;
        }
        if (multis == 0 || nf == 0) cont = false;           // qsft.py:204-205
        // the reference also subtracts after its last round, but nothing reads the bins afterwards: skip unless the q^n
        // guard needs the count
        const bool do_link = cont || a.guard_can_bind;
        if (nf > 0)
            for (int p = 0; p < a.world; ++p) kl_link<NW>(a, rd, p, round, do_link);
        for (int p = 0; p < a.world; ++p) used[p] = now[p];
        if (cont || a.guard_can_bind) {
            kl_grid_barrier(a.gbar, epoch);
            if (a.guard_can_bind) num_peeling = (double)__ldcg(a.counters + 2);
            const long long listed = (long long)__ldcg(a.dcount + (round & 1));
            nd = listed < a.max_dirty ? listed : a.max_dirty;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.counters[5] = (unsigned long long)round;
        a.counters[7] = (unsigned long long)used[a.rank];
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

// Slots per chunk: large enough that the global counter is touched rarely, small enough that the slots left unused at the end
// of a round (at most one chunk per candidate warp and round) stay below ~C B / 16 per round.
inline int kl_chunk(const PeelDev& d, int grid) {
    const long long c = (long long)d.C * d.B / ((long long)grid * KL_NC * 16);
    return c < 4 ? 4 : c > KL_CHUNK ? KL_CHUNK : (int)c;
}

// tile geometry for a shared-memory budget: the widest tile (<= 128 bins, no wider than the group needs) that leaves at
// least `min_stages` stages beside the candidate warps' private column buffers (KlArgs.priv_bytes each: four items x all
// delay rows).  Returns false when even 16-bin tiles do not fit (very long delay-row lists: the caller uses the host-driven
// rounds).
inline bool kl_geometry(const PeelDev& d, int budget, int min_stages, KlArgs* a) {
    const long long priv = (((long long)d.R * d.P_src * 4 * 8) + 127) & ~127ll;
    const long long avail = (long long)budget - KL_CTRL_BYTES - priv * KL_NC;
    for (int W = KL_MAXW; W >= 16; W >>= 1) {
        if (W > 16 && (long long)(W >> 1) >= d.B) continue;
        const int box = ((W >> 4) * d.P_src * 128 + 1023) & ~1023;
        const long long stage = ((long long)d.R * box + W * 4 + 1023) & ~1023ll;
        const long long n = avail / stage;
        if (n >= min_stages) {
            a->W = W;
            a->lgW = W == 128 ? 7 : W == 64 ? 6 : W == 32 ? 5 : 4;
            a->box = box;
            a->stage_bytes = (int)stage;
            a->nstages = n > KL_MAX_STAGES ? KL_MAX_STAGES : (int)n;
            a->priv_bytes = (int)priv;
            return true;
        }
    }
    return false;
}

}  // namespace (device part; the CPU emulation cuts here)

// ---- launch (one translation unit per NW: k4_peel_loop_nw*.cu) -------------------------------------------------------------
#ifndef QSFT_EMU
namespace {

template <int NW>
int kl_launch(const KlArgs& a, const KlBlocks& blk, const KlMaps& maps, bool use_tma, size_t smem, int grid, cudaStream_t st) {
    void* params[] = {(void*)&a, (void*)&blk, (void*)&maps};
    const void* fn = use_tma ? (const void*)k4_peel_loop_kernel<NW, true> : (const void*)k4_peel_loop_kernel<NW, false>;
    QSFT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QSFT_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(use_tma ? KL_CT + 32 : KL_CT), params, smem, st));
    g_qsft_launches.fetch_add(1, std::memory_order_relaxed);
    return QSFT_OK;
}

}  // namespace
#endif
