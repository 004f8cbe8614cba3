// K3: batched b-dimensional length-q DFT (tensor product F_q^{(x)b}, no inter-axis twiddles), forward sign
// e^{-2 pi i/q}, scaled by 1/q^b, in place on complex64.  Replaces qsft/utils.py:31-36 (scipy.fft.fftn / q^n) as
// called from qsft/input_signal_subsampled.py:264-266.  HBM bound: each pass reads and writes every element once;
// a pass stages a tile in shared memory and runs radix-q butterflies over a group of axes.
//
// Axes are addressed by "level" u = stride exponent: element index j = sum_u digit_u q^u (digit_{b-1} is the most
// significant = first axis of the C-order reshape [q]*b).
#include "common.cuh"

#include <stdlib.h>

#include "k3_shared.cuh"

namespace {

constexpr int K3_THREADS = 256;
constexpr int K3_TILE = 4096;  // complex elements staged per CTA (32 KB)

template <int Q>
struct Dft {
    // generic small prime-ish radix with a register twiddle table tw[m] = e^{-2 pi i m / Q}
    __device__ static __forceinline__ void run(float2 (&x)[Q], const float2* tw) {
        float2 y[Q];
#pragma unroll
        for (int m = 0; m < Q; ++m) {
            float2 acc = x[0];
#pragma unroll
            for (int j = 1; j < Q; ++j) {
                const float2 w = tw[(j * m) % Q];
                acc.x = fmaf(x[j].x, w.x, fmaf(-x[j].y, w.y, acc.x));
                acc.y = fmaf(x[j].x, w.y, fmaf(x[j].y, w.x, acc.y));
            }
            y[m] = acc;
        }
#pragma unroll
        for (int m = 0; m < Q; ++m) x[m] = y[m];
    }
};
template <>
struct Dft<2> {
    __device__ static __forceinline__ void run(float2 (&x)[2], const float2*) {
        float2 a = x[0], b = x[1];
        x[0] = make_float2(a.x + b.x, a.y + b.y);
        x[1] = make_float2(a.x - b.x, a.y - b.y);
    }
};
template <>
struct Dft<4> {
    __device__ static __forceinline__ void run(float2 (&x)[4], const float2*) {
        // forward: w = -i
        float2 s02 = make_float2(x[0].x + x[2].x, x[0].y + x[2].y);
        float2 d02 = make_float2(x[0].x - x[2].x, x[0].y - x[2].y);
        float2 s13 = make_float2(x[1].x + x[3].x, x[1].y + x[3].y);
        float2 d13 = make_float2(x[1].x - x[3].x, x[1].y - x[3].y);
        x[0] = make_float2(s02.x + s13.x, s02.y + s13.y);
        x[2] = make_float2(s02.x - s13.x, s02.y - s13.y);
        x[1] = make_float2(d02.x + d13.y, d02.y - d13.x);  // d02 - i*d13
        x[3] = make_float2(d02.x - d13.y, d02.y + d13.x);  // d02 + i*d13
    }
};

// One pass over levels [a, a + r).  Tile = q^r rows (stride q^a) x W contiguous elements; for a == 0, W == 1 and
// `outer` consecutive q^r blocks are processed per CTA.  Shared layout: e = (o * q^r + t) * W + w.
// POW2 (q = 2, 4): every division / modulo below is a shift / mask (lgW = log2 W, lgq = log2 q) and global accesses
// are 16-byte (two complex) vectors.
template <int Q, bool POW2>
__global__ void __launch_bounds__(K3_THREADS)
k3_pass_kernel(float2* __restrict__ x, long long B, int q, int r, long long qa, int W, int lgW, int outer, int rows /*q^r*/,
               long long tiles_per_block, long long tile0, float scale) {
    extern __shared__ float2 s[];
    __shared__ float2 s_tw[Q > 0 ? Q : QSFT_MAX_Q];
    const int tid = threadIdx.x;
    const int T = outer * rows * W;
    const int qq = Q > 0 ? Q : q;
    constexpr int lgq = (Q == 2) ? 1 : 2;
    if (tid < qq) {
        float sn, cs;
        sincospif(-2.0f * (float)tid / (float)qq, &sn, &cs);
        s_tw[tid] = make_float2(cs, sn);
    }
    // locate the tile
    const long long tile = tile0 + blockIdx.x;
    const long long blk = tile / tiles_per_block;      // which length-B block of the batch
    const long long tin = tile - blk * tiles_per_block;
    float2* base = x + blk * B;
    long long g0;
    if (qa == 1) {
        g0 = tin * (long long)T;                        // contiguous run
    } else {
        const long long mids = qa / W;                  // W-chunks inside one stride-q^a row segment
        const long long high = tin / mids, mid = tin - high * mids;
        g0 = high * qa * rows + mid * W;
    }
    // load
    if (POW2) {
        float4* s4 = reinterpret_cast<float4*>(s);
        if (qa == 1) {
            const float4* g4 = reinterpret_cast<const float4*>(base + g0);
            for (int e = tid; e < T / 2; e += K3_THREADS) s4[e] = g4[e];
        } else {
            for (int e = tid; e < T / 2; e += K3_THREADS) {
                const int t = (2 * e) >> lgW, w = (2 * e) & (W - 1);
                s4[e] = *reinterpret_cast<const float4*>(base + g0 + (long long)t * qa + w);
            }
        }
    } else if (qa == 1) {
        for (int e = tid; e < T; e += K3_THREADS) s[e] = base[g0 + e];
    } else {
        for (int e = tid; e < T; e += K3_THREADS) {
            int t = e / W, w = e - t * W;
            s[e] = base[g0 + (long long)t * qa + w];
        }
    }
    __syncthreads();
    float2 tw[Q > 0 ? Q : 1];
    if constexpr (Q > 0) {
#pragma unroll
        for (int m = 0; m < Q; ++m) tw[m] = s_tw[m];
    }
    // butterflies
    int stride = W, lgs = lgW;
    const int nbf = T / qq;
    for (int u = 0; u < r; ++u) {
        for (int i = tid; i < nbf; i += K3_THREADS) {
            int e0;
            if (POW2) {
                const int hi = i >> lgs, lo = i & (stride - 1);
                e0 = (hi << (lgs + lgq)) + lo;
            } else {
                const int hi = i / stride, lo = i - hi * stride;
                e0 = hi * stride * qq + lo;
            }
            if constexpr (Q > 0) {
                float2 v[Q];
#pragma unroll
                for (int m = 0; m < Q; ++m) v[m] = s[e0 + m * stride];
                Dft<Q>::run(v, tw);
#pragma unroll
                for (int m = 0; m < Q; ++m) s[e0 + m * stride] = v[m];
            } else {
                float2 v[QSFT_MAX_Q];
                for (int m = 0; m < q; ++m) v[m] = s[e0 + m * stride];
                for (int m = 0; m < q; ++m) {
                    float2 acc = v[0];
                    int idx = 0;
                    for (int j = 1; j < q; ++j) {
                        idx += m;
                        if (idx >= q) idx -= q;
                        const float2 w = s_tw[idx];
                        acc.x = fmaf(v[j].x, w.x, fmaf(-v[j].y, w.y, acc.x));
                        acc.y = fmaf(v[j].x, w.y, fmaf(v[j].y, w.x, acc.y));
                    }
                    s[e0 + m * stride] = acc;
                }
            }
        }
        __syncthreads();
        stride *= qq;
        lgs += lgq;
    }
    // store
    if (POW2) {
        const float4* s4 = reinterpret_cast<const float4*>(s);
        for (int e = tid; e < T / 2; e += K3_THREADS) {
            float4 v = s4[e];
            v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
            if (qa == 1) {
                reinterpret_cast<float4*>(base + g0)[e] = v;
            } else {
                const int t = (2 * e) >> lgW, w = (2 * e) & (W - 1);
                *reinterpret_cast<float4*>(base + g0 + (long long)t * qa + w) = v;
            }
        }
    } else if (qa == 1) {
        for (int e = tid; e < T; e += K3_THREADS) {
            float2 v = s[e];
            base[g0 + e] = make_float2(v.x * scale, v.y * scale);
        }
    } else {
        for (int e = tid; e < T; e += K3_THREADS) {
            int t = e / W, w = e - t * W;
            float2 v = s[e];
            base[g0 + (long long)t * qa + w] = make_float2(v.x * scale, v.y * scale);
        }
    }
}

// ---- q = 4 fast pass: radix-16 steps in registers ---------------------------------------------------------
// Tile = 4096 complex elements, 256 threads, 16 elements per thread and step.  The r <= 6 levels of the pass sit at
// base-4 digit positions [p0, p0 + r) of the tile index e; they are processed two at a time (radix-16 butterfly in
// registers), the first step reads global memory directly, the last one writes it directly, steps in between
// exchange through shared memory with the XOR swizzle e ^ ((e >> 4) & 15), which makes every half-warp access
// (16 lanes x 8 bytes) bank-conflict free for all digit positions.
__device__ __forceinline__ int k3_swz(int e) { return e ^ ((e >> 4) & 15); }

__device__ __forceinline__ void k3_r4(float2& a, float2& b, float2& c, float2& d) {
    const float2 s02 = make_float2(a.x + c.x, a.y + c.y), d02 = make_float2(a.x - c.x, a.y - c.y);
    const float2 s13 = make_float2(b.x + d.x, b.y + d.y), d13 = make_float2(b.x - d.x, b.y - d.y);
    a = make_float2(s02.x + s13.x, s02.y + s13.y);
    c = make_float2(s02.x - s13.x, s02.y - s13.y);
    b = make_float2(d02.x + d13.y, d02.y - d13.x);   // d02 - i d13
    d = make_float2(d02.x - d13.y, d02.y + d13.x);   // d02 + i d13
}

template <bool STRIDED>
__device__ __forceinline__ void k3_q4_tile(float2* __restrict__ s, float2* __restrict__ base, long long tin, int r, int p0,
                                           long long qa, int lgW, float scale, const float2* xroot, const K3Peers& peers) {
    const int tau = threadIdx.x;
    const int W = 1 << lgW;
    long long g0;
    if (!STRIDED) {
        g0 = tin * 4096ll;
    } else {
        const long long rows = 4096 >> lgW;
        const long long mids = qa >> lgW;
        const long long high = tin / mids, mid = tin - high * mids;
        g0 = high * qa * rows + mid * W;
    }
    auto gptr = [&](int e) -> float2* {
        return STRIDED ? base + g0 + (long long)(e >> lgW) * qa + (e & (W - 1)) : base + g0 + e;
    };
    const int nsteps = (r + 1) >> 1;
    for (int st = 0; st < nsteps; ++st) {
        const int p = p0 + 2 * st;
        const bool pair = (2 * st + 1) < r;
        const bool first = (st == 0), last = (st == nsteps - 1);
        const int sh = 2 * p, lowmask = (1 << sh) - 1, step = 1 << sh;
        float2 v[16];
        int eb[4];
        if (pair) {
            eb[0] = ((tau >> sh) << (sh + 4)) | (tau & lowmask);
            if (first) {
                if (!STRIDED && p == 0) {
                    const float4* g4 = reinterpret_cast<const float4*>(base + g0 + eb[0]);
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        const float4 t4 = __ldcg(g4 + m);
                        v[2 * m] = make_float2(t4.x, t4.y);
                        v[2 * m + 1] = make_float2(t4.z, t4.w);
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < 16; ++m) v[m] = __ldcg(gptr(eb[0] + m * step));
                }
            } else {
#pragma unroll
                for (int m = 0; m < 16; ++m) v[m] = s[k3_swz(eb[0] + m * step)];
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) k3_r4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);      // digit p
#pragma unroll
            for (int h = 0; h < 4; ++h) k3_r4(v[h], v[h + 4], v[h + 8], v[h + 12]);                      // digit p + 1
            if (last) {
#pragma unroll
                for (int m = 0; m < 16; ++m)
                    k3_store(gptr(eb[0] + m * step), make_float2(v[m].x * scale, v[m].y * scale), xroot, peers);
            } else {
#pragma unroll
                for (int m = 0; m < 16; ++m) s[k3_swz(eb[0] + m * step)] = v[m];
            }
        } else {
            // single level: four radix-4 butterflies per thread
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int beta = tau + 256 * k;
                eb[k] = ((beta >> sh) << (sh + 2)) | (beta & lowmask);
#pragma unroll
                for (int m = 0; m < 4; ++m)
                    v[4 * k + m] = first ? __ldcg(gptr(eb[k] + m * step)) : s[k3_swz(eb[k] + m * step)];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) k3_r4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    if (last) k3_store(gptr(eb[k] + m * step), make_float2(v[4 * k + m].x * scale, v[4 * k + m].y * scale), xroot, peers);
                    else s[k3_swz(eb[k] + m * step)] = v[4 * k + m];
                }
            }
        }
        if (!last) __syncthreads();
    }
}

template <bool STRIDED>
__global__ void __launch_bounds__(256)
k3_q4_fast_kernel(float2* __restrict__ x, long long B, int r, int p0, long long qa, int lgW, long long tiles_per_block,
                  long long tile0, float scale, K3Peers peers) {
    __shared__ float2 s[4096];
    const long long tile = tile0 + blockIdx.x;
    const long long blk = tile / tiles_per_block;
    k3_q4_tile<STRIDED>(s, x + blk * B, tile - blk * tiles_per_block, r, p0, qa, lgW, scale, x, peers);
}

// Both passes of a two-pass transform (4^7 .. 4^12 points) in ONE launch.  Work items (tiles) are handed out through an
// atomic ticket; a strided-pass tile of block k waits on a per-block counter until all contiguous-pass tiles of block k
// have been written, so the intermediate of a block is consumed from L2: DRAM sees one read and one write of the data.
//
// Ticket order (k3_ticket_decode): contiguous tiles run `lag` blocks AHEAD of the strided tiles,
//     C(0) .. C(lag-1) | C(lag) S(0) | C(lag+1) S(1) | ... | C(nb-1) S(nb-1-lag) | S(nb-lag) .. S(nb-1)
// so (after the start-up blocks) lag * (tiles1 + tiles2) tickets lie between a block's last contiguous tile and its first
// strided tile.  With
// that distance >= the number of co-resident CTAs a strided tile never finds its block unfinished; with the plain order
// (lag = 0: C(k) S(k) back to back) the strided tiles of a block become resident while its contiguous tiles still run and
// about half of the resident CTAs only spin.  Every dependency of a ticket has a LOWER ticket, and lower tickets are
// held by CTAs that are already resident or finished -> no deadlock for any lag.
// MINB = 4: 64 registers (one 8-byte spill) -> 4 CTAs per SM (default); MINB = 3: the 80 registers ptxas takes
// unconstrained -> 3 CTAs per SM (QSFT_K3_CTAS=3, kept for the A/B measurement)
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
k3_q4_twopass_kernel(float2* __restrict__ x, long long B, int r1, int r2, long long qa2, int lgW2, int tiles1, int tiles2,
                     unsigned int* __restrict__ done /* [nblocks] counters + [1] ticket */, long long nblocks, int lag,
                     float scale, K3Peers peers) {
    __shared__ float2 s[4096];
    __shared__ unsigned int s_ticket;
    // tickets, not blockIdx: every lower ticket is then guaranteed to be held by a CTA that is already resident (or has
    // finished), which is what makes the wait below deadlock free
    if (threadIdx.x == 0) s_ticket = atomicAdd(done + nblocks, 1u);
    __syncthreads();
    const K3Ticket tk = k3_ticket_decode(s_ticket, nblocks, tiles1, tiles2, lag);
    const long long blk = tk.blk;
    const int t = tk.strided ? tk.t + tiles1 : tk.t;
    float2* base = x + blk * B;
    if (t < tiles1) {
        K3Peers none;
        none.n = 0;
    none.mc = nullptr;
    none.per = 0;
    none.B = 0;
    none.lgB = none.lgper = -1;
        k3_q4_tile<false>(s, base, t, r1, 0, 1, 0, 1.0f, x, none);
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(done + blk, 1u);
        }
    } else {
        if (threadIdx.x == 0) {
            unsigned int seen;
            do {
#ifdef QSFT_EMU   // CPU execution by tests/emu runs the CTAs one after the other in ticket order: an unfinished dependency
                  // could never complete, so it is reported instead of awaited (test infrastructure; never in the product build)
                seen = *reinterpret_cast<volatile unsigned int*>(done + blk);
                if (seen < (unsigned)tiles1) {
                    emu_fail("k3 two-pass: a strided tile was scheduled before its block's contiguous tiles finished");
                    seen = (unsigned)tiles1;
                }
#else
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(done + blk) : "memory");
                if (seen < (unsigned)tiles1) __nanosleep(64);
#endif
            } while (seen < (unsigned)tiles1);
        }
        __syncthreads();
        k3_q4_tile<true>(s, base, t - tiles1, r2, lgW2 / 2, qa2, lgW2, scale, x, peers);
    }
}

struct PassPlan {
    int a, r, W, lgW, outer, rows, T;
    long long qa, tiles_per_block;
    float scale;
};

PassPlan plan_pass(long long B, int q, int a, int r, float scale) {
    PassPlan p{};
    p.a = a; p.r = r; p.scale = scale;
    p.qa = ipow64(q, a);
    p.rows = (int)ipow64(q, r);
    p.W = 1; p.outer = 1; p.lgW = 0;
    if (a == 0) {
        // as many whole q^r blocks as fit (and exist) in one tile
        long long fit = K3_TILE / p.rows, have = B / p.rows, o = 1;
        while (o * q <= fit && (have % (o * q)) == 0) o *= q;
        p.outer = (int)o;
    } else {
        int w = 0;
        while (w < a && (long long)p.W * q * p.rows <= K3_TILE) {
            p.W *= q;
            ++w;
        }
        while ((1 << p.lgW) < p.W) ++p.lgW;
    }
    p.T = p.outer * p.rows * p.W;
    p.tiles_per_block = B / p.T;
    return p;
}

template <int Q>
int launch_pass(float2* x, long long B, int q, const PassPlan& p, long long blk0, long long nblk, cudaStream_t st) {
    const long long tiles = p.tiles_per_block * nblk;
    QSFT_CHECK_ARG(tiles <= 0x7fffffffLL, "too many tiles");
    // vector path: q = 2 / 4, tile and run lengths even (always true for W >= 2 or contiguous tiles of >= 2 elements)
    const bool pow2 = (Q == 2 || Q == 4) && (p.T % 2 == 0) && (p.qa == 1 || p.W >= 2) && (B % 2 == 0);
    const size_t smem = (size_t)p.T * sizeof(float2);
    if (pow2) {
        if constexpr (Q == 2 || Q == 4)
            k3_pass_kernel<Q, true><<<(unsigned)tiles, K3_THREADS, smem, st>>>(x, B, q, p.r, p.qa, p.W, p.lgW, p.outer, p.rows,
                                                                            p.tiles_per_block, blk0 * p.tiles_per_block, p.scale);
    } else {
        k3_pass_kernel<Q, false><<<(unsigned)tiles, K3_THREADS, smem, st>>>(x, B, q, p.r, p.qa, p.W, p.lgW, p.outer, p.rows,
                                                                         p.tiles_per_block, blk0 * p.tiles_per_block, p.scale);
    }
    QSFT_LAUNCHED();
    return QSFT_OK;
}

// peers.n > 0 only for the last pass of a transform and only honoured by the q = 4 kernels (returns 1 in *fused then)
int launch_pass_q(float2* x, long long B, int q, const PassPlan& p, long long blk0, long long nblk, cudaStream_t st,
                  const K3Peers& peers, bool* fused) {
    if (fused) *fused = false;
    if (q == 4 && p.T == 4096 && p.r <= 6) {
        if (fused) *fused = true;
        const long long tiles = p.tiles_per_block * nblk;
        QSFT_CHECK_ARG(tiles <= 0x7fffffffLL, "too many tiles");
        if (p.a == 0)
            k3_q4_fast_kernel<false><<<(unsigned)tiles, 256, 0, st>>>(x, B, p.r, 0, 1, 0, p.tiles_per_block,
                                                                      blk0 * p.tiles_per_block, p.scale, peers);
        else
            k3_q4_fast_kernel<true><<<(unsigned)tiles, 256, 0, st>>>(x, B, p.r, p.lgW / 2, p.qa, p.lgW, p.tiles_per_block,
                                                                     blk0 * p.tiles_per_block, p.scale, peers);
        QSFT_LAUNCHED();
        return QSFT_OK;
    }
    switch (q) {
        case 2: return launch_pass<2>(x, B, q, p, blk0, nblk, st);
        case 3: return launch_pass<3>(x, B, q, p, blk0, nblk, st);
        case 4: return launch_pass<4>(x, B, q, p, blk0, nblk, st);
        case 5: return launch_pass<5>(x, B, q, p, blk0, nblk, st);
        case 7: return launch_pass<7>(x, B, q, p, blk0, nblk, st);
        default: return launch_pass<0>(x, B, q, p, blk0, nblk, st);
    }
}

}  // namespace

// plain copy of the finished transform to the peers, for the paths whose last pass cannot store remotely itself
// (complex64 granularity: rows of odd length q^b are only 8-byte aligned)
__global__ void k3_bcast_copy_kernel(const float2* __restrict__ x, long long n, K3Peers peers) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float2 v = x[i];
        if (peers.per > 0) {
            const long long jj = peers.lgB >= 0 ? (i & (peers.B - 1)) : i % peers.B;
            const int own = (int)(peers.lgper >= 0 ? (jj >> peers.lgper) : jj / peers.per);
            if (peers.p[own] != x) peers.p[own][i] = v;
            continue;
        }
        if (peers.mc != nullptr) {
            asm volatile("multimem.st.weak.global.v2.f32 [%0], {%1, %2};" ::"l"(peers.mc + i), "f"(v.x), "f"(v.y) : "memory");
            continue;
        }
        for (int r = 0; r < peers.n; ++r) peers.p[r][i] = v;
    }
}

// How many blocks the contiguous pass runs ahead of the strided pass (see k3_q4_twopass_kernel): enough tickets between
// a block's two passes to cover every co-resident CTA, but no more intermediate data than stays comfortably in L2.
// QSFT_K3_LAG overrides (0 = the plain block-by-block order).
static int k3_twopass_ctas() { return 4; }                    // CTAs per SM the two-pass kernel is compiled for

static int k3_twopass_lag(long long B, int tiles1, int tiles2) {
    if (const char* e = getenv("QSFT_K3_LAG")) {
        const int v = atoi(e);
        if (v >= 0 && v <= 4096) return v;
    }
    const int minb = k3_twopass_ctas();
    static int resident_by_minb[5] = {0, 0, 0, 0, 0};         // co-resident CTAs of the kernel on this device
    int& resident = resident_by_minb[minb];
    if (resident == 0) {
        int per_sm = 0;
        const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k3_q4_twopass_kernel<4>, 256, 0);
        if (e != cudaSuccess || per_sm < 1) {
            (void)cudaGetLastError();
            per_sm = minb;
        }
        resident = per_sm * qsft_num_sms();
    }
    const long long per = (long long)tiles1 + tiles2;
    long long lag = (resident + per - 1) / per;
    const long long l2_rows = (32ll << 20) / (B * (long long)sizeof(float2));    // intermediates kept live in L2
    if (lag > l2_rows) lag = l2_rows;
    if (lag < 1) lag = 1;
    return (int)lag;
}

static int gwht_impl(float* x, int64_t batch, int q, int b, const K3Peers& peers, void* stream) {
    QSFT_CHECK_ARG(q >= 2 && q <= QSFT_MAX_Q, "q=%d out of range", q);
    QSFT_CHECK_ARG(b >= 0 && b <= QSFT_MAX_B, "b=%d out of range", b);
    QSFT_CHECK_ARG(batch >= 0, "negative batch");
    if (batch == 0 || b == 0) return QSFT_OK;
    QSFT_CHECK_ARG(x != nullptr, "null pointer");
    double Bd = 1;
    for (int i = 0; i < b; ++i) Bd *= q;
    QSFT_CHECK_ARG(Bd <= 4e12, "q^b too large");
    const long long B = ipow64(q, b);
    // q = 4, 4^6 .. 4^10 points: the TMA pipeline of k3_gwht_tma.cu (QSFT_K3_IMPL=1 keeps the register-staged kernels below
    // as a cross-check)
    if (q == 4 && b >= 6 && b <= 10) {
        const char* impl = getenv("QSFT_K3_IMPL");
        if (!(impl && atoi(impl) == 1)) {
            const int rc = qsft_k3_q4_tma(x, batch, b, peers, (cudaStream_t)stream);
            if (rc != QSFT_EUNSUPPORTED) return rc;
        }
    }
    // levels per pass: as many as fit in a tile, spread evenly over the passes
    int cap = 0;
    {
        long long t = 1;
        while (t * q <= K3_TILE) {
            t *= q;
            ++cap;
        }
    }
    QSFT_CHECK_ARG(cap >= 1, "q too large for the tile");
    const int passes = (b + cap - 1) / cap;
    QSFT_CHECK_ARG(passes <= 8, "too many passes");
    const float inv = (float)(1.0 / (double)B);
    cudaStream_t st = (cudaStream_t)stream;
    float2* xx = reinterpret_cast<float2*>(x);
    PassPlan plans[8];
    int a = 0;
    for (int p = 0; p < passes; ++p) {
        // first pass as large as the tile allows (contiguous, cheapest), the remaining levels spread evenly
        const int r = (p == 0) ? ((b < cap) ? b : cap) : (b - a + (passes - p) - 1) / (passes - p);
        plans[p] = plan_pass(B, q, a, r, (p == passes - 1) ? inv : 1.0f);
        a += r;
    }
    if (q == 4 && passes == 2 && plans[0].T == 4096 && plans[1].T == 4096 && plans[0].r <= 6 && plans[1].r <= 6 &&
        plans[0].tiles_per_block + plans[1].tiles_per_block <= (1 << 20) &&
        batch * (plans[0].tiles_per_block + plans[1].tiles_per_block) <= 0x7fffffffLL) {
        unsigned int* done = nullptr;
        QSFT_CUDA(qsft_scratch_alloc((void**)&done, (size_t)(batch + 1) * sizeof(unsigned int), st));
        QSFT_CUDA(cudaMemsetAsync(done, 0, (size_t)(batch + 1) * sizeof(unsigned int), st));
        const int t1 = (int)plans[0].tiles_per_block, t2 = (int)plans[1].tiles_per_block;
        const int lag = k3_twopass_lag(B, t1, t2);
        k3_q4_twopass_kernel<4><<<(unsigned)(batch * (t1 + t2)), 256, 0, st>>>(xx, B, plans[0].r, plans[1].r, plans[1].qa,
                                                                                 plans[1].lgW, t1, t2, done, (long long)batch, lag, inv, peers);
        QSFT_LAUNCHED();
        QSFT_CUDA(cudaFreeAsync(done, st));
        return QSFT_OK;
    }
    // Multi-pass transforms are run chunk by chunk (a few length-B blocks at a time, all passes back to back) so that
    // the intermediate of a chunk is still in the 126 MB L2 when the next pass reads it: DRAM sees ~1 read + 1 write.
    long long chunk = batch;
    if (passes > 1) {
        const long long l2_budget = 40ll << 20;
        chunk = l2_budget / (B * (long long)sizeof(float2));
        if (chunk < 1) chunk = 1;
        if (chunk > batch) chunk = batch;
    }
    K3Peers none;
    none.n = 0;
    none.mc = nullptr;
    none.per = 0;
    none.B = 0;
    none.lgB = none.lgper = -1;
    bool all_fused = true;
    for (long long blk0 = 0; blk0 < batch; blk0 += chunk) {
        const long long nblk = (batch - blk0 < chunk) ? (batch - blk0) : chunk;
        for (int p = 0; p < passes; ++p) {
            bool fused = false;
            const bool lastp = (p == passes - 1);
            if (int rc = launch_pass_q(xx, B, q, plans[p], blk0, nblk, st, lastp ? peers : none, lastp ? &fused : nullptr)) return rc;
            if (lastp && !fused) all_fused = false;
        }
    }
    if (peers.n > 0 && !all_fused) {
        k3_bcast_copy_kernel<<<(unsigned)(8 * qsft_num_sms()), 256, 0, st>>>(xx, batch * B, peers);
        QSFT_LAUNCHED();
    }
    return QSFT_OK;
}

extern "C" int qsft_gwht_batch(float* x, int64_t batch, int q, int b, void* stream) {
    K3Peers none;
    none.n = 0;
    none.mc = nullptr;
    none.per = 0;
    none.B = 0;
    none.lgB = none.lgper = -1;
    return gwht_impl(x, batch, q, b, none, stream);
}

extern "C" int qsft_gwht_batch_mcast(float* x, int64_t batch, int q, int b, float* mc_x, void* stream) {
    QSFT_CHECK_ARG(mc_x != nullptr && ((uintptr_t)mc_x & 7) == 0, "bad multicast pointer");
    K3Peers peers;
    peers.n = 1;                             // "has peers": selects the kernels' peer-store instantiation
    for (int r = 0; r < 8; ++r) peers.p[r] = nullptr;
    peers.per = 0;
    peers.B = 0;
    peers.lgB = peers.lgper = -1;
    peers.mc = reinterpret_cast<float2*>(mc_x);
    return gwht_impl(x, batch, q, b, peers, stream);
}

extern "C" int qsft_gwht_batch_scatter(float* x, int64_t batch, int q, int b, float* const* rank_x, int world, int rank,
                                       int64_t per_bins, void* stream) {
    QSFT_CHECK_ARG(world >= 2 && world <= 8 && rank >= 0 && rank < world && rank_x != nullptr, "bad rank / world");
    QSFT_CHECK_ARG(per_bins > 0 && per_bins * world >= ipow64(q, b), "per_bins * world must cover q^b");
    K3Peers peers;
    peers.n = 1;                             // "has peers": selects the kernels' peer-store instantiation
    peers.mc = nullptr;
    peers.B = ipow64(q, b);
    peers.per = per_bins;
    auto lg2 = [](long long v) { int l = 0; while ((1ll << l) < v) ++l; return (1ll << l) == v ? l : -1; };
    peers.lgB = lg2(peers.B);
    peers.lgper = lg2(per_bins);
    for (int r = 0; r < 8; ++r) peers.p[r] = r < world ? reinterpret_cast<float2*>(rank_x[r]) : nullptr;
    for (int r = 0; r < world; ++r) QSFT_CHECK_ARG(rank_x[r] != nullptr, "null buffer pointer");
    QSFT_CHECK_ARG(rank_x[rank] == x, "rank_x[rank] must be x itself");
    return gwht_impl(x, batch, q, b, peers, stream);
}

extern "C" int qsft_gwht_batch_bcast(float* x, int64_t batch, int q, int b, float* const* peer_x, int n_peers, void* stream) {
    QSFT_CHECK_ARG(n_peers >= 0 && n_peers <= 7, "n_peers must be in [0, 7]");
    QSFT_CHECK_ARG(n_peers == 0 || peer_x != nullptr, "null peer list");
    K3Peers peers;
    peers.mc = nullptr;
    peers.per = 0;
    peers.B = 0;
    peers.lgB = peers.lgper = -1;
    peers.p[7] = nullptr;
    peers.n = n_peers;
    for (int r = 0; r < 7; ++r) peers.p[r] = r < n_peers ? reinterpret_cast<float2*>(peer_x[r]) : nullptr;
    for (int r = 0; r < n_peers; ++r) QSFT_CHECK_ARG(peer_x[r] != nullptr, "null peer pointer");
    return gwht_impl(x, batch, q, b, peers, stream);
}

// Host-side view of the two-pass kernel's ticket order (verification helper: tests prove on the CPU that the order is a
// bijection onto the tiles and that every strided tile's dependencies have lower tickets, i.e. the kernel cannot deadlock).
extern "C" int qsft_k3_ticket_decode(uint32_t ticket, int64_t nblocks, int tiles1, int tiles2, int lag, int64_t* blk,
                                     int* tile, int* strided) {
    QSFT_CHECK_ARG(nblocks >= 1 && tiles1 >= 1 && tiles2 >= 1 && lag >= 0, "bad shape");
    QSFT_CHECK_ARG((long long)ticket < nblocks * ((long long)tiles1 + tiles2), "ticket out of range");
    QSFT_CHECK_ARG(blk && tile && strided, "null pointer");
    const K3Ticket t = k3_ticket_decode(ticket, nblocks, tiles1, tiles2, lag);
    *blk = t.blk;
    *tile = t.t;
    *strided = t.strided ? 1 : 0;
    return QSFT_OK;
}
