// K3 for q = 4, 4^6 .. 4^10 points: the batched base-4 DFT (qsft/utils.py:31-36 as called from
// qsft/input_signal_subsampled.py:264-266) as a persistent, warp-specialised TMA pipeline.
//
// The transform needs two passes over the data (a 4^10 row is 8 MB): levels 0..5 on contiguous 4096-element tiles, the
// remaining r2 = b - 6 levels on tiles of 4^r2 rows (stride 4096 elements) x W = 4^(6 - r2) contiguous elements.  Both
// passes run in ONE launch; the tiles are handed out through an atomic ticket whose order keeps the contiguous pass
// `lag` blocks ahead of the strided pass (k3_ticket_decode, k3_shared.cuh), so the intermediate is consumed from L2 and
// DRAM sees one read and one write of the data.
//
// CTA = 1 producer warp + 8 consumer warps, two CTAs per SM, ring of 2 x 32 KB tiles per CTA:
//   producer   fetches a ticket, waits for the tile's dependencies (strided tiles: the per-block counter of finished
//              contiguous tiles), and issues ONE cp.async.bulk.tensor per tile.  Both tile shapes are boxes of the same 3-D
//              view {16 elements = 128 B, 256 chunks, batch * B / 4096 runs} of the buffer and land in shared memory as
//              256 rows x 128 B with the 128-byte TMA swizzle, element e = t * W + w at row e >> 4.
//   consumers  radix-16 butterflies (two base-4 levels per step) in registers, in place in the tile: the swizzle makes the
//              16-byte accesses of the (0,1) step and the 8-byte accesses of all higher steps bank-conflict free; the last
//              step stores to global memory directly (16 lanes = one 128-byte line), so the stage is released as soon as
//              its last shared-memory read is done and the loads of the next tiles overlap the butterflies.
#include "common.cuh"

#include <stdlib.h>

#include "k3_shared.cuh"
#include "tma.cuh"

namespace {

constexpr int KT_MAX_STAGES = 6;                    // ring depth is a launch parameter (2 .. 6 tiles of 32 KB per CTA)
constexpr int KT_TILE = 4096;                       // complex elements per tile
constexpr int KT_TILE_BYTES = KT_TILE * 8;
constexpr int KT_CONSUMERS = 256;
constexpr int KT_THREADS = KT_CONSUMERS + 64;         // + producer warp + publisher warp
__host__ __device__ constexpr size_t kt_smem(int stages) { return 1024 + (size_t)stages * KT_TILE_BYTES + 256; }

struct KtInfo {
    long long blk;                                   // -1: no more work
    int t;
    int strided;
};

// element index -> position in the 128-byte-swizzled tile (16-byte chunk index ^= row & 7)
__device__ __forceinline__ int kt_swz(int e) { return e ^ (((e >> 4) & 7) << 1); }

// radix-4 butterfly on packed fp32 pairs (FADD2 / FFMA2: one instruction per complex add; a - b = fma(b, -1, a) and the
// +-i rotations = fma(swap(b), (+-1, -+1), a) are exactly the scalar adds / subtracts, so results are bit-identical to the
// register-staged kernels of k3_gwht.cu)
__device__ __forceinline__ void kt_r4(float2& a, float2& b, float2& c, float2& d) {
    const float2 m1 = make_float2(-1.f, -1.f);
    const float2 s02 = __fadd2_rn(a, c), d02 = __ffma2_rn(c, m1, a);
    const float2 s13 = __fadd2_rn(b, d), d13 = __ffma2_rn(d, m1, b);
    const float2 sw = make_float2(d13.y, d13.x);
    a = __fadd2_rn(s02, s13);
    c = __ffma2_rn(s13, m1, s02);
    b = __ffma2_rn(sw, make_float2(1.f, -1.f), d02);   // d02 - i d13
    d = __ffma2_rn(sw, make_float2(-1.f, 1.f), d02);   // d02 + i d13
}

__device__ __forceinline__ void kt_r16(float2 (&v)[16]) {
#pragma unroll
    for (int g = 0; g < 4; ++g) kt_r4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
#pragma unroll
    for (int h = 0; h < 4; ++h) kt_r4(v[h], v[h + 4], v[h + 8], v[h + 12]);
}

// A contiguous-pass tile is "published" -- its block's counter of finished tiles bumped after a gpu-scope fence -- so that
// the strided pass may load it.  The fence waits for the stores just issued (a microsecond): done by the consumer warps it
// cost them a fifth of their time.  A dedicated PUBLISHER warp does it instead: every consumer warp arrives on the stage's
// `stored` mbarrier once its stores of the tile are issued (release at CTA scope, no waiting), the publisher waits for the
// eight arrivals, fences and bumps the counter.
template <bool PEERS>
__device__ __forceinline__ void kt_store(float2* dst, float2 v, const float2* xroot, const K3Peers& peers) {
    if (PEERS) k3_store(dst, v, xroot, peers);
    else *dst = v;
}

__device__ __forceinline__ void kt_bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(KT_CONSUMERS) : "memory"); }

// One tile: levels at base-4 digit positions [P0, P0 + R) of the tile index e.  STRIDED: the tile is 4^R rows (stride 4096
// elements) x W = 4^P0 contiguous elements, e = t * W + w; otherwise 4096 contiguous elements.
template <bool STRIDED, int R, int P0, bool PEERS>
__device__ __forceinline__ void kt_tile(float2* __restrict__ s, float2* __restrict__ gbase, float scale, const float2* xroot,
                                        const K3Peers& peers, uint64_t* empty_bar, uint64_t* stored_bar) {
    const int tau = threadIdx.x;                     // consumer threads are 0 .. 255
    constexpr int lgW = STRIDED ? 2 * P0 : 0;
    constexpr int W = 1 << lgW;
    constexpr int nsteps = (R + 1) >> 1;
#pragma unroll
    for (int st = 0; st < nsteps; ++st) {
        const int p = P0 + 2 * st;
        const bool pair = (2 * st + 1) < R;
        const bool last = (st == nsteps - 1);
        const int sh = 2 * p, lowmask = (1 << sh) - 1, step = 1 << sh;
        // global address of element e0 + m * step (last step only): base + m * gstep
        const long long gstep = STRIDED ? ((long long)(step >> lgW) << 12) : (long long)step;
        float2 v[16];
        if (pair) {
            const int eb = ((tau >> sh) << (sh + 4)) | (tau & lowmask);
            if (p == 0) {
                const float4* s4 = reinterpret_cast<const float4*>(s) + tau * 8;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 t4 = s4[c ^ (tau & 7)];
                    v[2 * c] = make_float2(t4.x, t4.y);
                    v[2 * c + 1] = make_float2(t4.z, t4.w);
                }
            } else {
#pragma unroll
                for (int m = 0; m < 16; ++m) v[m] = s[kt_swz(eb + m * step)];
            }
            kt_r16(v);
            if (last) {
                __syncwarp();
                if ((tau & 31) == 0) tma::mbar_arrive(empty_bar);     // this warp no longer reads the stage
                float2* g = STRIDED ? gbase + (long long)(eb >> lgW) * KT_TILE + (eb & (W - 1)) : gbase + eb;
#pragma unroll
                for (int m = 0; m < 16; ++m) kt_store<PEERS>(g + m * gstep, make_float2(v[m].x * scale, v[m].y * scale), xroot, peers);
            } else if (p == 0) {
                float4* s4 = reinterpret_cast<float4*>(s) + tau * 8;
#pragma unroll
                for (int c = 0; c < 8; ++c) s4[c ^ (tau & 7)] = make_float4(v[2 * c].x, v[2 * c].y, v[2 * c + 1].x, v[2 * c + 1].y);
            } else {
#pragma unroll
                for (int m = 0; m < 16; ++m) s[kt_swz(eb + m * step)] = v[m];
            }
        } else {
            // single level: four radix-4 butterflies per thread
            int eb[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int beta = tau + 256 * k;
                eb[k] = ((beta >> sh) << (sh + 2)) | (beta & lowmask);
#pragma unroll
                for (int m = 0; m < 4; ++m) v[4 * k + m] = s[kt_swz(eb[k] + m * step)];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) kt_r4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            if (last) {
                __syncwarp();
                if ((tau & 31) == 0) tma::mbar_arrive(empty_bar);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float2* g = STRIDED ? gbase + (long long)(eb[k] >> lgW) * KT_TILE + (eb[k] & (W - 1)) : gbase + eb[k];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    if (last) kt_store<PEERS>(g + m * gstep, make_float2(v[4 * k + m].x * scale, v[4 * k + m].y * scale), xroot, peers);
                    else s[kt_swz(eb[k] + m * step)] = v[4 * k + m];
                }
            }
        }
        if (!last) kt_bar_consumers();
    }
    __syncwarp();
    if ((tau & 31) == 0) tma::mbar_arrive(stored_bar);                    // this warp's stores of the tile are issued
}

template <bool PEERS>
__global__ void __launch_bounds__(KT_THREADS, 2)
k3_q4_tma_kernel(const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2, float2* __restrict__ x,
                 long long B, int r1, int r2, int tiles1, int tiles2, unsigned int* __restrict__ done /* [nblocks] + ticket */,
                 long long nblocks, int lag, int nstages, int nofence, float scale1, float scale2, K3Peers peers) {
    extern __shared__ __align__(1024) uint8_t kt_raw[];
    uint8_t* base = kt_raw + ((1024u - (tma::smem_u32(kt_raw) & 1023u)) & 1023u);      // 1024-byte aligned (TMA swizzle atom)
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)nstages * KT_TILE_BYTES);
    uint64_t* empty = full + KT_MAX_STAGES;
    uint64_t* stored = empty + KT_MAX_STAGES;
    KtInfo* info = reinterpret_cast<KtInfo*>(stored + KT_MAX_STAGES);
    const int warp = threadIdx.x >> 5;
    const int rows = (int)(B / KT_TILE);                                     // 4096-element runs per block = 4^r2
    const int lgW = 2 * (6 - r2);
    const long long total = nblocks * ((long long)tiles1 + tiles2);

    if (threadIdx.x == 0) {
        for (int i = 0; i < nstages; ++i) {
            tma::mbar_init(&full[i], 1);
            tma::mbar_init(&empty[i], KT_CONSUMERS / 32 + 1);               // the consumer warps + the publisher
            tma::mbar_init(&stored[i], KT_CONSUMERS / 32);
        }
        tma::mbar_fence_init();
    }
    __syncthreads();

    if (warp == KT_CONSUMERS / 32 + 1) {
        // ---- publisher ------------------------------------------------------------------------------------------
        if ((threadIdx.x & 31) == 0) {
            for (unsigned it = 0;; ++it) {
                const int stage = (int)(it % (unsigned)nstages);
                const uint32_t ph = (it / (unsigned)nstages) & 1u;
                tma::mbar_wait(&full[stage], ph);
                const long long blk = info[stage].blk;
                const bool contiguous = info[stage].strided == 0;
                if (blk < 0) break;
                tma::mbar_arrive(&empty[stage]);                       // the stage's description has been read
                tma::mbar_wait(&stored[stage], ph);
                if (contiguous && tiles2 != 0) {
                    if (!nofence) __threadfence();
                    atomicAdd(done + blk, (unsigned int)(KT_CONSUMERS / 32));
                }
            }
        }
    } else if (warp == KT_CONSUMERS / 32) {
        // ---- producer --------------------------------------------------------------------------------------------
        if ((threadIdx.x & 31) == 0) {
            tma::prefetch_map(&tm1);
            tma::prefetch_map(&tm2);
            // tickets, not blockIdx: every lower ticket is held by a CTA that is already resident (or finished), which is
            // what makes the dependency wait below deadlock free.  The ticket of the NEXT tile is drawn before this one's
            // stage is waited for, so the round trip of the atomic overlaps the wait.
            unsigned int ticket_next = atomicAdd(done + nblocks, 1u);
            for (unsigned it = 0;; ++it) {
                const int stage = (int)(it % (unsigned)nstages);
                const uint32_t ph = (it / (unsigned)nstages) & 1u;
                const unsigned int ticket = ticket_next;
                if ((long long)ticket < total) ticket_next = atomicAdd(done + nblocks, 1u);
                tma::mbar_wait(&empty[stage], ph ^ 1u);
                if ((long long)ticket >= total) {
                    info[stage].blk = -1;
                    tma::mbar_arrive(&full[stage]);
                    break;
                }
                K3Ticket tk;
                if (tiles2 == 0) {
                    tk.blk = ticket / (unsigned int)tiles1;
                    tk.t = (int)(ticket - (unsigned int)(tk.blk * tiles1));
                    tk.strided = false;
                } else {
                    tk = k3_ticket_decode(ticket, nblocks, tiles1, tiles2, lag);
                }
                info[stage].blk = tk.blk;
                info[stage].t = tk.t;
                info[stage].strided = tk.strided ? 1 : 0;
                uint8_t* dst = base + (size_t)stage * KT_TILE_BYTES;
                if (!tk.strided) {
                    tma::mbar_expect_tx(&full[stage], KT_TILE_BYTES);
                    tma::load_3d(dst, &tm1, 0, 0, (int)(tk.blk * rows + tk.t), &full[stage]);
                } else {
                    const unsigned int need = (unsigned int)tiles1 * (KT_CONSUMERS / 32);
                    unsigned int seen;
                    for (;;) {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(done + tk.blk) : "memory");
                        if (seen >= need) break;
                        __nanosleep(64);
                    }
                    tma::fence_proxy_async_global();
                    tma::mbar_expect_tx(&full[stage], KT_TILE_BYTES);
                    tma::load_3d(dst, &tm2, 0, tk.t << (lgW > 4 ? lgW - 4 : 0), (int)(tk.blk * rows), &full[stage]);
                }
            }
        }
    } else {
        // ---- consumers -------------------------------------------------------------------------------------------
        for (unsigned it = 0;; ++it) {
            const int stage = (int)(it % (unsigned)nstages);
            const uint32_t ph = (it / (unsigned)nstages) & 1u;
            tma::mbar_wait(&full[stage], ph);
            const long long blk = info[stage].blk;
            if (blk < 0) break;
            const int t = info[stage].t;
            const bool strided = info[stage].strided != 0;
            float2* s = reinterpret_cast<float2*>(base + (size_t)stage * KT_TILE_BYTES);
            float2* xb = x + blk * B;
            if (!strided) {
                if (tiles2 == 0)
                    kt_tile<false, 6, 0, PEERS>(s, xb + (long long)t * KT_TILE, scale1, x, peers, &empty[stage], &stored[stage]);
                else
                    kt_tile<false, 6, 0, false>(s, xb + (long long)t * KT_TILE, scale1, x, peers, &empty[stage], &stored[stage]);
            } else {
                float2* gb = xb + ((long long)t << lgW);
                switch (r2) {
                    case 4: kt_tile<true, 4, 2, PEERS>(s, gb, scale2, x, peers, &empty[stage], &stored[stage]); break;
                    case 3: kt_tile<true, 3, 3, PEERS>(s, gb, scale2, x, peers, &empty[stage], &stored[stage]); break;
                    case 2: kt_tile<true, 2, 4, PEERS>(s, gb, scale2, x, peers, &empty[stage], &stored[stage]); break;
                    default: kt_tile<true, 1, 5, PEERS>(s, gb, scale2, x, peers, &empty[stage], &stored[stage]); break;
                }
            }
        }
    }
}

}  // namespace

// x (batch, 4^b) complex64 in place, 6 <= b <= 10, batch * 4^b / 4096 < 2^31.  Returns QSFT_EUNSUPPORTED when the shape is
// outside that range (the caller falls back to the register-staged kernels of k3_gwht.cu).
int qsft_k3_q4_tma(float* xf, int64_t batch, int b, const K3Peers& peers_in, cudaStream_t st) {
    const int n_peers = peers_in.n;
    if (b < 6 || b > 10) return QSFT_EUNSUPPORTED;
    const long long B = ipow64(4, b);
    const long long runs = batch * (B / KT_TILE);
    if (runs >= 0x7fffffffLL || ((uintptr_t)xf & 15) != 0) return QSFT_EUNSUPPORTED;
    const int r1 = 6, r2 = b - 6;
    const int tiles1 = (int)(B / KT_TILE);
    const int tiles2 = r2 ? (int)(KT_TILE >> (2 * (6 - r2))) : 0;             // 4096 / W
    const int W16 = r2 ? ((1 << (2 * (6 - r2))) / 16) : 1;                    // 16-element chunks per row of a strided tile
    if (r2 && (1 << (2 * (6 - r2))) < 16) return QSFT_EUNSUPPORTED;
    CUtensorMap tm1, tm2;
    // 3-D view: 32 floats (16 elements, 128 B) x 256 chunks (one 4096-element run) x runs
    const cuuint64_t dims[3] = {32, 256, (cuuint64_t)runs};
    const cuuint64_t strides[2] = {128, (cuuint64_t)KT_TILE_BYTES};
    const cuuint32_t box1[3] = {32, 256, 1};
    if (int rc = tma::make_map(&tm1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, xf, dims, strides, box1, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    const cuuint32_t box2[3] = {32, (cuuint32_t)W16, (cuuint32_t)(r2 ? (256 / W16) : 1)};
    if (int rc = tma::make_map(&tm2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, xf, dims, strides, r2 ? box2 : box1, CU_TENSOR_MAP_SWIZZLE_128B))
        return rc;
    // ring depth (QSFT_K3_STAGES = 2 .. 6 for measurements), measured per shape with the publisher warp in place (r3g):
    //   4^6: 0.052 / 0.054 ms   4^7: 0.086 / 0.080   4^8: 0.169 / 0.151   4^9: 0.173 / 0.165   4^10: 0.182 / 0.190   (2 / 3 tiles)
    // Three CTAs per SM (64 registers, no spills) with two-tile rings, measured on one box (r4b): 4^7 0.078, 4^8 0.151, 4^9 0.168
    // -- what the third stage already gives -- and 4^10 0.200 against 0.177 ms: more tiles in flight do not help the two-pass
    // shape, whose tile rate is set by the shared-memory traffic of the butterfly steps (160 KB per 32 KB tile).  Not kept.
    int nstages = (b >= 7 && b <= 9) ? 3 : 2, nofence = 0;
    if (const char* e = getenv("QSFT_K3_STAGES"))
        if (atoi(e) >= 2 && atoi(e) <= KT_MAX_STAGES) nstages = atoi(e);
    if (const char* e = getenv("QSFT_K3_NOFENCE")) nofence = atoi(e) != 0;
    static int ctas_by_stages[KT_MAX_STAGES + 1] = {0};
    int& ctas = ctas_by_stages[nstages];
    const size_t smem = kt_smem(nstages);
    if (ctas == 0) {
        QSFT_CUDA(cudaFuncSetAttribute(k3_q4_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kt_smem(KT_MAX_STAGES)));
        QSFT_CUDA(cudaFuncSetAttribute(k3_q4_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kt_smem(KT_MAX_STAGES)));
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k3_q4_tma_kernel<true>, KT_THREADS, smem) != cudaSuccess || per_sm < 1) {
            (void)cudaGetLastError();
            per_sm = 1;
        }
        ctas = per_sm * qsft_num_sms();
    }
    const long long total = batch * ((long long)tiles1 + tiles2);
    const int grid = (int)(total < ctas ? total : ctas);
    // contiguous pass `lag` blocks ahead: enough tickets between a block's two passes to cover every tile in flight, but no
    // more intermediate data than stays comfortably in L2
    int lag = 1;
    if (tiles2) {
        const long long per = (long long)tiles1 + tiles2;
        long long l = ((long long)ctas * nstages + per - 1) / per;
        const long long l2_rows = (48ll << 20) / (B * 8);
        if (l > l2_rows) l = l2_rows;
        if (l < 1) l = 1;
        // 4^10 (two tiles per CTA in flight cover 1.16 blocks): one more block of lead measured 0.176 against 0.181 ms (r4n: lag
        // 1 / 2 / 3 / 4 / 6 = 0.212 / 0.181 / 0.176 / 0.190 / 0.208 ms for 41 rows; the producer's poll interval made no difference)
        if (b == 10 && l == 2) l = 3;
        lag = (int)l;
    }
    unsigned int* done = nullptr;
    QSFT_CUDA(qsft_scratch_alloc((void**)&done, (size_t)(batch + 1) * sizeof(unsigned int), st));
    QSFT_CUDA(cudaMemsetAsync(done, 0, (size_t)(batch + 1) * sizeof(unsigned int), st));
    const K3Peers peers = peers_in;
    const float inv = (float)(1.0 / (double)B);
    if (n_peers > 0)
        k3_q4_tma_kernel<true><<<grid, KT_THREADS, smem, st>>>(tm1, tm2, reinterpret_cast<float2*>(xf), B, r1, r2, tiles1, tiles2, done,
                                                               (long long)batch, lag, nstages, nofence, r2 ? 1.0f : inv, inv, peers);
    else
        k3_q4_tma_kernel<false><<<grid, KT_THREADS, smem, st>>>(tm1, tm2, reinterpret_cast<float2*>(xf), B, r1, r2, tiles1, tiles2, done,
                                                                (long long)batch, lag, nstages, nofence, r2 ? 1.0f : inv, inv, peers);
    QSFT_LAUNCHED();
    QSFT_CUDA(cudaFreeAsync(done, st));
    return QSFT_OK;
}
