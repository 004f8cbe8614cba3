// K2 (tensor-core variant): synthetic sparse signal evaluation as a tcgen05 int8 GEMM with a fused epilogue.
//   D[m, s] = <qdig[m], loc[s]>            (UTCIMMA, kind::i8, int32 accumulators in TMEM, K = ld)
//   out[m]  = sum_s a_s * w^(D[m, s] mod q)  (epilogue: tcgen05.ld -> mod q -> root-of-unity table in smem -> complex add)
// q = 4 fast path: the support digits are pre-scaled by 8 (values 0..24), so (D & 24) is directly the byte offset of
// the rotated strength a_s * i^t inside a 32-byte table entry: LOP3 + LDS.64 + 2 FADD per (query, support) pair.
// Replaces synt_exp/synt_src/synthetic_signal.py:100-118 (exp(Q @ 2 pi i locq / q) @ strengths).
//
// One CTA owns 128 query rows (UMMA M = 128, cta_group::1) and streams the support in tiles of 256 rows (UMMA N = 256)
// through a TMA / mbarrier ring; two 256-column accumulators ping-pong in TMEM so the MMA of tile i+1 overlaps the
// epilogue of tile i.  Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc), 2..5 = epilogue (one TMEM lane
// quarter each).  Operand tiles are K-major rows of `LD` bytes with the LD-byte TMA/UMMA swizzle.
#include <cuda.h>

#include <utility>

#include "common.cuh"

namespace {

constexpr int TC_BM = 128;          // query rows per CTA
constexpr int TC_BN = 256;          // support rows per tile
constexpr int TC_EPI_WARPS = 8;     // two warps per TMEM lane quarter, each reduces half of the tile's columns
constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32;
constexpr int TC_THREADS = 64 + TC_EPI_THREADS;
constexpr int TC_COLS_PER_WARP = TC_BN / (TC_EPI_WARPS / 4);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive columns of 32-bit accumulators -> 32 registers per thread
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// q = 4 epilogue step for column (CH*32 + J) of this warp's column range: the table entry of a column is 32 bytes
// (a, ia, -a, -ia); `v & 24` selects the rotation, `base` (32-byte aligned, so OR == ADD) is the table address of the
// warp's first column and the column offset is an immediate of the LDS: LOP3 + LDS.64 + 2 FADD per pair.
template <int OFF>
__device__ __forceinline__ float2 lds_f2_off(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(addr), "n"(OFF) : "memory");
    return v;
}
template <int CH, int J>
__device__ __forceinline__ void q4_step(uint32_t v, uint32_t base, float (&pr)[4], float (&pi)[4]) {
    const float2 w = lds_f2_off<(CH * 32 + J) * 32>((v & 24u) | base);
    pr[J & 3] += w.x;
    pi[J & 3] += w.y;
}
template <int CH, int... J>
__device__ __forceinline__ void q4_chunk(const uint32_t (&r)[32], uint32_t base, float (&pr)[4], float (&pi)[4],
                                         std::integer_sequence<int, J...>) {
    (q4_step<CH, J>(r[J], base, pr, pi), ...);
}

// K-major operand tile, rows of LD bytes, LD-byte swizzle: SBO = 8 rows * LD bytes, LBO unused (=1), version 1
template <int LD>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    constexpr uint64_t layout = (LD == 128) ? 2ull : (LD == 64) ? 4ull : 6ull;
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * LD) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

template <int LD>
struct TcSmem {
    static constexpr int kStages = (LD == 128) ? 4 : (LD == 64) ? 6 : 12;   // >= 114 KB smem: one CTA (one 512-column TMEM allocation) per SM
    static constexpr int kABytes = TC_BM * LD;
    static constexpr int kBBytes = TC_BN * LD;
    static constexpr size_t kBytes = 1024 /*align slack*/ + kABytes + (size_t)kStages * kBBytes +
                                     (QSFT_MAX_Q + 1) * 8 + 32 * 8 + 16 + TC_BM * 16;
};

template <int LD, bool Q4>
__global__ void __launch_bounds__(TC_THREADS, 1)
k2_eval_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const float2* __restrict__ strengths, long long N, long long S, int q, uint32_t qmagic,
                  float2* __restrict__ out) {
    using L = TcSmem<LD>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = base;
    uint8_t* sB = sA + L::kABytes;
    // strength tables for the two accumulators: static so that their shared address is a link-time constant
    // (the epilogue's LDS then needs no address arithmetic beyond one LOP3)
    __shared__ __align__(1024) float2 sTab[2 * TC_BN * 4];
    float2* sTw = reinterpret_cast<float2*>(sB + (size_t)L::kStages * L::kBBytes);
    double2* sRed = reinterpret_cast<double2*>(sTw + QSFT_MAX_Q + 1);           // [TC_BM] partial sums of column half 1
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + TC_BM);
    uint64_t* full = bars;                       // [kStages]
    uint64_t* empty = bars + L::kStages;         // [kStages]
    uint64_t* tfull = bars + 2 * L::kStages;     // [2]
    uint64_t* tempty = tfull + 2;                // [2]
    uint64_t* afull = tempty + 2;                // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m0 = (long long)blockIdx.x * TC_BM;
    const int ntiles = (int)((S + TC_BN - 1) / TC_BN);

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < L::kStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], TC_EPI_WARPS);
        }
        mbar_init(afull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (!Q4) {
        for (int t = threadIdx.x; t < q; t += TC_THREADS) {
            float sn, cs;
            sincospif(2.0f * (float)t / (float)q, &sn, &cs);
            sTw[t] = make_float2(cs, sn);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(afull, L::kABytes);
            tma_load_2d(sA, &tmA, 0, (int)m0, afull);
            for (int it = 0; it < ntiles; ++it) {
                const int stage = it % L::kStages;
                const uint32_t ph = (uint32_t)(it / L::kStages) & 1u;
                mbar_wait(&empty[stage], ph ^ 1u);
                mbar_expect_tx(&full[stage], L::kBBytes);
                tma_load_2d(sB + (size_t)stage * L::kBBytes, &tmB, 0, it * TC_BN, &full[stage]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D = S32, A = B = signed int8, both K-major, N = 256, M = 128
            constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                                       ((uint32_t)(TC_BM >> 4) << 24);
            const uint64_t adesc0 = make_desc<LD>(smem_u32(sA));
            mbar_wait(afull, 0);
            for (int it = 0; it < ntiles; ++it) {
                const int stage = it % L::kStages;
                const uint32_t ph = (uint32_t)(it / L::kStages) & 1u;
                const int acc = it & 1;
                const uint32_t aph = (uint32_t)(it >> 1) & 1u;
                mbar_wait(&tempty[acc], aph ^ 1u);
                mbar_wait(&full[stage], ph);
                tc_fence_after();
                const uint64_t bdesc0 = make_desc<LD>(smem_u32(sB + (size_t)stage * L::kBBytes));
#pragma unroll
                for (int k = 0; k < LD / 32; ++k) {
                    // advance 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
                    umma_i8(tmem_base + (uint32_t)(acc * TC_BN), adesc0 + (uint64_t)(2 * k), bdesc0 + (uint64_t)(2 * k), idesc,
                            k > 0 ? 1u : 0u);
                }
                umma_commit(&empty[stage]);   // smem stage may be refilled once these MMAs retire
                umma_commit(&tfull[acc]);     // accumulator ready for the epilogue
            }
        }
    } else {
        const int e = threadIdx.x - 64;                 // 0..TC_EPI_THREADS-1 inside the epilogue group
        const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;               // which column half of the tile this warp reduces
        const int row = quarter * 32 + lane;
        const int c0 = half * TC_COLS_PER_WARP;
        double xr = 0.0, xi = 0.0;
        for (int it = 0; it < ntiles; ++it) {
            const int acc = it & 1;
            const uint32_t aph = (uint32_t)(it >> 1) & 1u;
            // stage the strengths (for q = 4: their four rotations a, ia, -a, -ia) of this tile
            float2* tab = sTab + acc * (TC_BN * 4);
#pragma unroll
            for (int h = 0; h < TC_BN / TC_EPI_THREADS; ++h) {
                const int c = e + h * TC_EPI_THREADS;
                const long long s = (long long)it * TC_BN + c;
                const float2 a = (s < S) ? __ldg(strengths + s) : make_float2(0.f, 0.f);
                if (Q4) {
                    float4* t4 = reinterpret_cast<float4*>(tab + c * 4);
                    t4[0] = make_float4(a.x, a.y, -a.y, a.x);
                    t4[1] = make_float4(-a.x, -a.y, a.y, -a.x);
                } else {
                    tab[c] = a;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
            float pr[4] = {0.f, 0.f, 0.f, 0.f}, pi[4] = {0.f, 0.f, 0.f, 0.f};   // independent chains hide FADD latency
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * TC_BN + c0);
            const uint32_t tabq = smem_u32(sTab) + (uint32_t)(acc * TC_BN + c0) * 32u;   // 32-byte aligned
            const float2* tabg = sTab + acc * (TC_BN * 4) + c0;
            uint32_t r0[32], r1[32];
            static_assert(TC_COLS_PER_WARP == 128, "epilogue below is written for four 32-column chunks per warp");
            auto generic_chunk = [&](const uint32_t (&r)[32], int ch) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const uint32_t v = r[j];
                    const uint32_t t = v - __umulhi(v, qmagic) * (uint32_t)q;
                    const float2 tw = sTw[t];
                    const float2 a = tabg[ch * 32 + j];
                    pr[j & 3] = fmaf(a.x, tw.x, fmaf(-a.y, tw.y, pr[j & 3]));
                    pi[j & 3] = fmaf(a.x, tw.y, fmaf(a.y, tw.x, pi[j & 3]));
                }
            };
            using Seq = std::make_integer_sequence<int, 32>;
            tmem_ld_32x32(taddr, r0);
            tmem_ld_wait();
            tmem_ld_32x32(taddr + 32u, r1);
            if (Q4) q4_chunk<0>(r0, tabq, pr, pi, Seq{}); else generic_chunk(r0, 0);
            tmem_ld_wait();
            tmem_ld_32x32(taddr + 64u, r0);
            if (Q4) q4_chunk<1>(r1, tabq, pr, pi, Seq{}); else generic_chunk(r1, 1);
            tmem_ld_wait();
            tmem_ld_32x32(taddr + 96u, r1);
            if (Q4) q4_chunk<2>(r0, tabq, pr, pi, Seq{}); else generic_chunk(r0, 2);
            tmem_ld_wait();
            if (Q4) q4_chunk<3>(r1, tabq, pr, pi, Seq{}); else generic_chunk(r1, 3);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            xr += (double)((pr[0] + pr[1]) + (pr[2] + pr[3]));
            xi += (double)((pi[0] + pi[1]) + (pi[2] + pi[3]));
        }
        // combine the two column halves of each row
        if (half == 1) sRed[row] = make_double2(xr, xi);
        asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
        if (half == 0 && m0 + row < N) {
            const double2 o = sRed[row];
            out[m0 + row] = make_float2((float)(xr + o.x), (float)(xi + o.y));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// q = 4 fast path: support digit rows scaled by 8 (see the epilogue).  q = 2 takes the same path with the digits scaled by 16:
// (-1)^t = i^(2 t), so (16 t) & 24 = 16 (t & 1) selects a or -a in the same table.
__global__ void scale8_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, long long words, int shift) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < words) out[i] = in[i] << shift;     // bytes are <= 3 (q = 4, shift 3) or <= 1 (q = 2, shift 4): no carry across byte lanes
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D uint8 tensor (rows, ld) -> box (box_rows, ld) with the ld-byte swizzle; OOB rows read as zero
int make_map(CUtensorMap* map, const void* ptr, long long rows, int ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        qsft_set_error("cuTensorMapEncodeTiled entry point not available");
        return QSFT_ECUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld};
    cuuint32_t box[2] = {(cuuint32_t)ld, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = ld == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : ld == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        qsft_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld ld=%d box=%d)", (int)r, rows, ld, box_rows);
        return QSFT_ECUDA;
    }
    return QSFT_OK;
}

template <int LD, bool Q4>
int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const float2* a, long long N, long long S, int q, float2* out,
              cudaStream_t st) {
    const size_t smem = TcSmem<LD>::kBytes;
    static bool attr_set = false;
    if (!attr_set) {
        QSFT_CUDA(cudaFuncSetAttribute(k2_eval_tc_kernel<LD, Q4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const uint32_t qmagic = (uint32_t)(((1ull << 32) + q - 1) / q);
    const long long blocks = (N + TC_BM - 1) / TC_BM;
    k2_eval_tc_kernel<LD, Q4><<<(unsigned)blocks, TC_THREADS, smem, st>>>(ma, mb, a, N, S, q, qmagic, out);
    QSFT_LAUNCHED();
    return QSFT_OK;
}

}  // namespace

bool qsft_eval_synth_tc_supported(int64_t N, int64_t S, int q, int n, int ld) {
    (void)n;
    if (!(ld == 32 || ld == 64 || ld == 128)) return false;
    if (N < 1 || S < 1 || N > 0x7fffffffLL || S > 0x7fffffffLL - TC_BN) return false;
    if (q < 2 || q > QSFT_MAX_Q) return false;
    return true;
}

int qsft_eval_synth_tc(const int8_t* qdig, int64_t N, const int8_t* loc, const float* strengths, int64_t S, int q, int n,
                       int ld, float* out, void* stream) {
    (void)n;
    QSFT_CHECK_ARG(((uintptr_t)qdig & 15) == 0 && ((uintptr_t)loc & 15) == 0, "digit buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const bool q4 = (q == 4 || q == 2);          // quarter-turn phases: the LOP3 + LDS.64 + 2 FADD epilogue
    int8_t* loc8 = nullptr;
    if (q4) {  // stream-ordered scratch copy of the support digits, scaled by 8 (q = 4) / 16 (q = 2)
        const long long words = S * (long long)ld / 4;
        QSFT_CUDA(qsft_scratch_alloc((void**)&loc8, (size_t)S * ld, st));
        scale8_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint32_t*>(loc),
                                                                       reinterpret_cast<uint32_t*>(loc8), words, q == 4 ? 3 : 4);
        QSFT_LAUNCHED();
    }
    CUtensorMap ma, mb;
    int rc = make_map(&ma, qdig, N, ld, TC_BM);
    if (!rc) rc = make_map(&mb, q4 ? loc8 : loc, S, ld, TC_BN);
    const float2* a = reinterpret_cast<const float2*>(strengths);
    float2* o = reinterpret_cast<float2*>(out);
    if (!rc) {
        if (ld == 32) rc = q4 ? launch_tc<32, true>(ma, mb, a, N, S, q, o, st) : launch_tc<32, false>(ma, mb, a, N, S, q, o, st);
        else if (ld == 64) rc = q4 ? launch_tc<64, true>(ma, mb, a, N, S, q, o, st) : launch_tc<64, false>(ma, mb, a, N, S, q, o, st);
        else rc = q4 ? launch_tc<128, true>(ma, mb, a, N, S, q, o, st) : launch_tc<128, false>(ma, mb, a, N, S, q, o, st);
    }
    if (loc8) cudaFreeAsync(loc8, st);
    return rc;
}
