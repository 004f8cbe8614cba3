// K2L: lattice-factorised synthetic evaluation (fused K1 + K2) for q = 4 (and q = 3, see "q = 3" below) -- SURVEY section 7,
// option (b).
//
// For the query lattice m = M l + d_p the phase splits as  <k_s, M l + d_p> = <h_s, l> + <k_s, d_p>,  h_s = M^T k_s mod q.
// With l = (l_hi, l_lo) (b1 + b2 digits) the samples of delay row p are a complex matrix product
//     X_p[l_hi, l_lo] = sum_s  A_p[l_hi, s] * Y[s, l_lo],   A_p = i^(<h_hi(s), l_hi> + e_ps),   Y = a_s i^<h_lo(s), l_lo>
// (the delay phase e_ps = <k_s, d_p> rides on the exact operand, so the 3-limb operand Y is shared by all delay rows)
// with K = S on the tensor cores and NO per-(query, support) epilogue.  A is exact in int8 (entries 0, +-1 after the
// real embedding [[Re, -Im], [Im, Re]]); Y_p is a rotation (sign / swap, exact) of a_s quantised to three balanced
// base-128 int8 limbs (21 bits + sign relative to max|a|), accumulated error-free in int32 (UTCIMMA kind::i8) and
// recombined in the epilogue.  Replaces synt_exp/synt_src/synthetic_signal.py:100-118 evaluated on the indices of
// qsft/input_signal_subsampled.py:183-206, for every delay row of one (M, D) block at once.
//
// GEMM tile: 128 rows of A' (64 l_hi x {re, im}) x 128 columns (l_lo) x 3 limb accumulators (384 TMEM columns);
// K' = 2S bytes streamed in 128-byte slabs through a 3-stage TMA/mbarrier ring (A slab 16 KB + 3 limb slabs 48 KB).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int LT_BM = 128, LT_BN = 128, LT_BK = 128, LT_STAGES = 3, LT_LIMBS = 3, LT_THREADS = 192;
constexpr int LT_STAGE_BYTES = LT_BM * LT_BK + LT_LIMBS * LT_BN * LT_BK;   // 64 KB
constexpr size_t LT_SMEM = 1024 + (size_t)LT_STAGES * LT_STAGE_BYTES + 256;
constexpr int LT_SCALE_BITS = 20;

__device__ __forceinline__ uint32_t lt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lt_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lt_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void lt_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lt_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LT_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra LT_WAIT_DONE;\n\t"
        "bra LT_WAIT_LOOP;\n\t"
        "LT_WAIT_DONE:\n\t"
        "}" ::"r"(lt_smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void lt_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            lt_smem_u32(dst)),
        "l"(map), "r"(lt_smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void lt_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(lt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void lt_umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void lt_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// K-major tile, 128-byte rows, 128-byte swizzle: SBO = 1024 bytes, LBO unused, descriptor version 1
__device__ __forceinline__ uint64_t lt_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// sum_i digit_i(a) * digit_i(b) mod 4 over nd base-4 digits packed two bits each.  With a = 2 a1 + a0, b = 2 b1 + b0 per
// digit, a b = a0 b0 + 2 (a1 b0 + a0 b1) mod 4, so the sum is popc(a0 & b0) + 2 popc((a1 & b0) ^ (a0 & b1)) mod 4: about ten
// integer instructions instead of a loop of seven per digit (checked against the loop for every pair up to nd = 6 and on
// random words; tests/test_emu_k2l_prep.py).  The b side is a block constant in every caller and gets hoisted.
__device__ __forceinline__ uint32_t dot4(uint32_t a, uint32_t b, int nd) {
    const uint32_t M = 0x55555555u & (nd >= 16 ? 0xffffffffu : ((1u << (2 * nd)) - 1u));
    const uint32_t a0 = a & M, a1 = (a >> 1) & M, b0 = b & M, b1 = (b >> 1) & M;
    return ((uint32_t)__popc(a0 & b0) + 2u * (uint32_t)__popc((a1 & b0) ^ (a0 & b1))) & 3u;
}

// q = 2 runs through the same machinery as q = 4: (-1)^t = i^(2 t), so lt_prep stores the bin-hash digits and delay phases
// DOUBLED (0 / 2) in the two-bit fields, and a lattice index l in [0, 2^nd) -- one BIT per digit -- is spread to one two-bit
// field per digit (values 0 / 1) before the dot product: dot4(2 h, spread(l)) = 2 <h, l> mod 4.  Row / column counts are
// 2^b1 / 2^b2 instead of 4^b1 / 4^b2; the GEMM kernels only ever see counts and tables.
__device__ __forceinline__ uint32_t lt_digits(uint32_t idx, int spread) {
    if (!spread) return idx;
    uint32_t x = idx & 0xffffu;
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

// ---- operand generation -----------------------------------------------------------------------------------
// per support element: bin hash halves (h_hi, h_lo) and the delay phases e[p][s] = <d_p, k_s> mod 4
__global__ void lt_prep_kernel(const int8_t* __restrict__ M, const int8_t* __restrict__ D, const int8_t* __restrict__ loc,
                               long long S, long long Se, int n, int b, int b1, int P, int ld, uint32_t* __restrict__ hhi,
                               uint32_t* __restrict__ hlo, uint8_t* __restrict__ e /* (P, Se), Se even */, int q, int fw = 2) {
    extern __shared__ int8_t lt_sm[];
    int8_t* sM = lt_sm;             // (n, b)
    int8_t* sD = lt_sm + n * b;     // (P, n)
    for (int i = threadIdx.x; i < n * b; i += blockDim.x) sM[i] = M[i];
    for (int i = threadIdx.x; i < P * n; i += blockDim.x) sD[i] = D[i];
    __syncthreads();
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int8_t k[QSFT_MAX_N];
    const int8_t* row = loc + (size_t)s * ld;
    for (int u = 0; u < n; ++u) k[u] = row[u];
    uint32_t hi = 0, lo = 0;
    const int mul = (q == 2) ? 2 : 1;                 // q = 2: phases in quarter turns (see lt_digits)
    for (int i = 0; i < b; ++i) {
        int acc = 0;
        for (int u = 0; u < n; ++u) acc += (int)sM[u * b + i] * (int)k[u];
        if (i < b1) hi = (hi << fw) | (uint32_t)(acc % q * mul);       // fw bits per digit: 2, or 3 for q = 5 / 7
        else lo = (lo << fw) | (uint32_t)(acc % q * mul);
    }
    hhi[s] = hi;
    hlo[s] = lo;
    for (int p = 0; p < P; ++p) {
        int acc = 0;
        for (int u = 0; u < n; ++u) acc += (int)sD[p * n + u] * (int)k[u];
        e[(size_t)p * Se + s] = (uint8_t)(acc % q * mul);
    }
}

__global__ void lt_amax_kernel(const float2* __restrict__ a, long long S, unsigned int* __restrict__ amax_bits) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.f;
    if (s < S) m = fmaxf(fabsf(a[s].x), fabsf(a[s].y));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(amax_bits, __float_as_uint(m));   // non-negative floats order like uints
}

// balanced base-128 limbs of round(x * scale): v = (l0 * 128 + l1) * 128 + l2, limbs in [-64, 64]
__device__ __forceinline__ void limbs3(float x, float scale, int (&l)[3]) {
    int v = __float2int_rn(x * scale);
    int l2 = ((v + 64) & 127) - 64;
    v = (v - l2) >> 7;
    int l1 = ((v + 64) & 127) - 64;
    v = (v - l1) >> 7;
    l[0] = v; l[1] = l1; l[2] = l2;
}

// alimb[s] = {re limbs 0..2, im limbs 0..2, pad, pad} as int8x8.
// pass 0: v = round(a * scale), scale = (2^20 - 1) / max|a|: absolute error <= 0.5 / scale per component, i.e. ~5e-7 max|a|
//         -- enough for 1e-5 RELATIVE accuracy of the recovered coefficients only while min|a| >= ~0.1 max|a|.
// pass 1: the residual a * scale - v (|.| <= 0.5, computed with one fma) scaled by 2^21 and sliced the same way; its GEMM is
//         accumulated onto the first one with inv_scale / 2^21.  Together 41 bits below max|a|: the error left is the fp32
//         rounding of the samples themselves.  Only run when the strengths span a wide range (see qsft_eval_synth_lattice_ex).
__global__ void lt_quant_kernel(const float2* __restrict__ a, long long S, const unsigned int* __restrict__ amax_bits,
                                float* __restrict__ inv_scale /* [2]: pass 0, pass 1 */, int2* __restrict__ alimb, int pass) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const float amax = __uint_as_float(*amax_bits);
    // a power of two: max|a| = m 2^e with m in [0.5, 1) -> scale = 2^(20 - e), |a scale| < 2^20, and 1 / scale is exact (the
    // two passes then add up to a to ~2^-41 max|a|; with scale = (2^20 - 1) / max|a| the rounding of the reciprocal would
    // leave a common relative error of 1e-7 on all coefficients)
    int ex = 0;
    if (amax > 0.f) (void)frexpf(amax, &ex);
    // three balanced limbs reach +-(64 * 128^2 + 64 * 128 + 64) = 2^20 + 8256: a maximum that is (just above) a power of two
    // -- unit-modulus strengths, the reference's default -- can take one more bit
    if (amax > 0.f && ldexpf(amax, LT_SCALE_BITS + 1 - ex) <= 1056832.0f) --ex;
    const float scale = amax > 0.f ? ldexpf(1.0f, LT_SCALE_BITS - ex) : 0.f;
    if (s == 0 && pass == 0) {
        inv_scale[0] = amax > 0.f ? ldexpf(1.0f, ex - LT_SCALE_BITS) : 0.f;
        inv_scale[1] = amax > 0.f ? ldexpf(1.0f, ex - LT_SCALE_BITS - 21) : 0.f;
    }
    if (s >= S) return;
    float x = a[s].x, y = a[s].y, sc = scale;
    if (pass == 1) {
        x = fmaf(x, scale, -(float)__float2int_rn(x * scale));          // residual of pass 0 in units of 1 / scale
        y = fmaf(y, scale, -(float)__float2int_rn(y * scale));
        sc = 2097152.0f;                                                 // 2^21: |x * sc| <= 2^20
    }
    int lr[3], li[3];
    limbs3(x, sc, lr);
    limbs3(y, sc, li);
    uint32_t w0 = (uint32_t)(lr[0] & 0xff) | ((uint32_t)(lr[1] & 0xff) << 8) | ((uint32_t)(lr[2] & 0xff) << 16);
    uint32_t w1 = (uint32_t)(li[0] & 0xff) | ((uint32_t)(li[1] & 0xff) << 8) | ((uint32_t)(li[2] & 0xff) << 16);
    alimb[s] = make_int2((int)w0, (int)w1);
}

// smallest non-zero max(|re|, |im|) of the strengths (dynamic range of the signal, decides the residual pass)
__global__ void lt_amin_kernel(const float2* __restrict__ a, long long S, unsigned int* __restrict__ amin_bits) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float m = 3.0e38f;
    if (s < S) {
        const float v = fmaxf(fabsf(a[s].x), fabsf(a[s].y));
        if (v > 0.f) m = v;
    }
    for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMin(amin_bits, __float_as_uint(m));
}

// A'[(p * Mhi + l_hi) * 2 + part][2 s + comp]:  part 0 (Re row): (er, -ei),  part 1 (Im row): (ei, er),
// (er, ei) = i^(<h_hi(s), l_hi> + e[p][s]).  One thread owns two support elements of one l_hi and walks over the
// delay rows; the four possible byte pairs of a row are packed in a 64-bit constant and picked by a shift.
__global__ void __launch_bounds__(256)
lt_agen_kernel(const uint32_t* __restrict__ hhi, const uint8_t* __restrict__ e, long long S, long long Se, int b1, int P,
               long long Mhi, long long Kp, uint32_t* __restrict__ A, int spread) {
    const long long pair = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // two support elements = 4 K' bytes
    const uint32_t lhi = blockIdx.y;
    if (pair * 4 >= Kp) return;
    const long long s0 = 2 * pair;
    const bool live0 = s0 < S, live1 = s0 + 1 < S;
    const uint32_t ldig = lt_digits(lhi, spread);
    const uint32_t t0 = live0 ? dot4(hhi[s0], ldig, b1) : 0u;
    const uint32_t t1 = live1 ? dot4(hhi[s0 + 1], ldig, b1) : 0u;
    // byte pairs (lo byte first) for rotation r = 0..3:  Re row (er, -ei) = 01 00 | 00 FF | FF 00 | 00 01
    //                                                    Im row (ei,  er) = 00 01 | 01 00 | 00 FF | FF 00
    constexpr unsigned long long kRe = 0x010000FFFF000001ull, kIm = 0x00FFFF0000010100ull;
    const size_t row_words = (size_t)Kp / 4;
    uint32_t* out = A + ((size_t)(2 * lhi)) * row_words + pair;
    const uint8_t* ep = e + s0;
    for (int p = 0; p < P; ++p) {
        uint32_t wre = 0, wim = 0;
        if (live0) {
            uint32_t r0 = t0, r1 = t1;
            if (live1) {
                const uint32_t ee = *reinterpret_cast<const uint16_t*>(ep + (size_t)p * Se);
                r0 += ee & 0xffu;
                r1 += ee >> 8;
            } else {
                r0 += ep[(size_t)p * Se];
            }
            r0 &= 3u;
            r1 &= 3u;
            wre = (uint32_t)(kRe >> (16 * r0)) & 0xffffu;
            wim = (uint32_t)(kIm >> (16 * r0)) & 0xffffu;
            if (live1) {
                wre |= ((uint32_t)(kRe >> (16 * r1)) & 0xffffu) << 16;
                wim |= ((uint32_t)(kIm >> (16 * r1)) & 0xffffu) << 16;
            }
        }
        uint32_t* o = out + (size_t)p * (size_t)(2 * Mhi) * row_words;
        o[0] = wre;
        o[row_words] = wim;
    }
}

// B'_l[l_lo][2 s + comp] = limb l of (Re, Im) of a_s * i^<h_lo(s), l_lo>  (shared by all delay rows).
// A rotation by i^r is a byte swap of the (x, y) pair (PRMT) followed by a per-byte conditional negate
// ((w ^ m) - m with SIMD-in-word subtract).  One thread owns two support elements and walks over LT_BGEN_LL consecutive
// l_lo: the limb bytes are loaded and packed once, only the two rotations change per l_lo.
constexpr int LT_BGEN_LL = 8;
static_assert(LT_BGEN_LL == 8, "lt_bgen_kernel steps through the three low bits of l_lo");
__global__ void __launch_bounds__(256)
lt_bgen_kernel(const uint32_t* __restrict__ hlo, const int2* __restrict__ alimb, long long S, int b2, long long Nlo,
               long long Kp, uint32_t* __restrict__ Bq, int spread) {
    const long long pair = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pair * 4 >= Kp) return;
    const long long s0 = 2 * pair;
    const bool live0 = s0 < S, live1 = s0 + 1 < S;
    uint32_t base[3] = {0, 0, 0};
    uint32_t h0w = 0, h1w = 0;
    if (live0) {
        const int2 w = alimb[s0];
        h0w = hlo[s0];
#pragma unroll
        for (int l = 0; l < 3; ++l)
            base[l] |= (((uint32_t)w.x >> (8 * l)) & 0xffu) | ((((uint32_t)w.y >> (8 * l)) & 0xffu) << 8);
    }
    if (live1) {
        const int2 w = alimb[s0 + 1];
        h1w = hlo[s0 + 1];
#pragma unroll
        for (int l = 0; l < 3; ++l)
            base[l] |= ((((uint32_t)w.x >> (8 * l)) & 0xffu) << 16) | ((((uint32_t)w.y >> (8 * l)) & 0xffu) << 24);
    }
    const size_t row_words = (size_t)Kp / 4;
    const long long llo_begin = (long long)blockIdx.y * LT_BGEN_LL;
    // <h_lo, l_lo> is additive in the three low bits of l_lo (llo_begin is a multiple of eight): two dot products per group of
    // rows instead of two per row (q = 4: bits 0 / 1 are digit 0, bit 2 the low bit of digit 1; q = 2: one digit per bit)
    const uint32_t ldig0 = lt_digits((uint32_t)llo_begin, spread);
    const uint32_t rb0 = dot4(h0w, ldig0, b2), rb1 = dot4(h1w, ldig0, b2);
    const uint32_t f00 = h0w & 3u, f01 = (h0w >> 2) & 3u, f02 = (h0w >> 4) & 3u;
    const uint32_t f10 = h1w & 3u, f11 = (h1w >> 2) & 3u, f12 = (h1w >> 4) & 3u;
    const uint32_t c00 = f00, c01 = spread ? f01 : 2u * f00, c02 = spread ? f02 : f01;
    const uint32_t c10 = f10, c11 = spread ? f11 : 2u * f10, c12 = spread ? f12 : f11;
#pragma unroll
    for (int j = 0; j < LT_BGEN_LL; ++j) {
        const long long llo = llo_begin + j;
        if (llo >= Nlo) break;
        // dead elements have zero limbs: their rotation does not matter
        const uint32_t r0 = (rb0 + ((j & 1) ? c00 : 0u) + ((j & 2) ? c01 : 0u) + ((j & 4) ? c02 : 0u)) & 3u;
        const uint32_t r1 = (rb1 + ((j & 1) ? c10 : 0u) + ((j & 2) ? c11 : 0u) + ((j & 4) ? c12 : 0u)) & 3u;
        // rotation r: swap (x, y) iff r & 1; negate byte 0 iff (r & 1) ^ (r >> 1); negate byte 1 iff r >> 1
        const uint32_t b0 = r0 & 1u, h0 = r0 >> 1, b1 = r1 & 1u, h1 = r1 >> 1;
        const uint32_t sel = (b0 ? 0x01u : 0x10u) | ((b1 ? 0x23u : 0x32u) << 8);
        const uint32_t m = ((0u - (b0 ^ h0)) & 0x000000ffu) | ((0u - h0) & 0x0000ff00u) |
                           ((0u - (b1 ^ h1)) & 0x00ff0000u) | ((0u - h1) & 0xff000000u);
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            const uint32_t sw = __byte_perm(base[l], 0u, sel);
            Bq[((size_t)l * Nlo + (size_t)llo) * row_words + pair] = __vsub4(sw ^ m, m);
        }
    }
}

// ---- packed phase tables for the in-kernel generation of A' (FUSED_A) -------------------------------------
// T[l_hi][w]: 16 two-bit fields <h_hi(s), l_hi> mod 4 for s = 16 w .. 16 w + 15;  E2[p][w]: the delay phases e[p][s] packed
// the same way.  A' is then a pure function of (T + E2) mod 4, generated slab by slab in shared memory by the GEMM CTAs.
// One thread owns the sixteen support elements of word w and walks over LT_TTAB_LL consecutive l_hi (the bin-hash words
// are loaded once); blockIdx.y enumerates groups of LT_TTAB_LL rows.
constexpr int LT_TTAB_LL = 64;
// field-wise (a + b) mod 4 on sixteen packed two-bit fields
__device__ __forceinline__ uint32_t lt_add4(uint32_t x, uint32_t y) {
    constexpr uint32_t H = 0xAAAAAAAAu;
    return ((x & ~H) + (y & ~H)) ^ ((x ^ y) & H);
}
// The sixty-four rows of a group differ from its first row l0 (a multiple of 64) only in the six low BITS of the lattice index,
// and <h, l> is additive in them: bit k of the index adds the word G[k] -- for q = 4 bits (2 d, 2 d + 1) belong to digit d and
// add H_d and 2 H_d, H_d = the sixteen elements' field d of h; for q = 2 (one bit per digit, lt_digits) bit d adds H_d -- so one
// row costs one packed add instead of sixteen popcount dot products (ncu r4l: 56 us per config-5 block before).
__global__ void __launch_bounds__(256)
lt_ttab_kernel(const uint32_t* __restrict__ hhi, long long S, int b1, long long Tw, long long Mhi, uint32_t* __restrict__ T,
               int spread) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= Tw) return;
    uint32_t h[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const long long s = 16 * w + i;
        h[i] = s < S ? hhi[s] : 0u;                     // dead elements: rotation 0, like the per-element guard before
    }
    const long long l0 = (long long)blockIdx.y * LT_TTAB_LL;
    uint32_t base = 0;
    {
        const uint32_t ldig = lt_digits((uint32_t)l0, spread);
#pragma unroll
        for (int i = 0; i < 16; ++i) base |= dot4(h[i], ldig, b1) << (2 * i);
    }
    uint32_t G[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) {
        uint32_t Hd = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) Hd |= ((h[i] >> (2 * d)) & 3u) << (2 * i);
        G[d] = Hd;
    }
    if (!spread) {                                      // digits 0 .. 2: (H_0, 2 H_0, H_1, 2 H_1, H_2, 2 H_2)
        const uint32_t H0 = G[0], H1 = G[1], H2 = G[2];
        G[0] = H0; G[1] = lt_add4(H0, H0);
        G[2] = H1; G[3] = lt_add4(H1, H1);
        G[4] = H2; G[5] = lt_add4(H2, H2);
    }
    uint32_t L[8];
    L[0] = 0u; L[1] = G[0]; L[2] = G[1]; L[3] = lt_add4(G[0], G[1]);
    L[4] = G[2]; L[5] = lt_add4(G[2], G[0]); L[6] = lt_add4(G[2], G[1]); L[7] = lt_add4(L[6], G[0]);
#pragma unroll
    for (int hi = 0; hi < 8; ++hi) {
        uint32_t th = base;
        if (hi & 1) th = lt_add4(th, G[3]);
        if (hi & 2) th = lt_add4(th, G[4]);
        if (hi & 4) th = lt_add4(th, G[5]);
#pragma unroll
        for (int lo = 0; lo < 8; ++lo) {
            const long long lhi = l0 + hi * 8 + lo;
            if (lhi < Mhi) T[(size_t)lhi * Tw + w] = lt_add4(th, L[lo]);
        }
    }
}

__global__ void __launch_bounds__(256)
lt_etab_kernel(const uint8_t* __restrict__ e, long long S, long long Se, int P, long long Tw, uint32_t* __restrict__ E2) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int p = blockIdx.y;
    if (w >= Tw) return;
    uint32_t word = 0;
    for (int i = 0; i < 16; ++i) {
        const long long s = 16 * w + i;
        if (s < S) word |= (uint32_t)(e[(size_t)p * Se + s] & 3u) << (2 * i);
    }
    E2[(size_t)p * Tw + w] = word;
}

// sixteen support elements of one A' row: rotations r16 (2 bits each) -> 16 compressed bytes (+-1) + 8 metadata nibbles
// (same encoding as lt_agen_sp_kernel).  PX = false: sign bits spread by a multiply, masked, scaled by 0xFE and xor-ed
// (six instructions per four elements).  PX = true (opt-in QSFT_LATTICE_EXPAND=1, unmeasured): the multiply moves sign bit
// 2j of the group straight to the top bit of byte j (shifts 7 + 6i: bit 2j + 7 + 6i = 8b + 7 only for i = j = b, and nothing
// else ever lands on or carries into 8b + 6 / 8b + 7), PRMT in sign-replication mode turns the bytes into 0x00 / 0xFF and an
// OR sets the low bit: five instructions, no mask of the product.  Both are checked exhaustively per group on the CPU.
// every byte -> 0x00 / 0xFF according to its top bit: PTX prmt in its generic form replicates the sign of the selected byte
// when bit 3 of the selector nibble is set.  (CUDA's __byte_perm only honours three selector bits per nibble -- the compiler
// masks 0xBA98 to 0x3210, an identity permutation -- so this has to be the PTX instruction.)
__device__ __forceinline__ uint32_t lt_sign_bytes(uint32_t x) {
#ifdef QSFT_EMU   // CPU execution by tests/emu (test infrastructure; never defined in the product build)
    uint32_t o = 0;
    for (int i = 0; i < 4; ++i)
        if ((x >> (8 * i + 7)) & 1u) o |= 0xFFu << (8 * i);
    return o;
#else
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(0u), "r"(0xBA98u));
    return d;
#endif
}

template <bool PX>
__device__ __forceinline__ void ts_expand(uint32_t r16, bool im, uint32_t* a4, uint32_t& e1) {
    const uint32_t lo = r16 & 0x55555555u, hi = (r16 >> 1) & 0x55555555u;
    const uint32_t neg = im ? hi : (lo ^ hi);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        if (PX) {
            a4[g] = lt_sign_bytes(((neg >> (8 * g)) & 0x55u) * 0x02082080u) | 0x01010101u;
        } else {
            const uint32_t m = (((neg >> (8 * g)) & 0x55u) * 0x00041041u) & 0x01010101u;
            a4[g] = 0x01010101u ^ (m * 0xFEu);
        }
    }
    e1 = (lo | 0x88888888u) ^ (im ? 0x55555555u : 0u);
}

// Phase words of the tables -> expansion.  PX = false: SWAR add of the sixteen 2-bit fields, then ts_expand (the measured
// default).  PX = true: the two bit planes of (t + e) mod 4 directly -- lo = (t ^ e) & L, hi = ((t ^ e) >> 1 & L) ^ (t & e & L)
// -- five instead of eight instructions before the expansion.
template <bool PX>
__device__ __forceinline__ void ts_phase_expand(uint32_t tw, uint32_t ew, bool im, uint32_t* a4, uint32_t& e1) {
    if (!PX) {
        constexpr uint32_t H = 0xAAAAAAAAu;
        const uint32_t r16 = ((tw & ~H) + (ew & ~H)) ^ ((tw ^ ew) & H);
        ts_expand<false>(r16, im, a4, e1);
    } else {
        constexpr uint32_t L = 0x55555555u;
        const uint32_t x = tw ^ ew;
        const uint32_t lo = x & L, hi = ((x >> 1) & L) ^ (tw & ew & L);
        const uint32_t neg = im ? hi : (lo ^ hi);
#pragma unroll
        for (int g = 0; g < 4; ++g)
            a4[g] = lt_sign_bytes(((neg >> (8 * g)) & 0x55u) * 0x02082080u) | 0x01010101u;
        e1 = (lo | 0x88888888u) ^ (im ? 0x55555555u : 0u);
    }
}

// ---- q = 3 ---------------------------------------------------------------------------------------------------
// The cube roots of unity are integers of Z[w], w^2 = -1 - w: w^t = (1, 0), (0, 1), (-1, -1) in the basis (1, w), and the
// product of (a1 + b1 w) with (u + v w) is (a1 u - b1 v) + (b1 u + (a1 - b1) v) w -- a 2 x 2 integer matrix with entries
// 0, +-1.  With that matrix in place of the real embedding of i^t, the SAME dense GEMM evaluates
//     C_p[l_hi, l_lo] = sum_s  w^(<h_hi(s), l_hi> + e_ps)  *  x_s w^<h_lo(s), l_lo>          (a number c_1 + c_w w)
// for a REAL strength x_s: A' rows = (1-part, w-part) per l_hi, K' = (u, v) per support element, limbs of x_s as for q = 4.
// Two launches (x = Re a, x = Im a) give the complex coefficients c_1, c_w, and
//     X = c_1 + c_w w,  w = -1/2 + i sqrt(3)/2
// is formed by lt_combine3_kernel.  No 2:4 structure here (three of the four matrix entries can be non-zero): dense MMA.
__device__ __forceinline__ uint32_t lt_pack3(uint32_t l, int nd) {        // base-3 digits of l, two bits each, digit 0 lowest
    uint32_t out = 0;
    for (int i = 0; i < nd; ++i) {
        out |= (l % 3u) << (2 * i);
        l /= 3u;
    }
    return out;
}
__device__ __forceinline__ uint32_t lt_dot3(uint32_t a, uint32_t b, int nd) {   // sum of digit products mod 3
    uint32_t acc = 0;
    for (int i = 0; i < nd; ++i) acc += ((a >> (2 * i)) & 3u) * ((b >> (2 * i)) & 3u);
    return acc % 3u;
}

// A'[(p * Mhi + l_hi) * 2 + e][2 s + k']: row e of the matrix of w^t, t = (<h_hi(s), l_hi> + e[p][s]) mod 3:
//   t = 0: (1, 0 | 0, 1)   t = 1: (0, -1 | 1, -1)   t = 2: (-1, 1 | -1, 0)
__global__ void __launch_bounds__(256)
lt_agen3_kernel(const uint32_t* __restrict__ hhi, const uint8_t* __restrict__ e, long long S, long long Se, int b1, int P,
                long long Mhi, long long Kp, uint32_t* __restrict__ A) {
    const long long pair = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // two support elements = 4 K' bytes
    const uint32_t lhi = blockIdx.y;
    if (pair * 4 >= Kp) return;
    const long long s0 = 2 * pair;
    const bool live0 = s0 < S, live1 = s0 + 1 < S;
    const uint32_t lp = lt_pack3(lhi, b1);
    const uint32_t t0 = live0 ? lt_dot3(hhi[s0], lp, b1) : 0u;
    const uint32_t t1 = live1 ? lt_dot3(hhi[s0 + 1], lp, b1) : 0u;
    // byte pairs (k' = 0 low byte) per t: row 0 = 01 00 | 00 FF | FF 01, row 1 = 00 01 | 01 FF | FF 00
    constexpr unsigned long long kR0 = 0x000001FFFF000001ull, kR1 = 0x000000FFFF010100ull;
    const size_t row_words = (size_t)Kp / 4;
    uint32_t* out = A + ((size_t)(2 * lhi)) * row_words + pair;
    for (int p = 0; p < P; ++p) {
        uint32_t w0 = 0, w1 = 0;
        if (live0) {
            const uint32_t r0 = (t0 + e[(size_t)p * Se + s0]) % 3u;
            w0 = (uint32_t)(kR0 >> (16 * r0)) & 0xffffu;
            w1 = (uint32_t)(kR1 >> (16 * r0)) & 0xffffu;
            if (live1) {
                const uint32_t r1 = (t1 + e[(size_t)p * Se + s0 + 1]) % 3u;
                w0 |= ((uint32_t)(kR0 >> (16 * r1)) & 0xffffu) << 16;
                w1 |= ((uint32_t)(kR1 >> (16 * r1)) & 0xffffu) << 16;
            }
        }
        uint32_t* o = out + (size_t)p * (size_t)(2 * Mhi) * row_words;
        o[0] = w0;
        o[row_words] = w1;
    }
}

// B'_l[l_lo][2 s + k'] = limb l of x_s * (alpha, beta)[k'], (alpha, beta) = w^<h_lo(s), l_lo> in the basis (1, w);
// x_s = Re a_s (part 0) or Im a_s (part 1).  Balanced limbs negate digit by digit, so -x has the negated limbs.
__global__ void __launch_bounds__(256)
lt_bgen3_kernel(const uint32_t* __restrict__ hlo, const int2* __restrict__ alimb, long long S, int b2, long long Nlo,
                long long Kp, int part, uint32_t* __restrict__ Bq) {
    const long long pair = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long llo = blockIdx.y;
    if (pair * 4 >= Kp) return;
    const long long s0 = 2 * pair;
    const uint32_t lp = lt_pack3((uint32_t)llo, b2);
    uint32_t words[3] = {0u, 0u, 0u};
    for (int h = 0; h < 2; ++h) {
        const long long s = s0 + h;
        if (s >= S) break;
        const int2 w = alimb[s];
        const uint32_t lw = part ? (uint32_t)w.y : (uint32_t)w.x;
        const uint32_t t = lt_dot3(hlo[s], lp, b2);
        const int al = t == 0 ? 1 : t == 1 ? 0 : -1, be = t == 0 ? 0 : t == 1 ? 1 : -1;
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            const int limb = (int)(int8_t)((lw >> (8 * l)) & 0xffu);
            words[l] |= (((uint32_t)(al * limb) & 0xffu) | (((uint32_t)(be * limb) & 0xffu) << 8)) << (16 * h);
        }
    }
    const size_t row_words = (size_t)Kp / 4;
#pragma unroll
    for (int l = 0; l < 3; ++l) Bq[((size_t)l * Nlo + (size_t)llo) * row_words + pair] = words[l];
}

// X = c_1 + c_w w with cre = (Re c_1, Re c_w) and cim = (Im c_1, Im c_w) per sample
__global__ void lt_combine3_kernel(const float2* __restrict__ cre, const float2* __restrict__ cim, long long N, float2* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float2 r = cre[i], m = cim[i];
    const float h = 0.86602540378443864676f;
    out[i] = make_float2(r.x - 0.5f * r.y - h * m.y, m.x - 0.5f * m.y + h * r.y);
}


// ---- odd primes q = 5, 7 (and, same construction, 3) -----------------------------------------------------------------------
// The q = 3 idea for any odd prime: Z[w] with basis (1, w, ..., w^(d-1)), d = q - 1, and w^d = -(1 + w + ... + w^(d-1)).  The
// coordinates of w^m are e_m for m < d and (-1, ..., -1) for m = d, so multiplication by w^t is the d x d integer matrix
//     M_t[r][c] = [(t + c) mod q == r] - [(t + c) mod q == d]            (entries 0, +-1),
// A' carries d rows per l_hi and d bytes per support element, B' the limbs of x_s times the coordinates of w^<h_lo, l_lo>, and
// the same dense GEMM gives the d coordinates of every sample for a REAL strength; two launches (Re a, Im a) and
// lt_combineq_kernel form X = sum_m (cre_m + i cim_m) w^m.  d is even, so rows (2 j, 2 j + 1) of one l_hi are one float2 of the
// GEMM's output and the kernel runs unchanged with Mhi * d / 2 "row pairs" per delay row: planes [p][l_hi][j][l_lo].
// 4 d^2 int8 MACs per pair and limb: 96 x 3 for q = 5, 216 x 3 for q = 7 (q = 3: 24 x 3) -- still far below the per-pair
// epilogue of the plain K = n kernel for q = 5, about even for q = 7.
template <int Q> struct LtQ {
    static constexpr int D = Q - 1;
    static constexpr int FW = Q <= 4 ? 2 : 3;                     // bits per packed digit
};
template <int Q>
__device__ __forceinline__ uint32_t lt_packq(uint32_t l, int nd) {          // base-q digits of l, FW bits each, digit 0 lowest
    uint32_t out = 0;
    for (int i = 0; i < nd; ++i) {
        out |= (l % (uint32_t)Q) << (LtQ<Q>::FW * i);
        l /= (uint32_t)Q;
    }
    return out;
}
template <int Q>
__device__ __forceinline__ uint32_t lt_dotq(uint32_t a, uint32_t b, int nd) {   // sum of digit products mod q
    constexpr int FW = LtQ<Q>::FW;
    constexpr uint32_t MK = (1u << FW) - 1u;
    uint32_t acc = 0;
    for (int i = 0; i < nd; ++i) acc += ((a >> (FW * i)) & MK) * ((b >> (FW * i)) & MK);
    return acc % (uint32_t)Q;
}

// A'[(p * Mhi + l_hi) * d + r][d s + c] = M_t[r][c], t = (<h_hi(s), l_hi> + e[p][s]) mod q.  One thread owns one 4-byte word of
// the K' range (its bytes belong to at most three support elements) of one l_hi and walks over the delay rows and the d rows.
template <int Q>
__global__ void __launch_bounds__(256)
lt_agenq_kernel(const uint32_t* __restrict__ hhi, const uint8_t* __restrict__ e, long long S, long long Se, int b1, int P,
                long long Mhi, long long Kp, uint32_t* __restrict__ A) {
    constexpr int D = LtQ<Q>::D;
    const long long word = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lhi = blockIdx.y;
    if (word * 4 >= Kp) return;
    const uint32_t lp = lt_packq<Q>(lhi, b1);
    long long sb[4];
    int cb[4], tb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long kb = word * 4 + i;
        sb[i] = kb / D;
        cb[i] = (int)(kb - sb[i] * D);
        tb[i] = sb[i] < S ? (int)lt_dotq<Q>(hhi[sb[i]], lp, b1) : -1;
    }
    const size_t row_words = (size_t)Kp / 4;
    for (int p = 0; p < P; ++p) {
        int m[4];                                                  // (t + c) mod q per byte, -1 for K padding
#pragma unroll
        for (int i = 0; i < 4; ++i) m[i] = tb[i] < 0 ? -1 : (tb[i] + (int)e[(size_t)p * Se + sb[i]] + cb[i]) % Q;
        uint32_t* o = A + ((size_t)p * (size_t)Mhi + lhi) * D * row_words + word;
#pragma unroll
        for (int r = 0; r < D; ++r) {
            uint32_t w = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int v = m[i] < 0 ? 0 : (m[i] == r ? 1 : 0) - (m[i] == D ? 1 : 0);
                w |= ((uint32_t)v & 0xffu) << (8 * i);
            }
            o[(size_t)r * row_words] = w;
        }
    }
}

// B'_l[l_lo][d s + c] = limb l of x_s times coordinate c of w^<h_lo(s), l_lo>; x_s = Re a_s (part 0) or Im a_s (part 1).
template <int Q>
__global__ void __launch_bounds__(256)
lt_bgenq_kernel(const uint32_t* __restrict__ hlo, const int2* __restrict__ alimb, long long S, int b2, long long Nlo,
                long long Kp, int part, uint32_t* __restrict__ Bq) {
    constexpr int D = LtQ<Q>::D;
    const long long word = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long llo = blockIdx.y;
    if (word * 4 >= Kp) return;
    const uint32_t lp = lt_packq<Q>((uint32_t)llo, b2);
    uint32_t words[3] = {0u, 0u, 0u};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long kb = word * 4 + i;
        const long long s = kb / D;
        if (s >= S) break;
        const int c = (int)(kb - s * D);
        const int t = (int)lt_dotq<Q>(hlo[s], lp, b2);
        const int coef = (t == c ? 1 : 0) - (t == D ? 1 : 0);
        const int2 w = alimb[s];
        const uint32_t lw = part ? (uint32_t)w.y : (uint32_t)w.x;
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            const int limb = (int)(int8_t)((lw >> (8 * l)) & 0xffu);
            words[l] |= ((uint32_t)(coef * limb) & 0xffu) << (8 * i);
        }
    }
    const size_t row_words = (size_t)Kp / 4;
#pragma unroll
    for (int l = 0; l < 3; ++l) Bq[((size_t)l * Nlo + (size_t)llo) * row_words + word] = words[l];
}

// X = sum_m (cre_m + i cim_m) w^m, w = exp(2 pi i / q); planes [(p, l_hi)][j][l_lo] float2 = coordinates (2 j, 2 j + 1)
template <int Q>
__global__ void lt_combineq_kernel(const float2* __restrict__ cre, const float2* __restrict__ cim, long long rows /* P * Mhi */,
                                   long long Nlo, float2* __restrict__ out) {
    constexpr int D = LtQ<Q>::D;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * Nlo) return;
    const long long row = i / Nlo, llo = i - row * Nlo;
    float xr = 0.f, xi = 0.f;
#pragma unroll
    for (int j = 0; j < D / 2; ++j) {
        const size_t at = ((size_t)row * (D / 2) + j) * (size_t)Nlo + (size_t)llo;
        const float2 r = cre[at], m = cim[at];
        float s0, c0, s1, c1;
        sincospif(2.0f * (float)(2 * j) / (float)Q, &s0, &c0);
        sincospif(2.0f * (float)(2 * j + 1) / (float)Q, &s1, &c1);
        xr += r.x * c0 - m.x * s0 + r.y * c1 - m.y * s1;
        xi += r.x * s0 + m.x * c0 + r.y * s1 + m.y * c1;
    }
    out[i] = make_float2(xr, xi);
}

// ---- the GEMM ---------------------------------------------------------------------------------------------
// Dense cross-check kernel (QSFT_LATTICE_SPARSE=0): A' materialised in HBM by lt_agen_kernel, cta_group::1.
// RAGGED (q = 3: 2 * Mhi and Nlo are powers of three): the row / column counts need not be multiples of the tile; the TMA
// boxes read zeros (or, between the limb blocks of B', a neighbour's rows) beyond the edge and the epilogue masks its stores.
template <bool RAGGED>
__global__ void __launch_bounds__(LT_THREADS, 1)
lt_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int nkb, int Mhi, int Nlo,
               const float* __restrict__ inv_scale_ptr, float2* __restrict__ out, int accumulate, long long total_rows) {
    extern __shared__ uint8_t lt_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)lt_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)LT_STAGES * LT_STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + LT_STAGES;
    uint64_t* tfull = bars + 2 * LT_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // n-tiles fastest: the CTAs sharing one A' m-tile (one delay row, 64 l_hi) are neighbours, so its slabs are fetched
    // from DRAM once; the small limb operand B' (3 * Nlo rows) is shared by every CTA and lives in L2
    const int ntile = blockIdx.x, mtile = blockIdx.y;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < LT_STAGES; ++i) {
            lt_mbar_init(&full[i], 1);
            lt_mbar_init(&empty[i], 1);
        }
        lt_mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(lt_smem_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int stage = kb % LT_STAGES;
                const uint32_t ph = (uint32_t)(kb / LT_STAGES) & 1u;
                lt_mbar_wait(&empty[stage], ph ^ 1u);
                lt_mbar_expect_tx(&full[stage], LT_STAGE_BYTES);
                uint8_t* st = base + (size_t)stage * LT_STAGE_BYTES;
                lt_tma_2d(st, &tmA, kb * LT_BK, mtile * LT_BM, &full[stage]);
#pragma unroll
                for (int l = 0; l < LT_LIMBS; ++l)
                    lt_tma_2d(st + LT_BM * LT_BK + l * (LT_BN * LT_BK), &tmB, kb * LT_BK, l * Nlo + ntile * LT_BN, &full[stage]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptors: D = S32, A = B = signed int8, K-major, M = 128; limbs 0 and 1 sit in adjacent
            // shared-memory tiles and adjacent TMEM columns, so they are one N = 256 MMA (A is read twice, not three times)
            constexpr uint32_t idesc_base = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(LT_BM >> 4) << 24);
            constexpr uint32_t idesc256 = idesc_base | ((uint32_t)((2 * LT_BN) >> 3) << 17);
            constexpr uint32_t idesc128 = idesc_base | ((uint32_t)(LT_BN >> 3) << 17);
            static_assert(LT_LIMBS == 3, "MMA issue below is written for three limbs");
            for (int kb = 0; kb < nkb; ++kb) {
                const int stage = kb % LT_STAGES;
                const uint32_t ph = (uint32_t)(kb / LT_STAGES) & 1u;
                lt_mbar_wait(&full[stage], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = lt_smem_u32(base + (size_t)stage * LT_STAGE_BYTES);
                const uint64_t adesc = lt_desc(sa);
                const uint64_t bdesc01 = lt_desc(sa + LT_BM * LT_BK);
                const uint64_t bdesc2 = lt_desc(sa + LT_BM * LT_BK + 2 * (LT_BN * LT_BK));
#pragma unroll
                for (int k = 0; k < LT_BK / 32; ++k) {
                    const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                    lt_umma_i8(tmem_base, adesc + (uint64_t)(2 * k), bdesc01 + (uint64_t)(2 * k), idesc256, acc);
                    lt_umma_i8(tmem_base + (uint32_t)(2 * LT_BN), adesc + (uint64_t)(2 * k), bdesc2 + (uint64_t)(2 * k), idesc128, acc);
                }
                lt_commit(&empty[stage]);
            }
            lt_commit(tfull);
        }
    } else {
        // epilogue: row r = 2 * l_hi_local + part lives in TMEM lane r; neighbouring lanes hold (Re, Im) of one l_hi
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const long long grow = (long long)mtile * LT_BM + row;          // (p * Mhi + l_hi) * 2 + part
        const int p = (int)(grow / (2 * Mhi));
        const int lhi = (int)(grow - (long long)p * 2 * Mhi) >> 1;
        const bool odd = lane & 1;
        const int llo0 = ntile * LT_BN;
        const double inv_scale = (double)(*inv_scale_ptr);
        lt_mbar_wait(tfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        float2* orow = out + ((size_t)p * Mhi + lhi) * Nlo + llo0;
        const bool row_ok = !RAGGED || grow < total_rows;
#pragma unroll 1
        for (int ch = 0; ch < LT_BN / 16; ++ch) {
            uint32_t a0[16], a1[16], a2[16];
            lt_ld16(taddr + (uint32_t)(0 * LT_BN + ch * 16), a0);
            lt_ld16(taddr + (uint32_t)(1 * LT_BN + ch * 16), a1);
            lt_ld16(taddr + (uint32_t)(2 * LT_BN + ch * 16), a2);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float val[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const long long v = ((long long)(int)a0[j] * 128 + (long long)(int)a1[j]) * 128 + (long long)(int)a2[j];
                val[j] = (float)((double)v * inv_scale);
            }
            // even lane (Re row) keeps columns 0..7, odd lane (Im row) keeps columns 8..15
            float2 o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float send = odd ? val[j] : val[8 + j];
                const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
                o[j] = odd ? make_float2(recv, val[8 + j]) : make_float2(val[j], recv);
            }
            if (RAGGED) {
                const int col0 = llo0 + ch * 16 + (odd ? 8 : 0);
                float2* dst = orow + ch * 16 + (odd ? 8 : 0);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (row_ok && col0 + j < Nlo) {
                        float2 w = o[j];
                        if (accumulate) {
                            const float2 pv = dst[j];
                            w = make_float2(pv.x + w.x, pv.y + w.y);
                        }
                        dst[j] = w;
                    }
                }
            } else {
                float4* dst = reinterpret_cast<float4*>(orow + ch * 16 + (odd ? 8 : 0));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 w = make_float4(o[2 * j].x, o[2 * j].y, o[2 * j + 1].x, o[2 * j + 1].y);
                    if (accumulate) {                               // residual pass: onto the first pass's samples
                        const float4 p = dst[j];
                        w = make_float4(p.x + w.x, p.y + w.y, p.z + w.z, p.w + w.w);
                    }
                    dst[j] = w;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- 2:4 structured-sparse variant ------------------------------------------------------------------------
// Every (Re, Im) byte pair of A' holds exactly one zero, so four consecutive K' bytes (two support elements) hold exactly two
// non-zeros: A' is a valid 2:4 sparse operand for tcgen05.mma.sp.  Stored compressed (one +-1 byte per row and support
// element) plus 4 bits of metadata per four logical bytes: half the tensor work and 0.625x the operand bytes, bit-exact.
// The GEMM runs on CTA pairs (cta_group::2): 256 rows of A' x 128 columns x 3 limbs per pair; each CTA stages its own 128
// compressed rows + metadata, ONE of limbs 0 / 1 and half of the limb-2 columns, so the limb operand is fetched once per pair.
constexpr int SP_BK = 256;                                   // logical K' bytes per stage (128 support elements)
constexpr int SP_B01_BYTES = 2 * 128 * 128;                  // this CTA's limb (0 or 1), two 128-byte K halves
constexpr int SP_B2_BYTES = 2 * 64 * 128;                    // this CTA's 64 columns of limb 2, two K halves

__device__ __forceinline__ uint32_t sp_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t sp_mapa(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void sp_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion is signalled on the pair leader's barrier (cluster address `bar_cluster`)
__device__ __forceinline__ void sp_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            lt_smem_u32(dst)),
        "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void sp_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     lt_smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
// ---- sparse variant with A' generated straight into tensor memory --------------------------------------------
// Measured (tools/sp_probe.cu): a sparse MMA pair (N = 256 + N = 128) costs 259 cycles when A' is read from shared memory but
// 192 = 128 + 64 cycles (the issue floor) when A' sits in tensor memory; tcgen05.cp does not overlap with the MMAs (268).
// So the four otherwise idle epilogue warps PRODUCE the compressed A' rows and their metadata from the packed phase tables
// (2 bits per row and support element, L2 resident) and write them with tcgen05.st into a double-buffered TMEM region:
// A' never exists in HBM or shared memory, the TMA ring carries only the limb operand (48 KB per CTA and stage).
constexpr int TS_STAGES = 4;
constexpr int TS_STAGE_BYTES = SP_B01_BYTES + SP_B2_BYTES;       // 48 KB
constexpr int TS_OFF_B2 = SP_B01_BYTES;
constexpr size_t TS_SMEM = 1024 + (size_t)TS_STAGES * TS_STAGE_BYTES + 256;
constexpr uint32_t TS_TMEM_A = 384, TS_TMEM_BUF = 40;            // per buffer: 32 columns of A' (4 k-steps) + 8 of metadata
constexpr int TS_NBUF = 3;
constexpr int TS_THREADS = 320;                                  // warps: 0 TMA, 1 MMA, 2..9 A' producers + epilogue                                       // TMEM operand buffers (384 + 3 * 40 = 504 columns)

__device__ __forceinline__ void ts_umma_i8(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t e_tmem, uint32_t idesc,
                                           uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.sp.cta_group::2.kind::i8 [%0], [%1], %2, [%3], %4, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(e_tmem), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void ts_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
            taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void ts_st4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3])
                 : "memory");
}

// PROBE (measurement aid, QSFT_LT_PROBE; results are garbage when non-zero): bit 0 = after the first ring lap no limb slabs are
// loaded (the MMAs re-read what the ring holds), bit 1 = after the first TMEM lap the producers write nothing: which of the
// two feeds keeps a stage above the 768-cycle issue floor of its eight MMAs.
template <int PROBE>
__global__ void __launch_bounds__(TS_THREADS, 1)
lt_gemm_spts_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmB2, int nkb, int Mhi, int Nlo,
                    int n_mtiles, const uint32_t* __restrict__ Ttab, const uint32_t* __restrict__ Etab, int Tw,
                    const float* __restrict__ inv_scale_ptr, float2* __restrict__ out, int accumulate) {
    extern __shared__ uint8_t lt_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)lt_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)TS_STAGES * TS_STAGE_BYTES);
    uint64_t* full = bars;                       // leader only: the limb slabs of both CTAs complete on it
    uint64_t* empty = bars + TS_STAGES;          // per CTA, released by the leader's multicast commit
    uint64_t* aready = bars + 2 * TS_STAGES;     // leader only: 2 CTAs x 8 producer warps have written TMEM buffer b
    uint64_t* tfree = aready + TS_NBUF;          // per CTA: the MMAs reading TMEM buffer b have completed
    uint64_t* tfull = tfree + TS_NBUF;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = sp_cluster_rank();
    // m-tiles fastest: A' is generated, not loaded, so the only streamed operand is B'; all resident pairs then walk the K
    // range of the SAME n-tile (77 MB at config 5, L2 resident even when the pairs drift apart) instead of all eight
    // (measured with n fastest: 90 GB of DRAM reads per launch, L2 hit 50 %)
    const int ntile = blockIdx.y;
    const int mtile = (int)blockIdx.x;           // = 2 * pair + rank: the cluster spans two consecutive x

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < TS_STAGES; ++i) {
            lt_mbar_init(&full[i], 1);
            lt_mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < TS_NBUF; ++i) {
            lt_mbar_init(&aready[i], 16);
            lt_mbar_init(&tfree[i], 1);
        }
        lt_mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(lt_smem_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    sp_cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int stage = kb % TS_STAGES;
                const uint32_t ph = (uint32_t)(kb / TS_STAGES) & 1u;
                lt_mbar_wait(&empty[stage], ph ^ 1u);
                if ((PROBE & 1) && kb >= TS_STAGES) {
                    if (rank == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(lt_smem_u32(&full[stage])) : "memory");
                    continue;
                }
                if (rank == 0) lt_mbar_expect_tx(&full[stage], 2 * TS_STAGE_BYTES);
                const uint32_t fl = sp_mapa(lt_smem_u32(&full[stage]), 0);
                uint8_t* st = base + (size_t)stage * TS_STAGE_BYTES;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    sp_tma_2d(st + h * 16384, &tmB, kb * SP_BK + h * 128, (int)rank * Nlo + ntile * LT_BN, fl);
                    sp_tma_2d(st + TS_OFF_B2 + h * 8192, &tmB2, kb * SP_BK + h * 128, 2 * Nlo + ntile * LT_BN + (int)rank * 64, fl);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc_base = (1u << 2) | (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 4) << 24);
            constexpr uint32_t idesc256 = idesc_base | ((256u >> 3) << 17);
            constexpr uint32_t idesc128 = idesc_base | ((128u >> 3) << 17);
            for (int kb = 0; kb < nkb; ++kb) {
                const int stage = kb % TS_STAGES;
                const uint32_t ph = (uint32_t)(kb / TS_STAGES) & 1u;
                const int buf = kb % TS_NBUF;
                lt_mbar_wait(&aready[buf], (uint32_t)(kb / TS_NBUF) & 1u);
                lt_mbar_wait(&full[stage], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = lt_smem_u32(base + (size_t)stage * TS_STAGE_BYTES);
                const uint32_t ta = tmem_base + TS_TMEM_A + TS_TMEM_BUF * (uint32_t)buf, te = ta + 32u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t acc = (kb > 0 || j > 0) ? 1u : 0u;
                    const uint64_t koff = (uint64_t)(4 * (j & 1));
                    const uint64_t b01 = lt_desc(sa + (j >> 1) * 16384) + koff;
                    const uint64_t b2 = lt_desc(sa + TS_OFF_B2 + (j >> 1) * 8192) + koff;
                    ts_umma_i8(tmem_base, ta + 8u * j, b01, te + 2u * j, idesc256, acc);
                    ts_umma_i8(tmem_base + 256u, ta + 8u * j, b2, te + 2u * j, idesc128, acc);
                }
                sp_commit_pair(&empty[stage]);
                sp_commit_pair(&tfree[buf]);
            }
            sp_commit_pair(tfull);
        }
    } else {
        // two warps per TMEM lane quarter: `half` 0 owns the first 64 support elements of a stage (k-steps 0, 1) and the
        // first four column chunks of the epilogue, `half` 1 the rest
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const bool odd = lane & 1;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        {
            // A' producer: this thread owns TMEM lane `row` = A' row (p, l_hi, part); per stage 64 support elements =
            // four packed words of T[l_hi] and E2[p]; t = (T + E2) mod 4 on sixteen 2-bit fields at once
            const int mt_eff = mtile < n_mtiles ? mtile : n_mtiles - 1;
            const long long grow_g = (long long)mt_eff * LT_BM + row;
            const int pg = (int)(grow_g / (2 * Mhi));
            const int lhig = (int)(grow_g - (long long)pg * 2 * Mhi) >> 1;
            const uint4* trow = reinterpret_cast<const uint4*>(Ttab + (size_t)lhig * Tw) + half;
            const uint4* erow = reinterpret_cast<const uint4*>(Etab + (size_t)pg * Tw) + half;
            const uint32_t ar0 = sp_mapa(lt_smem_u32(&aready[0]), 0);
            uint4 tw0 = trow[0], ew0 = erow[0];
            for (int kb = 0; kb < nkb; ++kb) {
                const int buf = kb % TS_NBUF;
                const uint32_t tw[4] = {tw0.x, tw0.y, tw0.z, tw0.w};
                const uint32_t ew[4] = {ew0.x, ew0.y, ew0.z, ew0.w};
                if (kb + 1 < nkb) {
                    tw0 = trow[2 * (kb + 1)];
                    ew0 = erow[2 * (kb + 1)];
                }
                uint32_t av[16], ev[4];
#pragma unroll
                for (int wi = 0; wi < 4; ++wi) ts_phase_expand<false>(tw[wi], ew[wi], odd, &av[4 * wi], ev[wi]);
                lt_mbar_wait(&tfree[buf], ((uint32_t)(kb / TS_NBUF) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t ta = lane_addr + TS_TMEM_A + TS_TMEM_BUF * (uint32_t)buf;
                if (!((PROBE & 2) && kb >= TS_NBUF)) {
                    ts_st16(ta + 16u * (uint32_t)half, av);
                    ts_st4(ta + 32u + 4u * (uint32_t)half, ev);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                // relaxed: the TMEM writes are ordered by wait::st + fence::before_thread_sync; a releasing arrive would also
                // drain the prefetched global loads (measured: MEMBAR.ALL.GPU + ERRBAR = 36 % of all stall samples)
                if (lane == 0)
                    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(ar0 + 8u * (uint32_t)buf) : "memory");
            }
        }
        // epilogue (as in the dense kernel)
        const long long grow = (long long)mtile * LT_BM + row;
        const int p = (int)(grow / (2 * Mhi));
        const int lhi = (int)(grow - (long long)p * 2 * Mhi) >> 1;
        const int llo0 = ntile * LT_BN;
        const double inv_scale = (double)(*inv_scale_ptr);
        lt_mbar_wait(tfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (mtile < n_mtiles) {
            float2* orow = out + ((size_t)p * Mhi + lhi) * Nlo + llo0;
#pragma unroll 1
            for (int ch = 4 * half; ch < 4 * half + 4; ++ch) {
                uint32_t a0[16], a1[16], a2[16];
                lt_ld16(lane_addr + (uint32_t)(0 * LT_BN + ch * 16), a0);
                lt_ld16(lane_addr + (uint32_t)(1 * LT_BN + ch * 16), a1);
                lt_ld16(lane_addr + (uint32_t)(2 * LT_BN + ch * 16), a2);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float val[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const long long v = ((long long)(int)a0[j] * 128 + (long long)(int)a1[j]) * 128 + (long long)(int)a2[j];
                    val[j] = (float)((double)v * inv_scale);
                }
                float2 o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float send = odd ? val[j] : val[8 + j];
                    const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
                    o[j] = odd ? make_float2(recv, val[8 + j]) : make_float2(val[j], recv);
                }
                float4* dst = reinterpret_cast<float4*>(orow + ch * 16 + (odd ? 8 : 0));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 w = make_float4(o[2 * j].x, o[2 * j].y, o[2 * j + 1].x, o[2 * j + 1].y);
                    if (accumulate) {                           // residual pass: onto the first pass's samples
                        const float4 p = dst[j];
                        w = make_float4(p.x + w.x, p.y + w.y, p.z + w.z, p.w + w.w);
                    }
                    dst[j] = w;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    sp_cluster_sync();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int lt_make_map(CUtensorMap* map, const void* ptr, long long rows, long long kbytes, int box_k = LT_BK, int box_rows = 128,
                bool swizzle = true) {
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<EncodeTiledFn>(p);
    }
    if (!enc) {
        qsft_set_error("cuTensorMapEncodeTiled entry point not available");
        return QSFT_ECUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)kbytes, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kbytes};
    cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        qsft_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld k=%lld)", (int)r, rows, kbytes);
        return QSFT_ECUDA;
    }
    return QSFT_OK;
}

}  // namespace

// split of the b lattice digits into row (b1) and column (b2) digits of the GEMM
static inline int lt_split_b1(int q, int b) { return q == 2 ? (b - 1) / 2 : b / 2; }

extern "C" int qsft_eval_lattice_supported(int q, int n, int b, int P, int64_t S) {
    if (n < 1 || n > QSFT_MAX_N || P < 1 || S < 1 || b < 1) return 0;
    const int b1 = lt_split_b1(q, b), b2 = b - b1;
    if (q == 3) {                                            // dense Z[w] variant, ragged tiles
        if (b1 < 3 || b2 < 4 || b2 > 10 || b > 20) return 0;   // packed digits fit 32 bits; 3 * 3^b2 >= 128 rows of B'
        const long long Mhi = ipow64(3, b1), Nlo = ipow64(3, b2);
        if ((long long)P * Mhi * 2 < LT_BM || (long long)P * Mhi * 2 >= 0x7fffffffLL || Nlo > 65535) return 0;
        return 1;
    }
    if (q == 5 || q == 7) {                                  // dense Z[w] variant with d = q - 1 rows / bytes per element
        const long long Mhi = ipow64(q, b1), Nlo = ipow64(q, b2);
        if (b1 < 1 || b2 > 10 || LT_LIMBS * Nlo < LT_BN || Nlo > 65535) return 0;   // 3-bit digits fit 32 bits; B' rows >= one tile
        if ((long long)P * Mhi * (q - 1) < LT_BM || (long long)P * Mhi * (q - 1) >= 0x7fffffffLL) return 0;
        if (Mhi * (q - 1) / 2 >= 0x7fffffffLL) return 0;
        return 1;
    }
    if (q == 2) {                                            // q = 4 machinery on 2^b1 x 2^b2 lattices (lt_digits)
        if (b1 < 6 || b2 < 8 || b > 28) return 0;            // 2 * 2^b1 >= 128 rows, 2^b2 >= 256 columns
        if ((long long)P * ipow64(2, b1) * 2 >= 0x7fffffffLL) return 0;
        return 1;
    }
    if (q != 4) return 0;
    if (b1 < 3 || b2 < 4 || b > 14) return 0;                // 2 * 4^b1 >= 128 rows, 4^b2 >= 256 columns, grid.y limits
    if ((long long)P * ipow64(4, b1) * 2 >= 0x7fffffffLL) return 0;
    return 1;
}

namespace {

// q = 3 (see "q = 3" above): prep as for q = 4, A' materialised in HBM per chunk of delay rows, two dense GEMM launches per
// pass (real / imaginary parts of the strengths) into two scratch planes, then the combination X = c_1 + c_w w.
int lt_eval_q3(const int8_t* M, const int8_t* D, const int8_t* loc, const float* strengths, int64_t S, int n, int b, int P, int ld,
               float* out, int residual_passes, cudaStream_t st) {
    const int q = 3, b1 = b / 2, b2 = b - b1;
    const long long Mhi = ipow64(3, b1), Nlo = ipow64(3, b2), Bn = Mhi * Nlo;
    const long long Kp = (2 * S + LT_BK - 1) / LT_BK * LT_BK;
    double budget_gb = 8.0;
    if (const char* env = getenv("QSFT_LATTICE_SCRATCH_GB")) {
        const double v = atof(env);
        if (v > 0.0) budget_gb = v;
    }
    long long Pc = (long long)(budget_gb * 1e9 / (2.0 * (double)Mhi * (double)Kp));
    if (Pc < 1) Pc = 1;
    if (Pc > P) Pc = P;
    while ((Pc * 2 * Mhi + LT_BM - 1) / LT_BM > 65535) --Pc;
    uint32_t *hhi = nullptr, *hlo = nullptr;
    uint8_t *e = nullptr, *A = nullptr, *Bq = nullptr;
    int2* alimb = nullptr;
    unsigned int* amax = nullptr;
    float2* planes = nullptr;
    int rc = QSFT_OK;
    auto alloc = [&](void** p, size_t bytes) {
        if (rc == QSFT_OK && qsft_scratch_alloc(p, bytes, st) != cudaSuccess) {
            qsft_set_error("cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
            rc = QSFT_ECUDA;
        }
    };
    const long long Se = (S + 3) & ~3ll;
    alloc((void**)&hhi, (size_t)S * 4);
    alloc((void**)&hlo, (size_t)S * 4);
    alloc((void**)&e, (size_t)P * Se);
    alloc((void**)&alimb, (size_t)S * 8);
    alloc((void**)&amax, 16);
    alloc((void**)&A, (size_t)Pc * 2 * Mhi * Kp);
    alloc((void**)&Bq, (size_t)2 * LT_LIMBS * Nlo * Kp);                     // real-part and imaginary-part operand
    alloc((void**)&planes, (size_t)2 * Pc * Bn * sizeof(float2));
    float* inv_scale = amax ? reinterpret_cast<float*>(amax + 2) : nullptr;
    if (rc == QSFT_OK) {
        const int T = 256;
        const unsigned sb = (unsigned)((S + T - 1) / T);
        const unsigned int init[2] = {0u, 0x7f7fffffu};
        cudaMemcpyAsync(amax, init, 8, cudaMemcpyHostToDevice, st);
        lt_prep_kernel<<<sb, T, (size_t)n * b + (size_t)P * n, st>>>(M, D, loc, S, Se, n, b, b1, P, ld, hhi, hlo, e, q);
        lt_amax_kernel<<<sb, T, 0, st>>>(reinterpret_cast<const float2*>(strengths), S, amax);
        g_qsft_launches.fetch_add(2, std::memory_order_relaxed);
        int passes = 1 + (residual_passes > 0 ? 1 : 0);
        if (residual_passes < 0) {
            lt_amin_kernel<<<sb, T, 0, st>>>(reinterpret_cast<const float2*>(strengths), S, amax + 1);
            g_qsft_launches.fetch_add(1, std::memory_order_relaxed);
            float mm[2] = {0.f, 0.f};
            if (cudaMemcpyAsync(mm, amax, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
                qsft_set_error("reading the strength range failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = QSFT_ECUDA;
            }
            if (mm[1] < 0.1f * mm[0]) passes = 2;
        }
        static bool attr = false;
        if (!attr && !rc) {
            if (cudaFuncSetAttribute(lt_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LT_SMEM) != cudaSuccess) {
                qsft_set_error("cudaFuncSetAttribute failed");
                rc = QSFT_ECUDA;
            }
            attr = true;
        }
        const unsigned pb = (unsigned)((Kp / 4 + T - 1) / T);
        uint8_t* Bre = Bq;
        uint8_t* Bim = Bq + (size_t)LT_LIMBS * Nlo * Kp;
        CUtensorMap ma, mbr, mbi;
        if (!rc) rc = lt_make_map(&mbr, Bre, LT_LIMBS * Nlo, Kp);
        if (!rc) rc = lt_make_map(&mbi, Bim, LT_LIMBS * Nlo, Kp);
        for (long long p0 = 0; p0 < P && !rc; p0 += Pc) {
            const long long pc = (P - p0 < Pc) ? (P - p0) : Pc;
            const long long rows = pc * 2 * Mhi;
            lt_agen3_kernel<<<dim3(pb, (unsigned)Mhi), T, 0, st>>>(hhi, e + (size_t)p0 * Se, S, Se, b1, (int)pc, Mhi, Kp,
                                                                   reinterpret_cast<uint32_t*>(A));
            g_qsft_launches.fetch_add(1, std::memory_order_relaxed);
            rc = lt_make_map(&ma, A, rows, Kp);
            if (rc) break;
            float2* cre = planes;
            float2* cim = planes + (size_t)pc * Bn;
            for (int pass = 0; pass < passes && !rc; ++pass) {
                lt_quant_kernel<<<sb, T, 0, st>>>(reinterpret_cast<const float2*>(strengths), S, amax, inv_scale, alimb, pass);
                lt_bgen3_kernel<<<dim3(pb, (unsigned)Nlo), T, 0, st>>>(hlo, alimb, S, b2, Nlo, Kp, 0, reinterpret_cast<uint32_t*>(Bre));
                lt_bgen3_kernel<<<dim3(pb, (unsigned)Nlo), T, 0, st>>>(hlo, alimb, S, b2, Nlo, Kp, 1, reinterpret_cast<uint32_t*>(Bim));
                dim3 grid((unsigned)((Nlo + LT_BN - 1) / LT_BN), (unsigned)((rows + LT_BM - 1) / LT_BM));
                lt_gemm_kernel<true><<<grid, LT_THREADS, LT_SMEM, st>>>(ma, mbr, (int)(Kp / LT_BK), (int)Mhi, (int)Nlo,
                                                                          (const float*)(inv_scale + pass), cre, pass, rows);
                lt_gemm_kernel<true><<<grid, LT_THREADS, LT_SMEM, st>>>(ma, mbi, (int)(Kp / LT_BK), (int)Mhi, (int)Nlo,
                                                                          (const float*)(inv_scale + pass), cim, pass, rows);
                g_qsft_launches.fetch_add(5, std::memory_order_relaxed);
            }
            const long long N = pc * Bn;
            lt_combine3_kernel<<<(unsigned)((N + T - 1) / T), T, 0, st>>>(cre, cim, N, reinterpret_cast<float2*>(out) + (size_t)p0 * Bn);
            g_qsft_launches.fetch_add(1, std::memory_order_relaxed);
            cudaError_t ce = cudaGetLastError();
            if (ce != cudaSuccess) {
                qsft_set_error("lattice GEMM (q = 3) launch failed: %s", cudaGetErrorString(ce));
                rc = QSFT_ECUDA;
            }
        }
    }
    void* frees[] = {hhi, hlo, e, alimb, amax, A, Bq, planes};
    for (void* p : frees)
        if (p) cudaFreeAsync(p, st);
    return rc;
}

// odd primes q = 5, 7 ("odd primes" above): as lt_eval_q3 with d = q - 1 rows per l_hi and d bytes per support element
template <int Q>
int lt_eval_qodd(const int8_t* M, const int8_t* D, const int8_t* loc, const float* strengths, int64_t S, int n, int b, int P, int ld,
               float* out, int residual_passes, cudaStream_t st) {
    constexpr int DQ = LtQ<Q>::D;
    const int q = Q, b1 = b / 2, b2 = b - b1;
    const long long Mhi = ipow64(Q, b1), Nlo = ipow64(Q, b2), Bn = Mhi * Nlo;
    const long long Kp = ((long long)DQ * S + LT_BK - 1) / LT_BK * LT_BK;
    double budget_gb = 8.0;
    if (const char* env = getenv("QSFT_LATTICE_SCRATCH_GB")) {
        const double v = atof(env);
        if (v > 0.0) budget_gb = v;
    }
    long long Pc = (long long)(budget_gb * 1e9 / ((double)DQ * (double)Mhi * (double)Kp + 8.0 * (double)DQ * (double)Bn));
    if (Pc < 1) Pc = 1;
    if (Pc > P) Pc = P;
    while ((Pc * DQ * Mhi + LT_BM - 1) / LT_BM > 65535) --Pc;
    uint32_t *hhi = nullptr, *hlo = nullptr;
    uint8_t *e = nullptr, *A = nullptr, *Bq = nullptr;
    int2* alimb = nullptr;
    unsigned int* amax = nullptr;
    float2* planes = nullptr;
    int rc = QSFT_OK;
    auto alloc = [&](void** p, size_t bytes) {
        if (rc == QSFT_OK && qsft_scratch_alloc(p, bytes, st) != cudaSuccess) {
            qsft_set_error("cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
            rc = QSFT_ECUDA;
        }
    };
    const long long Se = (S + 3) & ~3ll;
    alloc((void**)&hhi, (size_t)S * 4);
    alloc((void**)&hlo, (size_t)S * 4);
    alloc((void**)&e, (size_t)P * Se);
    alloc((void**)&alimb, (size_t)S * 8);
    alloc((void**)&amax, 16);
    alloc((void**)&A, (size_t)Pc * DQ * Mhi * Kp);
    alloc((void**)&Bq, (size_t)2 * LT_LIMBS * Nlo * Kp);                     // real-part and imaginary-part operand
    alloc((void**)&planes, (size_t)2 * Pc * Bn * (DQ / 2) * sizeof(float2));
    float* inv_scale = amax ? reinterpret_cast<float*>(amax + 2) : nullptr;
    if (rc == QSFT_OK) {
        const int T = 256;
        const unsigned sb = (unsigned)((S + T - 1) / T);
        const unsigned int init[2] = {0u, 0x7f7fffffu};
        cudaMemcpyAsync(amax, init, 8, cudaMemcpyHostToDevice, st);
        lt_prep_kernel<<<sb, T, (size_t)n * b + (size_t)P * n, st>>>(M, D, loc, S, Se, n, b, b1, P, ld, hhi, hlo, e, q, LtQ<Q>::FW);
        lt_amax_kernel<<<sb, T, 0, st>>>(reinterpret_cast<const float2*>(strengths), S, amax);
        g_qsft_launches.fetch_add(2, std::memory_order_relaxed);
        int passes = 1 + (residual_passes > 0 ? 1 : 0);
        if (residual_passes < 0) {
            lt_amin_kernel<<<sb, T, 0, st>>>(reinterpret_cast<const float2*>(strengths), S, amax + 1);
            g_qsft_launches.fetch_add(1, std::memory_order_relaxed);
            float mm[2] = {0.f, 0.f};
            if (cudaMemcpyAsync(mm, amax, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
                qsft_set_error("reading the strength range failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = QSFT_ECUDA;
            }
            if (mm[1] < 0.1f * mm[0]) passes = 2;
        }
        static bool attr = false;   // (one flag per instantiation; setting the attribute twice is harmless)
        if (!attr && !rc) {
            if (cudaFuncSetAttribute(lt_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LT_SMEM) != cudaSuccess) {
                qsft_set_error("cudaFuncSetAttribute failed");
                rc = QSFT_ECUDA;
            }
            attr = true;
        }
        const unsigned pb = (unsigned)((Kp / 4 + T - 1) / T);
        uint8_t* Bre = Bq;
        uint8_t* Bim = Bq + (size_t)LT_LIMBS * Nlo * Kp;
        CUtensorMap ma, mbr, mbi;
        if (!rc) rc = lt_make_map(&mbr, Bre, LT_LIMBS * Nlo, Kp);
        if (!rc) rc = lt_make_map(&mbi, Bim, LT_LIMBS * Nlo, Kp);
        for (long long p0 = 0; p0 < P && !rc; p0 += Pc) {
            const long long pc = (P - p0 < Pc) ? (P - p0) : Pc;
            const long long rows = pc * DQ * Mhi;
            lt_agenq_kernel<Q><<<dim3(pb, (unsigned)Mhi), T, 0, st>>>(hhi, e + (size_t)p0 * Se, S, Se, b1, (int)pc, Mhi, Kp,
                                                                   reinterpret_cast<uint32_t*>(A));
            g_qsft_launches.fetch_add(1, std::memory_order_relaxed);
            rc = lt_make_map(&ma, A, rows, Kp);
            if (rc) break;
            float2* cre = planes;
            float2* cim = planes + (size_t)pc * Bn * (DQ / 2);
            for (int pass = 0; pass < passes && !rc; ++pass) {
                lt_quant_kernel<<<sb, T, 0, st>>>(reinterpret_cast<const float2*>(strengths), S, amax, inv_scale, alimb, pass);
                lt_bgenq_kernel<Q><<<dim3(pb, (unsigned)Nlo), T, 0, st>>>(hlo, alimb, S, b2, Nlo, Kp, 0, reinterpret_cast<uint32_t*>(Bre));
                lt_bgenq_kernel<Q><<<dim3(pb, (unsigned)Nlo), T, 0, st>>>(hlo, alimb, S, b2, Nlo, Kp, 1, reinterpret_cast<uint32_t*>(Bim));
                dim3 grid((unsigned)((Nlo + LT_BN - 1) / LT_BN), (unsigned)((rows + LT_BM - 1) / LT_BM));
                lt_gemm_kernel<true><<<grid, LT_THREADS, LT_SMEM, st>>>(ma, mbr, (int)(Kp / LT_BK), (int)(Mhi * (DQ / 2)), (int)Nlo,
                                                                          (const float*)(inv_scale + pass), cre, pass, rows);
                lt_gemm_kernel<true><<<grid, LT_THREADS, LT_SMEM, st>>>(ma, mbi, (int)(Kp / LT_BK), (int)(Mhi * (DQ / 2)), (int)Nlo,
                                                                          (const float*)(inv_scale + pass), cim, pass, rows);
                g_qsft_launches.fetch_add(5, std::memory_order_relaxed);
            }
            const long long N = pc * Bn;
            lt_combineq_kernel<Q><<<(unsigned)((N + T - 1) / T), T, 0, st>>>(cre, cim, pc * Mhi, Nlo, reinterpret_cast<float2*>(out) + (size_t)p0 * Bn);
            g_qsft_launches.fetch_add(1, std::memory_order_relaxed);
            cudaError_t ce = cudaGetLastError();
            if (ce != cudaSuccess) {
                qsft_set_error("lattice GEMM (odd prime q) launch failed: %s", cudaGetErrorString(ce));
                rc = QSFT_ECUDA;
            }
        }
    }
    void* frees[] = {hhi, hlo, e, alimb, amax, A, Bq, planes};
    for (void* p : frees)
        if (p) cudaFreeAsync(p, st);
    return rc;
}


}  // namespace

// residual_passes: 0 = one GEMM pass (20 bits below max|a|), 1 = a second pass over the quantisation residual accumulated
// onto the first (41 bits; twice the tensor work), -1 = decide here: second pass iff min|a| < 0.1 max|a| (reads two floats
// back, i.e. synchronises the stream once).
extern "C" int qsft_eval_synth_lattice_ex(const int8_t* M, const int8_t* D, const int8_t* loc, const float* strengths, int64_t S,
                                          int q, int n, int b, int P, int ld, float* out, int residual_passes, void* stream) {
    QSFT_CHECK_ARG(qsft_eval_lattice_supported(q, n, b, P, S), "lattice evaluation supports q = 4 (7 <= b <= 14), q = 3 (7 <= b <= 20), q = 2 (14 <= b <= 28) and q = 5 / 7 (see qsft_eval_lattice_supported) only");
    QSFT_CHECK_ARG(M && D && loc && strengths && out, "null pointer");
    QSFT_CHECK_ARG(ld >= n && ld % 16 == 0, "bad ld");
    QSFT_CHECK_ARG(residual_passes >= -1 && residual_passes <= 1, "residual_passes must be -1 (auto), 0 or 1");
    cudaStream_t st = (cudaStream_t)stream;
    if (q == 3) return lt_eval_q3(M, D, loc, strengths, S, n, b, P, ld, out, residual_passes, st);
    if (q == 5) return lt_eval_qodd<5>(M, D, loc, strengths, S, n, b, P, ld, out, residual_passes, st);
    if (q == 7) return lt_eval_qodd<7>(M, D, loc, strengths, S, n, b, P, ld, out, residual_passes, st);
    const int b1 = lt_split_b1(q, b), b2 = b - b1;
    const int spread = (q == 2) ? 1 : 0;                 // q = 2: one bit per lattice digit (lt_digits)
    const long long Mhi = ipow64(q, b1), Nlo = ipow64(q, b2);
    // Default: 2:4 structured-sparse A' generated straight into tensor memory (tcgen05.mma.sp on CTA pairs; A' never exists
    // in HBM or shared memory).  QSFT_LATTICE_SPARSE=0 selects the dense cross-check kernel: A' materialised in HBM
    // (2 * Mhi * Kp bytes per delay row), delay rows processed in chunks that keep it under a scratch budget (default 32 GB,
    // QSFT_LATTICE_SCRATCH_GB overrides).  Both are exact integer arithmetic and produce the same bits.
    bool sparse_ts = true;
    if (const char* env = getenv("QSFT_LATTICE_SPARSE")) sparse_ts = atoi(env) != 0;
    const long long kalign = sparse_ts ? SP_BK : LT_BK;
    const long long Kp = (2 * S + kalign - 1) / kalign * kalign;
    double budget_gb = 32.0;
    if (const char* env = getenv("QSFT_LATTICE_SCRATCH_GB")) {
        const double v = atof(env);
        if (v > 0.0) budget_gb = v;
    }
    const double per_row = 2.0 * (double)Mhi * (double)Kp;
    long long Pc = sparse_ts ? P : (long long)(budget_gb * 1e9 / per_row);
    if (Pc < 1) Pc = 1;
    if (Pc > P) Pc = P;
    while (Pc * 2 * Mhi / LT_BM > 65535) --Pc;          // grid.y limit
    const long long Tw = Kp / 32;                       // packed phase words (16 support elements each) per row
    // stream-ordered workspace
    uint32_t *hhi = nullptr, *hlo = nullptr, *Ttab = nullptr, *Etab = nullptr;
    uint8_t* e = nullptr;
    int2* alimb = nullptr;
    unsigned int* amax = nullptr;
    uint8_t *A = nullptr, *Bq = nullptr;
    int rc = QSFT_OK;
    auto alloc = [&](void** p, size_t bytes) {
        if (rc == QSFT_OK && qsft_scratch_alloc(p, bytes, st) != cudaSuccess) {
            qsft_set_error("cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
            rc = QSFT_ECUDA;
        }
    };
    alloc((void**)&hhi, (size_t)S * 4);
    alloc((void**)&hlo, (size_t)S * 4);
    const long long Se = (S + 3) & ~3ll;     // even row stride so that e[p][s0], e[p][s0+1] is one aligned 16-bit load
    alloc((void**)&e, (size_t)P * Se);
    alloc((void**)&alimb, (size_t)S * 8);
    alloc((void**)&amax, 16);                // [0] max, [1] min (bit patterns), [2..3] inv_scale of pass 0 / 1
    if (sparse_ts) {
        alloc((void**)&Ttab, (size_t)Mhi * Tw * 4);
        alloc((void**)&Etab, (size_t)P * Tw * 4);
    } else {
        alloc((void**)&A, (size_t)Pc * 2 * Mhi * Kp);
    }
    alloc((void**)&Bq, (size_t)LT_LIMBS * Nlo * Kp);
    float* inv_scale = amax ? reinterpret_cast<float*>(amax + 2) : nullptr;
    if (rc == QSFT_OK) {
        const int T = 256;
        const unsigned sb = (unsigned)((S + T - 1) / T);
        const unsigned int init[2] = {0u, 0x7f7fffffu};
        cudaMemcpyAsync(amax, init, 8, cudaMemcpyHostToDevice, st);
        lt_prep_kernel<<<sb, T, (size_t)n * b + (size_t)P * n, st>>>(M, D, loc, S, Se, n, b, b1, P, ld, hhi, hlo, e, q);
        lt_amax_kernel<<<sb, T, 0, st>>>(reinterpret_cast<const float2*>(strengths), S, amax);
        g_qsft_launches.fetch_add(2, std::memory_order_relaxed);
        int passes = 1 + (residual_passes > 0 ? 1 : 0);
        if (residual_passes < 0) {
            lt_amin_kernel<<<sb, T, 0, st>>>(reinterpret_cast<const float2*>(strengths), S, amax + 1);
            g_qsft_launches.fetch_add(1, std::memory_order_relaxed);
            float mm[2] = {0.f, 0.f};
            if (cudaMemcpyAsync(mm, amax, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
                qsft_set_error("reading the strength range failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = QSFT_ECUDA;
            }
            if (mm[1] < 0.1f * mm[0]) passes = 2;
        }
        const unsigned pb = (unsigned)((Kp / 4 + T - 1) / T);
        const unsigned wb = (unsigned)((Tw + T - 1) / T);
        if (sparse_ts && !rc) {
            lt_ttab_kernel<<<dim3(wb, (unsigned)((Mhi + LT_TTAB_LL - 1) / LT_TTAB_LL)), T, 0, st>>>(hhi, S, b1, Tw, Mhi, Ttab, spread);
            lt_etab_kernel<<<dim3(wb, (unsigned)P), T, 0, st>>>(e, S, Se, P, Tw, Etab);
            g_qsft_launches.fetch_add(2, std::memory_order_relaxed);
        }
        CUtensorMap ma, mb, mb2;
        if (!rc) rc = lt_make_map(&mb, Bq, LT_LIMBS * Nlo, Kp);
        if (!rc && sparse_ts) rc = lt_make_map(&mb2, Bq, LT_LIMBS * Nlo, Kp, LT_BK, 64);
        ma = mb;   // placeholder when A' is generated in the kernel
        if (!rc) {
            static bool attr = false;
            if (!attr) {
                if (cudaFuncSetAttribute(lt_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LT_SMEM) != cudaSuccess ||
                    cudaFuncSetAttribute(lt_gemm_spts_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_SMEM) != cudaSuccess ||
                    cudaFuncSetAttribute(lt_gemm_spts_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_SMEM) != cudaSuccess ||
                    cudaFuncSetAttribute(lt_gemm_spts_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_SMEM) != cudaSuccess ||
                    cudaFuncSetAttribute(lt_gemm_spts_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_SMEM) != cudaSuccess) {
                    qsft_set_error("cudaFuncSetAttribute failed");
                    rc = QSFT_ECUDA;
                }
                attr = true;
            }
        }
        for (int pass = 0; pass < passes && !rc; ++pass) {
            // the limb operand B' of this pass (shared by all delay rows)
            lt_quant_kernel<<<sb, T, 0, st>>>(reinterpret_cast<const float2*>(strengths), S, amax, inv_scale, alimb, pass);
            lt_bgen_kernel<<<dim3(pb, (unsigned)((Nlo + LT_BGEN_LL - 1) / LT_BGEN_LL)), T, 0, st>>>(hlo, alimb, S, b2, Nlo, Kp,
                                                                                                     reinterpret_cast<uint32_t*>(Bq), spread);
            g_qsft_launches.fetch_add(2, std::memory_order_relaxed);
            for (long long p0 = 0; p0 < P && !rc; p0 += Pc) {
                const long long pc = (P - p0 < Pc) ? (P - p0) : Pc;
                float2* o = reinterpret_cast<float2*>(out) + (size_t)p0 * Mhi * Nlo;
                if (sparse_ts) {
                    const long long n_mt = pc * 2 * Mhi / LT_BM, n_mt_pad = (n_mt + 1) & ~1ll;
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3((unsigned)n_mt_pad, (unsigned)(Nlo / LT_BN));
                    cfg.blockDim = dim3(TS_THREADS);
                    cfg.dynamicSmemBytes = TS_SMEM;
                    cfg.stream = st;
                    cudaLaunchAttribute cattr[1];
                    cattr[0].id = cudaLaunchAttributeClusterDimension;
                    cattr[0].val.clusterDim.x = 2;
                    cattr[0].val.clusterDim.y = 1;
                    cattr[0].val.clusterDim.z = 1;
                    cfg.attrs = cattr;
                    cfg.numAttrs = 1;
                    int probe = 0;
                    if (const char* env = getenv("QSFT_LT_PROBE")) probe = atoi(env) & 3;
                    auto kern = probe == 0 ? lt_gemm_spts_kernel<0> : probe == 1 ? lt_gemm_spts_kernel<1>
                              : probe == 2 ? lt_gemm_spts_kernel<2> : lt_gemm_spts_kernel<3>;
                    cudaLaunchKernelEx(&cfg, kern, mb, mb2, (int)(Kp / SP_BK), (int)Mhi, (int)Nlo, (int)n_mt,
                                       (const uint32_t*)Ttab, (const uint32_t*)(Etab + (size_t)p0 * Tw), (int)Tw,
                                       (const float*)(inv_scale + pass), o, pass);
                    g_qsft_launches.fetch_add(1, std::memory_order_relaxed);
                } else {
                    dim3 grid((unsigned)(Nlo / LT_BN), (unsigned)(pc * 2 * Mhi / LT_BM));
                    lt_agen_kernel<<<dim3(pb, (unsigned)Mhi), T, 0, st>>>(hhi, e + (size_t)p0 * Se, S, Se, b1, (int)pc, Mhi, Kp,
                                                                          reinterpret_cast<uint32_t*>(A), spread);
                    rc = lt_make_map(&ma, A, pc * 2 * Mhi, Kp);
                    if (rc) break;
                    lt_gemm_kernel<false><<<grid, LT_THREADS, LT_SMEM, st>>>(ma, mb, (int)(Kp / LT_BK), (int)Mhi, (int)Nlo,
                                                                             (const float*)(inv_scale + pass), o, pass, pc * 2 * Mhi);
                    g_qsft_launches.fetch_add(2, std::memory_order_relaxed);
                }
                cudaError_t ce = cudaGetLastError();
                if (ce != cudaSuccess) {
                    qsft_set_error("lattice GEMM launch failed: %s", cudaGetErrorString(ce));
                    rc = QSFT_ECUDA;
                }
            }
        }
    }
    void* frees[] = {hhi, hlo, e, alimb, amax, A, Bq, Ttab, Etab};
    for (void* p : frees)
        if (p) cudaFreeAsync(p, st);
    return rc;
}

extern "C" int qsft_eval_synth_lattice(const int8_t* M, const int8_t* D, const int8_t* loc, const float* strengths,
                                       int64_t S, int q, int n, int b, int P, int ld, float* out, void* stream) {
    return qsft_eval_synth_lattice_ex(M, D, loc, strengths, S, q, n, b, P, ld, out, -1, stream);
}
