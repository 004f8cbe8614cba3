// K1: query lattice (M l + d) mod q -> decimal index / digit rows, plus the index codecs.
// Replaces qsft/input_signal_subsampled.py:183-206 and qsft/utils.py:74-84,107-108.  HBM-write bound.
#include "common.cuh"

namespace {

constexpr int K1_THREADS = 128;

// 128-bit Horner step: acc = acc * q + d
template <int LIMBS>
__device__ __forceinline__ void horner(uint64_t& hi, uint64_t& lo, uint32_t q, uint32_t d) {
    if (LIMBS == 2) {
        uint64_t carry = __umul64hi(lo, (uint64_t)q);
        hi = hi * q + carry;
    }
    lo = lo * q + d;
    if (LIMBS == 2) hi += (lo < (uint64_t)d) ? 1 : 0;
}

// Cooperative, coalesced copy of `rows` staged digit rows (row stride `ldw+1` words in smem) to global memory
// where the rows are contiguous (row stride ldw words).
__device__ __forceinline__ void flush_rows(const uint32_t* s, uint32_t* g, int rows, int ldw) {
    const int total = rows * ldw;
    for (int w = threadIdx.x; w < total; w += blockDim.x) {
        int r = w / ldw, c = w - r * ldw;
        g[w] = s[r * (ldw + 1) + c];
    }
}

// grid.x: tiles of K1_THREADS lattice points, grid.y: chunks of delay rows
template <int LIMBS>
__global__ void __launch_bounds__(K1_THREADS)
k1_lattice_kernel(const int8_t* __restrict__ M, const int8_t* __restrict__ D, int q, int n, int b, int P,
                  int p_per_block, long long B, uint64_t* __restrict__ out_idx, int8_t* __restrict__ out_dig, int ld) {
    extern __shared__ uint32_t smem_u32[];
    const int ldw = ld / 4;
    // layout: sML[n][K1_THREADS] bytes | sD[p_per_block * n] bytes | sM[n*b] bytes | sOut[K1_THREADS][ldw+1] words
    uint8_t* sML = reinterpret_cast<uint8_t*>(smem_u32);
    uint8_t* sD = sML + (size_t)n * K1_THREADS;
    uint8_t* sM = sD + (size_t)p_per_block * n;
    size_t off = ((size_t)n * K1_THREADS + (size_t)p_per_block * n + (size_t)n * b + 3) & ~(size_t)3;
    uint32_t* sOut = reinterpret_cast<uint32_t*>(sML + off);

    const int tid = threadIdx.x;
    const long long l0 = (long long)blockIdx.x * K1_THREADS;
    const long long l = l0 + tid;
    const int p0 = blockIdx.y * p_per_block;
    const int p1 = min(P, p0 + p_per_block);
    const int rows = (int)min((long long)K1_THREADS, B - l0);

    for (int i = tid; i < n * b; i += K1_THREADS) sM[i] = (uint8_t)M[i];
    for (int i = tid; i < (p1 - p0) * n; i += K1_THREADS) sD[i] = (uint8_t)D[(size_t)p0 * n + i];
    __syncthreads();

    if (l < B) {
        // digits of l, MSB first (itertools.product order, utils.py:107-108)
        uint8_t ldig[QSFT_MAX_B];
        long long v = l;
        for (int j = b - 1; j >= 0; --j) {
            ldig[j] = (uint8_t)(v % q);
            v /= q;
        }
        for (int i = 0; i < n; ++i) {
            uint32_t acc = 0;
            for (int j = 0; j < b; ++j) acc += (uint32_t)sM[i * b + j] * ldig[j];
            sML[i * K1_THREADS + tid] = (uint8_t)(acc % q);
        }
    }
    // no sync needed: each thread only reads its own sML column

    for (int p = p0; p < p1; ++p) {
        if (l < B) {
            uint64_t hi = 0, lo = 0;
            uint32_t word = 0;
            const uint8_t* d = sD + (size_t)(p - p0) * n;
            for (int i = 0; i < n; ++i) {
                uint32_t dg = (uint32_t)sML[i * K1_THREADS + tid] + d[i];
                dg = dg >= (uint32_t)q ? dg - q : dg;
                horner<LIMBS>(hi, lo, (uint32_t)q, dg);
                if (out_dig) {
                    word |= dg << (8 * (i & 3));
                    if ((i & 3) == 3) {
                        sOut[tid * (ldw + 1) + (i >> 2)] = word;
                        word = 0;
                    }
                }
            }
            if (out_dig) {
                if (n & 3) sOut[tid * (ldw + 1) + (n >> 2)] = word;
                for (int w = (n + 3) / 4; w < ldw; ++w) sOut[tid * (ldw + 1) + w] = 0;
            }
            if (out_idx) {
                uint64_t* o = out_idx + ((size_t)p * B + l) * LIMBS;
                if (LIMBS == 2) {
                    *reinterpret_cast<ulonglong2*>(o) = make_ulonglong2(hi, lo);
                } else {
                    o[0] = lo;
                }
            }
        }
        if (out_dig) {
            __syncthreads();
            flush_rows(sOut, reinterpret_cast<uint32_t*>(out_dig + ((size_t)p * B + l0) * ld), rows, ldw);
            __syncthreads();
        }
    }
}

// ---- q = 2^w fast path ---------------------------------------------------------------------------------
// Digits are w-bit fields of the 128-bit index itself (MSB first: digit i sits at bit w (n-1-i)), so the whole
// lattice point is bit arithmetic: M l = sum_j l_j * Mcol_j and + d_p are field-wise adds mod 2^w (SWAR, no carries
// between fields), and the decimal index IS the packed value.  ~10 integer ops per index instead of ~40 per digit.
struct U128 {
    uint64_t hi, lo;
};

__device__ __forceinline__ U128 field_add(U128 a, U128 b, U128 H) {
    // per-field (a + b) mod 2^w: add without the field's top bit, then xor the top bits back in
    U128 r;
    r.lo = ((a.lo & ~H.lo) + (b.lo & ~H.lo)) ^ ((a.lo ^ b.lo) & H.lo);
    r.hi = ((a.hi & ~H.hi) + (b.hi & ~H.hi)) ^ ((a.hi ^ b.hi) & H.hi);
    return r;
}

__device__ __forceinline__ uint32_t field_get(U128 v, int pos, uint32_t mask) {   // bits [pos, pos + w)
    // w divides 64 on this path, so a field never straddles the two words
    return (uint32_t)((pos >= 64) ? (v.hi >> (pos - 64)) : (v.lo >> pos)) & mask;
}

__global__ void __launch_bounds__(K1_THREADS)
k1_lattice_pow2_kernel(const int8_t* __restrict__ M, const int8_t* __restrict__ D, int w, int n, int b, int P,
                       int p_per_block, long long B, int limbs, uint64_t* __restrict__ out_idx,
                       int8_t* __restrict__ out_dig, int ld) {
    extern __shared__ uint32_t smem_u32[];
    const int ldw = ld / 4;
    U128* sMc = reinterpret_cast<U128*>(smem_u32);          // [b] packed columns of M
    U128* sD = sMc + b;                                      // [p_per_block] packed delay rows
    uint32_t* sOut = reinterpret_cast<uint32_t*>(sD + p_per_block);   // [K1_THREADS][ldw + 1]
    const int tid = threadIdx.x;
    const long long l0 = (long long)blockIdx.x * K1_THREADS;
    const long long l = l0 + tid;
    const int p0 = blockIdx.y * p_per_block;
    const int p1 = min(P, p0 + p_per_block);
    const int rows = (int)min((long long)K1_THREADS, B - l0);
    const uint32_t q1 = (1u << w) - 1;

    auto pack = [&](auto digit_of) {
        U128 v{0, 0};
        for (int i = 0; i < n; ++i) {
            const int pos = w * (n - 1 - i);
            const uint64_t dg = (uint64_t)(digit_of(i) & q1);
            if (pos >= 64) v.hi |= dg << (pos - 64);
            else v.lo |= dg << pos;
        }
        return v;
    };
    for (int j = tid; j < b; j += K1_THREADS) sMc[j] = pack([&](int i) { return (uint32_t)M[i * b + j]; });
    for (int p = p0 + tid; p < p1; p += K1_THREADS) sD[p - p0] = pack([&](int i) { return (uint32_t)D[(size_t)p * n + i]; });
    __syncthreads();
    // top bit of every field
    U128 H{0, 0};
    for (int i = 0; i < n; ++i) {
        const int pos = w * (n - 1 - i) + (w - 1);
        if (pos >= 64) H.hi |= 1ull << (pos - 64); else H.lo |= 1ull << pos;
    }
    U128 ml{0, 0};
    if (l < B) {
        long long v = l;
        for (int j = b - 1; j >= 0; --j) {          // l_j, MSB first (itertools.product order)
            const int lj = (int)(v & q1);
            v >>= w;
            const U128 col = sMc[j];
            for (int t = 0; t < lj; ++t) ml = field_add(ml, col, H);
        }
    }
    for (int p = p0; p < p1; ++p) {
        if (l < B) {
            const U128 idx = field_add(ml, sD[p - p0], H);
            if (out_idx) {
                uint64_t* o = out_idx + ((size_t)p * B + l) * limbs;
                if (limbs == 2) *reinterpret_cast<ulonglong2*>(o) = make_ulonglong2(idx.hi, idx.lo);
                else o[0] = idx.lo;
            }
            if (out_dig) {
                uint32_t* row = sOut + tid * (ldw + 1);
                for (int wd = 0; wd < ldw; ++wd) {
                    uint32_t word = 0;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int i = 4 * wd + t;
                        if (i < n) word |= field_get(idx, w * (n - 1 - i), q1) << (8 * t);
                    }
                    row[wd] = word;
                }
            }
        }
        if (out_dig) {
            __syncthreads();
            flush_rows(sOut, reinterpret_cast<uint32_t*>(out_dig + ((size_t)p * B + l0) * ld), rows, ldw);
            __syncthreads();
        }
    }
}

// ---- codecs ------------------------------------------------------------------------------------------
constexpr int CODEC_THREADS = 128;

template <int LIMBS>
__global__ void __launch_bounds__(CODEC_THREADS)
dec_to_qary_kernel(const uint64_t* __restrict__ idx, long long N, int q, int n, int g, uint32_t Qg,
                   int8_t* __restrict__ dig, int ld) {
    extern __shared__ uint32_t smem_u32[];
    const int ldw = ld / 4;
    uint8_t* sOut = reinterpret_cast<uint8_t*>(smem_u32);  // rows of (ldw+1)*4 bytes
    const int tid = threadIdx.x;
    const long long r0 = (long long)blockIdx.x * CODEC_THREADS;
    const long long r = r0 + tid;
    const int rows = (int)min((long long)CODEC_THREADS, N - r0);
    uint8_t* row = sOut + (size_t)tid * (ldw + 1) * 4;
    if (r < N) {
        for (int w = 0; w <= ldw; ++w) reinterpret_cast<uint32_t*>(row)[w] = 0;
        uint32_t v[4] = {0, 0, 0, 0};  // little endian 32-bit limbs
        if (LIMBS == 2) {
            uint64_t hi = idx[r * 2], lo = idx[r * 2 + 1];
            v[0] = (uint32_t)lo; v[1] = (uint32_t)(lo >> 32); v[2] = (uint32_t)hi; v[3] = (uint32_t)(hi >> 32);
        } else {
            uint64_t lo = idx[r];
            v[0] = (uint32_t)lo; v[1] = (uint32_t)(lo >> 32);
        }
        int i = n - 1;
        while (i >= 0) {
            // divide the 128-bit value by Qg = q^g, remainder -> g digits
            uint64_t rem = 0;
#pragma unroll
            for (int k = 2 * LIMBS - 1; k >= 0; --k) {
                uint64_t cur = (rem << 32) | v[k];
                v[k] = (uint32_t)(cur / Qg);
                rem = cur - (uint64_t)v[k] * Qg;
            }
            uint32_t c = (uint32_t)rem;
            for (int t = 0; t < g && i >= 0; ++t, --i) {
                uint32_t dq = c / (uint32_t)q;
                row[i] = (uint8_t)(c - dq * q);
                c = dq;
            }
        }
    }
    __syncthreads();
    flush_rows(reinterpret_cast<uint32_t*>(sOut), reinterpret_cast<uint32_t*>(dig + (size_t)r0 * ld), rows, ldw);
}

template <int LIMBS>
__global__ void __launch_bounds__(CODEC_THREADS)
qary_to_dec_kernel(const int8_t* __restrict__ dig, int ld, long long N, int q, int n, uint64_t* __restrict__ idx) {
    const long long r = (long long)blockIdx.x * CODEC_THREADS + threadIdx.x;
    if (r >= N) return;
    const uint32_t* row = reinterpret_cast<const uint32_t*>(dig + (size_t)r * ld);
    uint64_t hi = 0, lo = 0;
    for (int w = 0; w < (n + 3) / 4; ++w) {
        uint32_t word = row[w];
        for (int t = 0; t < 4 && w * 4 + t < n; ++t) horner<LIMBS>(hi, lo, (uint32_t)q, (word >> (8 * t)) & 0xff);
    }
    if (LIMBS == 2) {
        *reinterpret_cast<ulonglong2*>(idx + r * 2) = make_ulonglong2(hi, lo);
    } else {
        idx[r] = lo;
    }
}

int check_qn(int q, int n, int limbs, int ld, bool need_ld) {
    QSFT_CHECK_ARG(q >= 2 && q <= QSFT_MAX_Q, "q=%d out of range [2,%d]", q, QSFT_MAX_Q);
    QSFT_CHECK_ARG(n >= 1 && n <= QSFT_MAX_N, "n=%d out of range [1,%d]", n, QSFT_MAX_N);
    QSFT_CHECK_ARG(limbs == 1 || limbs == 2, "limbs must be 1 or 2");
    QSFT_CHECK_ARG(index_fits(q, n, limbs), "q^n (q=%d, n=%d) does not fit in %d x 64 bits", q, n, limbs);
    if (need_ld) QSFT_CHECK_ARG(ld >= n && ld % 16 == 0, "ld=%d must be >= n and a multiple of 16", ld);
    return QSFT_OK;
}

}  // namespace

extern "C" int qsft_query_lattice(const int8_t* M, const int8_t* D, int q, int n, int b, int P, uint64_t* out_idx,
                                  int limbs, int8_t* out_dig, int ld, void* stream) {
    if (int rc = check_qn(q, n, out_idx ? limbs : 2, ld, out_dig != nullptr)) return rc;
    QSFT_CHECK_ARG(b >= 1 && b <= QSFT_MAX_B && b <= n + 64, "b=%d out of range", b);
    QSFT_CHECK_ARG(P >= 1, "P must be >= 1");
    QSFT_CHECK_ARG(M && D && (out_idx || out_dig), "null pointer");
    double Bd = 1;
    for (int i = 0; i < b; ++i) Bd *= q;
    QSFT_CHECK_ARG(Bd <= 9e15, "q^b too large");
    const long long B = ipow64(q, b);
    const long long tiles = (B + K1_THREADS - 1) / K1_THREADS;
    // enough blocks to fill the GPU: split the delay rows when there are few lattice tiles
    int p_chunks = 1;
    const long long want = 4LL * qsft_num_sms();
    if (tiles < want) p_chunks = (int)min((long long)P, (want + tiles - 1) / tiles);
    int p_per_block = (P + p_chunks - 1) / p_chunks;
    p_chunks = (P + p_per_block - 1) / p_per_block;
    const int ldw = out_dig ? ld / 4 : 0;
    size_t smem = (((size_t)n * K1_THREADS + (size_t)p_per_block * n + (size_t)n * b + 3) & ~(size_t)3) +
                  (size_t)K1_THREADS * (ldw + 1) * 4;
    QSFT_CHECK_ARG(smem <= 200 * 1024, "delay block too large for shared memory (%zu bytes)", smem);
    QSFT_CHECK_ARG(p_chunks <= 65535, "too many delay chunks");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)tiles, (unsigned)p_chunks);
    if (q == 2 || q == 4 || q == 16) {   // q = 2^w with w | 64: bit-field fast path
        int w = 0;
        while ((1 << w) < q) ++w;
        const size_t sm2 = (size_t)(b + p_per_block) * sizeof(U128) + (size_t)K1_THREADS * (ldw + 1) * 4;
        QSFT_CHECK_ARG(sm2 <= 200 * 1024, "delay block too large for shared memory (%zu bytes)", sm2);
        if (sm2 > 48 * 1024) QSFT_CUDA(cudaFuncSetAttribute(k1_lattice_pow2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
        k1_lattice_pow2_kernel<<<grid, K1_THREADS, sm2, st>>>(M, D, w, n, b, P, p_per_block, B, out_idx ? limbs : 2, out_idx, out_dig, ld);
        QSFT_LAUNCHED();
        return QSFT_OK;
    }
    if (limbs == 2 || !out_idx) {
        if (smem > 48 * 1024) QSFT_CUDA(cudaFuncSetAttribute(k1_lattice_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k1_lattice_kernel<2><<<grid, K1_THREADS, smem, st>>>(M, D, q, n, b, P, p_per_block, B, out_idx, out_dig, ld);
    } else {
        if (smem > 48 * 1024) QSFT_CUDA(cudaFuncSetAttribute(k1_lattice_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k1_lattice_kernel<1><<<grid, K1_THREADS, smem, st>>>(M, D, q, n, b, P, p_per_block, B, out_idx, out_dig, ld);
    }
    QSFT_LAUNCHED();
    return QSFT_OK;
}

extern "C" int qsft_dec_to_qary(const uint64_t* idx, int limbs, int64_t N, int q, int n, int8_t* dig, int ld, void* stream) {
    if (int rc = check_qn(q, n, limbs, ld, true)) return rc;
    QSFT_CHECK_ARG(N >= 0 && idx && dig, "bad arguments");
    if (N == 0) return QSFT_OK;
    int g = 0;
    uint64_t Qg = 1;
    while (Qg * (uint64_t)q < (1ull << 32) && g < n) {
        Qg *= q;
        ++g;
    }
    const long long blocks = (N + CODEC_THREADS - 1) / CODEC_THREADS;
    size_t smem = (size_t)CODEC_THREADS * (ld / 4 + 1) * 4;
    cudaStream_t st = (cudaStream_t)stream;
    if (limbs == 2)
        dec_to_qary_kernel<2><<<(unsigned)blocks, CODEC_THREADS, smem, st>>>(idx, N, q, n, g, (uint32_t)Qg, dig, ld);
    else
        dec_to_qary_kernel<1><<<(unsigned)blocks, CODEC_THREADS, smem, st>>>(idx, N, q, n, g, (uint32_t)Qg, dig, ld);
    QSFT_LAUNCHED();
    return QSFT_OK;
}

extern "C" int qsft_qary_to_dec(const int8_t* dig, int ld, int64_t N, int q, int n, uint64_t* idx, int limbs, void* stream) {
    if (int rc = check_qn(q, n, limbs, ld, true)) return rc;
    QSFT_CHECK_ARG(N >= 0 && idx && dig, "bad arguments");
    if (N == 0) return QSFT_OK;
    const long long blocks = (N + CODEC_THREADS - 1) / CODEC_THREADS;
    cudaStream_t st = (cudaStream_t)stream;
    if (limbs == 2)
        qary_to_dec_kernel<2><<<(unsigned)blocks, CODEC_THREADS, 0, st>>>(dig, ld, N, q, n, idx);
    else
        qary_to_dec_kernel<1><<<(unsigned)blocks, CODEC_THREADS, 0, st>>>(dig, ld, N, q, n, idx);
    QSFT_LAUNCHED();
    return QSFT_OK;
}
