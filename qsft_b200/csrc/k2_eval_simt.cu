// K2 (SIMT variant): synthetic sparse signal evaluation with packed-int8 dot products (dp4a) on CUDA cores.
//   out[m] = sum_s a_s * w^(<qdig[m], loc[s]> mod q)
// Replaces synt_exp/synt_src/synthetic_signal.py:100-118.  General (any q <= 127, any n <= 128, arbitrary queries);
// the tensor-core variant (k2_eval_tc.cu) takes over for shapes it supports.
#include "common.cuh"

namespace {

constexpr int K2_THREADS = 256;
constexpr int K2_QPT = 2;       // queries per thread (register tile)
constexpr int K2_TS = 256;      // support rows staged per shared-memory tile

__device__ __forceinline__ int dp4a_s32(uint32_t a, uint32_t b, int c) {
#ifdef QSFT_EMU   // CPU execution of this kernel source by tests/emu (test infrastructure; never defined in the product build)
    for (int i = 0; i < 4; ++i) c += (int)((a >> (8 * i)) & 255u) * (int)((b >> (8 * i)) & 255u);
    return c;
#else
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#endif
}

// Q == 4: rotation table a_s * i^t staged per tile (exact: sign / swap).  Q == 0: generic q, twiddle LUT + complex FMA.
template <int NW, int Q>
__global__ void __launch_bounds__(K2_THREADS)
k2_eval_simt_kernel(const int8_t* __restrict__ qdig, long long N, const int8_t* __restrict__ loc,
                    const float2* __restrict__ strengths, long long S, int q, uint32_t qmagic, int ld,
                    float2* __restrict__ out) {
    __shared__ uint32_t sK[K2_TS][NW];
    __shared__ float2 sA[K2_TS * (Q == 4 ? 4 : 1)];
    __shared__ float2 sTw[QSFT_MAX_Q + 1];

    const int tid = threadIdx.x;
    const long long m0 = ((long long)blockIdx.x * K2_THREADS + tid) * K2_QPT;

    uint32_t qw[K2_QPT][NW];
#pragma unroll
    for (int u = 0; u < K2_QPT; ++u) {
        const long long m = m0 + u;
#pragma unroll
        for (int w = 0; w < NW; ++w)
            qw[u][w] = (m < N) ? reinterpret_cast<const uint32_t*>(qdig + (size_t)m * ld)[w] : 0u;
    }
    if (Q == 0) {
        for (int t = tid; t < q; t += K2_THREADS) {
            float s, c;
            sincospif(2.0f * (float)t / (float)q, &s, &c);
            sTw[t] = make_float2(c, s);
        }
    }
    double xr[K2_QPT], xi[K2_QPT];
#pragma unroll
    for (int u = 0; u < K2_QPT; ++u) xr[u] = xi[u] = 0.0;

    for (long long s0 = 0; s0 < S; s0 += K2_TS) {
        __syncthreads();
        for (int i = tid; i < K2_TS * NW; i += K2_THREADS) {
            int r = i / NW, w = i - r * NW;
            long long s = s0 + r;
            sK[r][w] = (s < S) ? reinterpret_cast<const uint32_t*>(loc + (size_t)s * ld)[w] : 0u;
        }
        for (int r = tid; r < K2_TS; r += K2_THREADS) {
            long long s = s0 + r;
            float2 a = (s < S) ? strengths[s] : make_float2(0.f, 0.f);
            if (Q == 4) {
                sA[r * 4 + 0] = a;
                sA[r * 4 + 1] = make_float2(-a.y, a.x);
                sA[r * 4 + 2] = make_float2(-a.x, -a.y);
                sA[r * 4 + 3] = make_float2(a.y, -a.x);
            } else {
                sA[r] = a;
            }
        }
        __syncthreads();
        float pr[K2_QPT], pi[K2_QPT];
#pragma unroll
        for (int u = 0; u < K2_QPT; ++u) pr[u] = pi[u] = 0.f;
        const int lim = (int)min((long long)K2_TS, S - s0);
#pragma unroll 2
        for (int r = 0; r < lim; ++r) {
            uint32_t kw[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) kw[w] = sK[r][w];
#pragma unroll
            for (int u = 0; u < K2_QPT; ++u) {
                int acc = 0;
#pragma unroll
                for (int w = 0; w < NW; ++w) acc = dp4a_s32(qw[u][w], kw[w], acc);
                if (Q == 4) {
                    float2 v = sA[r * 4 + (acc & 3)];
                    pr[u] += v.x;
                    pi[u] += v.y;
                } else {
                    uint32_t t = (uint32_t)acc - __umulhi((uint32_t)acc, qmagic) * (uint32_t)q;
                    float2 tw = sTw[t];
                    float2 a = sA[r];
                    pr[u] = fmaf(a.x, tw.x, fmaf(-a.y, tw.y, pr[u]));
                    pi[u] = fmaf(a.x, tw.y, fmaf(a.y, tw.x, pi[u]));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < K2_QPT; ++u) {
            xr[u] += (double)pr[u];
            xi[u] += (double)pi[u];
        }
    }
#pragma unroll
    for (int u = 0; u < K2_QPT; ++u)
        if (m0 + u < N) out[m0 + u] = make_float2((float)xr[u], (float)xi[u]);
}

template <int NW>
int launch_nw(const int8_t* qdig, long long N, const int8_t* loc, const float2* a, long long S, int q, int ld,
              float2* out, cudaStream_t st) {
    const long long per_block = (long long)K2_THREADS * K2_QPT;
    const long long blocks = (N + per_block - 1) / per_block;
    const uint32_t qmagic = (uint32_t)(((1ull << 32) + q - 1) / q);
    if (q == 4)
        k2_eval_simt_kernel<NW, 4><<<(unsigned)blocks, K2_THREADS, 0, st>>>(qdig, N, loc, a, S, q, qmagic, ld, out);
    else
        k2_eval_simt_kernel<NW, 0><<<(unsigned)blocks, K2_THREADS, 0, st>>>(qdig, N, loc, a, S, q, qmagic, ld, out);
    QSFT_LAUNCHED();
    return QSFT_OK;
}

}  // namespace

int qsft_eval_synth_simt(const int8_t* qdig, int64_t N, const int8_t* loc, const float* strengths, int64_t S, int q,
                         int n, int ld, float* out, void* stream) {
    const int need = (n + 3) / 4;
    cudaStream_t st = (cudaStream_t)stream;
    const float2* a = reinterpret_cast<const float2*>(strengths);
    float2* o = reinterpret_cast<float2*>(out);
#define QSFT_K2_CASE(NWT) \
    if (need <= NWT) return launch_nw<NWT>(qdig, N, loc, a, S, q, ld, o, st);
    QSFT_K2_CASE(2) QSFT_K2_CASE(4) QSFT_K2_CASE(6) QSFT_K2_CASE(8) QSFT_K2_CASE(10) QSFT_K2_CASE(12)
    QSFT_K2_CASE(14) QSFT_K2_CASE(16) QSFT_K2_CASE(20) QSFT_K2_CASE(24) QSFT_K2_CASE(28) QSFT_K2_CASE(32)
#undef QSFT_K2_CASE
    qsft_set_error("n=%d too large", n);
    return QSFT_EINVAL;
}

extern "C" int qsft_eval_synth(const int8_t* qdig, int64_t N, const int8_t* loc, const float* strengths, int64_t S,
                               int q, int n, int ld, float* out, int impl, void* stream) {
    QSFT_CHECK_ARG(q >= 2 && q <= QSFT_MAX_Q, "q=%d out of range", q);
    QSFT_CHECK_ARG(n >= 1 && n <= QSFT_MAX_N, "n=%d out of range", n);
    QSFT_CHECK_ARG(ld >= n && ld % 16 == 0, "ld=%d must be >= n and a multiple of 16", ld);
    QSFT_CHECK_ARG(N >= 0 && S >= 0, "negative size");
    QSFT_CHECK_ARG(impl >= 0 && impl <= 2, "impl must be 0, 1 or 2");
    if (N == 0) return QSFT_OK;
    QSFT_CHECK_ARG(qdig && out && (S == 0 || (loc && strengths)), "null pointer");
    if (S == 0) {
        QSFT_CUDA(cudaMemsetAsync(out, 0, (size_t)N * 8, (cudaStream_t)stream));
        return QSFT_OK;
    }
    bool tc_ok = qsft_eval_synth_tc_supported(N, S, q, n, ld);
    if (impl == 2 && !tc_ok) {
        qsft_set_error("tcgen05 evaluation kernel does not support this shape (N=%lld S=%lld q=%d n=%d ld=%d)",
                       (long long)N, (long long)S, q, n, ld);
        return QSFT_EUNSUPPORTED;
    }
    if (impl == 2 || (impl == 0 && tc_ok)) return qsft_eval_synth_tc(qdig, N, loc, strengths, S, q, n, ld, out, stream);
    return qsft_eval_synth_simt(qdig, N, loc, strengths, S, q, n, ld, out, stream);
}
