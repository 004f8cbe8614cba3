// K3: batched b-dimensional length-q DFT (tensor product F_q^{(x)b}, no inter-axis twiddles), forward sign
// e^{-2 pi i/q}, scaled by 1/q^b, in place on complex64.  Replaces qsft/utils.py:31-36 (scipy.fft.fftn / q^n) as
// called from qsft/input_signal_subsampled.py:264-266.  HBM bound: each pass reads and writes every element once;
// a pass stages a tile in shared memory and runs radix-q butterflies over a group of axes.
//
// Axes are addressed by "level" u = stride exponent: element index j = sum_u digit_u q^u (digit_{b-1} is the most
// significant = first axis of the C-order reshape [q]*b).
#include "common.cuh"

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
template <int Q>
__global__ void __launch_bounds__(K3_THREADS)
k3_pass_kernel(float2* __restrict__ x, long long B, int q, int r, long long qa, int W, int outer, int rows /*q^r*/,
               long long tiles_per_block, float scale) {
    extern __shared__ float2 s[];
    __shared__ float2 s_tw[Q > 0 ? Q : QSFT_MAX_Q];
    const int tid = threadIdx.x;
    const int T = outer * rows * W;
    const int qq = Q > 0 ? Q : q;
    if (tid < qq) {
        float sn, cs;
        sincospif(-2.0f * (float)tid / (float)qq, &sn, &cs);
        s_tw[tid] = make_float2(cs, sn);
    }
    // locate the tile
    const long long tile = blockIdx.x;
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
    if (qa == 1) {
        for (int e = tid; e < T; e += K3_THREADS) s[e] = base[g0 + e];
    } else {
        for (int e = tid; e < T; e += K3_THREADS) {
            int t = e / W, w = e - t * W;
            s[e] = base[g0 + (long long)t * qa + w];
        }
    }
    __syncthreads();
    float2 tw[Q > 0 ? Q : 1];
    if (Q > 0) {
#pragma unroll
        for (int m = 0; m < (Q > 0 ? Q : 1); ++m) tw[m] = s_tw[m];
    }
    // butterflies
    int stride = W;
    const int nbf = T / qq;
    for (int u = 0; u < r; ++u) {
        for (int i = tid; i < nbf; i += K3_THREADS) {
            const int hi = i / stride, lo = i - hi * stride;
            const int e0 = hi * stride * qq + lo;
            if (Q > 0) {
                float2 v[Q > 0 ? Q : 1];
#pragma unroll
                for (int m = 0; m < (Q > 0 ? Q : 1); ++m) v[m] = s[e0 + m * stride];
                Dft<(Q > 0 ? Q : 2)>::run(reinterpret_cast<float2(&)[Q > 0 ? Q : 2]>(v), tw);
#pragma unroll
                for (int m = 0; m < (Q > 0 ? Q : 1); ++m) s[e0 + m * stride] = v[m];
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
    }
    // store
    if (qa == 1) {
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

template <int Q>
int launch_pass(float2* x, long long batch, long long B, int q, int a, int r, int cap, float scale, cudaStream_t st) {
    const long long qa = ipow64(q, a);
    const int rows = (int)ipow64(q, r);
    int W = 1, outer = 1;
    if (a == 0) {
        // as many whole q^r blocks as fit (and exist) in one tile
        long long fit = K3_TILE / rows;
        long long have = B / rows;
        long long o = 1;
        while (o * q <= fit && (have % (o * q)) == 0) o *= q;
        outer = (int)o;
    } else {
        int w = 0;
        while (w < a && (long long)W * q * rows <= K3_TILE) {
            W *= q;
            ++w;
        }
    }
    const int T = outer * rows * W;
    const long long tiles_per_block = B / T;
    const long long tiles = tiles_per_block * batch;
    QSFT_CHECK_ARG(tiles <= 0x7fffffffLL, "too many tiles");
    (void)cap;
    k3_pass_kernel<Q><<<(unsigned)tiles, K3_THREADS, (size_t)T * sizeof(float2), st>>>(x, B, q, r, qa, W, outer, rows,
                                                                                     tiles_per_block, scale);
    QSFT_LAUNCHED();
    return QSFT_OK;
}

}  // namespace

extern "C" int qsft_gwht_batch(float* x, int64_t batch, int q, int b, void* stream) {
    QSFT_CHECK_ARG(q >= 2 && q <= QSFT_MAX_Q, "q=%d out of range", q);
    QSFT_CHECK_ARG(b >= 0 && b <= QSFT_MAX_B, "b=%d out of range", b);
    QSFT_CHECK_ARG(batch >= 0, "negative batch");
    if (batch == 0 || b == 0) return QSFT_OK;
    QSFT_CHECK_ARG(x != nullptr, "null pointer");
    double Bd = 1;
    for (int i = 0; i < b; ++i) Bd *= q;
    QSFT_CHECK_ARG(Bd <= 4e12, "q^b too large");
    const long long B = ipow64(q, b);
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
    const float inv = (float)(1.0 / (double)B);
    cudaStream_t st = (cudaStream_t)stream;
    float2* xx = reinterpret_cast<float2*>(x);
    int a = 0;
    for (int p = 0; p < passes; ++p) {
        int r = (b - a + (passes - p) - 1) / (passes - p);
        const float scale = (p == passes - 1) ? inv : 1.0f;
        int rc;
        switch (q) {
            case 2: rc = launch_pass<2>(xx, batch, B, q, a, r, cap, scale, st); break;
            case 3: rc = launch_pass<3>(xx, batch, B, q, a, r, cap, scale, st); break;
            case 4: rc = launch_pass<4>(xx, batch, B, q, a, r, cap, scale, st); break;
            case 5: rc = launch_pass<5>(xx, batch, B, q, a, r, cap, scale, st); break;
            case 7: rc = launch_pass<7>(xx, batch, B, q, a, r, cap, scale, st); break;
            default: rc = launch_pass<0>(xx, batch, B, q, a, r, cap, scale, st); break;
        }
        if (rc) return rc;
        a += r;
    }
    return QSFT_OK;
}
