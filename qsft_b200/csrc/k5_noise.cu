// Measurement noise on the device: independent N(0, sd^2) on the real and imaginary part of every bin.  Replaces the host
// loop of SyntheticSubsampledSignal.get_MDU (synt_exp/synt_src/synthetic_signal.py:120-130: np.random.normal per (i, j)
// block -- 6 GB of host normals at config 5, R = 3) for runs that do not need the reference's NumPy random stream.
// Counter-based Philox4x32-10: element e of the buffer always gets the same value for the same (seed, offset), whatever
// the grid, so ranks that share a seed add the same noise to their copies of U.
#include <curand_kernel.h>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
k5_add_noise_kernel(float4* __restrict__ x, long long n4 /* float4 = two complex bins */, float sd, unsigned long long seed,
                    unsigned long long offset) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)i + offset, 0ull, &st);      // one subsequence per float4
    const float4 g = curand_normal4(&st);
    float4 v = x[i];
    v.x = fmaf(sd, g.x, v.x);
    v.y = fmaf(sd, g.y, v.y);
    v.z = fmaf(sd, g.z, v.z);
    v.w = fmaf(sd, g.w, v.w);
    x[i] = v;
}

__global__ void k5_add_noise_tail_kernel(float* __restrict__ x, long long first, long long n, float sd, unsigned long long seed,
                                         unsigned long long offset) {
    const long long i = first + threadIdx.x;
    if (i >= n) return;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)(first / 4) + offset, 0ull, &st);
    const float4 g = curand_normal4(&st);
    const float gv[4] = {g.x, g.y, g.z, g.w};
    x[i] = fmaf(sd, gv[threadIdx.x & 3], x[i]);
}

}  // namespace

extern "C" int qsft_add_noise(float* U, int64_t n_bins, float sd, uint64_t seed, uint64_t offset, void* stream) {
    QSFT_CHECK_ARG(n_bins >= 0 && (n_bins == 0 || U), "null pointer");
    QSFT_CHECK_ARG(((uintptr_t)U & 15) == 0, "U must be 16-byte aligned");
    QSFT_CHECK_ARG(sd >= 0.f, "negative standard deviation");
    if (n_bins == 0 || sd == 0.f) return QSFT_OK;
    const long long nf = 2 * n_bins, n4 = nf / 4;
    cudaStream_t st = (cudaStream_t)stream;
    if (n4 > 0) {
        k5_add_noise_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<float4*>(U), n4, sd, seed, offset);
        QSFT_LAUNCHED();
    }
    if (nf > 4 * n4) {
        k5_add_noise_tail_kernel<<<1, 4, 0, st>>>(U, 4 * n4, nf, sd, seed, offset);
        QSFT_LAUNCHED();
    }
    return QSFT_OK;
}
