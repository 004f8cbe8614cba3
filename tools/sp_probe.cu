// Probe for tcgen05.mma.sp.kind::i8 (2:4 structured-sparse A, metadata in TMEM): one CTA, M = 128, N = 128, K = 256
// logical (4 MMAs), operands written to shared memory by the threads (128-byte swizzle), metadata copied with
// tcgen05.cp 128x128b.  Checks the metadata conventions against a CPU product.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o sp_probe tools/sp_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint64_t desc_plain(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const uint8_t* __restrict__ Ac /* [128][128] */, const uint8_t* __restrict__ E /* [2][128][16] */,
             const uint8_t* __restrict__ B /* [128][256] */, int* __restrict__ D /* [128][128] */, int e_lbo, int ts) {
    extern __shared__ uint8_t raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = base;                 // 16 KB
    uint8_t* sB = base + 16384;         // 2 x 16 KB (K halves)
    uint8_t* sE = base + 49152;         // 2 x 2 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + 49152 + 4096);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 8; i += 128) {          // 16-byte chunks of A (128 rows x 8 chunks)
        int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(Ac + r * 128 + c * 16);
    }
    for (int i = tid; i < 128 * 16; i += 128) {         // B rows are 256 bytes: two boxes of 128 bytes
        int r = i >> 4, c = i & 15, h = c >> 3, cc = c & 7;
        *reinterpret_cast<uint4*>(sB + h * 16384 + r * 128 + ((cc ^ (r & 7)) << 4)) =
            *reinterpret_cast<const uint4*>(B + r * 256 + c * 16);
    }
    for (int i = tid; i < 256; i += 128) reinterpret_cast<uint4*>(sE)[i] = reinterpret_cast<const uint4*>(E)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(su32(slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = *slot;
    if (tid == 0) {
        const uint32_t te = tm + 128;
        for (int c = 0; c < 2; ++c) {
            const uint64_t ed = desc_plain(su32(sE + c * 2048), (uint32_t)e_lbo, 128);
            asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(te + 4 * c), "l"(ed) : "memory");
        }
        // D = S32, A = B = signed int8, sparse, K-major, N = 128, M = 128
        const uint32_t idesc = (1u << 2) | (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        for (int j = 0; j < 4; ++j) {
            const uint64_t ad = desc_sw128(su32(sA)) + (uint64_t)(2 * j);                              // 32 compressed bytes
            const uint64_t bd = desc_sw128(su32(sB + (j >> 1) * 16384)) + (uint64_t)(4 * (j & 1));     // 64 bytes
            const uint32_t tej = te + 4 * (j >> 1) + 2 * (j & 1);
            const uint32_t acc = j > 0;
            if (ts) {       // A slice (128 rows x 32 compressed bytes) staged in TMEM columns 144 + 8 j
                const uint32_t ta = tm + 144 + 8 * j;
                asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(ta), "l"(ad) : "memory");
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                    "tcgen05.mma.sp.cta_group::1.kind::i8 [%0], [%1], %2, [%3], %4, p;\n\t}" ::"r"(tm),
                    "r"(ta), "l"(bd), "r"(tej), "r"(idesc), "r"(acc)
                    : "memory");
            } else
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                "tcgen05.mma.sp.cta_group::1.kind::i8 [%0], %1, %2, [%3], %4, p;\n\t}" ::"r"(tm),
                "l"(ad), "l"(bd), "r"(tej), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(bar)) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra DN;\n\tbra W;\n\tDN:\n\t}" ::"r"(
            su32(bar))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = tid;
    for (int ch = 0; ch < 8; ++ch) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tm + ((uint32_t)(warp * 32) << 16) + ch * 16)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[row * 128 + ch * 16 + j] = (int)r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u) : "memory");
}

// Timing: `iters` k-steps, each one N = 256 MMA into columns 0..255 and one N = 128 MMA into columns 256..383 (the
// production pattern), on resident operands (results meaningless); cycles per k-step by clock64.
//   sp: sparse (K = 64) or dense (K = 32);  ts: A read from TMEM;  cp: one tcgen05.cp 128x256b of the A slice per k-step
__global__ void __launch_bounds__(128, 1) rate_kernel(int sp, int ts, int cp, int iters, long long* __restrict__ cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + 16384 + 65536 + 4096);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (16384 + 65536 + 4096) / 4; i += 128) reinterpret_cast<uint32_t*>(base)[i] = 0x88888888u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(su32(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = *slot;
    if (tid == 0) {
        const uint32_t te = tm + 384, ta0 = tm + 400;
        const uint64_t ed = desc_plain(su32(base + 16384 + 65536), 128, 128);
        asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(te), "l"(ed) : "memory");
        const uint64_t ad = desc_sw128(su32(base)), bd = desc_sw128(su32(base + 16384)), bd2 = desc_sw128(su32(base + 16384 + 32768));
        for (int j = 0; j < 4; ++j) asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(ta0 + 8 * j), "l"(ad + 2 * j) : "memory");
        const uint32_t ib = (sp ? (1u << 2) : 0u) | (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
        const uint32_t i256 = ib | ((256u >> 3) << 17), i128 = ib | ((128u >> 3) << 17);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint64_t a = ad + (uint64_t)(2 * (it & 3)), kb = (uint64_t)((sp ? 4 : 2) * (it & 1));
            const uint32_t ta = ta0 + 8 * (it & 3), tej = te + 2 * (it & 1);
            if (cp) asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(ta), "l"(a) : "memory");
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t d = tm + (h ? 256u : 0u), idesc = h ? i128 : i256;
                const uint64_t b = (h ? bd2 : bd) + kb;
                if (sp && ts)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                                 "tcgen05.mma.sp.cta_group::1.kind::i8 [%0], [%1], %2, [%3], %4, p;\n\t}" ::"r"(d), "r"(ta), "l"(b),
                                 "r"(tej), "r"(idesc), "r"(1u) : "memory");
                else if (sp)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                                 "tcgen05.mma.sp.cta_group::1.kind::i8 [%0], %1, %2, [%3], %4, p;\n\t}" ::"r"(d), "l"(a), "l"(b),
                                 "r"(tej), "r"(idesc), "r"(1u) : "memory");
                else if (ts)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(ta), "l"(b), "r"(idesc),
                                 "r"(1u) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc),
                                 "r"(1u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(bar)) : "memory");
        asm volatile(
            "{\n\t.reg .pred P1;\n\tW2:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra DN2;\n\tbra W2;\n\tDN2:\n\t}" ::"r"(
                su32(bar))
            : "memory");
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
    const int M = 128, N = 128, K = 256;
    std::vector<int8_t> A(M * K, 0), B(N * K);
    std::vector<uint8_t> i0(M * K / 4), i1(M * K / 4);
    srand(7);
    for (int m = 0; m < M; ++m)
        for (int c = 0; c < K / 4; ++c) {
            int a = rand() % 4, b = rand() % 4;
            while (b == a) b = rand() % 4;
            if (a > b) { int t = a; a = b; b = t; }
            i0[m * (K / 4) + c] = a; i1[m * (K / 4) + c] = b;
            int va = rand() % 7 - 3, vb = rand() % 7 - 3;
            if (!va) va = 1;
            if (!vb) vb = -2;
            A[m * K + 4 * c + a] = va; A[m * K + 4 * c + b] = vb;
        }
    for (auto& v : B) v = rand() % 11 - 5;
    std::vector<int> ref(M * N);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            int s = 0;
            for (int k = 0; k < K; ++k) s += (int)A[m * K + k] * (int)B[n * K + k];
            ref[m * N + n] = s;
        }
    std::vector<uint8_t> Ac(M * K / 2);
    for (int m = 0; m < M; ++m)
        for (int c = 0; c < K / 4; ++c) {
            Ac[m * (K / 2) + 2 * c] = (uint8_t)A[m * K + 4 * c + i0[m * (K / 4) + c]];
            Ac[m * (K / 2) + 2 * c + 1] = (uint8_t)A[m * K + 4 * c + i1[m * (K / 4) + c]];
        }
    uint8_t *dA, *dE, *dB; int* dD;
    cudaMalloc(&dA, Ac.size()); cudaMalloc(&dE, 4096); cudaMalloc(&dB, B.size()); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, Ac.data(), Ac.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    for (int variant = 0; variant < 6; ++variant) {
        if (variant == 1 || variant == 3) continue;
        // variant bit 0: nibble = idx0 | idx1 << 2 (0) or idx1 | idx0 << 2 (1); bit 1: descriptor LBO 128 (0) or 16 (1)
        std::vector<uint8_t> E(4096, 0);
        for (int m = 0; m < M; ++m)
            for (int c = 0; c < K / 4; ++c) {
                int a = i0[m * (K / 4) + c], b = i1[m * (K / 4) + c];
                int nib = (variant & 1) ? (b | (a << 2)) : (a | (b << 2));
                int chunk = c / 32, cc = c % 32;                     // 128 logical K per 16-byte row
                E[chunk * 2048 + m * 16 + cc / 2] |= (uint8_t)(nib << (4 * (cc & 1)));
            }
        cudaMemcpy(dE, E.data(), 4096, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0xff, M * N * 4);
        probe_kernel<<<1, 128, 60000>>>(dA, dE, dB, dD, (variant & 2) ? 16 : 128, variant >= 4);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
        std::vector<int> got(M * N);
        cudaMemcpy(got.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int i = 0; i < M * N; ++i)
            if (got[i] != ref[i]) { if (first < 0) first = i; ++bad; }
        printf("variant %d%s: %d / %d mismatches", variant, variant >= 4 ? " (A via tcgen05.cp in TMEM)" : "", bad, M * N);
        if (first >= 0) printf(" (first at row %d col %d: got %d want %d)", first / N, first % N, got[first], ref[first]);
        printf("\n");
    }
    long long* dc;
    cudaMalloc(&dc, 148 * 8);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    for (int grid : {1, 148})
        for (int mode = 0; mode < 6; ++mode) {
            const int sp = mode >= 3, ts = (mode % 3) >= 1, cp = (mode % 3) == 2;
            const int iters = 8192;
            rate_kernel<<<grid, 128, 100000>>>(sp, ts, cp, iters, dc);
            rate_kernel<<<grid, 128, 100000>>>(sp, ts, cp, iters, dc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("rate mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
            long long hc[148];
            cudaMemcpy(hc, dc, grid * 8, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < grid; ++i) mx = hc[i] > mx ? hc[i] : mx;
            printf("grid %3d %s A from %s%s: %.1f cycles / k-step (N=256 + N=128)\n", grid, sp ? "sparse K=64" : "dense  K=32",
                   ts ? "TMEM" : "smem", cp ? " + tcgen05.cp per k-step" : "", (double)mx / iters);
        }
    return 0;
}
