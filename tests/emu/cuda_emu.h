// Minimal CPU execution model for the SIMT kernels of libqsft_b200 -- TEST INFRASTRUCTURE ONLY (never linked into the
// product, which has no CPU path).  A kernel's device source is compiled by g++ against this header; a launch runs the
// grid block by block, every CUDA thread of a block as an OS thread:
//   __syncthreads            -> barrier over the block's live threads
//   __shfl*_sync / __ballot_sync / __syncwarp (full masks, warp-uniform control flow, as in our kernels)
//                            -> per-warp slot exchange bracketed by warp barriers
//   __shared__               -> function-local statics (build_emu.py rewrites the keyword), one block at a time
//   atomics                  -> GCC __atomic builtins
// It checks LOGIC (indexing, reductions, decisions), not the memory model or performance.
#pragma once
#include <cuda_runtime.h>

#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <cmath>
#include <functional>
#include <thread>
#include <vector>

#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __launch_bounds__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#ifndef __noinline__
#define __noinline__
#endif

namespace emu {

struct Warp {
    pthread_barrier_t bar;
    unsigned long long slot[32];
};
struct Block {
    pthread_barrier_t bar;
    int live;
};
struct Ctx {
    uint3 tid, bid;
    dim3 bdim, gdim;
    Warp* warp;
    Block* block;
    int lane;
};
inline Ctx*& cur() {
    static thread_local Ctx* c = nullptr;
    return c;
}

// Runs kernel body `fn` (a lambda calling the __global__ function with its arguments) over the grid: one OS thread per
// CUDA thread of a block, created once per launch; the threads walk through the blocks together, one block at a time
// (shared memory is a set of statics), separated by a barrier.
inline void launch(dim3 grid, dim3 block, const std::function<void()>& fn) {
    const int nthreads = (int)(block.x * block.y * block.z);
    const int nwarps = (nthreads + 31) / 32;
    Block blk;
    pthread_barrier_init(&blk.bar, nullptr, nthreads);
    pthread_barrier_t step;
    pthread_barrier_init(&step, nullptr, nthreads);
    std::vector<Warp> warps(nwarps);
    for (int w = 0; w < nwarps; ++w) {
        const int lanes = (w == nwarps - 1 && nthreads % 32) ? nthreads % 32 : 32;
        pthread_barrier_init(&warps[w].bar, nullptr, lanes);
    }
    std::vector<std::thread> ths;
    ths.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t) {
        ths.emplace_back([&, t]() {
            Ctx c;
            c.tid = make_uint3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            c.bdim = block;
            c.gdim = grid;
            c.warp = &warps[t / 32];
            c.block = &blk;
            c.lane = t % 32;
            cur() = &c;
            for (unsigned bz = 0; bz < grid.z; ++bz)
                for (unsigned by = 0; by < grid.y; ++by)
                    for (unsigned bx = 0; bx < grid.x; ++bx) {
                        c.bid = make_uint3(bx, by, bz);
                        fn();
                        pthread_barrier_wait(&step);          // next block only when every thread has left this one
                    }
            cur() = nullptr;
        });
    }
    for (auto& th : ths) th.join();
    for (int w = 0; w < nwarps; ++w) pthread_barrier_destroy(&warps[w].bar);
    pthread_barrier_destroy(&step);
    pthread_barrier_destroy(&blk.bar);
}

template <typename T>
inline T shfl_from(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    Ctx* c = cur();
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    c->warp->slot[c->lane] = raw;
    pthread_barrier_wait(&c->warp->bar);
    raw = c->warp->slot[src & 31];
    pthread_barrier_wait(&c->warp->bar);
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}

}  // namespace emu

#define threadIdx (emu::cur()->tid)
#define blockIdx (emu::cur()->bid)
#define blockDim (emu::cur()->bdim)
#define gridDim (emu::cur()->gdim)

// NOTE: a thread that returns from the kernel while other threads of its block still reach __syncthreads would hang a
// real GPU too; our kernels only return block- or warp-uniformly after their last block barrier.
inline void __syncthreads() { pthread_barrier_wait(&emu::cur()->block->bar); }
inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&emu::cur()->warp->bar); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

template <typename T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    const int lane = emu::cur()->lane;
    return emu::shfl_from(v, (lane & ~(width - 1)) | (src & (width - 1)));
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int mask, int width = 32) {
    (void)width;
    return emu::shfl_from(v, emu::cur()->lane ^ mask);
}
inline unsigned __ballot_sync(unsigned, int pred) {
    emu::Ctx* c = emu::cur();
    c->warp->slot[c->lane] = pred ? 1ull : 0ull;
    pthread_barrier_wait(&c->warp->bar);
    unsigned out = 0;
    for (int l = 0; l < 32; ++l) out |= (unsigned)(c->warp->slot[l] & 1ull) << l;
    pthread_barrier_wait(&c->warp->bar);
    return out;
}

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
template <typename T>
inline T __ldg(const T* p) { return *p; }
template <typename T>
inline T __ldcg(const T* p) { return *p; }
inline void sincospi(double x, double* s, double* c) {
    *s = sin(M_PI * x);
    *c = cos(M_PI * x);
}
inline void sincospif(float x, float* s, float* c) {
    *s = (float)sin(M_PI * (double)x);
    *c = (float)cos(M_PI * (double)x);
}

inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline float atomicAdd(float* p, float v) {
    unsigned* u = reinterpret_cast<unsigned*>(p);
    unsigned old = __atomic_load_n(u, __ATOMIC_SEQ_CST), nw;
    float f;
    do {
        memcpy(&f, &old, 4);
        f += v;
        memcpy(&nw, &f, 4);
    } while (!__atomic_compare_exchange_n(u, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    memcpy(&f, &old, 4);
    return f;
}
inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline int atomicCAS(int* p, int cmp, int val) {
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}

// dynamic shared memory of the block being executed (build_emu.py rewrites `extern __shared__ ... name[];`)
inline unsigned char* emu_dyn_smem() {
    alignas(128) static unsigned char buf[232 * 1024];
    return buf;
}

// CUDA's global-namespace math overloads used by the kernels
inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }

// failure channel for conditions a real GPU would wait on forever
inline const char*& emu_failure() {
    static const char* msg = nullptr;
    return msg;
}
inline void emu_fail(const char* msg) { emu_failure() = msg; }

inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}

// integer / conversion intrinsics used by the operand-generation kernels
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    const unsigned long long pool = ((unsigned long long)b << 32) | a;
    unsigned out = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned n = (sel >> (4 * i)) & 0x7u;                 // CUDA's __byte_perm honours THREE selector bits per nibble
        const unsigned byte = (unsigned)(pool >> (8 * n)) & 0xffu;  // (no sign-replication mode, unlike PTX prmt)
        out |= byte << (8 * i);
    }
    return out;
}
inline unsigned __vsub4(unsigned a, unsigned b) {
    unsigned out = 0;
    for (int i = 0; i < 4; ++i) out |= (((a >> (8 * i)) - (b >> (8 * i))) & 0xffu) << (8 * i);
    return out;
}
inline unsigned atomicMax(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline unsigned atomicMin(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline int __float2int_rn(float x) { return (int)nearbyintf(x); }
