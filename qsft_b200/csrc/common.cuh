// Shared helpers for libqsft_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/qsft_b200.h"

#define QSFT_MAX_N 128      // digits per index (q^n must fit 128 bits)
#define QSFT_MAX_B 32
#define QSFT_MAX_Q 127      // digits are int8

void qsft_set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_qsft_launches;

#define QSFT_CHECK_ARG(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            qsft_set_error(__VA_ARGS__);          \
            return QSFT_EINVAL;                   \
        }                                         \
    } while (0)

#define QSFT_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            qsft_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return QSFT_ECUDA;                                                            \
        }                                                                                 \
    } while (0)

// call after every kernel launch: counts the launch and surfaces launch-configuration errors
#define QSFT_LAUNCHED()                                                                   \
    do {                                                                                  \
        g_qsft_launches.fetch_add(1, std::memory_order_relaxed);                          \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) {                                                         \
            qsft_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return QSFT_ECUDA;                                                            \
        }                                                                                 \
    } while (0)

static inline int64_t ipow64(int64_t q, int e) {
    int64_t r = 1;
    for (int i = 0; i < e; ++i) r *= q;
    return r;
}

// true when every index in [0, q^n) fits in `limbs` 64-bit words, i.e. q^n <= 2^(64*limbs)
static inline bool index_fits(int q, int n, int limbs) {
    unsigned __int128 v = 1;
    const unsigned __int128 top = (~(unsigned __int128)0) / (unsigned)q;
    for (int i = 0; i < n; ++i) {
        if (v > top) {  // v*q > 2^128 - 1; only acceptable if it is exactly 2^128 on the last step
            bool pow2 = (q & (q - 1)) == 0;
            int lg = 0;
            while ((1 << lg) < q) ++lg;
            return i == n - 1 && limbs == 2 && pow2 && v == ((unsigned __int128)1 << (128 - lg));
        }
        v *= (unsigned)q;
    }
    if (limbs == 1) return v <= ((unsigned __int128)1 << 64);
    return true;
}

__host__ __device__ static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

int qsft_num_sms();
// K2 variants (k2_eval_simt.cu / k2_eval_tc.cu), dispatched by qsft_eval_synth
int qsft_eval_synth_tc(const int8_t* qdig, int64_t N, const int8_t* loc, const float* strengths, int64_t S, int q, int n,
                       int ld, float* out, void* stream);
bool qsft_eval_synth_tc_supported(int64_t N, int64_t S, int q, int n, int ld);
cudaError_t qsft_scratch_alloc(void** p, size_t bytes, cudaStream_t st);
struct PeelDev;
struct UniqOut;
// K4: persistent on-device peel loop (k4_peel_loop.cu); QSFT_EUNSUPPORTED when the shape is not handled
struct KlShardHost {
    int rank, world;
    void* const* peers;
    unsigned int epoch;
};
int qsft_peel_loop(const PeelDev& d, const float* const* blocks, int64_t ldU, int64_t* find_cj, int8_t* find_k, float* find_rho,
                   int32_t* find_round, int32_t* find_id, int64_t max_finds, unsigned long long* counters, const UniqOut* uo,
                   int64_t* n_finds_out, int64_t* n_uniq_out, int* n_rounds_out, cudaStream_t st, const KlShardHost* shard = nullptr);
int64_t qsft_peel_loop_workspace_bytes(const PeelDev& d, int64_t max_finds);
