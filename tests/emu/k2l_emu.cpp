// Host driver of the operand-generation kernels of the lattice-factorised evaluation (lt_prep / lt_amax / lt_quant /
// lt_bgen / lt_ttab / lt_etab in qsft_b200/csrc/k2_eval_lattice.cu) under the CPU execution shim -- TEST INFRASTRUCTURE
// ONLY.  The GEMM kernels themselves (tcgen05 / TMA) cannot be emulated; these kernels produce everything they consume.
#include "cuda_emu.h"
#define QSFT_EMU 1
#include "../../qsft_b200/csrc/common.cuh"
#include "_gen/k2l_device.inc"

// odd primes through the generic Z[w] generators (q = 3 as well: cross-check of the specialised q = 3 kernels)
template <int Q>
static int emu_agenq(const uint32_t* hhi, const uint8_t* e, long long S, long long Se, int b1, int P, long long Mhi, long long Kp, uint32_t* A) {
    const int T = 256;
    emu::launch(dim3((unsigned)((Kp / 4 + T - 1) / T), (unsigned)Mhi), dim3(T), [&]() { lt_agenq_kernel<Q>(hhi, e, S, Se, b1, P, Mhi, Kp, A); });
    return 0;
}
template <int Q>
static int emu_bgenq(const uint32_t* hlo, const int32_t* alimb, long long S, int b2, long long Nlo, long long Kp, int part, uint32_t* Bq) {
    const int T = 256;
    emu::launch(dim3((unsigned)((Kp / 4 + T - 1) / T), (unsigned)Nlo), dim3(T),
                [&]() { lt_bgenq_kernel<Q>(hlo, reinterpret_cast<const int2*>(alimb), S, b2, Nlo, Kp, part, Bq); });
    return 0;
}
template <int Q>
static int emu_combineq(const float* cre, const float* cim, long long rows, long long Nlo, float* out) {
    const int T = 256;
    emu::launch(dim3((unsigned)((rows * Nlo + T - 1) / T)), dim3(T), [&]() {
        lt_combineq_kernel<Q>(reinterpret_cast<const float2*>(cre), reinterpret_cast<const float2*>(cim), rows, Nlo, reinterpret_cast<float2*>(out));
    });
    return 0;
}

extern "C" {

int emu_lt_prep(const int8_t* M, const int8_t* D, const int8_t* loc, long long S, long long Se, int n, int b, int b1, int P,
                int ld, uint32_t* hhi, uint32_t* hlo, uint8_t* e, int q) {
    const int T = 256;
    const int fw = q <= 4 ? 2 : 3;
    emu::launch(dim3((unsigned)((S + T - 1) / T)), dim3(T), [&]() { lt_prep_kernel(M, D, loc, S, Se, n, b, b1, P, ld, hhi, hlo, e, q, fw); });
    return 0;
}

// q = 3: dense operands over Z[w] and the final combination
int emu_lt_agen3(const uint32_t* hhi, const uint8_t* e, long long S, long long Se, int b1, int P, long long Mhi, long long Kp,
                 uint32_t* A) {
    const int T = 256;
    emu::launch(dim3((unsigned)((Kp / 4 + T - 1) / T), (unsigned)Mhi), dim3(T),
                [&]() { lt_agen3_kernel(hhi, e, S, Se, b1, P, Mhi, Kp, A); });
    return 0;
}
int emu_lt_bgen3(const uint32_t* hlo, const int32_t* alimb, long long S, int b2, long long Nlo, long long Kp, int part, uint32_t* Bq) {
    const int T = 256;
    emu::launch(dim3((unsigned)((Kp / 4 + T - 1) / T), (unsigned)Nlo), dim3(T),
                [&]() { lt_bgen3_kernel(hlo, reinterpret_cast<const int2*>(alimb), S, b2, Nlo, Kp, part, Bq); });
    return 0;
}
int emu_lt_combine3(const float* cre, const float* cim, long long N, float* out) {
    const int T = 256;
    emu::launch(dim3((unsigned)((N + T - 1) / T)), dim3(T), [&]() {
        lt_combine3_kernel(reinterpret_cast<const float2*>(cre), reinterpret_cast<const float2*>(cim), N, reinterpret_cast<float2*>(out));
    });
    return 0;
}

int emu_lt_agenq(int q, const uint32_t* hhi, const uint8_t* e, long long S, long long Se, int b1, int P, long long Mhi, long long Kp,
                 uint32_t* A) {
    return q == 3 ? emu_agenq<3>(hhi, e, S, Se, b1, P, Mhi, Kp, A) : q == 5 ? emu_agenq<5>(hhi, e, S, Se, b1, P, Mhi, Kp, A)
                  : q == 7 ? emu_agenq<7>(hhi, e, S, Se, b1, P, Mhi, Kp, A) : -1;
}
int emu_lt_bgenq(int q, const uint32_t* hlo, const int32_t* alimb, long long S, int b2, long long Nlo, long long Kp, int part, uint32_t* Bq) {
    return q == 3 ? emu_bgenq<3>(hlo, alimb, S, b2, Nlo, Kp, part, Bq) : q == 5 ? emu_bgenq<5>(hlo, alimb, S, b2, Nlo, Kp, part, Bq)
                  : q == 7 ? emu_bgenq<7>(hlo, alimb, S, b2, Nlo, Kp, part, Bq) : -1;
}
int emu_lt_combineq(int q, const float* cre, const float* cim, long long rows, long long Nlo, float* out) {
    return q == 3 ? emu_combineq<3>(cre, cim, rows, Nlo, out) : q == 5 ? emu_combineq<5>(cre, cim, rows, Nlo, out)
                  : q == 7 ? emu_combineq<7>(cre, cim, rows, Nlo, out) : -1;
}

// pass 0: limbs of round(a * scale); pass 1: limbs of the quantisation residual (inv_scale holds two floats)
int emu_lt_quant(const float* a, long long S, float* inv_scale, int32_t* alimb /* (S, 2) */, int pass) {
    const int T = 256;
    unsigned int amax = 0;
    const unsigned sb = (unsigned)((S + T - 1) / T);
    emu::launch(dim3(sb), dim3(T), [&]() { lt_amax_kernel(reinterpret_cast<const float2*>(a), S, &amax); });
    emu::launch(dim3(sb), dim3(T), [&]() {
        lt_quant_kernel(reinterpret_cast<const float2*>(a), S, &amax, inv_scale, reinterpret_cast<int2*>(alimb), 0);
    });
    if (pass == 1) emu::launch(dim3(sb), dim3(T), [&]() {
        lt_quant_kernel(reinterpret_cast<const float2*>(a), S, &amax, inv_scale, reinterpret_cast<int2*>(alimb), 1);
    });
    return 0;
}

int emu_lt_bgen(const uint32_t* hlo, const int32_t* alimb, long long S, int b2, long long Nlo, long long Kp, uint32_t* Bq,
                int spread) {
    const int T = 256;
    emu::launch(dim3((unsigned)((Kp / 4 + T - 1) / T), (unsigned)((Nlo + LT_BGEN_LL - 1) / LT_BGEN_LL)), dim3(T),
                [&]() { lt_bgen_kernel(hlo, reinterpret_cast<const int2*>(alimb), S, b2, Nlo, Kp, Bq, spread); });
    return 0;
}

int emu_lt_ttab(const uint32_t* hhi, long long S, int b1, long long Tw, long long Mhi, uint32_t* T_, int spread) {
    const int T = 256;
    emu::launch(dim3((unsigned)((Tw + T - 1) / T), (unsigned)((Mhi + LT_TTAB_LL - 1) / LT_TTAB_LL)), dim3(T),
                [&]() { lt_ttab_kernel(hhi, S, b1, Tw, Mhi, T_, spread); });
    return 0;
}

int emu_lt_etab(const uint8_t* e, long long S, long long Se, int P, long long Tw, uint32_t* E2) {
    const int T = 256;
    emu::launch(dim3((unsigned)((Tw + T - 1) / T), (unsigned)P), dim3(T), [&]() { lt_etab_kernel(e, S, Se, P, Tw, E2); });
    return 0;
}

int emu_lt_agen(const uint32_t* hhi, const uint8_t* e, long long S, long long Se, int b1, int P, long long Mhi, long long Kp,
                uint32_t* A, int spread) {
    const int T = 256;
    emu::launch(dim3((unsigned)((Kp / 4 + T - 1) / T), (unsigned)Mhi), dim3(T),
                [&]() { lt_agen_kernel(hhi, e, S, Se, b1, P, Mhi, Kp, A, spread); });
    return 0;
}

// the A' expansion of the tensor-memory GEMM kernel (ts_expand<PX>) for n words: a4 (n, 4), e1 (n)
int emu_ts_expand(const uint32_t* r16, long long n, int im, int px, uint32_t* a4, uint32_t* e1) {
    for (long long i = 0; i < n; ++i) {
        if (px) ts_expand<true>(r16[i], im != 0, a4 + 4 * i, e1[i]);
        else ts_expand<false>(r16[i], im != 0, a4 + 4 * i, e1[i]);
    }
    return 0;
}

// the producers' whole per-word step: table words (t, e) -> expansion of (t + e) mod 4
int emu_ts_phase_expand(const uint32_t* tw, const uint32_t* ew, long long n, int im, int px, uint32_t* a4, uint32_t* e1) {
    for (long long i = 0; i < n; ++i) {
        if (px) ts_phase_expand<true>(tw[i], ew[i], im != 0, a4 + 4 * i, e1[i]);
        else ts_phase_expand<false>(tw[i], ew[i], im != 0, a4 + 4 * i, e1[i]);
    }
    return 0;
}
}
