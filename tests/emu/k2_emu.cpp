// Host driver of the SIMT evaluation kernel (K2, dp4a variant) under the CPU execution shim -- TEST INFRASTRUCTURE ONLY.
// Mirrors qsft_eval_synth_simt / launch_nw in qsft_b200/csrc/k2_eval_simt.cu.
#include "cuda_emu.h"
#define QSFT_EMU 1
#include "../../qsft_b200/csrc/common.cuh"
#include "_gen/k2_device.inc"

namespace {
template <int NW>
void launch_nw_emu(const int8_t* qdig, long long N, const int8_t* loc, const float2* a, long long S, int q, int ld, float2* out) {
    const long long per_block = (long long)K2_THREADS * K2_QPT;
    const long long blocks = (N + per_block - 1) / per_block;
    const uint32_t qmagic = (uint32_t)(((1ull << 32) + q - 1) / q);
    if (q == 4)
        emu::launch(dim3((unsigned)blocks), dim3(K2_THREADS), [&]() { k2_eval_simt_kernel<NW, 4>(qdig, N, loc, a, S, q, qmagic, ld, out); });
    else
        emu::launch(dim3((unsigned)blocks), dim3(K2_THREADS), [&]() { k2_eval_simt_kernel<NW, 0>(qdig, N, loc, a, S, q, qmagic, ld, out); });
}
}  // namespace

extern "C" int emu_eval_synth(const int8_t* qdig, long long N, const int8_t* loc, const float* strengths, long long S, int q,
                              int n, int ld, float* out) {
    const int need = (n + 3) / 4;
    const float2* a = reinterpret_cast<const float2*>(strengths);
    float2* o = reinterpret_cast<float2*>(out);
#define CASE(NWT) \
    if (need <= NWT) { launch_nw_emu<NWT>(qdig, N, loc, a, S, q, ld, o); return 0; }
    CASE(2) CASE(4) CASE(6) CASE(8) CASE(10) CASE(12) CASE(14) CASE(16) CASE(20) CASE(24) CASE(28) CASE(32)
#undef CASE
    return -1;
}
