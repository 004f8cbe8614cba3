// Host driver of the K1 kernels (query lattice + index codecs) under the CPU execution shim -- TEST INFRASTRUCTURE ONLY.
// Mirrors the launch logic of qsft_query_lattice / qsft_dec_to_qary / qsft_qary_to_dec in qsft_b200/csrc/k1_lattice.cu.
#include "cuda_emu.h"
#define QSFT_EMU 1
#include "../../qsft_b200/csrc/common.cuh"

#include <algorithm>

void qsft_set_error(const char*, ...) {}
#include "_gen/k1_device.inc"

extern "C" {

int emu_query_lattice(const int8_t* M, const int8_t* D, int q, int n, int b, int P, uint64_t* out_idx, int limbs,
                      int8_t* out_dig, int ld, int num_sms) {
    if (int rc = check_qn(q, n, out_idx ? limbs : 2, ld, out_dig != nullptr)) return rc;
    const long long B = ipow64(q, b);
    const long long tiles = (B + K1_THREADS - 1) / K1_THREADS;
    int p_chunks = 1;
    const long long want = 4LL * num_sms;
    if (tiles < want) p_chunks = (int)std::min((long long)P, (want + tiles - 1) / tiles);
    int p_per_block = (P + p_chunks - 1) / p_chunks;
    p_chunks = (P + p_per_block - 1) / p_per_block;
    dim3 grid((unsigned)tiles, (unsigned)p_chunks);
    if (q == 2 || q == 4 || q == 16) {
        int w = 0;
        while ((1 << w) < q) ++w;
        emu::launch(grid, dim3(K1_THREADS), [&]() {
            k1_lattice_pow2_kernel(M, D, w, n, b, P, p_per_block, B, out_idx ? limbs : 2, out_idx, out_dig, ld);
        });
    } else if (limbs == 2 || !out_idx) {
        emu::launch(grid, dim3(K1_THREADS), [&]() { k1_lattice_kernel<2>(M, D, q, n, b, P, p_per_block, B, out_idx, out_dig, ld); });
    } else {
        emu::launch(grid, dim3(K1_THREADS), [&]() { k1_lattice_kernel<1>(M, D, q, n, b, P, p_per_block, B, out_idx, out_dig, ld); });
    }
    return 0;
}

int emu_dec_to_qary(const uint64_t* idx, int limbs, long long N, int q, int n, int8_t* dig, int ld) {
    if (int rc = check_qn(q, n, limbs, ld, true)) return rc;
    if (N == 0) return 0;
    int g = 0;
    uint64_t Qg = 1;
    while (Qg * (uint64_t)q < (1ull << 32) && g < n) {
        Qg *= q;
        ++g;
    }
    const long long blocks = (N + CODEC_THREADS - 1) / CODEC_THREADS;
    if (limbs == 2)
        emu::launch(dim3((unsigned)blocks), dim3(CODEC_THREADS), [&]() { dec_to_qary_kernel<2>(idx, N, q, n, g, (uint32_t)Qg, dig, ld); });
    else
        emu::launch(dim3((unsigned)blocks), dim3(CODEC_THREADS), [&]() { dec_to_qary_kernel<1>(idx, N, q, n, g, (uint32_t)Qg, dig, ld); });
    return 0;
}

int emu_qary_to_dec(const int8_t* dig, int ld, long long N, int q, int n, uint64_t* idx, int limbs) {
    if (int rc = check_qn(q, n, limbs, ld, true)) return rc;
    if (N == 0) return 0;
    const long long blocks = (N + CODEC_THREADS - 1) / CODEC_THREADS;
    if (limbs == 2)
        emu::launch(dim3((unsigned)blocks), dim3(CODEC_THREADS), [&]() { qary_to_dec_kernel<2>(dig, ld, N, q, n, idx); });
    else
        emu::launch(dim3((unsigned)blocks), dim3(CODEC_THREADS), [&]() { qary_to_dec_kernel<1>(dig, ld, N, q, n, idx); });
    return 0;
}
}
