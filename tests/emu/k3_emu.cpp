// Host driver of the K3 kernels under the CPU execution shim (tests/emu/cuda_emu.h) -- TEST INFRASTRUCTURE ONLY.
// Mirrors the planning of gwht_impl / launch_pass_q in qsft_b200/csrc/k3_gwht.cu (the pass plans themselves come from the
// product's own plan_pass) so that the DEVICE source, including the ticket order of the single-launch two-pass
// transform, can be exercised without a GPU.
#include "cuda_emu.h"
#define QSFT_EMU 1
#include "../../qsft_b200/csrc/common.cuh"
#include "_gen/k3_device.inc"

namespace {

template <int Q>
void run_pass(float2* x, long long B, int q, const PassPlan& p, long long blk0, long long nblk) {
    const long long tiles = p.tiles_per_block * nblk;
    const bool pow2 = (Q == 2 || Q == 4) && (p.T % 2 == 0) && (p.qa == 1 || p.W >= 2) && (B % 2 == 0);
    if (pow2) {
        if constexpr (Q == 2 || Q == 4)
            emu::launch(dim3((unsigned)tiles), dim3(K3_THREADS), [&]() {
                k3_pass_kernel<Q, true>(x, B, q, p.r, p.qa, p.W, p.lgW, p.outer, p.rows, p.tiles_per_block,
                                        blk0 * p.tiles_per_block, p.scale);
            });
    } else {
        emu::launch(dim3((unsigned)tiles), dim3(K3_THREADS), [&]() {
            k3_pass_kernel<Q, false>(x, B, q, p.r, p.qa, p.W, p.lgW, p.outer, p.rows, p.tiles_per_block,
                                     blk0 * p.tiles_per_block, p.scale);
        });
    }
}

void run_pass_q(float2* x, long long B, int q, const PassPlan& p, long long blk0, long long nblk, const K3Peers& peers) {
    if (q == 4 && p.T == 4096 && p.r <= 6) {
        const long long tiles = p.tiles_per_block * nblk;
        if (p.a == 0)
            emu::launch(dim3((unsigned)tiles), dim3(256), [&]() {
                k3_q4_fast_kernel<false>(x, B, p.r, 0, 1, 0, p.tiles_per_block, blk0 * p.tiles_per_block, p.scale, peers);
            });
        else
            emu::launch(dim3((unsigned)tiles), dim3(256), [&]() {
                k3_q4_fast_kernel<true>(x, B, p.r, p.lgW / 2, p.qa, p.lgW, p.tiles_per_block, blk0 * p.tiles_per_block, p.scale, peers);
            });
        return;
    }
    switch (q) {
        case 2: run_pass<2>(x, B, q, p, blk0, nblk); break;
        case 3: run_pass<3>(x, B, q, p, blk0, nblk); break;
        case 4: run_pass<4>(x, B, q, p, blk0, nblk); break;
        case 5: run_pass<5>(x, B, q, p, blk0, nblk); break;
        case 7: run_pass<7>(x, B, q, p, blk0, nblk); break;
        default: run_pass<0>(x, B, q, p, blk0, nblk); break;
    }
}

}  // namespace

// x (batch, q^b) complex64 in place; lag < 0 -> separate launches per pass even when the two-pass kernel applies;
// peer0 (may be NULL): a second buffer that receives the final stores (the fused all-gather path, one emulated peer).
// Returns 0, or 1 when the two-pass kernel found a strided tile scheduled before its dependencies (would hang a GPU).
extern "C" int emu_gwht(float* xf, long long batch, int q, int b, int lag, float* peer0, int* used_twopass) {
    float2* x = reinterpret_cast<float2*>(xf);
    emu_failure() = nullptr;
    *used_twopass = 0;
    if (batch == 0 || b == 0) return 0;
    const long long B = ipow64(q, b);
    int cap = 0;
    for (long long t = 1; t * q <= K3_TILE; t *= q) ++cap;
    const int passes = (b + cap - 1) / cap;
    const float inv = (float)(1.0 / (double)B);
    PassPlan plans[8];
    int a = 0;
    for (int p = 0; p < passes; ++p) {
        const int r = (p == 0) ? ((b < cap) ? b : cap) : (b - a + (passes - p) - 1) / (passes - p);
        plans[p] = plan_pass(B, q, a, r, (p == passes - 1) ? inv : 1.0f);
        a += r;
    }
    K3Peers peers;
    peers.n = peer0 ? 1 : 0;
    peers.mc = nullptr;
    peers.per = 0;
    peers.B = 0;
    peers.lgB = peers.lgper = -1;
    for (int r = 0; r < 8; ++r) peers.p[r] = nullptr;
    peers.p[0] = reinterpret_cast<float2*>(peer0);
    K3Peers none;
    none.n = 0;
    none.mc = nullptr;
    none.per = 0;
    none.B = 0;
    none.lgB = none.lgper = -1;
    if (lag >= 0 && q == 4 && passes == 2 && plans[0].T == 4096 && plans[1].T == 4096 && plans[0].r <= 6 && plans[1].r <= 6) {
        std::vector<unsigned int> done((size_t)batch + 1, 0u);
        const int t1 = (int)plans[0].tiles_per_block, t2 = (int)plans[1].tiles_per_block;
        *used_twopass = 1;
        emu::launch(dim3((unsigned)(batch * (t1 + t2))), dim3(256), [&]() {
            k3_q4_twopass_kernel<4>(x, B, plans[0].r, plans[1].r, plans[1].qa, plans[1].lgW, t1, t2, done.data(), batch, lag, inv, peers);
        });
        return emu_failure() ? 1 : 0;
    }
    for (int p = 0; p < passes; ++p) run_pass_q(x, B, q, plans[p], 0, batch, (p == passes - 1) ? peers : none);
    if (peers.n > 0 && !(q == 4 && plans[passes - 1].T == 4096 && plans[passes - 1].r <= 6))
        memcpy(peer0, xf, (size_t)batch * B * sizeof(float2));          // k3_bcast_copy_kernel's job
    return 0;
}
