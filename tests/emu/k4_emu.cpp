// Host driver of the K4 kernels under the CPU execution shim (tests/emu/cuda_emu.h) -- TEST INFRASTRUCTURE ONLY.
// Mirrors the launch logic of qsft_b200/csrc/k4_peel.cu (make_dev, the NW dispatch, the round loop of qsft_peel) so
// that the DEVICE source of the product can be exercised without a GPU.
#include "cuda_emu.h"
#define QSFT_EMU 1
#include "../../qsft_b200/csrc/common.cuh"
#include "_gen/k4_device.inc"

namespace {

int fill_dev(PeelDev* d, int q, int n, int b, int C, int P, int P_src, int channel, int source, int rs_t, int rs_s, int ld,
             float cutoff, const int8_t* MT, const int8_t* D, const int32_t* rs_exp, const int32_t* rs_log) {
    memset(d, 0, sizeof(*d));
    d->q = q; d->n = n; d->b = b; d->C = C; d->P = P; d->P_src = P_src; d->R = P / P_src;
    d->channel = channel; d->source = source; d->rs_t = rs_t; d->rs_s = rs_s; d->ld = ld;
    d->B = ipow64(q, b);
    d->thresh = (double)cutoff * (double)P;
    d->invP = 1.0 / (double)P;
    d->qmagic = (unsigned int)(((1ull << 32) + q - 1) / q);
    d->MT = MT; d->D = D; d->rs_exp = rs_exp; d->rs_log = rs_log;
    d->rs_order = source ? (int)ipow64(q, rs_s) : 0;
    const char* fd = getenv("QSFT_K4_FASTDET");
    d->fastdet = (fd && atoi(fd) != 0) ? 1 : 0;
    return 0;
}

int nw_of(int ld) { return ld / 4 <= 4 ? 4 : ld / 4 <= 8 ? 8 : ld / 4 <= 16 ? 16 : 32; }

#define NW_SWITCH(nw, CALL)                    \
    switch (nw) {                              \
        case 4: { constexpr int NW = 4; CALL; } break;   \
        case 8: { constexpr int NW = 8; CALL; } break;   \
        case 16: { constexpr int NW = 16; CALL; } break; \
        default: { constexpr int NW = 32; CALL; } break; \
    }

void classify(const PeelDev& d, const float2* U, long long jb, long long je, long long* cj, int8_t* fk, float2* rho,
              int32_t* frd, int32_t* fid, long long maxf, int round, unsigned long long* counters) {
    const int nw = nw_of(d.ld);
    dim3 grid((unsigned)((je - jb + K4_THREADS - 1) / K4_THREADS), (unsigned)d.C);
    NW_SWITCH(nw, emu::launch(grid, dim3(K4_THREADS), [&]() {
                  k4_classify_kernel<NW>(d, U, jb, je, cj, fk, rho, frd, fid, maxf, round, counters);
              }));
}

// The persistent loop kernel (k4_peel_loop.cu, plain-copy variant) as ONE block: mirrors the host side of qsft_peel_loop.
// maxW caps the tile width (128 = the product's choice) so that the narrower tile shapes are exercised too.
int peel_loop(const PeelDev& d, const float2* U, long long* cj, int8_t* fk, float2* rho, int32_t* frd, int32_t* fid,
              long long maxf, const UniqOut& uo, unsigned long long* counters, int maxW, bool in_stage) {
    if (d.C * d.R > KL_MAX_BLOCKS || d.P_src > 256) return -3;
    KlArgs a{};
    a.d = d;
    a.ldU = d.B;
    (void)in_stage;
    if (!kl_geometry(d, 228 * 1024, 1, &a)) return -3;
    while (a.W > maxW) {                                       // narrower tiles on request
        a.W >>= 1;
        a.lgW -= 1;
        a.box = ((a.W >> 4) * d.P_src * 128 + 1023) & ~1023;
        a.stage_bytes = (d.R * a.box + a.W * 4 + 1023) & ~1023;
    }
    a.nstages = 1;
    a.chunk = kl_chunk(d, 1);
    a.rank = 0; a.world = 1; a.jb = 0; a.je = d.B; a.seg = maxf;
    KlBlocks blk{};
    for (int c = 0; c < d.C; ++c)
        for (int r = 0; r < d.R; ++r) blk.p[c * d.R + r] = U + ((size_t)c * d.P + (size_t)r * d.P_src) * d.B;
    std::vector<int32_t> head((size_t)d.C * d.B, 0), next((size_t)maxf * d.C, 0);
    std::vector<unsigned int> zres((size_t)d.C * d.B, 0u);
    std::vector<uint8_t> cls((size_t)d.C * d.B, 0);
    std::vector<long long> dirty((size_t)d.C * d.B, 0);
    std::vector<float> fres((size_t)maxf, 0.f);
    unsigned long long dcount[2] = {0ull, 0ull};
    a.zres = zres.data(); a.cls = cls.data(); a.dirty = dirty.data(); a.max_dirty = (long long)d.C * d.B;
    a.find_res = fres.data(); a.dcount = dcount;
    std::vector<unsigned long long> multi(16, 0);
    unsigned int gbar = 0;
    int dflag = 0;
    emu::launch(dim3(1), dim3(256), [&]() { kl_dstruct_kernel(d, &dflag); });
    a.find_cj = cj; a.find_k = fk; a.find_rho = rho; a.find_round = frd; a.find_id = fid; a.max_finds = maxf;
    a.head = head.data(); a.next = next.data(); a.uo = uo; a.has_uniq = 1; a.counters = counters; a.multi = multi.data();
    a.gbar = &gbar; a.dstruct = &dflag; a.max_rounds = 15;
    a.peeling_max = pow((double)d.q, (double)d.n);
    a.guard_can_bind = a.peeling_max <= 15.0 * (double)d.C * (double)d.B ? 1 : 0;
    a.rel_floor = 1e-10f;
    const int nw = nw_of(d.ld);
    NW_SWITCH(nw, emu::launch(dim3(1), dim3(KL_CT), [&]() { k4_peel_loop_kernel<NW, false>(a, blk); }));
    return 0;
}

}  // namespace

extern "C" {

// one classification pass (qsft_peel_classify); impl 1 = default kernel, 2 = k4_classify_v2_kernel
int emu_classify(int q, int n, int b, int C, int P, int P_src, int channel, int source, int rs_t, int rs_s, int ld,
                 float cutoff, const int8_t* MT, const int8_t* D, const int32_t* rs_exp, const int32_t* rs_log,
                 const float* U, long long* find_cj, int8_t* find_k, float* find_rho, int32_t* find_round, int32_t* find_id,
                 long long max_finds, int round, unsigned long long* counters, int impl, long long j_begin, long long j_end) {
    PeelDev d;
    fill_dev(&d, q, n, b, C, P, P_src, channel, source, rs_t, rs_s, ld, cutoff, MT, D, rs_exp, rs_log);
    if (j_end < 0) j_end = d.B;
    (void)impl;
    classify(d, reinterpret_cast<const float2*>(U), j_begin, j_end, find_cj, find_k, reinterpret_cast<float2*>(find_rho), find_round,
             find_id, max_finds, round, counters);
    return 0;
}

// whole peel loop (qsft_peel): classify / reduce / apply rounds with the reference's stop rule
int emu_peel(int q, int n, int b, int C, int P, int P_src, int channel, int source, int rs_t, int rs_s, int ld, float cutoff,
             const int8_t* MT, const int8_t* D, const int32_t* rs_exp, const int32_t* rs_log, float* U, long long* find_cj,
             int8_t* find_k, float* find_rho, int32_t* find_round, int32_t* find_id, long long max_finds,
             int32_t* seen0, int8_t* uk, float* usum, int32_t* ucnt, long long* ukey, int32_t* unext, long long max_uniq,
             int impl, long long* n_finds_out, long long* n_uniq_out, int* n_rounds_out) {
    PeelDev d;
    fill_dev(&d, q, n, b, C, P, P_src, channel, source, rs_t, rs_s, ld, cutoff, MT, D, rs_exp, rs_log);
    unsigned long long counters[8] = {0};
    memset(seen0, 0, (size_t)d.B * sizeof(int32_t));
    const int nw = nw_of(d.ld);
    if (impl >= 2) {            // on-device loop: 2 = product's tile width, bins in shared memory; 3 = 32-bin tiles;
                                // 4 = as 2 (kept for the tests' parametrisation)
        UniqOut uo{seen0, uk, usum, ucnt, ukey, unext, max_uniq};
        if (int rc = peel_loop(d, reinterpret_cast<const float2*>(U), find_cj, find_k, reinterpret_cast<float2*>(find_rho), find_round,
                               find_id, max_finds, uo, counters, impl == 3 ? 32 : 128, impl == 4))
            return rc;
        if (counters[6]) return -1;
        if ((long long)counters[4] > max_uniq) return -2;
        *n_finds_out = (long long)counters[7];
        *n_uniq_out = (long long)counters[4];
        *n_rounds_out = (int)counters[5];
        return 0;
    }
    long long total = 0;
    int round = 0;
    bool cont = true;
    const double peeling_max = pow((double)d.q, (double)d.n);      // qsft.py:151
    while (cont && (double)counters[2] < peeling_max && round < 15) {
        ++round;
        counters[1] = 0;
        classify(d, reinterpret_cast<const float2*>(U), 0, d.B, find_cj, find_k, reinterpret_cast<float2*>(find_rho),
                 find_round, find_id, max_finds, round, counters);
        const long long now = (long long)counters[0], multis = (long long)counters[1];
        if (now > max_finds) return -1;
        const long long nf = now - total;
        if (multis == 0 || nf == 0) cont = false;
        if (nf > 0) {
            NW_SWITCH(nw, emu::launch(dim3((unsigned)((nf + K4_THREADS - 1) / K4_THREADS)), dim3(K4_THREADS), [&]() {
                          k4_reduce_kernel<NW>(d, find_cj, find_k, reinterpret_cast<const float2*>(find_rho), find_id, total, nf,
                                               now, round, seen0, uk, usum, ucnt, ukey, unext, max_uniq, counters);
                      }));
        }
        if (nf > 0 && (cont || peeling_max <= 15.0 * (double)d.C * (double)d.B)) {
            const int wpb = K4_THREADS / 32;
            NW_SWITCH(nw, emu::launch(dim3((unsigned)((nf + wpb - 1) / wpb)), dim3(K4_THREADS), [&]() {
                          k4_apply_kernel<NW>(d, reinterpret_cast<float2*>(U), 0, d.B, find_cj, find_k,
                                              reinterpret_cast<const float2*>(find_rho), find_id, total, nf, now, 1, counters + 2);
                      }));
        }
        total = now;
    }
    if ((long long)counters[4] > max_uniq) return -2;
    *n_finds_out = total;
    *n_uniq_out = (long long)counters[4];
    *n_rounds_out = round;
    return 0;
}

int emu_detect(const float* cols, long long N, int q, int n, int P, int P_src, int channel, int source, int rs_t, int rs_s,
               const int32_t* rs_exp, const int32_t* rs_log, int8_t* k_out, int ld_out) {
    PeelDev d;
    memset(&d, 0, sizeof(d));
    d.q = q; d.n = source ? n : P_src - 1; d.P = P; d.P_src = P_src; d.R = P / P_src; d.channel = channel; d.source = source;
    if (source) { d.rs_t = rs_t; d.rs_s = rs_s; d.rs_exp = rs_exp; d.rs_log = rs_log; d.rs_order = (int)ipow64(q, rs_s); }
    const char* fd = getenv("QSFT_K4_FASTDET");
    d.fastdet = (fd && atoi(fd) != 0) ? 1 : 0;
    const int wpb = K4_THREADS / 32;
    emu::launch(dim3((unsigned)((N + wpb - 1) / wpb)), dim3(K4_THREADS),
                [&]() { k4_detect_kernel(d, reinterpret_cast<const float2*>(cols), N, k_out, ld_out); });
    return 0;
}

int emu_mle(const float* cols, long long N, int P, const float* S, int K, int32_t* k_sel, float* residual) {
    const int wpb = K4_THREADS / 32;
    emu::launch(dim3((unsigned)((N + wpb - 1) / wpb)), dim3(K4_THREADS), [&]() {
        k4_mle_kernel(reinterpret_cast<const float2*>(cols), N, P, reinterpret_cast<const float2*>(S), K, k_sel, residual);
    });
    return 0;
}
}
