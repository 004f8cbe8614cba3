// Types shared by the K3 kernels (k3_gwht.cu, k3_gwht_tma.cu).  Also compiled by the CPU emulation (tests/emu).
#pragma once

// (plain data shared across translation units)
struct K3Peers {
    float2* p[8];                            // gather: the n peers' buffers; scatter: every rank's buffer, indexed by rank
    int n;
    float2* mc;                              // multicast alias of xroot (0: unicast peer stores)
    long long B, per;                        // scatter: points per row, bins per rank (0: gather)
    int lgB, lgper;                          // their log2 when they are powers of two, else -1 (then divisions)
};

// K3, q = 4, 6 <= b <= 10: TMA pipeline (k3_gwht_tma.cu); QSFT_EUNSUPPORTED when the shape is not handled
#ifndef QSFT_EMU
int qsft_k3_q4_tma(float* x, int64_t batch, int b, const K3Peers& peers, cudaStream_t st);
#endif

namespace {

// Peer copies of the output, written by the final store of the last pass into the ranks' symmetric U buffers:
//   * gather (every rank gets every row: replicated / host-driven peeling): ONE multimem.st to the buffers' NVLS multicast
//     address (`mc` != nullptr: NVSwitch replicates the store to every rank, this one included; the GPU sends each byte
//     once), or, where no multicast mapping exists, up to 7 unicast P2P stores (each byte leaves the GPU once per peer);
//   * scatter (`per` > 0; bin-sharded on-device peeling): rank r only ever reads the bins [r * per, (r + 1) * per) of a row,
//     so an element goes to the ONE rank that owns its bin (p[owner], this rank's own buffer included) -- an all-to-all
//     instead of an all-gather: 1 / world of the NVLink traffic and nothing written that nobody reads.

__device__ __forceinline__ void k3_store(float2* dst, float2 v, const float2* xroot, const K3Peers& peers) {
    if (peers.per > 0) {
        const long long off = dst - xroot;
        const long long j = peers.lgB >= 0 ? (off & (peers.B - 1)) : off % peers.B;
        peers.p[peers.lgper >= 0 ? (j >> peers.lgper) : j / peers.per][off] = v;
        return;
    }
#ifndef QSFT_EMU
    if (peers.mc != nullptr) {
        asm volatile("multimem.st.weak.global.v2.f32 [%0], {%1, %2};" ::"l"(peers.mc + (dst - xroot)), "f"(v.x), "f"(v.y) : "memory");
        return;
    }
#endif
    *dst = v;
    if (peers.n > 0) {
        const long long off = dst - xroot;
#pragma unroll
        for (int r = 0; r < 7; ++r)
            if (r < peers.n) peers.p[r][off] = v;
    }
}

struct K3Ticket {
    long long blk;
    int t;          // tile inside its pass
    bool strided;
};

__host__ __device__ inline K3Ticket k3_ticket_decode(unsigned int ticket, long long nblocks, int tiles1, int tiles2, int lag) {
    K3Ticket o;
    const long long L = lag < nblocks ? (lag < 0 ? 0 : lag) : nblocks;
    const long long head = L * tiles1;                       // C(0) .. C(L-1)
    if ((long long)ticket < head) {
        o.blk = ticket / (unsigned int)tiles1;
        o.t = (int)(ticket - (unsigned int)(o.blk * tiles1));
        o.strided = false;
        return o;
    }
    const long long u = (long long)ticket - head;
    const long long per = (long long)tiles1 + tiles2;
    const long long groups = nblocks - L;                    // group g: C(g + L) then S(g)
    if (u < groups * per) {
        const long long g = u / per;
        const int v = (int)(u - g * per);
        o.strided = v >= tiles1;
        o.blk = o.strided ? g : g + L;
        o.t = o.strided ? v - tiles1 : v;
        return o;
    }
    const long long w = u - groups * per;                    // tail: S(nb - L) .. S(nb - 1)
    const long long g = w / tiles2;
    o.blk = groups + g;
    o.t = (int)(w - g * tiles2);
    o.strided = true;
    return o;
}

}  // namespace
