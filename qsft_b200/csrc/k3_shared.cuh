// Types shared by the K3 kernels (k3_gwht.cu, k3_gwht_tma.cu).  Also compiled by the CPU emulation (tests/emu).
#pragma once

namespace {

// Peer copies of the output (fused transform + all-gather): the final store of the last pass also writes the element to
// the same offset of the peers' symmetric U buffers -- either with ONE multimem.st to the buffers' NVLS multicast address
// (`mc` != nullptr: NVSwitch replicates the store to every rank, this one included; the GPU sends each byte once), or, where
// no multicast mapping exists, with up to 7 unicast P2P stores (each byte leaves the GPU once per peer).
struct K3Peers {
    float2* p[7];
    int n;
    float2* mc;                              // multicast alias of xroot (0: unicast peer stores)
};

__device__ __forceinline__ void k3_store(float2* dst, float2 v, const float2* xroot, const K3Peers& peers) {
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
