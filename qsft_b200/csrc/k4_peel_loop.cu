// K4, default path: host side of the persistent on-device peel loop (kernels: k4_peel_loop.cuh, instantiated per digit-word
// count NW in k4_peel_loop_nw8.cu / _nw16.cu / _nw32.cu so that the translation units compile in parallel).
#include "k4_peel_loop.cuh"

int qsft_kl_launch_nw8(const KlArgs& a, const KlBlocks& blk, const KlMaps& maps, bool use_tma, size_t smem, int grid,
                       cudaStream_t st);
int qsft_kl_launch_nw16(const KlArgs& a, const KlBlocks& blk, const KlMaps& maps, bool use_tma, size_t smem, int grid,
                        cudaStream_t st);
int qsft_kl_launch_nw32(const KlArgs& a, const KlBlocks& blk, const KlMaps& maps, bool use_tma, size_t smem, int grid,
                        cudaStream_t st);

// Whole peel loop on the device.  blocks[c * R + r] -> (P_src, ldU) complex64 rows of group c, repeat r (device pointers,
// host array).  Outputs as qsft_peel.  Returns QSFT_EUNSUPPORTED when the shape does not fit this kernel (C * R > 16,
// P_src > 256, tile larger than the shared memory).
// layout of a sharded workspace: control block, then the find list and the find table
namespace {
struct KlWsLayout {
    long long off_ctl, off_cj, off_k, off_rho, off_round, off_id, off_res, bytes;
};
KlWsLayout kl_ws_layout(const PeelDev& d, long long max_finds) {
    auto up = [](long long v) { return (v + 255) & ~255ll; };
    KlWsLayout L;
    L.off_ctl = 0;
    L.off_cj = up(8192 > (long long)sizeof(KlCtl) ? 8192 : (long long)sizeof(KlCtl));
    L.off_k = L.off_cj + up(max_finds * 8);
    L.off_rho = L.off_k + up(max_finds * d.ld);
    L.off_round = L.off_rho + up(max_finds * 8);
    L.off_id = L.off_round + up(max_finds * 4);
    L.off_res = L.off_id + up((long long)d.C * d.B * 4);
    L.bytes = L.off_res + up(max_finds * 4);
    return L;
}
}  // namespace

int64_t qsft_peel_loop_workspace_bytes(const PeelDev& d, int64_t max_finds) { return kl_ws_layout(d, max_finds).bytes; }

int qsft_peel_loop(const PeelDev& d, const float* const* blocks, int64_t ldU, int64_t* find_cj, int8_t* find_k, float* find_rho,
                   int32_t* find_round, int32_t* find_id, int64_t max_finds, unsigned long long* counters, const UniqOut* uo,
                   int64_t* n_finds_out, int64_t* n_uniq_out, int* n_rounds_out, cudaStream_t st, const KlShardHost* shard) {
    const int nblk = d.C * d.R;
    if (nblk > KL_MAX_BLOCKS || d.P_src > 256 || d.C > KL_MAX_BLOCKS) return QSFT_EUNSUPPORTED;
    static int sms = 0, smem_max = 0;
    if (!sms) {
        int dev = 0;
        QSFT_CUDA(cudaGetDevice(&dev));
        int coop = 0;
        QSFT_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        QSFT_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        sms = coop ? qsft_num_sms() : -1;
    }
    if (sms < 0) return QSFT_EUNSUPPORTED;
    KlArgs a{};
    a.d = d;
    a.ldU = ldU;
    bool use_tma = (d.B % 16 == 0) && (ldU % 2 == 0) && (getenv("QSFT_K4_NO_TMA") == nullptr);
    for (int i = 0; i < nblk; ++i)
        if ((uintptr_t)blocks[i] & 15) use_tma = false;
    const int budget = smem_max - 1024 - 2048;             // alignment slack + the kernel's static shared memory
    if (!kl_geometry(d, budget, use_tma ? 2 : 1, &a)) return QSFT_EUNSUPPORTED;
    KlBlocks blk{};
    for (int i = 0; i < nblk; ++i) blk.p[i] = reinterpret_cast<const float2*>(blocks[i]);
    KlMaps hm;
    memset(&hm, 0, sizeof(hm));
    for (int i = 0; i < nblk && use_tma; ++i) {
        const cuuint64_t dims[3] = {32, (cuuint64_t)(d.B / 16), (cuuint64_t)d.P_src};
        const cuuint64_t strides[2] = {128, (cuuint64_t)ldU * 8};
        const cuuint32_t box[3] = {32, (cuuint32_t)(a.W >> 4), (cuuint32_t)d.P_src};
        if (tma::make_map(&hm.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, blocks[i], dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) !=
            QSFT_OK)
            use_tma = false;                                // a shape the tensor map cannot express: plain copies instead
    }
    if (!use_tma) a.nstages = 1;
    const size_t smem = (size_t)a.nstages * a.stage_bytes + KL_CTRL_BYTES + (size_t)KL_NC * a.priv_bytes + 1024;
    // workspace: grid barrier, dirty-list counters, delay-structure flag, round counters | list heads | zres | bin classes
    // (all of these zeroed) | list links | dirty list | residuals of the finds
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t CB = (size_t)d.C * d.B;
    const size_t head_off = 256, zres_off = head_off + up(CB * 4), cls_off = zres_off + up(CB * 4), next_off = cls_off + up(CB);
    const size_t dirty_off = next_off + up((size_t)max_finds * d.C * 4), res_off = dirty_off + up(CB * 8);
    uint8_t* ws = nullptr;
    QSFT_CUDA(qsft_scratch_alloc((void**)&ws, res_off + up((size_t)max_finds * 4), st));
    QSFT_CUDA(cudaMemsetAsync(ws, 0, next_off, st));
    a.gbar = reinterpret_cast<unsigned int*>(ws);
    a.dcount = reinterpret_cast<unsigned long long*>(ws + 8);
    int* dflag = reinterpret_cast<int*>(ws + 64);
    a.multi = reinterpret_cast<unsigned long long*>(ws + 128);                 // 16 slots
    a.dstruct = dflag;
    a.head = reinterpret_cast<int32_t*>(ws + head_off);
    a.zres = reinterpret_cast<unsigned int*>(ws + zres_off);
    a.cls = ws + cls_off;
    a.next = reinterpret_cast<int32_t*>(ws + next_off);
    a.dirty = reinterpret_cast<long long*>(ws + dirty_off);
    a.max_dirty = (long long)CB;
    a.find_res = reinterpret_cast<float*>(ws + res_off);
    a.rank = 0;
    a.world = 1;
    a.jb = 0;
    a.je = d.B;
    a.seg = max_finds;
    if (shard != nullptr && shard->world > 1) {
        if (shard->world > 8 || shard->rank < 0 || shard->rank >= shard->world) {
            cudaFreeAsync(ws, st);
            qsft_set_error("bad shard (rank %d of %d; at most 8 ranks)", shard->rank, shard->world);
            return QSFT_EINVAL;
        }
        // the find list and the find table live in this rank's symmetric workspace (same layout on every rank)
        const KlWsLayout L = kl_ws_layout(d, max_finds);
        uint8_t* mine = static_cast<uint8_t*>(shard->peers[shard->rank]);
        find_cj = reinterpret_cast<int64_t*>(mine + L.off_cj);
        find_k = reinterpret_cast<int8_t*>(mine + L.off_k);
        find_rho = reinterpret_cast<float*>(mine + L.off_rho);
        find_round = reinterpret_cast<int32_t*>(mine + L.off_round);
        find_id = reinterpret_cast<int32_t*>(mine + L.off_id);
        a.find_res = reinterpret_cast<float*>(mine + L.off_res);
        a.rank = shard->rank;
        a.world = shard->world;
        const long long per = ((d.B + shard->world - 1) / shard->world + 127) & ~127ll;     // whole tiles per rank
        a.jb = per * shard->rank < d.B ? per * shard->rank : d.B;
        a.je = a.jb + per < d.B ? a.jb + per : d.B;
        a.seg = max_finds / shard->world;
        for (int p = 0; p < 8; ++p) a.sh.peer[p] = p < shard->world ? static_cast<uint8_t*>(shard->peers[p]) : nullptr;
        a.sh.off_cj = L.off_cj; a.sh.off_k = L.off_k; a.sh.off_rho = L.off_rho; a.sh.off_round = L.off_round;
        a.sh.off_id = L.off_id; a.sh.off_res = L.off_res; a.sh.off_ctl = L.off_ctl;
        a.sh.epoch = shard->epoch;
    }
    a.find_cj = (long long*)find_cj;
    a.find_k = find_k;
    a.find_rho = reinterpret_cast<float2*>(find_rho);
    a.find_round = find_round;
    a.find_id = find_id;
    a.max_finds = max_finds;
    if (uo) a.uo = *uo;
    a.has_uniq = uo ? 1 : 0;
    a.counters = counters;
    a.max_rounds = 15;
    a.chunk = kl_chunk(d, sms);
    if (const char* mr = getenv("QSFT_K4_MAX_ROUNDS"))       // measurement aid (tools/microbench.py): cost of the first rounds alone
        if (atoi(mr) >= 1 && atoi(mr) < 15) a.max_rounds = atoi(mr);
    a.peeling_max = pow((double)d.q, (double)d.n);
    a.guard_can_bind = a.peeling_max <= 15.0 * (double)d.C * (double)d.B ? 1 : 0;
    a.rel_floor = 1e-10f;
    QSFT_CUDA(cudaMemsetAsync(counters, 0, 8 * sizeof(unsigned long long), st));
    if (uo) QSFT_CUDA(cudaMemsetAsync(uo->seen0, 0, (size_t)d.B * sizeof(int32_t), st));
    kl_dstruct_kernel<<<1, 256, 0, st>>>(d, dflag);
    QSFT_LAUNCHED();
    // QSFT_K4_TIMING=1 (measurement aid): duration of the cooperative kernel alone on stderr, once the call has synchronised
    static thread_local cudaEvent_t tev[2] = {nullptr, nullptr};
    const bool timing = getenv("QSFT_K4_TIMING") != nullptr;
    if (timing) {
        if (!tev[0]) {
            QSFT_CUDA(cudaEventCreate(&tev[0]));
            QSFT_CUDA(cudaEventCreate(&tev[1]));
        }
        QSFT_CUDA(cudaEventRecord(tev[0], st));
    }
    int rc;
    const int nw = d.ld / 4;
    if (nw <= 8) rc = qsft_kl_launch_nw8(a, blk, hm, use_tma, smem, sms, st);
    else if (nw <= 16) rc = qsft_kl_launch_nw16(a, blk, hm, use_tma, smem, sms, st);
    else rc = qsft_kl_launch_nw32(a, blk, hm, use_tma, smem, sms, st);
    if (rc != QSFT_OK) {
        cudaFreeAsync(ws, st);
        return rc;
    }
    if (n_rounds_out == nullptr) {
        // asynchronous call: queued, not waited for.  The caller reads `counters` (device) when it needs the outcome:
        // [4] distinct k, [5] rounds, [6] != 0: find buffer too small, [7] find slots used
        QSFT_CUDA(cudaFreeAsync(ws, st));
        return QSFT_OK;
    }
    static thread_local unsigned long long* host = nullptr;
    if (!host) QSFT_CUDA(cudaMallocHost(&host, 8 * sizeof(unsigned long long)));
    QSFT_CUDA(cudaMemcpyAsync(host, counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    QSFT_CUDA(cudaFreeAsync(ws, st));
    if (timing) QSFT_CUDA(cudaEventRecord(tev[1], st));
    QSFT_CUDA(cudaStreamSynchronize(st));                   // the only synchronisation of the peel: result sizes
    if (timing) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, tev[0], tev[1]);
        fprintf(stderr, "[qsft_peel_loop] rank %d/%d kernel %.3f ms, rounds %llu, slots %llu\n", a.rank, a.world, ms, host[5], host[7]);
    }
    if (host[6]) {
        qsft_set_error("find buffer too small: %llu finds > max_finds=%lld", host[0], (long long)max_finds);
        return QSFT_EINVAL;
    }
    if (uo && (long long)host[4] > uo->max_uniq) {
        qsft_set_error("unique buffer too small: %llu > %lld", host[4], uo->max_uniq);
        return QSFT_EINVAL;
    }
    *n_finds_out = (int64_t)host[7];
    if (n_uniq_out) *n_uniq_out = (int64_t)host[4];
    *n_rounds_out = (int)host[5];
    return QSFT_OK;
}
