// Error reporting, launch counter, device queries.
#include <stdarg.h>
#include <string.h>

#include <thread>
#include <vector>

#include "common.cuh"

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_qsft_launches{0};

void qsft_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int qsft_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

// Stream-ordered scratch memory: keep freed blocks cached in the default pool (the default release threshold of 0
// hands memory back to the driver at every synchronisation, which makes the next cudaMallocAsync cost milliseconds).
cudaError_t qsft_scratch_alloc(void** p, size_t bytes, cudaStream_t st) {
    static bool tuned = false;
    if (!tuned) {
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        tuned = true;
    }
    return cudaMallocAsync(p, bytes, st);
}

namespace {
template <typename T>
void pack_rows(const T* src, int64_t r0, int64_t r1, int n, int64_t row_stride, int64_t col_stride, int8_t* dst, int ld) {
    for (int64_t r = r0; r < r1; ++r) {
        const T* s = src + r * row_stride;
        int8_t* d = dst + r * ld;
        if (col_stride == 1) {
            for (int u = 0; u < n; ++u) d[u] = (int8_t)s[u];
        } else {
            for (int u = 0; u < n; ++u) d[u] = (int8_t)s[(int64_t)u * col_stride];
        }
        for (int u = n; u < ld; ++u) d[u] = 0;
    }
}
}  // namespace

extern "C" {
// Host staging of a digit table (the support locq, delay matrices): `rows` rows of n integer digits (elem_bytes = 1 / 2 /
// 4 / 8, strides in elements, either orientation) -> int8 rows of ld bytes, zero padded, written straight into the caller's
// (pinned) staging buffer so that ONE DMA uploads the table in its device layout.  Runs on `threads` host threads of its own
// (torchrun sets OMP_NUM_THREADS=1, which made the same cast in torch / NumPy take 3-4 ms per rank for the 1e5 x 40 support
// of config 5 -- host time nothing on the GPU can overlap in a synchronous transform).  Host logic only: no device compute.
int qsft_host_pack_digits(const void* src, int elem_bytes, int64_t rows, int n, int64_t row_stride, int64_t col_stride,
                          int8_t* dst, int ld, int threads) {
    QSFT_CHECK_ARG(src && dst, "null pointer");
    QSFT_CHECK_ARG(rows >= 0 && n >= 0 && ld >= n, "bad shape");
    QSFT_CHECK_ARG(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 1, 2, 4 or 8");
    if (threads < 1) threads = 1;
    if (threads > 16) threads = 16;
    const int64_t min_rows = 4096;                                  // below this a thread costs more than it saves
    const int nt = (int)std::min<int64_t>(threads, std::max<int64_t>(1, rows / min_rows));
    auto work = [&](int64_t r0, int64_t r1) {
        switch (elem_bytes) {
            case 1: pack_rows((const int8_t*)src, r0, r1, n, row_stride, col_stride, dst, ld); break;
            case 2: pack_rows((const int16_t*)src, r0, r1, n, row_stride, col_stride, dst, ld); break;
            case 4: pack_rows((const int32_t*)src, r0, r1, n, row_stride, col_stride, dst, ld); break;
            default: pack_rows((const int64_t*)src, r0, r1, n, row_stride, col_stride, dst, ld); break;
        }
    };
    if (nt <= 1) {
        work(0, rows);
        return QSFT_OK;
    }
    std::vector<std::thread> pool;
    const int64_t per = (rows + nt - 1) / nt;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work, std::min(rows, t * per), std::min(rows, (t + 1) * per));
    work(0, std::min(rows, per));
    for (auto& th : pool) th.join();
    return QSFT_OK;
}

const char* qsft_last_error(void) { return g_err; }
int qsft_version(void) { return 100; }
int64_t qsft_launch_count(void) { return g_qsft_launches.load(); }
void qsft_reset_launch_count(void) { g_qsft_launches.store(0); }
}
