// Error reporting, launch counter, device queries.
#include <stdarg.h>
#include <string.h>

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

extern "C" {
const char* qsft_last_error(void) { return g_err; }
int qsft_version(void) { return 100; }
int64_t qsft_launch_count(void) { return g_qsft_launches.load(); }
void qsft_reset_launch_count(void) { g_qsft_launches.store(0); }
}
