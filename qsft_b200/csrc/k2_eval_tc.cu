// placeholder until the tcgen05 kernel lands
#include "common.cuh"
bool qsft_eval_synth_tc_supported(int64_t, int64_t, int, int, int) { return false; }
int qsft_eval_synth_tc(const int8_t*, int64_t, const int8_t*, const float*, int64_t, int, int, int, float*, void*) {
    qsft_set_error("tcgen05 evaluation kernel not built");
    return QSFT_EUNSUPPORTED;
}
