// k4_peel_loop kernels for k digit rows of at most 8 32-bit words (see k4_peel_loop.cuh)
#include "k4_peel_loop.cuh"

int qsft_kl_launch_nw8(const KlArgs& a, const KlBlocks& blk, const KlMaps& maps, bool use_tma, size_t smem, int grid,
                       cudaStream_t st) {
    return kl_launch<8>(a, blk, maps, use_tma, smem, grid, st);
}
