// engine_strongsort.cu - instantiates the fused strongsort frame-step kernels (one per compiled shape) and their launchers.
#include "engine_launch.h"
#include "strongsort_kernel.cuh"

namespace mot {

template <int I>
static cudaError_t ss_set_smem(size_t bytes) {
    constexpr BtShape sh = kSsShapes[I];
    return cudaFuncSetAttribute(strongsort_step_kernel<sh.cap, sh.d_max, sh.e_cap>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
template <int I>
static void ss_launch_one(int grid, size_t smem, cudaStream_t st, const SsArgs& a) {
    constexpr BtShape sh = kSsShapes[I];
    strongsort_step_kernel<sh.cap, sh.d_max, sh.e_cap><<<grid, kSsThreads, smem, st>>>(a);
}
cudaError_t ss_prepare(int shape, size_t smem) { return shape == 0 ? ss_set_smem<0>(smem) : ss_set_smem<1>(smem); }
void ss_launch(int shape, int grid, size_t smem, cudaStream_t st, const SsArgs& a) {
    if (shape == 0) ss_launch_one<0>(grid, smem, st, a); else ss_launch_one<1>(grid, smem, st, a);
}
static_assert(kNumSsShapes == 2, "update the StrongSORT dispatch switches");

}  // namespace mot
