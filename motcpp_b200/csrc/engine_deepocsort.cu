// engine_deepocsort.cu - instantiates the fused DeepOC-SORT frame-step kernels (the OC-SORT kernel text with its
// appearance branches compiled in, one per compiled shape) and their launchers.
#include "engine_launch.h"
#include "ocsort_kernel.cuh"

namespace mot {

template <int I>
static cudaError_t deepoc_set_smem(size_t bytes) {
    constexpr OcShape sh = kOcShapes[I];
    return cudaFuncSetAttribute(deepocsort_step_kernel<sh.cap, sh.d_max, sh.e_cap>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
template <int I>
static void deepoc_launch_one(int grid, size_t smem, cudaStream_t st, const OcArgs& a) {
    constexpr OcShape sh = kOcShapes[I];
    deepocsort_step_kernel<sh.cap, sh.d_max, sh.e_cap><<<grid, kOcThreads, smem, st>>>(a);
}
cudaError_t deepoc_prepare(int shape, size_t smem) {
    switch (shape) {
        case 0: return deepoc_set_smem<0>(smem);
        case 1: return deepoc_set_smem<1>(smem);
        default: return deepoc_set_smem<2>(smem);
    }
}
void deepoc_launch(int shape, int grid, size_t smem, cudaStream_t st, const OcArgs& a) {
    switch (shape) {
        case 0: deepoc_launch_one<0>(grid, smem, st, a); break;
        case 1: deepoc_launch_one<1>(grid, smem, st, a); break;
        default: deepoc_launch_one<2>(grid, smem, st, a); break;
    }
}
static_assert(kNumOcShapes == 3, "update the DeepOC-SORT dispatch switches");

}  // namespace mot
