// engine_boosttrack.cu - instantiates the fused BoostTrack frame-step kernels (one per compiled shape) and their launchers.
#include "engine_launch.h"
#include "boosttrack_kernel.cuh"

namespace mot {

template <int I>
static cudaError_t boost_set_smem(size_t bytes) {
    constexpr BtShape sh = kBtShapes[I];
    return cudaFuncSetAttribute(boosttrack_step_kernel<sh.cap, sh.d_max, sh.e_cap>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
template <int I>
static void boost_launch_one(int grid, size_t smem, cudaStream_t st, const BoostArgs& a) {
    constexpr BtShape sh = kBtShapes[I];
    boosttrack_step_kernel<sh.cap, sh.d_max, sh.e_cap><<<grid, kBoostThreads, smem, st>>>(a);
}
cudaError_t boost_prepare(int shape, size_t smem) {
    switch (shape) {
        case 0: return boost_set_smem<0>(smem);
        case 1: return boost_set_smem<1>(smem);
        case 2: return boost_set_smem<2>(smem);
        default: return boost_set_smem<3>(smem);
    }
}
void boost_launch(int shape, int grid, size_t smem, cudaStream_t st, const BoostArgs& a) {
    switch (shape) {
        case 0: boost_launch_one<0>(grid, smem, st, a); break;
        case 1: boost_launch_one<1>(grid, smem, st, a); break;
        case 2: boost_launch_one<2>(grid, smem, st, a); break;
        default: boost_launch_one<3>(grid, smem, st, a); break;
    }
}
static_assert(kNumBtShapes == 4, "update the BoostTrack dispatch switches");

}  // namespace mot
