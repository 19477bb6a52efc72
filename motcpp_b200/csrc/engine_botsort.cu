// engine_botsort.cu - instantiates the fused botsort frame-step kernels (one per compiled shape) and their launchers.
#include "engine_launch.h"
#include "botsort_kernel.cuh"

namespace mot {

template <int I>
static cudaError_t bot_set_smem(size_t bytes) {
    constexpr BtShape sh = kBotShapes[I];
    return cudaFuncSetAttribute(botsort_step_kernel<sh.cap, sh.d_max, sh.e_cap>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
template <int I>
static void bot_launch_one(int grid, size_t smem, cudaStream_t st, const BotArgs& a) {
    constexpr BtShape sh = kBotShapes[I];
    botsort_step_kernel<sh.cap, sh.d_max, sh.e_cap><<<grid, kBotThreads, smem, st>>>(a);
}
cudaError_t bot_prepare(int shape, size_t smem) {
    switch (shape) {
        case 0: return bot_set_smem<0>(smem);
        case 1: return bot_set_smem<1>(smem);
        default: return bot_set_smem<2>(smem);
    }
}
void bot_launch(int shape, int grid, size_t smem, cudaStream_t st, const BotArgs& a) {
    switch (shape) {
        case 0: bot_launch_one<0>(grid, smem, st, a); break;
        case 1: bot_launch_one<1>(grid, smem, st, a); break;
        default: bot_launch_one<2>(grid, smem, st, a); break;
    }
}
static_assert(kNumBotShapes == 3, "update the BoT-SORT dispatch switches");

}  // namespace mot
