// engine_bytetrack.cu - instantiates the fused bytetrack frame-step kernels (one per compiled shape) and their launchers.
#include "engine_launch.h"
#include "bytetrack_kernel.cuh"

namespace mot {

// the 1024-thread variant exists for the C2 / C5 shape only (index 1); other shapes always run 512 threads
constexpr int kBtWideShape = 1;

template <int I>
static cudaError_t bt_set_smem(size_t bytes) {
    constexpr BtShape sh = kBtShapes[I];
    return cudaFuncSetAttribute(bytetrack_step_kernel<sh.cap, sh.d_max, sh.e_cap>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
template <int I>
static void bt_launch_one(int grid, size_t smem, cudaStream_t st, const BtArgs& a) {
    constexpr BtShape sh = kBtShapes[I];
    bytetrack_step_kernel<sh.cap, sh.d_max, sh.e_cap><<<grid, kBtThreads, smem, st>>>(a);
}
int bt_threads(int shape, int n_streams, int n_sms) { return (shape == kBtWideShape && n_streams <= n_sms) ? kBtThreadsWide : kBtThreads; }
cudaError_t bt_prepare(int shape, size_t smem, int threads) {
    if (threads == kBtThreadsWide) {
        constexpr BtShape sh = kBtShapes[kBtWideShape];
        return cudaFuncSetAttribute(bytetrack_step_kernel<sh.cap, sh.d_max, sh.e_cap, kBtThreadsWide>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    switch (shape) {
        case 0: return bt_set_smem<0>(smem);
        case 1: return bt_set_smem<1>(smem);
        case 2: return bt_set_smem<2>(smem);
        default: return bt_set_smem<3>(smem);
    }
}
void bt_launch(int shape, int grid, size_t smem, cudaStream_t st, const BtArgs& a, int threads) {
    if (threads == kBtThreadsWide) {
        constexpr BtShape sh = kBtShapes[kBtWideShape];
        bytetrack_step_kernel<sh.cap, sh.d_max, sh.e_cap, kBtThreadsWide><<<grid, kBtThreadsWide, smem, st>>>(a);
        return;
    }
    switch (shape) {
        case 0: bt_launch_one<0>(grid, smem, st, a); break;
        case 1: bt_launch_one<1>(grid, smem, st, a); break;
        case 2: bt_launch_one<2>(grid, smem, st, a); break;
        default: bt_launch_one<3>(grid, smem, st, a); break;
    }
}
static_assert(kNumBtShapes == 4, "update the dispatch switches");

}  // namespace mot
