// engine_sort.cu - instantiates the fused sort frame-step kernels (one per compiled shape) and their launchers.
#include "engine_launch.h"
#include "sort_kernel.cuh"

namespace mot {

template <int I>
static cudaError_t sort_set_smem(size_t bytes) {
    constexpr BtShape sh = kBtShapes[I];
    return cudaFuncSetAttribute(sort_step_kernel<sh.cap, sh.d_max, sh.e_cap>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
template <int I>
static void sort_launch_one(int grid, size_t smem, cudaStream_t st, const SortArgs& a) {
    constexpr BtShape sh = kBtShapes[I];
    sort_step_kernel<sh.cap, sh.d_max, sh.e_cap><<<grid, kSortThreads, smem, st>>>(a);
}
cudaError_t sort_prepare(int shape, size_t smem) {
    switch (shape) {
        case 0: return sort_set_smem<0>(smem);
        case 1: return sort_set_smem<1>(smem);
        case 2: return sort_set_smem<2>(smem);
        default: return sort_set_smem<3>(smem);
    }
}
void sort_launch(int shape, int grid, size_t smem, cudaStream_t st, const SortArgs& a) {
    switch (shape) {
        case 0: sort_launch_one<0>(grid, smem, st, a); break;
        case 1: sort_launch_one<1>(grid, smem, st, a); break;
        case 2: sort_launch_one<2>(grid, smem, st, a); break;
        default: sort_launch_one<3>(grid, smem, st, a); break;
    }
}

}  // namespace mot
