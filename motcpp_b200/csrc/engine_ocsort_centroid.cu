// engine_ocsort_centroid.cu - the OC-SORT frame-step kernels for asso_func = "centroid" (own translation unit: builds in
// parallel with the default "iou" kernels, which keep the association function as a compile-time constant).
#include "engine_launch.h"
#include "ocsort_kernel.cuh"

namespace mot {

template <int I>
static cudaError_t occ_set_smem(size_t bytes) {
    constexpr OcShape sh = kOcShapes[I];
    return cudaFuncSetAttribute(ocsort_centroid_step_kernel<sh.cap, sh.d_max, sh.e_cap>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
template <int I>
static void occ_launch_one(int grid, size_t smem, cudaStream_t st, const OcArgs& a) {
    constexpr OcShape sh = kOcShapes[I];
    ocsort_centroid_step_kernel<sh.cap, sh.d_max, sh.e_cap><<<grid, kOcThreads, smem, st>>>(a);
}
cudaError_t oc_centroid_prepare(int shape, size_t smem) {
    switch (shape) {
        case 0: return occ_set_smem<0>(smem);
        case 1: return occ_set_smem<1>(smem);
        default: return occ_set_smem<2>(smem);
    }
}
void oc_centroid_launch(int shape, int grid, size_t smem, cudaStream_t st, const OcArgs& a) {
    switch (shape) {
        case 0: occ_launch_one<0>(grid, smem, st, a); break;
        case 1: occ_launch_one<1>(grid, smem, st, a); break;
        default: occ_launch_one<2>(grid, smem, st, a); break;
    }
}
static_assert(kNumOcShapes == 3, "update the OC-SORT (centroid) dispatch switches");

}  // namespace mot
