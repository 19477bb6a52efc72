// TEST INFRASTRUCTURE: builds the product's kernel sources against the SIMT emulator
// (tests/cpusim/cpusim.hpp) into tests/cpusim/libmotb200_cpusim.so.  Never shipped, never
// loaded by motcpp_b200/.
#define MOT_CPUSIM 1
#include "cpusim.hpp"
#include "../../motcpp_b200/csrc/kernels_lap.cuh"

#include <vector>

extern "C" {

// one problem, host pointers; block size is a parameter so tests can cover several shapes
int sim_lap(const float* cost, int n, int m, int ld, float thresh, int* row2col, int* col2row,
            int e_cap, int threads) {
    const int n_max = n > 0 ? n : 1, m_max = m > 0 ? m : 1;
    std::vector<unsigned char> gs(mot::lap_gscratch_bytes(n_max, m_max) + 64);
    mot::LapBatchArgs a{};
    a.cost = cost; a.stride_cost = 0; a.n_rows = nullptr; a.n_cols = nullptr;
    a.n = n; a.m = m; a.ld = ld; a.thresh = thresh;
    a.row2col = row2col; a.col2row = col2row; a.gscratch = gs.data();
    a.n_max = n_max; a.m_max = m_max; a.e_cap = e_cap; a.n_problems = 1;
    const size_t smem = mot::lap_smem_bytes(n_max, m_max, e_cap);
    cpusim::launch(dim3(1), dim3(threads), smem, [=] { mot::lap_dense_kernel(a); });
    return 0;
}

}  // extern "C"
