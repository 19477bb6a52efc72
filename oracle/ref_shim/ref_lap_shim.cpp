// TEST INFRASTRUCTURE.  C entry point around the reference's REAL LAP solver,
// compiled in place from /root/reference/include/motcpp/association/lap_solver.hpp
// (never vendored; see oracle/Makefile target `_ref/libref_lap.so`).
//
// Mirrors what utils::linear_assignment does around it
// (reference src/utils/matching.cpp:14-60): float cost -> double, solve, repack.
#include <Eigen/Dense>
#include <motcpp/association/lap_solver.hpp>

extern "C" {

// cost: row-major (n x m) fp32 with leading dimension ld.
// row2col[n], col2row[m]: -1 when unmatched.  Returns number of matches.
int ref_linear_assignment(const float* cost, int n, int m, int ld, float thresh,
                          int* row2col, int* col2row) {
    for (int i = 0; i < n; ++i) row2col[i] = -1;
    for (int j = 0; j < m; ++j) col2row[j] = -1;
    if (n == 0 || m == 0) return 0;
    Eigen::MatrixXd c(n, m);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) c(i, j) = static_cast<double>(cost[(size_t)i * ld + j]);
    std::vector<std::vector<int>> matches;
    std::vector<int> ua, ub;
    trackers::association::LAPSolver::linearAssignment(c, static_cast<double>(thresh), matches, ua, ub);
    for (const auto& mm : matches) {
        row2col[mm[0]] = mm[1];
        col2row[mm[1]] = mm[0];
    }
    return static_cast<int>(matches.size());
}

}  // extern "C"
