// TEST INFRASTRUCTURE (oracle): restatement of the reference's linear-assignment path.
//   utils::linear_assignment          src/utils/matching.cpp:14-60
//   LAPSolver::linearAssignment/lapjv include/motcpp/association/lap_solver.hpp:251-332
//   lapjv_internal and helpers        include/motcpp/association/lap_solver.hpp:36-231
//
// The reference solves an (n+m) x (n+m) dense Jonker-Volgenant problem on
//     [ C      L/2 ]
//     [ L/2     0  ]          L = thresh
// in fp64.  This file re-expresses the same three JV phases (column reduction with reduction
// transfer, two sweeps of augmenting row reduction, shortest augmenting paths) over a flat
// row-major matrix.  Tie-breaking is part of the contract: scan orders, strict/non-strict
// comparisons and the 1e6 sentinel follow the cited lines, so results equal the reference's
// even when the optimum is not unique.  Pinned against the real header through
// oracle/_ref/libref_lap.so (tests/test_oracle_lap.py).
#include "oracle.h"

#include <cstddef>
#include <vector>

namespace {

constexpr double kBig = 1000000.0;   // lap_solver.hpp:24

class DenseJV {
public:
    DenseJV(int n, std::vector<double> cost) : n_(n), c_(std::move(cost)), x_(n, -1), y_(n, 0), v_(n, kBig) {}

    void solve() {
        std::vector<int> free_rows;
        column_reduction(free_rows);                                  // :36-72
        for (int sweep = 0; sweep < 2 && !free_rows.empty(); ++sweep)  // :221-224
            augmenting_row_reduction(free_rows);                      // :74-113
        for (int f : free_rows) augment_from(f);                      // :195-211
    }
    const std::vector<int>& row_to_col() const { return x_; }
    const std::vector<int>& col_to_row() const { return y_; }

private:
    double c(int i, int j) const { return c_[(size_t)i * n_ + j]; }

    // lap_solver.hpp:36-72
    void column_reduction(std::vector<int>& free_rows) {
        for (int i = 0; i < n_; ++i)
            for (int j = 0; j < n_; ++j)
                if (c(i, j) < v_[j]) { v_[j] = c(i, j); y_[j] = i; }   // strict: lowest row wins ties
        std::vector<char> sole(n_, 1);
        for (int j = n_ - 1; j >= 0; --j) {                            // right-to-left claim
            const int i = y_[j];
            if (x_[i] < 0) x_[i] = j;
            else { sole[i] = 0; y_[j] = -1; }
        }
        for (int i = 0; i < n_; ++i) {
            if (x_[i] < 0) { free_rows.push_back(i); continue; }
            if (!sole[i]) continue;
            const int j = x_[i];                                      // reduction transfer
            double second = kBig;
            for (int k = 0; k < n_; ++k) {
                if (k == j) continue;
                const double red = c(i, k) - v_[k];
                if (red < second) second = red;
            }
            v_[j] -= second;
        }
    }

    // lap_solver.hpp:74-113.  `free_rows` is rewritten in place with the rows still free.
    void augmenting_row_reduction(std::vector<int>& free_rows) {
        const unsigned n = (unsigned)n_;
        const unsigned total = (unsigned)free_rows.size();
        std::vector<int> fr(free_rows);          // working copy: entries before `cur` get overwritten
        unsigned cur = 0, rounds = 0;
        int kept = 0;
        while (cur < total) {
            ++rounds;
            const int i = fr[cur++];
            // best (j1,u1) and runner-up (j2,u2) reduced costs of row i
            int j1 = 0, j2 = -1;
            double u1 = c(i, 0) - v_[0], u2 = kBig;
            for (int j = 1; j < n_; ++j) {
                const double red = c(i, j) - v_[j];
                if (red < u2) {
                    if (red >= u1) { u2 = red; j2 = j; }
                    else { u2 = u1; u1 = red; j2 = j1; j1 = j; }
                }
            }
            int owner = y_[j1];
            const double lowered = v_[j1] - (u2 - u1);
            const bool strictly_lower = lowered < v_[j1];
            if (rounds < cur * n) {                                   // unsigned arithmetic as in :101
                if (strictly_lower) v_[j1] = lowered;
                else if (owner >= 0 && j2 >= 0) { j1 = j2; owner = y_[j2]; }
                if (owner >= 0) {
                    if (strictly_lower) fr[--cur] = owner;            // re-process the displaced row now
                    else fr[kept++] = owner;
                }
            } else if (owner >= 0) {
                fr[kept++] = owner;
            }
            x_[i] = j1;
            y_[j1] = i;
        }
        fr.resize(kept);
        free_rows.swap(fr);
    }

    // lap_solver.hpp:115-211 (find_path_dense + the augmentation loop of _ca_dense)
    void augment_from(int start) {
        std::vector<int> order(n_), pred(n_, start);
        std::vector<double> dist(n_);
        for (int j = 0; j < n_; ++j) { order[j] = j; dist[j] = c(start, j) - v_[j]; }
        int lo = 0, hi = 0, settled = 0, sink = -1;
        while (sink < 0) {
            if (lo == hi) {                                           // open the next distance level
                settled = lo;
                hi = lo + 1;
                double level = dist[order[lo]];
                for (int k = hi; k < n_; ++k) {
                    const int j = order[k];
                    if (dist[j] <= level) {
                        if (dist[j] < level) { hi = lo; level = dist[j]; }
                        order[k] = order[hi];
                        order[hi++] = j;
                    }
                }
                for (int k = lo; k < hi; ++k)
                    if (y_[order[k]] < 0) sink = order[k];             // last free column of the level
            }
            if (sink < 0) {                                           // relax from the level's columns
                int slo = lo, shi = hi;
                bool hit = false;
                while (slo != shi && !hit) {
                    const int jq = order[slo++];
                    const int i = y_[jq];
                    const double level = dist[jq];
                    const double base = c(i, jq) - v_[jq] - level;
                    for (int k = shi; k < n_; ++k) {
                        const int j = order[k];
                        const double cand = c(i, j) - v_[j] - base;
                        if (cand < dist[j]) {
                            dist[j] = cand;
                            pred[j] = i;
                            if (cand == level) {
                                if (y_[j] < 0) { sink = j; hit = true; break; }
                                order[k] = order[shi];
                                order[shi++] = j;
                            }
                        }
                    }
                }
                if (!hit) { lo = slo; hi = shi; }                     // on a hit the caller's lo/hi stay put (:152)
            }
        }
        const double level = dist[order[lo]];
        for (int k = 0; k < settled; ++k) {
            const int j = order[k];
            v_[j] += dist[j] - level;
        }
        int j = sink, i;
        do {                                                          // flip the path (:203-208)
            i = pred[j];
            y_[j] = i;
            const int prev = x_[i];
            x_[i] = j;
            j = prev;
        } while (i != start);
    }

    int n_;
    std::vector<double> c_;
    std::vector<int> x_, y_;
    std::vector<double> v_;
};

}  // namespace

namespace {
// biased: adds the tie-break infinitesimal -(64 j + (i j mod 64)) 2^-50 to every real entry (see orc_linear_assignment_biased)
int solve_extended(const float* cost, int n, int m, int ld, float thresh, bool biased, int* row2col, int* col2row, int dup_first = 0) {
    for (int i = 0; i < n; ++i) row2col[i] = -1;
    for (int j = 0; j < m; ++j) col2row[j] = -1;
    if (n == 0 || m == 0) return 0;                                   // matching.cpp:20-28
    const int N = n + m;
    const double half = (double)thresh / 2.0;                         // lap_solver.hpp:299
    std::vector<double> ext((size_t)N * N);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double val;
            if (i < n && j < m) {
                val = (double)cost[(size_t)i * ld + j];                   // matching.cpp:31 cast
                if (biased) val += -(double)(64 * j + ((i * j) & 63)) * 0x1p-50;
                if (dup_first > 0 && i >= dup_first) val += (double)(j + 1) * 0x1p-50;
            }
            else if (i >= n && j >= m) val = 0.0;
            else val = half;
            ext[(size_t)i * N + j] = val;
        }
    DenseJV jv(N, std::move(ext));
    jv.solve();
    int matches = 0;
    for (int i = 0; i < n; ++i) {                                     // lap_solver.hpp:326-331
        const int j = jv.row_to_col()[i];
        row2col[i] = (j >= m) ? -1 : j;
        if (row2col[i] >= 0) ++matches;
    }
    for (int j = 0; j < m; ++j) {
        const int i = jv.col_to_row()[j];
        col2row[j] = (i >= n) ? -1 : i;
    }
    return matches;
}
}  // namespace

extern "C" int orc_linear_assignment(const float* cost, int n, int m, int ld, float thresh,
                                     int* row2col, int* col2row) {
    return solve_extended(cost, n, m, ld, thresh, false, row2col, col2row);
}

// NOT the reference's behaviour: the same problem with an infinitesimal added to every real entry in fp64,
//   cost(i,j) - (64 j + (i j mod 64)) 2^-50,
// far below the 2^-26 granularity of sums of fp32 costs of magnitude >= 0.25, so it only decides between
// otherwise exactly tied optima: it prefers the HIGHER column, and makes "two rows on two identical columns"
// unique as well.  This is the tie-break the CUDA OC-SORT kernel applies to the reference's bit-identical
// "twin" tracks (DESIGN.md "Ties"); tests use it to separate kernel-logic parity from LAPJV's scan-order
// tie-breaking.
extern "C" int orc_linear_assignment_biased(const float* cost, int n, int m, int ld, float thresh,
                                            int* row2col, int* col2row) {
    return solve_extended(cost, n, m, ld, thresh, true, row2col, col2row);
}

// NOT the reference's behaviour: rows i >= n_first (the second copies of StrongSORT's duplicated track rows, see
// oracle/strongsort.cpp "q1") get + (j + 1) 2^-50 added in fp64.  Between exactly tied optima this gives a lone
// candidate detection to the FIRST copy and, of two candidates, the lower-indexed one to the second copy - the CUDA
// StrongSORT kernel's rule for problems too large for its on-device LAPJV.
extern "C" int orc_linear_assignment_rowdup(const float* cost, int n, int m, int ld, float thresh, int n_first,
                                            int* row2col, int* col2row) {
    return solve_extended(cost, n, m, ld, thresh, false, row2col, col2row, n_first);
}
