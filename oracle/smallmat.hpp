// TEST INFRASTRUCTURE (oracle).  Tiny dense fp32 matrix used to restate the reference's
// Eigen expressions as plain loops.  Products accumulate in ascending k starting from the
// first term, one rounding per multiply and one per add (no FMA: build with
// -ffp-contract=off).  This is the "arithmetic contract" the CUDA kernels reproduce
// bit-for-bit.
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

namespace orc {

struct Mat {
    int r = 0, c = 0;
    std::vector<float> d;
    Mat() = default;
    Mat(int rows, int cols, float fill = 0.0f) : r(rows), c(cols), d(static_cast<size_t>(rows) * cols, fill) {}
    float& operator()(int i, int j) { return d[static_cast<size_t>(i) * c + j]; }
    float operator()(int i, int j) const { return d[static_cast<size_t>(i) * c + j]; }
    static Mat identity(int n) {
        Mat m(n, n);
        for (int i = 0; i < n; ++i) m(i, i) = 1.0f;
        return m;
    }
    static Mat from(const float* p, int rows, int cols) {
        Mat m(rows, cols);
        for (size_t k = 0; k < m.d.size(); ++k) m.d[k] = p[k];
        return m;
    }
    void to(float* p) const {
        for (size_t k = 0; k < d.size(); ++k) p[k] = d[k];
    }
};

// C = A * B, ascending-k accumulation
inline Mat mul(const Mat& a, const Mat& b) {
    Mat o(a.r, b.c);
    for (int i = 0; i < a.r; ++i)
        for (int j = 0; j < b.c; ++j) {
            float acc = a(i, 0) * b(0, j);
            for (int k = 1; k < a.c; ++k) acc = acc + a(i, k) * b(k, j);
            o(i, j) = acc;
        }
    return o;
}

// C = A * B^T
inline Mat mul_bt(const Mat& a, const Mat& b) {
    Mat o(a.r, b.r);
    for (int i = 0; i < a.r; ++i)
        for (int j = 0; j < b.r; ++j) {
            float acc = a(i, 0) * b(j, 0);
            for (int k = 1; k < a.c; ++k) acc = acc + a(i, k) * b(j, k);
            o(i, j) = acc;
        }
    return o;
}

inline Mat add(const Mat& a, const Mat& b) {
    Mat o(a.r, a.c);
    for (size_t k = 0; k < o.d.size(); ++k) o.d[k] = a.d[k] + b.d[k];
    return o;
}

inline Mat sub(const Mat& a, const Mat& b) {
    Mat o(a.r, a.c);
    for (size_t k = 0; k < o.d.size(); ++k) o.d[k] = a.d[k] - b.d[k];
    return o;
}

inline Mat diag_sq(const float* std, int n) {
    Mat o(n, n);
    for (int i = 0; i < n; ++i) o(i, i) = std[i] * std[i];
    return o;
}

// Lower Cholesky factor of a symmetric n x n matrix (reads the lower triangle), unblocked,
// column by column (the order Eigen's LLT uses for small matrices):
//   x = S(k,k) - sum_{p<k} L(k,p)^2 ; L(k,k) = sqrt(x)
//   L(i,k) = (S(i,k) - sum_{p<k} L(i,p) L(k,p)) / L(k,k)
// Returns false when a pivot is <= 0 (Eigen reports NumericalIssue).
inline bool cholesky_lower(const Mat& s, Mat& l) {
    const int n = s.r;
    l = Mat(n, n);
    for (int k = 0; k < n; ++k) {
        float x = s(k, k);
        if (k > 0) {
            float acc = l(k, 0) * l(k, 0);
            for (int p = 1; p < k; ++p) acc = acc + l(k, p) * l(k, p);
            x = x - acc;
        }
        if (!(x > 0.0f)) return false;
        const float lkk = std::sqrt(x);
        l(k, k) = lkk;
        for (int i = k + 1; i < n; ++i) {
            float v = s(i, k);
            if (k > 0) {
                float acc = l(i, 0) * l(k, 0);
                for (int p = 1; p < k; ++p) acc = acc + l(i, p) * l(k, p);
                v = v - acc;
            }
            l(i, k) = v / lkk;
        }
    }
    return true;
}

// Solve (L L^T) x = b in place: forward then backward substitution.
inline void cholesky_solve(const Mat& l, float* b) {
    const int n = l.r;
    for (int k = 0; k < n; ++k) {
        float v = b[k];
        if (k > 0) {
            float acc = l(k, 0) * b[0];
            for (int p = 1; p < k; ++p) acc = acc + l(k, p) * b[p];
            v = v - acc;
        }
        b[k] = v / l(k, k);
    }
    for (int k = n - 1; k >= 0; --k) {
        float v = b[k];
        if (k < n - 1) {
            float acc = l(k + 1, k) * b[k + 1];
            for (int p = k + 2; p < n; ++p) acc = acc + l(p, k) * b[p];
            v = v - acc;
        }
        b[k] = v / l(k, k);
    }
}

// General inverse by LU with partial pivoting (what Eigen's .inverse() does for a dynamic
// MatrixXf): factor PA = LU, then solve for each unit vector.
inline Mat inverse_lu(const Mat& a) {
    const int n = a.r;
    Mat lu = a;
    std::vector<int> perm(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        float best = std::fabs(lu(k, k));
        for (int i = k + 1; i < n; ++i) {
            const float v = std::fabs(lu(i, k));
            if (v > best) { best = v; piv = i; }
        }
        if (piv != k) {
            for (int j = 0; j < n; ++j) { const float t = lu(k, j); lu(k, j) = lu(piv, j); lu(piv, j) = t; }
            const int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
        }
        for (int i = k + 1; i < n; ++i) {
            lu(i, k) = lu(i, k) / lu(k, k);
            for (int j = k + 1; j < n; ++j) lu(i, j) = lu(i, j) - lu(i, k) * lu(k, j);
        }
    }
    Mat inv(n, n);
    std::vector<float> y(n);
    for (int col = 0; col < n; ++col) {
        for (int i = 0; i < n; ++i) {
            float v = (perm[i] == col) ? 1.0f : 0.0f;
            for (int p = 0; p < i; ++p) v = v - lu(i, p) * y[p];
            y[i] = v;
        }
        for (int i = n - 1; i >= 0; --i) {
            float v = y[i];
            for (int p = i + 1; p < n; ++p) v = v - lu(i, p) * y[p];
            y[i] = v / lu(i, i);
        }
        for (int i = 0; i < n; ++i) inv(i, col) = y[i];
    }
    return inv;
}

}  // namespace orc
