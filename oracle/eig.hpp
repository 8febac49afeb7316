// oracle/eig.hpp — TEST INFRASTRUCTURE ONLY.
//
// Symmetric eigen-decomposition by Householder tridiagonalisation followed by
// implicit-shift QL — the same family of algorithm as Eigen's
// SelfAdjointEigenSolver used by project_to_psd (utils/eigen_ext.tpp:75-77;
// SURVEY Appendix B.3): eigenvalues ascending, orthonormal eigenvectors,
// reads the lower triangle.  (Classic EISPACK tred2/tql2 formulation.)
#pragma once
#include <cmath>
#include <vector>
#include <algorithm>

namespace oracle {

// A: n x n col-major (ld = n), symmetric (lower triangle read).
// On return d[0..n) ascending eigenvalues, V (n x n col-major) eigenvectors in columns.
inline bool eig_sym(int n, const double* A, double* d, double* V)
{
    std::vector<double> e(n);
    auto v = [&](int r, int c) -> double& { return V[r + n * c]; };
    for (int c = 0; c < n; c++)
        for (int r = 0; r < n; r++) v(r, c) = r >= c ? A[r + n * c] : A[c + n * r];

    // --- Householder reduction to tridiagonal form
    for (int j = 0; j < n; j++) d[j] = v(n - 1, j);
    for (int i = n - 1; i > 0; i--) {
        double scale = 0.0, h = 0.0;
        for (int k = 0; k < i; k++) scale += std::abs(d[k]);
        if (scale == 0.0) {
            e[i] = d[i - 1];
            for (int j = 0; j < i; j++) {
                d[j] = v(i - 1, j);
                v(i, j) = 0.0;
                v(j, i) = 0.0;
            }
        } else {
            for (int k = 0; k < i; k++) {
                d[k] /= scale;
                h += d[k] * d[k];
            }
            double f = d[i - 1];
            double g = std::sqrt(h);
            if (f > 0) g = -g;
            e[i] = scale * g;
            h -= f * g;
            d[i - 1] = f - g;
            for (int j = 0; j < i; j++) e[j] = 0.0;
            for (int j = 0; j < i; j++) {
                f = d[j];
                v(j, i) = f;
                g = e[j] + v(j, j) * f;
                for (int k = j + 1; k <= i - 1; k++) {
                    g += v(k, j) * d[k];
                    e[k] += v(k, j) * f;
                }
                e[j] = g;
            }
            f = 0.0;
            for (int j = 0; j < i; j++) {
                e[j] /= h;
                f += e[j] * d[j];
            }
            const double hh = f / (h + h);
            for (int j = 0; j < i; j++) e[j] -= hh * d[j];
            for (int j = 0; j < i; j++) {
                f = d[j];
                g = e[j];
                for (int k = j; k <= i - 1; k++) v(k, j) -= (f * e[k] + g * d[k]);
                d[j] = v(i - 1, j);
                v(i, j) = 0.0;
            }
        }
        d[i] = h;
    }
    // accumulate transformations
    for (int i = 0; i < n - 1; i++) {
        v(n - 1, i) = v(i, i);
        v(i, i) = 1.0;
        const double h = d[i + 1];
        if (h != 0.0) {
            for (int k = 0; k <= i; k++) d[k] = v(k, i + 1) / h;
            for (int j = 0; j <= i; j++) {
                double g = 0.0;
                for (int k = 0; k <= i; k++) g += v(k, i + 1) * v(k, j);
                for (int k = 0; k <= i; k++) v(k, j) -= g * d[k];
            }
        }
        for (int k = 0; k <= i; k++) v(k, i + 1) = 0.0;
    }
    for (int j = 0; j < n; j++) {
        d[j] = v(n - 1, j);
        v(n - 1, j) = 0.0;
    }
    v(n - 1, n - 1) = 1.0;
    e[0] = 0.0;

    // --- implicit QL
    for (int i = 1; i < n; i++) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    double f = 0.0, tst1 = 0.0;
    const double eps = std::pow(2.0, -52.0);
    for (int l = 0; l < n; l++) {
        tst1 = std::max(tst1, std::abs(d[l]) + std::abs(e[l]));
        int m = l;
        while (m < n) {
            if (std::abs(e[m]) <= eps * tst1) break;
            m++;
        }
        if (m > l) {
            int iter = 0;
            do {
                if (++iter > 200) return false;
                double g = d[l];
                double p = (d[l + 1] - g) / (2.0 * e[l]);
                double r = std::hypot(p, 1.0);
                if (p < 0) r = -r;
                d[l] = e[l] / (p + r);
                d[l + 1] = e[l] * (p + r);
                const double dl1 = d[l + 1];
                double h = g - d[l];
                for (int i = l + 2; i < n; i++) d[i] -= h;
                f += h;
                p = d[m];
                double c = 1.0, c2 = c, c3 = c, el1 = e[l + 1], s = 0.0, s2 = 0.0;
                for (int i = m - 1; i >= l; i--) {
                    c3 = c2;
                    c2 = c;
                    s2 = s;
                    g = c * e[i];
                    h = c * p;
                    r = std::hypot(p, e[i]);
                    e[i + 1] = s * r;
                    s = e[i] / r;
                    c = p / r;
                    p = c * d[i] - s * g;
                    d[i + 1] = h + s * (c * g + s * d[i]);
                    for (int k = 0; k < n; k++) {
                        h = v(k, i + 1);
                        v(k, i + 1) = s * v(k, i) + c * h;
                        v(k, i) = c * v(k, i) - s * h;
                    }
                }
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                e[l] = s * p;
                d[l] = c * p;
            } while (std::abs(e[l]) > eps * tst1);
        }
        d[l] = d[l] + f;
        e[l] = 0.0;
    }
    // sort ascending
    for (int i = 0; i < n - 1; i++) {
        int k = i;
        double p = d[i];
        for (int j = i + 1; j < n; j++)
            if (d[j] < p) {
                k = j;
                p = d[j];
            }
        if (k != i) {
            d[k] = d[i];
            d[i] = p;
            for (int j = 0; j < n; j++) std::swap(v(j, i), v(j, k));
        }
    }
    return true;
}

// project_to_psd (utils/eigen_ext.tpp:56-108): mode 0 NONE, 1 CLAMP, 2 ABS.
// A is n x n col-major with leading dimension lda; result overwrites A.
//
// Deviation that cannot be pinned without Eigen itself: when a stencil point is not involved in the
// local matrix (all of its 3 rows / columns are exactly zero, e.g. an edge-edge collision whose
// distance type is a vertex-edge pair), the eigen-solver's Householder reflectors mix that
// coordinate in and the reconstruction leaves rounding noise (<= 1e-15 ||A||) there, in an
// ordering-dependent way.  Such rows span an invariant subspace of the projection, so they are
// removed before the solve and stay EXACTLY zero (no entry in the assembled sparse matrix).
inline void project_to_psd(int n, double* A, int lda, int mode)
{
    if (mode == 0) return;
    if (n % 3 == 0 && n > 3) {
        std::vector<int> used;
        for (int p = 0; p < n / 3; p++) {
            bool any = false;
            for (int r = 3 * p; r < 3 * p + 3; r++)
                for (int c = 0; c < n; c++) any |= A[r + lda * c] != 0.0 || A[c + lda * r] != 0.0;
            if (any) used.push_back(p);
        }
        if (int(used.size()) < n / 3) {
            const int m = 3 * int(used.size());
            if (m == 0) return;
            std::vector<double> sub(m * m);
            auto g = [&](int i) { return 3 * used[i / 3] + i % 3; };
            for (int c = 0; c < m; c++)
                for (int r = 0; r < m; r++) sub[r + m * c] = A[g(r) + lda * g(c)];
            project_to_psd(m, sub.data(), m, mode);
            for (int c = 0; c < m; c++)
                for (int r = 0; r < m; r++) A[g(r) + lda * g(c)] = sub[r + m * c];
            return;
        }
    }
    std::vector<double> M(n * n), d(n), V(n * n);
    for (int c = 0; c < n; c++)
        for (int r = 0; r < n; r++) M[r + n * c] = A[r + lda * c];
    if (!eig_sym(n, M.data(), d.data(), V.data())) throw std::runtime_error("unable to project matrix onto positive semi-definite cone");
    if (d[0] >= 0.0) return; // A returned unchanged, bit for bit (:84-86)
    for (int i = 0; i < n; i++) {
        if (d[i] < 0.0) {
            d[i] = mode == 1 ? 0.0 : std::abs(d[i]);
        } else {
            break;
        }
    }
    for (int c = 0; c < n; c++)
        for (int r = 0; r < n; r++) {
            double s = 0;
            for (int k = 0; k < n; k++) s += V[r + n * k] * d[k] * V[c + n * k];
            A[r + lda * c] = s;
        }
}

} // namespace oracle
