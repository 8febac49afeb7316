// oracle/selftest.cpp — TEST INFRASTRUCTURE ONLY.
//
// Pins the oracle's hand-derived derivatives against second-order automatic
// differentiation of the reference's value formulas, its eigen-solver against
// reconstruction / known answers (tests/src/tests/utils/test_utils.cpp:27-48),
// and the Morton code against the reference header's own output
// (morton_3D(0.5,0.25,0.75) == 0x5600000000000000).
// Exit code 0 = all checks passed.  Run by tests/test_oracle_selftest.py.
#include "geom.hpp"
#include "eig.hpp"
#include "ccd.hpp"
#include "hyperdual.hpp"
#include <cstdio>
#include <random>

using namespace oracle;
using H12 = HD<12>;
using HV = HV3<12>;

static int failures = 0;
#define CHECK(cond, ...)                                                                                              \
    do {                                                                                                              \
        if (!(cond)) {                                                                                                \
            failures++;                                                                                               \
            std::printf("FAIL %s:%d: %s | ", __FILE__, __LINE__, #cond);                                              \
            std::printf(__VA_ARGS__);                                                                                 \
            std::printf("\n");                                                                                        \
        }                                                                                                             \
    } while (0)

static HV hv(const V3& p, int point) { return { H12::var(p.x, 3 * point), H12::var(p.y, 3 * point + 1), H12::var(p.z, 3 * point + 2) }; }

// reference value formulas in autodiff arithmetic
static H12 ad_pp(const HV* x) { return dot(x[1] - x[0], x[1] - x[0]); }
static H12 ad_pl(const HV* x)
{
    const HV c = cross(x[1] - x[0], x[2] - x[0]);
    return dot(c, c) / dot(x[2] - x[1], x[2] - x[1]);
}
static H12 ad_plane(const HV* x)
{
    const HV n = cross(x[2] - x[1], x[3] - x[1]);
    const H12 s = dot(x[0] - x[1], n);
    return s * s / dot(n, n);
}
static H12 ad_ll(const HV* x)
{
    const HV n = cross(x[1] - x[0], x[3] - x[2]);
    const H12 s = dot(x[2] - x[0], n);
    return s * s / dot(n, n);
}
static H12 ad_cross(const HV* x)
{
    const HV n = cross(x[1] - x[0], x[3] - x[2]);
    return dot(n, n);
}

static void compare(const char* name, const Deriv& D, const H12& ad, int npts, double tol = 1e-10)
{
    const int n = 3 * npts;
    double gn = 0, gd = 0, hn = 0, hd = 0;
    for (int i = 0; i < n; i++) {
        gn += ad.g[i] * ad.g[i];
        gd += (D.g[i] - ad.g[i]) * (D.g[i] - ad.g[i]);
        for (int j = 0; j < n; j++) {
            hn += ad.h[i * 12 + j] * ad.h[i * 12 + j];
            const double e = D.h(i, j) - ad.h[i * 12 + j];
            hd += e * e;
        }
    }
    CHECK(std::abs(D.val - ad.v) <= tol * std::abs(ad.v), "%s value %g vs %g", name, D.val, ad.v);
    CHECK(std::sqrt(gd) <= tol * std::sqrt(gn), "%s gradient rel err %g", name, std::sqrt(gd / gn));
    CHECK(std::sqrt(hd) <= tol * std::sqrt(hn), "%s hessian rel err %g", name, std::sqrt(hd / hn));
    // symmetry
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++)
            CHECK(std::abs(D.h(i, j) - D.h(j, i)) <= 1e-12 * std::sqrt(hn), "%s hessian asymmetric at %d,%d", name, i, j);
}

int main()
{
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(-1, 1);
    for (int trial = 0; trial < 200; trial++) {
        const double scale = std::pow(10.0, -3.0 * (trial % 3)); // also exercise small stencils
        V3 x[4];
        HV hx[4];
        for (int k = 0; k < 4; k++) {
            x[k] = { scale * U(rng), scale * U(rng), scale * U(rng) };
            hx[k] = hv(x[k], k);
        }
        Deriv D;
        point_point_deriv(x[0], x[1], D);
        compare("point_point", D, ad_pp(hx), 2);
        point_line_deriv(x[0], x[1], x[2], D);
        compare("point_line", D, ad_pl(hx), 3);
        point_plane_deriv(x[0], x[1], x[2], x[3], D);
        compare("point_plane", D, ad_plane(hx), 4);
        line_line_deriv(x[0], x[1], x[2], x[3], D);
        compare("line_line", D, ad_ll(hx), 4);
        edge_edge_cross_squarednorm_deriv(x[0], x[1], x[2], x[3], D);
        compare("cross_sqnorm", D, ad_cross(hx), 4);

        // embedding tables: every dtype's embedded derivative equals the
        // autodiff of the primitive on the permuted points
        for (int t = 0; t < 9; t++) {
            const Embed em = embed_edge_edge(EE(t));
            embed_deriv(em, x, D);
            HV px[4];
            for (int a = 0; a < em.n; a++) px[a] = hx[em.idx[a]];
            const H12 ad = em.prim == PRIM_PP ? ad_pp(px) : em.prim == PRIM_PL ? ad_pl(px) : ad_ll(px);
            compare("edge_edge embed", D, ad, 4);
        }
        for (int t = 0; t < 7; t++) {
            const Embed em = embed_point_triangle(PT(t));
            embed_deriv(em, x, D);
            HV px[4];
            for (int a = 0; a < em.n; a++) px[a] = hx[em.idx[a]];
            const H12 ad = em.prim == PRIM_PP ? ad_pp(px) : em.prim == PRIM_PL ? ad_pl(px) : ad_plane(px);
            compare("point_triangle embed", D, ad, 4);
        }
    }

    // eigen solver: reconstruction + orthogonality on random symmetric matrices
    for (int n : { 2, 3, 6, 9, 12 }) {
        for (int trial = 0; trial < 50; trial++) {
            std::vector<double> A(n * n), d(n), V(n * n);
            for (int i = 0; i < n; i++)
                for (int j = 0; j <= i; j++) A[i + n * j] = A[j + n * i] = U(rng);
            CHECK(eig_sym(n, A.data(), d.data(), V.data()), "eig_sym failed n=%d", n);
            double err = 0, orth = 0, nrm = 0;
            for (int i = 0; i < n; i++)
                for (int j = 0; j < n; j++) {
                    double s = 0, o = 0;
                    for (int k = 0; k < n; k++) {
                        s += V[i + n * k] * d[k] * V[j + n * k];
                        o += V[k + n * i] * V[k + n * j];
                    }
                    err += (s - A[i + n * j]) * (s - A[i + n * j]);
                    nrm += A[i + n * j] * A[i + n * j];
                    orth += (o - (i == j)) * (o - (i == j));
                }
            CHECK(std::sqrt(err) <= 1e-13 * std::sqrt(nrm) * n, "eig reconstruction n=%d err=%g", n, std::sqrt(err / nrm));
            CHECK(std::sqrt(orth) <= 1e-13 * n, "eig orthogonality n=%d err=%g", n, std::sqrt(orth));
            for (int i = 1; i < n; i++) CHECK(d[i - 1] <= d[i], "eigenvalues not ascending");
        }
    }
    // project_to_psd known answers (tests/src/tests/utils/test_utils.cpp:27-48)
    {
        double I2[4] = { 1, 0, 0, 1 };
        project_to_psd(2, I2, 2, 1);
        CHECK(I2[0] == 1 && I2[1] == 0 && I2[2] == 0 && I2[3] == 1, "psd(I) != I");
        double N2[4] = { -1, 0, 0, -1 };
        project_to_psd(2, N2, 2, 1);
        for (double v : N2) CHECK(std::abs(v) < 1e-15, "psd(-I) != 0");
        double A2[4] = { 2, 1, 1, 2 };
        project_to_psd(2, A2, 2, 1);
        CHECK(A2[0] == 2 && A2[1] == 1 && A2[2] == 1 && A2[3] == 2, "psd([[2,1],[1,2]]) changed");
        double B2[4] = { 1, 2, 2, 1 }; // eigenvalues -1, 3 -> clamp gives 1.5*[[1,1],[1,1]]
        project_to_psd(2, B2, 2, 1);
        for (double v : B2) CHECK(std::abs(v - 1.5) < 1e-14, "psd clamp value %g", v);
        double C2[4] = { 1, 2, 2, 1 }; // abs: eigenvalues 1, 3 -> [[2,1],[1,2]]
        project_to_psd(2, C2, 2, 2);
        CHECK(std::abs(C2[0] - 2) < 1e-14 && std::abs(C2[1] - 1) < 1e-14, "psd abs value");
    }
    // TI root finder: head-on point-triangle, analytic toi = 0.5 - ms-ish
    {
        const V3 s[4] = { { 0.25, 0.25, 1 }, { 0, 0, 0 }, { 1, 0, 0 }, { 0, 1, 0 } };
        const V3 e[4] = { { 0.25, 0.25, -1 }, { 0, 0, 0 }, { 1, 0, 0 }, { 0, 1, 0 } };
        TightInclusionCCD ti;
        double toi = -1;
        const bool hit = ti.point_triangle_ccd(s, e, toi, 0.0, 1.0);
        CHECK(hit, "TI point-triangle head-on missed");
        CHECK(toi <= 0.5 && toi > 0.5 - 1e-3, "TI toi %g", toi);
        AdditiveCCD ac;
        double toi2 = -1;
        CHECK(ac.point_triangle_ccd(s, e, toi2, 0.0, 1.0), "ACCD missed");
        CHECK(toi2 <= 0.5 && toi2 > 0.4, "ACCD toi %g", toi2);
        const V3 e_far[4] = { { 0.25, 0.25, 0.5 }, { 0, 0, 0 }, { 1, 0, 0 }, { 0, 1, 0 } };
        CHECK(!ti.point_triangle_ccd(s, e_far, toi, 0.0, 1.0), "TI false positive");
    }
    if (failures == 0) std::printf("oracle selftest: all checks passed\n");
    return failures ? 1 : 0;
}
