// oracle/geom.hpp — TEST INFRASTRUCTURE ONLY (CPU restatement, never shipped).
//
// Plain C++17 restatement of the reference's distance layer
// (src/ipc/distance/*.cpp of ipc-toolkit v1.6.0).  The reference's gradients
// and Hessians are MATLAB-generated scalar code (namespace autogen); here they
// are written as closed-form vector calculus of the same functions and are
// pinned in oracle/selftest.cpp against second-order automatic differentiation
// (oracle/hyperdual.hpp) of the reference's value formulas, finite differences
// and the reference's own known-answer tests.
//
// Floating point: compiled with -ffp-contract=off.  Three-term sums follow
// Eigen's unrolled reduction order  e0 + (e1 + e2)  (Eigen redux_novec_unroller
// splits a length-3 reduction as 1 + 2) so that dot products and squared norms
// round like the reference's Eigen expressions do without FMA contraction.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <limits>
#include <stdexcept>

namespace oracle {

struct V3 {
    double x, y, z;
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3 operator*(double s, V3 a) { return { s * a.x, s * a.y, s * a.z }; }
inline V3 operator*(V3 a, double s) { return { a.x * s, a.y * s, a.z * s }; }
inline bool operator==(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline double sum3(double a, double b, double c) { return a + (b + c); }
inline double dot(V3 a, V3 b) { return sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline double sqnorm(V3 a) { return dot(a, a); }
inline V3 cross(V3 a, V3 b)
{
    return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
}

// ---------------------------------------------------------------------------
// Distance types: distance/distance_type.hpp:14-55 (uint8 values in order)
enum PE : uint8_t { PE_P_E0 = 0, PE_P_E1, PE_P_E, PE_AUTO };
enum PT : uint8_t { PT_P_T0 = 0, PT_P_T1, PT_P_T2, PT_P_E0, PT_P_E1, PT_P_E2, PT_P_T, PT_AUTO };
enum EE : uint8_t { EE_EA0_EB0 = 0, EE_EA0_EB1, EE_EA1_EB0, EE_EA1_EB1, EE_EA_EB0, EE_EA_EB1, EE_EA0_EB, EE_EA1_EB, EE_EA_EB, EE_AUTO };

// distance/distance_type.cpp:10-35
inline PE point_edge_distance_type(V3 p, V3 e0, V3 e1)
{
    const V3 e = e1 - e0;
    const double e_length_sqr = sqnorm(e);
    if (e_length_sqr == 0) {
        return PE_P_E0; // degenerate edge: arbitrary end-point (reference warns)
    }
    const double ratio = dot(e, p - e0) / e_length_sqr;
    if (ratio < 0) {
        return PE_P_E0;
    } else if (ratio > 1) {
        return PE_P_E1;
    }
    return PE_P_E;
}

// Eigen's LDLT (pivoted, in place, lower) on a symmetric 2x2 [[a,b],[b,c]]
// followed by solve(rhs): SURVEY Appendix B.2.  Pivot = first largest |diag|;
// D entries with |d| <= DBL_MIN give a zero solution component.
inline void ldlt2_solve(double a, double b, double c, double r0, double r1, double& x0, double& x1)
{
    const bool swap = std::abs(c) > std::abs(a); // first max wins ties
    if (swap) {
        std::swap(a, c);
        std::swap(r0, r1);
    }
    double l = 0, d0 = a, d1 = c;
    if (std::abs(d0) > 0) {
        l = b / d0;
        d1 = c - l * (d0 * l);
    } else {
        // entire diagonal is zero (k == 0 and invalid pivot): D = 0
        d1 = 0;
        l = 0;
    }
    // forward substitution with unit lower L
    double y0 = r0, y1 = r1 - l * y0;
    const double tol = std::numeric_limits<double>::min();
    y0 = std::abs(d0) > tol ? y0 / d0 : 0.0;
    y1 = std::abs(d1) > tol ? y1 / d1 : 0.0;
    // backward substitution with L^T
    const double z1 = y1, z0 = y0 - l * z1;
    if (swap) {
        x0 = z1;
        x1 = z0;
    } else {
        x0 = z0;
        x1 = z1;
    }
}

// distance/distance_type.cpp:37-83
inline PT point_triangle_distance_type(V3 p, V3 t0, V3 t1, V3 t2)
{
    const V3 normal = cross(t1 - t0, t2 - t0);
    const V3 tv[3] = { t0, t1, t2 };
    double param0[3], param1[3];
    for (int k = 0; k < 3; k++) {
        const V3 b0 = tv[(k + 1) % 3] - tv[k];
        const V3 b1 = cross(b0, normal);
        const V3 rel = p - tv[k];
        ldlt2_solve(dot(b0, b0), dot(b0, b1), dot(b1, b1), dot(b0, rel), dot(b1, rel), param0[k], param1[k]);
        if (param0[k] > 0.0 && param0[k] < 1.0 && param1[k] >= 0.0) {
            return PT(PT_P_E0 + k);
        }
    }
    if (param0[0] <= 0.0 && param0[2] >= 1.0) {
        return PT_P_T0;
    } else if (param0[1] <= 0.0 && param0[0] >= 1.0) {
        return PT_P_T1;
    } else if (param0[2] <= 0.0 && param0[1] >= 1.0) {
        return PT_P_T2;
    }
    return PT_P_T;
}

// distance/distance_type.cpp:170-211
inline EE edge_edge_parallel_distance_type(V3 ea0, V3 ea1, V3 eb0, V3 eb1)
{
    const V3 ea = ea1 - ea0;
    const double alpha = dot(eb0 - ea0, ea) / sqnorm(ea);
    const double beta = dot(eb1 - ea0, ea) / sqnorm(ea);
    uint8_t eac, ebc; // 0: E*0, 1: E*1, 2: interior
    if (alpha < 0) {
        eac = (0 <= beta && beta <= 1) ? 2 : 0;
        ebc = (beta <= alpha) ? 0 : (beta <= 1 ? 1 : 2);
    } else if (alpha > 1) {
        eac = (0 <= beta && beta <= 1) ? 2 : 1;
        ebc = (beta >= alpha) ? 0 : (0 <= beta ? 1 : 2);
    } else {
        eac = 2;
        ebc = 0;
    }
    return EE(ebc < 2 ? (eac << 1 | ebc) : (6 + eac));
}

// distance/distance_type.cpp:85-168
inline EE edge_edge_distance_type(V3 ea0, V3 ea1, V3 eb0, V3 eb1)
{
    constexpr double PARALLEL_THRESHOLD = 2.5e-16;
    const V3 u = ea1 - ea0, v = eb1 - eb0, w = ea0 - eb0;
    const double a = sqnorm(u), b = dot(u, v), c = sqnorm(v), d = dot(u, w), e = dot(v, w);
    const double D = a * c - b * b;

    if (a == 0.0 && c == 0.0) {
        return EE_EA0_EB0;
    } else if (a == 0.0) {
        return EE_EA0_EB;
    } else if (c == 0.0) {
        return EE_EA_EB0;
    }

    const double parallel_tolerance = PARALLEL_THRESHOLD * a * c;
    const double cross_sqnorm = sqnorm(cross(u, v));
    if (cross_sqnorm < parallel_tolerance) {
        return edge_edge_parallel_distance_type(ea0, ea1, eb0, eb1);
    }

    EE default_case = EE_EA_EB;
    const double sN = (b * e - c * d);
    double tN, tD;
    if (sN <= 0.0) {
        tN = e;
        tD = c;
        default_case = EE_EA0_EB;
    } else if (sN >= D) {
        tN = e + b;
        tD = c;
        default_case = EE_EA1_EB;
    } else {
        tN = (a * e - b * d);
        tD = D;
        if (tN > 0.0 && tN < tD && cross_sqnorm < parallel_tolerance) {
            if (sN < D / 2) {
                tN = e;
                tD = c;
                default_case = EE_EA0_EB;
            } else {
                tN = e + b;
                tD = c;
                default_case = EE_EA1_EB;
            }
        }
    }

    if (tN <= 0.0) {
        if (-d <= 0.0) {
            return EE_EA0_EB0;
        } else if (-d >= a) {
            return EE_EA1_EB0;
        }
        return EE_EA_EB0;
    } else if (tN >= tD) {
        if ((-d + b) <= 0.0) {
            return EE_EA0_EB1;
        } else if ((-d + b) >= a) {
            return EE_EA1_EB1;
        }
        return EE_EA_EB1;
    }
    return default_case;
}

// ---------------------------------------------------------------------------
// Squared distances (values): point_point.cpp:5-9, point_line.cpp:5-23,
// point_plane.cpp:10-26 (+ geometry/normal.hpp:149-168: NORMALISED normal),
// line_line.cpp:7-16
inline double point_point_distance(V3 p0, V3 p1) { return sqnorm(p1 - p0); }
inline double point_line_distance(V3 p, V3 e0, V3 e1)
{
    return sqnorm(cross(e0 - p, e1 - p)) / sqnorm(e1 - e0);
}
inline double point_plane_distance(V3 p, V3 t0, V3 t1, V3 t2)
{
    const V3 n = cross(t1 - t0, t2 - t0);
    const double len = std::sqrt(sqnorm(n));
    const V3 nh = { n.x / len, n.y / len, n.z / len }; // Eigen normalized(): n / norm
    const double s = dot(p - t0, nh);
    return s * s / sqnorm(nh);
}
inline double line_line_distance(V3 ea0, V3 ea1, V3 eb0, V3 eb1)
{
    const V3 n = cross(ea1 - ea0, eb1 - eb0);
    const double s = dot(eb0 - ea0, n);
    return s * s / sqnorm(n);
}

// ---------------------------------------------------------------------------
// Primitive kinds and the stencil embedding tables used by the dispatchers
// (point_edge.cpp:10-122, point_triangle.cpp:11-186, edge_edge.cpp:13-221):
// each distance type is a PP / PL / plane-like primitive over a subset of the
// stencil's points; idx lists which stencil points feed the primitive, in
// the primitive's argument order.
enum Prim : uint8_t { PRIM_PP, PRIM_PL, PRIM_PLANE, PRIM_LL };
struct Embed {
    Prim prim;
    int n;      // points used
    int idx[4]; // stencil point of primitive argument k
};
inline Embed embed_point_edge(PE t)
{
    switch (t) {
    case PE_P_E0: return { PRIM_PP, 2, { 0, 1, 0, 0 } };
    case PE_P_E1: return { PRIM_PP, 2, { 0, 2, 0, 0 } };
    case PE_P_E: return { PRIM_PL, 3, { 0, 1, 2, 0 } };
    default: throw std::invalid_argument("Invalid distance type for point-edge distance!");
    }
}
inline Embed embed_point_triangle(PT t)
{
    switch (t) {
    case PT_P_T0: return { PRIM_PP, 2, { 0, 1, 0, 0 } };
    case PT_P_T1: return { PRIM_PP, 2, { 0, 2, 0, 0 } };
    case PT_P_T2: return { PRIM_PP, 2, { 0, 3, 0, 0 } };
    case PT_P_E0: return { PRIM_PL, 3, { 0, 1, 2, 0 } };
    case PT_P_E1: return { PRIM_PL, 3, { 0, 2, 3, 0 } };
    case PT_P_E2: return { PRIM_PL, 3, { 0, 3, 1, 0 } };
    case PT_P_T: return { PRIM_PLANE, 4, { 0, 1, 2, 3 } };
    default: throw std::invalid_argument("Invalid distance type for point-triangle distance!");
    }
}
inline Embed embed_edge_edge(EE t)
{
    switch (t) {
    case EE_EA0_EB0: return { PRIM_PP, 2, { 0, 2, 0, 0 } };
    case EE_EA0_EB1: return { PRIM_PP, 2, { 0, 3, 0, 0 } };
    case EE_EA1_EB0: return { PRIM_PP, 2, { 1, 2, 0, 0 } };
    case EE_EA1_EB1: return { PRIM_PP, 2, { 1, 3, 0, 0 } };
    case EE_EA_EB0: return { PRIM_PL, 3, { 2, 0, 1, 0 } };
    case EE_EA_EB1: return { PRIM_PL, 3, { 3, 0, 1, 0 } };
    case EE_EA0_EB: return { PRIM_PL, 3, { 0, 2, 3, 0 } };
    case EE_EA1_EB: return { PRIM_PL, 3, { 1, 2, 3, 0 } };
    case EE_EA_EB: return { PRIM_LL, 4, { 0, 1, 2, 3 } };
    default: throw std::invalid_argument("Invalid distance type for edge-edge distance!");
    }
}

inline double prim_value(const Embed& em, const V3* x)
{
    const int* i = em.idx;
    switch (em.prim) {
    case PRIM_PP: return point_point_distance(x[i[0]], x[i[1]]);
    case PRIM_PL: return point_line_distance(x[i[0]], x[i[1]], x[i[2]]);
    case PRIM_PLANE: return point_plane_distance(x[i[0]], x[i[1]], x[i[2]], x[i[3]]);
    default: return line_line_distance(x[i[0]], x[i[1]], x[i[2]], x[i[3]]);
    }
}

inline double point_edge_distance(V3 p, V3 e0, V3 e1, PE t = PE_AUTO)
{
    if (t == PE_AUTO) t = point_edge_distance_type(p, e0, e1);
    const V3 x[3] = { p, e0, e1 };
    return prim_value(embed_point_edge(t), x);
}
inline double point_triangle_distance(V3 p, V3 t0, V3 t1, V3 t2, PT t = PT_AUTO)
{
    if (t == PT_AUTO) t = point_triangle_distance_type(p, t0, t1, t2);
    const V3 x[4] = { p, t0, t1, t2 };
    return prim_value(embed_point_triangle(t), x);
}
inline double edge_edge_distance(V3 ea0, V3 ea1, V3 eb0, V3 eb1, EE t = EE_AUTO)
{
    if (t == EE_AUTO) t = edge_edge_distance_type(ea0, ea1, eb0, eb1);
    const V3 x[4] = { ea0, ea1, eb0, eb1 };
    return prim_value(embed_edge_edge(t), x);
}

// ---------------------------------------------------------------------------
// Derivatives.  A primitive is a function of m difference vectors
// (each = x[plus] - x[minus]); its gradient / Hessian are formed in
// difference coordinates and pulled back to the primitive's points.
struct Blk3 { // 3x3 block, row-major
    double a[9];
};
inline Blk3 outer(V3 u, V3 v)
{
    return { { u.x * v.x, u.x * v.y, u.x * v.z, u.y * v.x, u.y * v.y, u.y * v.z, u.z * v.x, u.z * v.y, u.z * v.z } };
}
inline Blk3 skew(V3 w) // [w]x : [w]x y = w x y
{
    return { { 0, -w.z, w.y, w.z, 0, -w.x, -w.y, w.x, 0 } };
}
inline Blk3 ident(double s) { return { { s, 0, 0, 0, s, 0, 0, 0, s } }; }
inline Blk3 operator+(Blk3 a, Blk3 b)
{
    Blk3 r;
    for (int i = 0; i < 9; i++) r.a[i] = a.a[i] + b.a[i];
    return r;
}
inline Blk3 operator-(Blk3 a, Blk3 b)
{
    Blk3 r;
    for (int i = 0; i < 9; i++) r.a[i] = a.a[i] - b.a[i];
    return r;
}
inline Blk3 operator*(double s, Blk3 a)
{
    for (int i = 0; i < 9; i++) a.a[i] *= s;
    return a;
}
inline Blk3 transpose(Blk3 a)
{
    return { { a.a[0], a.a[3], a.a[6], a.a[1], a.a[4], a.a[7], a.a[2], a.a[5], a.a[8] } };
}

// Local derivative container for up to 4 points: g[12], H[12*12] col-major
// with leading dimension 12 (only the leading 3n x 3n part is meaningful).
struct Deriv {
    double val;
    double g[12];
    double H[144];
    void zero()
    {
        val = 0;
        std::memset(g, 0, sizeof g);
        std::memset(H, 0, sizeof H);
    }
    double& h(int r, int c) { return H[r + 12 * c]; }
    double h(int r, int c) const { return H[r + 12 * c]; }
};

// Pull back (value, gd[3m], Hd blocks[m][m]) from difference vectors
// d_k = x[plus[k]] - x[minus[k]] onto points 0..np-1.
inline void pullback(int m, const int* plus, const int* minus, const V3* gd, const Blk3 (*Hd)[3], Deriv& out)
{
    for (int k = 0; k < m; k++) {
        for (int c = 0; c < 3; c++) {
            out.g[3 * plus[k] + c] += gd[k][c];
            out.g[3 * minus[k] + c] -= gd[k][c];
        }
    }
    for (int k = 0; k < m; k++) {
        for (int l = 0; l < m; l++) {
            const Blk3& B = Hd[k][l];
            const int pi[2] = { plus[k], minus[k] }, pj[2] = { plus[l], minus[l] };
            for (int si = 0; si < 2; si++) {
                for (int sj = 0; sj < 2; sj++) {
                    const double sgn = (si == sj) ? 1.0 : -1.0;
                    for (int r = 0; r < 3; r++)
                        for (int c = 0; c < 3; c++)
                            out.h(3 * pi[si] + r, 3 * pj[sj] + c) += sgn * B.a[3 * r + c];
                }
            }
        }
    }
}

// d^2 = |p1 - p0|^2  (point_point.cpp:11-40)
inline void point_point_deriv(V3 p0, V3 p1, Deriv& out)
{
    out.zero();
    const V3 r = p0 - p1;
    out.val = sqnorm(p1 - p0);
    for (int c = 0; c < 3; c++) {
        out.g[c] = 2.0 * r[c];
        out.g[3 + c] = -out.g[c];
        out.h(c, c) = 2.0;
        out.h(3 + c, 3 + c) = 2.0;
        out.h(c, 3 + c) = out.h(3 + c, c) = -2.0;
    }
}

// d^2 = |a x b|^2 / |b - a|^2,  a = e0 - p, b = e1 - p  (points: p, e0, e1)
inline void point_line_deriv(V3 p, V3 e0, V3 e1, Deriv& out)
{
    out.zero();
    const V3 a = e0 - p, b = e1 - p, c = cross(a, b), e = e1 - e0;
    const double N = sqnorm(c), L = sqnorm(e);
    out.val = N / L;
    // differences: d0 = a (e0 - p), d1 = b (e1 - p); e = d1 - d0
    const V3 gNa = 2.0 * cross(b, c), gNb = 2.0 * cross(c, a);
    const V3 gLa = -2.0 * e, gLb = 2.0 * e;
    const double aa = sqnorm(a), bb = sqnorm(b), ab = dot(a, b);
    Blk3 HN[2][2];
    HN[0][0] = 2.0 * (ident(bb) - outer(b, b));
    HN[1][1] = 2.0 * (ident(aa) - outer(a, a));
    HN[0][1] = 2.0 * (2.0 * outer(a, b) - outer(b, a) - ident(ab));
    HN[1][0] = transpose(HN[0][1]);
    const V3 gN[2] = { gNa, gNb }, gL[2] = { gLa, gLb };
    const double HLs[2][2] = { { 2, -2 }, { -2, 2 } };
    const double iL = 1.0 / L, NL2 = N / (L * L), NL3 = 2.0 * N / (L * L * L);
    V3 gd[3];
    Blk3 Hd[3][3];
    for (int k = 0; k < 2; k++) {
        gd[k] = iL * gN[k] - NL2 * gL[k];
        for (int l = 0; l < 2; l++) {
            Hd[k][l] = iL * HN[k][l] - (iL * iL) * (outer(gN[k], gL[l]) + outer(gL[k], gN[l])) - ident(NL2 * HLs[k][l])
                + NL3 * outer(gL[k], gL[l]);
        }
    }
    const int plus[2] = { 1, 2 }, minus[2] = { 0, 0 };
    pullback(2, plus, minus, gd, Hd, out);
}

// T(q,u,v) = (q . (u x v))^2 / |u x v|^2 in difference coordinates; shared by
// point-plane (q = p - t0, u = t1 - t0, v = t2 - t0) and line-line
// (q = eb0 - ea0, u = ea1 - ea0, v = eb1 - eb0).
inline double triple_deriv(V3 q, V3 u, V3 v, V3 gd[3], Blk3 Hd[3][3])
{
    const V3 n = cross(u, v);
    const double s = dot(q, n), M = sqnorm(n);
    const V3 gs[3] = { n, cross(v, q), cross(q, u) };
    const V3 zero = { 0, 0, 0 };
    const V3 gM[3] = { zero, 2.0 * cross(v, n), 2.0 * cross(n, u) };
    Blk3 Hs[3][3], HM[3][3];
    const Blk3 Z = ident(0);
    Hs[0][0] = Z, Hs[1][1] = Z, Hs[2][2] = Z;
    Hs[0][1] = -1.0 * skew(v), Hs[0][2] = skew(u);
    Hs[1][0] = skew(v), Hs[1][2] = -1.0 * skew(q);
    Hs[2][0] = -1.0 * skew(u), Hs[2][1] = skew(q);
    const double uu = sqnorm(u), vv = sqnorm(v), uv = dot(u, v);
    for (int k = 0; k < 3; k++) HM[0][k] = HM[k][0] = Z;
    HM[1][1] = 2.0 * (ident(vv) - outer(v, v));
    HM[2][2] = 2.0 * (ident(uu) - outer(u, u));
    HM[1][2] = 2.0 * (2.0 * outer(u, v) - outer(v, u) - ident(uv));
    HM[2][1] = transpose(HM[1][2]);
    const double iM = 1.0 / M, c1 = 2.0 * s * iM, c2 = s * s * iM * iM, c3 = 2.0 * s * iM * iM,
                 c4 = 2.0 * s * s * iM * iM * iM;
    for (int k = 0; k < 3; k++) {
        gd[k] = c1 * gs[k] - c2 * gM[k];
        for (int l = 0; l < 3; l++) {
            Hd[k][l] = (2.0 * iM) * outer(gs[k], gs[l]) + c1 * Hs[k][l] - c3 * (outer(gs[k], gM[l]) + outer(gM[k], gs[l]))
                - c2 * HM[k][l] + c4 * outer(gM[k], gM[l]);
        }
    }
    return s * s * iM;
}

inline void point_plane_deriv(V3 p, V3 t0, V3 t1, V3 t2, Deriv& out)
{
    out.zero();
    V3 gd[3];
    Blk3 Hd[3][3];
    triple_deriv(p - t0, t1 - t0, t2 - t0, gd, Hd);
    out.val = point_plane_distance(p, t0, t1, t2);
    const int plus[3] = { 0, 2, 3 }, minus[3] = { 1, 1, 1 };
    pullback(3, plus, minus, gd, Hd, out);
}

inline void line_line_deriv(V3 ea0, V3 ea1, V3 eb0, V3 eb1, Deriv& out)
{
    out.zero();
    V3 gd[3];
    Blk3 Hd[3][3];
    triple_deriv(eb0 - ea0, ea1 - ea0, eb1 - eb0, gd, Hd);
    out.val = line_line_distance(ea0, ea1, eb0, eb1);
    const int plus[3] = { 2, 1, 3 }, minus[3] = { 0, 0, 2 };
    pullback(3, plus, minus, gd, Hd, out);
}

// Evaluate an embedded primitive and scatter its derivatives into the stencil.
inline void embed_deriv(const Embed& em, const V3* x, Deriv& out)
{
    Deriv loc;
    const int* i = em.idx;
    switch (em.prim) {
    case PRIM_PP: point_point_deriv(x[i[0]], x[i[1]], loc); break;
    case PRIM_PL: point_line_deriv(x[i[0]], x[i[1]], x[i[2]], loc); break;
    case PRIM_PLANE: point_plane_deriv(x[i[0]], x[i[1]], x[i[2]], x[i[3]], loc); break;
    default: line_line_deriv(x[i[0]], x[i[1]], x[i[2]], x[i[3]], loc); break;
    }
    out.zero();
    out.val = loc.val;
    for (int a = 0; a < em.n; a++) {
        for (int r = 0; r < 3; r++) out.g[3 * i[a] + r] = loc.g[3 * a + r];
        for (int b = 0; b < em.n; b++)
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) out.h(3 * i[a] + r, 3 * i[b] + c) = loc.h(3 * a + r, 3 * b + c);
    }
}

// ---------------------------------------------------------------------------
// Edge-edge mollifier: distance/edge_edge_mollifier.cpp:7-135,194-202
inline double edge_edge_cross_squarednorm(V3 ea0, V3 ea1, V3 eb0, V3 eb1)
{
    return sqnorm(cross(ea1 - ea0, eb1 - eb0));
}
// s = |u x v|^2, u = ea1 - ea0, v = eb1 - eb0 (points ea0, ea1, eb0, eb1)
inline void edge_edge_cross_squarednorm_deriv(V3 ea0, V3 ea1, V3 eb0, V3 eb1, Deriv& out)
{
    out.zero();
    const V3 u = ea1 - ea0, v = eb1 - eb0, n = cross(u, v);
    out.val = sqnorm(n);
    const double uu = sqnorm(u), vv = sqnorm(v), uv = dot(u, v);
    V3 gd[3] = { 2.0 * cross(v, n), 2.0 * cross(n, u), { 0, 0, 0 } };
    Blk3 Hd[3][3];
    Hd[0][0] = 2.0 * (ident(vv) - outer(v, v));
    Hd[1][1] = 2.0 * (ident(uu) - outer(u, u));
    Hd[0][1] = 2.0 * (2.0 * outer(u, v) - outer(v, u) - ident(uv));
    Hd[1][0] = transpose(Hd[0][1]);
    const int plus[2] = { 1, 3 }, minus[2] = { 0, 2 };
    pullback(2, plus, minus, gd, Hd, out);
}
inline double edge_edge_mollifier(double x, double eps_x)
{
    if (x < eps_x) {
        const double q = x / eps_x;
        return (-q + 2) * q;
    }
    return 1;
}
inline double edge_edge_mollifier_gradient(double x, double eps_x)
{
    if (x < eps_x) {
        const double one_div_eps_x = 1 / eps_x;
        return 2 * one_div_eps_x * std::fma(-one_div_eps_x, x, 1); // reference uses fma()
    }
    return 0;
}
inline double edge_edge_mollifier_hessian(double x, double eps_x) { return x < eps_x ? -2 / (eps_x * eps_x) : 0; }
inline double edge_edge_mollifier_threshold(V3 ea0r, V3 ea1r, V3 eb0r, V3 eb1r)
{
    return 1e-3 * sqnorm(ea0r - ea1r) * sqnorm(eb0r - eb1r);
}

// ---------------------------------------------------------------------------
// Barrier: barrier/barrier.cpp:11-43
inline double barrier(double d, double dhat)
{
    if (d <= 0.0) return std::numeric_limits<double>::infinity();
    if (d >= dhat) return 0;
    const double t = d - dhat;
    return -t * t * std::log(d / dhat);
}
inline double barrier_first_derivative(double d, double dhat)
{
    if (d <= 0.0 || d >= dhat) return 0.0;
    return (dhat - d) * (2 * std::log(d / dhat) - dhat / d + 1);
}
inline double barrier_second_derivative(double d, double dhat)
{
    if (d <= 0.0 || d >= dhat) return 0.0;
    const double dhat_d = dhat / d;
    return (dhat_d + 2) * dhat_d - 2 * std::log(d / dhat) - 3;
}

} // namespace oracle
