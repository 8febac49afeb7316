// geom.cuh — device-side distance layer of the contact path.
//
// Replaces (reference src/ipc/): distance/distance_type.cpp:10-211,
// distance/{point_point,point_line,point_plane,line_line,point_edge,
// point_triangle,edge_edge}.cpp, distance/edge_edge_mollifier.cpp:7-135,194-202,
// barrier/barrier.cpp:11-43.
//
// The reference's gradients/Hessians are MATLAB-generated scalar code; here
// every primitive is differentiated in closed form in its DIFFERENCE vectors
// (a squared distance only depends on differences of its points) and pulled
// back to the stencil with +-1 block scatters.
//
// Rounding contract for classification (it decides the collision SET, which
// must match the CPU oracle bit for bit): this library is compiled with
// -fmad=false so * and + are never contracted, and three-term sums are
// evaluated as e0 + (e1 + e2) like Eigen's unrolled reductions.  Explicit
// fma() is used only where values (not decisions) are produced.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>

namespace ipcb {

struct d3 {
    double x, y, z;
};
__host__ __device__ inline d3 mk3(double x, double y, double z) { return d3 { x, y, z }; }
__host__ __device__ inline d3 operator+(d3 a, d3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
__host__ __device__ inline d3 operator-(d3 a, d3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
__host__ __device__ inline d3 operator*(double s, d3 a) { return { s * a.x, s * a.y, s * a.z }; }
__host__ __device__ inline double sum3(double a, double b, double c) { return a + (b + c); }
__host__ __device__ inline double dot(d3 a, d3 b) { return sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
__host__ __device__ inline double sqn(d3 a) { return dot(a, a); }
__host__ __device__ inline d3 cross(d3 a, d3 b)
{
    return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
}
__host__ __device__ inline bool same(d3 a, d3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
__host__ __device__ inline double comp(d3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
// one vertex = one 32-byte sector: two read-only 16-byte loads (x,y) and (z,pad)
__device__ inline d3 load_vertex(const double4* X, int i)
{
    const double2* p = reinterpret_cast<const double2*>(X + i);
    const double2 a = __ldg(p), b = __ldg(p + 1);
    return { a.x, a.y, b.x };
}

// distance types: distance/distance_type.hpp:14-55
enum : uint8_t { PE_E0 = 0, PE_E1 = 1, PE_E = 2 };
enum : uint8_t { PT_T0 = 0, PT_T1, PT_T2, PT_E0, PT_E1, PT_E2, PT_T };
enum : uint8_t { EE_A0B0 = 0, EE_A0B1, EE_A1B0, EE_A1B1, EE_AB0, EE_AB1, EE_A0B, EE_A1B, EE_AB };

// ---- classification ---------------------------------------------------------
__host__ __device__ inline int point_edge_type(d3 p, d3 e0, d3 e1) // distance_type.cpp:10-35
{
    const d3 e = e1 - e0;
    const double len2 = sqn(e);
    if (len2 == 0) return PE_E0;
    const double ratio = dot(e, p - e0) / len2;
    return ratio < 0 ? PE_E0 : (ratio > 1 ? PE_E1 : PE_E);
}

// pivoted 2x2 LDL^T solve with Eigen's conventions
__host__ __device__ inline void ldlt2(double a, double b, double c, double r0, double r1, double& x0, double& x1)
{
    const bool swp = fabs(c) > fabs(a);
    if (swp) {
        double t = a;
        a = c;
        c = t;
        t = r0;
        r0 = r1;
        r1 = t;
    }
    double l = 0, d0 = a, d1 = c;
    if (fabs(d0) > 0) {
        l = b / d0;
        d1 = c - l * (d0 * l);
    } else {
        d1 = 0;
    }
    double y0 = r0, y1 = r1 - l * y0;
    const double tiny = 2.2250738585072014e-308;
    y0 = fabs(d0) > tiny ? y0 / d0 : 0.0;
    y1 = fabs(d1) > tiny ? y1 / d1 : 0.0;
    const double z1 = y1, z0 = y0 - l * z1;
    x0 = swp ? z1 : z0;
    x1 = swp ? z0 : z1;
}

__host__ __device__ inline int point_triangle_type(d3 p, d3 t0, d3 t1, d3 t2) // distance_type.cpp:37-83
{
    const d3 n = cross(t1 - t0, t2 - t0);
    double a0, a1, a2, b;
    {
        const d3 e = t1 - t0, f = cross(e, n), r = p - t0;
        ldlt2(dot(e, e), dot(e, f), dot(f, f), dot(e, r), dot(f, r), a0, b);
        if (a0 > 0.0 && a0 < 1.0 && b >= 0.0) return PT_E0;
    }
    {
        const d3 e = t2 - t1, f = cross(e, n), r = p - t1;
        ldlt2(dot(e, e), dot(e, f), dot(f, f), dot(e, r), dot(f, r), a1, b);
        if (a1 > 0.0 && a1 < 1.0 && b >= 0.0) return PT_E1;
    }
    {
        const d3 e = t0 - t2, f = cross(e, n), r = p - t2;
        ldlt2(dot(e, e), dot(e, f), dot(f, f), dot(e, r), dot(f, r), a2, b);
        if (a2 > 0.0 && a2 < 1.0 && b >= 0.0) return PT_E2;
    }
    if (a0 <= 0.0 && a2 >= 1.0) return PT_T0;
    if (a1 <= 0.0 && a0 >= 1.0) return PT_T1;
    if (a2 <= 0.0 && a1 >= 1.0) return PT_T2;
    return PT_T;
}

__host__ __device__ inline int edge_edge_parallel_type(d3 ea0, d3 ea1, d3 eb0, d3 eb1) // distance_type.cpp:170-211
{
    const d3 ea = ea1 - ea0;
    const double alpha = dot(eb0 - ea0, ea) / sqn(ea);
    const double beta = dot(eb1 - ea0, ea) / sqn(ea);
    int eac, ebc;
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
    return ebc < 2 ? (eac << 1 | ebc) : (6 + eac);
}

__host__ __device__ inline int edge_edge_type(d3 ea0, d3 ea1, d3 eb0, d3 eb1) // distance_type.cpp:85-168
{
    const d3 u = ea1 - ea0, v = eb1 - eb0, w = ea0 - eb0;
    const double a = sqn(u), b = dot(u, v), c = sqn(v), d = dot(u, w), e = dot(v, w);
    const double D = a * c - b * b;
    if (a == 0.0 && c == 0.0) return EE_A0B0;
    if (a == 0.0) return EE_A0B;
    if (c == 0.0) return EE_AB0;
    const double par_tol = 2.5e-16 * a * c;
    const double cr = sqn(cross(u, v));
    if (cr < par_tol) return edge_edge_parallel_type(ea0, ea1, eb0, eb1);
    int dflt = EE_AB;
    const double sN = b * e - c * d;
    double tN, tD;
    if (sN <= 0.0) {
        tN = e, tD = c, dflt = EE_A0B;
    } else if (sN >= D) {
        tN = e + b, tD = c, dflt = EE_A1B;
    } else {
        tN = a * e - b * d, tD = D;
        if (tN > 0.0 && tN < tD && cr < par_tol) {
            if (sN < D / 2) {
                tN = e, tD = c, dflt = EE_A0B;
            } else {
                tN = e + b, tD = c, dflt = EE_A1B;
            }
        }
    }
    if (tN <= 0.0) {
        if (-d <= 0.0) return EE_A0B0;
        if (-d >= a) return EE_A1B0;
        return EE_AB0;
    } else if (tN >= tD) {
        if ((-d + b) <= 0.0) return EE_A0B1;
        if ((-d + b) >= a) return EE_A1B1;
        return EE_AB1;
    }
    return dflt;
}

// ---- primitives (values) ------------------------------------------------------
__host__ __device__ inline double pp_dist(d3 a, d3 b) { return sqn(b - a); }
__host__ __device__ inline double pl_dist(d3 p, d3 e0, d3 e1) { return sqn(cross(e0 - p, e1 - p)) / sqn(e1 - e0); }
__host__ __device__ inline double plane_dist(d3 p, d3 t0, d3 t1, d3 t2)
{
    // point_plane.cpp:10-26 with the NORMALISED triangle normal (geometry/normal.hpp:149-168)
    const d3 n = cross(t1 - t0, t2 - t0);
    const double len = sqrt(sqn(n));
    const d3 nh = { n.x / len, n.y / len, n.z / len };
    const double s = dot(p - t0, nh);
    return s * s / sqn(nh);
}
__host__ __device__ inline double ll_dist(d3 ea0, d3 ea1, d3 eb0, d3 eb1)
{
    const d3 n = cross(ea1 - ea0, eb1 - eb0);
    const double s = dot(eb0 - ea0, n);
    return s * s / sqn(n);
}

// A distance type = one primitive over a subset of the stencil's 4 points.
// prim: 0 PP, 1 PL, 2 plane, 3 LL; i0..i3 = stencil point feeding argument k.
struct Sub {
    int prim, i0, i1, i2, i3;
};
__host__ __device__ inline Sub sub_point_edge(int t)
{
    return t == PE_E0 ? Sub { 0, 0, 1, 0, 0 } : (t == PE_E1 ? Sub { 0, 0, 2, 0, 0 } : Sub { 1, 0, 1, 2, 0 });
}
__host__ __device__ inline Sub sub_point_triangle(int t)
{
    switch (t) {
    case PT_T0: return { 0, 0, 1, 0, 0 };
    case PT_T1: return { 0, 0, 2, 0, 0 };
    case PT_T2: return { 0, 0, 3, 0, 0 };
    case PT_E0: return { 1, 0, 1, 2, 0 };
    case PT_E1: return { 1, 0, 2, 3, 0 };
    case PT_E2: return { 1, 0, 3, 1, 0 };
    default: return { 2, 0, 1, 2, 3 };
    }
}
__host__ __device__ inline Sub sub_edge_edge(int t)
{
    switch (t) {
    case EE_A0B0: return { 0, 0, 2, 0, 0 };
    case EE_A0B1: return { 0, 0, 3, 0, 0 };
    case EE_A1B0: return { 0, 1, 2, 0, 0 };
    case EE_A1B1: return { 0, 1, 3, 0, 0 };
    case EE_AB0: return { 1, 2, 0, 1, 0 };
    case EE_AB1: return { 1, 3, 0, 1, 0 };
    case EE_A0B: return { 1, 0, 2, 3, 0 };
    case EE_A1B: return { 1, 1, 2, 3, 0 };
    default: return { 3, 0, 1, 2, 3 };
    }
}
__host__ __device__ inline double sub_value(const Sub& s, const d3* x)
{
    switch (s.prim) {
    case 0: return pp_dist(x[s.i0], x[s.i1]);
    case 1: return pl_dist(x[s.i0], x[s.i1], x[s.i2]);
    case 2: return plane_dist(x[s.i0], x[s.i1], x[s.i2], x[s.i3]);
    default: return ll_dist(x[s.i0], x[s.i1], x[s.i2], x[s.i3]);
    }
}

// ---- barrier (barrier/barrier.cpp:11-43) ------------------------------------------
__host__ __device__ inline double barrier_f(double d, double dhat)
{
    if (d <= 0.0) return INFINITY;
    if (d >= dhat) return 0.0;
    const double t = d - dhat;
    return -t * t * log(d / dhat);
}
__host__ __device__ inline double barrier_df(double d, double dhat)
{
    if (d <= 0.0 || d >= dhat) return 0.0;
    return (dhat - d) * (2 * log(d / dhat) - dhat / d + 1);
}
__host__ __device__ inline double barrier_ddf(double d, double dhat)
{
    if (d <= 0.0 || d >= dhat) return 0.0;
    const double q = dhat / d;
    return (q + 2) * q - 2 * log(d / dhat) - 3;
}

// ---- mollifier scalars (edge_edge_mollifier.cpp:43-73,194-202) ---------------------------
__host__ __device__ inline double moll(double x, double eps)
{
    if (x < eps) {
        const double q = x / eps;
        return (-q + 2) * q;
    }
    return 1.0;
}
__host__ __device__ inline double moll_d(double x, double eps)
{
    if (x < eps) {
        const double ie = 1 / eps;
        return 2 * ie * fma(-ie, x, 1.0);
    }
    return 0.0;
}
__host__ __device__ inline double moll_dd(double x, double eps) { return x < eps ? -2 / (eps * eps) : 0.0; }
__host__ __device__ inline double moll_threshold(d3 a0, d3 a1, d3 b0, d3 b1) { return 1e-3 * sqn(a0 - a1) * sqn(b0 - b1); }

// ============================================================================
// Derivatives.  LocalDeriv holds the stencil gradient (12) and the stencil
// Hessian as 4x4 blocks of 3x3 (row-major inside a block): H[(bi*4+bj)*9 + 3r+c].
struct LocalDeriv {
    double g[12];
    double H[144];
};

struct M3 {
    double m[9];
};
__device__ inline M3 outer3(d3 u, d3 v)
{
    return { { u.x * v.x, u.x * v.y, u.x * v.z, u.y * v.x, u.y * v.y, u.y * v.z, u.z * v.x, u.z * v.y, u.z * v.z } };
}
__device__ inline M3 skew3(d3 w) { return { { 0, -w.z, w.y, w.z, 0, -w.x, -w.y, w.x, 0 } }; }
__device__ inline M3 diag3(double s) { return { { s, 0, 0, 0, s, 0, 0, 0, s } }; }
__device__ inline M3 tr3(const M3& a) { return { { a.m[0], a.m[3], a.m[6], a.m[1], a.m[4], a.m[7], a.m[2], a.m[5], a.m[8] } }; }
// r = ca*a + cb*b + cc*c
__device__ inline M3 lin3(double ca, const M3& a, double cb, const M3& b, double cc, const M3& c)
{
    M3 r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.m[i] = ca * a.m[i] + cb * b.m[i] + cc * c.m[i];
    return r;
}

// scatter a difference-coordinate gradient / Hessian block onto stencil points:
// difference k = x[plus_k] - x[minus_k]
__device__ inline void scatter_g(LocalDeriv& L, int plus, int minus, d3 g)
{
    L.g[3 * plus] += g.x, L.g[3 * plus + 1] += g.y, L.g[3 * plus + 2] += g.z;
    L.g[3 * minus] -= g.x, L.g[3 * minus + 1] -= g.y, L.g[3 * minus + 2] -= g.z;
}
__device__ inline void scatter_H(LocalDeriv& L, int pk, int mk, int pl, int ml, const M3& B)
{
#pragma unroll
    for (int i = 0; i < 9; i++) {
        L.H[(pk * 4 + pl) * 9 + i] += B.m[i];
        L.H[(mk * 4 + ml) * 9 + i] += B.m[i];
        L.H[(pk * 4 + ml) * 9 + i] -= B.m[i];
        L.H[(mk * 4 + pl) * 9 + i] -= B.m[i];
    }
}
__device__ inline void zero_local(LocalDeriv& L)
{
#pragma unroll 1
    for (int i = 0; i < 12; i++) L.g[i] = 0;
#pragma unroll 1
    for (int i = 0; i < 144; i++) L.H[i] = 0;
}

// d2 = |x[i1] - x[i0]|^2
__device__ inline double pp_deriv(const d3* x, int i0, int i1, LocalDeriv& L)
{
    const d3 r = x[i1] - x[i0]; // difference: plus i1, minus i0
    scatter_g(L, i1, i0, 2.0 * r);
    scatter_H(L, i1, i0, i1, i0, diag3(2.0));
    return sqn(r);
}

// |a x b|^2 derivatives wrt (a, b): gradient (ga, gb) and Hessian blocks
__device__ inline void cross_sq_deriv(d3 a, d3 b, d3 c /* = a x b */, d3& ga, d3& gb, M3& Haa, M3& Hab, M3& Hbb)
{
    ga = 2.0 * cross(b, c);
    gb = 2.0 * cross(c, a);
    const double aa = sqn(a), bb = sqn(b), ab = dot(a, b);
    Haa = lin3(2.0 * bb, diag3(1.0), -2.0, outer3(b, b), 0.0, diag3(0.0));
    Hbb = lin3(2.0 * aa, diag3(1.0), -2.0, outer3(a, a), 0.0, diag3(0.0));
    Hab = lin3(4.0, outer3(a, b), -2.0, outer3(b, a), -2.0 * ab, diag3(1.0));
}

// d2 = |a x b|^2 / |b - a|^2 with a = e0 - p, b = e1 - p
__device__ inline double pl_deriv(const d3* x, int ip, int ie0, int ie1, LocalDeriv& L)
{
    const d3 a = x[ie0] - x[ip], b = x[ie1] - x[ip], c = cross(a, b), e = x[ie1] - x[ie0];
    const double N = sqn(c), Lq = sqn(e);
    d3 gN[2];
    M3 Haa, Hab, Hbb;
    cross_sq_deriv(a, b, c, gN[0], gN[1], Haa, Hab, Hbb);
    const d3 gL[2] = { -2.0 * e, 2.0 * e };
    const double iL = 1.0 / Lq, NL2 = N * iL * iL, NL3 = 2.0 * N * iL * iL * iL;
    const int plus[2] = { ie0, ie1 };
    const M3* HN[2][2] = { { &Haa, &Hab }, { nullptr, &Hbb } };
    const M3 Hba = tr3(Hab);
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const d3 gk = { iL * gN[k].x - NL2 * gL[k].x, iL * gN[k].y - NL2 * gL[k].y, iL * gN[k].z - NL2 * gL[k].z };
        scatter_g(L, plus[k], ip, gk);
#pragma unroll
        for (int l = 0; l < 2; l++) {
            const M3& hn = (k == 1 && l == 0) ? Hba : *HN[k][l];
            const double hl = (k == l) ? 2.0 : -2.0;
            M3 B = lin3(iL, hn, -iL * iL, outer3(gN[k], gL[l]), -iL * iL, outer3(gL[k], gN[l]));
            B = lin3(1.0, B, -NL2 * hl, diag3(1.0), NL3, outer3(gL[k], gL[l]));
            scatter_H(L, plus[k], ip, plus[l], ip, B);
        }
    }
    return N / Lq;
}

// T(q,u,v) = (q . (u x v))^2 / |u x v|^2; differences given as (plus, minus) index pairs
__device__ inline void triple_deriv(d3 q, d3 u, d3 v, const int plus[3], const int minus[3], LocalDeriv& L)
{
    const d3 n = cross(u, v);
    const double s = dot(q, n), M = sqn(n);
    const d3 gs[3] = { n, cross(v, q), cross(q, u) };
    d3 gMu, gMv;
    M3 Huu, Huv, Hvv;
    cross_sq_deriv(u, v, n, gMu, gMv, Huu, Huv, Hvv);
    const d3 gM[3] = { { 0, 0, 0 }, gMu, gMv };
    const double iM = 1.0 / M, c1 = 2.0 * s * iM, c2 = s * s * iM * iM, c3 = 2.0 * s * iM * iM, c4 = 2.0 * s * s * iM * iM * iM;
    const M3 Z = diag3(0.0);
    const M3 Hvu = tr3(Huv);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const d3 gk = { c1 * gs[k].x - c2 * gM[k].x, c1 * gs[k].y - c2 * gM[k].y, c1 * gs[k].z - c2 * gM[k].z };
        scatter_g(L, plus[k], minus[k], gk);
#pragma unroll
        for (int l = 0; l < 3; l++) {
            // Hessian of s (trilinear): [q][u] = -[v]x, [q][v] = [u]x, [u][v] = -[q]x, antisymmetric pattern
            M3 Hs = Z;
            if (k == 0 && l == 1) Hs = lin3(-1.0, skew3(v), 0, Z, 0, Z);
            if (k == 0 && l == 2) Hs = skew3(u);
            if (k == 1 && l == 0) Hs = skew3(v);
            if (k == 1 && l == 2) Hs = lin3(-1.0, skew3(q), 0, Z, 0, Z);
            if (k == 2 && l == 0) Hs = lin3(-1.0, skew3(u), 0, Z, 0, Z);
            if (k == 2 && l == 1) Hs = skew3(q);
            const M3& HMkl = (k == 1 && l == 1) ? Huu : (k == 2 && l == 2) ? Hvv : (k == 1 && l == 2) ? Huv : (k == 2 && l == 1) ? Hvu : Z;
            M3 B = lin3(2.0 * iM, outer3(gs[k], gs[l]), c1, Hs, -c2, HMkl);
            B = lin3(1.0, B, -c3, outer3(gs[k], gM[l]), -c3, outer3(gM[k], gs[l]));
            B = lin3(1.0, B, c4, outer3(gM[k], gM[l]), 0, Z);
            scatter_H(L, plus[k], minus[k], plus[l], minus[l], B);
        }
    }
}

__device__ inline double plane_deriv(const d3* x, int ip, int i0, int i1, int i2, LocalDeriv& L)
{
    const int plus[3] = { ip, i1, i2 }, minus[3] = { i0, i0, i0 };
    triple_deriv(x[ip] - x[i0], x[i1] - x[i0], x[i2] - x[i0], plus, minus, L);
    return plane_dist(x[ip], x[i0], x[i1], x[i2]);
}
__device__ inline double ll_deriv(const d3* x, int a0, int a1, int b0, int b1, LocalDeriv& L)
{
    const int plus[3] = { b0, a1, b1 }, minus[3] = { a0, a0, b0 };
    triple_deriv(x[b0] - x[a0], x[a1] - x[a0], x[b1] - x[b0], plus, minus, L);
    return ll_dist(x[a0], x[a1], x[b0], x[b1]);
}
// value + derivatives of a sub-primitive accumulated into L (L must be zeroed by the caller)
__device__ inline double sub_deriv(const Sub& s, const d3* x, LocalDeriv& L)
{
    switch (s.prim) {
    case 0: return pp_deriv(x, s.i0, s.i1, L);
    case 1: return pl_deriv(x, s.i0, s.i1, s.i2, L);
    case 2: return plane_deriv(x, s.i0, s.i1, s.i2, s.i3, L);
    default: return ll_deriv(x, s.i0, s.i1, s.i2, s.i3, L);
    }
}
// ---- gradient-only forms (registers only) ---------------------------------------
// Same closed forms as the *_deriv functions above without the Hessian blocks.  gp[k] receives the
// gradient w.r.t. the primitive's k-th ARGUMENT point (y[0..np)); returns the squared distance.
__device__ __forceinline__ void cross_sq_grad(d3 a, d3 b, d3 c /* = a x b */, d3& ga, d3& gb)
{
    ga = 2.0 * cross(b, c);
    gb = 2.0 * cross(c, a);
}
__device__ __forceinline__ void triple_grad(d3 q, d3 u, d3 v, d3& gq, d3& gu, d3& gv)
{
    const d3 n = cross(u, v);
    const double s = dot(q, n), M = sqn(n);
    d3 gMu, gMv;
    cross_sq_grad(u, v, n, gMu, gMv);
    const double iM = 1.0 / M, c1 = 2.0 * s * iM, c2 = s * s * iM * iM;
    const d3 gsu = cross(v, q), gsv = cross(q, u);
    gq = c1 * n;
    gu = { c1 * gsu.x - c2 * gMu.x, c1 * gsu.y - c2 * gMu.y, c1 * gsu.z - c2 * gMu.z };
    gv = { c1 * gsv.x - c2 * gMv.x, c1 * gsv.y - c2 * gMv.y, c1 * gsv.z - c2 * gMv.z };
}
// prim: 0 PP (y0,y1), 1 PL (p, e0, e1), 2 plane (p, t0, t1, t2), 3 LL (a0, a1, b0, b1)
__device__ __forceinline__ void prim_grad(int prim, const d3* y, d3* gp)
{
    const d3 zero = { 0, 0, 0 };
    gp[0] = gp[1] = gp[2] = gp[3] = zero;
    if (prim == 0) {
        const d3 r = y[1] - y[0];
        gp[1] = 2.0 * r;
        gp[0] = zero - gp[1];
    } else if (prim == 1) {
        const d3 a = y[1] - y[0], b = y[2] - y[0], c = cross(a, b), e = y[2] - y[1];
        const double N = sqn(c), Lq = sqn(e);
        d3 ga, gb;
        cross_sq_grad(a, b, c, ga, gb);
        const double iL = 1.0 / Lq, NL2 = N * iL * iL;
        const d3 gL0 = -2.0 * e, gL1 = 2.0 * e;
        gp[1] = { iL * ga.x - NL2 * gL0.x, iL * ga.y - NL2 * gL0.y, iL * ga.z - NL2 * gL0.z };
        gp[2] = { iL * gb.x - NL2 * gL1.x, iL * gb.y - NL2 * gL1.y, iL * gb.z - NL2 * gL1.z };
        gp[0] = (zero - gp[1]) - gp[2];
    } else if (prim == 2) {
        d3 gq, gu, gv;
        triple_grad(y[0] - y[1], y[2] - y[1], y[3] - y[1], gq, gu, gv);
        gp[0] = gq, gp[2] = gu, gp[3] = gv;
        gp[1] = ((zero - gq) - gu) - gv;
    } else {
        d3 gq, gu, gv; // q = b0 - a0, u = a1 - a0, v = b1 - b0
        triple_grad(y[2] - y[0], y[1] - y[0], y[3] - y[2], gq, gu, gv);
        gp[0] = (zero - gq) - gu;
        gp[1] = gu;
        gp[2] = gq - gv;
        gp[3] = gv;
    }
}

// s = |(ea1-ea0) x (eb1-eb0)|^2 on stencil points 0..3
__device__ inline double cross_sqnorm_deriv(const d3* x, LocalDeriv& L)
{
    const d3 u = x[1] - x[0], v = x[3] - x[2], n = cross(u, v);
    d3 gu, gv;
    M3 Huu, Huv, Hvv;
    cross_sq_deriv(u, v, n, gu, gv, Huu, Huv, Hvv);
    scatter_g(L, 1, 0, gu);
    scatter_g(L, 3, 2, gv);
    scatter_H(L, 1, 0, 1, 0, Huu);
    scatter_H(L, 3, 2, 3, 2, Hvv);
    scatter_H(L, 1, 0, 3, 2, Huv);
    scatter_H(L, 3, 2, 1, 0, tr3(Huv));
    return sqn(n);
}

} // namespace ipcb
