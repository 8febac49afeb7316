// hessian_fast.cuh — PSD-projected local barrier Hessians in an analytic subspace.
//
// A squared distance is phi(x) = min_theta |r(x,theta)|^2 with r = sum_i c_i x_i the vector between
// the closest points (c = barycentric coefficients, summing to zero) and theta the p <= 2 closest-point
// parameters.  With tau_j = dr/dtheta_j (edge directions) and e_j = d tau_j / dx (a +-1 pattern over
// the points), implicit differentiation of the minimiser gives
//
//     grad phi = 2 c (x) r ,     Hess phi = 2 (c c^T) (x) I3  -  2 A G^-1 A^T ,
//     a_j = c (x) tau_j + e_j (x) r ,   G_jk = tau_j . tau_k ,
//
// so the local matrix  w [ f'' grad grad^T + f' Hess ]  (normal_potential.cpp:196-204) lives in the
// span of the 3 + p ORTHONORMAL vectors  c^ (x) e_x, c^ (x) e_y, c^ (x) e_z, eps_1 (x) r^, eps_2 (x) r^
// (eps_j: the e_j made orthonormal and orthogonal to c^).  Projecting the (3+p) x (3+p) coordinate
// matrix to the PSD cone projects the full 6/9/12-dimensional matrix exactly (project_to_psd,
// utils/eigen_ext.tpp:56-108), and a 3x3 / 4x4 / 5x5 cyclic Jacobi fits in registers.
// Edge-edge collisions with an ACTIVE mollifier have extra terms and take the general path.
#pragma once
#include "geom.cuh"

namespace ipcb {

// Jacobi rotation parameters for the pivot (app, apq, aqq): the smaller rotation angle phi with tan(2 phi) = h / d,
// d = aqq - app, h = 2 apq.  Half-angle form with two reciprocal square roots and neither a division nor a square
// root (FP64 div / sqrt are ~20-instruction sequences each, rsqrt about half of that):
//   cos(2 phi) = |d| / sqrt(d^2 + h^2),  c = cos(phi) = sqrt((1 + cos 2phi) / 2),  s = sin(2 phi) / (2 c),  t = s / c.
__device__ __forceinline__ void jacobi_params(double app, double apq, double aqq, double& t, double& c, double& s)
{
    const double d = aqq - app, h = 2.0 * apq;
    const double rq = rsqrt(fma(d, d, h * h));
    const double c2 = fma(0.5 * fabs(d), rq, 0.5); // cos^2(phi) in [1/2, 1]
    const double rc = rsqrt(c2);                   // 1 / cos(phi)
    c = c2 * rc;
    s = copysign(0.5 * h * rq, d * h) * rc;        // sign(d) * h carries the sign of tan(2 phi)
    t = s * rc;
}
// apply the rotation in the (p, q) plane to A (upper triangle) and accumulate it into V
template <int N> __device__ __forceinline__ void jacobi_apply(double (&A)[N][N], double (&V)[N][N], int p, int q, double t, double c, double s)
{
    const double apq = A[p][q];
    A[p][p] = fma(-t, apq, A[p][p]);
    A[q][q] = fma(t, apq, A[q][q]);
    A[p][q] = 0.0;
#pragma unroll
    for (int k = 0; k < N; k++) {
        if (k != p && k != q) {
            double& akp = k < p ? A[k][p] : A[p][k];
            double& akq = k < q ? A[k][q] : A[q][k];
            const double x = akp, y = akq;
            akp = fma(c, x, -s * y);
            akq = fma(s, x, c * y);
        }
    }
#pragma unroll
    for (int k = 0; k < N; k++) {
        const double x = V[k][p], y = V[k][q];
        V[k][p] = fma(c, x, -s * y);
        V[k][q] = fma(s, x, c * y);
    }
}

// Jacobi eigen-solver, fully unrolled in registers: A symmetric (upper triangle used), V eigenvectors in columns.
// Parallel (round-robin) ordering: the pivots of one round are disjoint, so their parameters (the serial part:
// a square root, a division and a reciprocal square root each) are independent instruction chains that the
// scheduler overlaps; rotations on disjoint planes commute, so they are applied one after the other.
template <int N> __device__ __forceinline__ void jacobi_reg(double (&A)[N][N], double (&V)[N][N])
{
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) V[i][j] = i == j ? 1.0 : 0.0;
    double tot = 0;
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = i; j < N; j++) tot = fma(A[i][j], A[i][j], tot);
    const double stop = tot * 1e-33;
    // an entry whose square is below stop / #pairs cannot keep the sweep loop alive: rotating it away is wasted work
    const double skip = stop / double(N * (N - 1) / 2);
    constexpr int M = (N % 2) ? N : N - 1; // rounds per sweep
#pragma unroll 1
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0;
#pragma unroll
        for (int p = 0; p < N; p++)
#pragma unroll
            for (int q = p + 1; q < N; q++) off = fma(A[p][q], A[p][q], off);
        if (off <= stop) break;
#pragma unroll
        for (int r = 0; r < M; r++) {
            // pairs of round r: i + j = r (mod M) among 0..M-1; for even N the index N-1 meets the one left over
            double t[N / 2], c[N / 2], s[N / 2];
            bool on[N / 2];
            int pp[N / 2], qq[N / 2];
            int np = 0;
#pragma unroll
            for (int i = 0; i < M; i++) {
                const int j = ((r - i) % M + M) % M;
                if (i < j) pp[np] = i, qq[np] = j, np++;
                else if (i == j && N % 2 == 0) pp[np] = i, qq[np] = N - 1, np++;
            }
#pragma unroll
            for (int k = 0; k < N / 2; k++) {
                const double apq = A[pp[k]][qq[k]];
                on[k] = apq * apq > skip;
                if (on[k]) jacobi_params(A[pp[k]][pp[k]], apq, A[qq[k]][qq[k]], t[k], c[k], s[k]);
            }
#pragma unroll
            for (int k = 0; k < N / 2; k++)
                if (on[k]) jacobi_apply<N>(A, V, pp[k], qq[k], t[k], c[k], s[k]);
        }
    }
}

struct FastGeom {
    d3 r;         // closest-point vector
    double cv[4]; // barycentric coefficients over the primitive's points
    d3 tau[2];    // dr/dtheta_j
    double ev[2][4];
    double g11, g12, g22; // G = tau_j . tau_k
};

// closest-point data of a sub-primitive over points y[0..np)
__device__ __forceinline__ int fast_geometry(int prim, const d3* y, FastGeom& g)
{
#pragma unroll
    for (int k = 0; k < 4; k++) g.cv[k] = 0, g.ev[0][k] = 0, g.ev[1][k] = 0;
    g.g11 = g.g22 = 1.0, g.g12 = 0.0;
    g.tau[0] = g.tau[1] = d3 { 0, 0, 0 };
    if (prim == 0) { // point - point
        g.r = y[0] - y[1];
        g.cv[0] = 1, g.cv[1] = -1;
        return 0;
    }
    if (prim == 1) { // point y0 - line (y1, y2)
        const d3 tl = y[2] - y[1], q = y[0] - y[1];
        const double ll = dot(tl, tl), lam = dot(q, tl) / ll;
        g.r = q - lam * tl;
        g.cv[0] = 1, g.cv[1] = lam - 1, g.cv[2] = -lam;
        g.tau[0] = -1.0 * tl;
        g.ev[0][1] = 1, g.ev[0][2] = -1;
        g.g11 = ll;
        return 1;
    }
    if (prim == 2) { // point y0 - plane (y1, y2, y3)
        const d3 ta = y[2] - y[1], tb = y[3] - y[1], q = y[0] - y[1];
        const double a = dot(ta, ta), b = dot(ta, tb), c = dot(tb, tb), b1 = dot(q, ta), b2 = dot(q, tb);
        const double det = a * c - b * b;
        const double la = (c * b1 - b * b2) / det, lb = (a * b2 - b * b1) / det;
        g.r = q - la * ta - lb * tb;
        g.cv[0] = 1, g.cv[1] = la + lb - 1, g.cv[2] = -la, g.cv[3] = -lb;
        g.tau[0] = -1.0 * ta, g.tau[1] = -1.0 * tb;
        g.ev[0][1] = 1, g.ev[0][2] = -1;
        g.ev[1][1] = 1, g.ev[1][3] = -1;
        g.g11 = a, g.g12 = b, g.g22 = c;
        return 2;
    }
    // line (y0, y1) - line (y2, y3)
    const d3 u = y[1] - y[0], v = y[3] - y[2], w0 = y[0] - y[2];
    const double a = dot(u, u), b = -dot(u, v), c = dot(v, v), b1 = -dot(w0, u), b2 = dot(w0, v);
    const double det = a * c - b * b;
    const double s = (c * b1 - b * b2) / det, t = (a * b2 - b * b1) / det;
    g.r = w0 + s * u - t * v;
    g.cv[0] = 1 - s, g.cv[1] = s, g.cv[2] = -(1 - t), g.cv[3] = -t;
    g.tau[0] = u, g.tau[1] = -1.0 * v;
    g.ev[0][0] = -1, g.ev[0][1] = 1;
    g.ev[1][2] = 1, g.ev[1][3] = -1;
    g.g11 = a, g.g12 = b, g.g22 = c;
    return 2;
}

// Projected coordinate matrix and basis of a primitive with P parameters.
template <int P> struct FastProj {
    double Mp[3 + P][3 + P]; // PSD-projected coordinate matrix
    double eps[2][4];        // orthonormal eps_j over the primitive's points
    double ch[4];            // c / |c|
    double rh[3];            // r / |r|
};

template <int P> __device__ __forceinline__ void fast_project(const FastGeom& g, double wf1, double wf2, int mode, FastProj<P>& pr)
{
    constexpr int N = 3 + P;
    const d3 r = g.r;
    const double d = sqrt(dot(r, r));
    pr.rh[0] = r.x / d, pr.rh[1] = r.y / d, pr.rh[2] = r.z / d;
    double cn2 = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) cn2 = fma(g.cv[k], g.cv[k], cn2);
    const double cn = sqrt(cn2);
#pragma unroll
    for (int k = 0; k < 4; k++) pr.ch[k] = g.cv[k] / cn;
    // orthonormal eps_j (orthogonal to ch) and the coordinates of a_j
    double tvec[2][3], bvec[2][2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
#pragma unroll
        for (int k = 0; k < 4; k++) pr.eps[j][k] = 0;
        tvec[j][0] = tvec[j][1] = tvec[j][2] = 0, bvec[j][0] = bvec[j][1] = 0;
    }
    if (P >= 1) {
        double al = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) al = fma(g.ev[0][k], pr.ch[k], al);
        double nn = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            pr.eps[0][k] = g.ev[0][k] - al * pr.ch[k];
            nn = fma(pr.eps[0][k], pr.eps[0][k], nn);
        }
        const double n1 = sqrt(nn);
#pragma unroll
        for (int k = 0; k < 4; k++) pr.eps[0][k] /= n1;
        tvec[0][0] = fma(cn, g.tau[0].x, al * r.x), tvec[0][1] = fma(cn, g.tau[0].y, al * r.y), tvec[0][2] = fma(cn, g.tau[0].z, al * r.z);
        bvec[0][0] = d * n1;
    }
    if (P >= 2) {
        double al = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) al = fma(g.ev[1][k], pr.ch[k], al);
        double w2[4], beta = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            w2[k] = g.ev[1][k] - al * pr.ch[k];
            beta = fma(w2[k], pr.eps[0][k], beta);
        }
        double nn = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            w2[k] -= beta * pr.eps[0][k];
            nn = fma(w2[k], w2[k], nn);
        }
        const double gam = sqrt(nn);
#pragma unroll
        for (int k = 0; k < 4; k++) pr.eps[1][k] = w2[k] / gam;
        tvec[1][0] = fma(cn, g.tau[1].x, al * r.x), tvec[1][1] = fma(cn, g.tau[1].y, al * r.y), tvec[1][2] = fma(cn, g.tau[1].z, al * r.z);
        bvec[1][0] = d * beta, bvec[1][1] = d * gam;
    }
    // G^-1
    double gi[2][2] = { { 0, 0 }, { 0, 0 } };
    if (P == 1) gi[0][0] = 1.0 / g.g11;
    if (P == 2) {
        const double det = g.g11 * g.g22 - g.g12 * g.g12;
        gi[0][0] = g.g22 / det, gi[1][1] = g.g11 / det, gi[0][1] = gi[1][0] = -g.g12 / det;
    }
    // coordinate matrix M = 4 w f'' cn^2 [r;0][r;0]^T + 2 w f' cn^2 diag(I3, 0) - 2 w f' A~ G^-1 A~^T
    double M[N][N], V[N][N];
    const double rr[3] = { r.x, r.y, r.z };
    double at[2][N]; // a~_j
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int i = 0; i < N; i++) at[j][i] = i < 3 ? tvec[j][i] : bvec[j][i - 3];
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = i; j < N; j++) {
            double v = 0;
            if (i < 3 && j < 3) v = 4.0 * wf2 * cn2 * rr[i] * rr[j] + (i == j ? 2.0 * wf1 * cn2 : 0.0);
            double k2 = 0;
#pragma unroll
            for (int a = 0; a < P; a++)
#pragma unroll
                for (int b = 0; b < P; b++) k2 = fma(gi[a][b] * at[a][i], at[b][j], k2);
            M[i][j] = v - 2.0 * wf1 * k2;
        }
    jacobi_reg<N>(M, V);
    double lam[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        lam[i] = M[i][i];
        if (lam[i] < 0.0) lam[i] = mode == IPCB_PSD_CLAMP ? 0.0 : -lam[i];
    }
    // M+ = V diag(lam) V^T
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = i; j < N; j++) {
            double acc = 0;
#pragma unroll
            for (int k = 0; k < N; k++) acc = fma(V[i][k] * lam[k], V[j][k], acc);
            pr.Mp[i][j] = pr.Mp[j][i] = acc;
        }
}

// 3x3 block (rows of point a, columns of point b; row-major) of the projected local matrix:
//   ch_a ch_b Mp_TT + ch_a (r^ (x) u_b) ... (see header)
template <int P> __device__ __forceinline__ void fast_block(const FastProj<P>& pr, int a, int b, double* blk)
{
    double ua[3] = { 0, 0, 0 }, ub[3] = { 0, 0, 0 }, sab = 0;
#pragma unroll
    for (int k = 0; k < P; k++) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            ua[c] = fma(pr.eps[k][a], pr.Mp[3 + k][c], ua[c]);
            ub[c] = fma(pr.eps[k][b], pr.Mp[c][3 + k], ub[c]);
        }
#pragma unroll
        for (int l = 0; l < P; l++) sab = fma(pr.eps[k][a] * pr.eps[l][b], pr.Mp[3 + k][3 + l], sab);
    }
    const double cab = pr.ch[a] * pr.ch[b];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++)
            blk[3 * r + c] = cab * pr.Mp[r][c] + pr.ch[a] * ub[r] * pr.rh[c] + pr.ch[b] * pr.rh[r] * ua[c] + sab * pr.rh[r] * pr.rh[c];
}

} // namespace ipcb
