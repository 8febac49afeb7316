// exact_orient3d.hpp — sign of the orientation determinant of four points, EXACTLY, on the host.
//
// ipc::has_intersections (ipc.cpp:105-166) asks igl::predicates::orient3d (Shewchuk's adaptive exact predicate, a
// third-party dependency that is not in the reference tree) on which side of a triangle's plane the two end points of an
// edge lie.  The CUDA kernel answers with a floating-point determinant and Shewchuk's static error bound; the few
// candidates whose determinant is below the bound come here.  The determinant is expanded into its 24 triple products of
// INPUT coordinates (no rounded differences), every product is formed exactly (two-product with FMA: 4 doubles) and the 96
// terms are summed as a non-overlapping floating-point expansion; the sign of the sum is the sign of its largest component.
#pragma once
#include <cmath>
#include <vector>

namespace ipcb_exact {

inline void two_sum(double a, double b, double& s, double& e)
{
    s = a + b;
    const double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}
inline void two_prod(double a, double b, double& p, double& e)
{
    p = a * b;
    e = std::fma(a, b, -p);
}
// h = e + b as a non-overlapping expansion (Shewchuk, Grow-Expansion with zero elimination)
inline void grow(std::vector<double>& e, double b)
{
    std::vector<double> h;
    h.reserve(e.size() + 1);
    double q = b;
    for (double ei : e) {
        double s, r;
        two_sum(q, ei, s, r);
        if (r != 0.0) h.push_back(r);
        q = s;
    }
    if (q != 0.0) h.push_back(q);
    e.swap(h);
}
inline void add_triple(std::vector<double>& acc, double sign, double x, double y, double z)
{
    double p, e, a, b, c, d;
    two_prod(x, y, p, e);
    two_prod(p, z, a, b);
    two_prod(e, z, c, d);
    for (double t : { d, c, b, a })
        if (t != 0.0) grow(acc, sign * t);
}
// det3 of rows P, Q, R added to acc with `sign`
inline void add_det3(std::vector<double>& acc, double sign, const double* P, const double* Q, const double* R)
{
    add_triple(acc, sign, P[0], Q[1], R[2]), add_triple(acc, -sign, P[0], Q[2], R[1]);
    add_triple(acc, -sign, P[1], Q[0], R[2]), add_triple(acc, sign, P[1], Q[2], R[0]);
    add_triple(acc, sign, P[2], Q[0], R[1]), add_triple(acc, -sign, P[2], Q[1], R[0]);
}
// sign (-1, 0, +1) of det [a - d; b - d; c - d] = det3(a,b,c) - det3(d,b,c) + det3(d,a,c) - det3(d,a,b)
inline int orient3d_sign(const double* a, const double* b, const double* c, const double* d)
{
    std::vector<double> acc;
    add_det3(acc, 1.0, a, b, c);
    add_det3(acc, -1.0, d, b, c);
    add_det3(acc, 1.0, d, a, c);
    add_det3(acc, -1.0, d, a, b);
    if (acc.empty()) return 0;
    return acc.back() > 0 ? 1 : -1;
}

} // namespace ipcb_exact
