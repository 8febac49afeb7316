// oracle/hyperdual.hpp — TEST INFRASTRUCTURE ONLY.
//
// Second-order forward-mode automatic differentiation in N variables.  Used by
// oracle/selftest.cpp to check the oracle's closed-form gradients / Hessians
// against the reference's VALUE formulas (distance/point_line.cpp:5-23,
// point_plane.cpp:10-26, line_line.cpp:7-16, edge_edge_mollifier.cpp:7-14)
// without trusting any hand derivation.
#pragma once
#include <array>
#include <cmath>

namespace oracle {

template <int N> struct HD {
    double v;
    std::array<double, N> g;
    std::array<double, N * N> h; // symmetric, row-major
    HD() : v(0)
    {
        g.fill(0);
        h.fill(0);
    }
    HD(double c) : v(c)
    {
        g.fill(0);
        h.fill(0);
    }
    static HD var(double x, int i)
    {
        HD r(x);
        r.g[i] = 1;
        return r;
    }
};
template <int N> HD<N> operator+(const HD<N>& a, const HD<N>& b)
{
    HD<N> r;
    r.v = a.v + b.v;
    for (int i = 0; i < N; i++) r.g[i] = a.g[i] + b.g[i];
    for (int i = 0; i < N * N; i++) r.h[i] = a.h[i] + b.h[i];
    return r;
}
template <int N> HD<N> operator-(const HD<N>& a, const HD<N>& b)
{
    HD<N> r;
    r.v = a.v - b.v;
    for (int i = 0; i < N; i++) r.g[i] = a.g[i] - b.g[i];
    for (int i = 0; i < N * N; i++) r.h[i] = a.h[i] - b.h[i];
    return r;
}
template <int N> HD<N> operator*(const HD<N>& a, const HD<N>& b)
{
    HD<N> r;
    r.v = a.v * b.v;
    for (int i = 0; i < N; i++) r.g[i] = a.g[i] * b.v + a.v * b.g[i];
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) r.h[i * N + j] = a.h[i * N + j] * b.v + a.g[i] * b.g[j] + a.g[j] * b.g[i] + a.v * b.h[i * N + j];
    return r;
}
// unary chain rule: f(a) with f', f''
template <int N> HD<N> chain(const HD<N>& a, double f, double df, double ddf)
{
    HD<N> r;
    r.v = f;
    for (int i = 0; i < N; i++) r.g[i] = df * a.g[i];
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) r.h[i * N + j] = df * a.h[i * N + j] + ddf * a.g[i] * a.g[j];
    return r;
}
template <int N> HD<N> inv(const HD<N>& a) { return chain(a, 1 / a.v, -1 / (a.v * a.v), 2 / (a.v * a.v * a.v)); }
template <int N> HD<N> operator/(const HD<N>& a, const HD<N>& b) { return a * inv(b); }
template <int N> HD<N> sqrt(const HD<N>& a)
{
    const double s = std::sqrt(a.v);
    return chain(a, s, 0.5 / s, -0.25 / (s * a.v));
}
template <int N> HD<N> log(const HD<N>& a) { return chain(a, std::log(a.v), 1 / a.v, -1 / (a.v * a.v)); }

template <int N> struct HV3 {
    HD<N> x, y, z;
};
template <int N> HV3<N> operator-(const HV3<N>& a, const HV3<N>& b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
template <int N> HD<N> dot(const HV3<N>& a, const HV3<N>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <int N> HV3<N> cross(const HV3<N>& a, const HV3<N>& b)
{
    return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
}

} // namespace oracle
