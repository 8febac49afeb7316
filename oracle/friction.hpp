// oracle/friction.hpp — TEST INFRASTRUCTURE ONLY (part of the CPU restatement, see oracle.cpp).
//
// Friction: the lagged tangential collision set and the smooth dissipative potential, restated from
// the reference (src/ipc/):
//   tangent/tangent_basis.cpp:17-48,72-92,120-134,156-171        tangent bases (3D)
//   tangent/closest_point.cpp:11-18,65-88,123-137                closest-point coordinates
//   tangent/relative_velocity.cpp                                 relative-velocity coefficients (Gamma)
//   friction/smooth_friction_mollifier.cpp, friction/smooth_mu.cpp
//   barrier/barrier_force_magnitude.cpp:7-15, potentials/barrier_potential.cpp:33-44   normal force magnitude
//   collisions/tangential/tangential_collisions.cpp:62-171       TangentialCollisions::build (isotropic coefficients)
//   potentials/tangential_potential.cpp:162-325                  per-collision energy / gradient / Hessian
// Out of scope (like the reference's force Jacobians and anisotropic "matchstick" coefficients): anything that
// differentiates the lagged quantities.
#pragma once
#include "eig.hpp"
#include "geom.hpp"

namespace oracle {

// Eigen's normalized(): v / sqrt(squaredNorm) when the squared norm is positive, v itself otherwise
inline V3 normalized(V3 v)
{
    const double n2 = sqnorm(v);
    if (!(n2 > 0)) return v;
    const double n = std::sqrt(n2);
    return { v.x / n, v.y / n, v.z / n };
}

// ---- tangent bases: P[0], P[1] = the two columns (tangent_basis.cpp)
inline void point_point_tangent_basis(V3 p0, V3 p1, V3 P[2])
{
    const V3 d = p1 - p0;
    const V3 cx = cross(V3 { 1, 0, 0 }, d), cy = cross(V3 { 0, 1, 0 }, d);
    if (sqnorm(cx) > sqnorm(cy)) {
        P[0] = normalized(cx), P[1] = normalized(cross(d, cx));
    } else {
        P[0] = normalized(cy), P[1] = normalized(cross(d, cy));
    }
}
inline void point_edge_tangent_basis(V3 p, V3 e0, V3 e1, V3 P[2])
{
    const V3 e = e1 - e0;
    P[0] = normalized(e), P[1] = normalized(cross(e, p - e0));
}
inline void edge_edge_tangent_basis(V3 ea0, V3 ea1, V3 eb0, V3 eb1, V3 P[2])
{
    const V3 ea = ea1 - ea0, normal = cross(ea, eb1 - eb0);
    P[0] = normalized(ea), P[1] = normalized(cross(normal, ea));
}
inline void point_triangle_tangent_basis(V3, V3 t0, V3 t1, V3 t2, V3 P[2])
{
    const V3 e0 = t1 - t0, normal = cross(e0, t2 - t0);
    P[0] = normalized(e0), P[1] = normalized(cross(normal, e0));
}

// ---- closest points (closest_point.cpp)
inline double point_edge_closest_point(V3 p, V3 e0, V3 e1)
{
    const V3 e = e1 - e0;
    return dot(p - e0, e) / sqnorm(e);
}
inline void edge_edge_closest_point(V3 ea0, V3 ea1, V3 eb0, V3 eb1, double beta[2])
{
    const V3 eb_to_ea = ea0 - eb0, ea = ea1 - ea0, eb = eb1 - eb0;
    const double a01 = -dot(eb, ea);
    ldlt2_solve(sqnorm(ea), a01, sqnorm(eb), -dot(eb_to_ea, ea), dot(eb_to_ea, eb), beta[0], beta[1]);
}
inline void point_triangle_closest_point(V3 p, V3 t0, V3 t1, V3 t2, double beta[2])
{
    const V3 b0 = t1 - t0, b1 = t2 - t0, q = p - t0;
    ldlt2_solve(dot(b0, b0), dot(b0, b1), dot(b1, b1), dot(b0, q), dot(b1, q), beta[0], beta[1]);
}

// ---- smooth friction mollifier (smooth_friction_mollifier.cpp) and smooth mu (smooth_mu.cpp)
inline double smooth_friction_f0(double y, double eps_v) { return std::abs(y) >= eps_v ? y : y * y * (1 - y / (3 * eps_v)) / eps_v + eps_v / 3; }
inline double smooth_friction_f1(double y, double eps_v)
{
    if (std::abs(y) >= eps_v) return 1;
    const double r = y / eps_v;
    return r * (2 - r);
}
inline double smooth_friction_f2(double y, double eps_v) { return std::abs(y) >= eps_v ? 0 : (2 - 2 * y / eps_v) / eps_v; }
inline double smooth_friction_f1_over_x(double y, double eps_v) { return std::abs(y) >= eps_v ? 1 / y : (2 - y / eps_v) / eps_v; }
inline double smooth_friction_f2_x_minus_f1_over_x3(double y, double eps_v)
{
    return std::abs(y) >= eps_v ? -1 / (y * y * y) : -1 / (y * eps_v * eps_v);
}
inline double smooth_mu(double y, double mu_s, double mu_k, double eps_v)
{
    if (mu_s == mu_k || std::abs(y) >= eps_v) return mu_k;
    const double z = std::abs(y) / eps_v;
    if (std::abs(y) < 0.5 * eps_v) return 2 * (mu_k - mu_s) * z * z + mu_s;
    return -2 * (mu_k - mu_s) * (z * (z - 2) + 1) + mu_k;
}
inline double smooth_mu_derivative(double y, double mu_s, double mu_k, double eps_v)
{
    if (mu_s == mu_k || std::abs(y) >= eps_v) return 0;
    const double z = std::abs(y) / eps_v;
    if (std::abs(y) < 0.5 * eps_v) return 4 * (mu_k - mu_s) * z / eps_v;
    return -4 * (mu_k - mu_s) * (z - 1) / eps_v;
}
inline double smooth_mu_f0(double y, double mu_s, double mu_k, double eps_v)
{
    if (mu_s == mu_k || std::abs(y) >= eps_v) return mu_k * smooth_friction_f0(y, eps_v);
    const double delta_mu = mu_k - mu_s, z = std::abs(y) / eps_v;
    if (std::abs(y) < 0.5 * eps_v)
        return y * z * (z * (z * (1 - 0.4 * z) * delta_mu - mu_s / 3.0) + mu_s) + (9.0 / 16.0) * eps_v * mu_k - (11.0 / 48.0) * eps_v * mu_s;
    return y * z * (z * (z * (0.4 * z - 2) * delta_mu + (3 * mu_k - (10.0 / 3.0) * mu_s)) + (2 * mu_s - mu_k)) + 0.6 * eps_v * mu_k
        - (4.0 / 15.0) * eps_v * mu_s;
}
inline double smooth_mu_f1(double y, double mu_s, double mu_k, double eps_v) { return smooth_mu(y, mu_s, mu_k, eps_v) * smooth_friction_f1(y, eps_v); }
inline double smooth_mu_f2(double y, double mu_s, double mu_k, double eps_v)
{
    return smooth_mu_derivative(y, mu_s, mu_k, eps_v) * smooth_friction_f1(y, eps_v) + smooth_mu(y, mu_s, mu_k, eps_v) * smooth_friction_f2(y, eps_v);
}
inline double smooth_mu_f1_over_x(double y, double mu_s, double mu_k, double eps_v)
{
    return smooth_mu(y, mu_s, mu_k, eps_v) * smooth_friction_f1_over_x(y, eps_v);
}
inline double smooth_mu_f2_x_minus_mu_f1_over_x3(double y, double mu_s, double mu_k, double eps_v)
{
    if (mu_s == mu_k || std::abs(y) >= eps_v) return mu_k * smooth_friction_f2_x_minus_f1_over_x3(y, eps_v);
    const double delta_mu = mu_k - mu_s, z = 1 / eps_v;
    if (std::abs(y) < 0.5 * eps_v) return z * z * (z * (8 - 6 * y * z) * delta_mu - mu_s / y);
    return z * z * (z * (6 * y * z - 16) * delta_mu + (9 * mu_k - 10 * mu_s) / y);
}

// ---- one lagged tangential collision (collisions/tangential/tangential_collision.hpp)
struct Tang {
    int32_t a, b;    // ids like the normal collision: VV (v0,v1); EV (edge,vertex); EE (ea,eb); FV (face,vertex)
    double weight;
    double normal_force;
    double mu_s, mu_k;
    double beta[2]; // closest point
    V3 P[2];        // tangent basis columns
};
// relative-velocity coefficients of the stencil points (relative_velocity.cpp): u_rel = sum_a gamma[a] v_a
inline int tangential_gamma(int kind, const double* beta, double* gamma)
{
    if (kind == 0) return gamma[0] = 1, gamma[1] = -1, 2;
    if (kind == 1) return gamma[0] = 1, gamma[1] = beta[0] - 1, gamma[2] = -beta[0], 3;
    if (kind == 2) return gamma[0] = 1 - beta[0], gamma[1] = beta[0], gamma[2] = beta[1] - 1, gamma[3] = -beta[1], 4;
    return gamma[0] = 1, gamma[1] = beta[0] + beta[1] - 1, gamma[2] = -beta[0], gamma[3] = -beta[1], 4;
}
// u = P^T Gamma v
inline void tangential_slip(int kind, const Tang& t, const V3* v, double u[2], double* gamma, int& n)
{
    n = tangential_gamma(kind, t.beta, gamma);
    V3 rel { 0, 0, 0 };
    for (int a = 0; a < n; a++) rel = rel + gamma[a] * v[a];
    u[0] = dot(t.P[0], rel), u[1] = dot(t.P[1], rel);
}
// tangential_potential.cpp:162-187
inline double friction_energy(int kind, const Tang& t, const V3* v, double eps_v)
{
    double u[2], gamma[4];
    int n;
    tangential_slip(kind, t, v, u, gamma, n);
    return t.weight * t.normal_force * smooth_mu_f0(std::sqrt(u[0] * u[0] + u[1] * u[1]), t.mu_s, t.mu_k, eps_v);
}
// :189-237 — g: 3 n doubles
inline void friction_gradient(int kind, const Tang& t, const V3* v, double eps_v, double* g)
{
    double u[2], gamma[4];
    int n;
    tangential_slip(kind, t, v, u, gamma, n);
    const double nu = std::sqrt(u[0] * u[0] + u[1] * u[1]);
    const double s = smooth_mu_f1_over_x(nu, t.mu_s, t.mu_k, eps_v) * (t.weight * t.normal_force);
    const V3 f = (s * u[0]) * t.P[0] + (s * u[1]) * t.P[1]; // P (s u)
    for (int a = 0; a < n; a++) g[3 * a] = gamma[a] * f.x, g[3 * a + 1] = gamma[a] * f.y, g[3 * a + 2] = gamma[a] * f.z;
}
// :239-325 — H: (3 n) x (3 n) col-major with leading dimension 12; every block (a, b) is gamma_a gamma_b P M P^T
inline void friction_hessian(int kind, const Tang& t, const V3* v, double eps_v, int psd_mode, double* H)
{
    double u[2], gamma[4];
    int n;
    tangential_slip(kind, t, v, u, gamma, n);
    const double nu = std::sqrt(u[0] * u[0] + u[1] * u[1]);
    const double f1ox = smooth_mu_f1_over_x(nu, t.mu_s, t.mu_k, eps_v);
    const double scale = t.weight * t.normal_force;
    double M[4] = { 0, 0, 0, 0 }; // 2 x 2 col-major
    if (nu > eps_v) { // is_dynamic: f1 = 1, f2 = 0: mu N f1/|u| (I - u u^T / |u|^2), PSD already
        if (!(psd_mode != 0 && scale <= 0)) {
            const double c = scale * f1ox / (nu * nu);
            const double p[2] = { -u[1], u[0] };
            M[0] = c * p[0] * p[0], M[1] = M[2] = c * p[0] * p[1], M[3] = c * p[1] * p[1];
        }
    } else if (nu == 0) {
        if (!(psd_mode != 0 && scale <= 0)) M[0] = M[3] = scale * f1ox;
    } else {
        const double f2 = smooth_mu_f2_x_minus_mu_f1_over_x3(nu, t.mu_s, t.mu_k, eps_v);
        M[0] = (f2 * u[0] * u[0] + f1ox) * scale, M[3] = (f2 * u[1] * u[1] + f1ox) * scale;
        M[1] = M[2] = (f2 * u[0] * u[1]) * scale;
        project_to_psd(2, M, 2, psd_mode);
    }
    // K = P M P^T
    const double Pm[2][3] = { { t.P[0].x, t.P[0].y, t.P[0].z }, { t.P[1].x, t.P[1].y, t.P[1].z } };
    double K[3][3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            K[r][c] = (Pm[0][r] * M[0] + Pm[1][r] * M[1]) * Pm[0][c] + (Pm[0][r] * M[2] + Pm[1][r] * M[3]) * Pm[1][c];
    for (int a = 0; a < n; a++)
        for (int b = 0; b < n; b++)
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) H[(3 * a + r) + 12 * (3 * b + c)] = (gamma[a] * gamma[b]) * K[r][c];
}

} // namespace oracle
