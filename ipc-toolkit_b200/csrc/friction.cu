// friction.cu — the lagged tangential collision set and the smooth friction potential.
//
// Replaces (reference src/ipc/): collisions/tangential/tangential_collisions.cpp:62-171 (TangentialCollisions::build),
// tangent/{tangent_basis,closest_point,relative_velocity}.cpp (the lagged geometry; the autogen Jacobians of those files
// serve the force Jacobians, which are outside this path), barrier/barrier_force_magnitude.cpp:7-15,
// friction/{smooth_friction_mollifier,smooth_mu}.cpp, potentials/tangential_potential.cpp:162-325 (per-collision energy /
// gradient / Hessian) and the assembly of potentials/potential.cpp:36-222.
//
// GPU formulation: one thread per normal collision builds its tangential record (struct of arrays: ids, weight, normal
// force, blended coefficients, closest point, the two basis columns); near-parallel edge-edge collisions are dropped
// with a scan (the reference skips them, order is kept).  The potential needs, per collision, only the 2-vector
// u = P^T sum_a gamma_a v_a: energy is a deterministic two-level reduction, the gradient scatters gamma_a P (s u) with
// FP64 atomics, and EVERY 3x3 vertex block of the local Hessian is gamma_a gamma_b K with ONE 3x3 matrix K = P M P^T —
// the inner 2x2 matrix M = scale [f1/|u| I + f2 u u^T] has the eigenvectors u and u-perp, so project_to_psd is analytic.
// The blocks go through the same record format and column assembly as the barrier Hessian (hessian_assembly.cuh).
#include "hessian_assembly.cuh"
#include <cub/device/device_scan.cuh>

namespace ipcb {

// ---- lagged geometry --------------------------------------------------------------------------------------------------
// Eigen's normalized(): v / sqrt(squaredNorm) when the squared norm is positive
__device__ inline d3 normalized3(d3 v)
{
    const double n2 = sqn(v);
    if (!(n2 > 0)) return v;
    const double n = sqrt(n2);
    return { v.x / n, v.y / n, v.z / n };
}
// tangent_basis.cpp:17-48
__device__ inline void pp_tangent_basis(d3 p0, d3 p1, d3* P)
{
    const d3 d = p1 - p0;
    const d3 cx = cross(mk3(1, 0, 0), d), cy = cross(mk3(0, 1, 0), d);
    if (sqn(cx) > sqn(cy)) P[0] = normalized3(cx), P[1] = normalized3(cross(d, cx));
    else P[0] = normalized3(cy), P[1] = normalized3(cross(d, cy));
}

struct TangOut {
    int2* ids;
    double *w, *N, *mus, *muk;
    double2* beta;
    double* P; // 6 per record
    unsigned char* keep;
};

// TangentialCollisions::build, one thread per normal collision of kind KIND (tangential_collisions.cpp:62-171)
template <int KIND>
__global__ void __launch_bounds__(256) k_tangential_build(CollView c, MeshView m, BarrierDev B, const double* __restrict__ mu_s,
                                                          const double* __restrict__ mu_k, TangOut o)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    int vid[4];
    d3 x[4];
    load_stencil(KIND, c.ids[i], m, vid, x);
    d3 P[2];
    double beta0 = 0, beta1 = 0, d2, ms, mk;
    bool keep = true;
    auto blend = [](double a, double b) { return (a + b) / 2; }; // default_blend_mu
    if (KIND == IPCB_VV) {
        pp_tangent_basis(x[0], x[1], P);
        d2 = pp_dist(x[0], x[1]);
        ms = blend(mu_s[vid[0]], mu_s[vid[1]]), mk = blend(mu_k[vid[0]], mu_k[vid[1]]);
    } else if (KIND == IPCB_EV) {
        const d3 e = x[2] - x[1];
        beta0 = dot(x[0] - x[1], e) / sqn(e); // closest_point.cpp:11-18
        P[0] = normalized3(e), P[1] = normalized3(cross(e, x[0] - x[1])); // tangent_basis.cpp:72-92
        d2 = sub_value(sub_point_edge(point_edge_type(x[0], x[1], x[2])), x); // distance type AUTO
        ms = blend((mu_s[vid[2]] - mu_s[vid[1]]) * beta0 + mu_s[vid[1]], mu_s[vid[0]]);
        mk = blend((mu_k[vid[2]] - mu_k[vid[1]]) * beta0 + mu_k[vid[1]], mu_k[vid[0]]);
    } else if (KIND == IPCB_EE) {
        const d3 ea = x[1] - x[0], eb = x[3] - x[2];
        keep = !(sqn(cross(ea, eb)) < c.eps[i]); // close to parallel: skipped (:123-126)
        const d3 eb_to_ea = x[0] - x[2];
        ldlt2(sqn(ea), -dot(eb, ea), sqn(eb), -dot(eb_to_ea, ea), dot(eb_to_ea, eb), beta0, beta1); // closest_point.cpp:65-88
        const d3 normal = cross(ea, eb);
        P[0] = normalized3(ea), P[1] = normalized3(cross(normal, ea)); // tangent_basis.cpp:120-134
        d2 = sub_value(sub_edge_edge(EE_AB), x); // known_dtype() == EA_EB (collisions/tangential/edge_edge.hpp:27-31)
        ms = blend((mu_s[vid[1]] - mu_s[vid[0]]) * beta0 + mu_s[vid[0]], (mu_s[vid[3]] - mu_s[vid[2]]) * beta1 + mu_s[vid[2]]);
        mk = blend((mu_k[vid[1]] - mu_k[vid[0]]) * beta0 + mu_k[vid[0]], (mu_k[vid[3]] - mu_k[vid[2]]) * beta1 + mu_k[vid[2]]);
    } else {
        const d3 b0 = x[2] - x[1], b1 = x[3] - x[1], q = x[0] - x[1];
        ldlt2(dot(b0, b0), dot(b0, b1), dot(b1, b1), dot(b0, q), dot(b1, q), beta0, beta1); // closest_point.cpp:123-137
        const d3 normal = cross(b0, b1);
        P[0] = normalized3(b0), P[1] = normalized3(cross(normal, b0)); // tangent_basis.cpp:156-171
        d2 = sub_value(sub_point_triangle(point_triangle_type(x[0], x[1], x[2], x[3])), x); // AUTO
        ms = blend(mu_s[vid[1]] + beta0 * (mu_s[vid[2]] - mu_s[vid[1]]) + beta1 * (mu_s[vid[3]] - mu_s[vid[1]]), mu_s[vid[0]]);
        mk = blend(mu_k[vid[1]] + beta0 * (mu_k[vid[2]] - mu_k[vid[1]]) + beta1 * (mu_k[vid[3]] - mu_k[vid[1]]), mu_k[vid[0]]);
    }
    // NormalPotential::force_magnitude: -kappa b'(d^2 - dmin^2) 2 d, scaled for the physical barrier
    double N = -B.kappa * barrier_df(d2 - B.dmin2, B.xhat) * 2 * sqrt(d2);
    if (B.physical) N *= B.scale;
    o.ids[i] = c.ids[i];
    o.w[i] = c.w[i], o.N[i] = N, o.mus[i] = ms, o.muk[i] = mk;
    o.beta[i] = make_double2(beta0, beta1);
    double* Pd = o.P + 6 * i;
    Pd[0] = P[0].x, Pd[1] = P[0].y, Pd[2] = P[0].z, Pd[3] = P[1].x, Pd[4] = P[1].y, Pd[5] = P[1].z;
    if (KIND == IPCB_EE) o.keep[i] = keep;
}
// compaction of the edge-edge records that were kept (order preserved)
__global__ void k_tangential_compact(int64_t n, const unsigned char* __restrict__ keep, const int* __restrict__ pos, TangOut in, TangOut out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const int p = pos[i];
    out.ids[p] = in.ids[i], out.w[p] = in.w[i], out.N[p] = in.N[i], out.mus[p] = in.mus[i], out.muk[p] = in.muk[i], out.beta[p] = in.beta[i];
    for (int k = 0; k < 6; k++) out.P[6 * size_t(p) + k] = in.P[6 * i + k];
}
__global__ void k_flags_to_int(int64_t n, const unsigned char* __restrict__ f, int* __restrict__ out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = f[i];
}

static TangOut tang_out(TangSet& t, unsigned char* keep)
{
    return { t.ids.p, t.w.p, t.N.p, t.mus.p, t.muk.p, t.beta.p, t.P.p, keep };
}
static void tang_reserve(TangSet& t, size_t n)
{
    n = std::max<size_t>(n, 1);
    t.ids.reserve(n), t.w.reserve(n), t.N.reserve(n), t.mus.reserve(n), t.muk.reserve(n), t.beta.reserve(n), t.P.reserve(6 * n);
}

void tangential_build(ipcb_ctx* ctx, const ipcb_barrier_params& bp, const double* d_mu_s, const double* d_mu_k)
{
    Stage st(ctx, "tangential_build");
    cudaStream_t s = ctx->stream;
    const BarrierDev B = make_barrier(bp, ctx->dmin);
    const MeshView m = mesh_view(ctx);
    for (int k = 0; k < 4; k++) {
        collisions_sort(ctx, k); // the tangential records follow the canonical order of the normal collisions
        const CollView c = view(ctx, k);
        TangSet& t = ctx->tang[k];
        t.count = c.n;
        if (c.n == 0) continue;
        tang_reserve(t, size_t(c.n));
        const unsigned g = grid_for(c.n, 256);
        if (k == IPCB_VV) k_tangential_build<IPCB_VV><<<g, 256, 0, s>>>(c, m, B, d_mu_s, d_mu_k, tang_out(t, nullptr));
        if (k == IPCB_EV) k_tangential_build<IPCB_EV><<<g, 256, 0, s>>>(c, m, B, d_mu_s, d_mu_k, tang_out(t, nullptr));
        if (k == IPCB_FV) k_tangential_build<IPCB_FV><<<g, 256, 0, s>>>(c, m, B, d_mu_s, d_mu_k, tang_out(t, nullptr));
        ctx->launches++;
        if (k == IPCB_EE) { // build into scratch, then keep the non-parallel ones in order
            TangSet& tmp = ctx->tang_tmp;
            tang_reserve(tmp, size_t(c.n));
            ctx->hflag.reserve(size_t(c.n)), ctx->hsel.reserve(2 * size_t(c.n) + 2);
            k_tangential_build<IPCB_EE><<<g, 256, 0, s>>>(c, m, B, d_mu_s, d_mu_k, tang_out(tmp, ctx->hflag.p));
            int* flags = ctx->hsel.p;
            int* pos = ctx->hsel.p + c.n + 1;
            k_flags_to_int<<<g, 256, 0, s>>>(c.n, ctx->hflag.p, flags);
            size_t bytes = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, bytes, flags, pos, int(c.n), s);
            ctx->cubtmp.reserve(bytes + 256);
            cub::DeviceScan::ExclusiveSum(ctx->cubtmp.p, bytes, flags, pos, int(c.n), s);
            k_tangential_compact<<<g, 256, 0, s>>>(c.n, ctx->hflag.p, pos, tang_out(tmp, nullptr), tang_out(t, nullptr));
            ctx->launches += 4;
            int last_pos = 0;
            unsigned char last_keep = 0;
            IPCB_CUDA(cudaMemcpyAsync(&last_pos, pos + (c.n - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
            IPCB_CUDA(cudaMemcpyAsync(&last_keep, ctx->hflag.p + (c.n - 1), 1, cudaMemcpyDeviceToHost, s));
            IPCB_CUDA(cudaStreamSynchronize(s));
            t.count = int64_t(last_pos) + int64_t(last_keep);
        }
    }
    IPCB_CUDA(cudaGetLastError());
    ctx->tang_valid = true;
}

// ---- smooth friction mollifier and smooth mu (friction/smooth_friction_mollifier.cpp, friction/smooth_mu.cpp) ---------
__device__ inline double sf_f0(double y, double e) { return fabs(y) >= e ? y : y * y * (1 - y / (3 * e)) / e + e / 3; }
__device__ inline double sf_f1_over_x(double y, double e) { return fabs(y) >= e ? 1 / y : (2 - y / e) / e; }
__device__ inline double sf_f2x_minus_f1_over_x3(double y, double e) { return fabs(y) >= e ? -1 / (y * y * y) : -1 / (y * e * e); }
__device__ inline double smooth_mu(double y, double mu_s, double mu_k, double e)
{
    if (mu_s == mu_k || fabs(y) >= e) return mu_k;
    const double z = fabs(y) / e;
    if (fabs(y) < 0.5 * e) return 2 * (mu_k - mu_s) * z * z + mu_s;
    return -2 * (mu_k - mu_s) * (z * (z - 2) + 1) + mu_k;
}
__device__ inline double smooth_mu_f0(double y, double mu_s, double mu_k, double e)
{
    if (mu_s == mu_k || fabs(y) >= e) return mu_k * sf_f0(y, e);
    const double delta_mu = mu_k - mu_s, z = fabs(y) / e;
    if (fabs(y) < 0.5 * e) return y * z * (z * (z * (1 - 0.4 * z) * delta_mu - mu_s / 3.0) + mu_s) + (9.0 / 16.0) * e * mu_k - (11.0 / 48.0) * e * mu_s;
    return y * z * (z * (z * (0.4 * z - 2) * delta_mu + (3 * mu_k - (10.0 / 3.0) * mu_s)) + (2 * mu_s - mu_k)) + 0.6 * e * mu_k - (4.0 / 15.0) * e * mu_s;
}
__device__ inline double smooth_mu_f1_over_x(double y, double mu_s, double mu_k, double e) { return smooth_mu(y, mu_s, mu_k, e) * sf_f1_over_x(y, e); }
__device__ inline double smooth_mu_f2x_minus_f1_over_x3(double y, double mu_s, double mu_k, double e)
{
    if (mu_s == mu_k || fabs(y) >= e) return mu_k * sf_f2x_minus_f1_over_x3(y, e);
    const double delta_mu = mu_k - mu_s, z = 1 / e;
    if (fabs(y) < 0.5 * e) return z * z * (z * (8 - 6 * y * z) * delta_mu - mu_s / y);
    return z * z * (z * (6 * y * z - 16) * delta_mu + (9 * mu_k - 10 * mu_s) / y);
}

// ---- per-collision slip --------------------------------------------------------------------------------------------
struct TangView {
    int kind;
    int64_t n;
    const int2* ids;
    const double *w, *N, *mus, *muk;
    const double2* beta;
    const double* P;
};
static TangView tang_view(const ipcb_ctx* ctx, int k)
{
    const TangSet& t = ctx->tang[k];
    return { k, t.count, t.ids.p, t.w.p, t.N.p, t.mus.p, t.muk.p, t.beta.p, t.P.p };
}
// relative-velocity coefficients (relative_velocity.cpp)
template <int KIND> __device__ inline void tang_gamma(double2 b, double* g)
{
    if (KIND == IPCB_VV) g[0] = 1, g[1] = -1;
    else if (KIND == IPCB_EV) g[0] = 1, g[1] = b.x - 1, g[2] = -b.x;
    else if (KIND == IPCB_EE) g[0] = 1 - b.x, g[1] = b.x, g[2] = b.y - 1, g[3] = -b.y;
    else g[0] = 1, g[1] = b.x + b.y - 1, g[2] = -b.x, g[3] = -b.y;
}
struct Slip {
    int vid[4];
    double gamma[4];
    d3 P0, P1;
    double u0, u1, nu, scale, mus, muk;
};
template <int KIND> __device__ inline Slip load_slip(const TangView& t, const MeshView& m, int64_t i)
{
    constexpr int NP = KIND == IPCB_VV ? 2 : (KIND == IPCB_EV ? 3 : 4);
    Slip s;
    d3 v[4];
    load_stencil(KIND, t.ids[i], m, s.vid, v); // m.X holds the velocities
    tang_gamma<KIND>(t.beta[i], s.gamma);
    d3 rel = mk3(0, 0, 0);
#pragma unroll
    for (int a = 0; a < NP; a++) rel = rel + s.gamma[a] * v[a];
    const double* P = t.P + 6 * i;
    s.P0 = mk3(P[0], P[1], P[2]), s.P1 = mk3(P[3], P[4], P[5]);
    s.u0 = dot(s.P0, rel), s.u1 = dot(s.P1, rel);
    s.nu = sqrt(s.u0 * s.u0 + s.u1 * s.u1);
    s.scale = t.w[i] * t.N[i];
    s.mus = t.mus[i], s.muk = t.muk[i];
    return s;
}

// ---- energy (tangential_potential.cpp:162-187) ------------------------------------------------------------------------
constexpr int FBLOCK = 256;
template <int KIND> __global__ void __launch_bounds__(FBLOCK) k_friction_energy(TangView t, MeshView m, double eps_v, double* __restrict__ partial)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double e = 0;
    if (i < t.n) {
        const Slip s = load_slip<KIND>(t, m, i);
        e = s.scale * smooth_mu_f0(s.nu, s.mus, s.muk, eps_v);
    }
    __shared__ double sm[FBLOCK];
    sm[threadIdx.x] = e;
    __syncthreads();
    for (int k = FBLOCK / 2; k > 0; k >>= 1) { // fixed-order tree: deterministic
        if (threadIdx.x < k) sm[threadIdx.x] += sm[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}
__global__ void k_friction_sum(int n, const double* __restrict__ partial, double* __restrict__ out)
{
    __shared__ double sm[1024];
    double s = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int k = blockDim.x / 2; k > 0; k >>= 1) {
        if (threadIdx.x < k) sm[threadIdx.x] += sm[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sm[0];
}
static void require_tangential(const ipcb_ctx* ctx)
{
    if (!ctx->tang_valid) throw Error("no tangential collision set has been built on this context");
}
void friction_energy(ipcb_ctx* ctx, double eps_v, double* d_out)
{
    require_tangential(ctx);
    Stage st(ctx, "friction_energy");
    cudaStream_t s = ctx->stream;
    const MeshView m = mesh_view(ctx);
    size_t nblocks = 0;
    for (int k = 0; k < 4; k++) nblocks += grid_for(ctx->tang[k].count, FBLOCK);
    ctx->dScalar.reserve(nblocks + 16);
    size_t off = 0;
    for (int k = 0; k < 4; k++) {
        const TangView t = tang_view(ctx, k);
        const unsigned g = grid_for(t.n, FBLOCK);
        if (!g) continue;
        if (k == IPCB_VV) k_friction_energy<IPCB_VV><<<g, FBLOCK, 0, s>>>(t, m, eps_v, ctx->dScalar.p + off);
        if (k == IPCB_EV) k_friction_energy<IPCB_EV><<<g, FBLOCK, 0, s>>>(t, m, eps_v, ctx->dScalar.p + off);
        if (k == IPCB_EE) k_friction_energy<IPCB_EE><<<g, FBLOCK, 0, s>>>(t, m, eps_v, ctx->dScalar.p + off);
        if (k == IPCB_FV) k_friction_energy<IPCB_FV><<<g, FBLOCK, 0, s>>>(t, m, eps_v, ctx->dScalar.p + off);
        off += g;
        ctx->launches++;
    }
    k_friction_sum<<<1, 1024, 0, s>>>(int(nblocks), ctx->dScalar.p, d_out);
    ctx->launches++;
    IPCB_CUDA(cudaGetLastError());
}

// ---- gradient (tangential_potential.cpp:189-237) ----------------------------------------------------------------------
template <int KIND> __global__ void __launch_bounds__(256) k_friction_gradient(TangView t, MeshView m, double eps_v, double* __restrict__ grad)
{
    constexpr int NP = KIND == IPCB_VV ? 2 : (KIND == IPCB_EV ? 3 : 4);
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= t.n) return;
    const Slip s = load_slip<KIND>(t, m, i);
    const double c = smooth_mu_f1_over_x(s.nu, s.mus, s.muk, eps_v) * s.scale;
    const d3 f = (c * s.u0) * s.P0 + (c * s.u1) * s.P1; // P (c u)
#pragma unroll
    for (int a = 0; a < NP; a++) {
        double* g = grad + 3 * (size_t)s.vid[a];
        const d3 ga = s.gamma[a] * f;
        if (ga.x != 0.0) atomicAdd(g, ga.x);
        if (ga.y != 0.0) atomicAdd(g + 1, ga.y);
        if (ga.z != 0.0) atomicAdd(g + 2, ga.z);
    }
}
void friction_gradient(ipcb_ctx* ctx, double eps_v, double* d_grad)
{
    require_tangential(ctx);
    Stage st(ctx, "friction_gradient");
    cudaStream_t s = ctx->stream;
    IPCB_CUDA(cudaMemsetAsync(d_grad, 0, sizeof(double) * 3 * size_t(ctx->nV), s));
    const MeshView m = mesh_view(ctx);
    for (int k = 0; k < 4; k++) {
        const TangView t = tang_view(ctx, k);
        const unsigned g = grid_for(t.n, 256);
        if (!g) continue;
        if (k == IPCB_VV) k_friction_gradient<IPCB_VV><<<g, 256, 0, s>>>(t, m, eps_v, d_grad);
        if (k == IPCB_EV) k_friction_gradient<IPCB_EV><<<g, 256, 0, s>>>(t, m, eps_v, d_grad);
        if (k == IPCB_EE) k_friction_gradient<IPCB_EE><<<g, 256, 0, s>>>(t, m, eps_v, d_grad);
        if (k == IPCB_FV) k_friction_gradient<IPCB_FV><<<g, 256, 0, s>>>(t, m, eps_v, d_grad);
        ctx->launches++;
    }
    IPCB_CUDA(cudaGetLastError());
}

// ---- Hessian (tangential_potential.cpp:239-325) -----------------------------------------------------------------------
// K = P M P^T with the inner 2x2 matrix M of the three branches; the "in between" branch projects M analytically:
// M = scale [f1/|u| I + f2 u u^T] has the eigenpairs (scale (f1/|u| + f2 |u|^2), u / |u|) and (scale f1/|u|, u-perp / |u|).
template <int KIND>
__global__ void __launch_bounds__(128) k_friction_hessian(TangView t, MeshView m, double eps_v, int psd_mode, int64_t gi0, int64_t inc0, HessOut out)
{
    constexpr int NP = KIND == IPCB_VV ? 2 : (KIND == IPCB_EV ? 3 : 4);
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= t.n) return;
    const Slip s = load_slip<KIND>(t, m, i);
    write_record<NP>(out, gi0 + i, inc0 + i * NP, s.vid);
    const double f1ox = smooth_mu_f1_over_x(s.nu, s.mus, s.muk, eps_v);
    double M00 = 0, M01 = 0, M11 = 0;
    if (s.nu > eps_v) { // is_dynamic: mu N f1/|u| (I - u u^T / |u|^2)
        if (!(psd_mode != IPCB_PSD_NONE && s.scale <= 0)) {
            const double c = s.scale * f1ox / (s.nu * s.nu);
            const double p0 = -s.u1, p1 = s.u0;
            M00 = c * p0 * p0, M01 = c * p0 * p1, M11 = c * p1 * p1;
        }
    } else if (s.nu == 0) {
        if (!(psd_mode != IPCB_PSD_NONE && s.scale <= 0)) M00 = M11 = s.scale * f1ox;
    } else {
        const double f2 = smooth_mu_f2x_minus_f1_over_x3(s.nu, s.mus, s.muk, eps_v);
        M00 = (f2 * s.u0 * s.u0 + f1ox) * s.scale, M11 = (f2 * s.u1 * s.u1 + f1ox) * s.scale, M01 = (f2 * s.u0 * s.u1) * s.scale;
        if (psd_mode != IPCB_PSD_NONE) {
            double l1 = s.scale * (f1ox + f2 * s.nu * s.nu), l2 = s.scale * f1ox; // along u, along u-perp
            if (l1 < 0 || l2 < 0) { // project_to_psd leaves a PSD matrix untouched (eigen_ext.tpp:84-86)
                l1 = l1 < 0 ? (psd_mode == IPCB_PSD_CLAMP ? 0.0 : -l1) : l1;
                l2 = l2 < 0 ? (psd_mode == IPCB_PSD_CLAMP ? 0.0 : -l2) : l2;
                const double a0 = s.u0 / s.nu, a1 = s.u1 / s.nu;
                M00 = l1 * a0 * a0 + l2 * a1 * a1, M11 = l1 * a1 * a1 + l2 * a0 * a0, M01 = (l1 - l2) * a0 * a1;
            }
        }
    }
    const double Pm[2][3] = { { s.P0.x, s.P0.y, s.P0.z }, { s.P1.x, s.P1.y, s.P1.z } };
    double K[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) K[3 * r + c] = (Pm[0][r] * M00 + Pm[1][r] * M01) * Pm[0][c] + (Pm[0][r] * M01 + Pm[1][r] * M11) * Pm[1][c];
    // upper-triangular vertex blocks: slot (column point a, row point b >= a) = gamma_a gamma_b K (K is symmetric up to
    // rounding; the stored block is rows of b, columns of a: entry 3 r + c = gamma_b gamma_a K[r][c])
    unsigned short masks[HSLOTS];
#pragma unroll
    for (int k = 0; k < HSLOTS; k++) masks[k] = 0;
    double* rec = out.blk + size_t(i) * (tri_count(NP) * 9);
#pragma unroll
    for (int a = 0; a < NP; a++)
#pragma unroll
        for (int b = a; b < NP; b++) {
            const double gg = s.gamma[b] * s.gamma[a];
            unsigned mask = 0;
#pragma unroll
            for (int k = 0; k < 9; k++) {
                const double v = gg * K[k];
                rec[tri_slot(NP, a, b) * 9 + k] = v;
                mask |= unsigned(v != 0.0) << k;
            }
            masks[a * 4 + b] = (unsigned short)mask;
            if (a != b) masks[b * 4 + a] = (unsigned short)mask_transpose(mask);
        }
    for (int k = 0; k < HSLOTS; k++) out.mask[(gi0 + i) * HSLOTS + k] = masks[k];
}
void friction_hessian(ipcb_ctx* ctx, double eps_v, int psd_mode)
{
    require_tangential(ctx);
    cudaStream_t s = ctx->stream;
    const int64_t nk[4] = { ctx->tang[0].count, ctx->tang[1].count, ctx->tang[2].count, ctx->tang[3].count };
    if (nk[0] + nk[1] + nk[2] + nk[3] == 0) return hessian_empty(ctx);
    const MeshView m = mesh_view(ctx);
    {
        Stage st(ctx, "friction_hessian_local");
        HessOut outs[4];
        hessian_records(ctx, nk, 0, ctx->nV, outs);
        const int64_t gi0[4] = { 0, nk[0], nk[0] + nk[1], nk[0] + nk[1] + nk[2] };
        const int64_t inc0[4] = { 0, 2 * nk[0], 2 * nk[0] + 3 * nk[1], 2 * nk[0] + 3 * nk[1] + 4 * nk[2] };
        for (int k = 0; k < 4; k++) {
            const TangView t = tang_view(ctx, k);
            const unsigned g = grid_for(t.n, 128);
            if (!g) continue;
            if (k == IPCB_VV) k_friction_hessian<IPCB_VV><<<g, 128, 0, s>>>(t, m, eps_v, psd_mode, gi0[k], inc0[k], outs[k]);
            if (k == IPCB_EV) k_friction_hessian<IPCB_EV><<<g, 128, 0, s>>>(t, m, eps_v, psd_mode, gi0[k], inc0[k], outs[k]);
            if (k == IPCB_EE) k_friction_hessian<IPCB_EE><<<g, 128, 0, s>>>(t, m, eps_v, psd_mode, gi0[k], inc0[k], outs[k]);
            if (k == IPCB_FV) k_friction_hessian<IPCB_FV><<<g, 128, 0, s>>>(t, m, eps_v, psd_mode, gi0[k], inc0[k], outs[k]);
            ctx->launches++;
        }
        IPCB_CUDA(cudaGetLastError());
    }
    hessian_assemble(ctx, nk);
}

} // namespace ipcb
