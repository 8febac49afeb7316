// potential.cu — BarrierPotential::operator() / gradient / hessian.
//
// Replaces (reference src/ipc/): potentials/potential.cpp:36-222 (assembly),
// potentials/normal_potential.cpp:127-232 (per-collision w*m(x)*f(d(x))
// calculus), potentials/barrier_potential.cpp:62-99 (barrier scalars),
// utils/eigen_ext.tpp:56-108 (project_to_psd), utils/local_to_global.hpp:21-45,
// 263-305 (scatter, exact-zero skipping).
//
// GPU formulation
//  * energy: one thread per collision, fixed-order two-level reduction
//    (deterministic, unlike the reference's tbb::parallel_reduce);
//  * gradient: one thread per collision, FP64 atomicAdd (RED) into the dense
//    3N vector (DOF = 3*vertex + axis);
//  * hessian: one thread per collision builds the local (3n x 3n) matrix in
//    block form.  PSD projection exploits translation invariance: a squared
//    distance only depends on differences of its n points, so the local matrix
//    is W S W^T with W = Helmert(n) ⊗ I3 (orthonormal columns).  S is only
//    3(n-1) x 3(n-1) (3 / 6 / 9 instead of 6 / 9 / 12) and is diagonalised with
//    cyclic Jacobi; projecting S projects the full matrix exactly.
//  * assembly works on 3x3 VERTEX blocks instead of scalar triplets: every
//    local matrix emits n^2 blocks keyed by (column vertex, row vertex) plus a
//    9-bit mask of its non-zero entries; blocks are radix-sorted (cub,
//    plumbing), run-length summed, and expanded to compressed columns with the
//    reference's exact pattern: an entry exists iff some local matrix had an
//    exactly non-zero value there (duplicates summed, never pruned).
#include "ctx.cuh"
#include "geom.cuh"
#include "hessian_fast.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace ipcb {

struct BarrierDev {
    double xhat;  // (2 dmin + dhat) dhat
    double dmin2; // dmin^2
    double kappa;
    double scale; // physical barrier factor dhat / xhat^2, else 1
    int physical;
    __device__ double f(double d2) const { return kappa * (physical ? barrier_f(d2 - dmin2, xhat) * scale : barrier_f(d2 - dmin2, xhat)); }
    __device__ double df(double d2) const { return kappa * (physical ? barrier_df(d2 - dmin2, xhat) * scale : barrier_df(d2 - dmin2, xhat)); }
    __device__ double ddf(double d2) const { return kappa * (physical ? barrier_ddf(d2 - dmin2, xhat) * scale : barrier_ddf(d2 - dmin2, xhat)); }
};
static BarrierDev make_barrier(const ipcb_barrier_params& bp, double dmin)
{
    BarrierDev b;
    b.xhat = (2 * dmin + bp.dhat) * bp.dhat;
    b.dmin2 = dmin * dmin;
    b.kappa = bp.stiffness;
    b.physical = bp.use_physical_barrier != 0;
    b.scale = b.physical ? bp.dhat / (b.xhat * b.xhat) : 1.0;
    return b;
}

struct CollView {
    int kind;
    int64_t n;
    const int2* ids;
    const double* w;
    const double* eps;
    const unsigned char* dt;
};
struct MeshView {
    const int2* E;
    const int4* F;
    const double4* X;
};

__device__ inline d3 ldx(const double4* X, int i) { return load_vertex(X, i); }
// stencil vertex ids + positions (candidates/*.cpp vertex_ids): VV [v0,v1]; EV [v,e0,e1];
// EE [ea0,ea1,eb0,eb1]; FV [v,f0,f1,f2]
__device__ inline int load_stencil(int kind, int2 id, const MeshView& m, int* vid, d3* x)
{
    int n;
    if (kind == IPCB_VV) {
        vid[0] = id.x, vid[1] = id.y, n = 2;
    } else if (kind == IPCB_EV) {
        const int2 e = __ldg(m.E + id.x);
        vid[0] = id.y, vid[1] = e.x, vid[2] = e.y, n = 3;
    } else if (kind == IPCB_EE) {
        const int2 ea = __ldg(m.E + id.x), eb = __ldg(m.E + id.y);
        vid[0] = ea.x, vid[1] = ea.y, vid[2] = eb.x, vid[3] = eb.y, n = 4;
    } else {
        const int4 f = __ldg(m.F + id.x);
        vid[0] = id.y, vid[1] = f.x, vid[2] = f.y, vid[3] = f.z, n = 4;
    }
    for (int k = 0; k < n; k++) x[k] = ldx(m.X, vid[k]);
    return n;
}
__device__ inline Sub collision_sub(int kind, int dt)
{
    // known distance types: collisions/normal/{edge_vertex,face_vertex}.hpp:28-32, edge_edge.hpp:96
    return kind == IPCB_VV ? Sub { 0, 0, 1, 0, 0 } : kind == IPCB_EV ? sub_point_edge(PE_E) : kind == IPCB_EE ? sub_edge_edge(dt) : sub_point_triangle(PT_T);
}

// ---------------------------------------------------------------------------
// energy (potential.cpp:36-56, normal_potential.cpp:127-135)
constexpr int EBLOCK = 256;
__global__ void __launch_bounds__(EBLOCK) k_energy(CollView c, MeshView m, BarrierDev B, double* __restrict__ partial)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double e = 0;
    if (i < c.n) {
        int vid[4];
        d3 x[4];
        load_stencil(c.kind, c.ids[i], m, vid, x);
        const double d = sub_value(collision_sub(c.kind, c.kind == IPCB_EE ? c.dt[i] : 0), x);
        double mol = 1.0;
        if (c.kind == IPCB_EE) mol = moll(sqn(cross(x[1] - x[0], x[3] - x[2])), c.eps[i]);
        e = c.w[i] * mol * B.f(d);
    }
    __shared__ double sm[EBLOCK];
    sm[threadIdx.x] = e;
    __syncthreads();
    for (int s = EBLOCK / 2; s > 0; s >>= 1) { // fixed-order tree: deterministic
        if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}
__global__ void k_sum_partials(int n, const double* __restrict__ partial, double* __restrict__ out)
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

static CollView view(const ipcb_ctx* ctx, int k)
{
    const CollisionSet& cs = ctx->coll[k];
    return { k, cs.count, cs.ids.p, cs.w.p, cs.eps.p, cs.dtype.p };
}
static MeshView mesh_view(const ipcb_ctx* ctx) { return { ctx->dE.p, ctx->dF.p, ctx->X0.p }; }

void barrier_energy(ipcb_ctx* ctx, const ipcb_barrier_params& bp, double* d_out)
{
    Stage st(ctx, "energy");
    cudaStream_t s = ctx->stream;
    const BarrierDev B = make_barrier(bp, ctx->dmin);
    size_t nblocks = 0;
    for (int k = 0; k < 4; k++) nblocks += grid_for(ctx->coll[k].count, EBLOCK);
    ctx->dScalar.reserve(nblocks + 16);
    size_t off = 0;
    for (int k = 0; k < 4; k++) {
        const unsigned g = grid_for(ctx->coll[k].count, EBLOCK);
        if (!g) continue;
        k_energy<<<g, EBLOCK, 0, s>>>(view(ctx, k), mesh_view(ctx), B, ctx->dScalar.p + off);
        off += g;
        ctx->launches++;
    }
    k_sum_partials<<<1, 1024, 0, s>>>(int(nblocks), ctx->dScalar.p, d_out);
    ctx->launches++;
    IPCB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// gradient (potential.cpp:58-94, normal_potential.cpp:137-170, local_to_global.hpp:21-45)
__global__ void __launch_bounds__(128) k_gradient(CollView c, MeshView m, BarrierDev B, double* __restrict__ grad)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    int vid[4];
    d3 x[4];
    const int n = load_stencil(c.kind, c.ids[i], m, vid, x);
    const double w = c.w[i];
    LocalDeriv D;
    zero_local(D);
    double mol = 1.0, s = 0.0, eps = 0.0;
    if (c.kind == IPCB_EE) {
        eps = c.eps[i];
        s = sqn(cross(x[1] - x[0], x[3] - x[2]));
        mol = moll(s, eps);
        if (mol <= 0) return; // gradient is exactly zero (normal_potential.cpp:143-147)
    }
    const double d = sub_deriv(collision_sub(c.kind, c.kind == IPCB_EE ? c.dt[i] : 0), x, D);
    const double gf = B.df(d);
    double g[12];
    if (c.kind != IPCB_EE) {
        for (int k = 0; k < 3 * n; k++) g[k] = (w * gf) * D.g[k];
    } else {
        const double f = B.f(d);
        for (int k = 0; k < 12; k++) g[k] = (w * mol * gf) * D.g[k];
        if (s < eps) {
            // second use of the local container for the mollifier's cross-norm derivative
            zero_local(D);
            cross_sqnorm_deriv(x, D);
            const double dm = moll_d(s, eps);
            for (int k = 0; k < 12; k++) g[k] = (w * f) * (dm * D.g[k]) + g[k];
        }
    }
    for (int a = 0; a < n; a++)
        for (int k = 0; k < 3; k++) {
            const double v = g[3 * a + k];
            if (v != 0.0) atomicAdd(grad + 3 * (size_t)vid[a] + k, v);
        }
}

void barrier_gradient(ipcb_ctx* ctx, const ipcb_barrier_params& bp, double* d_grad)
{
    Stage st(ctx, "gradient");
    cudaStream_t s = ctx->stream;
    const BarrierDev B = make_barrier(bp, ctx->dmin);
    IPCB_CUDA(cudaMemsetAsync(d_grad, 0, sizeof(double) * 3 * size_t(ctx->nV), s));
    for (int k = 0; k < 4; k++) {
        if (!ctx->coll[k].count) continue;
        k_gradient<<<grid_for(ctx->coll[k].count, 128), 128, 0, s>>>(view(ctx, k), mesh_view(ctx), B, d_grad);
        ctx->launches++;
    }
    IPCB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// PSD projection of W S W^T through S (see header)
template <int NP> struct Helmert {
    // column a (0..NP-2): entries i <= a: 1/sqrt((a+1)(a+2)); i = a+1: -(a+1)/sqrt((a+1)(a+2)); else 0
    __device__ static double at(int i, int a)
    {
        const double r = rsqrt(double((a + 1) * (a + 2)));
        return i <= a ? r : (i == a + 1 ? -double(a + 1) * r : 0.0);
    }
};

// cyclic Jacobi on a symmetric NR x NR matrix A (row-major); V receives the eigenvectors (columns)
template <int NR> __device__ inline void jacobi(double* A, double* V)
{
    for (int i = 0; i < NR; i++)
        for (int j = 0; j < NR; j++) V[i * NR + j] = i == j ? 1.0 : 0.0;
    double total = 0;
    for (int i = 0; i < NR * NR; i++) total = fma(A[i], A[i], total);
    const double stop = total * 1e-32; // off-diagonal Frobenius^2 below (1e-16 ||A||)^2
    for (int sweep = 0; sweep < 40; sweep++) {
        double off = 0;
        for (int p = 0; p < NR; p++)
            for (int q = p + 1; q < NR; q++) off = fma(A[p * NR + q], A[p * NR + q], off);
        if (2 * off <= stop) break;
        for (int p = 0; p < NR - 1; p++) {
            for (int q = p + 1; q < NR; q++) {
                const double apq = A[p * NR + q];
                if (apq == 0.0) continue;
                const double app = A[p * NR + p], aqq = A[q * NR + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
                const double c = rsqrt(fma(t, t, 1.0)), s = t * c;
                for (int k = 0; k < NR; k++) { // columns p, q
                    const double akp = A[k * NR + p], akq = A[k * NR + q];
                    A[k * NR + p] = fma(c, akp, -s * akq);
                    A[k * NR + q] = fma(s, akp, c * akq);
                }
                for (int k = 0; k < NR; k++) { // rows p, q
                    const double apk = A[p * NR + k], aqk = A[q * NR + k];
                    A[p * NR + k] = fma(c, apk, -s * aqk);
                    A[q * NR + k] = fma(s, apk, c * aqk);
                }
                for (int k = 0; k < NR; k++) {
                    const double vkp = V[k * NR + p], vkq = V[k * NR + q];
                    V[k * NR + p] = fma(c, vkp, -s * vkq);
                    V[k * NR + q] = fma(s, vkp, c * vkq);
                }
            }
        }
    }
}

// H: 4 x 4 blocks (9 doubles per block, row-major); the projection acts on the NP stencil points
// pt[0..NP) that the local matrix really involves — all other rows / columns are exactly zero and
// must stay exactly zero (they are not entries of the reference's sparse matrix)
template <int NP> __device__ inline void project_psd(double* H, int mode, const int* pt)
{
    constexpr int M = NP - 1, NR = 3 * M;
    double S[NR * NR], V[NR * NR];
    // S_ab = sum_ij Hel[i][a] Hel[j][b] H_ij  (3x3 blocks)
    for (int a = 0; a < M; a++)
        for (int b = 0; b < M; b++)
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) {
                    double acc = 0;
                    for (int i = 0; i < NP; i++) {
                        const double hia = Helmert<NP>::at(i, a);
                        if (hia == 0.0) continue;
                        double row = 0;
                        for (int j = 0; j < NP; j++) row = fma(Helmert<NP>::at(j, b), H[(pt[i] * 4 + pt[j]) * 9 + 3 * r + c], row);
                        acc = fma(hia, row, acc);
                    }
                    S[(3 * a + r) * NR + 3 * b + c] = acc;
                }
    // symmetrise (the local matrix is symmetric up to rounding; Eigen reads one triangle)
    for (int i = 0; i < NR; i++)
        for (int j = i + 1; j < NR; j++) {
            const double v = 0.5 * (S[i * NR + j] + S[j * NR + i]);
            S[i * NR + j] = S[j * NR + i] = v;
        }
    jacobi<NR>(S, V);
    double lam[NR];
    double lmin = INFINITY;
    for (int i = 0; i < NR; i++) {
        lam[i] = S[i * NR + i];
        lmin = fmin(lmin, lam[i]);
    }
    if (lmin >= 0.0) return; // A is returned unchanged (eigen_ext.tpp:84-86)
    for (int i = 0; i < NR; i++)
        if (lam[i] < 0.0) lam[i] = mode == IPCB_PSD_CLAMP ? 0.0 : fabs(lam[i]);
    // S+ = V diag(lam) V^T (into S)
    for (int i = 0; i < NR; i++)
        for (int j = i; j < NR; j++) {
            double acc = 0;
            for (int k = 0; k < NR; k++) acc = fma(V[i * NR + k] * lam[k], V[j * NR + k], acc);
            S[i * NR + j] = S[j * NR + i] = acc;
        }
    // H_ij = sum_ab Hel[i][a] Hel[j][b] S+_ab
    for (int i = 0; i < NP; i++)
        for (int j = 0; j < NP; j++)
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) {
                    double acc = 0;
                    for (int a = 0; a < M; a++) {
                        const double hia = Helmert<NP>::at(i, a);
                        if (hia == 0.0) continue;
                        double row = 0;
                        for (int b = 0; b < M; b++) row = fma(Helmert<NP>::at(j, b), S[(3 * a + r) * NR + 3 * b + c], row);
                        acc = fma(hia, row, acc);
                    }
                    H[(pt[i] * 4 + pt[j]) * 9 + 3 * r + c] = acc;
                }
}

// ---------------------------------------------------------------------------
// local Hessians -> vertex blocks (potential.cpp:96-154, normal_potential.cpp:172-232)
// `list` (optional): indices of the collisions to process (the ones the fast path handed over)
template <int KIND>
__global__ void __launch_bounds__(128)
    k_hessian_local(CollView c, MeshView m, BarrierDev B, int psd_mode, int64_t block_offset, unsigned long long* __restrict__ hkey,
                    double* __restrict__ hval, unsigned short* __restrict__ hmask, const int* __restrict__ list, int64_t nlist)
{
    constexpr int NP = KIND == IPCB_VV ? 2 : (KIND == IPCB_EV ? 3 : 4);
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (tid >= (list ? nlist : c.n)) return;
    const int64_t i = list ? list[tid] : tid;
    int vid[4];
    d3 x[4];
    load_stencil(KIND, c.ids[i], m, vid, x);
    const double w = c.w[i];
    LocalDeriv D;
    zero_local(D);
    const double d = sub_deriv(collision_sub(KIND, KIND == IPCB_EE ? c.dt[i] : 0), x, D);
    bool project = psd_mode != IPCB_PSD_NONE;
    if (KIND != IPCB_EE) {
        const double a = w * B.ddf(d), b = w * B.df(d);
        for (int bi = 0; bi < NP; bi++)
            for (int bj = 0; bj < NP; bj++)
                for (int r = 0; r < 3; r++)
                    for (int cc = 0; cc < 3; cc++) {
                        double& h = D.H[(bi * 4 + bj) * 9 + 3 * r + cc];
                        h = a * D.g[3 * bi + r] * D.g[3 * bj + cc] + b * h;
                    }
    } else {
        const double eps = c.eps[i];
        const double s = sqn(cross(x[1] - x[0], x[3] - x[2]));
        const double mol = moll(s, eps);
        const double f = B.f(d);
        LocalDeriv S;
        zero_local(S);
        double dm = 0, ddm = 0;
        if (s < eps) {
            cross_sqnorm_deriv(x, S);
            dm = moll_d(s, eps), ddm = moll_dd(s, eps);
        }
        if (mol <= 0) { // (w f) Hess(m), returned WITHOUT projection (normal_potential.cpp:184-189)
            for (int bi = 0; bi < 4; bi++)
                for (int bj = 0; bj < 4; bj++)
                    for (int r = 0; r < 3; r++)
                        for (int cc = 0; cc < 3; cc++) {
                            const int e = (bi * 4 + bj) * 9 + 3 * r + cc;
                            D.H[e] = (w * f) * ((dm * S.H[e]) + ((ddm * S.g[3 * bi + r]) * S.g[3 * bj + cc]));
                        }
            project = false;
        } else {
            const double gf = B.df(d), hf = B.ddf(d), wm = w * mol;
            for (int bi = 0; bi < 4; bi++)
                for (int bj = 0; bj < 4; bj++)
                    for (int r = 0; r < 3; r++)
                        for (int cc = 0; cc < 3; cc++) {
                            const int e = (bi * 4 + bj) * 9 + 3 * r + cc;
                            const int gi = 3 * bi + r, gj = 3 * bj + cc;
                            const double hess_m = (dm * S.H[e]) + ((ddm * S.g[gi]) * S.g[gj]);
                            const double gm_i = dm * S.g[gi], gm_j = dm * S.g[gj];
                            const double cross_ij = (w * gf) * D.g[gi] * gm_j, cross_ji = (w * gf) * D.g[gj] * gm_i;
                            D.H[e] = (w * f) * hess_m + cross_ij + cross_ji + (wm * hf) * D.g[gi] * D.g[gj] + (wm * gf) * D.H[e];
                        }
        }
    }
    if (project) {
        const int all[4] = { 0, 1, 2, 3 };
        if (KIND == IPCB_EE) {
            // an edge-edge collision whose distance type is a vertex-vertex / vertex-edge pair and whose
            // mollifier is inactive only involves 2 / 3 of its 4 points: the other points' rows are
            // exactly zero, span an invariant subspace of the projection and must stay exactly zero
            int pt[4], np = 0;
            for (int p = 0; p < 4; p++) {
                bool any = false;
                for (int q = 0; q < 4; q++)
                    for (int k = 0; k < 9; k++) any |= D.H[(p * 4 + q) * 9 + k] != 0.0 || D.H[(q * 4 + p) * 9 + k] != 0.0;
                if (any) pt[np++] = p;
            }
            if (np == 4) project_psd<4>(D.H, psd_mode, pt);
            else if (np == 3) project_psd<3>(D.H, psd_mode, pt);
            else if (np == 2) project_psd<2>(D.H, psd_mode, pt);
        } else {
            project_psd<NP>(D.H, psd_mode, all);
        }
    }
    // emit NP*NP vertex blocks keyed (column vertex, row vertex)
    const int64_t base = block_offset + i * (NP * NP);
    for (int bi = 0; bi < NP; bi++)
        for (int bj = 0; bj < NP; bj++) {
            const int64_t e = base + bi * NP + bj;
            hkey[e] = ((unsigned long long)(unsigned)vid[bj] << 32) | (unsigned)vid[bi];
            unsigned short mask = 0;
            for (int k = 0; k < 9; k++) {
                const double v = D.H[(bi * 4 + bj) * 9 + k];
                hval[e * 9 + k] = v;
                mask |= (v != 0.0) << k; // exact zeros are not entries (local_to_global.hpp:290-291)
            }
            hmask[e] = mask;
        }
}

// PSD-projected local Hessians through the analytic 3+p dimensional subspace (hessian_fast.cuh).
// Edge-edge collisions whose mollifier is active at X are appended to `slow` for the general kernel.
template <int KIND>
__global__ void __launch_bounds__(128)
    k_hessian_fast(CollView c, MeshView m, BarrierDev B, int psd_mode, int64_t block_offset, unsigned long long* __restrict__ hkey,
                   double* __restrict__ hval, unsigned short* __restrict__ hmask, int* __restrict__ slow, unsigned long long* slow_count)
{
    constexpr int NP = KIND == IPCB_VV ? 2 : (KIND == IPCB_EV ? 3 : 4);
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool is_slow = false;
    if (i < c.n) {
        int vid[4];
        d3 x[4];
        load_stencil(KIND, c.ids[i], m, vid, x);
        const Sub sb = collision_sub(KIND, KIND == IPCB_EE ? c.dt[i] : 0);
        if (KIND == IPCB_EE) is_slow = sqn(cross(x[1] - x[0], x[3] - x[2])) < c.eps[i]; // mollifier active
        if (!is_slow) {
            const int np = sb.prim == 0 ? 2 : (sb.prim == 1 ? 3 : 4);
            const int pi[4] = { sb.i0, sb.i1, sb.i2, sb.i3 };
            d3 y[4];
            for (int k = 0; k < np; k++) y[k] = x[pi[k]];
            FastGeom g;
            fast_geometry(sb.prim, y, g);
            const double d2 = sub_value(sb, x); // the same value the energy / gradient use
            const double w = c.w[i];
            const double wf1 = w * B.df(d2), wf2 = w * B.ddf(d2);
            const int64_t base = block_offset + i * (NP * NP);
            // keys (and zero blocks for the stencil points the primitive does not involve)
            bool used[4] = { false, false, false, false };
            for (int k = 0; k < np; k++) used[pi[k]] = true;
            for (int bi = 0; bi < NP; bi++)
                for (int bj = 0; bj < NP; bj++) {
                    const int64_t e = base + bi * NP + bj;
                    hkey[e] = ((unsigned long long)(unsigned)vid[bj] << 32) | (unsigned)vid[bi];
                    if (!(used[bi] && used[bj])) {
                        for (int k = 0; k < 9; k++) hval[e * 9 + k] = 0.0;
                        hmask[e] = 0;
                    }
                }
            auto emit = [&](int a, int b, const double* blk) {
                const int64_t e = base + pi[a] * NP + pi[b];
                unsigned short mask = 0;
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    hval[e * 9 + k] = blk[k];
                    mask |= (blk[k] != 0.0) << k;
                }
                hmask[e] = mask;
            };
            if (sb.prim == 0) fast_projected_blocks<0>(g, np, wf1, wf2, psd_mode, emit);
            else if (sb.prim == 1) fast_projected_blocks<1>(g, np, wf1, wf2, psd_mode, emit);
            else fast_projected_blocks<2>(g, np, wf1, wf2, psd_mode, emit);
        }
    }
    const unsigned mball = __ballot_sync(0xffffffffu, is_slow);
    if (mball) {
        const int lane = threadIdx.x & 31;
        unsigned long long basep = 0;
        if (lane == __ffs(mball) - 1) basep = atomicAdd(slow_count, (unsigned long long)__popc(mball));
        basep = __shfl_sync(0xffffffffu, basep, __ffs(mball) - 1);
        if (is_slow) slow[basep + __popc(mball & ((1u << lane) - 1))] = int(i);
    }
}

__global__ void k_iota_h(int64_t n, int* __restrict__ idx)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) idx[i] = int(i);
}
__global__ void k_block_heads(int64_t n, const unsigned long long* __restrict__ key, int* __restrict__ head)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}
// sum every run of equal keys into one unique block (setFromTriplets sums duplicates)
__global__ void k_block_reduce(int64_t n, const unsigned long long* __restrict__ key, const int* __restrict__ idx,
                               const int* __restrict__ head, const int* __restrict__ upos, const double* __restrict__ hval,
                               const unsigned short* __restrict__ hmask, unsigned long long* __restrict__ ukey,
                               double* __restrict__ ublk, unsigned short* __restrict__ umask)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n || !head[i]) return;
    const unsigned long long k = key[i];
    double acc[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    unsigned short mask = 0;
    for (int64_t j = i; j < n && key[j] == k; j++) {
        const int e = idx[j];
        mask |= hmask[e];
#pragma unroll
        for (int q = 0; q < 9; q++) acc[q] += hval[(size_t)e * 9 + q];
    }
    const int u = upos[i];
    ukey[u] = k;
    umask[u] = mask;
#pragma unroll
    for (int q = 0; q < 9; q++) ublk[(size_t)u * 9 + q] = acc[q];
}
// colptr[v] = first unique block whose column vertex is >= v  (v in 0..nV)
__global__ void k_colptr(int nV, int nU, const unsigned long long* __restrict__ ukey, int* __restrict__ colptr)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v > nV) return;
    const unsigned long long target = (unsigned long long)(unsigned)v << 32;
    int lo = 0, hi = nU;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ukey[mid] < target) lo = mid + 1;
        else hi = mid;
    }
    colptr[v] = lo;
}
// per unique block and scalar column l: number of entries, stored l-major inside its block column
__global__ void k_block_counts(int nU, const unsigned long long* __restrict__ ukey, const unsigned short* __restrict__ umask,
                               const int* __restrict__ colptr, int* __restrict__ cnt)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nU) return;
    const int vj = int(ukey[u] >> 32);
    const int s = colptr[vj], nb = colptr[vj + 1] - s;
    const unsigned m = umask[u];
#pragma unroll
    for (int l = 0; l < 3; l++) cnt[3 * (size_t)s + l * nb + (u - s)] = __popc(m & (0x49u << l));
}
__global__ void k_fill_csc(int nU, const unsigned long long* __restrict__ ukey, const unsigned short* __restrict__ umask,
                           const double* __restrict__ ublk, const int* __restrict__ colptr, const int* __restrict__ scan,
                           int* __restrict__ inner, double* __restrict__ vals)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nU) return;
    const unsigned long long k = ukey[u];
    const int vj = int(k >> 32), vi = int(k & 0xffffffffu);
    const int s = colptr[vj], nb = colptr[vj + 1] - s;
    const unsigned m = umask[u];
#pragma unroll
    for (int l = 0; l < 3; l++) {
        int p = scan[3 * (size_t)s + l * nb + (u - s)];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            if (m & (1u << (3 * r + l))) {
                inner[p] = 3 * vi + r;
                vals[p] = ublk[(size_t)u * 9 + 3 * r + l];
                p++;
            }
        }
    }
}
__global__ void k_outer(int nV, const int* __restrict__ colptr, const int* __restrict__ scan, int* __restrict__ outer)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v > nV) return;
    if (v == nV) {
        outer[3 * (size_t)nV] = scan[3 * (size_t)colptr[nV]];
        return;
    }
    const int s = colptr[v], nb = colptr[v + 1] - s;
#pragma unroll
    for (int l = 0; l < 3; l++) outer[3 * (size_t)v + l] = scan[3 * (size_t)s + l * nb];
}
__global__ void k_zero_int(int64_t n, int* p)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

void barrier_hessian(ipcb_ctx* ctx, const ipcb_barrier_params& bp, int psd_mode)
{
    cudaStream_t s = ctx->stream;
    const BarrierDev B = make_barrier(bp, ctx->dmin);
    const int nV = ctx->nV;
    const int np2[4] = { 4, 9, 16, 16 };
    int64_t nblk = 0, offs[4];
    for (int k = 0; k < 4; k++) {
        offs[k] = nblk;
        nblk += ctx->coll[k].count * np2[k];
    }
    ctx->outer.reserve(3 * size_t(nV) + 1);
    ctx->nnz = 0;
    if (nblk == 0) { // empty ndof x ndof matrix (potential.cpp:107-109)
        k_zero_int<<<grid_for(3 * size_t(nV) + 1, 256), 256, 0, s>>>(3 * int64_t(nV) + 1, ctx->outer.p);
        ctx->launches++;
        return;
    }
    if (nblk > 0x7fffffffll) throw Error("Hessian has more than 2^31 local blocks; shard the collision set");
    {
        Stage st(ctx, "hessian_local");
        ctx->hkey.reserve(nblk), ctx->hkey_sorted.reserve(nblk), ctx->hidx.reserve(nblk), ctx->hidx_sorted.reserve(nblk);
        ctx->hval.reserve(9 * size_t(nblk)), ctx->hmask.reserve(nblk);
        const MeshView m = mesh_view(ctx);
        unsigned long long* hk = ctx->hkey.p;
        double* hv = ctx->hval.p;
        unsigned short* hm = ctx->hmask.p;
        const int64_t n0 = ctx->coll[0].count, n1 = ctx->coll[1].count, n2 = ctx->coll[2].count, n3 = ctx->coll[3].count;
        static const bool force_general = getenv("IPCB_HESSIAN_GENERAL") != nullptr; // A/B switch for tests and profiles
        if (psd_mode == IPCB_PSD_NONE || force_general) {
            if (n0) k_hessian_local<IPCB_VV><<<grid_for(n0, 128), 128, 0, s>>>(view(ctx, 0), m, B, psd_mode, offs[0], hk, hv, hm, nullptr, 0);
            if (n1) k_hessian_local<IPCB_EV><<<grid_for(n1, 128), 128, 0, s>>>(view(ctx, 1), m, B, psd_mode, offs[1], hk, hv, hm, nullptr, 0);
            if (n2) k_hessian_local<IPCB_EE><<<grid_for(n2, 128), 128, 0, s>>>(view(ctx, 2), m, B, psd_mode, offs[2], hk, hv, hm, nullptr, 0);
            if (n3) k_hessian_local<IPCB_FV><<<grid_for(n3, 128), 128, 0, s>>>(view(ctx, 3), m, B, psd_mode, offs[3], hk, hv, hm, nullptr, 0);
            ctx->launches += 4;
        } else {
            unsigned long long* slow_count = ctx->dCounters.p + 5;
            ctx->hhead.reserve(std::max<int64_t>(n2, 1)); // scratch: the slow list (not yet needed by the assembly)
            IPCB_CUDA(cudaMemsetAsync(slow_count, 0, sizeof(unsigned long long), s));
            if (n0) k_hessian_fast<IPCB_VV><<<grid_for(n0, 128), 128, 0, s>>>(view(ctx, 0), m, B, psd_mode, offs[0], hk, hv, hm, ctx->hhead.p, slow_count);
            if (n1) k_hessian_fast<IPCB_EV><<<grid_for(n1, 128), 128, 0, s>>>(view(ctx, 1), m, B, psd_mode, offs[1], hk, hv, hm, ctx->hhead.p, slow_count);
            if (n3) k_hessian_fast<IPCB_FV><<<grid_for(n3, 128), 128, 0, s>>>(view(ctx, 3), m, B, psd_mode, offs[3], hk, hv, hm, ctx->hhead.p, slow_count);
            if (n2) {
                k_hessian_fast<IPCB_EE><<<grid_for(n2, 128), 128, 0, s>>>(view(ctx, 2), m, B, psd_mode, offs[2], hk, hv, hm, ctx->hhead.p, slow_count);
                IPCB_CUDA(cudaMemcpyAsync(&ctx->pinned.p[11], slow_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
                IPCB_CUDA(cudaStreamSynchronize(s));
                const int64_t nslow = ctx->pinned.p[11];
                if (nslow) {
                    // the list lives in hhead, which the assembly overwrites later: process it now
                    k_hessian_local<IPCB_EE><<<grid_for(nslow, 128), 128, 0, s>>>(view(ctx, 2), m, B, psd_mode, offs[2], hk, hv, hm, ctx->hhead.p, nslow);
                    ctx->launches++;
                }
            }
            ctx->launches += 4;
        }
        IPCB_CUDA(cudaGetLastError());
    }
    Stage st(ctx, "hessian_assemble");
    int vbits = 0;
    while ((1ll << vbits) < std::max(nV, 2)) vbits++;
    k_iota_h<<<grid_for(nblk, 256), 256, 0, s>>>(nblk, ctx->hidx.p);
    ctx->hhead.reserve(nblk), ctx->hpos.reserve(nblk + 1);
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, ctx->hkey.p, ctx->hkey_sorted.p, ctx->hidx.p, ctx->hidx_sorted.p, nblk, 0, 32 + vbits, s);
    cub::DeviceScan::ExclusiveSum(nullptr, b2, ctx->hhead.p, ctx->hpos.p, nblk, s);
    ctx->cubtmp.reserve(std::max(b1, b2) + 1024);
    cub::DeviceRadixSort::SortPairs(ctx->cubtmp.p, b1, ctx->hkey.p, ctx->hkey_sorted.p, ctx->hidx.p, ctx->hidx_sorted.p, nblk, 0, 32 + vbits,
                                    s);
    k_block_heads<<<grid_for(nblk, 256), 256, 0, s>>>(nblk, ctx->hkey_sorted.p, ctx->hhead.p);
    cub::DeviceScan::ExclusiveSum(ctx->cubtmp.p, b2, ctx->hhead.p, ctx->hpos.p, nblk, s);
    IPCB_CUDA(cudaMemcpyAsync(&ctx->pinned.p[8], ctx->hpos.p + (nblk - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    IPCB_CUDA(cudaMemcpyAsync(&ctx->pinned.p[10], ctx->hhead.p + (nblk - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    IPCB_CUDA(cudaStreamSynchronize(s));
    // number of runs = heads before the last element + (1 if the last element starts a run)
    const int nU = *reinterpret_cast<int*>(&ctx->pinned.p[8]) + *reinterpret_cast<int*>(&ctx->pinned.p[10]);
    ctx->ukey.reserve(nU), ctx->ublk.reserve(9 * size_t(nU)), ctx->umask.reserve(nU);
    k_block_reduce<<<grid_for(nblk, 256), 256, 0, s>>>(nblk, ctx->hkey_sorted.p, ctx->hidx_sorted.p, ctx->hhead.p, ctx->hpos.p, ctx->hval.p,
                                                       ctx->hmask.p, ctx->ukey.p, ctx->ublk.p, ctx->umask.p);
    ctx->hcolptr.reserve(size_t(nV) + 2);
    k_colptr<<<grid_for(size_t(nV) + 1, 256), 256, 0, s>>>(nV, nU, ctx->ukey.p, ctx->hcolptr.p);
    const size_t ncnt = 3 * size_t(nU) + 1;
    ctx->hcnt.reserve(ncnt), ctx->hscan.reserve(ncnt);
    k_zero_int<<<grid_for(ncnt, 256), 256, 0, s>>>(int64_t(ncnt), ctx->hcnt.p);
    k_block_counts<<<grid_for(nU, 256), 256, 0, s>>>(nU, ctx->ukey.p, ctx->umask.p, ctx->hcolptr.p, ctx->hcnt.p);
    size_t b3 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b3, ctx->hcnt.p, ctx->hscan.p, int(ncnt), s);
    ctx->cubtmp.reserve(b3);
    cub::DeviceScan::ExclusiveSum(ctx->cubtmp.p, b3, ctx->hcnt.p, ctx->hscan.p, int(ncnt), s);
    IPCB_CUDA(cudaMemcpyAsync(&ctx->pinned.p[9], ctx->hscan.p + (ncnt - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    IPCB_CUDA(cudaStreamSynchronize(s));
    ctx->nnz = *reinterpret_cast<int*>(&ctx->pinned.p[9]);
    ctx->inner.reserve(ctx->nnz), ctx->vals.reserve(ctx->nnz);
    k_fill_csc<<<grid_for(nU, 256), 256, 0, s>>>(nU, ctx->ukey.p, ctx->umask.p, ctx->ublk.p, ctx->hcolptr.p, ctx->hscan.p, ctx->inner.p,
                                                 ctx->vals.p);
    k_outer<<<grid_for(size_t(nV) + 1, 256), 256, 0, s>>>(nV, ctx->hcolptr.p, ctx->hscan.p, ctx->outer.p);
    ctx->launches += 9 + 12;
    IPCB_CUDA(cudaGetLastError());
}

} // namespace ipcb
