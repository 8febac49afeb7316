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
#include "hessian_assembly.cuh"
#include "hessian_fast.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <memory>

namespace ipcb {

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

void barrier_energy(ipcb_ctx* ctx, const ipcb_barrier_params& bp, double* d_out)
{
    Stage st(ctx, "energy");
    cudaStream_t s = ctx->stream;
    const BarrierDev B = make_barrier(bp, ctx->dmin);
    size_t nblocks = 0;
    for (int k = 0; k < 4; k++) nblocks += grid_for(view_slice(ctx, k).n, EBLOCK);
    ctx->dScalar.reserve(nblocks + 16);
    size_t off = 0;
    for (int k = 0; k < 4; k++) {
        const CollView c = view_slice(ctx, k);
        const unsigned g = grid_for(c.n, EBLOCK);
        if (!g) continue;
        k_energy<<<g, EBLOCK, 0, s>>>(c, mesh_view(ctx), B, ctx->dScalar.p + off);
        off += g;
        ctx->launches++;
    }
    k_sum_partials<<<1, 1024, 0, s>>>(int(nblocks), ctx->dScalar.p, d_out);
    ctx->launches++;
    IPCB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// gradient (potential.cpp:58-94, normal_potential.cpp:137-170, local_to_global.hpp:21-45)
// One thread per collision, registers only: the distance gradient comes from the gradient-only closed
// forms (geom.cuh prim_grad), the stencil is resolved at compile time for VV / EV / FV and through
// the distance type for EE.
template <int KIND> __global__ void __launch_bounds__(256) k_gradient(CollView c, MeshView m, BarrierDev B, double* __restrict__ grad)
{
    constexpr int NP = KIND == IPCB_VV ? 2 : (KIND == IPCB_EV ? 3 : 4);
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    int vid[4];
    d3 x[4];
    load_stencil(KIND, c.ids[i], m, vid, x);
    const double w = c.w[i];
    d3 G[4];
    if (KIND != IPCB_EE) {
        // VV: PP(0,1); EV: PL(0,1,2); FV: plane(0,1,2,3) — argument order == stencil order
        const int prim = KIND == IPCB_VV ? 0 : (KIND == IPCB_EV ? 1 : 2);
        const double d = sub_value(collision_sub(KIND, 0), x);
        prim_grad(prim, x, G);
        const double sc = w * B.df(d);
#pragma unroll
        for (int a = 0; a < NP; a++) G[a] = sc * G[a];
    } else {
        const double eps = c.eps[i];
        const double s = sqn(cross(x[1] - x[0], x[3] - x[2]));
        const double mol = moll(s, eps);
        if (mol <= 0) return; // gradient is exactly zero (normal_potential.cpp:143-147)
        const Sub sb = sub_edge_edge(c.dt[i]);
        const double d = sub_value(sb, x);
        const int pi[4] = { sb.i0, sb.i1, sb.i2, sb.i3 };
        const int np = sb.prim == 0 ? 2 : (sb.prim == 1 ? 3 : 4);
        d3 y[4], gp[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int p = pi[k];
            y[k] = p == 0 ? x[0] : (p == 1 ? x[1] : (p == 2 ? x[2] : x[3]));
        }
        prim_grad(sb.prim, y, gp);
        const double sc = w * mol * B.df(d);
#pragma unroll
        for (int a = 0; a < 4; a++) {
            d3 acc = { 0, 0, 0 };
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (k < np && pi[k] == a) acc = acc + gp[k];
            G[a] = sc * acc;
        }
        if (s < eps) { // + (w f) m'(s) grad s
            const d3 u = x[1] - x[0], v = x[3] - x[2];
            d3 gu, gv;
            cross_sq_grad(u, v, cross(u, v), gu, gv);
            const double wf = w * B.f(d), dm = moll_d(s, eps);
            const d3 mu = wf * (dm * gu), mv = wf * (dm * gv);
            G[0] = G[0] - mu, G[1] = G[1] + mu, G[2] = G[2] - mv, G[3] = G[3] + mv;
        }
    }
#pragma unroll
    for (int a = 0; a < NP; a++) {
        double* g = grad + 3 * (size_t)vid[a];
        if (G[a].x != 0.0) atomicAdd(g, G[a].x);
        if (G[a].y != 0.0) atomicAdd(g + 1, G[a].y);
        if (G[a].z != 0.0) atomicAdd(g + 2, G[a].z);
    }
}

void barrier_gradient(ipcb_ctx* ctx, const ipcb_barrier_params& bp, double* d_grad)
{
    Stage st(ctx, "gradient");
    cudaStream_t s = ctx->stream;
    const BarrierDev B = make_barrier(bp, ctx->dmin);
    IPCB_CUDA(cudaMemsetAsync(d_grad, 0, sizeof(double) * 3 * size_t(ctx->nV), s));
    const MeshView m = mesh_view(ctx);
    const CollView c0 = view_slice(ctx, 0), c1 = view_slice(ctx, 1), c2 = view_slice(ctx, 2), c3 = view_slice(ctx, 3);
    if (c0.n) k_gradient<IPCB_VV><<<grid_for(c0.n, 256), 256, 0, s>>>(c0, m, B, d_grad), ctx->launches++;
    if (c1.n) k_gradient<IPCB_EV><<<grid_for(c1.n, 256), 256, 0, s>>>(c1, m, B, d_grad), ctx->launches++;
    if (c2.n) k_gradient<IPCB_EE><<<grid_for(c2.n, 256), 256, 0, s>>>(c2, m, B, d_grad), ctx->launches++;
    if (c3.n) k_gradient<IPCB_FV><<<grid_for(c3.n, 256), 256, 0, s>>>(c3, m, B, d_grad), ctx->launches++;
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
// Output of the local kernels: one record per collision, addressed by the GLOBAL collision index
// gi (VV first, then EV, EE, FV).  Block slot (a, b) = a * 4 + b holds the 3x3 block whose COLUMNS
// belong to stencil point a and whose ROWS belong to point b (row-major), i.e. what the column of
// vertex vid[a] needs in compressed-column order.
template <int KIND>
__global__ void __launch_bounds__(128)
    k_hessian_local(CollView c, MeshView m, BarrierDev B, int psd_mode, int64_t gi0, int64_t inc0, HessOut out, const int* __restrict__ list,
                    int64_t nlist, const int* __restrict__ sel, int64_t nsel)
{
    constexpr int NP = KIND == IPCB_VV ? 2 : (KIND == IPCB_EV ? 3 : 4);
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (tid >= (list ? nlist : nsel)) return;
    const int64_t t = list ? list[tid] : tid; // record index within the kind
    const int64_t i = sel ? sel[t] : t;       // collision index within the kind
    int vid[4];
    d3 x[4];
    load_stencil(KIND, c.ids[i], m, vid, x);
    // listed collisions already have their record; a collision without an owned vertex contributes nothing
    if (!list && write_record<NP>(out, gi0 + t, inc0 + t * NP, vid) == 0) return;
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
    // emit the upper-triangular vertex blocks: slot (column point bj, row point bi), bj <= bi
    const int64_t gi = gi0 + t;
    unsigned short masks[HSLOTS];
#pragma unroll
    for (int k = 0; k < HSLOTS; k++) masks[k] = 0;
    double* rec = out.blk + size_t(t) * (tri_count(NP) * 9);
    for (int bj = 0; bj < NP; bj++)
        for (int bi = bj; bi < NP; bi++) {
            unsigned mask = 0;
            for (int k = 0; k < 9; k++) {
                const double v = D.H[(bi * 4 + bj) * 9 + k];
                rec[tri_slot(NP, bj, bi) * 9 + k] = v;
                mask |= unsigned(v != 0.0) << k; // exact zeros are not entries (local_to_global.hpp:290-291)
            }
            masks[bj * 4 + bi] = (unsigned short)mask;
            // a diagonal block keeps its OWN mask: its exact-zero pattern need not be symmetric at rounding level, and
            // the numeric pass decides from the values it reads — mask and values must describe the same block
            if (bi != bj) masks[bi * 4 + bj] = (unsigned short)mask_transpose(mask);
        }
    for (int k = 0; k < HSLOTS; k++) out.mask[gi * HSLOTS + k] = masks[k];
}

// PSD-projected local Hessians through the analytic 3+p dimensional subspace (hessian_fast.cuh).
// Edge-edge collisions whose mollifier is active at X (or whose distance type is not edge-edge) are
// appended to `slow` for the general kernel.
//
// Stores: a thread owns one collision = 216 / 432 / 720 contiguous bytes of (upper-triangular) blocks.  Writing them from
// registers would make every 8-byte store of a warp hit 32 different sectors.  !BULK (vertex-vertex records, and the A/B
// alternative for the others): each row of the triangle is staged in shared memory (odd stride: conflict-free) and written
// out by the whole warp with consecutive lanes on consecutive addresses.
//
// BULK (3- and 4-point kinds): the record of a collision is 432 / 720 contiguous, 16-byte aligned bytes, so the thread stages it in
// its own shared-memory slot and hands it to the TMA unit (cp.async.bulk.global.shared::cta): one instruction per 288- or 432-byte
// piece instead of a warp-wide copy loop (which was 30 % of the kernel's instructions); the store drains while the thread computes
// the next piece.  A 4-point record goes in two pieces — row 0 (4 blocks) and rows 1..3 (6 blocks) — so that 432 bytes per thread
// (54 KB per block, four blocks per SM) are enough.
constexpr int BULK_STRIDE = 54; // doubles per thread
template <int KIND, int MINB, bool BULK>
__global__ void __launch_bounds__(128, MINB)
    k_hessian_fast(CollView c, MeshView m, BarrierDev B, int psd_mode, int64_t gi0, int64_t inc0, HessOut out, int* __restrict__ slow,
                   unsigned long long* slow_count, const int* __restrict__ sel, int64_t nsel)
{
    constexpr int NP = KIND == IPCB_VV ? 2 : (KIND == IPCB_EV ? 3 : 4);
    constexpr int P = KIND == IPCB_VV ? 0 : (KIND == IPCB_EV ? 1 : 2);
    constexpr int PRIM = KIND == IPCB_VV ? 0 : (KIND == IPCB_EV ? 1 : (KIND == IPCB_FV ? 2 : 3));
    static_assert(!BULK || NP >= 3, "a vertex-vertex record (216 bytes) is not a multiple of 16 bytes");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // t = record index within the kind; i = collision index (sel: the collisions touching the rank's row block)
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool valid = t < nsel;
    const int64_t i = valid && sel ? sel[t] : t;
    bool is_slow = false;
    unsigned own = 0; // stencil points whose vertex (column) this rank owns
    FastProj<P> pr;
    if (valid) {
        int vid[4];
        d3 x[4];
        load_stencil(KIND, c.ids[i], m, vid, x);
        own = write_record<NP>(out, gi0 + t, inc0 + t * NP, vid);
        if (KIND == IPCB_EE) is_slow = own != 0 && (c.dt[i] != EE_AB || sqn(cross(x[1] - x[0], x[3] - x[2])) < c.eps[i]);
        if (!is_slow && own != 0) {
            FastGeom g;
            fast_geometry(PRIM, x, g);
            const double d2 = sub_value(collision_sub(KIND, EE_AB), x); // the same value the energy / gradient use
            const double w = c.w[i];
            fast_project<P>(g, w * B.df(d2), w * B.ddf(d2), psd_mode, pr);
        }
    }
    const bool emit = valid && !is_slow && own != 0;
    const unsigned emask_any = __ballot_sync(0xffffffffu, emit);
    const unsigned smask = __ballot_sync(0xffffffffu, is_slow);
    if (smask) {
        unsigned long long basep = 0;
        if (lane == __ffs(smask) - 1) basep = atomicAdd(slow_count, (unsigned long long)__popc(smask));
        basep = __shfl_sync(0xffffffffu, basep, __ffs(smask) - 1);
        if (is_slow) slow[basep + __popc(smask & ((1u << lane) - 1))] = int(t);
    }
    if (emask_any == 0) return;
    unsigned mpack[HSLOTS / 2];
#pragma unroll
    for (int k = 0; k < HSLOTS / 2; k++) mpack[k] = 0;
    if constexpr (BULK) {
        extern __shared__ __align__(128) double bulk_stage[];
        double* my = bulk_stage + threadIdx.x * BULK_STRIDE;
        const unsigned my_s = unsigned(__cvta_generic_to_shared(my));
        double* rec = out.blk + t * (tri_count(NP) * 9);
        constexpr int NPH = NP == 4 ? 2 : 1;
#pragma unroll
        for (int ph = 0; ph < NPH; ph++) {
            const int a_lo = NP == 4 ? ph : 0, a_hi = NP == 4 ? (ph == 0 ? 1 : 4) : NP;
            // rows a_lo.. are needed by an owned column >= a_lo (directly, or transposed)
            const bool need = emit && (own >> a_lo) != 0u;
            if (ph > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // the slot has been read
            if (need) {
                int off = 0;
#pragma unroll
                for (int a = a_lo; a < a_hi; a++)
#pragma unroll
                    for (int b = a; b < NP; b++) {
                        double blk[9];
                        fast_block<P>(pr, b, a, blk);
                        unsigned mask = 0;
#pragma unroll
                        for (int k = 0; k < 9; k++) {
                            my[off + k] = blk[k];
                            mask |= unsigned(blk[k] != 0.0) << k;
                        }
                        off += 9;
                        const int slot = a * 4 + b, mirror = b * 4 + a;
                        mpack[slot >> 1] |= mask << (16 * (slot & 1));
                        if (b != a) mpack[mirror >> 1] |= mask_transpose(mask) << (16 * (mirror & 1));
                    }
                asm volatile("fence.proxy.async;" ::: "memory"); // the staged values become visible to the async proxy
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(rec + tri_slot(NP, a_lo, a_lo) * 9), "r"(my_s),
                             "r"(off * 8)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (emit) {
            uint4* mp = reinterpret_cast<uint4*>(out.mask + (gi0 + t) * HSLOTS);
            mp[0] = make_uint4(mpack[0], mpack[1], mpack[2], mpack[3]);
            mp[1] = make_uint4(mpack[4], mpack[5], mpack[6], mpack[7]);
        }
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // shared memory must outlive the copy
    } else {
        constexpr int ROW = NP * 9, PAD = ROW | 1;
        __shared__ double stage[4][32][PAD];
        const int64_t tw = t - lane; // record index (within the kind) of lane 0's collision
#pragma unroll
        for (int a = 0; a < NP; a++) { // column point: its row of the upper triangle, blocks (a, b >= a), is contiguous
            const bool emit_a = emit && (own >> a) != 0u; // needed by an owned column a, or by an owned column b > a (transposed)
            const unsigned emask = __ballot_sync(0xffffffffu, emit_a);
            if (emask == 0) continue;
            const int row = (NP - a) * 9; // doubles in this row
            if (emit_a) {
#pragma unroll
                for (int b = a; b < NP; b++) { // row point
                    double blk[9];
                    fast_block<P>(pr, b, a, blk);
                    unsigned mask = 0;
#pragma unroll
                    for (int k = 0; k < 9; k++) {
                        stage[warp][lane][(b - a) * 9 + k] = blk[k];
                        mask |= unsigned(blk[k] != 0.0) << k;
                    }
                    const int slot = a * 4 + b, mirror = b * 4 + a;
                    mpack[slot >> 1] |= mask << (16 * (slot & 1));
                    if (b != a) mpack[mirror >> 1] |= mask_transpose(mask) << (16 * (mirror & 1));
                }
            }
            __syncwarp();
            {
                int cl = 0, j = lane; // flat index lane + 32 * it over the warp's 32 x row staged doubles, without a division
                while (j >= row) j -= row, cl++;
                double* dst = out.blk + tw * (tri_count(NP) * 9) + tri_slot(NP, a, a) * 9;
#pragma unroll 4
                for (int it = 0; it < row; it++) {
                    if ((emask >> cl) & 1u) dst[size_t(cl) * (tri_count(NP) * 9) + j] = stage[warp][cl][j];
                    j += 32;
                    while (j >= row) j -= row, cl++;
                }
            }
            __syncwarp();
        }
        if (emit) {
            uint4* mp = reinterpret_cast<uint4*>(out.mask + (gi0 + t) * HSLOTS);
            mp[0] = make_uint4(mpack[0], mpack[1], mpack[2], mpack[3]);
            mp[1] = make_uint4(mpack[4], mpack[5], mpack[6], mpack[7]);
        }
    }
}

// ---------------------------------------------------------------------------
// Assembly into compressed columns (== compressed rows of the symmetric matrix), replacing
// local_hessian_to_global_triplets + setFromTriplets (local_to_global.hpp:263-305, potential.cpp:218).
//
// The reference sorts 144 scalar triplets per collision globally.  Here only the (vertex, collision)
// INCIDENCES are sorted globally (4 per collision, radix sort on the vertex bits); everything else
// happens per column vertex inside one warp: the column's row blocks are sorted by row vertex in
// shared memory (bitonic), runs of equal row vertex become one unique block whose pattern is the
// OR of the exact-non-zero masks (pass 1: counts), and after one prefix sum over the scalar
// columns the blocks are gathered (72 contiguous bytes each), run-summed with a segmented warp
// scan and written straight into inner / values (pass 2).  Summation order is fixed by the sort.
constexpr int WARP_CAP = 512;  // items a warp sorts in shared memory
constexpr int CTA_CAP = 8192;  // items a block sorts in shared memory; beyond: global scratch

__device__ inline int lower_bound_hi(const unsigned long long* __restrict__ key, int lo, int hi, unsigned v)
{
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (unsigned(key[mid] >> 32) < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
__device__ inline int lower_bound_lo(const unsigned long long* __restrict__ key, int lo, int hi, unsigned v)
{
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (unsigned(key[mid]) < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
// per column vertex: first incidence and number of row-block items (2 / 3 / 4 per VV / EV / other incidence)
// colb[v] = (b1, b2): first edge-vertex incidence and first 4-point incidence of the column (incidences are ordered
// VV, EV, then the 4-point kinds) — computed once here instead of by every lane of the column's warp
__global__ void k_col_ranges(int nV, int nInc, const unsigned long long* __restrict__ inc, unsigned ref_ev, unsigned ref_ee,
                             int* __restrict__ colinc, int* __restrict__ colR, int2* __restrict__ colb)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v > nV) return;
    const int s = lower_bound_hi(inc, 0, nInc, unsigned(v));
    colinc[v] = s;
    int R = 0;
    if (v < nV) {
        const int e = lower_bound_hi(inc, s, nInc, unsigned(v) + 1u);
        const int b1 = lower_bound_lo(inc, s, e, ref_ev), b2 = lower_bound_lo(inc, b1, e, ref_ee);
        R = 2 * (b1 - s) + 3 * (b2 - b1) + 4 * (e - b2);
        colb[v] = make_int2(b1, b2);
    }
    colR[v] = R;
}

template <int NT> __device__ inline void group_sync()
{
    if (NT == 32) __syncwarp();
    else __syncthreads();
}
// bitonic sort of n (power of two) keys by NT cooperating threads
template <int NT> __device__ inline void bitonic_sort(unsigned long long* keys, int n, int t)
{
    for (int k = 2; k <= n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int q = t; q < (n >> 1); q += NT) {
                const int i = ((q & ~(j - 1)) << 1) | (q & (j - 1)); // index with bit j clear
                const int p = i | j;
                const unsigned long long a = keys[i], b = keys[p];
                const bool up = (i & k) == 0;
                if ((a > b) == up) keys[i] = b, keys[p] = a;
            }
            group_sync<NT>();
        }
}

struct SymArgs {
    int nV;
    const unsigned long long* inc;
    const int* colinc;
    const int* colR;
    const int* itemoff;
    const int2* colb;
    const int4* vid;
    const unsigned short* mask;
    unsigned ref_ev, ref_ee;
    unsigned blk_base[3]; // first block of the VV / EV / 4-point records (in blocks of 9 doubles)
    unsigned* sref; // per item (grouped by unique block): (block index << 1) | read-transposed
    int2* udesc;    // per unique block of a column, at itemoff[v] + u: (first item within the column, row vertex)
    int* colU;      // unique blocks per column
    int* cnt;       // entries per scalar column
};

// one incidence = one collision seen from one of its stencil points: its row blocks
struct IncItems {
    unsigned gi, a;
    int np;
    int vi[4];
    unsigned short mk[4];
};
__device__ inline IncItems load_incidence(const SymArgs& A, int q, int b1, int b2)
{
    IncItems it;
    const unsigned ref = unsigned(A.inc[q]); // gi * 4 + a
    it.gi = ref >> 2, it.a = ref & 3u;
    it.np = q < b1 ? 2 : (q < b2 ? 3 : 4); // incidences are ordered VV, EV, then the 4-point kinds
    const int4 vv = A.vid[it.gi];
    const uint2 mm = *reinterpret_cast<const uint2*>(A.mask + size_t(it.gi) * HSLOTS + it.a * 4);
    it.vi[0] = vv.x, it.vi[1] = vv.y, it.vi[2] = vv.z, it.vi[3] = vv.w;
    it.mk[0] = (unsigned short)(mm.x & 0xffffu), it.mk[1] = (unsigned short)(mm.x >> 16);
    it.mk[2] = (unsigned short)(mm.y & 0xffffu), it.mk[3] = (unsigned short)(mm.y >> 16);
    return it;
}

// where the column of point it.a finds its row block b: the stored upper-triangular slot (min, max), transposed if a > b
__device__ __forceinline__ unsigned block_ref(const SymArgs& A, const IncItems& it, int b)
{
    const int a = int(it.a), lo = min(a, b), hi = max(a, b);
    unsigned idx;
    if (it.np == 2) idx = A.blk_base[0] + it.gi * 3u + unsigned(tri_slot(2, lo, hi));
    else if (it.np == 3) idx = A.blk_base[1] + (it.gi - (A.ref_ev >> 2)) * 6u + unsigned(tri_slot(3, lo, hi));
    else idx = A.blk_base[2] + (it.gi - (A.ref_ee >> 2)) * 10u + unsigned(tri_slot(4, lo, hi));
    return (idx << 1) | unsigned(a > b);
}

// ---- pass 1, general path: sort all row blocks of the column (NT cooperating threads) ---------------------------------
// keys / refs / masks: room for npow2 / R / R entries; scan: NT ints (shared)
template <int NT>
__device__ inline void column_symbolic_sort(const SymArgs& A, int v, int t, unsigned long long* keys, unsigned* refs, unsigned short* masks,
                                            int* scan)
{
    const int s = A.colinc[v], e = A.colinc[v + 1], R = A.colR[v];
    int npow2 = 1;
    while (npow2 < R) npow2 <<= 1;
    const int2 bb = A.colb[v];
    const int b1 = bb.x, b2 = bb.y;
    for (int q = s + t; q < e; q += NT) {
        const IncItems it = load_incidence(A, q, b1, b2);
        const int slot0 = q < b1 ? 2 * (q - s) : (q < b2 ? 2 * (b1 - s) + 3 * (q - b1) : 2 * (b1 - s) + 3 * (b2 - b1) + 4 * (q - b2));
#pragma unroll
        for (int b = 0; b < 4; b++)
            if (b < it.np) {
                const int slot = slot0 + b;
                keys[slot] = ((unsigned long long)(unsigned)it.vi[b] << 32) | unsigned(slot);
                refs[slot] = block_ref(A, it, b);
                masks[slot] = it.mk[b];
            }
    }
    for (int q = R + t; q < npow2; q += NT) keys[q] = ~0ull;
    group_sync<NT>();
    bitonic_sort<NT>(keys, npow2, t);
    // unique blocks: thread t owns the contiguous items [lo, hi)
    const int L = (R + NT - 1) / NT, lo = min(R, t * L), hi = min(R, lo + L);
    int heads = 0;
    for (int q = lo; q < hi; q++) heads += q == 0 || unsigned(keys[q] >> 32) != unsigned(keys[q - 1] >> 32);
    scan[t] = heads;
    group_sync<NT>();
    int u = 0, total = 0;
    for (int k = 0; k < NT; k++) {
        const int c = scan[k];
        u += k < t ? c : 0;
        total += c;
    }
    const int ioff = A.itemoff[v];
    int c0 = 0, c1 = 0, c2 = 0;
    for (int q = lo; q < hi; q++) {
        const unsigned long long k = keys[q];
        const unsigned row = unsigned(k >> 32);
        A.sref[ioff + q] = refs[unsigned(k)];
        if (q == 0 || unsigned(keys[q - 1] >> 32) != row) {
            unsigned mk = 0;
            for (int j = q; j < R && unsigned(keys[j] >> 32) == row; j++) mk |= masks[unsigned(keys[j])];
            c0 += __popc(mk & 0x49u), c1 += __popc(mk & 0x92u), c2 += __popc(mk & 0x124u);
            A.udesc[ioff + u] = make_int2(q, int(row));
            u++;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c0 += __shfl_xor_sync(0xffffffffu, c0, o);
        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    }
    if (NT == 32) {
        if (t == 0) A.cnt[3 * size_t(v)] = c0, A.cnt[3 * size_t(v) + 1] = c1, A.cnt[3 * size_t(v) + 2] = c2, A.colU[v] = total;
    } else {
        group_sync<NT>(); // scan[] is reused for the block reduction
        if ((t & 31) == 0) scan[3 * (t >> 5)] = c0, scan[3 * (t >> 5) + 1] = c1, scan[3 * (t >> 5) + 2] = c2;
        group_sync<NT>();
        if (t < 3) {
            int sum = 0;
            for (int w = 0; w < NT / 32; w++) sum += scan[3 * w + t];
            A.cnt[3 * size_t(v) + t] = sum;
        }
        if (t == 0) A.colU[v] = total;
        group_sync<NT>();
    }
}

// ---- pass 1, fast path (one warp): de-duplicate the row vertices in a shared-memory hash table, sort only the
// unique ones, then place the items with a stable counting sort (deterministic order inside every run) ----------------
constexpr int HT = 128; // hash slots per warp
struct HashSmem {
    int key[HT];
    unsigned msk[HT];
    int cnt[HT];
    int base[HT];
    unsigned long long ukey[HT];
};
__device__ inline int hash_slot(int vi) { return int((unsigned(vi) * 2654435761u) >> 25); } // 7 bits

// returns false when the column has more unique row vertices than the table holds
__device__ inline bool column_symbolic_hash(const SymArgs& A, int v, int lane, HashSmem& H)
{
    const int s = A.colinc[v], e = A.colinc[v + 1], R = A.colR[v];
    const int2 bb = A.colb[v];
    const int b1 = bb.x, b2 = bb.y;
    for (int k = lane; k < HT; k += 32) H.key[k] = -1, H.msk[k] = 0, H.cnt[k] = 0;
    __syncwarp();
    // A. insert every row vertex; OR the patterns, count the items.  The first two incidences of every lane stay in
    // registers for the placement pass (a column has ~46 incidences on the dense scenes: no second gather)
    bool overflow = false;
    IncItems keep0, keep1;
    keep0.np = keep1.np = 0;
    for (int q0 = s; q0 < e; q0 += 32) {
        const int q = q0 + lane;
        if (q < e) {
            const IncItems it = load_incidence(A, q, b1, b2);
            if (q < s + 32) keep0 = it;
            else if (q < s + 64) keep1 = it;
#pragma unroll
            for (int b = 0; b < 4; b++)
                if (b < it.np) {
                    int h = hash_slot(it.vi[b]), probe = 0;
                    for (; probe < HT; probe++) {
                        const int old = atomicCAS(&H.key[h], -1, it.vi[b]);
                        if (old == -1 || old == it.vi[b]) break;
                        h = (h + 1) & (HT - 1);
                    }
                    if (probe == HT) overflow = true;
                    else atomicOr(&H.msk[h], unsigned(it.mk[b])), atomicAdd(&H.cnt[h], 1);
                }
        }
        // a full table makes every further item probe all of it: leave at once (a 9 345-item column spent 1 ms here)
        if (__any_sync(0xffffffffu, overflow)) return false;
    }
    __syncwarp();
    // B. unique row vertices, sorted
    int U = 0;
    for (int k0 = 0; k0 < HT; k0 += 32) {
        const int k = k0 + lane;
        const bool occ = H.key[k] != -1;
        const unsigned m = __ballot_sync(0xffffffffu, occ);
        if (occ) H.ukey[U + __popc(m & ((1u << lane) - 1))] = ((unsigned long long)(unsigned)H.key[k] << 32) | unsigned(k);
        U += __popc(m);
    }
    int npow2 = 1;
    while (npow2 < U) npow2 <<= 1;
    for (int k = U + lane; k < npow2; k += 32) H.ukey[k] = ~0ull;
    __syncwarp();
    bitonic_sort<32>(H.ukey, npow2, lane);
    // C. run starts, descriptors, pattern counts
    const int ioff = A.itemoff[v];
    int c0 = 0, c1 = 0, c2 = 0, run = 0;
    for (int u0 = 0; u0 < U; u0 += 32) {
        const int u = u0 + lane;
        int slot = 0, c = 0;
        if (u < U) slot = int(unsigned(H.ukey[u])), c = H.cnt[slot];
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (u < U) {
            const int start = run + incl - c;
            H.base[slot] = start;
            A.udesc[ioff + u] = make_int2(start, int(unsigned(H.ukey[u] >> 32)));
            const unsigned mk = H.msk[slot];
            c0 += __popc(mk & 0x49u), c1 += __popc(mk & 0x92u), c2 += __popc(mk & 0x124u);
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c0 += __shfl_xor_sync(0xffffffffu, c0, o);
        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    }
    if (lane == 0) A.cnt[3 * size_t(v)] = c0, A.cnt[3 * size_t(v) + 1] = c1, A.cnt[3 * size_t(v) + 2] = c2, A.colU[v] = U;
    __syncwarp();
    // D. stable placement: chunks of 32 incidences in order, point by point; items of a chunk that fall into the
    // same run are ranked by lane
    for (int q0 = s; q0 < e; q0 += 32) {
        const int q = q0 + lane;
        IncItems it;
        it.np = 0;
        if (q < e) it = q0 == s ? keep0 : (q0 == s + 32 ? keep1 : load_incidence(A, q, b1, b2));
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const bool valid = b < it.np;
            const unsigned vm = __ballot_sync(0xffffffffu, valid);
            if (valid) {
                int h = hash_slot(it.vi[b]);
                while (H.key[h] != it.vi[b]) h = (h + 1) & (HT - 1);
                const unsigned peers = __match_any_sync(vm, h);
                const int rank = __popc(peers & ((1u << lane) - 1));
                const int pos = H.base[h];
                __syncwarp(vm);
                if (rank == 0) H.base[h] = pos + __popc(peers);
                A.sref[ioff + pos + rank] = block_ref(A, it, b);
            }
            __syncwarp();
        }
    }
    (void)R;
    return true;
}

constexpr int SYM_WARPS = 4;
union WarpSmem {
    HashSmem h;
    struct {
        unsigned long long keys[WARP_CAP];
        unsigned refs[WARP_CAP];
        unsigned short masks[WARP_CAP];
        int scan[32];
    } srt;
};
// order (optional): the columns are visited in the MORTON order of their vertices (the broad phase's vertex order)
// instead of by vertex id.  A stored upper-triangular block is read by the columns of BOTH its vertices; vertices in
// contact are neighbours in space, not in id (the next cloth layer is 63 K ids away), so a spatial visiting order lets the
// second read (and the shared id / mask records) hit L2.  The output position of a column does not depend on the order.
__global__ void __launch_bounds__(32 * SYM_WARPS)
    k_hess_symbolic(SymArgs A, int warp_cap, int use_hash, int* __restrict__ big, unsigned long long* nbig, const int* __restrict__ active,
                    const int* __restrict__ nactive)
{
    __shared__ WarpSmem sm[SYM_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = blockIdx.x * SYM_WARPS + warp;
    if (w >= *nactive) return; // the counts of every other column were zeroed by the caller
    const int v = active[w];
    const int R = A.colR[v];
    if (R == 0) {
        if (lane < 3) A.cnt[3 * size_t(v) + lane] = 0;
        if (lane == 0) A.colU[v] = 0;
        return;
    }
    // columns far beyond what 128 hash slots can de-duplicate go straight to the block kernel (its hash has 4 096 slots)
    if (use_hash && R <= 8 * WARP_CAP && column_symbolic_hash(A, v, lane, sm[warp].h)) return;
    __syncwarp();
    if (R > warp_cap) { // handed to the block-per-column kernel
        if (lane == 0) big[atomicAdd(nbig, 1ull)] = v;
        return;
    }
    column_symbolic_sort<32>(A, v, lane, sm[warp].srt.keys, sm[warp].srt.refs, sm[warp].srt.masks, sm[warp].srt.scan);
}

constexpr int BIG_THREADS = 256;
constexpr size_t BIG_SMEM = size_t(CTA_CAP) * (8 + 4 + 2);
// ---- pass 1, block-wide hash path: the same scheme as column_symbolic_hash for a column with up to HB unique row
// vertices and ANY number of items (a high-valence vertex inside a contact region: BASELINE config 2's sphere pole owns
// 9 345 row blocks on 268 unique rows — sorting those items in global scratch took 2 ms, more than the rest of the
// scene's symbolic pass).  Only the unique rows are sorted; items are placed chunk by chunk, warp after warp, so the
// order inside every run is the incidence order (deterministic).
constexpr int HB = 4096; // hash slots per block: 24 B each = 96 KB of the kernel's dynamic shared memory
static_assert(size_t(HB) * 24 <= BIG_SMEM, "the block-wide hash table lives in the sort buffers");
__device__ inline int hash_slot_big(int vi) { return int((unsigned(vi) * 2654435761u) >> 20); } // 12 bits
__device__ inline bool column_symbolic_hash_cta(const SymArgs& A, int v, int t, char* smem, int* scan)
{
    unsigned long long* ukey = reinterpret_cast<unsigned long long*>(smem);
    int* key = reinterpret_cast<int*>(smem + size_t(HB) * 8);
    unsigned* msk = reinterpret_cast<unsigned*>(smem + size_t(HB) * 12);
    int* cnt = reinterpret_cast<int*>(smem + size_t(HB) * 16);
    int* base = reinterpret_cast<int*>(smem + size_t(HB) * 20);
    __shared__ int nU;
    const int s = A.colinc[v], e = A.colinc[v + 1];
    const int2 bb = A.colb[v];
    const int b1 = bb.x, b2 = bb.y;
    for (int k = t; k < HB; k += BIG_THREADS) key[k] = -1, msk[k] = 0, cnt[k] = 0;
    if (t == 0) nU = 0;
    __syncthreads();
    bool overflow = false;
    for (int q = s + t; q < e; q += BIG_THREADS) {
        const IncItems it = load_incidence(A, q, b1, b2);
#pragma unroll
        for (int b = 0; b < 4; b++)
            if (b < it.np) {
                int h = hash_slot_big(it.vi[b]), probe = 0;
                for (; probe < HB; probe++) {
                    const int old = atomicCAS(&key[h], -1, it.vi[b]);
                    if (old == -1 || old == it.vi[b]) break;
                    h = (h + 1) & (HB - 1);
                }
                if (probe == HB) overflow = true;
                else atomicOr(&msk[h], unsigned(it.mk[b])), atomicAdd(&cnt[h], 1);
            }
    }
    if (__syncthreads_or(overflow)) return false;
    // unique row vertices (any order: sorted next)
    for (int k = t; k < HB; k += BIG_THREADS)
        if (key[k] != -1) ukey[atomicAdd(&nU, 1)] = ((unsigned long long)(unsigned)key[k] << 32) | unsigned(k);
    __syncthreads();
    const int U = nU;
    int npow2 = 1;
    while (npow2 < U) npow2 <<= 1;
    for (int k = U + t; k < npow2; k += BIG_THREADS) ukey[k] = ~0ull;
    __syncthreads();
    bitonic_sort<BIG_THREADS>(ukey, npow2, t);
    // run starts (thread t owns the contiguous unique rows [lo, hi)), descriptors, pattern counts
    const int L = (U + BIG_THREADS - 1) / BIG_THREADS, lo = min(U, t * L), hi = min(U, lo + L);
    int mine = 0;
    for (int u = lo; u < hi; u++) mine += cnt[unsigned(ukey[u])];
    scan[t] = mine;
    __syncthreads();
    int start = 0;
    for (int k = 0; k < t; k++) start += scan[k];
    const int ioff = A.itemoff[v];
    int c0 = 0, c1 = 0, c2 = 0;
    for (int u = lo; u < hi; u++) {
        const int slot = int(unsigned(ukey[u]));
        base[slot] = start;
        A.udesc[ioff + u] = make_int2(start, int(unsigned(ukey[u] >> 32)));
        start += cnt[slot];
        const unsigned mk = msk[slot];
        c0 += __popc(mk & 0x49u), c1 += __popc(mk & 0x92u), c2 += __popc(mk & 0x124u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c0 += __shfl_xor_sync(0xffffffffu, c0, o);
        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    }
    __syncthreads(); // scan[] is reused for the block reduction
    if ((t & 31) == 0) scan[3 * (t >> 5)] = c0, scan[3 * (t >> 5) + 1] = c1, scan[3 * (t >> 5) + 2] = c2;
    __syncthreads();
    if (t < 3) {
        int sum = 0;
        for (int w = 0; w < BIG_THREADS / 32; w++) sum += scan[3 * w + t];
        A.cnt[3 * size_t(v) + t] = sum;
    }
    if (t == 0) A.colU[v] = U;
    // stable placement: chunks of BIG_THREADS incidences in order, point by point, warp after warp
    const int lane = t & 31, warp = t >> 5;
    for (int q0 = s; q0 < e; q0 += BIG_THREADS) {
        const int q = q0 + t;
        IncItems it;
        it.np = 0;
        if (q < e) it = load_incidence(A, q, b1, b2);
        for (int b = 0; b < 4; b++) {
            int h = 0;
            const bool valid = b < it.np;
            if (valid) {
                h = hash_slot_big(it.vi[b]);
                while (key[h] != it.vi[b]) h = (h + 1) & (HB - 1);
            }
            for (int w = 0; w < BIG_THREADS / 32; w++) {
                if (warp == w) {
                    const unsigned vm = __ballot_sync(0xffffffffu, valid);
                    if (valid) {
                        const unsigned peers = __match_any_sync(vm, h);
                        const int rank = __popc(peers & ((1u << lane) - 1));
                        const int pos = base[h];
                        __syncwarp(vm);
                        if (rank == 0) base[h] = pos + __popc(peers);
                        A.sref[ioff + pos + rank] = block_ref(A, it, b);
                    }
                }
                __syncthreads();
            }
        }
    }
    return true;
}

// columns the warp kernel could not take: one block per column, looping over the list; columns beyond CTA_CAP
// sort in a per-block global scratch region (need[0] reports the size required if it is too small)
__global__ void __launch_bounds__(BIG_THREADS)
    k_hess_symbolic_big(SymArgs A, int cta_cap, int use_hash, const int* __restrict__ big, const unsigned long long* nbig, char* scratch,
                        unsigned long long scratch_items, unsigned long long* need)
{
    extern __shared__ __align__(16) char smem[];
    __shared__ int scan[BIG_THREADS];
    const unsigned long long n = *nbig;
    for (unsigned long long idx = blockIdx.x; idx < n; idx += gridDim.x) {
        const int v = big[idx];
        const int R = A.colR[v];
        // columns the shared-memory sort cannot hold (and all of them under the test hook use_hash == 2): block-wide hash
        if (use_hash && (R > cta_cap || use_hash == 2)) {
            const bool done = column_symbolic_hash_cta(A, v, threadIdx.x, smem, scan);
            __syncthreads();
            if (done) continue;
        }
        int npow2 = 1;
        while (npow2 < R) npow2 <<= 1;
        char* base = smem;
        size_t cap = CTA_CAP;
        if (R > cta_cap) {
            if ((unsigned long long)npow2 > scratch_items) {
                if (threadIdx.x == 0) atomicMax(need, (unsigned long long)npow2);
                continue; // the host grows the scratch and repeats the assembly
            }
            base = scratch + size_t(blockIdx.x) * scratch_items * 14;
            cap = scratch_items;
        }
        unsigned long long* keys = reinterpret_cast<unsigned long long*>(base);
        unsigned* refs = reinterpret_cast<unsigned*>(base + cap * 8);
        unsigned short* masks = reinterpret_cast<unsigned short*>(base + cap * 12);
        column_symbolic_sort<BIG_THREADS>(A, v, threadIdx.x, keys, refs, masks, scan);
        __syncthreads();
    }
}

// ---- pass 2: one warp per column; 9 lanes per unique block (one per entry of the 3x3 block), three blocks at a time.
// A lane walks its block's run in the fixed item order, adding entry k of every gathered block (the 9 lanes of a
// group read 72 contiguous bytes) and noting whether any addend was non-zero: that IS the reference's pattern
// (local_to_global.hpp:290-291).  Positions inside the three scalar columns follow from one ballot.
// The loads form a dependent chain (descriptor -> block references -> blocks); it is software-pipelined: while the
// blocks of one group of runs are in flight, the descriptors and references of the next group are fetched.
// NUM_BATCH: block references fetched ahead per run (12 by default; IPCB_NUM_BATCH=8 / 16 select the other variants)
template <int NUM_BATCH> struct RunRefs {
    int start, len, row;
    unsigned ref[NUM_BATCH];
};
template <int NUM_BATCH> __device__ __forceinline__ RunRefs<NUM_BATCH> load_run(const int2* __restrict__ ud, const unsigned* __restrict__ sr, int u, int U, int R, bool lane_ok)
{
    RunRefs<NUM_BATCH> rr;
    rr.start = 0, rr.len = 0, rr.row = 0;
    if (lane_ok && u < U) {
        const int2 d = ud[u];
        rr.start = d.x, rr.row = d.y;
        rr.len = (u + 1 < U ? ud[u + 1].x : R) - d.x;
    }
#pragma unroll
    for (int x = 0; x < NUM_BATCH; x++) rr.ref[x] = x < rr.len ? sr[rr.start + x] : 0u;
    return rr;
}
// the two halves of load_run, for a three-deep pipeline: descriptors two rounds ahead, references one round ahead
struct RunDesc {
    int start, len, row;
};
__device__ __forceinline__ RunDesc load_desc(const int2* __restrict__ ud, int u, int U, int R, bool lane_ok)
{
    RunDesc d { 0, 0, 0 };
    if (lane_ok && u < U) {
        const int2 a = ud[u];
        d.start = a.x, d.row = a.y;
        d.len = (u + 1 < U ? ud[u + 1].x : R) - a.x;
    }
    return d;
}
template <int NUM_BATCH> __device__ __forceinline__ RunRefs<NUM_BATCH> load_refs(const unsigned* __restrict__ sr, const RunDesc& d)
{
    RunRefs<NUM_BATCH> rr;
    rr.start = d.start, rr.len = d.len, rr.row = d.row;
#pragma unroll
    for (int x = 0; x < NUM_BATCH; x++) rr.ref[x] = x < d.len ? sr[d.start + x] : 0u;
    return rr;
}
template <int NUM_BATCH, int REM, int COOP_UNROLL = 1>
__global__ void __launch_bounds__(32 * SYM_WARPS)
    k_hess_numeric(int nV, const int* __restrict__ colR, const int* __restrict__ colU, const int* __restrict__ itemoff,
                   const unsigned* __restrict__ sref, const int2* __restrict__ udesc, const double* __restrict__ blk,
                   const int* __restrict__ outer, int* __restrict__ inner, double* __restrict__ vals, const int* __restrict__ active, const int* __restrict__ nactive,
                   int big_items, int* __restrict__ big, unsigned long long* nbig)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * SYM_WARPS + (threadIdx.x >> 5);
    if (w >= *nactive) return;
    const int v = active[w]; // spatial visiting order of the columns (see k_hess_symbolic)
    const int U = colU[v];
    if (U == 0) return;
    const int R = colR[v], ioff = itemoff[v];
    if (R > big_items) { // a giant column (high-valence vertex in contact): one block instead of one warp (k_hess_numeric_big)
        if (lane == 0) big[atomicAdd(nbig, 1ull)] = v;
        return;
    }
    const int g = lane / 9, k = lane - 9 * g, l = k % 3, r = k / 3, kt = 3 * l + r; // kt: the same entry of the transposed block
    const bool lane_ok = g < 3;
    const unsigned colmask = 0x1249249u << l; // lanes of the same scalar column
    const int2* ud = udesc + ioff;
    const unsigned* sr = sref + ioff;
    int base = lane_ok ? outer[3 * size_t(v) + l] : 0;
    // three-deep pipeline: while the blocks of round r are in flight, the references of round r + 1 (their descriptors arrived one
    // round ago) and the descriptors of round r + 2 are requested — one memory latency per round instead of two chained ones
    RunRefs<NUM_BATCH> cur = load_run<NUM_BATCH>(ud, sr, g, U, R, lane_ok);
    RunDesc dnext = load_desc(ud, 3 + g, U, R, lane_ok);
    for (int u0 = 0; u0 < U; u0 += 3) {
        double val[NUM_BATCH];
#pragma unroll
        for (int x = 0; x < NUM_BATCH; x++) {
            val[x] = 0.0;
            if (x < cur.len) val[x] = __ldg(blk + size_t(cur.ref[x] >> 1) * 9 + ((cur.ref[x] & 1u) ? kt : k));
        }
        const RunRefs<NUM_BATCH> nxt = load_refs<NUM_BATCH>(sr, dnext);
        dnext = load_desc(ud, u0 + 6 + g, U, R, lane_ok);
        double acc = 0.0;
        bool nz = false;
#pragma unroll
        for (int x = 0; x < NUM_BATCH; x++)
            if (x < cur.len) {
                acc += val[x];
                nz |= val[x] != 0.0;
            }
        // long runs (every column has one: its diagonal block gathers one block per incidence, ~46 on the dense scenes)
        if constexpr (REM == 0) {
            // cooperative remainder: the three lane groups share the rest of every long run of the round (contiguous thirds,
            // combined in group order: a fixed association, so the sum is reproducible) instead of idling behind its owner
            unsigned longmask = __ballot_sync(0xffffffffu, lane_ok && k == 0 && cur.len > NUM_BATCH);
            while (longmask) {
                const int src = __ffs(longmask) - 1; // lane 9 * (owner group)
                longmask &= longmask - 1;
                const int st = __shfl_sync(0xffffffffu, cur.start, src), ln = __shfl_sync(0xffffffffu, cur.len, src);
                const int part = (ln - NUM_BATCH + 2) / 3;
                const int lo = NUM_BATCH + g * part, hi = min(ln, lo + part);
                double p = 0.0;
                bool pz = false;
                // COOP_UNROLL gathers in flight, branch-free (the index is clamped, a slot past the end adds +0.0): the lanes of a
                // group agree on lo / hi, the groups do not, and per-slot branches would serialise them
                for (int j = lo; j < hi; j += COOP_UNROLL) {
                    unsigned rf[COOP_UNROLL];
                    double wv[COOP_UNROLL];
#pragma unroll
                    for (int x = 0; x < COOP_UNROLL; x++) rf[x] = sr[st + min(j + x, hi - 1)];
#pragma unroll
                    for (int x = 0; x < COOP_UNROLL; x++) wv[x] = __ldg(blk + size_t(rf[x] >> 1) * 9 + ((rf[x] & 1u) ? kt : k));
#pragma unroll
                    for (int x = 0; x < COOP_UNROLL; x++) {
                        const bool in = j + x < hi;
                        p += in ? wv[x] : 0.0;
                        pz |= in && wv[x] != 0.0;
                    }
                }
                const double p0 = __shfl_sync(0xffffffffu, p, k), p1 = __shfl_sync(0xffffffffu, p, 9 + k), p2 = __shfl_sync(0xffffffffu, p, 18 + k);
                const unsigned zm = __ballot_sync(0xffffffffu, pz);
                if (lane_ok && 9 * g == src) {
                    acc += p0, acc += p1, acc += p2;
                    nz |= (((zm >> k) | (zm >> (9 + k)) | (zm >> (18 + k))) & 1u) != 0;
                }
            }
        } else {
            for (int j = NUM_BATCH; j < cur.len; j += REM) {
                double w[REM];
#pragma unroll
                for (int x = 0; x < REM; x++) {
                    w[x] = 0.0;
                    if (j + x < cur.len) {
                        const unsigned ref = sr[cur.start + j + x];
                        w[x] = __ldg(blk + size_t(ref >> 1) * 9 + ((ref & 1u) ? kt : k));
                    }
                }
#pragma unroll
                for (int x = 0; x < REM; x++)
                    if (j + x < cur.len) {
                        acc += w[x];
                        nz |= w[x] != 0.0;
                    }
            }
        }
        const bool present = cur.len > 0 && nz;
        const unsigned pm = __ballot_sync(0xffffffffu, present) & colmask;
        if (present) {
            const int p = base + __popc(pm & ((1u << lane) - 1));
            inner[p] = 3 * cur.row + r;
            vals[p] = acc;
        }
        base += __popc(pm);
        cur = nxt;
    }
}

// ---- pass 2, column-lane variant: 3 lanes per unique block (one per scalar column of the block), ten blocks at a time.  A lane
// sums the three entries of its column over the run, so the reference decode / address arithmetic is paid once per three gathered
// entries and a warp covers ten blocks per round instead of three.  Same item order per entry as k_hess_numeric: bit-identical.
template <int NUM_BATCH>
__global__ void __launch_bounds__(32 * SYM_WARPS)
    k_hess_numeric_col(int nV, const int* __restrict__ colR, const int* __restrict__ colU, const int* __restrict__ itemoff,
                       const unsigned* __restrict__ sref, const int2* __restrict__ udesc, const double* __restrict__ blk,
                       const int* __restrict__ outer, int* __restrict__ inner, double* __restrict__ vals, const int* __restrict__ active, const int* __restrict__ nactive,
                       int big_items, int* __restrict__ big, unsigned long long* nbig)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * SYM_WARPS + (threadIdx.x >> 5);
    if (w >= *nactive) return;
    const int v = active[w];
    const int U = colU[v];
    if (U == 0) return;
    const int R = colR[v], ioff = itemoff[v];
    if (R > big_items) {
        if (lane == 0) big[atomicAdd(nbig, 1ull)] = v;
        return;
    }
    const int g = lane / 3, l = lane - 3 * g;
    const bool lane_ok = lane < 30;
    const unsigned colmask = 0x09249249u << l; // lanes of the same scalar column (bits l, l + 3, ..., l + 27)
    const unsigned below = (1u << lane) - 1;
    const int2* ud = udesc + ioff;
    const unsigned* sr = sref + ioff;
    int base = lane_ok ? outer[3 * size_t(v) + l] : 0;
    RunRefs<NUM_BATCH> cur = load_run<NUM_BATCH>(ud, sr, g, U, R, lane_ok);
    for (int u0 = 0; u0 < U; u0 += 10) {
        double a0[NUM_BATCH], a1[NUM_BATCH], a2[NUM_BATCH];
#pragma unroll
        for (int x = 0; x < NUM_BATCH; x++) {
            a0[x] = a1[x] = a2[x] = 0.0;
            if (x < cur.len) {
                const bool tr = cur.ref[x] & 1u; // stored by the other vertex: entry (r, l) is the stored (l, r)
                const double* b = blk + size_t(cur.ref[x] >> 1) * 9 + (tr ? 3 * l : l);
                const int st = tr ? 1 : 3;
                a0[x] = __ldg(b), a1[x] = __ldg(b + st), a2[x] = __ldg(b + 2 * st);
            }
        }
        const RunRefs<NUM_BATCH> nxt = load_run<NUM_BATCH>(ud, sr, u0 + 10 + g, U, R, lane_ok);
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        bool z0 = false, z1 = false, z2 = false;
#pragma unroll
        for (int x = 0; x < NUM_BATCH; x++)
            if (x < cur.len) {
                s0 += a0[x], s1 += a1[x], s2 += a2[x];
                z0 |= a0[x] != 0.0, z1 |= a1[x] != 0.0, z2 |= a2[x] != 0.0;
            }
        for (int j = NUM_BATCH; j < cur.len; j++) {
            const unsigned ref = sr[cur.start + j];
            const bool tr = ref & 1u;
            const double* b = blk + size_t(ref >> 1) * 9 + (tr ? 3 * l : l);
            const int st = tr ? 1 : 3;
            const double w0 = __ldg(b), w1 = __ldg(b + st), w2 = __ldg(b + 2 * st);
            s0 += w0, s1 += w1, s2 += w2;
            z0 |= w0 != 0.0, z1 |= w1 != 0.0, z2 |= w2 != 0.0;
        }
        const bool have = cur.len > 0;
        const unsigned m0 = __ballot_sync(0xffffffffu, have && z0) & colmask;
        const unsigned m1 = __ballot_sync(0xffffffffu, have && z1) & colmask;
        const unsigned m2 = __ballot_sync(0xffffffffu, have && z2) & colmask;
        if (have) { // inside a scalar column: blocks in order, rows ascending inside a block
            int p = base + __popc(m0 & below) + __popc(m1 & below) + __popc(m2 & below);
            if (z0) inner[p] = 3 * cur.row, vals[p] = s0, p++;
            if (z1) inner[p] = 3 * cur.row + 1, vals[p] = s1, p++;
            if (z2) inner[p] = 3 * cur.row + 2, vals[p] = s2;
        }
        base += __popc(m0) + __popc(m1) + __popc(m2);
        cur = nxt;
    }
}

// giant columns: one block per column.  Phase 1: the 8 warps sum 24 unique blocks at a time (same 9-lanes-per-block layout and the
// same item order as k_hess_numeric, so the values are bit-identical) into shared memory, with the 9-bit non-zero pattern of every
// block; phase 2: positions inside the three scalar columns from a block-wide scan of the pattern counts, then the stores.
// Columns with more than NUMERIC_UCAP unique blocks are processed in segments that carry the three running positions.
constexpr int NUMERIC_UCAP = 1024;
constexpr int NBIG_THREADS = 1024; // 32 warps x 3 blocks = 96 unique blocks per round
__global__ void __launch_bounds__(NBIG_THREADS)
    k_hess_numeric_big(const int* __restrict__ big, const unsigned long long* nbig, const int* __restrict__ colR, const int* __restrict__ colU,
                       const int* __restrict__ itemoff, const unsigned* __restrict__ sref, const int2* __restrict__ udesc,
                       const double* __restrict__ blk, const int* __restrict__ outer, int* __restrict__ inner, double* __restrict__ vals,
                       int ucap)
{
    extern __shared__ __align__(16) char smem[];
    double* acc_s = reinterpret_cast<double*>(smem);                                 // NUMERIC_UCAP x 9
    unsigned short* pat = reinterpret_cast<unsigned short*>(acc_s + NUMERIC_UCAP * 9); // NUMERIC_UCAP
    __shared__ int wtot[3 * (NBIG_THREADS / 32)];
    __shared__ int run_base[3];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int g = lane / 9, k = lane - 9 * g, l = k % 3, r = k / 3, kt = 3 * l + r;
    const bool lane_ok = g < 3;
    const unsigned long long n = *nbig;
    for (unsigned long long idx = blockIdx.x; idx < n; idx += gridDim.x) {
        const int v = big[idx];
        const int U = colU[v], R = colR[v], ioff = itemoff[v];
        const int2* ud = udesc + ioff;
        const unsigned* sr = sref + ioff;
        if (t < 3) run_base[t] = outer[3 * size_t(v) + t];
        __syncthreads();
        for (int seg = 0; seg < U; seg += ucap) {
            const int nseg = min(ucap, U - seg);
            for (int u0 = 3 * warp; u0 < nseg; u0 += 3 * (NBIG_THREADS / 32)) {
                const int u = seg + u0 + g;
                double acc = 0.0;
                bool nz = false;
                const bool has = lane_ok && u0 + g < nseg;
                if (has) {
                    const int start = ud[u].x, len = (u + 1 < U ? ud[u + 1].x : R) - start;
                    int j = 0;
                    for (; j + 8 <= len; j += 8) { // eight gathers in flight, added in item order
                        unsigned rf[8];
                        double wv[8];
#pragma unroll
                        for (int x = 0; x < 8; x++) rf[x] = sr[start + j + x];
#pragma unroll
                        for (int x = 0; x < 8; x++) wv[x] = __ldg(blk + size_t(rf[x] >> 1) * 9 + ((rf[x] & 1u) ? kt : k));
#pragma unroll
                        for (int x = 0; x < 8; x++) acc += wv[x], nz |= wv[x] != 0.0;
                    }
                    for (; j < len; j++) {
                        const unsigned ref = sr[start + j];
                        const double w = __ldg(blk + size_t(ref >> 1) * 9 + ((ref & 1u) ? kt : k));
                        acc += w;
                        nz |= w != 0.0;
                    }
                    acc_s[(u0 + g) * 9 + k] = acc;
                }
                const unsigned pm = __ballot_sync(0xffffffffu, has && nz);
                if (has && k == 0) pat[u0 + g] = (unsigned short)((pm >> (9 * g)) & 0x1ffu);
            }
            __syncthreads();
            // entries per scalar column of the blocks thread t owns ([lo, hi) of the segment): warp scan + scan of the warp totals
            const int L = (nseg + NBIG_THREADS - 1) / NBIG_THREADS, lo = min(nseg, t * L), hi = min(nseg, lo + L);
            int c[3] = { 0, 0, 0 };
            for (int u = lo; u < hi; u++) {
                const unsigned m = pat[u];
                c[0] += __popc(m & 0x49u), c[1] += __popc(m & 0x92u), c[2] += __popc(m & 0x124u);
            }
            int incl[3] = { c[0], c[1], c[2] };
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
#pragma unroll
                for (int x = 0; x < 3; x++) {
                    const int up = __shfl_up_sync(0xffffffffu, incl[x], o);
                    if (lane >= o) incl[x] += up;
                }
            if (lane == 31) wtot[3 * warp] = incl[0], wtot[3 * warp + 1] = incl[1], wtot[3 * warp + 2] = incl[2];
            __syncthreads();
            int p[3], tot[3] = { 0, 0, 0 };
#pragma unroll
            for (int x = 0; x < 3; x++) p[x] = run_base[x] + incl[x] - c[x];
            for (int q = 0; q < NBIG_THREADS / 32; q++)
#pragma unroll
                for (int x = 0; x < 3; x++) {
                    const int cq = wtot[3 * q + x];
                    p[x] += q < warp ? cq : 0;
                    tot[x] += cq;
                }
            for (int u = lo; u < hi; u++) {
                const unsigned m = pat[u];
                const int row = ud[seg + u].y;
#pragma unroll
                for (int e = 0; e < 9; e++) // e = 3 r + l: rows ascend inside a scalar column
                    if (m & (1u << e)) {
                        const int x = e % 3;
                        inner[p[x]] = 3 * row + e / 3;
                        vals[p[x]] = acc_s[u * 9 + e];
                        p[x]++;
                    }
            }
            __syncthreads();
            if (t < 3) run_base[t] += tot[t];
            __syncthreads();
        }
    }
}

// ---- row block of a sharded Hessian: which collisions touch an owned vertex ---------------------------------------
__device__ inline int stencil_ids_only(int kind, int2 id, const MeshView& m, int* vid)
{
    if (kind == IPCB_VV) {
        vid[0] = id.x, vid[1] = id.y;
        return 2;
    }
    if (kind == IPCB_EV) {
        const int2 e = __ldg(m.E + id.x);
        vid[0] = id.y, vid[1] = e.x, vid[2] = e.y;
        return 3;
    }
    if (kind == IPCB_EE) {
        const int2 ea = __ldg(m.E + id.x), eb = __ldg(m.E + id.y);
        vid[0] = ea.x, vid[1] = ea.y, vid[2] = eb.x, vid[3] = eb.y;
        return 4;
    }
    const int4 f = __ldg(m.F + id.x);
    vid[0] = id.y, vid[1] = f.x, vid[2] = f.y, vid[3] = f.z;
    return 4;
}
__global__ void k_touch_flags(CollView c, MeshView m, int v_lo, int v_hi, unsigned char* __restrict__ flag)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    int vid[4];
    const int n = stencil_ids_only(c.kind, c.ids[i], m, vid);
    bool any = false;
    for (int a = 0; a < n; a++) any |= vid[a] >= v_lo && vid[a] < v_hi;
    flag[i] = any;
}

// the ids part of the Hessian records (stencil vertex ids, incidences): needs no arithmetic, so it runs — with the incidence sort
// behind it — on a side stream while the local-Hessian kernels compute the blocks
template <int KIND>
__global__ void k_write_records(CollView c, MeshView m, int64_t gi0, int64_t inc0, HessOut out, const int* __restrict__ sel, int64_t nsel)
{
    constexpr int NP = KIND == IPCB_VV ? 2 : (KIND == IPCB_EV ? 3 : 4);
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nsel) return;
    const int64_t i = sel ? sel[t] : t;
    int vid[4];
    stencil_ids_only(KIND, c.ids[i], m, vid);
    write_record<NP>(out, gi0 + t, inc0 + t * NP, vid);
}

// ---- balanced row blocks of a sharded Hessian: per vertex, the number of 3x3 blocks its column receives ----------
template <int KIND> __global__ void k_vertex_load(CollView c, MeshView m, int* __restrict__ load)
{
    constexpr int NP = KIND == IPCB_VV ? 2 : (KIND == IPCB_EV ? 3 : 4);
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    const int2 id = c.ids[i];
    int vid[4];
    if (KIND == IPCB_VV) {
        vid[0] = id.x, vid[1] = id.y;
    } else if (KIND == IPCB_EV) {
        const int2 e = __ldg(m.E + id.x);
        vid[0] = id.y, vid[1] = e.x, vid[2] = e.y;
    } else if (KIND == IPCB_EE) {
        const int2 ea = __ldg(m.E + id.x), eb = __ldg(m.E + id.y);
        vid[0] = ea.x, vid[1] = ea.y, vid[2] = eb.x, vid[3] = eb.y;
    } else {
        const int4 f = __ldg(m.F + id.x);
        vid[0] = id.y, vid[1] = f.x, vid[2] = f.y, vid[3] = f.z;
    }
#pragma unroll
    for (int a = 0; a < NP; a++) atomicAdd(load + vid[a], NP);
}
// bounds[r] = first vertex whose exclusive prefix load reaches r * total / world
__global__ void k_balance_bounds(int nV, const int* __restrict__ prefix, int world, int* __restrict__ bounds)
{
    const int r = threadIdx.x;
    if (r > world) return;
    const long long total = prefix[nV];
    const long long want = total * r / world;
    int lo = 0, hi = nV;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (prefix[mid] < want) lo = mid + 1;
        else hi = mid;
    }
    bounds[r] = r == 0 ? 0 : (r == world ? nV : lo);
}
void hessian_balanced_row_blocks(ipcb_ctx* ctx, int world, int32_t* bounds)
{
    cudaStream_t s = ctx->stream;
    const int nV = ctx->nV;
    if (world < 1 || world > 1024) throw Error("bad world size");
    ctx->hcolR.reserve(size_t(nV) + 2), ctx->hitemoff.reserve(size_t(nV) + 2 + 1032);
    IPCB_CUDA(cudaMemsetAsync(ctx->hcolR.p, 0, sizeof(int) * (size_t(nV) + 1), s));
    const MeshView m = mesh_view(ctx);
    const int64_t n0 = ctx->coll[0].count, n1 = ctx->coll[1].count, n2 = ctx->coll[2].count, n3 = ctx->coll[3].count;
    if (n0) k_vertex_load<IPCB_VV><<<grid_for(n0, 256), 256, 0, s>>>(view(ctx, 0), m, ctx->hcolR.p), ctx->launches++;
    if (n1) k_vertex_load<IPCB_EV><<<grid_for(n1, 256), 256, 0, s>>>(view(ctx, 1), m, ctx->hcolR.p), ctx->launches++;
    if (n2) k_vertex_load<IPCB_EE><<<grid_for(n2, 256), 256, 0, s>>>(view(ctx, 2), m, ctx->hcolR.p), ctx->launches++;
    if (n3) k_vertex_load<IPCB_FV><<<grid_for(n3, 256), 256, 0, s>>>(view(ctx, 3), m, ctx->hcolR.p), ctx->launches++;
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, ctx->hcolR.p, ctx->hitemoff.p, nV + 1, s);
    ctx->cubtmp.reserve(bytes + 1024);
    cub::DeviceScan::ExclusiveSum(ctx->cubtmp.p, bytes, ctx->hcolR.p, ctx->hitemoff.p, nV + 1, s);
    int* d_bounds = ctx->hitemoff.p + nV + 2;
    k_balance_bounds<<<1, 1024, 0, s>>>(nV, ctx->hitemoff.p, world, d_bounds);
    ctx->launches += 3;
    IPCB_CUDA(cudaGetLastError());
    IPCB_CUDA(cudaMemcpyAsync(bounds, d_bounds, sizeof(int) * (world + 1), cudaMemcpyDeviceToHost, s));
    IPCB_CUDA(cudaStreamSynchronize(s));
    for (int r = 1; r <= world; r++) bounds[r] = std::max(bounds[r], bounds[r - 1]);
}

__global__ void k_zero_int(int64_t n, int* p)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

void hessian_empty(ipcb_ctx* ctx)
{
    ctx->outer.reserve(3 * size_t(ctx->nV) + 1);
    ctx->nnz = 0;
    k_zero_int<<<grid_for(3 * size_t(ctx->nV) + 1, 256), 256, 0, ctx->stream>>>(3 * int64_t(ctx->nV) + 1, ctx->outer.p);
    ctx->launches++;
}

bool hess_counting_placement();
void hessian_records(ipcb_ctx* ctx, const int64_t nk[4], int v_lo, int v_hi, HessOut outs[4])
{
    const int64_t n0 = nk[0], n1 = nk[1], n2 = nk[2], n3 = nk[3];
    const int64_t ncoll = n0 + n1 + n2 + n3;
    const int64_t ninc = 2 * n0 + 3 * n1 + 4 * (n2 + n3);
    const int64_t nitems = 4 * n0 + 9 * n1 + 16 * (n2 + n3);
    // stored blocks (upper triangles): 3 / 6 / 10 per VV / EV / 4-point record
    int64_t blk0[4];
    hess_block_offsets(nk, blk0);
    const int64_t nblocks = blk0[3] + 10 * n3;
    if (ncoll >= (int64_t(1) << 27) || nitems > 0x7fffffffll || nblocks > 0x7fffffffll)
        throw Error("Hessian: more than 2^27 collisions / 2^31 local blocks on one device; shard the collision set");
    ctx->hvid.reserve(ncoll), ctx->hmask.reserve(size_t(ncoll) * HSLOTS), ctx->hblk.reserve(size_t(nblocks) * 9);
    ctx->hkey.reserve(ninc), ctx->hkey_sorted.reserve(ninc);
    unsigned long long* cnt = nullptr;
    if (hess_counting_placement()) {
        ctx->hcount.reserve(size_t(ctx->nV) + 2);
        IPCB_CUDA(cudaMemsetAsync(ctx->hcount.p, 0, (size_t(ctx->nV) + 2) * sizeof(unsigned long long), ctx->stream));
        cnt = ctx->hcount.p;
    }
    for (int k = 0; k < 4; k++)
        outs[k] = HessOut { ctx->hvid.p, ctx->hmask.p, ctx->hblk.p + size_t(blk0[k]) * 9, ctx->hkey.p, v_lo, v_hi, ctx->nV, 0, cnt };
}

// Grouping of the incidences by column vertex.  Default: counting placement — the kernels that write the records count the
// incidences per vertex (one 64-bit atomic each, three 21-bit fields by record size), a scan gives the column ranges, a scatter
// pass places every incidence in its column and record-size segment with one more atomic, and each column's few dozen
// entries are sorted in shared memory, which restores exactly the order a stable global sort by vertex produces (collision
// order: the summation order of the numeric pass stays reproducible).  IPCB_HESS_RADIX_INCIDENCES: the global radix sort
// (three passes over 8-byte keys) + binary searches for the ranges, kept as the A/B and test alternative.
constexpr int COLSORT_WARP_CAP = 512; // incidences a warp sorts in shared memory (up to 256: in registers)
bool hess_counting_placement()
{
    return getenv("IPCB_HESS_RADIX_INCIDENCES") == nullptr; // read per call: the tests switch it
}
__global__ void k_col_from_counts(int nV, const unsigned long long* __restrict__ cnt, int* __restrict__ colcount, int* __restrict__ colR,
                                  unsigned long long* maxcount, int report_above)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    int mine = 0;
    if (v <= nV) {
        const unsigned long long c = cnt[v];
        const int n0 = int(c & 0x1fffffull), n1 = int((c >> 21) & 0x1fffffull), n2 = int((c >> 42) & 0x1fffffull);
        colcount[v] = n0 + n1 + n2;
        colR[v] = v < nV ? 2 * n0 + 3 * n1 + 4 * n2 : 0;
        if (v < nV) mine = n0 + n1 + n2;
    } else if (v == nV + 1) {
        colcount[v] = 0; // the scan over nV + 2 entries reads it
    }
    for (int o = 16; o > 0; o >>= 1) mine = max(mine, __shfl_xor_sync(0xffffffffu, mine, o));
    if ((threadIdx.x & 31) == 0 && mine > report_above) atomicMax(maxcount, (unsigned long long)mine); // rare: large columns only
}
__global__ void k_col_bounds(int nV, const unsigned long long* __restrict__ cnt, const int* __restrict__ colinc, int2* __restrict__ colb)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    const unsigned long long c = cnt[v];
    const int n0 = int(c & 0x1fffffull), n1 = int((c >> 21) & 0x1fffffull);
    colb[v] = make_int2(colinc[v] + n0, colinc[v] + n0 + n1);
}
__global__ void k_scatter_incidences(int64_t n, const unsigned long long* __restrict__ key, unsigned ref_ev, unsigned ref_ee,
                                     const unsigned long long* __restrict__ cnt, const int* __restrict__ colinc, unsigned long long* cursor,
                                     unsigned long long* __restrict__ out, unsigned v_none)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = key[i];
    const unsigned v = unsigned(k >> 32), ref = unsigned(k);
    if (v >= v_none) return; // an incidence of a vertex outside the rank's row block: parked, never read
    const int cls = ref < ref_ev ? 0 : (ref < ref_ee ? 1 : 2);
    const unsigned long long c = cnt[v];
    const int seg = cls == 0 ? 0 : (cls == 1 ? int(c & 0x1fffffull) : int(c & 0x1fffffull) + int((c >> 21) & 0x1fffffull));
    const unsigned long long old = atomicAdd(cursor + v, 1ull << (21 * cls));
    out[colinc[v] + seg + int((old >> (21 * cls)) & 0x1fffffull)] = k;
}
// bitonic network over 32 E values, E per lane (element e = lane + 32 h): partners at distance j < 32 by shuffle, at j >= 32
// inside the lane; ascending where (e & k) == 0
template <int E> __device__ __forceinline__ void warp_bitonic(unsigned (&a)[E], int lane)
{
#pragma unroll
    for (int k = 2; k <= 32 * E; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int hj = j / 32;
#pragma unroll
                for (int h = 0; h < E; h++)
                    if ((h & hj) == 0) {
                        const bool up = ((32 * h) & k) == 0;
                        const unsigned lo = min(a[h], a[h | hj]), hi = max(a[h], a[h | hj]);
                        a[h] = up ? lo : hi, a[h | hj] = up ? hi : lo;
                    }
            } else {
                const bool lower = (lane & j) == 0;
#pragma unroll
                for (int h = 0; h < E; h++) {
                    const unsigned b = __shfl_xor_sync(0xffffffffu, a[h], j);
                    const bool up = k < 32 ? (lane & k) == 0 : ((32 * h) & k) == 0;
                    a[h] = (lower == up) ? min(a[h], b) : max(a[h], b);
                }
            }
        }
}
template <int E> __device__ __forceinline__ void sort_column_registers(unsigned long long* col, int L, int lane)
{
    const unsigned long long hi = col[0] & 0xffffffff00000000ull;
    unsigned a[E];
#pragma unroll
    for (int h = 0; h < E; h++) a[h] = lane + 32 * h < L ? unsigned(col[lane + 32 * h]) : 0xffffffffu;
    warp_bitonic<E>(a, lane);
#pragma unroll
    for (int h = 0; h < E; h++)
        if (lane + 32 * h < L) col[lane + 32 * h] = hi | a[h];
}
// every active column's incidences into ascending order (all of a column share the vertex bits: the order is the collision order)
__global__ void __launch_bounds__(32 * SYM_WARPS)
    k_sort_columns(const int* __restrict__ active, const int* __restrict__ nactive, const int* __restrict__ colinc, unsigned long long* inc,
                   int* __restrict__ big, unsigned long long* nbig, int warp_cap)
{
    __shared__ unsigned long long keys[SYM_WARPS][COLSORT_WARP_CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = blockIdx.x * SYM_WARPS + warp;
    if (w >= *nactive) return;
    const int v = active[w];
    const int s = colinc[v], L = colinc[v + 1] - s;
    if (L <= 1) return;
    if (L > warp_cap) {
        if (lane == 0) big[atomicAdd(nbig, 1ull)] = v;
        return;
    }
    // the common cases (a few dozen to a few hundred incidences): the 32-bit references (the vertex bits are the column's)
    // sorted in registers, E per lane, by a bitonic network over shuffles
    if (L <= 64) {
        sort_column_registers<2>(inc + s, L, lane);
        return;
    }
    if (L <= 128) {
        sort_column_registers<4>(inc + s, L, lane);
        return;
    }
    if (L <= 256) {
        sort_column_registers<8>(inc + s, L, lane);
        return;
    }
    int npow2 = 2;
    while (npow2 < L) npow2 <<= 1;
    for (int q = lane; q < npow2; q += 32) keys[warp][q] = q < L ? inc[s + q] : ~0ull;
    __syncwarp();
    bitonic_sort<32>(keys[warp], npow2, lane);
    for (int q = lane; q < L; q += 32) inc[s + q] = keys[warp][q];
}
// columns beyond the warp's shared memory: one block each (up to COLSORT_CTA_CAP entries)
constexpr int COLSORT_CTA_CAP = 16384;
static int colsort_cta_cap() // test hook: a small scene reaches the redo with the radix sort
{
    const char* e = getenv("IPCB_HESS_COLSORT_CTA_CAP");
    return e ? std::min(COLSORT_CTA_CAP, std::max(2, atoi(e))) : COLSORT_CTA_CAP;
}
__global__ void __launch_bounds__(BIG_THREADS)
    k_sort_columns_big(const int* __restrict__ big, const unsigned long long* nbig, const int* __restrict__ colinc, unsigned long long* inc, int cap)
{
    extern __shared__ __align__(16) char smem[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);
    const unsigned long long n = *nbig;
    for (unsigned long long idx = blockIdx.x; idx < n; idx += gridDim.x) {
        const int v = big[idx];
        const int s = colinc[v], L = colinc[v + 1] - s;
        int npow2 = 2;
        while (npow2 < L) npow2 <<= 1;
        if (L > cap) continue; // reported through maxcount: the assembly is redone with the radix sort
        for (int q = threadIdx.x; q < npow2; q += BIG_THREADS) keys[q] = q < L ? inc[s + q] : ~0ull;
        __syncthreads();
        bitonic_sort<BIG_THREADS>(keys, npow2, threadIdx.x);
        for (int q = threadIdx.x; q < L; q += BIG_THREADS) inc[s + q] = keys[q];
        __syncthreads();
    }
}

struct ActiveColumn {
    const int* colR;
    __device__ bool operator()(int v) const { return colR[v] > 0; }
};

// Assembly of the per-collision records into compressed columns (see the comment block above k_col_ranges)
void hessian_assemble(ipcb_ctx* ctx, const int64_t nk[4])
{
    hessian_assemble_prepare(ctx, nk, ctx->stream);
    if (!hessian_assemble_finish(ctx, nk)) {
        hessian_assemble_prepare(ctx, nk, ctx->stream, true);
        hessian_assemble_finish(ctx, nk);
    }
}

void hessian_assemble_prepare(ipcb_ctx* ctx, const int64_t nk[4], cudaStream_t s, bool force_radix)
{
    const int nV = ctx->nV;
    const int64_t n0 = nk[0], n1 = nk[1], n2 = nk[2], n3 = nk[3];
    const int64_t gi0[4] = { 0, n0, n0 + n1, n0 + n1 + n2 };
    const int64_t ninc = 2 * n0 + 3 * n1 + 4 * (n2 + n3);
    const int64_t nitems = 4 * n0 + 9 * n1 + 16 * (n2 + n3);
    int64_t blk0[4];
    hess_block_offsets(nk, blk0);
    ctx->outer.reserve(3 * size_t(nV) + 1);
    // stage timers (only when ctx->timing is on): the three kernels of the assembly are timed one by one
    std::unique_ptr<Stage> st(new Stage(ctx, "hess_incidences", s));
    int vbits = 1;
    while ((1ll << vbits) <= nV) vbits++; // the value nV marks incidences of vertices outside the rank's row block
    // 1. incidences grouped by vertex (stable: each column keeps the collision order)
    size_t b1 = 0, b2 = 0, b3 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, b1, ctx->hkey.p, ctx->hkey_sorted.p, ninc, 32, 32 + vbits, s);
    ctx->hcolinc.reserve(size_t(nV) + 2), ctx->hcolR.reserve(size_t(nV) + 2), ctx->hitemoff.reserve(size_t(nV) + 2);
    ctx->hcnt.reserve(3 * size_t(nV) + 1), ctx->hcolU.reserve(size_t(nV) + 1);
    cub::DeviceScan::ExclusiveSum(nullptr, b2, ctx->hcolR.p, ctx->hitemoff.p, nV + 2, s);
    cub::DeviceScan::ExclusiveSum(nullptr, b3, ctx->hcnt.p, ctx->outer.p, 3 * nV + 1, s);
    ctx->cubtmp.reserve(std::max(b1, std::max(b2, b3)) + 1024);
    const unsigned ref_ev = unsigned(gi0[1] * 4), ref_ee = unsigned(gi0[2] * 4);
    ctx->hcolb.reserve(size_t(nV) + 1);
    const bool counting = hess_counting_placement() && !force_radix;
    ctx->hess_counting_used = counting;
    if (counting) {
        // counts (written with the records) -> column sizes.  A column beyond what one block sorts in shared memory is left
        // unsorted and reported through `maxcount`; hessian_assemble_finish reads it at its own synchronisation point and the
        // assembly is then redone with the radix sort (hkey still holds the unsorted incidences) — no extra host round trip
        unsigned long long* maxcount = ctx->dCounters.p + 26;
        IPCB_CUDA(cudaMemsetAsync(maxcount, 0, sizeof(unsigned long long), s));
        // hitemoff doubles as the per-column incidence count until the item scan overwrites it
        k_col_from_counts<<<grid_for(size_t(nV) + 2, 256), 256, 0, s>>>(nV, ctx->hcount.p, ctx->hitemoff.p, ctx->hcolR.p, maxcount,
                                                                        std::min(COLSORT_WARP_CAP, colsort_cta_cap()));
        ctx->launches++;
    }
    if (counting) {
        Stage kt(ctx, "k:place(incidences)", s);
        cub::DeviceScan::ExclusiveSum(ctx->cubtmp.p, b2, ctx->hitemoff.p, ctx->hcolinc.p, nV + 2, s); // colinc[nV + 1] = all incidences
        k_col_bounds<<<grid_for(nV, 256), 256, 0, s>>>(nV, ctx->hcount.p, ctx->hcolinc.p, ctx->hcolb.p);
        ctx->hcursor.reserve(size_t(nV) + 2);
        IPCB_CUDA(cudaMemsetAsync(ctx->hcursor.p, 0, (size_t(nV) + 2) * sizeof(unsigned long long), s));
        k_scatter_incidences<<<grid_for(ninc, 256), 256, 0, s>>>(ninc, ctx->hkey.p, ref_ev, ref_ee, ctx->hcount.p, ctx->hcolinc.p, ctx->hcursor.p,
                                                               ctx->hkey_sorted.p, unsigned(nV));
        ctx->launches += 4;
    } else {
        {
            Stage kt(ctx, "k:radix_sort(incidences)", s);
            cub::DeviceRadixSort::SortKeys(ctx->cubtmp.p, b1, ctx->hkey.p, ctx->hkey_sorted.p, ninc, 32, 32 + vbits, s);
        }
        ctx->launches += 2 + (vbits + 7) / 8;
        // 2. column ranges
        k_col_ranges<<<grid_for(size_t(nV) + 1, 256), 256, 0, s>>>(nV, int(ninc), ctx->hkey_sorted.p, ref_ev, ref_ee, ctx->hcolinc.p, ctx->hcolR.p,
                                                                ctx->hcolb.p);
        ctx->launches++;
    }
    // item offsets
    cub::DeviceScan::ExclusiveSum(ctx->cubtmp.p, b2, ctx->hcolR.p, ctx->hitemoff.p, nV + 1, s);
    ctx->launches += 2;
    // 2b. the columns that have anything to assemble, in the Morton order of their vertices (from the last broad phase on this
    // context; any order is valid).  On a row block of a sharded Hessian three quarters (N = 4) of the columns are empty, on a
    // scene like C2 most vertices are not in contact: the per-column kernels run one warp per ACTIVE column, so that every
    // resident block is full of working warps (with one warp per vertex the working warps were 1 in 4 and the passes did not
    // scale with the rank count).
    static const bool id_order = getenv("IPCB_HESS_COLUMN_ID_ORDER") != nullptr; // A/B switch
    const int* order = (!id_order && ctx->vorder_valid && ctx->vtree.n == nV) ? ctx->vtree.ord_sorted.p : nullptr;
    ctx->hactive.reserve(size_t(nV) + 1);
    int* nactive = reinterpret_cast<int*>(ctx->dCounters.p + 25);
    {
        const ActiveColumn is_active { ctx->hcolR.p };
        size_t b4 = 0;
        if (order) cub::DeviceSelect::If(nullptr, b4, order, ctx->hactive.p, nactive, nV, is_active, s);
        else cub::DeviceSelect::If(nullptr, b4, cub::CountingInputIterator<int>(0), ctx->hactive.p, nactive, nV, is_active, s);
        ctx->hseltmp.reserve(b4 + 256);
        if (order) cub::DeviceSelect::If(ctx->hseltmp.p, b4, order, ctx->hactive.p, nactive, nV, is_active, s);
        else cub::DeviceSelect::If(ctx->hseltmp.p, b4, cub::CountingInputIterator<int>(0), ctx->hactive.p, nactive, nV, is_active, s);
        IPCB_CUDA(cudaMemsetAsync(ctx->hcnt.p, 0, (3 * size_t(nV) + 1) * sizeof(int), s));
        IPCB_CUDA(cudaMemsetAsync(ctx->hcolU.p, 0, (size_t(nV) + 1) * sizeof(int), s));
        ctx->launches += 2;
    }
    if (counting) { // the placed incidences of every active column into collision order
        Stage kt(ctx, "k:sort_columns(incidences)", s);
        ctx->hbig.reserve(size_t(nV) + 1);
        unsigned long long* nbig = ctx->dCounters.p + 27;
        IPCB_CUDA(cudaMemsetAsync(nbig, 0, sizeof(unsigned long long), s));
        int cs_cap = COLSORT_WARP_CAP; // test hook: small scenes reach the block-per-column sort
        if (const char* e = getenv("IPCB_HESS_COLSORT_WARP_CAP")) cs_cap = std::min(COLSORT_WARP_CAP, std::max(1, atoi(e)));
        k_sort_columns<<<grid_for(size_t(nV), SYM_WARPS), 32 * SYM_WARPS, 0, s>>>(ctx->hactive.p, nactive, ctx->hcolinc.p, ctx->hkey_sorted.p,
                                                                                ctx->hbig.p, nbig, cs_cap);
        if (!ctx->colsort_attr_set) {
            IPCB_CUDA(cudaFuncSetAttribute(k_sort_columns_big, cudaFuncAttributeMaxDynamicSharedMemorySize, COLSORT_CTA_CAP * 8));
            ctx->colsort_attr_set = true;
        }
        k_sort_columns_big<<<NUM_SMS, BIG_THREADS, COLSORT_CTA_CAP * 8, s>>>(ctx->hbig.p, nbig, ctx->hcolinc.p, ctx->hkey_sorted.p, colsort_cta_cap());
        ctx->launches += 2;
    }
    (void)nitems, (void)blk0;
}

bool hessian_assemble_finish(ipcb_ctx* ctx, const int64_t nk[4])
{
    cudaStream_t s = ctx->stream;
    const int nV = ctx->nV;
    const int64_t n0 = nk[0], n1 = nk[1], n2 = nk[2], n3 = nk[3];
    const int64_t gi0[4] = { 0, n0, n0 + n1, n0 + n1 + n2 };
    const int64_t nitems = 4 * n0 + 9 * n1 + 16 * (n2 + n3);
    int64_t blk0[4];
    hess_block_offsets(nk, blk0);
    const unsigned ref_ev = unsigned(gi0[1] * 4), ref_ee = unsigned(gi0[2] * 4);
    int* nactive = reinterpret_cast<int*>(ctx->dCounters.p + 25);
    size_t b3 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b3, ctx->hcnt.p, ctx->outer.p, 3 * nV + 1, s);
    std::unique_ptr<Stage> st(new Stage(ctx, "hess_symbolic"));
    // 3. pass 1: per-column sort by row vertex, pattern counts
    ctx->hsref.reserve(nitems);
    ctx->hbig.reserve(size_t(nV) + 1);
    unsigned long long* nbig = ctx->dCounters.p + 6;
    unsigned long long* need = ctx->dCounters.p + 7;
    ctx->hudesc.reserve(nitems), ctx->hcolU.reserve(size_t(nV) + 1);
    const SymArgs A { nV, ctx->hkey_sorted.p, ctx->hcolinc.p, ctx->hcolR.p, ctx->hitemoff.p, ctx->hcolb.p, ctx->hvid.p, ctx->hmask.p, ref_ev, ref_ee,
                      { unsigned(blk0[0]), unsigned(blk0[1]), unsigned(blk0[2]) }, ctx->hsref.p, ctx->hudesc.p, ctx->hcolU.p, ctx->hcnt.p };
    if (!ctx->hess_attr_set) { // per device
        IPCB_CUDA(cudaFuncSetAttribute(k_hess_symbolic_big, cudaFuncAttributeMaxDynamicSharedMemorySize, int(BIG_SMEM)));
        IPCB_CUDA(cudaFuncSetAttribute(k_hess_numeric_big, cudaFuncAttributeMaxDynamicSharedMemorySize, int(NUMERIC_UCAP) * (72 + 2)));
        ctx->hess_attr_set = true;
    }
    const int big_grid = NUM_SMS;
    // test hooks: lower the hand-over thresholds so that small scenes exercise the block / global-scratch paths
    int warp_cap = WARP_CAP, cta_cap = CTA_CAP;
    if (const char* e = getenv("IPCB_HESS_WARP_CAP")) warp_cap = std::min(WARP_CAP, std::max(1, atoi(e)));
    if (const char* e = getenv("IPCB_HESS_CTA_CAP")) cta_cap = std::min(CTA_CAP, std::max(1, atoi(e)));
    const int use_hash = getenv("IPCB_HESS_NO_HASH") ? 0 : 1;
    // block-wide hash for columns beyond the shared-memory sort; IPCB_HESS_BIG_HASH=0 disables it (the global-scratch sort then
    // takes those columns), =2 sends every column of the block kernel through it (test hook)
    int big_hash = use_hash;
    if (const char* e = getenv("IPCB_HESS_BIG_HASH")) big_hash = atoi(e);
    for (int attempt = 0;; attempt++) {
        IPCB_CUDA(cudaMemsetAsync(nbig, 0, 2 * sizeof(unsigned long long), s));
        k_hess_symbolic<<<grid_for(size_t(nV) + 1, SYM_WARPS), 32 * SYM_WARPS, 0, s>>>(A, warp_cap, use_hash, ctx->hbig.p, nbig, ctx->hactive.p, nactive);
        k_hess_symbolic_big<<<big_grid, BIG_THREADS, BIG_SMEM, s>>>(A, cta_cap, big_hash, ctx->hbig.p, nbig, ctx->hscratch.p, ctx->hscratch_items,
                                                                    need);
        ctx->launches += 2;
        // 4. scalar column pointers, nnz
        cub::DeviceScan::ExclusiveSum(ctx->cubtmp.p, b3, ctx->hcnt.p, ctx->outer.p, 3 * nV + 1, s);
        ctx->launches += 2;
        IPCB_CUDA(cudaMemcpyAsync(&ctx->pinned.p[9], ctx->outer.p + 3 * size_t(nV), sizeof(int), cudaMemcpyDeviceToHost, s));
        IPCB_CUDA(cudaMemcpyAsync(&ctx->pinned.p[10], need, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        if (ctx->hess_counting_used)
            IPCB_CUDA(cudaMemcpyAsync(&ctx->pinned.p[12], ctx->dCounters.p + 26, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        IPCB_CUDA(cudaStreamSynchronize(s));
        if (ctx->hess_counting_used && ctx->pinned.p[12] > (unsigned long long)colsort_cta_cap()) return false;
        const size_t want = size_t(ctx->pinned.p[10]);
        if (want == 0) break;
        if (attempt > 0) throw Error("Hessian: sort scratch for a huge column could not be provided");
        ctx->hscratch.reserve(want * 14 * size_t(big_grid)); // a vertex with more than CTA_CAP row blocks: sort in global memory
        ctx->hscratch_items = want;
    }
    ctx->nnz = *reinterpret_cast<int*>(&ctx->pinned.p[9]);
    ctx->inner.reserve(std::max<int64_t>(ctx->nnz, 1)), ctx->vals.reserve(std::max<int64_t>(ctx->nnz, 1));
    st.reset();
    st.reset(new Stage(ctx, "hess_numeric"));
    // 5. pass 2: gather, run-sum, write compressed columns
    const char* nb_env = getenv("IPCB_NUM_BATCH");
    const int nb = nb_env ? atoi(nb_env) : 12; // measured on C3: 2.43 ms (8), 2.24 ms (12)
    const unsigned ngrid = grid_for(nV, SYM_WARPS);
    // columns with more items than this go to the block-per-column kernel (test hook: IPCB_HESS_NUMERIC_BIG lowers it)
    int big_items = 4096;
    if (const char* e = getenv("IPCB_HESS_NUMERIC_BIG")) big_items = std::max(1, atoi(e));
    int ucap = NUMERIC_UCAP; // unique blocks per shared-memory segment (test hook: IPCB_HESS_NUMERIC_UCAP)
    if (const char* e = getenv("IPCB_HESS_NUMERIC_UCAP")) ucap = std::min(NUMERIC_UCAP, std::max(3, atoi(e)));
    unsigned long long* nbig2 = ctx->dCounters.p + 15;
    IPCB_CUDA(cudaMemsetAsync(nbig2, 0, sizeof(unsigned long long), s));
#define IPCB_NUMERIC(KERNEL, ...)                                                                                                          \
    KERNEL<__VA_ARGS__><<<ngrid, 32 * SYM_WARPS, 0, s>>>(nV, ctx->hcolR.p, ctx->hcolU.p, ctx->hitemoff.p, ctx->hsref.p, ctx->hudesc.p, ctx->hblk.p, \
                                                         ctx->outer.p, ctx->inner.p, ctx->vals.p, ctx->hactive.p, nactive, big_items, ctx->hbig.p, nbig2)
    // IPCB_NUMERIC_LANES=9: one lane per block entry (three blocks per round); =3: one lane per block column (ten blocks per round)
    // IPCB_NUM_REM: the remainder of a long run: 0 (default) shared by the three lane groups, n > 0: by its owner, n gathers at a
    // time.  Measured on C3 (ms, NUM_BATCH 12): 0: 2.34, 1: 2.43, 4: 2.75, 8: 4.44; shorter first batches lose (0 / 8: 2.62, 0 / 6: 2.83)
    const char* nl_env = getenv("IPCB_NUMERIC_LANES");
    const int nlanes = nl_env ? atoi(nl_env) : 9;
    const int rem = getenv("IPCB_NUM_REM") ? atoi(getenv("IPCB_NUM_REM")) : 0;
    if (nlanes == 3) {
        const int nb3 = nb_env ? atoi(nb_env) : 6;
        if (nb3 >= 8) IPCB_NUMERIC(k_hess_numeric_col, 8);
        else if (nb3 >= 6) IPCB_NUMERIC(k_hess_numeric_col, 6);
        else IPCB_NUMERIC(k_hess_numeric_col, 4);
    } else if (nb == 16) IPCB_NUMERIC(k_hess_numeric, 16, 1);
    else if (nb == 8 && rem == 0) IPCB_NUMERIC(k_hess_numeric, 8, 0);
    else if (nb == 6 && rem == 0) IPCB_NUMERIC(k_hess_numeric, 6, 0);
    else if (rem == 0) {
        const int cu = getenv("IPCB_NUM_COOP") ? atoi(getenv("IPCB_NUM_COOP")) : 1; // gathers in flight in the cooperative remainder
        if (cu >= 4) IPCB_NUMERIC(k_hess_numeric, 12, 0, 4);
        else if (cu >= 2) IPCB_NUMERIC(k_hess_numeric, 12, 0, 2);
        else IPCB_NUMERIC(k_hess_numeric, 12, 0, 1);
    }
    else if (nb == 8) IPCB_NUMERIC(k_hess_numeric, 8, 1);
    else if (rem >= 8) IPCB_NUMERIC(k_hess_numeric, 12, 8);
    else if (rem >= 4) IPCB_NUMERIC(k_hess_numeric, 12, 4);
    else IPCB_NUMERIC(k_hess_numeric, 12, 1);
#undef IPCB_NUMERIC
    constexpr size_t NUMERIC_SMEM = size_t(NUMERIC_UCAP) * (72 + 2);
    k_hess_numeric_big<<<NUM_SMS, NBIG_THREADS, NUMERIC_SMEM, s>>>(ctx->hbig.p, nbig2, ctx->hcolR.p, ctx->hcolU.p, ctx->hitemoff.p, ctx->hsref.p,
                                                                 ctx->hudesc.p, ctx->hblk.p, ctx->outer.p, ctx->inner.p, ctx->vals.p, ucap);
    ctx->launches += 2;
    ctx->launches++;
    IPCB_CUDA(cudaGetLastError());
    return true;
}

void barrier_hessian(ipcb_ctx* ctx, const ipcb_barrier_params& bp, int psd_mode)
{
    cudaStream_t s = ctx->stream;
    const BarrierDev B = make_barrier(bp, ctx->dmin);
    const int nV = ctx->nV;
    const int v_lo = ctx->row_hi < 0 ? 0 : std::max(0, ctx->row_lo), v_hi = ctx->row_hi < 0 ? nV : std::min(nV, ctx->row_hi);
    const MeshView m = mesh_view(ctx);
    // n_k: records of kind k = all its collisions, or (row block of a sharded Hessian) the ones touching an owned vertex
    int64_t nk[4] = { ctx->coll[0].count, ctx->coll[1].count, ctx->coll[2].count, ctx->coll[3].count };
    const int* sel[4] = { nullptr, nullptr, nullptr, nullptr };
    ctx->outer.reserve(3 * size_t(nV) + 1);
    ctx->nnz = 0;
    if ((v_lo > 0 || v_hi < nV) && nk[0] + nk[1] + nk[2] + nk[3] > 0) {
        Stage st(ctx, "hess_select");
        const int64_t all = nk[0] + nk[1] + nk[2] + nk[3];
        ctx->hflag.reserve(all), ctx->hsel.reserve(all);
        cudaStream_t where[4] = { ctx->aux[0], ctx->aux[1], s, ctx->aux[2] };
        unsigned long long* d_cnt = ctx->dCounters.p + 24;
        IPCB_CUDA(cudaMemsetAsync(d_cnt, 0, 4 * sizeof(unsigned long long), s));
        ctx->fork();
        int64_t off = 0;
        for (int k = 0; k < 4; k++) {
            if (nk[k] == 0) continue;
            unsigned char* flag = ctx->hflag.p + off;
            int* list = ctx->hsel.p + off;
            sel[k] = list;
            off += nk[k];
            k_touch_flags<<<grid_for(nk[k], 256), 256, 0, where[k]>>>(view(ctx, k), m, v_lo, v_hi, flag);
            size_t bytes = 0;
            cub::CountingInputIterator<int> iota(0);
            int* d_num = reinterpret_cast<int*>(d_cnt + k);
            cub::DeviceSelect::Flagged(nullptr, bytes, iota, flag, list, d_num, int(nk[k]), where[k]);
            ctx->coll[k].cubtmp.reserve(bytes + 256);
            cub::DeviceSelect::Flagged(ctx->coll[k].cubtmp.p, bytes, iota, flag, list, d_num, int(nk[k]), where[k]);
            ctx->launches += 3;
        }
        for (int k = 0; k < ipcb_ctx::NAUX; k++) ctx->join(k);
        IPCB_CUDA(cudaGetLastError());
        IPCB_CUDA(cudaMemcpyAsync(&ctx->pinned.p[28], d_cnt, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        IPCB_CUDA(cudaStreamSynchronize(s));
        for (int k = 0; k < 4; k++) nk[k] = nk[k] ? int64_t(ctx->pinned.p[28 + k] & 0xffffffffll) : 0;
    }
    const int64_t n0 = nk[0], n1 = nk[1], n2 = nk[2], n3 = nk[3];
    const int64_t gi0[4] = { 0, n0, n0 + n1, n0 + n1 + n2 };
    const int64_t inc0[4] = { 0, 2 * n0, 2 * n0 + 3 * n1, 2 * n0 + 3 * n1 + 4 * n2 };
    if (n0 + n1 + n2 + n3 == 0) { // empty ndof x ndof matrix (potential.cpp:107-109)
        hessian_empty(ctx);
        return;
    }
    {
        Stage st(ctx, "hessian_local");
        HessOut outs[4];
        hessian_records(ctx, nk, v_lo, v_hi, outs);
        // The incidences only depend on the ids: they could be written and grouped by vertex (radix sort, column ranges, active
        // columns) on a side stream WHILE the local-Hessian kernels compute the blocks (IPCB_HESS_OVERLAP).  Measured on C3:
        // no gain (16.29 vs 16.30 ms) — k_hessian_fast holds every SM's register file (4 blocks x 128 threads x 128 registers),
        // so the side stream's kernels wait for its blocks to retire anyway.  Off by default.
        static const bool overlap = getenv("IPCB_HESS_OVERLAP") != nullptr;
        if (overlap) {
            ctx->fork();
            cudaStream_t side = ctx->aux[2];
            if (n0) k_write_records<IPCB_VV><<<grid_for(n0, 256), 256, 0, side>>>(view(ctx, 0), m, gi0[0], inc0[0], outs[0], sel[0], n0);
            if (n1) k_write_records<IPCB_EV><<<grid_for(n1, 256), 256, 0, side>>>(view(ctx, 1), m, gi0[1], inc0[1], outs[1], sel[1], n1);
            if (n2) k_write_records<IPCB_EE><<<grid_for(n2, 256), 256, 0, side>>>(view(ctx, 2), m, gi0[2], inc0[2], outs[2], sel[2], n2);
            if (n3) k_write_records<IPCB_FV><<<grid_for(n3, 256), 256, 0, side>>>(view(ctx, 3), m, gi0[3], inc0[3], outs[3], sel[3], n3);
            ctx->launches += 4;
            hessian_assemble_prepare(ctx, nk, side);
            for (int k = 0; k < 4; k++) outs[k].records_done = 1;
        }
        static const bool force_general = getenv("IPCB_HESSIAN_GENERAL") != nullptr; // A/B switch for tests and profiles
        if (psd_mode == IPCB_PSD_NONE || force_general) {
            if (n0) k_hessian_local<IPCB_VV><<<grid_for(n0, 128), 128, 0, s>>>(view(ctx, 0), m, B, psd_mode, gi0[0], inc0[0], outs[0], nullptr, 0, sel[0], n0), ctx->launches++;
            if (n1) k_hessian_local<IPCB_EV><<<grid_for(n1, 128), 128, 0, s>>>(view(ctx, 1), m, B, psd_mode, gi0[1], inc0[1], outs[1], nullptr, 0, sel[1], n1), ctx->launches++;
            if (n2) k_hessian_local<IPCB_EE><<<grid_for(n2, 128), 128, 0, s>>>(view(ctx, 2), m, B, psd_mode, gi0[2], inc0[2], outs[2], nullptr, 0, sel[2], n2), ctx->launches++;
            if (n3) k_hessian_local<IPCB_FV><<<grid_for(n3, 128), 128, 0, s>>>(view(ctx, 3), m, B, psd_mode, gi0[3], inc0[3], outs[3], nullptr, 0, sel[3], n3), ctx->launches++;
        } else {
            unsigned long long* slow_count = ctx->dCounters.p + 5;
            ctx->hslow.reserve(std::max<int64_t>(n2, 1));
            IPCB_CUDA(cudaMemsetAsync(slow_count, 0, sizeof(unsigned long long), s));
            // the four kinds write disjoint records: the small ones run beside the edge-edge kernel
            ctx->fork();
            static const bool dense = getenv("IPCB_HFAST_SPARSE") == nullptr; // 4 resident blocks per SM (small spill) for the 4-point kinds; A/B switch
            std::unique_ptr<Stage> kt(new Stage(ctx, "k:k_hessian_fast<VV>", ctx->aux[0]));
            if (n0) k_hessian_fast<IPCB_VV, 4, false><<<grid_for(n0, 128), 128, 0, ctx->aux[0]>>>(view(ctx, 0), m, B, psd_mode, gi0[0], inc0[0], outs[0], ctx->hslow.p, slow_count, sel[0], n0), ctx->launches++;
            kt.reset(), kt.reset(new Stage(ctx, "k:k_hessian_fast<EV>", ctx->aux[0]));
            // IPCB_HFAST_STAGED: the warp-wide staged copy instead of the TMA bulk stores (A/B switch, tests)
            static const bool bulk = getenv("IPCB_HFAST_STAGED") == nullptr;
            constexpr size_t BULK_SMEM = size_t(128) * BULK_STRIDE * sizeof(double);
            if (bulk && !ctx->hfast_attr_set) { // per device
                IPCB_CUDA(cudaFuncSetAttribute(k_hessian_fast<IPCB_EV, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(BULK_SMEM)));
                IPCB_CUDA(cudaFuncSetAttribute(k_hessian_fast<IPCB_FV, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(BULK_SMEM)));
                IPCB_CUDA(cudaFuncSetAttribute(k_hessian_fast<IPCB_EE, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(BULK_SMEM)));
                ctx->hfast_attr_set = true;
            }
#define IPCB_HFAST(KIND, MINB, BULK, K, STREAM)                                                                                            \
    k_hessian_fast<KIND, MINB, BULK><<<grid_for(nk[K], 128), 128, BULK ? BULK_SMEM : 0, STREAM>>>(view(ctx, K), m, B, psd_mode, gi0[K], inc0[K], \
                                                                                                  outs[K], ctx->hslow.p, slow_count, sel[K], nk[K])
            if (n1) {
                if (bulk) IPCB_HFAST(IPCB_EV, 4, true, 1, ctx->aux[0]);
                else IPCB_HFAST(IPCB_EV, 4, false, 1, ctx->aux[0]);
                ctx->launches++;
            }
            kt.reset(), kt.reset(new Stage(ctx, "k:k_hessian_fast<FV>", ctx->aux[1]));
            if (n3) {
                if (bulk) IPCB_HFAST(IPCB_FV, 4, true, 3, ctx->aux[1]);
                else if (dense) IPCB_HFAST(IPCB_FV, 4, false, 3, ctx->aux[1]);
                else IPCB_HFAST(IPCB_FV, 3, false, 3, ctx->aux[1]);
                ctx->launches++;
            }
            kt.reset(), kt.reset(new Stage(ctx, "k:k_hessian_fast<EE>", s));
            if (n2) {
                if (bulk) IPCB_HFAST(IPCB_EE, 4, true, 2, s);
                else if (dense) IPCB_HFAST(IPCB_EE, 4, false, 2, s);
                else IPCB_HFAST(IPCB_EE, 3, false, 2, s);
#undef IPCB_HFAST
                ctx->launches++;
                kt.reset();
                IPCB_CUDA(cudaMemcpyAsync(&ctx->pinned.p[11], slow_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
                IPCB_CUDA(cudaStreamSynchronize(s));
                const int64_t nslow = ctx->pinned.p[11];
                if (nslow) {
                    Stage ks(ctx, "k:k_hessian_local<EE>(mollified)", s);
                    // a short list (a few thousand mollified pairs): one warp per block spreads it over every SM's FP64 pipe
                    const int bs = nslow < int64_t(NUM_SMS) * 512 ? 32 : 128;
                    k_hessian_local<IPCB_EE><<<grid_for(nslow, bs), bs, 0, s>>>(view(ctx, 2), m, B, psd_mode, gi0[2], inc0[2], outs[2], ctx->hslow.p, nslow, sel[2], n2);
                    ctx->launches++;
                }
            }
            kt.reset();
            ctx->join(0);
            ctx->join(1);
        }
        IPCB_CUDA(cudaGetLastError());
        if (overlap) ctx->join(2);
        else hessian_assemble_prepare(ctx, nk, s);
    }
    if (!hessian_assemble_finish(ctx, nk)) {
        hessian_assemble_prepare(ctx, nk, s, true);
        hessian_assemble_finish(ctx, nk);
    }
}

} // namespace ipcb
