// ccd.cu — narrow-phase CCD and the collision-free step size.
//
// Replaces (reference src/ipc/): candidates/candidates.cpp:252-292
// (Candidates::compute_collision_free_stepsize with its shared earliest-TOI
// pruning), ccd/tight_inclusion_ccd.cpp:33-336 (strategy wrapper),
// ccd/additive_ccd.cpp:71-325, ccd/check_initial_distance.hpp.  The
// Tight-Inclusion root finder itself (third-party ticcd v1.0.6, not in the
// reference tree) is re-designed for the GPU from the published algorithm:
//
//   * one thread per candidate gathers its 8 points (one 32-byte sector per
//     vertex), evaluates the initial distance, the domain tolerances and the
//     floating-point filter, and tests the root box [0,1]^3 — most candidates
//     end here;
//   * survivors refine IN REGISTERS: an earliest-time-first depth-first
//     subdivision over dyadic (t,u,v) boxes with a small per-thread stack.  The
//     result is the minimum lower time bound over all terminal boxes, which is
//     what the sequential breadth-first search of the library returns (it
//     returns the first terminal box in (level, t) order);
//   * like the reference's `std::atomic<double> earliest_toi`
//     (candidates.cpp:267-286) a global bound holds the minimum of the FINAL
//     times of impact published so far; every thread re-reads it while refining,
//     so late queries only search [0, bound];
//   * a query that exceeds its in-register unit budget spills its stack into a
//     GLOBAL interval-subdivision queue that is processed level by level by all
//     SMs (warp-aggregated atomic pushes), so a pathological query is refined by
//     thousands of threads instead of stalling one warp;
//   * queries whose TOI is below SMALL_TOI are re-run with the reference's
//     no-zero-TOI refinement (tight_inclusion_ccd.cpp:59-72 + ticcd's loop).
#include "ctx.cuh"
#include "geom.cuh"

namespace ipcb {

ipcb_ccd_params resolve_ccd(const ipcb_ccd_params* p)
{
    ipcb_ccd_params r = p ? *p : ipcb_ccd_params { IPCB_CCD_TIGHT_INCLUSION, 0, 0, 0 };
    if (r.kind == IPCB_CCD_ADDITIVE) {
        if (r.max_iterations == 0) r.max_iterations = 10'000'000;
        if (r.conservative_rescaling <= 0) r.conservative_rescaling = 0.9;
    } else {
        if (r.tolerance <= 0) r.tolerance = 1e-6;
        if (r.max_iterations == 0) r.max_iterations = 10'000'000;
        if (r.conservative_rescaling <= 0) r.conservative_rescaling = 0.8;
    }
    return r;
}

// where queries come from: resident candidates + positions, or raw 12-double stencils
struct QuerySource {
    int kind;
    int64_t n;
    const int2* cand;
    const int2* E;
    const int4* F;
    const double4* X0;
    const double4* X1;
    const double* raw0;
    const double* raw1;
};
// several kinds behind one concatenated index space [off[k], off[k+1]) (the step-size search looks at the
// face-vertex and edge-edge candidates in the same launches)
struct MultiSource {
    int nk;
    int kind[4];
    long long off[5];
    const int2* cand[4];
    const int2* E;
    const int4* F;
    const double4* X0;
    const double4* X1;
    const double* raw0;
    const double* raw1;
};
__device__ inline QuerySource locate(const MultiSource& m, long long g, int64_t& i)
{
    int k = 0;
#pragma unroll
    for (int j = 1; j < 4; j++)
        if (j < m.nk && g >= m.off[j]) k = j;
    i = g - m.off[k];
    return QuerySource { m.kind[k], m.off[k + 1] - m.off[k], m.cand[k], m.E, m.F, m.X0, m.X1, m.raw0, m.raw1 };
}

// stencil at t0 / t1; returns number of points (2, 3, 4)
__device__ inline int load_query(const QuerySource& q, int64_t i, d3* a, d3* b)
{
    const int n = q.kind == IPCB_VV ? 2 : (q.kind == IPCB_EV ? 3 : 4);
    if (q.cand) {
        const int2 c = q.cand[i];
        int vid[4];
        if (q.kind == IPCB_VV) {
            vid[0] = c.x, vid[1] = c.y;
        } else if (q.kind == IPCB_EV) {
            const int2 e = __ldg(q.E + c.x);
            vid[0] = c.y, vid[1] = e.x, vid[2] = e.y;
        } else if (q.kind == IPCB_EE) {
            const int2 ea = __ldg(q.E + c.x), eb = __ldg(q.E + c.y);
            vid[0] = ea.x, vid[1] = ea.y, vid[2] = eb.x, vid[3] = eb.y;
        } else {
            const int4 f = __ldg(q.F + c.x);
            vid[0] = c.y, vid[1] = f.x, vid[2] = f.y, vid[3] = f.z;
        }
        for (int k = 0; k < n; k++) a[k] = load_vertex(q.X0, vid[k]), b[k] = load_vertex(q.X1, vid[k]);
    } else {
        for (int k = 0; k < n; k++) {
            a[k] = { q.raw0[12 * i + 3 * k], q.raw0[12 * i + 3 * k + 1], q.raw0[12 * i + 3 * k + 2] };
            b[k] = { q.raw1[12 * i + 3 * k], q.raw1[12 * i + 3 * k + 1], q.raw1[12 * i + 3 * k + 2] };
        }
    }
    return n;
}
// squared distance with automatic distance type
__device__ inline double auto_distance(int kind, const d3* x)
{
    if (kind == IPCB_VV) return pp_dist(x[0], x[1]);
    if (kind == IPCB_EV) return sub_value(sub_point_edge(point_edge_type(x[0], x[1], x[2])), x);
    if (kind == IPCB_EE) return sub_value(sub_edge_edge(edge_edge_type(x[0], x[1], x[2], x[3])), x);
    return sub_value(sub_point_triangle(point_triangle_type(x[0], x[1], x[2], x[3])), x);
}

__device__ inline void atomic_min_double(unsigned long long* addr, double v) // v >= 0
{
    atomicMin(addr, (unsigned long long)__double_as_longlong(v));
}
__host__ __device__ inline double load_bound(const unsigned long long* addr)
{
    const unsigned long long bits = *reinterpret_cast<const volatile unsigned long long*>(addr);
    double v;
    memcpy(&v, &bits, sizeof v);
    return v;
}

struct CcdOut {
    unsigned long long* bound; // global earliest FINAL time of impact (bit pattern), or nullptr
    unsigned char* hit;        // per query, or nullptr
    double* toi;               // per query, or nullptr
};
__device__ inline void report(const CcdOut& o, int64_t i, bool hit, double toi)
{
    if (o.hit) {
        o.hit[i] = hit;
        o.toi[i] = hit ? toi : INFINITY;
    }
    if (o.bound && hit) atomic_min_double(o.bound, toi);
}

// ===========================================================================
// Additive CCD: ccd/additive_ccd.cpp:71-325
__global__ void __launch_bounds__(128)
    k_additive(QuerySource q, double min_distance, double tmax_in, long long max_iterations, double rescale, CcdOut out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= q.n) return;
    d3 x[4], x1[4];
    const int n = load_query(q, i, x, x1);
    const double ms2 = min_distance * min_distance;
    if (auto_distance(q.kind, x) <= ms2) { // initial distance <= d_min: toi = 0 (:139-145 etc.)
        report(out, i, true, 0.0);
        return;
    }
    d3 dx[4];
    d3 mean = { 0, 0, 0 };
    for (int k = 0; k < n; k++) {
        dx[k] = x1[k] - x[k];
        mean = mean + dx[k];
    }
    mean = { mean.x / n, mean.y / n, mean.z / n };
    for (int k = 0; k < n; k++) dx[k] = dx[k] - mean;
    double lp;
    if (q.kind == IPCB_VV) lp = sqrt(sqn(dx[0])) + sqrt(sqn(dx[1]));
    else if (q.kind == IPCB_EV) lp = sqrt(sqn(dx[0])) + sqrt(fmax(sqn(dx[1]), sqn(dx[2])));
    else if (q.kind == IPCB_FV) lp = sqrt(sqn(dx[0])) + sqrt(fmax(fmax(sqn(dx[1]), sqn(dx[2])), sqn(dx[3])));
    else lp = sqrt(fmax(sqn(dx[0]), sqn(dx[1]))) + sqrt(fmax(sqn(dx[2]), sqn(dx[3])));
    if (lp == 0) {
        report(out, i, false, 0);
        return;
    }
    auto dist = [&](const d3* y) {
        double d = auto_distance(q.kind, y);
        if (q.kind == IPCB_EE && d - ms2 <= 0) // far away nearly parallel edges (:303-318)
            d = fmin(fmin(sqn(y[0] - y[2]), sqn(y[0] - y[3])), fmin(sqn(y[1] - y[2]), sqn(y[1] - y[3])));
        return d;
    };
    double tmax = out.bound ? fmin(tmax_in, load_bound(out.bound)) : tmax_in;
    double d_sq = dist(x), d = sqrt(d_sq);
    double d_func = d_sq - ms2;
    const double gap = (1 - rescale) * d_func / (d + min_distance);
    double toi = 0;
    bool hit = true;
    for (long long it = 0; max_iterations < 0 || it < max_iterations; ++it) {
        const double lower = rescale * d_func / ((d + min_distance) * lp);
        for (int k = 0; k < n; k++) x[k] = x[k] + lower * dx[k];
        d_sq = dist(x);
        d = sqrt(d_sq);
        d_func = d_sq - ms2;
        if (toi > 0 && d_func / (d + min_distance) < gap) break;
        toi += lower;
        if (toi > tmax) {
            hit = false;
            break;
        }
        if (out.bound && (it & 15) == 15) tmax = fmin(tmax, load_bound(out.bound)); // fresher shared bound: same result, less work
    }
    report(out, i, hit, toi);
}

// ===========================================================================
// Tight Inclusion
constexpr double SMALL_TOI = 1e-6; // tight_inclusion_ccd.hpp:19

struct TIParams { // parameters of one root-finder run
    double tol[3]; // domain tolerances (t, u, v)
    double err[3]; // floating-point filter per coordinate
    double ms;     // minimum separation
    double co_tol; // co-domain tolerance (delta)
    double tmax;
};
struct TIQuery { // a spilled query (global queue)
    double s[12], e[12];
    TIParams P;
    int is_vf;
    int pad;
    long long src;
};
struct TIBox { // dyadic box: [n / 2^k, (n+1) / 2^k] per dimension
    unsigned tn, un, vn;
    unsigned char tk, uk, vk, pad;
};
struct TIUnit {
    int q;
    TIBox b;
};

__host__ __device__ inline double linf3(d3 a) { return fmax(fmax(fabs(a.x), fabs(a.y)), fabs(a.z)); }
__host__ __device__ inline d3 getp(const double* p, int k) { return { p[3 * k], p[3 * k + 1], p[3 * k + 2] }; }

// domain tolerances from the co-domain tolerance: delta / (3 * max edge length of the swept primitive)
__host__ __device__ inline void ti_tolerances(const double* s, const double* e, int is_vf, double co_tol, double* tol)
{
    d3 p000, p001, p011, p010, p100, p101, p111, p110;
    const d3 s0 = getp(s, 0), s1 = getp(s, 1), s2 = getp(s, 2), s3 = getp(s, 3);
    const d3 e0 = getp(e, 0), e1 = getp(e, 1), e2 = getp(e, 2), e3 = getp(e, 3);
    if (is_vf) {
        p000 = s0 - s1, p001 = s0 - s3, p011 = s0 - (s2 + s3 - s1), p010 = s0 - s2;
        p100 = e0 - e1, p101 = e0 - e3, p111 = e0 - (e2 + e3 - e1), p110 = e0 - e2;
    } else {
        p000 = s0 - s2, p001 = s0 - s3, p011 = s1 - s3, p010 = s1 - s2;
        p100 = e0 - e2, p101 = e0 - e3, p111 = e1 - e3, p110 = e1 - e2;
    }
    const double dl = 3 * fmax(fmax(linf3(p100 - p000), linf3(p101 - p001)), fmax(linf3(p111 - p011), linf3(p110 - p010)));
    const double e0l = 3 * fmax(fmax(linf3(p010 - p000), linf3(p110 - p100)), fmax(linf3(p111 - p101), linf3(p011 - p001)));
    const double e1l = 3 * fmax(fmax(linf3(p001 - p000), linf3(p101 - p100)), fmax(linf3(p111 - p110), linf3(p011 - p010)));
    tol[0] = co_tol / dl;
    tol[1] = co_tol / e0l;
    tol[2] = co_tol / e1l;
}
__host__ __device__ inline void ti_error(const double* s, const double* e, int is_vf, bool using_ms, double* err)
{
    const double filter = using_ms ? (is_vf ? 7.549516567451064e-15 : 7.105427357601002e-15)
                                   : (is_vf ? 6.661338147750939e-15 : 6.217248937900877e-15);
    for (int c = 0; c < 3; c++) {
        double mx = 0;
        for (int k = 0; k < 4; k++) mx = fmax(mx, fmax(fabs(s[3 * k + c]), fabs(e[3 * k + c])));
        const double delta = fmax(mx, 1.0);
        err[c] = filter * delta * delta * delta;
    }
}
// co-domain box of the (multilinear) root function over a (t,u,v) box against the eps-cube
// co-domain box of the (multilinear) root function over a (t,u,v) box against the eps-cube
__host__ __device__ inline bool ti_inclusion(const double* s, const double* e, int is_vf, const TIParams& P, const double* tt, const double* uu,
                                    const double* vv, bool& box_in, double* true_tol)
{
    box_in = true;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double vmin = INFINITY, vmax = -INFINITY;
#pragma unroll
        for (int a = 0; a < 2; a++) {
            const double t = tt[a];
            const double p0 = (e[c] - s[c]) * t + s[c];
            const double p1 = (e[3 + c] - s[3 + c]) * t + s[3 + c];
            const double p2 = (e[6 + c] - s[6 + c]) * t + s[6 + c];
            const double p3 = (e[9 + c] - s[9 + c]) * t + s[9 + c];
#pragma unroll
            for (int b = 0; b < 2; b++)
#pragma unroll
                for (int d = 0; d < 2; d++) {
                    double val;
                    if (is_vf) {
                        const double pt = (p2 - p1) * uu[b] + (p3 - p1) * vv[d] + p1;
                        val = p0 - pt;
                    } else {
                        const double va = (p1 - p0) * uu[b] + p0;
                        const double vb = (p3 - p2) * vv[d] + p2;
                        val = va - vb;
                    }
                    vmin = fmin(vmin, val);
                    vmax = fmax(vmax, val);
                }
        }
        const double lim = P.err[c] + P.ms;
        true_tol[c] = vmax - vmin;
        if (vmin > lim || vmax < -lim) return false;
        if (vmin < -lim || vmax > lim) box_in = false;
    }
    return true;
}

__host__ __device__ inline void box_bounds(const TIBox& b, double* tt, double* uu, double* vv)
{
    tt[0] = ldexp(double(b.tn), -int(b.tk)), tt[1] = ldexp(double(b.tn + 1), -int(b.tk));
    uu[0] = ldexp(double(b.un), -int(b.uk)), uu[1] = ldexp(double(b.un + 1), -int(b.uk));
    vv[0] = ldexp(double(b.vn), -int(b.vk)), vv[1] = ldexp(double(b.vn + 1), -int(b.vk));
}

// One subdivision step on a live box.  Returns 0: rejected, 1: terminal (its lower time bound is
// a time of impact), 2: split into nchild children (earliest first).
__host__ __device__ inline int ti_step(const double* s, const double* e, int is_vf, const TIParams& P, const TIBox& b, const double* tt,
                              const double* uu, const double* vv, TIBox* child, int& nchild)
{
    bool box_in;
    double true_tol[3];
    nchild = 0;
    if (!ti_inclusion(s, e, is_vf, P, tt, uu, vv, box_in, true_tol)) return 0;
    const double w[3] = { tt[1] - tt[0], uu[1] - uu[0], vv[1] - vv[0] };
    const bool tol_cond = true_tol[0] <= P.co_tol && true_tol[1] <= P.co_tol && true_tol[2] <= P.co_tol;
    const bool cond1 = w[0] <= P.tol[0] && w[1] <= P.tol[1] && w[2] <= P.tol[2];
    if (cond1 || tol_cond || box_in) return 1;
    int split = 0;
    double bestv = -INFINITY;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double r = w[k] > P.tol[k] ? w[k] / P.tol[k] : -INFINITY;
        if (r > bestv) bestv = r, split = k; // first maximum wins
    }
    const int kk = split == 0 ? b.tk : (split == 1 ? b.uk : b.vk);
    if (kk >= 30) return 1; // bisection depth exhausted: stop conservatively
    TIBox c0 = b, c1 = b;
    if (split == 0) {
        c0.tn = 2 * b.tn, c1.tn = 2 * b.tn + 1, c0.tk = c1.tk = b.tk + 1;
        child[nchild++] = c0; // the first half always overlaps [0, tmax]
        if (P.tmax == 1.0 || ldexp(double(c1.tn), -int(c1.tk)) <= P.tmax) child[nchild++] = c1;
    } else if (split == 1) {
        c0.un = 2 * b.un, c1.un = 2 * b.un + 1, c0.uk = c1.uk = b.uk + 1;
        child[nchild++] = c0;
        if (!is_vf || ldexp(double(c1.un), -int(c1.uk)) + vv[0] <= 1.0) child[nchild++] = c1; // u + v <= 1
    } else {
        c0.vn = 2 * b.vn, c1.vn = 2 * b.vn + 1, c0.vk = c1.vk = b.vk + 1;
        child[nchild++] = c0;
        if (!is_vf || ldexp(double(c1.vn), -int(c1.vk)) + uu[0] <= 1.0) child[nchild++] = c1;
    }
    return 2;
}

// separating direction test over the window [0, tmax] x [0,1]^2 with a uniform margin (see ti_cull (2)).
// n.F at the 8 corners of the window only needs the projections of the 4 points at t = 0 and t = tmax:
// edge-edge corners are a_i - b_j, vertex-face corners p - {t0, t1, t2, t1 + t2 - t0}.
__host__ __device__ inline bool ti_separated(const d3* p0, const d3* e4, int is_vf, double tmax, double margin)
{
    d3 axes[4];
    int na;
    if (is_vf) {
        const d3 n = cross(p0[2] - p0[1], p0[3] - p0[1]);
        axes[0] = n, axes[1] = cross(p0[2] - p0[1], n), axes[2] = cross(p0[3] - p0[2], n), axes[3] = cross(p0[1] - p0[3], n);
        na = 4;
    } else {
        const d3 ua = p0[1] - p0[0], ub = p0[3] - p0[2], n = cross(ua, ub);
        axes[0] = n, axes[1] = cross(ua, n), axes[2] = cross(ub, n);
        na = 3;
    }
    for (int a = 0; a < na; a++) {
        const d3 n = axes[a];
        const double l1 = (fabs(n.x) + fabs(n.y) + fabs(n.z)) * margin;
        double mn = INFINITY, mx = -INFINITY;
#pragma unroll
        for (int t = 0; t < 2; t++) {
            double pr[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const double d0 = dot(n, p0[k]);
                pr[k] = t == 0 ? d0 : (dot(n, e4[k]) - d0) * tmax + d0;
            }
            double lo, hi;
            if (is_vf) {
                const double q = pr[2] + pr[3] - pr[1];
                lo = pr[0] - fmax(fmax(pr[1], pr[2]), fmax(pr[3], q));
                hi = pr[0] - fmin(fmin(pr[1], pr[2]), fmin(pr[3], q));
            } else {
                lo = fmin(pr[0], pr[1]) - fmax(pr[2], pr[3]);
                hi = fmax(pr[0], pr[1]) - fmin(pr[2], pr[3]);
            }
            mn = fmin(mn, lo), mx = fmax(mx, hi);
        }
        if (mn > l1 || mx < -l1) return true;
    }
    return false;
}

// Level-0 culling, consistent with the root finder: returns true when NO box inside
// [0,tmax] x [0,1]^2 can ever be reported.
//  (1) the co-domain box of the root clipped to the search window misses the eps-cube (every
//      sub-box's co-domain box is contained in it);
//  (2) a separating direction n: n.F is multilinear in (t,u,v), so it is bounded by its 8 corner
//      values; if all of them exceed ||n||_1 (ms + err + delta) in magnitude with one sign, no
//      point of the window has |F_c| <= ms + err_c + delta for all c, and a terminal box (co-domain
//      width <= delta, or inside the cube) cannot exist.
__host__ __device__ inline bool ti_cull(const double* s, const double* e, int is_vf, const TIParams& P, double tmax)
{
    const double tt[2] = { 0.0, tmax }, unit[2] = { 0.0, 1.0 };
    bool box_in;
    double tw[3];
    if (!ti_inclusion(s, e, is_vf, P, tt, unit, unit, box_in, tw)) return true;
    d3 p0[4], e4[4];
#pragma unroll
    for (int k = 0; k < 4; k++) p0[k] = getp(s, k), e4[k] = getp(e, k);
    const double margin = (fmax(fmax(P.err[0], P.err[1]), P.err[2]) + P.ms + P.co_tol) * (1.0 + 1e-9);
    return ti_separated(p0, e4, is_vf, tmax, margin);
}

struct TIQueue {
    TIQuery* queries;
    unsigned long long* qtoi; // per spilled query: earliest terminal t so far (bit pattern)
    int* qflags;              // bit1: in the no-zero-toi refinement; bit2: finished
    unsigned long long* qcount; // boxes evaluated for the query in its current run (effort cap)
    unsigned long long* nq;
    unsigned long long qcap;
    TIUnit* in;
    TIUnit* outq;
    unsigned long long* nout;
    unsigned long long ucap;
};

constexpr int DFS_STACK = 96;

// Earliest-first depth-first refinement in registers.  Returns true when the search finished;
// false when the unit budget was exhausted (stack[0..sp) then holds the unexplored boxes).
__host__ __device__ inline bool ti_dfs(const double* s, const double* e, int is_vf, TIParams& P, const unsigned long long* bound, int budget,
                              TIBox* stack, int& sp, double& best)
{
    int iter = 0;
    while (sp > 0) {
        const TIBox b = stack[--sp];
        double tt[2], uu[2], vv[2];
        box_bounds(b, tt, uu, vv);
        if (tt[0] >= best) continue; // TOI_SKIP pruning of the sequential search
        if (bound && (iter & 3) == 0) P.tmax = fmin(P.tmax, load_bound(bound));
        if (tt[0] > P.tmax) continue;
        tt[1] = fmin(tt[1], fmin(best, P.tmax)); // only impacts before the best one so far (and inside the window) matter
        if (++iter > budget || sp + 2 > DFS_STACK) {
            stack[sp++] = b;
            return false;
        }
        TIBox child[2];
        int nchild;
        const int r = ti_step(s, e, is_vf, P, b, tt, uu, vv, child, nchild);
        if (r == 1) best = fmin(best, tt[0]);
        else if (r == 2) {
            if (nchild == 2) stack[sp++] = child[1];
            stack[sp++] = child[0]; // earliest / first half on top
        }
    }
    return true;
}

// warp-aggregated slot reservation (count slots per wanting lane)
__device__ inline unsigned long long warp_reserve(unsigned long long* counter, bool want, int count)
{
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (m == 0) return 0;
    const int lane = threadIdx.x & 31;
    const int mine = want ? count : 0;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(counter, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + (incl - mine);
}

// Pre-filter, one thread per candidate with few registers (high occupancy; the positions are L2
// resident): a candidate whose swept stencil is separated along one of the primitive's own
// directions by more than the LARGEST margin any root-finder run on it could use
//   err (floating-point filter, ms > 0 variant) + min_distance + 1e-4 (cap of the minimum effective
//   distance, tight_inclusion_ccd.cpp:48) + tolerance (>= the adjusted co-domain tolerance)
// can neither start closer than min_distance nor produce a terminal box (see ti_cull), so it is
// dropped here; survivors are compacted into `list` for the full per-query kernel.
__global__ void __launch_bounds__(256, 3)
    k_ti_filter(MultiSource ms, int stride, int prev_stride, double min_distance, double tmax_in, double tolerance, CcdOut out,
                int* __restrict__ list, unsigned long long* nlist, const float* __restrict__ scene)
{
    const unsigned long long* bound = out.bound;
    // thread t looks at candidate g = t * stride, unless an earlier (coarser) phase already did: g % prev_stride == 0
    const long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * stride;
    bool keep = false;
    if (g < ms.off[ms.nk] && !(prev_stride > 0 && g % prev_stride == 0)) {
        int64_t i;
        const QuerySource q = locate(ms, g, i);
        keep = true;
        if (q.kind >= IPCB_EE) { // point-point / point-edge queries are few (codimensional): always kept
            d3 a[4], b[4];
            load_query(q, i, a, b);
            // largest |coordinate| for the floating-point filter: the scene box of the swept broad phase bounds every
            // query's own maximum (a larger value only widens the margin: conservative); raw queries compute their own
            double mx = 1.0;
            if (scene) {
#pragma unroll
                for (int k = 0; k < 6; k++) mx = fmax(mx, fabs(double(__ldg(scene + k))));
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    mx = fmax(mx, fmax(fmax(fmax(fabs(a[k].x), fabs(a[k].y)), fabs(a[k].z)), fmax(fmax(fabs(b[k].x), fabs(b[k].y)), fabs(b[k].z))));
            }
            const int is_vf = q.kind == IPCB_FV;
            const double err = (is_vf ? 7.549516567451064e-15 : 7.105427357601002e-15) * mx * mx * mx;
            const double margin = (err + min_distance + 1e-4 + tolerance) * (1.0 + 1e-9);
            const double tmax = bound ? fmin(tmax_in, load_bound(bound)) : tmax_in;
            keep = !ti_separated(a, b, is_vf, tmax, margin);
            if (!keep) report(out, i, false, 0.0);
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (m) {
        const int lane = threadIdx.x & 31;
        unsigned long long base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(nlist, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (keep) list[base + __popc(m & ((1u << lane) - 1))] = int(g);
    }
}

// ---------------------------------------------------------------------------
// The same pre-filter in FP32 on positions re-centred at the scene centre and packed as ONE 32-byte sector per vertex
// (t0 | t1).  The separating-direction test holds for ANY direction n (n.F is multilinear, its extrema over the window
// are at the 8 corners), so only the projections have to be bounded: their FP32 error (input rounding + three products
// and sums per dot product, a few units of 2^-24 |n|_1 max|coordinate|) is added to the margin — 64 * 2^-23 * extent,
// orders of magnitude below the 1e-4 the margin already contains.  Half the gathered sectors, half the registers, the
// FP32 pipes instead of the FP64 one; survivors (~0.1 %) go through the exact FP64 set-up as before, so the answers
// do not change.
struct f3 {
    float x, y, z;
};
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
__device__ __forceinline__ float dotf(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 crossf(f3 a, f3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
__global__ void k_pack_f32(int n, const double4* __restrict__ X0, const double4* __restrict__ X1, const float* __restrict__ scene,
                           float4* __restrict__ XF)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double cx = 0.5 * (double(scene[0]) + double(scene[3])), cy = 0.5 * (double(scene[1]) + double(scene[4])),
                 cz = 0.5 * (double(scene[2]) + double(scene[5]));
    const double4 a = X0[i], b = X1[i];
    XF[2 * size_t(i)] = make_float4(float(a.x - cx), float(a.y - cy), float(a.z - cz), 0.f);
    XF[2 * size_t(i) + 1] = make_float4(float(b.x - cx), float(b.y - cy), float(b.z - cz), 0.f);
}
__device__ __forceinline__ bool ti_separated_f32(const f3* p0, const f3* e4, int is_vf, float tmax, float margin)
{
    f3 axes[4];
    int na;
    if (is_vf) {
        const f3 n = crossf(p0[2] - p0[1], p0[3] - p0[1]);
        axes[0] = n, axes[1] = crossf(p0[2] - p0[1], n), axes[2] = crossf(p0[3] - p0[2], n), axes[3] = crossf(p0[1] - p0[3], n);
        na = 4;
    } else {
        const f3 ua = p0[1] - p0[0], ub = p0[3] - p0[2], n = crossf(ua, ub);
        axes[0] = n, axes[1] = crossf(ua, n), axes[2] = crossf(ub, n);
        na = 3;
    }
    for (int a = 0; a < na; a++) {
        const f3 n = axes[a];
        const float l1 = (fabsf(n.x) + fabsf(n.y) + fabsf(n.z)) * margin;
        float mn = INFINITY, mx = -INFINITY;
#pragma unroll
        for (int t = 0; t < 2; t++) {
            float pr[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float d0 = dotf(n, p0[k]);
                pr[k] = t == 0 ? d0 : (dotf(n, e4[k]) - d0) * tmax + d0;
            }
            float lo, hi;
            if (is_vf) {
                const float q = pr[2] + pr[3] - pr[1];
                lo = pr[0] - fmaxf(fmaxf(pr[1], pr[2]), fmaxf(pr[3], q));
                hi = pr[0] - fminf(fminf(pr[1], pr[2]), fminf(pr[3], q));
            } else {
                lo = fminf(pr[0], pr[1]) - fmaxf(pr[2], pr[3]);
                hi = fmaxf(pr[0], pr[1]) - fminf(pr[2], pr[3]);
            }
            mn = fminf(mn, lo), mx = fmaxf(mx, hi);
        }
        if (mn > l1 || mx < -l1) return true;
    }
    return false;
}
__global__ void __launch_bounds__(256, 4)
    k_ti_filter32(MultiSource ms, int stride, double min_distance, double tmax_in, double tolerance, const unsigned long long* __restrict__ bound,
                  int* __restrict__ list, unsigned long long* nlist, const float* __restrict__ scene, const float4* __restrict__ XF)
{
    const long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * stride;
    bool keep = false;
    if (g < ms.off[ms.nk]) {
        int64_t i;
        const QuerySource q = locate(ms, g, i);
        keep = true;
        if (q.kind >= IPCB_EE) {
            const int2 c = q.cand[i];
            int vid[4];
            if (q.kind == IPCB_EE) {
                const int2 ea = __ldg(q.E + c.x), eb = __ldg(q.E + c.y);
                vid[0] = ea.x, vid[1] = ea.y, vid[2] = eb.x, vid[3] = eb.y;
            } else {
                const int4 f = __ldg(q.F + c.x);
                vid[0] = c.y, vid[1] = f.x, vid[2] = f.y, vid[3] = f.z;
            }
            f3 a[4], b[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float4 u = __ldg(XF + 2 * size_t(vid[k])), v = __ldg(XF + 2 * size_t(vid[k]) + 1);
                a[k] = { u.x, u.y, u.z }, b[k] = { v.x, v.y, v.z };
            }
            // margins: the FP64 filter's (floating-point filter of the root finder from the largest |coordinate| of the
            // scene, min_distance, the 1e-4 cap, the tolerance) plus the FP32 projection error for the re-centred extent
            double mxc = 1.0, ext = 0.0;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const double lo = double(__ldg(scene + k)), hi = double(__ldg(scene + 3 + k));
                mxc = fmax(mxc, fmax(fabs(lo), fabs(hi)));
                ext = fmax(ext, 0.5 * (hi - lo));
            }
            const int is_vf = q.kind == IPCB_FV;
            const double err = (is_vf ? 7.549516567451064e-15 : 7.105427357601002e-15) * mxc * mxc * mxc;
            const double margin = (err + min_distance + 1e-4 + tolerance) * (1.0 + 1e-6) + 64.0 * 1.1920929e-7 * ext;
            const double tmax = bound ? fmin(tmax_in, load_bound(bound)) : tmax_in;
            keep = !ti_separated_f32(a, b, is_vf, __double2float_ru(tmax), __double2float_ru(margin));
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (m) {
        const int lane = threadIdx.x & 31;
        unsigned long long base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(nlist, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (keep) list[base + __popc(m & ((1u << lane) - 1))] = int(g);
    }
}

// Query set-up shared by the per-thread and the per-warp search (tight_inclusion_ccd.cpp:222-336 and
// ccd_strategy :33-75).  Returns false when the query is already answered (and reported).
__device__ inline bool ti_setup(const QuerySource& q, int64_t i, double min_distance, double tmax_in, double tolerance, double rescale,
                                const CcdOut& out, double* s, double* e, TIParams& P, int& is_vf, double& tmax0)
{
    d3 a[4], b[4];
    const int n = load_query(q, i, a, b);
    const double d0 = sqrt(auto_distance(q.kind, a));
    bool moving = false;
    for (int k = 0; k < n; k++) moving |= !same(a[k], b[k]);
    if (d0 <= min_distance) { // check_initial_distance: toi = 0 (also the no-motion answer)
        report(out, i, true, 0.0);
        return false;
    }
    if (!moving) {
        report(out, i, false, 0.0);
        return false;
    }
    // points handed to the root finder: VV / EV are degenerate edge-edge queries (:104,:177)
    d3 s4[4], e4[4];
    if (q.kind == IPCB_VV) {
        s4[0] = s4[1] = a[0], s4[2] = s4[3] = a[1];
        e4[0] = e4[1] = b[0], e4[2] = e4[3] = b[1];
    } else if (q.kind == IPCB_EV) {
        s4[0] = s4[1] = a[0], s4[2] = a[1], s4[3] = a[2];
        e4[0] = e4[1] = b[0], e4[2] = b[1], e4[3] = b[2];
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) s4[k] = a[k], e4[k] = b[k];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        s[3 * k] = s4[k].x, s[3 * k + 1] = s4[k].y, s[3 * k + 2] = s4[k].z;
        e[3 * k] = e4[k].x, e[3 * k + 1] = e4[k].y, e[3 * k + 2] = e4[k].z;
    }
    is_vf = q.kind == IPCB_FV;
    // the search window is fixed when the query starts, like `tmax = earliest_toi.load()` (candidates.cpp:270)
    tmax0 = out.bound ? fmin(tmax_in, load_bound(out.bound)) : tmax_in;
    double med = (1.0 - rescale) * (d0 - min_distance); // minimum effective distance (:45-49)
    med = fmin(med, 1e-4);
    med += min_distance;
    P.ms = med;
    P.co_tol = fmin(0.5 * d0, tolerance); // adjusted tolerance (:245-246)
    P.tmax = tmax0;
    ti_tolerances(s, e, is_vf, P.co_tol, P.tol);
    ti_error(s, e, is_vf, P.ms > 0, P.err);
    return true;
}
// no-zero-toi refinement step (ticcd's loop behind tight_inclusion_ccd.cpp:59-72): returns true when another round is needed
__device__ inline bool ti_shrink(const double* s, const double* e, int is_vf, TIParams& P, double best)
{
    if (!(best == 0.0 && P.co_tol > 1e-300)) return false;
    if (10 * P.co_tol < P.ms) {
        P.ms *= 0.5;
    } else {
        P.co_tol *= 0.5;
        ti_tolerances(s, e, is_vf, P.co_tol, P.tol);
    }
    return true;
}

// Stage 2, one THREAD per pre-filtered query (persistent grid, the list length is read on the device): exact set-up
// and level-0 cull, then a short earliest-first depth-first search in registers.  It answers most queries with
// thousands of them in flight; a query that needs more than `budget` boxes in any run is deferred to the
// warp-cooperative kernel.
__global__ void __launch_bounds__(128)
    k_ti_query(MultiSource ms, const int* __restrict__ list, const unsigned long long* __restrict__ nlist, unsigned long long* next,
               double min_distance, double tmax_in, double tolerance, double rescale, int budget, int* __restrict__ hard,
               unsigned long long* nhard, CcdOut out)
{
    const int lane = threadIdx.x & 31;
    const unsigned long long n = *nlist;
    for (;;) {
        unsigned long long w0 = 0;
        if (lane == 0) w0 = atomicAdd(next, 32ull);
        w0 = __shfl_sync(0xffffffffu, w0, 0);
        if (w0 >= n) return;
        const unsigned long long w = w0 + lane;
        bool defer = false;
        int g = 0;
        if (w < n) {
            g = list[w];
            int64_t i;
            const QuerySource q = locate(ms, g, i);
            double s[12], e[12];
            TIParams P;
            int is_vf;
            double tmax0;
            if (ti_setup(q, i, min_distance, tmax_in, tolerance, rescale, out, s, e, P, is_vf, tmax0)) {
                TIBox stack[DFS_STACK];
                int sp = 0;
                double best = INFINITY;
                // ---- first query: minimum effective distance, no_zero_toi = false
                bool done = true;
                if (!ti_cull(s, e, is_vf, P, tmax0)) {
                    stack[0] = TIBox { 0, 0, 0, 0, 0, 0, 0 };
                    sp = 1;
                    done = ti_dfs(s, e, is_vf, P, out.bound, budget, stack, sp, best);
                }
                if (!done) {
                    defer = true;
                } else if (best < SMALL_TOI) {
                    // ---- second query: ms = min_distance, no_zero_toi = true, shrinking until toi != 0
                    P.ms = min_distance;
                    P.tmax = tmax0;
                    ti_error(s, e, is_vf, P.ms > 0, P.err);
                    for (int round = 0; round < 200; round++) {
                        best = INFINITY;
                        stack[0] = TIBox { 0, 0, 0, 0, 0, 0, 0 };
                        sp = 1;
                        if (!ti_dfs(s, e, is_vf, P, nullptr, budget, stack, sp, best)) {
                            defer = true;
                            break;
                        }
                        if (!ti_shrink(s, e, is_vf, P, best)) break;
                    }
                    if (!defer) report(out, i, best < INFINITY, best * rescale);
                } else {
                    report(out, i, best < INFINITY, best);
                }
            }
        }
        const unsigned long long slot = warp_reserve(nhard, defer, 1);
        if (defer) hard[slot] = g;
    }
}

// One WARP per surviving query: the 32 lanes pop up to 32 boxes from a shared-memory stack, evaluate them in
// parallel and push the children (earliest first on top), sharing the best terminal time.  The result (minimum
// lower time bound over all terminal boxes not pruned by it) does not depend on the evaluation order.  If the
// stack overflows, the query is handed to the global level-synchronous queue.
constexpr int WSTACK = 1024;
__device__ unsigned long long g_dbg[8]; // diagnostics: [0] warp searches, [1] warp iterations, [2] max iterations of one search
__device__ inline bool ti_warp_search(const double* s, const double* e, int is_vf, TIParams& P, const unsigned long long* bound, TIBox* stk,
                                      int cap, long long max_boxes, double& best, int lane)
{
    int sp = 1, iters = 0;
    if (lane == 0) stk[0] = TIBox { 0, 0, 0, 0, 0, 0, 0 };
    __syncwarp();
    best = INFINITY;
    while (sp > 0) {
        const int n = min(sp, 32);
        const bool have = lane < n;
        TIBox b = TIBox { 0, 0, 0, 0, 0, 0, 0 };
        if (have) b = stk[sp - 1 - lane];
        sp -= n;
        __syncwarp();
        if (bound) P.tmax = fmin(P.tmax, load_bound(bound));
        int r = 0, nchild = 0;
        TIBox child[2];
        double t0 = INFINITY;
        if (have) {
            double tt[2], uu[2], vv[2];
            box_bounds(b, tt, uu, vv);
            if (tt[0] < best && tt[0] <= P.tmax) {
                tt[1] = fmin(tt[1], fmin(best, P.tmax)); // only impacts before the best one so far (and inside the window) matter
                r = ti_step(s, e, is_vf, P, b, tt, uu, vv, child, nchild);
                if (r == 1) t0 = tt[0];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t0 = fmin(t0, __shfl_xor_sync(0xffffffffu, t0, o));
        best = fmin(best, t0);
        const int mine = r == 2 ? nchild : 0;
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (sp + total > cap) return false;
        // lane 0 held the top of the stack: its children go back on top
        const int off = sp + (total - incl);
        if (mine == 2) stk[off] = child[1], stk[off + 1] = child[0];
        else if (mine == 1) stk[off] = child[0];
        sp += total;
        __syncwarp();
        iters++;
        if (max_boxes >= 0 && (long long)iters * 32 >= max_boxes) {
            // effort cap (max_iterations of the reference's root finder): stop refining and answer with the earliest
            // time any unexplored box could still hold an impact — conservative, like ticcd when it runs out of iterations
            double tmin = INFINITY;
            for (int k = lane; k < sp; k += 32) {
                const TIBox& u = stk[k];
                tmin = fmin(tmin, ldexp(double(u.tn), -int(u.tk)));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tmin = fmin(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
            if (tmin <= P.tmax) best = fmin(best, tmin);
            break;
        }
    }
    if (lane == 0) atomicAdd(&g_dbg[0], 1ull), atomicAdd(&g_dbg[1], (unsigned long long)iters), atomicMax(&g_dbg[2], (unsigned long long)iters);
    return true;
}

constexpr int WARP_SEARCH_WARPS = 2;
// Persistent grid: every warp draws the next entry of the survivor list from a device counter until the list (whose
// length the pre-filter left in *nlist) is exhausted — no host round trip between the filter and the search.
__global__ void __launch_bounds__(32 * WARP_SEARCH_WARPS, 5)
    k_ti_warp(MultiSource ms, const int* __restrict__ list, const unsigned long long* __restrict__ nlist, unsigned long long* next,
              double min_distance, double tmax_in, double tolerance, double rescale, int cap, long long max_boxes, TIQueue Z, CcdOut out)
{
    __shared__ TIBox stk[WARP_SEARCH_WARPS][WSTACK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long n = *nlist;
    for (;;) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(next, 1ull);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n) return;
        int64_t i;
        const QuerySource q = locate(ms, list[w], i);
        double s[12], e[12];
        TIParams P;
        int is_vf;
        double tmax0;
        // every lane computes the same set-up (an early answer is reported by all lanes with identical values)
        if (!ti_setup(q, i, min_distance, tmax_in, tolerance, rescale, out, s, e, P, is_vf, tmax0)) continue;
        double best = INFINITY;
        int flags = 0;
        bool ok = true;
        if (!ti_cull(s, e, is_vf, P, tmax0)) ok = ti_warp_search(s, e, is_vf, P, out.bound, stk[warp], cap, max_boxes, best, lane);
        if (ok && best < SMALL_TOI) {
            P.ms = min_distance;
            P.tmax = tmax0;
            ti_error(s, e, is_vf, P.ms > 0, P.err);
            flags = 2;
            for (int round = 0; round < 200; round++) {
                ok = ti_warp_search(s, e, is_vf, P, nullptr, stk[warp], cap, max_boxes, best, lane);
                if (!ok || !ti_shrink(s, e, is_vf, P, best)) break;
            }
            if (ok && lane == 0) report(out, i, best < INFINITY, best * rescale);
        } else if (ok) {
            if (lane == 0) report(out, i, best < INFINITY, best);
        }
        if (!ok && lane == 0) { // stack overflow: restart this run from its root in the global queue
            const unsigned long long slot = atomicAdd(Z.nq, 1ull);
            const unsigned long long u = atomicAdd(Z.nout, 1ull);
            if (slot < Z.qcap && u < Z.ucap) {
                TIQuery& Q = Z.queries[slot];
#pragma unroll
                for (int k = 0; k < 12; k++) Q.s[k] = s[k], Q.e[k] = e[k];
                if (flags == 0) P.tmax = tmax0;
                Q.P = P;
                Q.is_vf = is_vf;
                Q.pad = 0;
                Q.src = i;
                Z.qtoi[slot] = 0x7ff0000000000000ull;
                Z.qflags[slot] = flags;
                Z.qcount[slot] = 0;
                Z.outq[u] = TIUnit { int(slot), TIBox { 0, 0, 0, 0, 0, 0, 0 } };
            }
            // a failed reservation is detected on the host (counter > capacity): the whole search is repeated with room
        }
        __syncwarp();
    }
}

// one level of the global queue
__global__ void __launch_bounds__(128)
    k_ti_level(TIQueue Z, unsigned long long nin, const unsigned long long* bound, long long max_boxes, int force_terminal, int last_attempt)
{
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    int nchild = 0;
    TIBox child[2];
    int qid = 0;
    double t0 = 0;
    if (i < nin) {
        const TIUnit u = Z.in[i];
        qid = u.q;
        const TIQuery& Q = Z.queries[u.q];
        double tt[2], uu[2], vv[2];
        box_bounds(u.b, tt, uu, vv);
        t0 = tt[0];
        const double best_q = __longlong_as_double((long long)*reinterpret_cast<volatile unsigned long long*>(Z.qtoi + u.q));
        bool live = tt[0] < best_q && tt[0] <= Q.P.tmax;
        // first-query units may also use the fresher global bound; refinement units keep their window
        if (live && bound && !(Z.qflags[u.q] & 2)) live = tt[0] <= load_bound(bound);
        if (live) {
            tt[1] = fmin(tt[1], fmin(best_q, Q.P.tmax));
            int r = ti_step(Q.s, Q.e, Q.is_vf, Q.P, u.b, tt, uu, vv, child, nchild);
            // effort cap per query (the reference root finder's max_iterations): further boxes are not refined, their
            // lower time bound is taken as a possible impact (conservative)
            if (r == 2 && max_boxes >= 0 && (long long)atomicAdd(Z.qcount + u.q, 1ull) >= max_boxes) r = 1, nchild = 0;
            if (r == 2 && force_terminal) r = 1, nchild = 0;
            if (r == 1) atomic_min_double(Z.qtoi + u.q, tt[0]);
        }
    }
    const unsigned long long slot = warp_reserve(Z.nout, nchild > 0, nchild);
    if (nchild > 0) {
        if (slot + nchild <= Z.ucap) {
            for (int k = 0; k < nchild; k++) Z.outq[slot + k] = TIUnit { qid, child[k] };
        } else if (last_attempt) {
            atomic_min_double(Z.qtoi + qid, t0); // queue still full after growing: stop refining this box conservatively
        }
        // otherwise the host grows the queue and runs this level again (nothing may be recorded here)
    }
}

// after the queue drained: queries of the first run with toi < SMALL_TOI enter the refinement (flags
// bit1) with ms = min_distance; refinement queries with toi == 0 shrink ms / tolerance and stay
// active; everything else reports its final time of impact.
__global__ void k_ti_finalize(int nq, TIQueue Z, double min_distance, double rescale, CcdOut out, unsigned long long* nactive)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const int flags = Z.qflags[i];
    if (flags & 4) return;
    TIQuery& Q = Z.queries[i];
    const double toi = __longlong_as_double((long long)Z.qtoi[i]);
    const bool hit = toi < INFINITY;
    if (!(flags & 2)) {
        if (hit && toi < SMALL_TOI) {
            Q.P.ms = min_distance;
            ti_error(Q.s, Q.e, Q.is_vf, Q.P.ms > 0, Q.P.err);
            Z.qtoi[i] = 0x7ff0000000000000ull;
            Z.qflags[i] = 2;
            Z.qcount[i] = 0;
            atomicAdd(nactive, 1ull);
        } else {
            report(out, Q.src, hit, toi);
            Z.qflags[i] = 4;
        }
    } else {
        if (hit && toi == 0.0 && Q.P.co_tol > 1e-300) {
            if (10 * Q.P.co_tol < Q.P.ms) {
                Q.P.ms *= 0.5;
            } else {
                Q.P.co_tol *= 0.5;
                ti_tolerances(Q.s, Q.e, Q.is_vf, Q.P.co_tol, Q.P.tol);
            }
            Z.qtoi[i] = 0x7ff0000000000000ull;
            Z.qcount[i] = 0;
            atomicAdd(nactive, 1ull);
        } else {
            report(out, Q.src, hit, toi * rescale);
            Z.qflags[i] = 4;
        }
    }
}
// root units for the queries that (re)start a refinement round
__global__ void k_ti_push_roots(int nq, const int* __restrict__ qflags, TIUnit* units, unsigned long long* count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool want = i < nq && (qflags[i] & 2) && !(qflags[i] & 4);
    const unsigned long long slot = warp_reserve(count, want, 1);
    if (want) units[slot] = TIUnit { i, TIBox { 0, 0, 0, 0, 0, 0, 0 } };
}
__global__ void k_set_bits(unsigned long long* dst, double v) { *dst = (unsigned long long)__double_as_longlong(v); }

struct TIWork {
    Buf<TIQuery> queries;
    Buf<unsigned long long> qtoi;
    Buf<int> qflags;
    Buf<unsigned long long> qcount;
    Buf<TIUnit> ua, ub;
    Buf<int> list; // candidates that survive the pre-filter
    Buf<int> hard; // queries deferred to the warp-cooperative search
    Buf<float4> XF; // FP32 positions (t0 | t1 per vertex, re-centred) for the FP32 pre-filter
};
// the scratch belongs to its context (allocated on the context's device, freed by ipcb_ctx_destroy)
void ti_work_free(TIWork* w) { delete w; }

static unsigned long long read_counter(ipcb_ctx* ctx, const unsigned long long* d)
{
    IPCB_CUDA(cudaMemcpyAsync(ctx->pinned.p, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
    return (unsigned long long)ctx->pinned.p[0];
}

// drain the global queue: `nunits` units are in W.ua.  When a level produces more children than the
// output buffer holds, the buffer is grown and the level is simply run again (a level only reads its
// input and does atomicMin on the per-query times, so it is idempotent).
static void ti_run_levels(ipcb_ctx* ctx, TIWork& W, TIQueue Z, unsigned long long nunits, const unsigned long long* bound, long long max_boxes)
{
    cudaStream_t s = ctx->stream;
    unsigned long long* cnt = ctx->dCounters.p + 2;
    Buf<TIUnit>* in = &W.ua;
    Buf<TIUnit>* outq = &W.ub;
    for (int level = 0; nunits > 0; level++) {
        for (int attempt = 0;; attempt++) {
            IPCB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), s));
            Z.in = in->p, Z.outq = outq->p, Z.nout = cnt, Z.ucap = outq->cap;
            k_ti_level<<<grid_for(nunits, 128), 128, 0, s>>>(Z, nunits, bound, max_boxes, level >= 120 ? 1 : 0, attempt >= 3 ? 1 : 0);
            ctx->launches++;
            IPCB_CUDA(cudaGetLastError());
            const unsigned long long produced = read_counter(ctx, cnt);
            if (produced <= outq->cap || attempt >= 3) {
                nunits = std::min<unsigned long long>(produced, outq->cap);
                break;
            }
            outq->reserve(size_t(produced) + size_t(produced) / 2); // contents need not be preserved
        }
        std::swap(in, outq);
    }
}

// in-register boxes per query before it is deferred to the warp-cooperative search, and the shared-memory stack
// capacity per warp before a query goes to the global queue (the environment variables are test hooks that force
// the later stages on small inputs)
static int dfs_budget()
{
    const char* e = getenv("IPCB_TI_BUDGET");
    return e ? std::max(1, atoi(e)) : 96;
}
static int warp_stack_cap()
{
    const char* e = getenv("IPCB_TI_WSTACK");
    return e ? std::min(WSTACK, std::max(2, atoi(e))) : WSTACK;
}

// Tight-Inclusion CCD over a (multi-kind) query source.
//   phase = pre-filter (most broad-phase candidates are separated along one of their own directions inside the
//   current search window) -> survivor list -> warp-cooperative search (persistent grid, device-side list length).
// With a shared bound (step-size search) strided SAMPLES of the candidates are probed first: the earliest time of
// impact they find shrinks the window of the final pass over all candidates — the std::atomic<double> earliest_toi
// idea of candidates.cpp:267-286 applied before any root finding — so the pre-filter of the bulk discards almost
// everything that cannot beat the bound.
// Nothing is read back until the end; stack overflows (rare) are collected in the global queue and drained last.
static void ti_run(ipcb_ctx* ctx, const MultiSource& ms, double min_distance, double tmax, const ipcb_ccd_params& p, CcdOut out)
{
    // the scene box bounds the coordinates of resident-candidate queries when the swept broad phase of THESE positions
    // built it (ccd_stepsize right after candidates_build); otherwise every query takes its own maximum
    const float* scene = (ms.cand[0] && ctx->scene_covers_positions) ? ctx->scene.p : nullptr;
    const int64_t total = ms.off[ms.nk];
    if (total == 0) return;
    if (total > 0x7fffffffll) throw Error("ccd: more than 2^31 candidates in one search");
    cudaStream_t s = ctx->stream;
    if (!ctx->ti_work) ctx->ti_work = new TIWork();
    TIWork& W = *ctx->ti_work;
    unsigned long long* nq_d = ctx->dCounters.p + 1;
    unsigned long long* cnt = ctx->dCounters.p + 2; // units in the global queue (directly after nq_d)
    unsigned long long* nactive_d = ctx->dCounters.p + 3;
    unsigned long long* nlist_d = ctx->dCounters.p + 9; // survivor count, its work counter, deferred count, its work counter
    unsigned long long* next_d = ctx->dCounters.p + 10;
    unsigned long long* nhard_d = ctx->dCounters.p + 11;
    unsigned long long* next2_d = ctx->dCounters.p + 12;
    size_t qcap = std::max<size_t>(W.queries.cap, 1024);
    size_t ucap = std::max<size_t>(W.ua.cap, size_t(1) << 16);
    W.list.reserve(total), W.hard.reserve(total);
    // FP32 pre-filter: resident candidates of the face-vertex / edge-edge kinds with the scene box of THESE positions,
    // no per-query outputs (a dropped query only needs reporting in the narrow-phase batch API); IPCB_TI_FILTER_F64: A/B
    static const bool force_f64 = getenv("IPCB_TI_FILTER_F64") != nullptr;
    bool use_f32 = scene != nullptr && !out.hit && !force_f64;
    for (int k = 0; k < ms.nk; k++) use_f32 = use_f32 && ms.kind[k] >= IPCB_EE;
    if (use_f32) {
        W.XF.reserve(2 * size_t(ctx->nV));
        k_pack_f32<<<grid_for(ctx->nV, 256), 256, 0, s>>>(ctx->nV, ctx->X0.p, ctx->X1.p, scene, W.XF.p);
        ctx->launches++;
    }
    // strided samples, coarse to fine (default: one sample of ~64K candidates, then everything else; IPCB_TI_GROWTH
    // inserts intermediate samples)
    int64_t sample = 65536;
    int growth = 1 << 20;
    if (const char* e = getenv("IPCB_TI_SAMPLE")) sample = std::max(1, atoi(e));
    if (const char* e = getenv("IPCB_TI_GROWTH")) growth = std::max(2, atoi(e));
    int strides[8], nphase = 0;
    if (out.bound && total / sample >= 4) {
        // a candidate that two phases both look at (strides that do not divide each other) is simply searched twice
        for (int64_t st = std::min<int64_t>(total / sample, 1 << 24); st > 1 && nphase < 7; st /= growth) strides[nphase++] = int(st);
    }
    strides[nphase++] = 1;
    unsigned long long nq = 0, nunits = 0;
    for (int attempt = 0;; attempt++) {
        W.queries.reserve(qcap), W.qtoi.reserve(qcap), W.qflags.reserve(qcap), W.qcount.reserve(qcap);
        W.ua.reserve(ucap), W.ub.reserve(ucap);
        qcap = std::min(std::min(W.queries.cap, W.qcount.cap), std::min(W.qtoi.cap, W.qflags.cap));
        ucap = std::min(W.ua.cap, W.ub.cap);
        IPCB_CUDA(cudaMemsetAsync(nq_d, 0, 2 * sizeof(unsigned long long), s)); // nq and the unit counter
        TIQueue Z { W.queries.p, W.qtoi.p, W.qflags.p, W.qcount.p, nq_d, (unsigned long long)qcap, nullptr, W.ua.p, cnt, (unsigned long long)ucap };
        // probe = true: a sample phase.  Its only purpose is a good bound, so queries that need more than the
        // in-register budget are NOT pursued (an expensive search above the final step size is wasted work, and a
        // single pathological one can cost milliseconds): every sampled candidate is looked at again by the final
        // phase, then inside a tight window.
        auto phase = [&](int st, bool probe) {
            IPCB_CUDA(cudaMemsetAsync(nlist_d, 0, 4 * sizeof(unsigned long long), s)); // list lengths and work counters
            {
                Stage kt(ctx, probe ? "k:k_ti_filter(sample)" : "k:k_ti_filter", s);
                if (use_f32)
                    k_ti_filter32<<<grid_for((total + st - 1) / st, 256), 256, 0, s>>>(ms, st, min_distance, tmax, p.tolerance, out.bound, W.list.p, nlist_d,
                                                                                     scene, W.XF.p);
                else
                    k_ti_filter<<<grid_for((total + st - 1) / st, 256), 256, 0, s>>>(ms, st, 0, min_distance, tmax, p.tolerance, out, W.list.p, nlist_d, scene);
            }
            {
                Stage kt(ctx, probe ? "k:k_ti_query(sample)" : "k:k_ti_query", s);
                k_ti_query<<<NUM_SMS * 2, 128, 0, s>>>(ms, W.list.p, nlist_d, next_d, min_distance, tmax, p.tolerance, p.conservative_rescaling,
                                                       dfs_budget(), W.hard.p, nhard_d, out);
            }
            ctx->launches += 2;
            if (!probe) {
                Stage kt(ctx, "k:k_ti_warp", s);
                k_ti_warp<<<NUM_SMS * 5, 32 * WARP_SEARCH_WARPS, 0, s>>>(ms, W.hard.p, nhard_d, next2_d, min_distance, tmax, p.tolerance,
                                                                        p.conservative_rescaling, warp_stack_cap(), (long long)p.max_iterations, Z, out);
                ctx->launches++;
            }
            static const bool debug = getenv("IPCB_DEBUG") != nullptr;
            if (debug) { // per-phase survivor count and time (synchronises; diagnostics only)
                const auto t0 = std::chrono::steady_clock::now();
                const unsigned long long nl = read_counter(ctx, nlist_d);
                const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                double b = 0;
                if (out.bound) {
                    const unsigned long long bits = read_counter(ctx, out.bound);
                    memcpy(&b, &bits, sizeof b);
                }
                unsigned long long dbg[8];
                cudaMemcpyFromSymbol(dbg, g_dbg, sizeof dbg);
                const unsigned long long zero[8] = { 0 };
                cudaMemcpyToSymbol(g_dbg, zero, sizeof zero);
                fprintf(stderr, "[ipcb]   phase stride %d: %llu survivors, %llu deferred, bound %.6f, waited %.3f ms; warp searches %llu, iterations %llu, max %llu\n",
                        st, nl, read_counter(ctx, nhard_d), b, ms, dbg[0], dbg[1], dbg[2]);
            }
        };
        for (int k = 0; k < nphase; k++) phase(strides[k], k + 1 < nphase);
        IPCB_CUDA(cudaGetLastError());
        nq = read_counter(ctx, nq_d);
        nunits = nq ? read_counter(ctx, cnt) : 0;
        static const bool debug = getenv("IPCB_DEBUG") != nullptr;
        if (debug)
            fprintf(stderr, "[ipcb] ti_run: %lld candidates, stride %d, attempt %d: %llu queries / %llu units in the global queue (cap %zu / %zu)\n",
                    (long long)total, strides[0], attempt, nq, nunits, qcap, ucap);
        if (nq <= qcap && nunits <= ucap) break;
        if (attempt > 3) throw Error("ccd: spill buffers overflow persisted");
        // reports are idempotent (atomicMin / identical values), so the search can simply be repeated with room
        qcap = std::max<size_t>(qcap, nq + nq / 4);
        ucap = std::max<size_t>(ucap, nq + nq / 4);
    }
    if (nq == 0) return;
    // ---- queries whose shared-memory stack overflowed: global level-synchronous queue
    TIQueue Z { W.queries.p, W.qtoi.p, W.qflags.p, W.qcount.p, nq_d, (unsigned long long)qcap, nullptr, nullptr, nullptr, (unsigned long long)ucap };
    ti_run_levels(ctx, W, Z, nunits, out.bound, (long long)p.max_iterations);
    for (int round = 0; round < 250; round++) {
        IPCB_CUDA(cudaMemsetAsync(nactive_d, 0, sizeof(unsigned long long), s));
        k_ti_finalize<<<grid_for(nq, 128), 128, 0, s>>>(int(nq), Z, min_distance, p.conservative_rescaling, out, nactive_d);
        ctx->launches++;
        if (read_counter(ctx, nactive_d) == 0) break;
        IPCB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), s));
        k_ti_push_roots<<<grid_for(nq, 128), 128, 0, s>>>(int(nq), W.qflags.p, W.ua.p, cnt);
        ctx->launches++;
        ti_run_levels(ctx, W, Z, read_counter(ctx, cnt), nullptr, (long long)p.max_iterations);
    }
}

static QuerySource cand_source(ipcb_ctx* ctx, int kind)
{
    return { kind, ctx->cand[kind].count, ctx->cand[kind].pairs.p, ctx->dE.p, ctx->dF.p, ctx->X0.p, ctx->X1.p, nullptr, nullptr };
}

static MultiSource multi_source(ipcb_ctx* ctx, std::initializer_list<int> kinds)
{
    MultiSource m {};
    m.E = ctx->dE.p, m.F = ctx->dF.p, m.X0 = ctx->X0.p, m.X1 = ctx->X1.p;
    m.off[0] = 0;
    for (int kind : kinds) {
        if (ctx->cand[kind].count == 0) continue;
        m.kind[m.nk] = kind;
        m.cand[m.nk] = ctx->cand[kind].pairs.p;
        m.off[m.nk + 1] = m.off[m.nk] + ctx->cand[kind].count;
        m.nk++;
    }
    return m;
}

// Candidates::compute_collision_free_stepsize (candidates.cpp:252-292): earliest TOI over the
// resident candidates, 1.0 when there are none; the result is left in *d_out (device)
void ccd_stepsize(ipcb_ctx* ctx, double min_distance, const ipcb_ccd_params& p, double* d_out)
{
    Stage st(ctx, "ccd_narrow");
    cudaStream_t s = ctx->stream;
    unsigned long long* bound = ctx->dCounters.p + 8;
    k_set_bits<<<1, 1, 0, s>>>(bound, 1.0);
    ctx->launches++;
    CcdOut out { bound, nullptr, nullptr };
    if (p.kind == IPCB_CCD_ADDITIVE) {
        for (int kind : { IPCB_VV, IPCB_EV, IPCB_FV, IPCB_EE }) {
            const QuerySource src = cand_source(ctx, kind);
            if (src.n == 0) continue;
            k_additive<<<grid_for(src.n, 128), 128, 0, s>>>(src, min_distance, 1.0, (long long)p.max_iterations, p.conservative_rescaling, out);
            ctx->launches++;
        }
        IPCB_CUDA(cudaGetLastError());
    } else {
        // the few codimensional point-point / point-edge candidates first, then faces and edges together
        ti_run(ctx, multi_source(ctx, { IPCB_VV, IPCB_EV }), min_distance, 1.0, p, out);
        ti_run(ctx, multi_source(ctx, { IPCB_FV, IPCB_EE }), min_distance, 1.0, p, out);
    }
    IPCB_CUDA(cudaMemcpyAsync(d_out, bound, sizeof(double), cudaMemcpyDeviceToDevice, s));
}

// ---- Candidates::compute_noncandidate_conservative_stepsize (candidates.cpp:294-338)
__global__ void k_mark_candidate_vertices(int kind, int64_t n, const int2* __restrict__ cand, const int2* __restrict__ E, const int4* __restrict__ F,
                                          unsigned char* __restrict__ flag)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int2 c = cand[i];
    if (kind == IPCB_VV) {
        flag[c.x] = 1, flag[c.y] = 1;
    } else if (kind == IPCB_EV) {
        const int2 e = __ldg(E + c.x);
        flag[c.y] = 1, flag[e.x] = 1, flag[e.y] = 1;
    } else if (kind == IPCB_EE) {
        const int2 ea = __ldg(E + c.x), eb = __ldg(E + c.y);
        flag[ea.x] = 1, flag[ea.y] = 1, flag[eb.x] = 1, flag[eb.y] = 1;
    } else {
        const int4 f = __ldg(F + c.x);
        flag[c.y] = 1, flag[f.x] = 1, flag[f.y] = 1, flag[f.z] = 1;
    }
}
// max over flagged vertices of |displacement| (X1 holds the displacements, or X1 - X0 is formed when X0 is given)
__global__ void k_max_displacement(int n, const unsigned char* __restrict__ flag, const double4* __restrict__ X0, const double4* __restrict__ X1,
                                   unsigned long long* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double d = 0;
    if (i < n && flag[i]) {
        double4 u = X1[i];
        if (X0) {
            const double4 a = X0[i];
            u.x -= a.x, u.y -= a.y, u.z -= a.z;
        }
        d = sqrt((u.x * u.x + u.y * u.y) + u.z * u.z); // Eigen's linear reduction of a dynamic row
    }
    for (int o = 16; o > 0; o >>= 1) d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0 && d > 0) atomicMax(out, (unsigned long long)__double_as_longlong(d));
}
// X0 == nullptr: ctx->X1 holds the displacements
double noncandidate_stepsize(ipcb_ctx* ctx, bool difference, double dhat)
{
    int64_t total = 0;
    for (auto& c : ctx->cand) total += c.count;
    if (total == 0) return 1.0; // no possible collisions: full step
    cudaStream_t s = ctx->stream;
    ctx->hflag.reserve(size_t(ctx->nV) + 1);
    IPCB_CUDA(cudaMemsetAsync(ctx->hflag.p, 0, size_t(ctx->nV) + 1, s));
    for (int k = 0; k < 4; k++) {
        const PairList& pl = ctx->cand[k];
        if (pl.count == 0) continue;
        k_mark_candidate_vertices<<<grid_for(pl.count, 256), 256, 0, s>>>(k, pl.count, pl.pairs.p, ctx->dE.p, ctx->dF.p, ctx->hflag.p);
        ctx->launches++;
    }
    unsigned long long* out = ctx->dCounters.p + 14;
    IPCB_CUDA(cudaMemsetAsync(out, 0, sizeof(unsigned long long), s));
    k_max_displacement<<<grid_for(ctx->nV, 256), 256, 0, s>>>(ctx->nV, ctx->hflag.p, difference ? ctx->X0.p : nullptr, ctx->X1.p, out);
    ctx->launches++;
    IPCB_CUDA(cudaGetLastError());
    const unsigned long long bits = read_counter(ctx, out);
    double m;
    memcpy(&m, &bits, sizeof m);
    return 0.5 * dhat / m;
}

// compute_collision_free_stepsize when a candidate list would exceed max_pairs() (SURVEY §7 hard part 7; BASELINE config
// 4: 10^8 .. 10^9 swept candidates): the query leaves of the rank's range are traversed in CHUNKS, every chunk's pairs
// go straight through the narrow phase and are dropped.  The earliest time of impact found so far prunes the later
// chunks exactly like the shared bound prunes later queries, so most of them end in the pre-filter.  A chunk that
// still overflows is halved; chunks grow again while their lists stay below a quarter of the limit.
void ccd_stepsize_streaming(ipcb_ctx* ctx, double min_distance, const ipcb_ccd_params& p, double* d_out)
{
    Stage st(ctx, "ccd_streaming");
    cudaStream_t s = ctx->stream;
    unsigned long long* bound = ctx->dCounters.p + 8;
    k_set_bits<<<1, 1, 0, s>>>(bound, 1.0);
    ctx->launches++;
    CcdOut out { bound, nullptr, nullptr };
    auto narrow = [&](std::initializer_list<int> kinds) {
        if (p.kind == IPCB_CCD_ADDITIVE) {
            for (int kind : kinds) {
                const QuerySource src = cand_source(ctx, kind);
                if (src.n == 0) continue;
                k_additive<<<grid_for(src.n, 128), 128, 0, s>>>(src, min_distance, 1.0, (long long)p.max_iterations, p.conservative_rescaling, out);
                ctx->launches++;
            }
            IPCB_CUDA(cudaGetLastError());
        } else {
            ti_run(ctx, multi_source(ctx, kinds), min_distance, 1.0, p, out);
        }
    };
    narrow({ IPCB_VV, IPCB_EV }); // the few codimensional candidates are resident
    for (int kind : { IPCB_FV, IPCB_EE }) {
        if (ctx->cand_overflow[kind] == 0) { // this kind fitted: its list is resident
            narrow({ kind });
            continue;
        }
        int range[2];
        traverse_chunk(ctx, kind, -1, -1, range);
        const double per_query = double(ctx->cand_overflow[kind]) / std::max(1, range[1] - range[0]);
        int len = std::max(64, int(0.5 * double(max_pairs()) / std::max(per_query, 1e-9)));
        for (int q = range[0]; q < range[1];) {
            const int qe = int(std::min<int64_t>(range[1], int64_t(q) + len));
            const unsigned long long found = traverse_chunk(ctx, kind, q, qe, nullptr);
            if (found > max_pairs()) {
                if (qe - q <= 1) throw Error("ccd: one query leaf has more candidates than IPCB_MAX_PAIRS allows");
                len = std::max(1, int(0.8 * double(qe - q) * double(max_pairs()) / double(found)));
                continue;
            }
            narrow({ kind });
            q = qe;
            if (found < max_pairs() / 4) len = int(std::min<int64_t>(int64_t(len) * 2, 1 << 30));
        }
        ctx->cand[kind].count = 0; // the last chunk is not the kind's candidate set
    }
    IPCB_CUDA(cudaMemcpyAsync(d_out, bound, sizeof(double), cudaMemcpyDeviceToDevice, s));
}

void ccd_narrow_phase(ipcb_ctx* ctx, int kind, int64_t n, const double* h_t0, const double* h_t1, double min_distance, double tmax,
                      const ipcb_ccd_params& p, uint8_t* h_hit, double* h_toi)
{
    if (n == 0) return;
    cudaStream_t s = ctx->stream;
    Buf<double> a, b, toi;
    Buf<unsigned char> hit;
    a.reserve(12 * n), b.reserve(12 * n), toi.reserve(n), hit.reserve(n);
    IPCB_CUDA(cudaMemcpyAsync(a.p, h_t0, sizeof(double) * 12 * n, cudaMemcpyHostToDevice, s));
    IPCB_CUDA(cudaMemcpyAsync(b.p, h_t1, sizeof(double) * 12 * n, cudaMemcpyHostToDevice, s));
    IPCB_CUDA(cudaMemsetAsync(hit.p, 0, n, s));
    CcdOut out { nullptr, hit.p, toi.p };
    if (p.kind == IPCB_CCD_ADDITIVE) {
        const QuerySource src { kind, n, nullptr, nullptr, nullptr, nullptr, nullptr, a.p, b.p };
        k_additive<<<grid_for(n, 128), 128, 0, s>>>(src, min_distance, tmax, (long long)p.max_iterations, p.conservative_rescaling, out);
        ctx->launches++;
        IPCB_CUDA(cudaGetLastError());
    } else {
        MultiSource m {};
        m.nk = 1, m.kind[0] = kind, m.off[0] = 0, m.off[1] = n, m.raw0 = a.p, m.raw1 = b.p;
        ti_run(ctx, m, min_distance, tmax, p, out);
    }
    IPCB_CUDA(cudaMemcpyAsync(h_hit, hit.p, n, cudaMemcpyDeviceToHost, s));
    IPCB_CUDA(cudaMemcpyAsync(h_toi, toi.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    IPCB_CUDA(cudaStreamSynchronize(s));
}

} // namespace ipcb
