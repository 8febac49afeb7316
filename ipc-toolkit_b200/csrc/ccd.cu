// ccd.cu — narrow-phase CCD and the collision-free step size.
//
// Replaces (reference src/ipc/): candidates/candidates.cpp:252-292
// (Candidates::compute_collision_free_stepsize with its shared earliest-TOI
// pruning), ccd/tight_inclusion_ccd.cpp:33-336 (strategy wrapper),
// ccd/additive_ccd.cpp:71-325, ccd/check_initial_distance.hpp.  The
// Tight-Inclusion root finder itself (third-party ticcd v1.0.6, not in the
// reference tree) is re-designed for the GPU from the published algorithm:
//
//   * level 0: one thread per candidate gathers its 8 points, evaluates the
//     initial distance, tolerances and the floating-point filter, and tests the
//     root box [0,1]^3; only survivors get a query record and a queue unit;
//   * a GLOBAL interval-subdivision queue: every unit is (query, dyadic t/u/v
//     box); one kernel launch processes a whole level of all live queries, tests
//     the 8-corner co-domain box against the eps-cube, records terminal boxes
//     with atomicMin on the query's TOI bit pattern, and pushes split children
//     with warp-aggregated atomics.  The earliest TOI over all terminal boxes is
//     exactly what the sequential breadth-first search returns (it returns the
//     first terminal box in (level, t) order), see DESIGN.md;
//   * cross-query pruning with a global bound mirrors the reference's atomic
//     earliest_toi (candidates.cpp:267-286);
//   * the rare queries whose TOI is below SMALL_TOI are re-run with the
//     reference's no-zero-TOI refinement (tight_inclusion_ccd.cpp:59-72).
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
__device__ inline d3 ldp(const double4* X, int i) { return load_vertex(X, i); }
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
        for (int k = 0; k < n; k++) a[k] = ldp(q.X0, vid[k]), b[k] = ldp(q.X1, vid[k]);
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
__device__ inline double load_bound(const unsigned long long* addr)
{
    return __longlong_as_double((long long)*reinterpret_cast<const volatile unsigned long long*>(addr));
}

struct CcdOut {
    unsigned long long* bound; // global earliest TOI (bit pattern), or nullptr
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
struct TIQuery {
    double s[12], e[12]; // the 4 points handed to the root finder at t = 0 / t = 1
    double tol[3];       // domain tolerances (t, u, v)
    double err[3];       // floating-point filter per coordinate
    double ms;           // minimum separation of the current run
    double co_tol;       // co-domain tolerance (delta)
    double tmax;
    double d0;           // initial distance
    int is_vf;
    int pad;
    long long src; // index of the originating candidate / query
};
struct TIUnit {
    int q;
    unsigned tn, un, vn; // dyadic numerators
    unsigned char tk, uk, vk, pad;
};

__device__ inline double linf3(d3 a) { return fmax(fmax(fabs(a.x), fabs(a.y)), fabs(a.z)); }
__device__ inline d3 getp(const double* p, int k) { return { p[3 * k], p[3 * k + 1], p[3 * k + 2] }; }

__device__ inline void ti_tolerances(const double* s, const double* e, int is_vf, double co_tol, double* tol)
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
__device__ inline void ti_error(const double* s, const double* e, int is_vf, bool using_ms, double* err)
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
// co-domain box of the root function over a (t,u,v) box vs the eps-cube
__device__ inline bool ti_inclusion(const TIQuery& Q, const double* tt, const double* uu, const double* vv, bool& box_in,
                                    double* true_tol)
{
    box_in = true;
    for (int c = 0; c < 3; c++) {
        double vmin = INFINITY, vmax = -INFINITY;
#pragma unroll
        for (int a = 0; a < 2; a++) {
            const double t = tt[a];
            const double p0 = (Q.e[c] - Q.s[c]) * t + Q.s[c];
            const double p1 = (Q.e[3 + c] - Q.s[3 + c]) * t + Q.s[3 + c];
            const double p2 = (Q.e[6 + c] - Q.s[6 + c]) * t + Q.s[6 + c];
            const double p3 = (Q.e[9 + c] - Q.s[9 + c]) * t + Q.s[9 + c];
#pragma unroll
            for (int b = 0; b < 2; b++)
#pragma unroll
                for (int d = 0; d < 2; d++) {
                    double val;
                    if (Q.is_vf) {
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
        const double lim = Q.err[c] + Q.ms;
        true_tol[c] = vmax - vmin;
        if (vmin > lim || vmax < -lim) return false;
        if (vmin < -lim || vmax > lim) box_in = false;
    }
    return true;
}

struct TIQueue {
    TIQuery* queries;
    unsigned long long* qtoi; // per query earliest terminal t (bit pattern), +inf when none
    int* qflags;              // per query: bit0 = excluded from the global bound
    unsigned long long* nq;   // survivor counter
    unsigned long long qcap;
    TIUnit* in;
    TIUnit* outq;
    unsigned long long* nout;
    unsigned long long ucap;
};

// warp-aggregated slot reservation
__device__ inline unsigned long long warp_reserve(unsigned long long* counter, bool want, int count)
{
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (m == 0) return 0;
    const int lane = threadIdx.x & 31;
    // every wanting lane asks for `count` slots (count is 1 or 2); prefix over lanes
    int mine = want ? count : 0;
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

// level 0 (tight_inclusion_ccd.cpp:222-336 + ccd_strategy :33-58 up to the first root-finder call)
__global__ void __launch_bounds__(128)
    k_ti_level0(QuerySource q, double min_distance, double tmax, double tolerance, double rescale, TIQueue Z, CcdOut out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool survive = false;
    TIQuery Q;
    if (i < q.n) {
        d3 a[4], b[4];
        const int n = load_query(q, i, a, b);
        const double d0 = sqrt(auto_distance(q.kind, a));
        bool moving = false;
        for (int k = 0; k < n; k++) moving |= !same(a[k], b[k]);
        if (d0 <= min_distance) { // check_initial_distance: toi = 0 (also the no-motion answer)
            report(out, i, true, 0.0);
        } else if (!moving) {
            report(out, i, false, 0.0);
        } else {
            // points handed to the root finder: VV / EV are degenerate edge-edge queries (:104,:177)
            d3 s4[4], e4[4];
            if (q.kind == IPCB_VV) {
                s4[0] = s4[1] = a[0], s4[2] = s4[3] = a[1];
                e4[0] = e4[1] = b[0], e4[2] = e4[3] = b[1];
            } else if (q.kind == IPCB_EV) {
                s4[0] = s4[1] = a[0], s4[2] = a[1], s4[3] = a[2];
                e4[0] = e4[1] = b[0], e4[2] = b[1], e4[3] = b[2];
            } else {
                for (int k = 0; k < 4; k++) s4[k] = a[k], e4[k] = b[k];
            }
            for (int k = 0; k < 4; k++) {
                Q.s[3 * k] = s4[k].x, Q.s[3 * k + 1] = s4[k].y, Q.s[3 * k + 2] = s4[k].z;
                Q.e[3 * k] = e4[k].x, Q.e[3 * k + 1] = e4[k].y, Q.e[3 * k + 2] = e4[k].z;
            }
            Q.is_vf = q.kind == IPCB_FV;
            Q.d0 = d0;
            Q.tmax = tmax;
            Q.src = i;
            Q.pad = 0;
            double med = (1.0 - rescale) * (d0 - min_distance); // min effective distance (:45-49)
            med = fmin(med, 1e-4);
            med += min_distance;
            Q.ms = med;
            Q.co_tol = fmin(0.5 * d0, tolerance); // adjusted tolerance (:245-246)
            ti_tolerances(Q.s, Q.e, Q.is_vf, Q.co_tol, Q.tol);
            ti_error(Q.s, Q.e, Q.is_vf, Q.ms > 0, Q.err);
            const double unit[2] = { 0.0, 1.0 };
            bool box_in;
            double tt[3];
            survive = ti_inclusion(Q, unit, unit, unit, box_in, tt);
            if (!survive) report(out, i, false, 0.0);
        }
    }
    const unsigned long long slot = warp_reserve(Z.nq, survive, 1);
    if (survive) {
        if (slot < Z.qcap) {
            Z.queries[slot] = Q;
            Z.qtoi[slot] = 0x7ff0000000000000ull;
            Z.qflags[slot] = 0;
        }
        // slot >= qcap is detected on the host (counter > capacity) and the pass is repeated
    }
}

// push the root unit of every query whose flags match (need_mask == 0: any) and avoid exclude_mask
__global__ void k_ti_push_roots(int nq, const int* __restrict__ qflags, int need_mask, int exclude_mask, TIUnit* units,
                                unsigned long long* count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool want = i < nq;
    if (want) {
        const int f = qflags[i];
        want = (need_mask == 0 || (f & need_mask)) && !(f & exclude_mask);
    }
    const unsigned long long slot = warp_reserve(count, want, 1);
    if (want) units[slot] = TIUnit { i, 0, 0, 0, 0, 0, 0, 0 };
}

constexpr double SMALL_TOI = 1e-6;

__global__ void __launch_bounds__(128)
    k_ti_level(TIQueue Z, unsigned long long nin, unsigned long long* bound, int use_bound, int force_terminal)
{
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    int nchild = 0;
    TIUnit child[2];
    if (i < nin) {
        const TIUnit u = Z.in[i];
        const TIQuery& Q = Z.queries[u.q];
        const double tt[2] = { ldexp(double(u.tn), -int(u.tk)), ldexp(double(u.tn + 1), -int(u.tk)) };
        const double uu[2] = { ldexp(double(u.un), -int(u.uk)), ldexp(double(u.un + 1), -int(u.uk)) };
        const double vv[2] = { ldexp(double(u.vn), -int(u.vk)), ldexp(double(u.vn + 1), -int(u.vk)) };
        const double best_q = __longlong_as_double((long long)*reinterpret_cast<volatile unsigned long long*>(Z.qtoi + u.q));
        bool live = tt[0] < best_q; // TOI_SKIP pruning of the sequential search
        if (live && use_bound && !(Z.qflags[u.q] & 1)) live = tt[0] < load_bound(bound);
        if (live) {
            bool box_in;
            double true_tol[3];
            if (ti_inclusion(Q, tt, uu, vv, box_in, true_tol)) {
                const double w[3] = { tt[1] - tt[0], uu[1] - uu[0], vv[1] - vv[0] };
                const bool tol_cond = true_tol[0] <= Q.co_tol && true_tol[1] <= Q.co_tol && true_tol[2] <= Q.co_tol;
                const bool cond1 = w[0] <= Q.tol[0] && w[1] <= Q.tol[1] && w[2] <= Q.tol[2];
                bool terminal = cond1 || tol_cond || box_in || force_terminal;
                int split = 0;
                if (!terminal) {
                    double bestv = -INFINITY;
                    for (int k = 0; k < 3; k++) {
                        const double r = w[k] > Q.tol[k] ? w[k] / Q.tol[k] : -INFINITY;
                        if (r > bestv) bestv = r, split = k;
                    }
                    const int kk = split == 0 ? u.tk : (split == 1 ? u.uk : u.vk);
                    if (kk >= 30) terminal = true; // bisection depth exhausted: stop conservatively
                }
                if (terminal) {
                    atomic_min_double(Z.qtoi + u.q, tt[0]);
                    if (use_bound && !(Z.qflags[u.q] & 1) && tt[0] >= SMALL_TOI) atomic_min_double(bound, tt[0]);
                } else {
                    TIUnit c0 = u, c1 = u;
                    if (split == 0) {
                        c0.tn = 2 * u.tn, c1.tn = 2 * u.tn + 1, c0.tk = c1.tk = u.tk + 1;
                        child[nchild++] = c0; // first half always overlaps [0, tmax]
                        const double mid = ldexp(double(c1.tn), -int(c1.tk));
                        if (Q.tmax == 1.0 || mid <= Q.tmax) child[nchild++] = c1;
                    } else if (split == 1) {
                        c0.un = 2 * u.un, c1.un = 2 * u.un + 1, c0.uk = c1.uk = u.uk + 1;
                        child[nchild++] = c0;
                        const double mid = ldexp(double(c1.un), -int(c1.uk));
                        if (!Q.is_vf || mid + vv[0] <= 1.0) child[nchild++] = c1; // u + v <= 1
                    } else {
                        c0.vn = 2 * u.vn, c1.vn = 2 * u.vn + 1, c0.vk = c1.vk = u.vk + 1;
                        child[nchild++] = c0;
                        const double mid = ldexp(double(c1.vn), -int(c1.vk));
                        if (!Q.is_vf || mid + uu[0] <= 1.0) child[nchild++] = c1;
                    }
                    // (the first half satisfies u + v <= 1 whenever its parent did)
                }
            }
        }
    }
    const unsigned long long slot = warp_reserve(Z.nout, nchild > 0, nchild);
    if (nchild > 0) {
        if (slot + nchild <= Z.ucap) {
            for (int k = 0; k < nchild; k++) Z.outq[slot + k] = child[k];
        } else {
            // queue full: stop refining this box conservatively (its lower time bound is a valid TOI)
            const TIUnit u = Z.in[i];
            const double t0 = ldexp(double(u.tn), -int(u.tk));
            atomic_min_double(Z.qtoi + u.q, t0);
        }
    }
}

// after a run: classify queries. mode 0 (after phase 1): flag queries with toi < SMALL_TOI for the
// no-zero-toi refinement (bit 1 = active, bit 0 = excluded from bound) and set their refinement
// parameters; others report.  mode 1 (refinement round): queries with toi == 0 shrink ms / tolerance
// and stay active; the rest report toi * rescale (tight_inclusion_ccd.cpp:59-72 and the ticcd
// no_zero_toi loop).
__global__ void k_ti_finalize(int nq, TIQueue Z, int mode, double min_distance, double rescale, CcdOut out, unsigned long long* nactive)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    TIQuery& Q = Z.queries[i];
    const double toi = __longlong_as_double((long long)Z.qtoi[i]);
    const bool hit = toi < INFINITY;
    int flags = Z.qflags[i];
    if (mode == 0) {
        if (flags & 4) return; // already finished in an earlier pass
        if (hit && toi < SMALL_TOI) {
            Q.ms = min_distance;
            ti_error(Q.s, Q.e, Q.is_vf, Q.ms > 0, Q.err);
            Z.qtoi[i] = 0x7ff0000000000000ull;
            Z.qflags[i] = 1 | 2;
            atomicAdd(nactive, 1ull);
        } else {
            report(out, Q.src, hit, toi);
            Z.qflags[i] = flags | 4;
        }
    } else {
        if (!(flags & 2)) return;
        if (hit && toi == 0.0 && Q.co_tol > 1e-300) {
            if (10 * Q.co_tol < Q.ms) {
                Q.ms *= 0.5;
            } else {
                Q.co_tol *= 0.5;
                ti_tolerances(Q.s, Q.e, Q.is_vf, Q.co_tol, Q.tol);
            }
            Z.qtoi[i] = 0x7ff0000000000000ull;
            atomicAdd(nactive, 1ull);
        } else {
            report(out, Q.src, hit, toi * rescale);
            Z.qflags[i] = (flags & ~2) | 4;
        }
    }
}

// redo pass: forget the phase-1 result of every non-refined query
__global__ void k_ti_reset(int nq, const int* __restrict__ qflags, unsigned long long* qtoi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq && !(qflags[i] & 1)) qtoi[i] = 0x7ff0000000000000ull;
}
__global__ void k_ti_report_plain(int nq, TIQueue Z, CcdOut out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq || (Z.qflags[i] & 1)) return;
    const double toi = __longlong_as_double((long long)Z.qtoi[i]);
    report(out, Z.queries[i].src, toi < INFINITY, toi);
}
__global__ void k_set_bits(unsigned long long* dst, const unsigned long long* src, double cap)
{
    // dst = min(*src, cap) as a non-negative double bit pattern
    const double v = src ? fmin(__longlong_as_double((long long)*src), cap) : cap;
    *dst = (unsigned long long)__double_as_longlong(v);
}

struct TIWork {
    Buf<TIQuery> queries;
    Buf<unsigned long long> qtoi;
    Buf<int> qflags;
    Buf<TIUnit> ua, ub;
};
static std::map<ipcb_ctx*, TIWork*> g_work; // one per context

static unsigned long long read_counter(ipcb_ctx* ctx, const unsigned long long* d)
{
    IPCB_CUDA(cudaMemcpyAsync(ctx->pinned.p, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
    return (unsigned long long)ctx->pinned.p[0];
}
static double bits_to_double(unsigned long long b)
{
    double d;
    memcpy(&d, &b, sizeof d);
    return d;
}

// run the queue to exhaustion starting from `nunits` units in W.ua
static void ti_run_levels(ipcb_ctx* ctx, TIWork& W, TIQueue Z, unsigned long long nunits, unsigned long long* prune)
{
    cudaStream_t s = ctx->stream;
    unsigned long long* cnt = ctx->dCounters.p + 2;
    TIUnit* in = W.ua.p;
    TIUnit* outq = W.ub.p;
    for (int level = 0; nunits > 0; level++) {
        IPCB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), s));
        Z.in = in, Z.outq = outq, Z.nout = cnt;
        k_ti_level<<<grid_for(nunits, 128), 128, 0, s>>>(Z, nunits, prune, prune ? 1 : 0, level >= 100 ? 1 : 0);
        ctx->launches++;
        IPCB_CUDA(cudaGetLastError());
        nunits = std::min(read_counter(ctx, cnt), Z.ucap);
        std::swap(in, outq);
    }
}

// Tight-Inclusion CCD over one query source.  out.bound != nullptr selects step-size mode: *out.bound
// only ever receives FINAL times of impact; a separate pruning bound (seeded from it) is tightened
// by phase-1 terminal boxes.  Because a query that is later refined (toi < SMALL_TOI) may have
// tightened the pruning bound with a non-final value, the result is accepted only if it is <= the
// final pruning bound; otherwise the non-refined queries are searched once more against the (valid)
// result itself.
static void ti_run(ipcb_ctx* ctx, const QuerySource& src, double min_distance, double tmax, const ipcb_ccd_params& p, CcdOut out)
{
    if (src.n == 0) return;
    cudaStream_t s = ctx->stream;
    TIWork*& wp = g_work[ctx];
    if (!wp) wp = new TIWork();
    TIWork& W = *wp;
    unsigned long long* nq_d = ctx->dCounters.p + 1;
    unsigned long long* cnt = ctx->dCounters.p + 2;
    unsigned long long* nactive_d = ctx->dCounters.p + 3;
    unsigned long long* prune_d = ctx->dCounters.p + 4;
    const bool step_mode = out.bound != nullptr;
    size_t qcap = std::max<size_t>(W.queries.cap, std::max<size_t>(4096, size_t(src.n) / 16));
    unsigned long long nq = 0;
    for (int attempt = 0;; attempt++) {
        W.queries.reserve(qcap), W.qtoi.reserve(qcap), W.qflags.reserve(qcap);
        qcap = std::min(W.queries.cap, std::min(W.qtoi.cap, W.qflags.cap));
        IPCB_CUDA(cudaMemsetAsync(nq_d, 0, sizeof(unsigned long long), s));
        TIQueue Z0 { W.queries.p, W.qtoi.p, W.qflags.p, nq_d, (unsigned long long)qcap, nullptr, nullptr, nullptr, 0 };
        k_ti_level0<<<grid_for(src.n, 128), 128, 0, s>>>(src, min_distance, tmax, p.tolerance, p.conservative_rescaling, Z0, out);
        ctx->launches++;
        IPCB_CUDA(cudaGetLastError());
        nq = read_counter(ctx, nq_d);
        if (nq <= qcap) break;
        if (attempt > 2) throw Error("ccd: query buffer overflow persisted");
        qcap = nq + nq / 8; // level-0 reports are idempotent, so the pass can simply be repeated with room
    }
    if (nq == 0) return;
    const size_t want_u = std::max<size_t>(W.ua.cap, std::max<size_t>(1 << 16, 8 * size_t(nq)));
    W.ua.reserve(want_u), W.ub.reserve(want_u);
    TIQueue Z { W.queries.p, W.qtoi.p, W.qflags.p, nq_d, (unsigned long long)qcap, nullptr, nullptr, nullptr,
                (unsigned long long)std::min(W.ua.cap, W.ub.cap) };
    auto push_roots = [&](int need_mask, int exclude_mask) {
        IPCB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), s));
        k_ti_push_roots<<<grid_for(nq, 128), 128, 0, s>>>(int(nq), W.qflags.p, need_mask, exclude_mask, W.ua.p, cnt);
        ctx->launches++;
        return read_counter(ctx, cnt);
    };
    // ---- phase 1: minimum effective distance, no_zero_toi = false (tight_inclusion_ccd.cpp:45-52)
    if (step_mode) {
        k_set_bits<<<1, 1, 0, s>>>(prune_d, out.bound, tmax);
        ctx->launches++;
    }
    ti_run_levels(ctx, W, Z, push_roots(0, 0), step_mode ? prune_d : nullptr);
    IPCB_CUDA(cudaMemsetAsync(nactive_d, 0, sizeof(unsigned long long), s));
    k_ti_finalize<<<grid_for(nq, 128), 128, 0, s>>>(int(nq), Z, 0, min_distance, p.conservative_rescaling, out, nactive_d);
    ctx->launches++;
    unsigned long long nactive = read_counter(ctx, nactive_d);
    const bool refined = nactive > 0;
    // ---- phase 2: no-zero-toi refinement rounds (ms = min_distance; :59-72 + ticcd's shrink loop)
    for (int round = 0; nactive > 0 && round < 100; round++) {
        ti_run_levels(ctx, W, Z, push_roots(2, 0), nullptr);
        IPCB_CUDA(cudaMemsetAsync(nactive_d, 0, sizeof(unsigned long long), s));
        k_ti_finalize<<<grid_for(nq, 128), 128, 0, s>>>(int(nq), Z, 1, min_distance, p.conservative_rescaling, out, nactive_d);
        ctx->launches++;
        nactive = read_counter(ctx, nactive_d);
    }
    if (step_mode && refined) {
        const double result = bits_to_double(read_counter(ctx, out.bound));
        const double pruned = bits_to_double(read_counter(ctx, prune_d));
        if (result > pruned) { // pruning bound was tightened by a value that did not stay final
            k_ti_reset<<<grid_for(nq, 128), 128, 0, s>>>(int(nq), W.qflags.p, W.qtoi.p);
            k_set_bits<<<1, 1, 0, s>>>(prune_d, out.bound, tmax);
            ctx->launches += 2;
            ti_run_levels(ctx, W, Z, push_roots(0, 1), prune_d);
            k_ti_report_plain<<<grid_for(nq, 128), 128, 0, s>>>(int(nq), Z, out);
            ctx->launches++;
        }
    }
}

static QuerySource cand_source(ipcb_ctx* ctx, int kind)
{
    return { kind, ctx->cand[kind].count, ctx->cand[kind].pairs.p, ctx->dE.p, ctx->dF.p, ctx->X0.p, ctx->X1.p, nullptr, nullptr };
}

static void run_ccd(ipcb_ctx* ctx, const QuerySource& src, double min_distance, double tmax, const ipcb_ccd_params& p, CcdOut out)
{
    if (src.n == 0) return;
    if (p.kind == IPCB_CCD_ADDITIVE) {
        k_additive<<<grid_for(src.n, 128), 128, 0, ctx->stream>>>(src, min_distance, tmax, (long long)p.max_iterations,
                                                                 p.conservative_rescaling, out);
        ctx->launches++;
        IPCB_CUDA(cudaGetLastError());
    } else {
        ti_run(ctx, src, min_distance, tmax, p, out);
    }
}

// Candidates::compute_collision_free_stepsize (candidates.cpp:252-292): earliest TOI over the
// resident candidates, 1.0 when there are none; the result is left in *d_out (device)
void ccd_stepsize(ipcb_ctx* ctx, double min_distance, const ipcb_ccd_params& p, double* d_out)
{
    Stage st(ctx, "ccd_narrow");
    cudaStream_t s = ctx->stream;
    unsigned long long* bound = ctx->dCounters.p + 8;
    k_set_bits<<<1, 1, 0, s>>>(bound, nullptr, 1.0);
    ctx->launches++;
    CcdOut out { bound, nullptr, nullptr };
    // the cheap kinds first: their hits tighten the bound for the big EE / FV sets
    for (int kind : { IPCB_VV, IPCB_EV, IPCB_FV, IPCB_EE }) run_ccd(ctx, cand_source(ctx, kind), min_distance, 1.0, p, out);
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
    QuerySource src { kind, n, nullptr, nullptr, nullptr, nullptr, nullptr, a.p, b.p };
    CcdOut out { nullptr, hit.p, toi.p };
    run_ccd(ctx, src, min_distance, tmax, p, out);
    IPCB_CUDA(cudaMemcpyAsync(h_hit, hit.p, n, cudaMemcpyDeviceToHost, s));
    IPCB_CUDA(cudaMemcpyAsync(h_toi, toi.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    IPCB_CUDA(cudaStreamSynchronize(s));
}

} // namespace ipcb
