// collisions.cu — NormalCollisions::build(candidates, mesh, V, dhat, dmin).
//
// Replaces (reference src/ipc/): collisions/normal/normal_collisions.cpp:38-158
// and collisions/normal/normal_collisions_builder.cpp:26-336 (classification,
// dhat filter, reduction of every candidate to the VV / EV / EE / FV collision
// of its closest feature pair) and :547-689 (merge of duplicates with weight
// accumulation, weight == 0 dropped) — IPC collision set type.
//
// GPU formulation: one thread per candidate reads its ids (coalesced int2),
// gathers the stencil (one 32-byte sector per vertex), classifies in FP64 with
// the oracle's exact operation order, and appends a (key, weight[, eps, dtype])
// record to one of four streams with warp-aggregated atomics.  Each stream is
// then radix-sorted by key and run-length merged (the GPU equivalent of the
// reference's per-thread hash maps + serial merge), which also makes the
// output order canonical: sorted by (id0, id1).
#include "ctx.cuh"
#include "geom.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace ipcb {

struct StreamOut {
    unsigned long long* key;
    double* w;
    double* eps;
    unsigned char* dt;
    unsigned long long* counter;
};
struct ClassifyArgs {
    const int2* cand;
    int64_t n;
    const int2* E;
    const int4* F;
    const int4* F2E;
    const double4* X;
    const double4* rest;
    const double* vArea;
    const double* eArea;
    double offset_sqr;
    int area;
    StreamOut out[4];
};

__device__ inline d3 ld3(const double4* X, int i) { return load_vertex(X, i); }
__device__ inline unsigned long long mkkey(int a, int b) { return ((unsigned long long)(unsigned)a << 32) | (unsigned)b; }

__global__ void k_ids_to_keys(int64_t n, const int2* __restrict__ ids, int unordered, unsigned long long* __restrict__ key)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int2 c = ids[i];
    key[i] = unordered ? mkkey(min(c.x, c.y), max(c.x, c.y)) : mkkey(c.x, c.y);
}

// warp-aggregated append of one record per participating lane
__device__ inline void append(const StreamOut& o, bool pred, unsigned long long key, double w, double eps, unsigned char dt)
{
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(o.counter, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (pred) {
        const unsigned long long p = base + __popc(m & ((1u << lane) - 1));
        o.key[p] = key;
        o.w[p] = w;
        if (o.eps) o.eps[p] = eps;
        if (o.dt) o.dt[p] = dt;
    }
}

// MINB: resident blocks per SM the register allocation aims for (4 = 64 registers, 50 % occupancy: the kernel waits on
// vertex gathers, more warps hide more latency; a handful of spilled bytes for EE / FV)
template <int KIND, int MINB> __global__ void __launch_bounds__(256, MINB) k_classify(ClassifyArgs a)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    // 0 = nothing, 1 = VV, 2 = EV, 3 = own kind (EE / FV)
    int dest = 0;
    unsigned long long key = 0;
    double w = 1, eps = 0;
    unsigned char dt = 0;
    if (i < a.n) {
        const int2 c = a.cand[i];
        if (KIND == IPCB_VV) { // builder.cpp:26-56
            const double d = pp_dist(ld3(a.X, c.x), ld3(a.X, c.y));
            if (d < a.offset_sqr) {
                dest = 1;
                key = mkkey(min(c.x, c.y), max(c.x, c.y));
                w = a.area ? 0.5 * (a.vArea[c.x] + a.vArea[c.y]) : 1.0;
            }
        } else if (KIND == IPCB_EV) { // builder.cpp:58-138
            const int2 e = __ldg(a.E + c.x);
            const d3 p = ld3(a.X, c.y), e0 = ld3(a.X, e.x), e1 = ld3(a.X, e.y);
            const int t = point_edge_type(p, e0, e1);
            const d3 x[3] = { p, e0, e1 };
            const double d = sub_value(sub_point_edge(t), x);
            if (d < a.offset_sqr) {
                w = a.area ? 0.5 * a.vArea[c.y] : 1.0;
                if (t == PE_E) {
                    dest = 2, key = mkkey(c.x, c.y);
                } else {
                    const int vj = t == PE_E0 ? e.x : e.y;
                    dest = 1, key = mkkey(min(c.y, vj), max(c.y, vj));
                }
            }
        } else if (KIND == IPCB_EE) { // builder.cpp:140-237
            const int2 ea = __ldg(a.E + c.x), eb = __ldg(a.E + c.y);
            const d3 x[4] = { ld3(a.X, ea.x), ld3(a.X, ea.y), ld3(a.X, eb.x), ld3(a.X, eb.y) };
            const int actual = edge_edge_type(x[0], x[1], x[2], x[3]);
            const double d = sub_value(sub_edge_edge(actual), x);
            if (d < a.offset_sqr) {
                eps = moll_threshold(ld3(a.rest, ea.x), ld3(a.rest, ea.y), ld3(a.rest, eb.x), ld3(a.rest, eb.y));
                const double cr = sqn(cross(x[1] - x[0], x[3] - x[2]));
                const int t = cr < eps ? EE_AB : actual;
                w = a.area ? 0.25 * (a.eArea[c.x] + a.eArea[c.y]) : 1.0;
                const int va[2] = { ea.x, ea.y }, vb[2] = { eb.x, eb.y };
                if (t <= EE_A1B1) {
                    const int vi = va[t >> 1], vj = vb[t & 1];
                    dest = 1, key = mkkey(min(vi, vj), max(vi, vj));
                } else if (t == EE_AB0 || t == EE_AB1) {
                    dest = 2, key = mkkey(c.x, vb[t - EE_AB0]);
                } else if (t == EE_A0B || t == EE_A1B) {
                    dest = 2, key = mkkey(c.y, va[t - EE_A0B]);
                } else {
                    dest = 3, key = mkkey(c.x, c.y), dt = (unsigned char)actual;
                }
            }
        } else { // FV, builder.cpp:239-336
            const int4 f = __ldg(a.F + c.x);
            const d3 x[4] = { ld3(a.X, c.y), ld3(a.X, f.x), ld3(a.X, f.y), ld3(a.X, f.z) };
            const int t = point_triangle_type(x[0], x[1], x[2], x[3]);
            const double d = sub_value(sub_point_triangle(t), x);
            if (d < a.offset_sqr) {
                w = a.area ? 0.25 * a.vArea[c.y] : 1.0;
                if (t <= PT_T2) {
                    const int vj = t == PT_T0 ? f.x : (t == PT_T1 ? f.y : f.z);
                    dest = 1, key = mkkey(min(c.y, vj), max(c.y, vj));
                } else if (t <= PT_E2) {
                    const int4 fe = __ldg(a.F2E + c.x);
                    const int ej = t == PT_E0 ? fe.x : (t == PT_E1 ? fe.y : fe.z);
                    dest = 2, key = mkkey(ej, c.y);
                } else {
                    dest = 3, key = mkkey(c.x, c.y);
                }
            }
        }
    }
    // warp-aggregated append to the (up to) three destination streams with ONE atomic round trip: lanes 0..2 reserve
    // the ranges of destinations 1..3 in the same instruction, instead of three dependent reserve-then-store rounds
    const int lane = threadIdx.x & 31;
    const unsigned m1 = __ballot_sync(0xffffffffu, dest == 1), m2 = __ballot_sync(0xffffffffu, dest == 2),
                   m3 = __ballot_sync(0xffffffffu, dest == 3);
    if ((m1 | m2 | m3) == 0) return;
    constexpr int OWN = KIND == IPCB_EE ? IPCB_EE : IPCB_FV; // stream of destination 3 (unused for VV / EV candidates)
    unsigned long long base = 0;
    {
        const unsigned mine = lane == 0 ? m1 : (lane == 1 ? m2 : m3);
        unsigned long long* counter = lane == 0 ? a.out[IPCB_VV].counter : (lane == 1 ? a.out[IPCB_EV].counter : a.out[OWN].counter);
        if (lane < 3 && mine) base = atomicAdd(counter, (unsigned long long)__popc(mine));
    }
    const unsigned long long b1 = __shfl_sync(0xffffffffu, base, 0), b2 = __shfl_sync(0xffffffffu, base, 1),
                             b3 = __shfl_sync(0xffffffffu, base, 2);
    if (dest == 0) return;
    const unsigned below = (1u << lane) - 1;
    const unsigned long long p = dest == 1 ? b1 + __popc(m1 & below) : (dest == 2 ? b2 + __popc(m2 & below) : b3 + __popc(m3 & below));
    unsigned long long* okey = dest == 1 ? a.out[IPCB_VV].key : (dest == 2 ? a.out[IPCB_EV].key : a.out[OWN].key);
    double* ow = dest == 1 ? a.out[IPCB_VV].w : (dest == 2 ? a.out[IPCB_EV].w : a.out[OWN].w);
    okey[p] = key;
    ow[p] = w;
    if (KIND == IPCB_EE && dest == 3) a.out[IPCB_EE].eps[p] = eps, a.out[IPCB_EE].dt[p] = dt;
}

// ---- merge: sorted keys -> unique records with accumulated weights ---------------
__global__ void k_iota(int64_t n, int* __restrict__ idx)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) idx[i] = int(i);
}
// head of a run (or every element when merge == 0); head threads sum the run's
// weights and flag the record as kept when the sum is non-zero (builder.cpp:668-688)
__global__ void k_runs(int64_t n, const unsigned long long* __restrict__ key, const int* __restrict__ idx,
                       const double* __restrict__ w_raw, int merge, int* __restrict__ keep, double* __restrict__ wsum)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = key[i];
    if (!merge) {
        keep[i] = 1;
        wsum[i] = w_raw[idx[i]];
        return;
    }
    if (i > 0 && key[i - 1] == k) {
        keep[i] = 0;
        return;
    }
    double s = 0;
    for (int64_t j = i; j < n && key[j] == k; j++) s += w_raw[idx[j]];
    wsum[i] = s;
    keep[i] = s != 0.0;
}
__global__ void k_emit_collisions(int64_t n, const unsigned long long* __restrict__ key, const int* __restrict__ idx,
                                  const int* __restrict__ keep, const int* __restrict__ pos, const double* __restrict__ wsum,
                                  const double* __restrict__ eps_raw, const unsigned char* __restrict__ dt_raw, int2* __restrict__ ids,
                                  double* __restrict__ w, double* __restrict__ eps, unsigned char* __restrict__ dt)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const int p = pos[i];
    const unsigned long long k = key[i];
    ids[p] = make_int2(int(k >> 32), int(k & 0xffffffffu));
    w[p] = wsum[i];
    if (eps_raw) {
        eps[p] = eps_raw[idx[i]];
        dt[p] = dt_raw[idx[i]];
    }
}

// Keys are (id0 << 32 | id1) with id0, id1 below these bounds per kind: the radix sort only has to look at
// the low bits of id1 and of id0 (two bit ranges would need two sorts; one range from 0 to 32 + bits(id0) with
// id1 < 2^32 is what cub offers, so the saving comes from the id0 bits only)
static int key_bits(const ipcb_ctx* ctx, int kind)
{
    const int m = kind == IPCB_VV ? ctx->nV : (kind == IPCB_FV ? ctx->nF : ctx->nE);
    int b = 1;
    while ((1ll << b) < m) b++;
    return 32 + b;
}

// enqueue the merge of one stream on `s`: sort by key, run heads + weight sums, scan, emit; the count is read back
// into pinned slots 20 + 2 * kind (pos of the last record) and 21 + 2 * kind (its head flag)
static void merge_stream_enqueue(ipcb_ctx* ctx, int kind, int64_t n, cudaStream_t s)
{
    CollisionSet& cs = ctx->coll[kind];
    cs.count = 0;
    if (n == 0) return;
    cs.idx_raw.reserve(n), cs.idx_sorted.reserve(n), cs.key_sorted.reserve(n), cs.head.reserve(n), cs.pos.reserve(n + 1);
    Buf<double>& wsum = cs.wsum;
    wsum.reserve(n);
    k_iota<<<grid_for(n, 256), 256, 0, s>>>(n, cs.idx_raw.p);
    size_t bytes = 0, bytes2 = 0;
    const int bits = key_bits(ctx, kind);
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, cs.key_raw.p, cs.key_sorted.p, cs.idx_raw.p, cs.idx_sorted.p, n, 0, bits, s);
    cub::DeviceScan::ExclusiveSum(nullptr, bytes2, cs.head.p, cs.pos.p, n, s);
    cs.cubtmp.reserve(std::max(bytes, bytes2));
    cub::DeviceRadixSort::SortPairs(cs.cubtmp.p, bytes, cs.key_raw.p, cs.key_sorted.p, cs.idx_raw.p, cs.idx_sorted.p, n, 0, bits, s);
    const int merge = kind != IPCB_FV; // fv collisions are appended, never merged (builder.cpp:659-661)
    k_runs<<<grid_for(n, 256), 256, 0, s>>>(n, cs.key_sorted.p, cs.idx_sorted.p, cs.w_raw.p, merge, cs.head.p, wsum.p);
    cub::DeviceScan::ExclusiveSum(cs.cubtmp.p, bytes2, cs.head.p, cs.pos.p, n, s);
    cs.ids.reserve(n), cs.w.reserve(n);
    if (kind == IPCB_EE) cs.eps.reserve(n), cs.dtype.reserve(n);
    k_emit_collisions<<<grid_for(n, 256), 256, 0, s>>>(n, cs.key_sorted.p, cs.idx_sorted.p, cs.head.p, cs.pos.p, wsum.p,
                                                       kind == IPCB_EE ? cs.eps_raw.p : nullptr, cs.dt_raw.p, cs.ids.p, cs.w.p,
                                                       cs.eps.p, cs.dtype.p);
    ctx->launches += 5 + (bits + 7) / 8 + 4;
    // count = pos[n-1] + head[n-1]
    IPCB_CUDA(cudaMemcpyAsync(&ctx->pinned.p[20 + 2 * kind], cs.pos.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    IPCB_CUDA(cudaMemcpyAsync(&ctx->pinned.p[21 + 2 * kind], cs.head.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
}

static void merge_streams(ipcb_ctx* ctx, const int64_t raw[4], bool disjoint = false);

void collisions_build(ipcb_ctx* ctx, double dhat, double dmin, int flags)
{
    if (flags & IPCB_SET_IMPROVED_MAX_APPROX)
        throw Error("CollisionSetType::IMPROVED_MAX_APPROX is not implemented in the CUDA library yet (IPC set type only)");
    cudaStream_t s = ctx->stream;
    int64_t total = 0;
    for (auto& c : ctx->cand) total += c.count;
    ctx->dmin = dmin;
    ctx->coll_valid = true;
    for (auto& c : ctx->coll) c.count = 0;
    if (total == 0) return;
    if (total > 0x7fffffffll) throw Error("more than 2^31 candidates in one collision build");
    {
        Stage st(ctx, "classify");
        // worst case every candidate lands in one stream
        const int64_t cap_vv = total, cap_ev = total - ctx->cand[IPCB_VV].count, cap_ee = ctx->cand[IPCB_EE].count,
                      cap_fv = ctx->cand[IPCB_FV].count;
        const int64_t caps[4] = { cap_vv, cap_ev, cap_ee, cap_fv };
        ClassifyArgs a;
        a.E = ctx->dE.p, a.F = ctx->dF.p, a.F2E = ctx->dF2E.p, a.X = ctx->X0.p, a.rest = ctx->dRest.p;
        a.vArea = ctx->dVArea.p, a.eArea = ctx->dEArea.p;
        a.offset_sqr = (dmin + dhat) * (dmin + dhat);
        a.area = (flags & IPCB_USE_AREA_WEIGHTING) ? 1 : 0;
        IPCB_CUDA(cudaMemsetAsync(ctx->dCounters.p, 0, 8 * sizeof(unsigned long long), s));
        for (int k = 0; k < 4; k++) {
            CollisionSet& cs = ctx->coll[k];
            cs.key_raw.reserve(caps[k]), cs.w_raw.reserve(caps[k]);
            if (k == IPCB_EE) cs.eps_raw.reserve(caps[k]), cs.dt_raw.reserve(caps[k]);
            a.out[k] = { cs.key_raw.p, cs.w_raw.p, k == IPCB_EE ? cs.eps_raw.p : nullptr, k == IPCB_EE ? cs.dt_raw.p : nullptr,
                         ctx->dCounters.p + 1 + k };
        }
        ctx->fork(); // the face-vertex candidates are classified concurrently with the edge-edge ones
        for (int k = 0; k < 4; k++) {
            a.cand = ctx->cand[k].pairs.p;
            a.n = ctx->cand[k].count;
            if (a.n == 0) continue;
            const unsigned grid = grid_for(a.n, 256);
            static const bool sparse = getenv("IPCB_CLASSIFY_SPARSE") != nullptr; // A/B switch: 3 resident blocks, no spill
            if (k == IPCB_VV) k_classify<IPCB_VV, 4><<<grid, 256, 0, s>>>(a);
            if (k == IPCB_EV) k_classify<IPCB_EV, 4><<<grid, 256, 0, s>>>(a);
            if (k == IPCB_EE && sparse) k_classify<IPCB_EE, 3><<<grid, 256, 0, s>>>(a);
            else if (k == IPCB_EE) k_classify<IPCB_EE, 4><<<grid, 256, 0, s>>>(a);
            if (k == IPCB_FV && sparse) k_classify<IPCB_FV, 3><<<grid, 256, 0, ctx->aux[0]>>>(a);
            else if (k == IPCB_FV) k_classify<IPCB_FV, 4><<<grid, 256, 0, ctx->aux[0]>>>(a);
            ctx->launches++;
        }
        ctx->join(0);
        IPCB_CUDA(cudaGetLastError());
        IPCB_CUDA(cudaMemcpyAsync(ctx->pinned.p, ctx->dCounters.p + 1, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        IPCB_CUDA(cudaStreamSynchronize(s));
    }
    const int64_t raw[4] = { ctx->pinned.p[0], ctx->pinned.p[1], ctx->pinned.p[2], ctx->pinned.p[3] };
    merge_streams(ctx, raw);
}

// records -> final arrays without sorting or merging (every record is kept)
__global__ void k_keep_all(int64_t n, const unsigned long long* __restrict__ key, const double* __restrict__ w_raw,
                           const double* __restrict__ eps_raw, const unsigned char* __restrict__ dt_raw, int2* __restrict__ ids,
                           double* __restrict__ w, double* __restrict__ eps, unsigned char* __restrict__ dt)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = key[i];
    ids[i] = make_int2(int(k >> 32), int(k & 0xffffffffu));
    w[i] = w_raw[i];
    if (eps_raw) eps[i] = eps_raw[i], dt[i] = dt_raw[i];
}

// disjoint: the builders come from disjoint candidate shards, so edge-edge and face-vertex records are unique across
// builders and only VV / EV records need uniting: EE / FV are concatenated in builder order (each builder's records
// are sorted; the canonical global order is restored on demand by collisions_sort)
static void merge_streams(ipcb_ctx* ctx, const int64_t raw[4], bool disjoint)
{
    cudaStream_t s = ctx->stream;
    Stage st(ctx, "merge_collisions");
    // the four streams are merged concurrently (their sorts are small and latency-bound): EE on the main stream
    cudaStream_t where[4] = { ctx->aux[0], ctx->aux[1], s, ctx->aux[2] };
    ctx->fork();
    for (int k = 0; k < 4; k++) {
        CollisionSet& cs = ctx->coll[k];
        cs.sorted = true;
        if (disjoint && (k == IPCB_EE || k == IPCB_FV)) {
            const int64_t n = raw[k];
            cs.count = n;
            cs.sorted = false;
            if (n == 0) continue;
            cs.ids.reserve(n), cs.w.reserve(n);
            if (k == IPCB_EE) cs.eps.reserve(n), cs.dtype.reserve(n);
            k_keep_all<<<grid_for(n, 256), 256, 0, where[k]>>>(n, cs.key_raw.p, cs.w_raw.p, k == IPCB_EE ? cs.eps_raw.p : nullptr, cs.dt_raw.p,
                                                             cs.ids.p, cs.w.p, cs.eps.p, cs.dtype.p);
            ctx->launches++;
        } else {
            merge_stream_enqueue(ctx, k, raw[k], where[k]);
        }
    }
    for (int k = 0; k < 4; k++) {
        if (raw[k] == 0 || !ctx->coll[k].sorted) continue;
        IPCB_CUDA(cudaStreamSynchronize(where[k]));
        ctx->coll[k].count = int64_t(*reinterpret_cast<int*>(&ctx->pinned.p[20 + 2 * k])) + int64_t(*reinterpret_cast<int*>(&ctx->pinned.p[21 + 2 * k]));
    }
    for (int k = 0; k < ipcb_ctx::NAUX; k++) ctx->join(k);
    IPCB_CUDA(cudaGetLastError());
}

// canonical order (sorted ids) of a kind that was concatenated by a disjoint merge: its own records go through the
// regular sort + run merge once (nothing is united: edge-edge / face-vertex records of disjoint shards are unique)
void collisions_sort(ipcb_ctx* ctx, int kind)
{
    CollisionSet& cs = ctx->coll[kind];
    if (cs.sorted || cs.count == 0) {
        cs.sorted = true;
        return;
    }
    cudaStream_t s = ctx->stream;
    const int64_t n = cs.count;
    cs.key_raw.reserve(n), cs.w_raw.reserve(n);
    if (kind == IPCB_EE) cs.eps_raw.reserve(n), cs.dt_raw.reserve(n);
    k_ids_to_keys<<<grid_for(n, 256), 256, 0, s>>>(n, cs.ids.p, kind == IPCB_VV || kind == IPCB_EE, cs.key_raw.p);
    IPCB_CUDA(cudaMemcpyAsync(cs.w_raw.p, cs.w.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
    if (kind == IPCB_EE) {
        IPCB_CUDA(cudaMemcpyAsync(cs.eps_raw.p, cs.eps.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
        IPCB_CUDA(cudaMemcpyAsync(cs.dt_raw.p, cs.dtype.p, n, cudaMemcpyDeviceToDevice, s));
    }
    ctx->launches++;
    merge_stream_enqueue(ctx, kind, n, s);
    IPCB_CUDA(cudaStreamSynchronize(s));
    cs.count = int64_t(*reinterpret_cast<int*>(&ctx->pinned.p[20 + 2 * kind])) + int64_t(*reinterpret_cast<int*>(&ctx->pinned.p[21 + 2 * kind]));
    cs.sorted = true;
}

// ---- public containers + NormalCollisionsBuilder::merge (normal_collisions.hpp:177-189, builder.cpp:547-689) -------
// collisions_clear / collisions_append / collisions_merge: records of several builders (the ranks of a sharded
// build) are appended to the raw streams as (key, weight[, eps, dtype]) and merged by the same sort + run-merge
// that finishes collisions_build.
// one launch appends all four kinds of a packed buffer (include/ipcb200.h: collisions_pack_dev layout)
struct UnpackArgs {
    const char* in;
    int64_t n[4], first[4]; // records per kind, first thread of the kind
    int64_t ids[4], w[4], eps, dt; // byte offsets inside the buffer
    unsigned long long* key[4];
    double* wout[4];
    double* eps_out;
    unsigned char* dt_out;
};
__global__ void k_unpack(UnpackArgs a, int64_t total)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int k = t >= a.first[3] ? 3 : (t >= a.first[2] ? 2 : (t >= a.first[1] ? 1 : 0));
    const int64_t i = t - a.first[k];
    const int2 c = reinterpret_cast<const int2*>(a.in + a.ids[k])[i];
    a.key[k][i] = (k == IPCB_VV || k == IPCB_EE) ? mkkey(min(c.x, c.y), max(c.x, c.y)) : mkkey(c.x, c.y);
    a.wout[k][i] = reinterpret_cast<const double*>(a.in + a.w[k])[i];
    if (k == IPCB_EE) {
        a.eps_out[i] = reinterpret_cast<const double*>(a.in + a.eps)[i];
        a.dt_out[i] = reinterpret_cast<const unsigned char*>(a.in + a.dt)[i];
    }
}
void collisions_append_packed_dev(ipcb_ctx* ctx, const void* d_buffer, const int64_t n[4], const int64_t ids_off[4], const int64_t w_off[4],
                                  int64_t eps_off, int64_t dt_off)
{
    cudaStream_t s = ctx->stream;
    UnpackArgs a;
    a.in = static_cast<const char*>(d_buffer);
    int64_t total = 0;
    for (int k = 0; k < 4; k++) {
        CollisionSet& cs = ctx->coll[k];
        const size_t have = size_t(cs.raw_count), want = have + size_t(n[k]);
        if (want > 0x7fffffffull) throw Error("more than 2^31 collision records of one kind");
        cs.key_raw.reserve_keep(want, have, s), cs.w_raw.reserve_keep(want, have, s);
        if (k == IPCB_EE) cs.eps_raw.reserve_keep(want, have, s), cs.dt_raw.reserve_keep(want, have, s);
        a.n[k] = n[k], a.first[k] = total, a.ids[k] = ids_off[k], a.w[k] = w_off[k];
        a.key[k] = cs.key_raw.p + have, a.wout[k] = cs.w_raw.p + have;
        if (k == IPCB_EE) a.eps_out = cs.eps_raw.p + have, a.dt_out = cs.dt_raw.p + have;
        total += n[k];
        cs.raw_count = int64_t(want);
    }
    a.eps = eps_off, a.dt = dt_off;
    if (total == 0) return;
    k_unpack<<<grid_for(total, 256), 256, 0, s>>>(a, total);
    ctx->launches++;
    IPCB_CUDA(cudaGetLastError());
}

void collisions_clear(ipcb_ctx* ctx)
{
    for (auto& c : ctx->coll) c.count = 0, c.raw_count = 0;
    ctx->coll_valid = false;
}
void collisions_append_dev(ipcb_ctx* ctx, int kind, int64_t n, const int32_t* d_ids, const double* d_w, const double* d_eps,
                           const uint8_t* d_dt)
{
    if (n == 0) return;
    cudaStream_t s = ctx->stream;
    CollisionSet& cs = ctx->coll[kind];
    const size_t have = size_t(cs.raw_count), want = have + size_t(n);
    if (want > 0x7fffffffull) throw Error("more than 2^31 collision records of one kind");
    cs.key_raw.reserve_keep(want, have, s), cs.w_raw.reserve_keep(want, have, s);
    if (kind == IPCB_EE) cs.eps_raw.reserve_keep(want, have, s), cs.dt_raw.reserve_keep(want, have, s);
    k_ids_to_keys<<<grid_for(n, 256), 256, 0, s>>>(n, reinterpret_cast<const int2*>(d_ids), kind == IPCB_VV || kind == IPCB_EE, cs.key_raw.p + have);
    ctx->launches++;
    IPCB_CUDA(cudaMemcpyAsync(cs.w_raw.p + have, d_w, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
    if (kind == IPCB_EE) {
        if (!d_eps || !d_dt) throw Error("edge-edge collision records need eps_x and dtype");
        IPCB_CUDA(cudaMemcpyAsync(cs.eps_raw.p + have, d_eps, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
        IPCB_CUDA(cudaMemcpyAsync(cs.dt_raw.p + have, d_dt, n, cudaMemcpyDeviceToDevice, s));
    }
    IPCB_CUDA(cudaGetLastError());
    cs.raw_count = int64_t(want);
}
void collisions_merge(ipcb_ctx* ctx, double dmin, int flags)
{
    ctx->dmin = dmin;
    ctx->coll_valid = true;
    const int64_t raw[4] = { ctx->coll[0].raw_count, ctx->coll[1].raw_count, ctx->coll[2].raw_count, ctx->coll[3].raw_count };
    for (auto& c : ctx->coll) c.count = 0, c.raw_count = 0;
    merge_streams(ctx, raw, (flags & IPCB_MERGE_DISJOINT_SHARDS) != 0);
}

// ---- compute_minimum_distance (normal_collisions.cpp:209-233) ----------------------
__device__ inline void stencil_points(int kind, int2 id, const int2* E, const int4* F, const double4* X, d3* x)
{
    if (kind == IPCB_VV) {
        x[0] = ld3(X, id.x), x[1] = ld3(X, id.y);
    } else if (kind == IPCB_EV) {
        const int2 e = __ldg(E + id.x);
        x[0] = ld3(X, id.y), x[1] = ld3(X, e.x), x[2] = ld3(X, e.y);
    } else if (kind == IPCB_EE) {
        const int2 ea = __ldg(E + id.x), eb = __ldg(E + id.y);
        x[0] = ld3(X, ea.x), x[1] = ld3(X, ea.y), x[2] = ld3(X, eb.x), x[3] = ld3(X, eb.y);
    } else {
        const int4 f = __ldg(F + id.x);
        x[0] = ld3(X, id.y), x[1] = ld3(X, f.x), x[2] = ld3(X, f.y), x[3] = ld3(X, f.z);
    }
}
__global__ void k_min_distance(int kind, int64_t n, const int2* __restrict__ ids, const unsigned char* __restrict__ dt,
                               const int2* E, const int4* F, const double4* X, unsigned long long* out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double d = INFINITY;
    if (i < n) {
        d3 x[4];
        stencil_points(kind, ids[i], E, F, X, x);
        const Sub s = kind == IPCB_VV ? Sub { 0, 0, 1, 0, 0 }
            : kind == IPCB_EV         ? sub_point_edge(PE_E)
            : kind == IPCB_EE         ? sub_edge_edge(dt[i])
                                      : sub_point_triangle(PT_T);
        d = sub_value(s, x);
    }
    for (int o = 16; o > 0; o >>= 1) d = fmin(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0 && d < INFINITY) atomicMin(out, (unsigned long long)__double_as_longlong(fmax(d, 0.0)));
}
double collisions_min_distance(ipcb_ctx* ctx)
{
    cudaStream_t s = ctx->stream;
    const unsigned long long inf_bits = 0x7ff0000000000000ull;
    IPCB_CUDA(cudaMemcpyAsync(ctx->dCounters.p + 8, &inf_bits, sizeof inf_bits, cudaMemcpyHostToDevice, s));
    for (int k = 0; k < 4; k++) {
        const CollisionSet& cs = ctx->coll[k];
        if (cs.count == 0) continue;
        k_min_distance<<<grid_for(cs.count, 256), 256, 0, s>>>(k, cs.count, cs.ids.p, cs.dtype.p, ctx->dE.p, ctx->dF.p, ctx->X0.p,
                                                              ctx->dCounters.p + 8);
        ctx->launches++;
    }
    unsigned long long bits = 0;
    IPCB_CUDA(cudaMemcpyAsync(&bits, ctx->dCounters.p + 8, sizeof bits, cudaMemcpyDeviceToHost, s));
    IPCB_CUDA(cudaStreamSynchronize(s));
    double d;
    memcpy(&d, &bits, sizeof d);
    return d;
}

} // namespace ipcb
