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
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

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

// record ids -> merge keys.  mode 0: ordered pair (edge-vertex, face-vertex); 1: unordered pair (vertex-vertex);
// 2: edge-edge, TYPED key (unordered edge pair, distance type, orientation bit — ee_typed_key below): a record whose
// edges are stored as (max, min) keeps its orientation, so its distance type never has to be re-expressed
__global__ void k_ids_to_keys(int64_t n, const int2* __restrict__ ids, int mode, const unsigned char* __restrict__ dt,
                              unsigned long long* __restrict__ key);

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
// [emu-begin merge]  (tests/test_kernel_emulation.py runs the kernels between these tags on the host, one lane per warp)
__global__ void k_iota(int64_t n, int* __restrict__ idx)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) idx[i] = int(i);
}
// head of a run (or every element when merge == 0); head threads sum the run's
// weights and flag the record as kept when the sum is non-zero (builder.cpp:668-688)
// Sum of the weights of one run of equal keys, ADDED IN ASCENDING ORDER OF VALUE: the records of a run arrive in the order of
// the warp-aggregated appends (not reproducible from run to run), and with area weighting they differ in value, so a sum
// in arrival order is not bit-reproducible.  Runs are short (one to a few records); the selection loop is O(length x
// distinct values).
template <class KeyEq> __device__ inline double run_weight_sum(int64_t i, int64_t n, const int* __restrict__ idx, const double* __restrict__ w_raw, KeyEq same)
{
    int64_t len = 1;
    while (i + len < n && same(i + len)) len++;
    if (len == 1) return w_raw[idx[i]];
    double s = 0, last = -INFINITY;
    for (int64_t done = 0; done < len;) {
        double m = INFINITY;
        int64_t c = 0;
        for (int64_t j = i; j < i + len; j++) {
            const double w = w_raw[idx[j]];
            if (w > last && w < m) m = w, c = 1;
            else if (w == m) c++;
        }
        for (int64_t r = 0; r < c; r++) s += m;
        last = m, done += c;
    }
    return s;
}
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
    const double s = run_weight_sum(i, n, idx, w_raw, [&](int64_t j) { return key[j] == k; });
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

// [emu-end merge]

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

static void merge_streams(ipcb_ctx* ctx, const int64_t raw[4], bool disjoint = false, bool typed_ee = false);
static void improved_max_approx_corrections(ipcb_ctx* ctx, double offset_sqr, bool area, int64_t raw[4]);
static void improved_max_approx_keys(ipcb_ctx* ctx, double offset_sqr, int64_t nu[4]);

void collisions_build(ipcb_ctx* ctx, double dhat, double dmin, int flags, bool may_defer)
{
    const bool improved = (flags & IPCB_SET_IMPROVED_MAX_APPROX) != 0;
    ctx->ima_pending = false;
    const bool defer = improved && ((flags & IPCB_DEFER_CORRECTIONS) || (ctx->shard_world > 1 && may_defer));
    if (improved && ctx->shard_world > 1 && !defer) // the corrections need the sub-element keys of every rank
        throw Error("CollisionSetType::IMPROVED_MAX_APPROX on a sharded context needs the deferred build (IPCB_DEFER_CORRECTIONS or "
                    "collisions_build_dev) followed by the exchange of the sub-element keys (collisions_corrections_*)");
    cudaStream_t s = ctx->stream;
    int64_t total = 0;
    for (auto& c : ctx->cand) total += c.count;
    ctx->dmin = dmin;
    ctx->coll_valid = true;
    for (auto& c : ctx->coll) c.count = 0;
    if (total == 0) {
        if (defer) { // this rank has nothing, the others may: it still takes part in the exchange
            for (int k = 0; k < 4; k++) ctx->ima_raw[k] = ctx->ima_nu[k] = 0;
            ctx->ima_area = (flags & IPCB_USE_AREA_WEIGHTING) != 0;
            ctx->ima_pending = true;
            ctx->coll_valid = false;
            IPCB_CUDA(cudaMemsetAsync(ctx->dCounters.p, 0, 8 * sizeof(unsigned long long), s));
        }
        return;
    }
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
            static const char* const names[4] = { "k:k_classify<VV>", "k:k_classify<EV>", "k:k_classify<EE>", "k:k_classify<FV>" };
            Stage kt(ctx, names[k], k == IPCB_FV ? ctx->aux[0] : s);
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
    int64_t raw[4] = { ctx->pinned.p[0], ctx->pinned.p[1], ctx->pinned.p[2], ctx->pinned.p[3] };
    if (defer) {
        // A sub-element pair can be derived from candidates of several ranks, and its correction records must be added ONCE:
        // the build stops after the rank's unique sub-element keys; the ranks exchange them (collisions_corrections_keys_dev /
        // _pack_dev), and collisions_corrections_apply_dev finishes the rank's set from the united lists.
        improved_max_approx_keys(ctx, (dmin + dhat) * (dmin + dhat), ctx->ima_nu);
        for (int k = 0; k < 4; k++) ctx->ima_raw[k] = raw[k];
        ctx->ima_area = (flags & IPCB_USE_AREA_WEIGHTING) != 0;
        ctx->ima_pending = true;
        ctx->coll_valid = false;
        return;
    }
    if (improved) improved_max_approx_corrections(ctx, (dmin + dhat) * (dmin + dhat), (flags & IPCB_USE_AREA_WEIGHTING) != 0, raw);
    // IPC set: every edge-edge / face-vertex record comes from its own candidate, so those two streams hold no
    // duplicates and need no merge — only vertex-vertex / edge-vertex records are united.  Their canonical order
    // (sorted ids) is what a caller SEES, not what the potential needs, so it could be restored lazily when somebody
    // fetches the records (collisions_sort; IPCB_LAZY_SORT=1).  Measured on C3 that is a wash — the 0.35 ms the two large
    // sorts cost come back as better locality of the energy / gradient / symbolic kernels on id-ordered records — and the
    // sorted order makes every later sum reproducible run to run, so the eager sort stays the default.
    // On a sharded context the rank's records are concatenated with the other ranks' ones right away (disjoint merge),
    // so sorting the shard first would be wasted work: lazy there.
    static const bool lazy = getenv("IPCB_LAZY_SORT") != nullptr;
    merge_streams(ctx, raw, !improved && (lazy || ctx->shard_world > 1), improved);
}

// records -> final arrays without sorting or merging (every record is kept)
// typed: edge-edge keys are [min edge : 28][max edge : 28][distance type : 4][stored as (max, min) : 1] (ee_typed_key)
__global__ void k_keep_all(int64_t n, const unsigned long long* __restrict__ key, const double* __restrict__ w_raw,
                           const double* __restrict__ eps_raw, const unsigned char* __restrict__ dt_raw, int typed, int2* __restrict__ ids,
                           double* __restrict__ w, double* __restrict__ eps, unsigned char* __restrict__ dt)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = key[i];
    if (typed) {
        const int lo = int(k >> 33), hi = int((k >> 5) & 0xfffffffull);
        ids[i] = (k & 1ull) ? make_int2(hi, lo) : make_int2(lo, hi);
    } else {
        ids[i] = make_int2(int(k >> 32), int(k & 0xffffffffu));
    }
    w[i] = w_raw[i];
    if (eps_raw) eps[i] = eps_raw[i], dt[i] = dt_raw[i];
}

// disjoint: the builders come from disjoint candidate shards, so edge-edge and face-vertex records are unique across
// builders and only VV / EV records need uniting: EE / FV are concatenated in builder order (each builder's records
// are sorted; the canonical global order is restored on demand by collisions_sort)
static void merge_ee_typed_enqueue(ipcb_ctx* ctx, int64_t n, cudaStream_t s);
static void merge_streams(ipcb_ctx* ctx, const int64_t raw[4], bool disjoint, bool typed_ee)
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
                                                             typed_ee && k == IPCB_EE, cs.ids.p, cs.w.p, cs.eps.p, cs.dtype.p);
            ctx->launches++;
        } else if (typed_ee && k == IPCB_EE) {
            merge_ee_typed_enqueue(ctx, raw[k], where[k]);
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
    const bool typed = kind == IPCB_EE && ctx->nE < (1 << 28); // beyond: only IPC records exist (always (min, max) ordered)
    k_ids_to_keys<<<grid_for(n, 256), 256, 0, s>>>(n, cs.ids.p, typed ? 2 : (kind == IPCB_VV || kind == IPCB_EE ? 1 : 0), cs.dtype.p, cs.key_raw.p);
    IPCB_CUDA(cudaMemcpyAsync(cs.w_raw.p, cs.w.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
    if (kind == IPCB_EE) {
        IPCB_CUDA(cudaMemcpyAsync(cs.eps_raw.p, cs.eps.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
        IPCB_CUDA(cudaMemcpyAsync(cs.dt_raw.p, cs.dtype.p, n, cudaMemcpyDeviceToDevice, s));
    }
    ctx->launches++;
    if (typed) merge_ee_typed_enqueue(ctx, n, s);
    else merge_stream_enqueue(ctx, kind, n, s);
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
__global__ void k_unpack(UnpackArgs a, int64_t total);
void collisions_append_packed_dev(ipcb_ctx* ctx, const void* d_buffer, const int64_t n[4], const int64_t ids_off[4], const int64_t w_off[4],
                                  int64_t eps_off, int64_t dt_off)
{
    cudaStream_t s = ctx->stream;
    UnpackArgs a;
    a.in = static_cast<const char*>(d_buffer);
    if (n[IPCB_EE] && ctx->nE >= (1 << 28)) throw Error("edge-edge collision records: more than 2^28 edges");
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
    if (kind == IPCB_EE && (!d_eps || !d_dt)) throw Error("edge-edge collision records need eps_x and dtype");
    if (kind == IPCB_EE && ctx->nE >= (1 << 28)) throw Error("edge-edge collision records: more than 2^28 edges");
    k_ids_to_keys<<<grid_for(n, 256), 256, 0, s>>>(n, reinterpret_cast<const int2*>(d_ids), kind == IPCB_EE ? 2 : (kind == IPCB_VV ? 1 : 0), d_dt,
                                                   cs.key_raw.p + have);
    ctx->launches++;
    IPCB_CUDA(cudaMemcpyAsync(cs.w_raw.p + have, d_w, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
    if (kind == IPCB_EE) {
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
    // appended edge-edge records carry TYPED keys: records are united on (unordered edge pair, distance type) like
    // EdgeEdgeNormalCollision::operator== (collisions/normal/edge_edge.cpp:123-142) and keep their stored orientation
    merge_streams(ctx, raw, (flags & IPCB_MERGE_DISJOINT_SHARDS) != 0, true);
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


// =====================================================================================================================
// CollisionSetType::IMPROVED_MAX_APPROX (normal_collisions.cpp:84-128).  After the IPC classification pass:
//  1. sub-element candidates are derived from the element candidates (candidates.cpp:584-695): the ACTIVE
//     vertex-vertex / edge-vertex pairs of every edge-vertex, edge-edge and face-vertex candidate — emitted as 64-bit
//     keys, radix-sorted and made unique (the reference sorts and std::unique's them);
//  2. one thread per unique pair appends its NEGATIVE / POSITIVE correction records to the same raw streams the
//     classification filled (builder.cpp:340-543), reading the mesh adjacencies as CSR;
//  3. the usual merge follows; edge-edge records are merged on (unordered edge pair, distance type) and keep their
//     orientation (collisions/normal/edge_edge.cpp:123-142): correction collisions carry vertex / edge distance types
//     and are not (min, max)-ordered.
// [emu-begin improved]
struct AdjView {
    const int *vv_off, *vv, *ve_off, *ve, *ev_off, *ev;
    const unsigned char* boundary;
};
__device__ inline bool csr_contains(const int* __restrict__ off, const int* __restrict__ val, int row, int x)
{
    int lo = off[row], hi = off[row + 1];
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int v = val[mid];
        if (v == x) return true;
        if (v < x) lo = mid + 1;
        else hi = mid;
    }
    return false;
}
// typed edge-edge key: [min edge : 28][max edge : 28][distance type : 4][stored as (max, min) : 1]
__device__ inline unsigned long long ee_typed_key(int a, int b, int dt)
{
    const unsigned long long lo = unsigned(min(a, b)), hi = unsigned(max(a, b));
    return (lo << 33) | (hi << 5) | ((unsigned long long)(dt & 15) << 1) | (unsigned long long)(a > b);
}
__device__ inline void emit_key(bool pred, unsigned long long key, unsigned long long* __restrict__ out, unsigned long long* counter)
{
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (pred) out[base + __popc(m & ((1u << lane) - 1))] = key;
}
__device__ inline double active_point_edge(const double4* X, int p, int2 e)
{
    const d3 x[3] = { ld3(X, p), ld3(X, e.x), ld3(X, e.y) };
    return sub_value(sub_point_edge(point_edge_type(x[0], x[1], x[2])), x);
}
// edge-edge candidates -> edge-vertex (candidates.cpp:666-695)
__global__ void __launch_bounds__(256) k_sub_from_ee(int64_t n, const int2* __restrict__ cand, const int2* __restrict__ E, const double4* X,
                                                     double offset_sqr, unsigned long long* out_ev, unsigned long long* cnt_ev)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const int2 c = valid ? cand[i] : make_int2(0, 0);
#pragma unroll
    for (int s = 0; s < 4; s++) {
        bool on = false;
        unsigned long long key = 0;
        if (valid) {
            const int ei = (s >> 1) ? c.y : c.x, ej = (s >> 1) ? c.x : c.y;
            const int2 f = __ldg(E + ej);
            const int vj = (s & 1) ? f.y : f.x;
            on = active_point_edge(X, vj, __ldg(E + ei)) < offset_sqr;
            key = mkkey(ei, vj);
        }
        emit_key(on, key, out_ev, cnt_ev);
    }
}
// face-vertex candidates -> edge-vertex (:640-664) and vertex-vertex (:624-638)
__global__ void __launch_bounds__(256) k_sub_from_fv(int64_t n, const int2* __restrict__ cand, const int2* __restrict__ E, const int4* __restrict__ F,
                                                     const int4* __restrict__ F2E, const double4* X, double offset_sqr, unsigned long long* out_ev,
                                                     unsigned long long* cnt_ev, unsigned long long* out_vv, unsigned long long* cnt_vv)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const int2 c = valid ? cand[i] : make_int2(0, 0);
    int4 f = make_int4(0, 0, 0, 0), fe = f;
    if (valid) f = __ldg(F + c.x), fe = __ldg(F2E + c.x);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        bool on_ev = false, on_vv = false;
        unsigned long long kev = 0, kvv = 0;
        if (valid) {
            const int ei = j == 0 ? fe.x : (j == 1 ? fe.y : fe.z), vj = j == 0 ? f.x : (j == 1 ? f.y : f.z);
            on_ev = active_point_edge(X, c.y, __ldg(E + ei)) < offset_sqr;
            kev = mkkey(ei, c.y);
            on_vv = pp_dist(ld3(X, c.y), ld3(X, vj)) < offset_sqr;
            kvv = mkkey(min(c.y, vj), max(c.y, vj));
        }
        emit_key(on_ev, kev, out_ev, cnt_ev);
        emit_key(on_vv, kvv, out_vv, cnt_vv);
    }
}
// edge-vertex candidates -> vertex-vertex (:584-622)
__global__ void __launch_bounds__(256) k_sub_from_ev(int64_t n, const int2* __restrict__ cand, const int2* __restrict__ E, const double4* X,
                                                     double offset_sqr, unsigned long long* out_vv, unsigned long long* cnt_vv)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const int2 c = valid ? cand[i] : make_int2(0, 0);
    const int2 e = valid ? __ldg(E + c.x) : make_int2(0, 0);
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int vj = j ? e.y : e.x;
        const bool on = valid && pp_dist(ld3(X, c.y), ld3(X, vj)) < offset_sqr;
        emit_key(on, mkkey(min(c.y, vj), max(c.y, vj)), out_vv, cnt_vv);
    }
}

struct CorrOut { // raw streams + their device counters (the ones the classification advanced)
    unsigned long long *key_vv, *key_ev, *key_ee;
    double *w_vv, *w_ev, *w_ee, *eps_ee;
    unsigned char* dt_ee;
    unsigned long long *cnt_vv, *cnt_ev, *cnt_ee;
};
__device__ inline void corr_add_vv(const CorrOut& o, int vi, int vj, double w)
{
    const unsigned long long p = atomicAdd(o.cnt_vv, 1ull);
    o.key_vv[p] = mkkey(min(vi, vj), max(vi, vj));
    o.w_vv[p] = w;
}
__device__ inline void corr_add_ev(const CorrOut& o, int ei, int vi, double w)
{
    const unsigned long long p = atomicAdd(o.cnt_ev, 1ull);
    o.key_ev[p] = mkkey(ei, vi);
    o.w_ev[p] = w;
}
// add_edge_vertex_collision(mesh, candidate, dtype, weight) (builder.cpp:108-138): reduced to the closest feature
__device__ inline void corr_add_ev_typed(const CorrOut& o, int ei, int2 e, int vi, int t, double w)
{
    if (t == PE_E0) corr_add_vv(o, vi, e.x, w);
    else if (t == PE_E1) corr_add_vv(o, vi, e.y, w);
    else corr_add_ev(o, ei, vi, w);
}
// builder.cpp:340-383 (positive = 0: candidates from edge-vertex pairs, negative weights) and :385-419 (positive = 1:
// candidates from face-vertex pairs, positive weights)
__global__ void k_corr_vv(int64_t n, const unsigned long long* __restrict__ uniq, AdjView A, const double* __restrict__ vArea, int area,
                          int positive, CorrOut o)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = int(uniq[i] >> 32), b = int(uniq[i] & 0xffffffffu);
    double w = 0;
#pragma unroll
    for (int dir = 0; dir < 2; dir++) {
        const int vi = dir ? b : a, vj = dir ? a : b;
        const bool incident = csr_contains(A.vv_off, A.vv, vj, vi);
        if (positive) {
            if (A.boundary[vj] || incident) continue; // boundary and incident vertices are skipped
            w += area ? 0.25 * vArea[vi] : 1.0;
        } else {
            const int amt = (A.vv_off[vj + 1] - A.vv_off[vj]) - int(incident);
            if (amt > 1) w += (1 - amt) * (area ? 0.5 * vArea[vi] : 1.0);
        }
    }
    if (w != 0.0) corr_add_vv(o, a, b, w);
}
// builder.cpp:421-455: edge-vertex pairs of face-vertex candidates, negative weights
__global__ void k_corr_ev_from_fv(int64_t n, const unsigned long long* __restrict__ uniq, AdjView A, const int2* __restrict__ E, const double4* X,
                                  const double* __restrict__ vArea, int area, CorrOut o)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int ei = int(uniq[i] >> 32), vi = int(uniq[i] & 0xffffffffu);
    const int amt = (A.ev_off[ei + 1] - A.ev_off[ei]) - int(csr_contains(A.ev_off, A.ev, ei, vi));
    if (amt <= 1) return;
    const double w = (1 - amt) * (area ? 0.25 * vArea[vi] : 1.0);
    const int2 e = __ldg(E + ei);
    corr_add_ev_typed(o, ei, e, vi, point_edge_type(ld3(X, vi), ld3(X, e.x), ld3(X, e.y)), w);
}
// builder.cpp:457-543: edge-vertex pairs (ea, p) of edge-edge candidates: a negative mollified edge-edge collision for
// every mollified edge at p, and the edge-vertex collision itself with weight (#non-mollified edges at p - 1) * w
__global__ void k_corr_ev_from_ee(int64_t n, const unsigned long long* __restrict__ uniq, AdjView A, const int2* __restrict__ E, const double4* X,
                                  const double4* rest, const double* __restrict__ eArea, int area, CorrOut o)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int ea = int(uniq[i] >> 32), p = int(uniq[i] & 0xffffffffu);
    const int2 e = __ldg(E + ea);
    const double w = area ? -0.25 * eArea[ea] : -1.0;
    const d3 a0 = ld3(X, e.x), a1 = ld3(X, e.y);
    const int t = point_edge_type(ld3(X, p), a0, a1);
    int nonmollified = 0;
    for (int q = A.ve_off[p]; q < A.ve_off[p + 1]; q++) {
        const int eb = A.ve[q];
        const int2 f = __ldg(E + eb);
        const int other = p == f.x ? f.y : f.x;
        if (other == e.x || other == e.y) continue;
        const double eps = moll_threshold(ld3(rest, e.x), ld3(rest, e.y), ld3(rest, f.x), ld3(rest, f.y));
        const double cr = sqn(cross(a1 - a0, ld3(X, f.y) - ld3(X, f.x)));
        if (cr >= eps) {
            nonmollified++;
            continue;
        }
        const int first = p == f.x; // is p the first vertex of eb
        const int dt = t == PE_E0 ? (first ? EE_A0B0 : EE_A0B1) : (t == PE_E1 ? (first ? EE_A1B0 : EE_A1B1) : (first ? EE_AB0 : EE_AB1));
        const unsigned long long pos = atomicAdd(o.cnt_ee, 1ull);
        o.key_ee[pos] = ee_typed_key(ea, eb, dt);
        o.w_ee[pos] = w;
        o.eps_ee[pos] = eps;
        o.dt_ee[pos] = (unsigned char)dt;
    }
    if (nonmollified == 1) return; // (rho - 1) = 0
    corr_add_ev_typed(o, ea, e, p, t, (nonmollified - 1) * w);
}
// the classification wrote plain (ea, eb) keys: retype them
__global__ void k_retype_ee_keys(int64_t n, unsigned long long* __restrict__ key, const unsigned char* __restrict__ dt)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = key[i];
    key[i] = ee_typed_key(int(k >> 32), int(k & 0xffffffffu), dt[i]);
}
// runs of equal (edge pair, distance type): bit 0 (the orientation) only orders the records of a run
__global__ void k_runs_ee_typed(int64_t n, const unsigned long long* __restrict__ key, const int* __restrict__ idx,
                                const double* __restrict__ w_raw, int* __restrict__ keep, double* __restrict__ wsum)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = key[i] >> 1;
    if (i > 0 && (key[i - 1] >> 1) == k) {
        keep[i] = 0;
        return;
    }
    const double s = run_weight_sum(i, n, idx, w_raw, [&](int64_t j) { return (key[j] >> 1) == k; });
    wsum[i] = s;
    keep[i] = s != 0.0;
}
__global__ void k_emit_ee_typed(int64_t n, const unsigned long long* __restrict__ key, const int* __restrict__ idx, const int* __restrict__ keep,
                                const int* __restrict__ pos, const double* __restrict__ wsum, const double* __restrict__ eps_raw,
                                int2* __restrict__ ids, double* __restrict__ w, double* __restrict__ eps, unsigned char* __restrict__ dt)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const int p = pos[i];
    const unsigned long long k = key[i];
    const int lo = int(k >> 33), hi = int((k >> 5) & 0xfffffffull);
    ids[p] = (k & 1ull) ? make_int2(hi, lo) : make_int2(lo, hi);
    w[p] = wsum[i];
    eps[p] = eps_raw[idx[i]];
    dt[p] = (unsigned char)((k >> 1) & 15ull);
}
// [emu-end improved]
__global__ void k_ids_to_keys(int64_t n, const int2* __restrict__ ids, int mode, const unsigned char* __restrict__ dt,
                              unsigned long long* __restrict__ key)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int2 c = ids[i];
    key[i] = mode == 2 ? ee_typed_key(c.x, c.y, dt[i]) : (mode == 1 ? mkkey(min(c.x, c.y), max(c.x, c.y)) : mkkey(c.x, c.y));
}
__global__ void k_unpack(UnpackArgs a, int64_t total)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int k = t >= a.first[3] ? 3 : (t >= a.first[2] ? 2 : (t >= a.first[1] ? 1 : 0));
    const int64_t i = t - a.first[k];
    const int2 c = reinterpret_cast<const int2*>(a.in + a.ids[k])[i];
    a.key[k][i] = k == IPCB_EE ? ee_typed_key(c.x, c.y, reinterpret_cast<const unsigned char*>(a.in + a.dt)[i])
                                : (k == IPCB_VV ? mkkey(min(c.x, c.y), max(c.x, c.y)) : mkkey(c.x, c.y));
    a.wout[k][i] = reinterpret_cast<const double*>(a.in + a.w[k])[i];
    if (k == IPCB_EE) {
        a.eps_out[i] = reinterpret_cast<const double*>(a.in + a.eps)[i];
        a.dt_out[i] = reinterpret_cast<const unsigned char*>(a.in + a.dt)[i];
    }
}
static void merge_ee_typed_enqueue(ipcb_ctx* ctx, int64_t n, cudaStream_t s)
{
    CollisionSet& cs = ctx->coll[IPCB_EE];
    cs.count = 0;
    if (n == 0) return;
    cs.idx_raw.reserve(n), cs.idx_sorted.reserve(n), cs.key_sorted.reserve(n), cs.head.reserve(n), cs.pos.reserve(n + 1);
    cs.wsum.reserve(n);
    k_iota<<<grid_for(n, 256), 256, 0, s>>>(n, cs.idx_raw.p);
    size_t bytes = 0, bytes2 = 0;
    const int bits = 61;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, cs.key_raw.p, cs.key_sorted.p, cs.idx_raw.p, cs.idx_sorted.p, n, 0, bits, s);
    cub::DeviceScan::ExclusiveSum(nullptr, bytes2, cs.head.p, cs.pos.p, n, s);
    cs.cubtmp.reserve(std::max(bytes, bytes2));
    cub::DeviceRadixSort::SortPairs(cs.cubtmp.p, bytes, cs.key_raw.p, cs.key_sorted.p, cs.idx_raw.p, cs.idx_sorted.p, n, 0, bits, s);
    k_runs_ee_typed<<<grid_for(n, 256), 256, 0, s>>>(n, cs.key_sorted.p, cs.idx_sorted.p, cs.w_raw.p, cs.head.p, cs.wsum.p);
    cub::DeviceScan::ExclusiveSum(cs.cubtmp.p, bytes2, cs.head.p, cs.pos.p, n, s);
    cs.ids.reserve(n), cs.w.reserve(n), cs.eps.reserve(n), cs.dtype.reserve(n);
    k_emit_ee_typed<<<grid_for(n, 256), 256, 0, s>>>(n, cs.key_sorted.p, cs.idx_sorted.p, cs.head.p, cs.pos.p, cs.wsum.p, cs.eps_raw.p, cs.ids.p,
                                                     cs.w.p, cs.eps.p, cs.dtype.p);
    ctx->launches += 5 + (bits + 7) / 8 + 4;
    IPCB_CUDA(cudaMemcpyAsync(&ctx->pinned.p[20 + 2 * IPCB_EE], cs.pos.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    IPCB_CUDA(cudaMemcpyAsync(&ctx->pinned.p[21 + 2 * IPCB_EE], cs.head.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
}

// init_adjacencies (collision_mesh.cpp:247-307) on the host, once per mesh, mirrored as CSR
static void ensure_adjacency(ipcb_ctx* ctx)
{
    if (ctx->adj_ready) return;
    const int nV = ctx->nV, nE = ctx->nE, nF = ctx->nF;
    std::vector<std::vector<int>> vv(nV), ve(nV), ev(nE);
    for (int i = 0; i < nE; i++) {
        const int a = ctx->hE[2 * size_t(i)], b = ctx->hE[2 * size_t(i) + 1];
        vv[a].push_back(b), vv[b].push_back(a);
        ve[a].push_back(i), ve[b].push_back(i);
    }
    for (int i = 0; i < nF; i++)
        for (int j = 0; j < 3; j++) ev[ctx->hF2E[3 * size_t(i) + j]].push_back(ctx->hF[3 * size_t(i) + (j + 2) % 3]);
    auto flatten = [&](std::vector<std::vector<int>>& rows, Buf<int>& off, Buf<int>& val, int* max_len) {
        std::vector<int> o(rows.size() + 1, 0), v;
        for (size_t r = 0; r < rows.size(); r++) {
            auto& a = rows[r];
            std::sort(a.begin(), a.end());
            a.erase(std::unique(a.begin(), a.end()), a.end());
            o[r + 1] = o[r] + int(a.size());
            v.insert(v.end(), a.begin(), a.end());
            if (max_len) *max_len = std::max(*max_len, int(a.size()));
        }
        off.reserve(o.size()), val.reserve(std::max<size_t>(v.size(), 1));
        IPCB_CUDA(cudaMemcpyAsync(off.p, o.data(), sizeof(int) * o.size(), cudaMemcpyHostToDevice, ctx->stream));
        if (!v.empty()) IPCB_CUDA(cudaMemcpyAsync(val.p, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice, ctx->stream));
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream)); // the host vectors go out of scope
    };
    ctx->adj_max_ve = 0;
    flatten(vv, ctx->adjVVoff, ctx->adjVV, nullptr);
    flatten(ve, ctx->adjVEoff, ctx->adjVE, &ctx->adj_max_ve);
    flatten(ev, ctx->adjEVoff, ctx->adjEV, nullptr);
    std::vector<unsigned char> boundary(std::max(nV, 1), 1); // a vertex of an edge shared by two triangles is interior (:283-292)
    for (int i = 0; i < nE; i++)
        if (ev[i].size() >= 2) boundary[ctx->hE[2 * size_t(i)]] = boundary[ctx->hE[2 * size_t(i) + 1]] = 0;
    ctx->adjBoundary.reserve(boundary.size());
    IPCB_CUDA(cudaMemcpyAsync(ctx->adjBoundary.p, boundary.data(), boundary.size(), cudaMemcpyHostToDevice, ctx->stream));
    IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->adj_ready = true;
}

// sort + unique of `n` keys in ctx->subkey -> out; returns the number of unique keys (host)
static int64_t unique_keys(ipcb_ctx* ctx, int64_t n, Buf<unsigned long long>& out)
{
    if (n == 0) return 0;
    cudaStream_t s = ctx->stream;
    ctx->subkey_sorted.reserve(n), out.reserve(n);
    size_t b1 = 0, b2 = 0;
    unsigned long long* d_num = ctx->dCounters.p + 28;
    cub::DeviceRadixSort::SortKeys(nullptr, b1, ctx->subkey.p, ctx->subkey_sorted.p, n, 0, 64, s);
    cub::DeviceSelect::Unique(nullptr, b2, ctx->subkey_sorted.p, out.p, reinterpret_cast<long long*>(d_num), n, s);
    ctx->cubtmp.reserve(std::max(b1, b2) + 256);
    cub::DeviceRadixSort::SortKeys(ctx->cubtmp.p, b1, ctx->subkey.p, ctx->subkey_sorted.p, n, 0, 64, s);
    cub::DeviceSelect::Unique(ctx->cubtmp.p, b2, ctx->subkey_sorted.p, out.p, reinterpret_cast<long long*>(d_num), n, s);
    ctx->launches += 12;
    long long h = 0;
    IPCB_CUDA(cudaMemcpyAsync(&h, d_num, sizeof h, cudaMemcpyDeviceToHost, s));
    IPCB_CUDA(cudaStreamSynchronize(s));
    return int64_t(h);
}

// step 1: the unique sub-element keys of this context's candidates: [0] VV of EV candidates, [1] EV of EE candidates, [2] EV and
// [3] VV of FV candidates -> ctx->subuniq[k], nu[k]
static void improved_max_approx_keys(ipcb_ctx* ctx, double offset_sqr, int64_t nu[4])
{
    cudaStream_t s = ctx->stream;
    if (ctx->nE >= (1 << 28)) throw Error("IMPROVED_MAX_APPROX: more than 2^28 edges");
    const int64_t nev = ctx->cand[IPCB_EV].count, nee = ctx->cand[IPCB_EE].count, nfv = ctx->cand[IPCB_FV].count;
    unsigned long long* cnt = ctx->dCounters.p + 29; // [0]: keys of the current list, [1]: second list of the face-vertex pass
    auto count_of = [&](int slot) {
        unsigned long long h = 0;
        IPCB_CUDA(cudaMemcpyAsync(&h, cnt + slot, sizeof h, cudaMemcpyDeviceToHost, s));
        IPCB_CUDA(cudaStreamSynchronize(s));
        return int64_t(h);
    };
    nu[0] = nu[1] = nu[2] = nu[3] = 0;
    if (nev) {
        ctx->subkey.reserve(2 * nev);
        IPCB_CUDA(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned long long), s));
        k_sub_from_ev<<<grid_for(nev, 256), 256, 0, s>>>(nev, ctx->cand[IPCB_EV].pairs.p, ctx->dE.p, ctx->X0.p, offset_sqr, ctx->subkey.p, cnt);
        nu[0] = unique_keys(ctx, count_of(0), ctx->subuniq[0]);
    }
    if (nee) {
        ctx->subkey.reserve(4 * nee);
        IPCB_CUDA(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned long long), s));
        k_sub_from_ee<<<grid_for(nee, 256), 256, 0, s>>>(nee, ctx->cand[IPCB_EE].pairs.p, ctx->dE.p, ctx->X0.p, offset_sqr, ctx->subkey.p, cnt);
        nu[1] = unique_keys(ctx, count_of(0), ctx->subuniq[1]);
    }
    if (nfv) {
        ctx->subkey.reserve(6 * nfv); // edge-vertex keys in the first half, vertex-vertex keys in the second
        IPCB_CUDA(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned long long), s));
        unsigned long long* second = ctx->subkey.p + 3 * nfv;
        k_sub_from_fv<<<grid_for(nfv, 256), 256, 0, s>>>(nfv, ctx->cand[IPCB_FV].pairs.p, ctx->dE.p, ctx->dF.p, ctx->dF2E.p, ctx->X0.p, offset_sqr,
                                                        ctx->subkey.p, cnt, second, cnt + 1);
        const int64_t n_ev = count_of(0), n_vv = count_of(1);
        nu[2] = unique_keys(ctx, n_ev, ctx->subuniq[2]);
        if (n_vv) { // move the second list to the front of the key buffer for the sort
            IPCB_CUDA(cudaMemcpyAsync(ctx->subkey.p, second, sizeof(unsigned long long) * n_vv, cudaMemcpyDeviceToDevice, s));
            nu[3] = unique_keys(ctx, n_vv, ctx->subuniq[3]);
        }
    }
    ctx->launches += 3;
    IPCB_CUDA(cudaGetLastError());
}

// steps 2 and 3: the correction records of the keys [first[k], first[k] + nu[k]) of ctx->subuniq[k], appended to the raw streams
static void improved_max_approx_apply(ipcb_ctx* ctx, bool area, int64_t raw[4], const int64_t first[4], const int64_t nu[4])
{
    cudaStream_t s = ctx->stream;
    ensure_adjacency(ctx);
    const AdjView A { ctx->adjVVoff.p, ctx->adjVV.p, ctx->adjVEoff.p, ctx->adjVE.p, ctx->adjEVoff.p, ctx->adjEV.p, ctx->adjBoundary.p };
    // ---- 2. room for the correction records behind the classification's records
    const int64_t more_vv = nu[0] + nu[1] + nu[2] + nu[3], more_ev = nu[1] + nu[2], more_ee = nu[1] * int64_t(std::max(ctx->adj_max_ve, 1));
    CollisionSet &vv = ctx->coll[IPCB_VV], &ev = ctx->coll[IPCB_EV], &ee = ctx->coll[IPCB_EE];
    vv.key_raw.reserve_keep(raw[IPCB_VV] + more_vv, raw[IPCB_VV], s), vv.w_raw.reserve_keep(raw[IPCB_VV] + more_vv, raw[IPCB_VV], s);
    ev.key_raw.reserve_keep(raw[IPCB_EV] + more_ev, raw[IPCB_EV], s), ev.w_raw.reserve_keep(raw[IPCB_EV] + more_ev, raw[IPCB_EV], s);
    ee.key_raw.reserve_keep(raw[IPCB_EE] + more_ee, raw[IPCB_EE], s), ee.w_raw.reserve_keep(raw[IPCB_EE] + more_ee, raw[IPCB_EE], s);
    ee.eps_raw.reserve_keep(raw[IPCB_EE] + more_ee, raw[IPCB_EE], s), ee.dt_raw.reserve_keep(raw[IPCB_EE] + more_ee, raw[IPCB_EE], s);
    if (raw[IPCB_EE]) k_retype_ee_keys<<<grid_for(raw[IPCB_EE], 256), 256, 0, s>>>(raw[IPCB_EE], ee.key_raw.p, ee.dt_raw.p);
    // ---- 3. corrections (the device counters 1 + kind still hold the raw counts of the classification)
    const CorrOut o { vv.key_raw.p, ev.key_raw.p, ee.key_raw.p, vv.w_raw.p, ev.w_raw.p, ee.w_raw.p, ee.eps_raw.p, ee.dt_raw.p,
                      ctx->dCounters.p + 1 + IPCB_VV, ctx->dCounters.p + 1 + IPCB_EV, ctx->dCounters.p + 1 + IPCB_EE };
    const int ar = area ? 1 : 0;
    if (nu[0]) k_corr_vv<<<grid_for(nu[0], 256), 256, 0, s>>>(nu[0], ctx->subuniq[0].p + first[0], A, ctx->dVArea.p, ar, 0, o);
    if (nu[1])
        k_corr_ev_from_ee<<<grid_for(nu[1], 256), 256, 0, s>>>(nu[1], ctx->subuniq[1].p + first[1], A, ctx->dE.p, ctx->X0.p, ctx->dRest.p, ctx->dEArea.p,
                                                              ar, o);
    if (nu[2]) k_corr_ev_from_fv<<<grid_for(nu[2], 256), 256, 0, s>>>(nu[2], ctx->subuniq[2].p + first[2], A, ctx->dE.p, ctx->X0.p, ctx->dVArea.p, ar, o);
    if (nu[3]) k_corr_vv<<<grid_for(nu[3], 256), 256, 0, s>>>(nu[3], ctx->subuniq[3].p + first[3], A, ctx->dVArea.p, ar, 1, o);
    ctx->launches += 5;
    IPCB_CUDA(cudaGetLastError());
    IPCB_CUDA(cudaMemcpyAsync(ctx->pinned.p, ctx->dCounters.p + 1, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    IPCB_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < 4; k++) raw[k] = ctx->pinned.p[k];
}

static void improved_max_approx_corrections(ipcb_ctx* ctx, double offset_sqr, bool area, int64_t raw[4])
{
    Stage st(ctx, "improved_max_approx");
    int64_t nu[4];
    const int64_t first[4] = { 0, 0, 0, 0 };
    improved_max_approx_keys(ctx, offset_sqr, nu);
    improved_max_approx_apply(ctx, area, raw, first, nu);
}

// ---- IMPROVED_MAX_APPROX on a sharded context: the exchange of the sub-element keys -------------------------------------------
void collisions_corrections_keys(ipcb_ctx* ctx, int64_t n[4])
{
    if (!ctx->ima_pending) throw Error("no deferred IMPROVED_MAX_APPROX build on this context");
    for (int k = 0; k < 4; k++) n[k] = ctx->ima_nu[k];
}
// the four lists one after the other
void collisions_corrections_pack(ipcb_ctx* ctx, unsigned long long* d_out)
{
    if (!ctx->ima_pending) throw Error("no deferred IMPROVED_MAX_APPROX build on this context");
    int64_t off = 0;
    for (int k = 0; k < 4; k++) {
        if (ctx->ima_nu[k])
            IPCB_CUDA(cudaMemcpyAsync(d_out + off, ctx->subuniq[k].p, sizeof(unsigned long long) * ctx->ima_nu[k], cudaMemcpyDeviceToDevice, ctx->stream));
        off += ctx->ima_nu[k];
    }
    IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
}
// d_keys: for each of the four lists the keys of ALL ranks (n[k] of them, duplicates allowed), list after list.  Every rank
// unites them (sort + unique: the same list everywhere) and adds the corrections of its slice of every list.
void collisions_corrections_apply(ipcb_ctx* ctx, const unsigned long long* d_keys, const int64_t n[4], int rank, int world)
{
    if (world < 1 || rank < 0 || rank >= world) throw Error("corrections_apply: bad slice");
    if (!ctx->ima_pending) throw Error("no deferred IMPROVED_MAX_APPROX build on this context");
    Stage st(ctx, "improved_max_approx");
    cudaStream_t s = ctx->stream;
    int64_t first[4], nu[4], raw[4], off = 0;
    for (int k = 0; k < 4; k++) {
        first[k] = nu[k] = 0;
        raw[k] = ctx->ima_raw[k];
        if (n[k]) {
            ctx->subkey.reserve(n[k]);
            IPCB_CUDA(cudaMemcpyAsync(ctx->subkey.p, d_keys + off, sizeof(unsigned long long) * n[k], cudaMemcpyDeviceToDevice, s));
            const int64_t u = unique_keys(ctx, n[k], ctx->subuniq[k]);
            first[k] = u * rank / world;
            nu[k] = u * (rank + 1) / world - first[k];
        }
        off += n[k];
    }
    // the append counters of the raw streams continue where the classification stopped (other calls may have used them since)
    unsigned long long h_raw[4] = { (unsigned long long)raw[0], (unsigned long long)raw[1], (unsigned long long)raw[2], (unsigned long long)raw[3] };
    IPCB_CUDA(cudaMemcpyAsync(ctx->dCounters.p + 1, h_raw, sizeof h_raw, cudaMemcpyHostToDevice, s));
    IPCB_CUDA(cudaStreamSynchronize(s));
    improved_max_approx_apply(ctx, ctx->ima_area, raw, first, nu);
    ctx->ima_pending = false;
    ctx->coll_valid = true;
    merge_streams(ctx, raw, false, true);
}

} // namespace ipcb
