// oracle/oracle.cpp — TEST INFRASTRUCTURE ONLY (CPU restatement, never shipped).
//
// CPU restatement of ipc-toolkit v1.6.0's per-step contact pipeline, exported
// through the host half of include/ipcb200.h with the prefix ipco_.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library; the product (libipcb200.so) never does.
//
// Each block cites the reference file:line (under src/ipc/) it follows.
// OpenMP parallelises the same loops the reference parallelises with TBB.
#define IPCB_ORACLE 1
#define IPCB_PREFIX ipco_
#include "../include/ipcb200.h"

#include "geom.hpp"
#include "friction.hpp"
#include "eig.hpp"
#include "ccd.hpp"

#include <omp.h>
#include <parallel/algorithm>
#include <array>
#include <atomic>
#include <cfloat>
#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>

using namespace oracle;

namespace {

thread_local std::string g_error;
int fail(const std::string& msg)
{
    g_error = msg;
    return 1;
}

struct Box {
    double lo[3], hi[3];
};
inline bool overlaps(const Box& a, const Box& b)
{
    // closed-interval test: aabb.cpp:29-33 (double) / lbvh.hpp:66-70 (float)
    return a.lo[0] <= b.hi[0] && b.lo[0] <= a.hi[0] && a.lo[1] <= b.hi[1] && b.lo[1] <= a.hi[1] && a.lo[2] <= b.hi[2]
        && b.lo[2] <= a.hi[2];
}
inline Box box_union(const Box& a, const Box& b)
{
    Box r;
    for (int c = 0; c < 3; c++) {
        r.lo[c] = std::min(a.lo[c], b.lo[c]);
        r.hi[c] = std::max(a.hi[c], b.hi[c]);
    }
    return r;
}

using Pair = std::array<int32_t, 2>;

struct Coll {
    int32_t a, b;
    double w;
    double eps_x;
    uint8_t dtype;
};

} // namespace

struct ipcb_ctx {
    int nV = 0, nE = 0, nF = 0;
    std::vector<V3> rest;
    std::vector<int32_t> E, F, F2E; // row-major nE x 2, nF x 3, nF x 3
    std::vector<int32_t> codim_vertices, codim_edges;
    std::vector<double> vertex_areas, edge_areas;
    // adjacencies (collision_mesh.cpp:247-307), sorted: vertex -> vertices, vertex -> edges, edge -> opposite vertices
    // of its faces; is_vertex_on_boundary
    std::vector<std::vector<int32_t>> vv_adj, ve_adj, ev_adj;
    std::vector<char> on_boundary;
    // broad phase
    int boxes_mode = IPCB_BOXES_FLOAT;
    bool built = false;
    std::vector<Box> vbox, ebox, fbox;
    std::vector<Pair> detected[6];
    int broad_method = 0; // 0 = auto (LBVH above 2048 boxes), 1 = brute force, 2 = LBVH
    // candidates + collisions
    std::vector<Pair> cand[4];
    std::vector<Coll> coll[4];
    std::vector<Coll> appended[4]; // collisions_append since the last collisions_clear
    // IMPROVED_MAX_APPROX over several builders (IPCB_DEFER_CORRECTIONS): the build stops after the classification and this
    // builder's unique sub-element pairs; collisions_corrections_apply finishes it from the pairs of all builders
    bool ima_pending = false, ima_area = false;
    double ima_offset_sqr = 0, ima_dmin = 0;
    std::vector<V3> ima_V;
    std::vector<Coll> ima_raw[4];
    std::vector<Pair> ima_keys[4]; // 0: VV of EV, 1: EV of EE, 2: EV of FV, 3: VV of FV candidates
    double dmin = 0;
    int coll_rank = 0, coll_world = 1; // energy / gradient: slice of every kind's collisions
    int row_lo = 0, row_hi = -1;       // Hessian: owned vertex range (row_hi < 0: all)
    // hessian
    std::vector<int32_t> outer, inner;
    std::vector<double> vals;
    // friction: the lagged tangential collision set (collisions/tangential/tangential_collisions.hpp)
    std::vector<oracle::Tang> tang[4];
    // CollisionMesh::can_collide (collision_mesh.hpp:338) as the intersection of the two descriptor-expressible
    // factories of collision_filter.hpp:113-143
    std::vector<int32_t> patch_ids; // make_vertex_patches_filter (empty: off)
    int32_t n_dynamic = -1;         // make_static_obstacle_filter (< 0: off)
    bool can_vertices_collide(int vi, int vj) const
    {
        return (patch_ids.empty() || patch_ids[vi] != patch_ids[vj]) && (n_dynamic < 0 || vi < n_dynamic || vj < n_dynamic);
    }
};

namespace {

// ---------------------------------------------------------------------------
// boxes: broad_phase/aabb.cpp:35-126, lbvh.cpp:29-41
Box vertex_box(V3 p, double r, int mode)
{
    Box b;
    const double inf = std::numeric_limits<double>::infinity();
    const float finf = std::numeric_limits<float>::infinity();
    for (int c = 0; c < 3; c++) {
        double lo = std::nextafter(p[c] - r, -inf);
        double hi = std::nextafter(p[c] + r, inf);
        if (mode == IPCB_BOXES_FLOAT) {
            lo = std::nextafter(float(lo), -finf);
            hi = std::nextafter(float(hi), finf);
        }
        b.lo[c] = lo;
        b.hi[c] = hi;
    }
    return b;
}

void build_boxes(ipcb_ctx* ctx, const std::vector<V3>& V0, const std::vector<V3>* V1, double r, int mode)
{
    const int nV = ctx->nV;
    ctx->boxes_mode = mode;
    ctx->vbox.resize(nV);
#pragma omp parallel for
    for (int i = 0; i < nV; i++) {
        Box b = vertex_box(V0[i], r, mode);
        if (V1) b = box_union(b, vertex_box((*V1)[i], r, mode)); // aabb.hpp:42-50
        ctx->vbox[i] = b;
    }
    ctx->ebox.resize(ctx->nE);
#pragma omp parallel for
    for (int i = 0; i < ctx->nE; i++) ctx->ebox[i] = box_union(ctx->vbox[ctx->E[2 * i]], ctx->vbox[ctx->E[2 * i + 1]]);
    ctx->fbox.resize(ctx->nF);
#pragma omp parallel for
    for (int i = 0; i < ctx->nF; i++)
        ctx->fbox[i] = box_union(box_union(ctx->vbox[ctx->F[3 * i]], ctx->vbox[ctx->F[3 * i + 1]]), ctx->vbox[ctx->F[3 * i + 2]]);
    ctx->built = true;
}

// ---------------------------------------------------------------------------
// Morton code: math/morton.hpp:23-63
inline uint64_t expand_bits_2(uint64_t v)
{
    v = (v | v << 32) & 0x1F00000000FFFFull;
    v = (v | v << 16) & 0x1F0000FF0000FFull;
    v = (v | v << 8) & 0x100F00F00F00F00Full;
    v = (v | v << 4) & 0x10C30C30C30C30C3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}
inline uint64_t morton_3D(double x, double y, double z)
{
    constexpr double scale = double(1ull << 21);
    x = std::clamp(x * scale, 0.0, scale - 1);
    y = std::clamp(y * scale, 0.0, scale - 1);
    z = std::clamp(z * scale, 0.0, scale - 1);
    return (expand_bits_2(uint64_t(x)) << 2) | (expand_bits_2(uint64_t(y)) << 1) | expand_bits_2(uint64_t(z));
}

// ---------------------------------------------------------------------------
// CPU LBVH (broad_phase/lbvh.cpp:134-330 builds with Apetrei's bottom-up pass;
// here the equivalent radix tree is built with Karras' per-node range search.
// The candidate SET does not depend on the tree: SURVEY §7 hard part 1).
struct LBVH {
    struct Node {
        Box box;
        int left, right; // children (index into nodes); leaf: left = -1, right = primitive id
        int last;        // last (rightmost) sorted leaf below this node
    };
    int n = 0;
    std::vector<Node> nodes; // [0, n-1) internal, [n-1, 2n-1) leaves in Morton order
    std::vector<int> order;  // sorted position -> primitive id

    void build(const std::vector<Box>& boxes)
    {
        n = int(boxes.size());
        nodes.clear();
        order.clear();
        if (n == 0) return;
        Box mesh = boxes[0];
        for (const Box& b : boxes) mesh = box_union(mesh, b); // broad_phase.cpp:93-125
        struct Key {
            uint64_t code;
            int id;
        };
        std::vector<Key> keys(n);
#pragma omp parallel for
        for (int i = 0; i < n; i++) { // lbvh.cpp:150-168
            double m[3];
            for (int c = 0; c < 3; c++) {
                const double w = mesh.hi[c] - mesh.lo[c];
                m[c] = w > 0 ? (0.5 * (boxes[i].lo[c] + boxes[i].hi[c]) - mesh.lo[c]) / w : 0.0;
            }
            keys[i] = { morton_3D(m[0], m[1], m[2]), i };
        }
        __gnu_parallel::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) {
            return a.code != b.code ? a.code < b.code : a.id < b.id;
        });
        order.resize(n);
        nodes.resize(2 * n - 1);
        for (int i = 0; i < n; i++) {
            order[i] = keys[i].id;
            nodes[n - 1 + i] = { boxes[keys[i].id], -1, keys[i].id, i };
        }
        if (n == 1) return;
        auto delta = [&](int i, int j) -> int { // lbvh.cpp:104-131
            if (j < 0 || j >= n) return -1;
            const uint64_t a = keys[i].code, b = keys[j].code;
            if (a == b) return 64 + __builtin_clz(unsigned(i) ^ unsigned(j));
            return __builtin_clzll(a ^ b);
        };
        std::vector<int> parent(2 * n - 1, -1);
#pragma omp parallel for
        for (int i = 0; i < n - 1; i++) {
            const int d = delta(i, i + 1) > delta(i, i - 1) ? 1 : -1;
            const int dmin = delta(i, i - d);
            int lmax = 2;
            while (delta(i, i + lmax * d) > dmin) lmax *= 2;
            int l = 0;
            for (int t = lmax / 2; t >= 1; t /= 2)
                if (delta(i, i + (l + t) * d) > dmin) l += t;
            const int j = i + l * d;
            const int dnode = delta(i, j);
            int s = 0;
            int t = l;
            do {
                t = (t + 1) >> 1;
                if (delta(i, i + (s + t) * d) > dnode) s += t;
            } while (t > 1);
            const int gamma = i + s * d + std::min(d, 0);
            const int lo = std::min(i, j), hi = std::max(i, j);
            const int left = (lo == gamma) ? (n - 1 + gamma) : gamma;
            const int right = (hi == gamma + 1) ? (n - 1 + gamma + 1) : gamma + 1;
            nodes[i].left = left;
            nodes[i].right = right;
            nodes[i].last = hi;
            parent[left] = i;
            parent[right] = i;
        }
        // bottom-up refit
        std::vector<std::atomic<int>> visits(n - 1);
        for (auto& v : visits) v.store(0);
#pragma omp parallel for
        for (int i = 0; i < n; i++) {
            int node = parent[n - 1 + i];
            while (node >= 0) {
                if (visits[node].fetch_add(1) == 0) break;
                nodes[node].box = box_union(nodes[nodes[node].left].box, nodes[nodes[node].right].box);
                node = parent[node];
            }
        }
    }

    // all leaves whose box overlaps q; triangular: only sorted leaves > query_leaf (lbvh.cpp:368-460)
    template <typename Emit> void query(const Box& q, int query_leaf, bool triangular, const Emit& emit) const
    {
        if (n == 0) return;
        if (n == 1) {
            if (!triangular && overlaps(nodes[0].box, q)) emit(nodes[0].right);
            return;
        }
        int stack[128];
        int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const Node& node = nodes[stack[--sp]];
            for (int child : { node.left, node.right }) {
                const Node& c = nodes[child];
                if (!overlaps(c.box, q)) continue;
                if (triangular && c.last <= query_leaf) continue;
                if (c.left < 0) {
                    emit(c.right);
                } else {
                    stack[sp++] = child;
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------
// detection over two box sets with a can_collide predicate
// (brute_force.cpp:15-61 / lbvh.cpp:692-797); pairs are (a in A, b in B);
// same-set detection reports each unordered pair once as (min, max).
template <typename CanCollide>
void detect_pairs(const ipcb_ctx* ctx, const std::vector<Box>& A, const std::vector<Box>& B, bool same, const CanCollide& can_collide,
                  std::vector<Pair>& out)
{
    out.clear();
    const int nA = int(A.size()), nB = int(B.size());
    if (nA == 0 || nB == 0 || (same && nA < 2)) return;
    const bool brute = ctx->broad_method == 1 || (ctx->broad_method == 0 && std::max(nA, nB) <= 2048);
    std::vector<std::vector<Pair>> local(omp_get_max_threads());
    if (brute) {
#pragma omp parallel for schedule(dynamic, 16)
        for (int i = 0; i < nA; i++) {
            auto& mine = local[omp_get_thread_num()];
            for (int j = same ? i + 1 : 0; j < nB; j++)
                if (overlaps(A[i], B[j]) && can_collide(i, j)) mine.push_back({ i, j });
        }
    } else {
        LBVH target;
        target.build(B);
        if (same) {
#pragma omp parallel for schedule(dynamic, 64)
            for (int qi = 0; qi < nA; qi++) {
                auto& mine = local[omp_get_thread_num()];
                const int i = target.order[qi];
                target.query(A[i], qi, true, [&](int j) {
                    if (can_collide(i, j)) mine.push_back({ std::min(i, j), std::max(i, j) });
                });
            }
        } else {
            LBVH source; // query leaves in Morton order like lbvh.cpp:633-690
            source.build(A);
#pragma omp parallel for schedule(dynamic, 64)
            for (int qi = 0; qi < nA; qi++) {
                auto& mine = local[omp_get_thread_num()];
                const int i = source.order[qi];
                target.query(A[i], -1, false, [&](int j) {
                    if (can_collide(i, j)) mine.push_back({ i, j });
                });
            }
        }
    }
    size_t total = 0;
    for (auto& l : local) total += l.size();
    out.reserve(total);
    for (auto& l : local) out.insert(out.end(), l.begin(), l.end()); // utils/merge_thread_local.hpp:21-84
}

void sort_pairs(std::vector<Pair>& p) { __gnu_parallel::sort(p.begin(), p.end()); }

// share-a-vertex rejection + can_vertices_collide on some pair of vertices of the two primitives (lbvh.cpp:801-873)
void broad_detect_kind(ipcb_ctx* ctx, int kind, std::vector<Pair>& out)
{
    const int32_t* E = ctx->E.data();
    const int32_t* F = ctx->F.data();
    // na / nb vertices of the two primitives: no shared vertex, and at least one pair of vertices that can collide
    auto ok = [ctx](const int32_t* a, int na, const int32_t* b, int nb) {
        bool any = false;
        for (int i = 0; i < na; i++)
            for (int j = 0; j < nb; j++) {
                if (a[i] == b[j]) return false;
                any |= ctx->can_vertices_collide(a[i], b[j]);
            }
        return any;
    };
    switch (kind) {
    case IPCB_VV:
        detect_pairs(ctx, ctx->vbox, ctx->vbox, true, [=](int a, int b) { return ctx->can_vertices_collide(a, b); }, out);
        break;
    case IPCB_EV:
        detect_pairs(ctx, ctx->ebox, ctx->vbox, false, [=](int e, int v) { return ok(E + 2 * e, 2, &v, 1); }, out);
        break;
    case IPCB_EE:
        detect_pairs(ctx, ctx->ebox, ctx->ebox, true, [=](int a, int b) { return ok(E + 2 * a, 2, E + 2 * b, 2); }, out);
        break;
    case IPCB_FV:
        detect_pairs(ctx, ctx->fbox, ctx->vbox, false, [=](int f, int v) { return ok(F + 3 * f, 3, &v, 1); }, out);
        break;
    case IPCB_EF:
        detect_pairs(ctx, ctx->ebox, ctx->fbox, false, [=](int e, int f) { return ok(E + 2 * e, 2, F + 3 * f, 3); }, out);
        break;
    case IPCB_FF:
        detect_pairs(ctx, ctx->fbox, ctx->fbox, true, [=](int fa, int fb) { return ok(F + 3 * fa, 3, F + 3 * fb, 3); }, out);
        break;
    }
    sort_pairs(out);
}

std::vector<V3> load_vertices(int n, const double* V, int ld)
{
    std::vector<V3> out(n);
    for (int i = 0; i < n; i++) out[i] = { V[i], V[i + ld], V[i + 2 * size_t(ld)] };
    return out;
}

// Candidates::build (candidates/candidates.cpp:43-222), 3D
void candidates_build(ipcb_ctx* ctx, const std::vector<V3>& V0, const std::vector<V3>* V1, double r)
{
    for (auto& c : ctx->cand) c.clear();
    build_boxes(ctx, V0, V1, r, IPCB_BOXES_FLOAT);
    broad_detect_kind(ctx, IPCB_EE, ctx->cand[IPCB_EE]); // broad_phase.cpp:79-91
    broad_detect_kind(ctx, IPCB_FV, ctx->cand[IPCB_FV]);
    const auto& cv = ctx->codim_vertices;
    const auto& ce = ctx->codim_edges;
    if (!cv.empty()) { // :66-77 codim vertices vs codim vertices
        std::vector<Box> vb(cv.size());
        for (size_t i = 0; i < cv.size(); i++) vb[i] = ctx->vbox[cv[i]];
        std::vector<Pair> vv;
        // the reference leaves broad_phase->can_vertices_collide = mesh.can_collide in place for this pass, whose vertex
        // ids are positions in codim_vertices (candidates.cpp:61,66-77): the filter sees those LOCAL ids — mirrored as is
        detect_pairs(ctx, vb, vb, true, [=](int a, int b) { return ctx->can_vertices_collide(a, b); }, vv);
        for (auto& p : vv) p = { std::min(cv[p[0]], cv[p[1]]), std::max(cv[p[0]], cv[p[1]]) };
        sort_pairs(vv);
        ctx->cand[IPCB_VV] = vv;
    }
    if (!cv.empty() && !ce.empty()) { // :83-116 codim edges vs codim vertices
        std::vector<Box> vb(cv.size()), eb(ce.size());
        for (size_t i = 0; i < cv.size(); i++) vb[i] = ctx->vbox[cv[i]];
        for (size_t i = 0; i < ce.size(); i++) eb[i] = ctx->ebox[ce[i]];
        std::vector<Pair> ev;
        // codim vertices are never endpoints of an edge, so no share-vertex case.  The filter of this pass is
        // make_codim_cross_filter(nCV) & mesh.can_collide on the ids of the re-indexed vertex set [codim vertices;
        // referenced vertices of the codim edges in ascending order] (candidates.cpp:83-108, igl::remove_unreferenced):
        // the cross filter always passes for (codim vertex, edge endpoint) and can_collide sees those LOCAL ids
        std::vector<int32_t> ref;
        for (int e : ce) ref.push_back(ctx->E[2 * e]), ref.push_back(ctx->E[2 * e + 1]);
        std::sort(ref.begin(), ref.end());
        ref.erase(std::unique(ref.begin(), ref.end()), ref.end());
        const int nCV = int(cv.size());
        auto local = [&](int v) { return nCV + int(std::lower_bound(ref.begin(), ref.end(), v) - ref.begin()); };
        std::vector<Pair> le(ce.size());
        for (size_t i = 0; i < ce.size(); i++) le[i] = { local(ctx->E[2 * ce[i]]), local(ctx->E[2 * ce[i] + 1]) };
        detect_pairs(
            ctx, eb, vb, false, [&](int e, int v) { return ctx->can_vertices_collide(v, le[e][0]) || ctx->can_vertices_collide(v, le[e][1]); },
            ev);
        for (auto& p : ev) p = { ce[p[0]], cv[p[1]] };
        sort_pairs(ev);
        ctx->cand[IPCB_EV] = ev;
    }
}

// ---------------------------------------------------------------------------
// stencils (candidates/*.cpp vertex_ids): VV [v0,v1]; EV [v,e0,e1];
// EE [ea0,ea1,eb0,eb1]; FV [v,f0,f1,f2]
int stencil_ids(const ipcb_ctx* ctx, int kind, int a, int b, int32_t ids[4])
{
    const int32_t* E = ctx->E.data();
    const int32_t* F = ctx->F.data();
    switch (kind) {
    case IPCB_VV: ids[0] = a, ids[1] = b; return 2;
    case IPCB_EV: ids[0] = b, ids[1] = E[2 * a], ids[2] = E[2 * a + 1]; return 3;
    case IPCB_EE: ids[0] = E[2 * a], ids[1] = E[2 * a + 1], ids[2] = E[2 * b], ids[3] = E[2 * b + 1]; return 4;
    default: ids[0] = b, ids[1] = F[3 * a], ids[2] = F[3 * a + 1], ids[3] = F[3 * a + 2]; return 4;
    }
}

// NormalCollisionsBuilder::merge (builder.cpp:604-689): records of all builders united, equal collisions
// merged with weight accumulation, weight == 0 dropped; canonical order = sorted by (a, b, dtype)
void merge_collisions(ipcb_ctx* ctx, int k, std::vector<Coll>& all)
{
    // edge-edge collisions are equal when their UNORDERED edge pair and their distance type agree
    // (collisions/normal/edge_edge.cpp:123-142); the stored orientation (the distance type refers to it) is kept.
    // With the IPC set type the first edge id is always the smaller one, so this is the plain (a, b, dtype) order.
    const bool unordered = k == IPCB_EE;
    auto lo = [unordered](const Coll& c) { return unordered ? std::min(c.a, c.b) : c.a; };
    auto hi = [unordered](const Coll& c) { return unordered ? std::max(c.a, c.b) : c.b; };
    __gnu_parallel::stable_sort(all.begin(), all.end(), [&](const Coll& x, const Coll& y) {
        if (lo(x) != lo(y)) return lo(x) < lo(y);
        if (hi(x) != hi(y)) return hi(x) < hi(y);
        if (x.dtype != y.dtype) return x.dtype < y.dtype;
        return x.a < y.a;
    });
    std::vector<Coll>& out = ctx->coll[k];
    out.clear();
    for (const Coll& c : all) {
        if (k != IPCB_FV && !out.empty() && lo(out.back()) == lo(c) && hi(out.back()) == hi(c) && out.back().dtype == c.dtype) {
            out.back().w += c.w;
        } else {
            out.push_back(c);
        }
    }
    if (k != IPCB_FV) {
        out.erase(std::remove_if(out.begin(), out.end(), [](const Coll& c) { return c.w == 0; }), out.end());
    }
}

// step 1 (candidates.cpp:584-695): the unique ACTIVE sub-element pairs of this context's candidates
// keys[0]: vertex-vertex of edge-vertex candidates, [1]: edge-vertex of edge-edge, [2]: edge-vertex and [3]: vertex-vertex of
// face-vertex candidates
template <typename Active> void improved_max_approx_keys(ipcb_ctx* ctx, const std::vector<V3>& V, const Active& is_active, std::vector<Pair> (&keys)[4])
{
    const int32_t* E = ctx->E.data();
    const int32_t* F = ctx->F.data();
    auto unique_unordered = [](std::vector<Pair>& p) { // VertexVertexCandidate: order and equality ignore orientation
        for (auto& q : p)
            if (q[0] > q[1]) std::swap(q[0], q[1]);
        std::sort(p.begin(), p.end());
        p.erase(std::unique(p.begin(), p.end()), p.end());
    };
    auto unique_ordered = [](std::vector<Pair>& p) {
        std::sort(p.begin(), p.end());
        p.erase(std::unique(p.begin(), p.end()), p.end());
    };
    for (auto& k : keys) k.clear();
    for (const Pair& c : ctx->cand[IPCB_EV]) // :584-622
        for (int j = 0; j < 2; j++) {
            const int vi = c[1], vj = E[2 * c[0] + j];
            if (is_active(point_point_distance(V[vi], V[vj]))) keys[0].push_back({ vi, vj });
        }
    unique_unordered(keys[0]);
    for (const Pair& c : ctx->cand[IPCB_EE]) // :666-695
        for (int i = 0; i < 2; i++) {
            const int ei = c[i], ej = c[1 - i];
            for (int j = 0; j < 2; j++) {
                const int vj = E[2 * ej + j];
                const V3 p = V[vj], e0 = V[E[2 * ei]], e1 = V[E[2 * ei + 1]];
                if (is_active(point_edge_distance(p, e0, e1, point_edge_distance_type(p, e0, e1)))) keys[1].push_back({ ei, vj });
            }
        }
    unique_ordered(keys[1]);
    for (const Pair& c : ctx->cand[IPCB_FV]) { // :624-664
        const int fi = c[0], vi = c[1];
        for (int j = 0; j < 3; j++) {
            const int ei = ctx->F2E[3 * fi + j];
            const V3 p = V[vi], e0 = V[E[2 * ei]], e1 = V[E[2 * ei + 1]];
            if (is_active(point_edge_distance(p, e0, e1, point_edge_distance_type(p, e0, e1)))) keys[2].push_back({ ei, vi });
            const int vj = F[3 * fi + j];
            if (is_active(point_point_distance(V[vi], V[vj]))) keys[3].push_back({ vi, vj });
        }
    }
    unique_ordered(keys[2]);
    unique_unordered(keys[3]);
}

// step 2 (builder.cpp:340-543): the NEGATIVE / POSITIVE correction collisions of the pairs [first[k], first[k] + count[k]) of
// keys[k], whose weight counts how often the IPC passes over-counted the pair
inline void improved_max_approx_apply(ipcb_ctx* ctx, const std::vector<V3>& V, bool area, const std::vector<Pair> (&keys)[4], const size_t first[4],
                                      const size_t count[4], std::vector<std::vector<Coll>> (&loc)[4])
{
    const int32_t* E = ctx->E.data();
    auto contains = [](const std::vector<int32_t>& sorted, int32_t v) { return std::binary_search(sorted.begin(), sorted.end(), v); };
    auto add_vv = [&](int vi, int vj, double w) { loc[IPCB_VV][0].push_back({ std::min(vi, vj), std::max(vi, vj), w, 0, 0 }); };
    auto add_ev = [&](int ei, int vi, double w) { loc[IPCB_EV][0].push_back({ ei, vi, w, 0, 0 }); };
    // add_edge_vertex_collision(mesh, candidate, dtype, weight) :108-138: reduced to the closest feature
    auto add_ev_typed = [&](int ei, int vi, PE dtype, double w) {
        if (dtype == PE_P_E0) add_vv(vi, E[2 * ei], w);
        else if (dtype == PE_P_E1) add_vv(vi, E[2 * ei + 1], w);
        else add_ev(ei, vi, w);
    };
    // ---- vertex-vertex pairs of edge-vertex candidates: negative corrections (builder.cpp:340-383)
    for (size_t i = first[0]; i < first[0] + count[0]; i++) {
        const Pair& c = keys[0][i];
        double w = 0;
        auto add_weight = [&](int vi, int vj) {
            const auto& inc = ctx->vv_adj[vj];
            const int amt = int(inc.size()) - int(contains(inc, vi));
            if (amt > 1) w += (1 - amt) * (area ? 0.5 * ctx->vertex_areas[vi] : 1.0); // / 2: double counting
        };
        add_weight(c[0], c[1]);
        add_weight(c[1], c[0]);
        if (w != 0) add_vv(c[0], c[1], w);
    }
    // ---- edge-vertex pairs of edge-edge candidates: negative corrections (builder.cpp:457-543)
    for (size_t i = first[1]; i < first[1] + count[1]; i++) {
        const Pair& c = keys[1][i];
        const int ea = c[0], p = c[1];
        const int ea0 = E[2 * ea], ea1 = E[2 * ea + 1];
        const double w = area ? -0.25 * ctx->edge_areas[ea] : -1.0; // / 4: double counting and PT + EE
        const PE dtype = point_edge_distance_type(V[p], V[ea0], V[ea1]);
        int nonmollified = 0;
        for (const int32_t eb : ctx->ve_adj[p]) {
            const int eb0 = E[2 * eb], eb1 = E[2 * eb + 1];
            const int q = p == eb0 ? eb1 : eb0;
            if (q == ea0 || q == ea1) continue;
            const double eps_x = edge_edge_mollifier_threshold(ctx->rest[ea0], ctx->rest[ea1], ctx->rest[eb0], ctx->rest[eb1]);
            if (edge_edge_cross_squarednorm(V[ea0], V[ea1], V[eb0], V[eb1]) >= eps_x) {
                nonmollified++;
                continue;
            }
            // a mollified edge-edge collision with the point-edge type lifted to the edge pair (ea, eb)
            EE ee;
            if (dtype == PE_P_E0) ee = p == eb0 ? EE_EA0_EB0 : EE_EA0_EB1;
            else if (dtype == PE_P_E1) ee = p == eb0 ? EE_EA1_EB0 : EE_EA1_EB1;
            else ee = p == eb0 ? EE_EA_EB0 : EE_EA_EB1;
            loc[IPCB_EE][0].push_back({ ea, eb, w, eps_x, uint8_t(ee) });
        }
        if (nonmollified == 1) continue; // (rho - 1) = 0
        add_ev_typed(ea, p, dtype, (nonmollified - 1) * w);
    }
    // ---- edge-vertex pairs of face-vertex candidates: negative (builder.cpp:421-455)
    for (size_t i = first[2]; i < first[2] + count[2]; i++) {
        const Pair& c = keys[2][i];
        const int ei = c[0], vi = c[1];
        const auto& inc = ctx->ev_adj[ei];
        const int amt = int(inc.size()) - int(contains(inc, vi));
        if (amt > 1) {
            const double w = (1 - amt) * (area ? 0.25 * ctx->vertex_areas[vi] : 1.0);
            add_ev_typed(ei, vi, point_edge_distance_type(V[vi], V[E[2 * ei]], V[E[2 * ei + 1]]), w);
        }
    }
    // ---- vertex-vertex pairs of face-vertex candidates: positive (builder.cpp:385-419)
    for (size_t i = first[3]; i < first[3] + count[3]; i++) {
        const Pair& c = keys[3][i];
        double w = 0;
        auto add_weight = [&](int vi, int vj) {
            if (ctx->on_boundary[vj] || contains(ctx->vv_adj[vj], vi)) return; // boundary and incident vertices are skipped
            w += area ? 0.25 * ctx->vertex_areas[vi] : 1.0;
        };
        add_weight(c[0], c[1]);
        add_weight(c[1], c[0]);
        if (w != 0) add_vv(c[0], c[1], w);
    }
}

// CollisionSetType::IMPROVED_MAX_APPROX (normal_collisions.cpp:84-128): after the IPC passes, sub-element candidates
// are derived from the element candidates (candidates.cpp:584-695: active vertex-vertex / edge-vertex pairs of every
// edge-vertex / edge-edge / face-vertex candidate, duplicates removed) and each adds a NEGATIVE or POSITIVE correction
// collision whose weight counts how often the IPC passes over-counted the pair (builder.cpp:340-543).
template <typename Active>
void improved_max_approx_corrections(ipcb_ctx* ctx, const std::vector<V3>& V, bool area, const Active& is_active, std::vector<std::vector<Coll>> (&loc)[4])
{
    std::vector<Pair> keys[4];
    improved_max_approx_keys(ctx, V, is_active, keys);
    const size_t first[4] = { 0, 0, 0, 0 }, count[4] = { keys[0].size(), keys[1].size(), keys[2].size(), keys[3].size() };
    improved_max_approx_apply(ctx, V, area, keys, first, count, loc);
}

// NormalCollisions::build(candidates, ...) — normal_collisions.cpp:38-158 with
// the IPC set type, builder.cpp:26-336 (classification + reduction) and
// :547-689 (merge with weight accumulation, weight == 0 dropped)
void collisions_build(ipcb_ctx* ctx, const std::vector<V3>& V, double dhat, double dmin, int flags)
{
    const bool area = flags & IPCB_USE_AREA_WEIGHTING;
    const double offset_sqr = (dmin + dhat) * (dmin + dhat);
    auto is_active = [&](double d_sqr) { return d_sqr < offset_sqr; };
    const int32_t* E = ctx->E.data();
    const int32_t* F = ctx->F.data();
    const int nt = omp_get_max_threads();
    std::vector<std::vector<Coll>> loc[4];
    for (auto& l : loc) l.resize(nt);
    auto add_vv = [&](int t, int vi, int vj, double w) { loc[IPCB_VV][t].push_back({ std::min(vi, vj), std::max(vi, vj), w, 0, 0 }); };
    auto add_ev = [&](int t, int ei, int vi, double w) { loc[IPCB_EV][t].push_back({ ei, vi, w, 0, 0 }); };

    const auto& vvc = ctx->cand[IPCB_VV];
#pragma omp parallel for
    for (size_t i = 0; i < vvc.size(); i++) { // builder.cpp:26-56
        const int t = omp_get_thread_num();
        const int vi = vvc[i][0], vj = vvc[i][1];
        if (!is_active(point_point_distance(V[vi], V[vj]))) continue;
        add_vv(t, vi, vj, area ? 0.5 * (ctx->vertex_areas[vi] + ctx->vertex_areas[vj]) : 1);
    }
    const auto& evc = ctx->cand[IPCB_EV];
#pragma omp parallel for
    for (size_t i = 0; i < evc.size(); i++) { // builder.cpp:58-138
        const int t = omp_get_thread_num();
        const int ei = evc[i][0], vi = evc[i][1];
        const V3 v = V[vi], e0 = V[E[2 * ei]], e1 = V[E[2 * ei + 1]];
        const PE dtype = point_edge_distance_type(v, e0, e1);
        if (!is_active(point_edge_distance(v, e0, e1, dtype))) continue;
        const double w = area ? 0.5 * ctx->vertex_areas[vi] : 1;
        switch (dtype) {
        case PE_P_E0: add_vv(t, vi, E[2 * ei], w); break;
        case PE_P_E1: add_vv(t, vi, E[2 * ei + 1], w); break;
        default: add_ev(t, ei, vi, w); break;
        }
    }
    const auto& eec = ctx->cand[IPCB_EE];
#pragma omp parallel for
    for (size_t i = 0; i < eec.size(); i++) { // builder.cpp:140-237
        const int t = omp_get_thread_num();
        const int eai = eec[i][0], ebi = eec[i][1];
        const int ea0i = E[2 * eai], ea1i = E[2 * eai + 1], eb0i = E[2 * ebi], eb1i = E[2 * ebi + 1];
        const V3 ea0 = V[ea0i], ea1 = V[ea1i], eb0 = V[eb0i], eb1 = V[eb1i];
        const EE actual_dtype = edge_edge_distance_type(ea0, ea1, eb0, eb1);
        if (!is_active(edge_edge_distance(ea0, ea1, eb0, eb1, actual_dtype))) continue;
        const double eps_x = edge_edge_mollifier_threshold(ctx->rest[ea0i], ctx->rest[ea1i], ctx->rest[eb0i], ctx->rest[eb1i]);
        const double ee_cross_norm_sqr = edge_edge_cross_squarednorm(ea0, ea1, eb0, eb1);
        const EE dtype = ee_cross_norm_sqr < eps_x ? EE_EA_EB : actual_dtype;
        const double w = area ? 0.25 * (ctx->edge_areas[eai] + ctx->edge_areas[ebi]) : 1;
        switch (dtype) {
        case EE_EA0_EB0: add_vv(t, ea0i, eb0i, w); break;
        case EE_EA0_EB1: add_vv(t, ea0i, eb1i, w); break;
        case EE_EA1_EB0: add_vv(t, ea1i, eb0i, w); break;
        case EE_EA1_EB1: add_vv(t, ea1i, eb1i, w); break;
        case EE_EA_EB0: add_ev(t, eai, eb0i, w); break;
        case EE_EA_EB1: add_ev(t, eai, eb1i, w); break;
        case EE_EA0_EB: add_ev(t, ebi, ea0i, w); break;
        case EE_EA1_EB: add_ev(t, ebi, ea1i, w); break;
        default: loc[IPCB_EE][t].push_back({ eai, ebi, w, eps_x, uint8_t(actual_dtype) }); break;
        }
    }
    const auto& fvc = ctx->cand[IPCB_FV];
#pragma omp parallel for
    for (size_t i = 0; i < fvc.size(); i++) { // builder.cpp:239-336
        const int t = omp_get_thread_num();
        const int fi = fvc[i][0], vi = fvc[i][1];
        const int f0i = F[3 * fi], f1i = F[3 * fi + 1], f2i = F[3 * fi + 2];
        const V3 v = V[vi], f0 = V[f0i], f1 = V[f1i], f2 = V[f2i];
        const PT dtype = point_triangle_distance_type(v, f0, f1, f2);
        if (!is_active(point_triangle_distance(v, f0, f1, f2, dtype))) continue;
        const double w = area ? 0.25 * ctx->vertex_areas[vi] : 1;
        switch (dtype) {
        case PT_P_T0: add_vv(t, vi, f0i, w); break;
        case PT_P_T1: add_vv(t, vi, f1i, w); break;
        case PT_P_T2: add_vv(t, vi, f2i, w); break;
        case PT_P_E0: add_ev(t, ctx->F2E[3 * fi + 0], vi, w); break;
        case PT_P_E1: add_ev(t, ctx->F2E[3 * fi + 1], vi, w); break;
        case PT_P_E2: add_ev(t, ctx->F2E[3 * fi + 2], vi, w); break;
        default: loc[IPCB_FV][t].push_back({ fi, vi, w, 0, 0 }); break;
        }
    }
    ctx->ima_pending = false;
    if ((flags & IPCB_SET_IMPROVED_MAX_APPROX) && (flags & IPCB_DEFER_CORRECTIONS)) {
        // one of several builders: stop after the classification and this builder's sub-element pairs
        improved_max_approx_keys(ctx, V, is_active, ctx->ima_keys);
        for (int k = 0; k < 4; k++) {
            ctx->ima_raw[k].clear();
            for (auto& l : loc[k]) ctx->ima_raw[k].insert(ctx->ima_raw[k].end(), l.begin(), l.end());
            ctx->coll[k].clear();
        }
        ctx->ima_V = V, ctx->ima_area = area, ctx->ima_offset_sqr = offset_sqr, ctx->ima_dmin = dmin;
        ctx->ima_pending = true;
        return;
    }
    if (flags & IPCB_SET_IMPROVED_MAX_APPROX) improved_max_approx_corrections(ctx, V, area, is_active, loc);
    for (int k = 0; k < 4; k++) {
        std::vector<Coll> all;
        for (auto& l : loc[k]) all.insert(all.end(), l.begin(), l.end());
        merge_collisions(ctx, k, all);
    }
    ctx->dmin = dmin; // normal_collisions.cpp:154-157
}

// the deferred build's second half: `keys` = the pairs of ALL builders (duplicates allowed); this builder adds the corrections
// of its slice [rank, world) of every united list
void collisions_corrections_apply(ipcb_ctx* ctx, std::vector<Pair> (&keys)[4], int rank, int world)
{
    size_t first[4], count[4];
    for (int k = 0; k < 4; k++) {
        std::sort(keys[k].begin(), keys[k].end());
        keys[k].erase(std::unique(keys[k].begin(), keys[k].end()), keys[k].end());
        const size_t u = keys[k].size();
        first[k] = u * size_t(rank) / size_t(world);
        count[k] = u * size_t(rank + 1) / size_t(world) - first[k];
    }
    std::vector<std::vector<Coll>> loc[4];
    for (int k = 0; k < 4; k++) loc[k].assign(1, ctx->ima_raw[k]);
    improved_max_approx_apply(ctx, ctx->ima_V, ctx->ima_area, keys, first, count, loc);
    for (int k = 0; k < 4; k++) merge_collisions(ctx, k, loc[k][0]);
    ctx->dmin = ctx->ima_dmin;
    ctx->ima_pending = false;
    for (auto& r : ctx->ima_raw) r.clear();
    ctx->ima_V.clear();
}

// ---------------------------------------------------------------------------
// per-collision calculus: potentials/normal_potential.cpp:127-232,
// barrier_potential.cpp:62-99
struct Barrier {
    double dhat, kappa;
    bool physical;
    double scale(double dmin) const
    {
        // physical barrier: dhat / units(x_hat) with ClampedLogBarrier::units(x) = x * x
        // (barrier/barrier.hpp:133-137, potentials/barrier_potential.cpp:68-70)
        const double xhat = (2 * dmin + dhat) * dhat;
        return physical ? dhat / (xhat * xhat) : 1.0;
    }
    double f(double d_sqr, double dmin) const
    {
        return kappa * (barrier(d_sqr - dmin * dmin, (2 * dmin + dhat) * dhat) * (physical ? scale(dmin) : 1.0));
    }
    double df(double d_sqr, double dmin) const
    {
        return kappa * (barrier_first_derivative(d_sqr - dmin * dmin, (2 * dmin + dhat) * dhat) * (physical ? scale(dmin) : 1.0));
    }
    double ddf(double d_sqr, double dmin) const
    {
        return kappa * (barrier_second_derivative(d_sqr - dmin * dmin, (2 * dmin + dhat) * dhat) * (physical ? scale(dmin) : 1.0));
    }
};

Embed collision_embed(int kind, const Coll& c)
{
    switch (kind) {
    case IPCB_VV: return { PRIM_PP, 2, { 0, 1, 0, 0 } };
    case IPCB_EV: return embed_point_edge(PE_P_E);            // collisions/normal/edge_vertex.hpp:28-32
    case IPCB_EE: return embed_edge_edge(EE(c.dtype));        // collisions/normal/edge_edge.hpp:96
    default: return embed_point_triangle(PT_P_T);             // collisions/normal/face_vertex.hpp:28-32
    }
}

double collision_energy(int kind, const Coll& c, const V3* x, const Barrier& B, double dmin)
{
    const double d = prim_value(collision_embed(kind, c), x);
    double m = 1;
    if (kind == IPCB_EE) m = edge_edge_mollifier(edge_edge_cross_squarednorm(x[0], x[1], x[2], x[3]), c.eps_x);
    return c.w * m * B.f(d, dmin);
}

// local gradient (n = 3 * npts)
void collision_gradient(int kind, const Coll& c, const V3* x, const Barrier& B, double dmin, double* g)
{
    const int n = 3 * (kind == IPCB_VV ? 2 : kind == IPCB_EV ? 3 : 4);
    double m = 1;
    double s = 0;
    if (kind == IPCB_EE) {
        s = edge_edge_cross_squarednorm(x[0], x[1], x[2], x[3]);
        m = edge_edge_mollifier(s, c.eps_x);
    }
    if (m <= 0) { // normal_potential.cpp:143-147
        for (int i = 0; i < n; i++) g[i] = 0;
        return;
    }
    Deriv D;
    embed_deriv(collision_embed(kind, c), x, D);
    const double f = B.f(D.val, dmin), grad_f = B.df(D.val, dmin);
    if (kind != IPCB_EE) {
        for (int i = 0; i < n; i++) g[i] = (c.w * grad_f) * D.g[i];
        return;
    }
    Deriv S;
    if (s < c.eps_x) {
        edge_edge_cross_squarednorm_deriv(x[0], x[1], x[2], x[3], S);
    } else {
        S.zero();
    }
    const double dm = edge_edge_mollifier_gradient(s, c.eps_x);
    for (int i = 0; i < n; i++) {
        const double grad_m = s < c.eps_x ? dm * S.g[i] : 0.0;
        g[i] = (c.w * f) * grad_m + (c.w * m * grad_f) * D.g[i];
    }
}

// local hessian (n x n col-major, ld 12), PSD-projected
void collision_hessian(int kind, const Coll& c, const V3* x, const Barrier& B, double dmin, int psd_mode, double* H)
{
    const int n = 3 * (kind == IPCB_VV ? 2 : kind == IPCB_EV ? 3 : 4);
    Deriv D;
    embed_deriv(collision_embed(kind, c), x, D);
    const double d = D.val;
    auto h = [&](int r, int cc) -> double& { return H[r + 12 * cc]; };
    if (kind != IPCB_EE) {
        const double grad_f = B.df(d, dmin), hess_f = B.ddf(d, dmin);
        for (int j = 0; j < n; j++)
            for (int i = 0; i < n; i++) h(i, j) = (c.w * hess_f) * D.g[i] * D.g[j] + (c.w * grad_f) * D.h(i, j);
    } else {
        const double s = edge_edge_cross_squarednorm(x[0], x[1], x[2], x[3]);
        const double m = edge_edge_mollifier(s, c.eps_x);
        Deriv S;
        double grad_m[12], hess_m[144];
        if (s < c.eps_x) {
            edge_edge_cross_squarednorm_deriv(x[0], x[1], x[2], x[3], S);
            const double dm = edge_edge_mollifier_gradient(s, c.eps_x), ddm = edge_edge_mollifier_hessian(s, c.eps_x);
            for (int i = 0; i < 12; i++) grad_m[i] = dm * S.g[i];
            for (int j = 0; j < 12; j++)
                for (int i = 0; i < 12; i++) hess_m[i + 12 * j] = (dm * S.h(i, j)) + ((ddm * S.g[i]) * S.g[j]);
        } else {
            for (double& v : grad_m) v = 0;
            for (double& v : hess_m) v = 0;
        }
        const double f = B.f(d, dmin);
        if (m <= 0) { // normal_potential.cpp:184-189
            for (int j = 0; j < 12; j++)
                for (int i = 0; i < 12; i++) h(i, j) = (c.w * f) * hess_m[i + 12 * j];
            return; // NOTE: the reference returns before project_to_psd here
        }
        const double grad_f = B.df(d, dmin), hess_f = B.ddf(d, dmin);
        const double weighted_m = c.w * m;
        for (int j = 0; j < 12; j++)
            for (int i = 0; i < 12; i++) {
                const double gfgm_ij = (c.w * grad_f) * D.g[i] * grad_m[j];
                const double gfgm_ji = (c.w * grad_f) * D.g[j] * grad_m[i];
                h(i, j) = (c.w * f) * hess_m[i + 12 * j] + gfgm_ij + gfgm_ji + (weighted_m * hess_f) * D.g[i] * D.g[j]
                    + (weighted_m * grad_f) * D.h(i, j);
            }
    }
    project_to_psd(n, H, 12, psd_mode);
}

template <typename Fn> void for_each_collision(const ipcb_ctx* ctx, const Fn& fn)
{
    for (int k = 0; k < 4; k++) {
        const auto& cs = ctx->coll[k];
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < cs.size(); i++) fn(k, cs[i]);
    }
}

ipcb_ccd_params resolve_ccd(const ipcb_ccd_params* p)
{
    ipcb_ccd_params r = p ? *p : ipcb_ccd_params { IPCB_CCD_TIGHT_INCLUSION, 0, 0, 0 };
    if (r.kind == IPCB_CCD_ADDITIVE) {
        if (r.max_iterations == 0) r.max_iterations = AdditiveCCD::DEFAULT_MAX_ITERATIONS;
        if (r.conservative_rescaling <= 0) r.conservative_rescaling = AdditiveCCD::DEFAULT_CONSERVATIVE_RESCALING;
    } else {
        if (r.tolerance <= 0) r.tolerance = TightInclusionCCD::DEFAULT_TOLERANCE;
        if (r.max_iterations == 0) r.max_iterations = TightInclusionCCD::DEFAULT_MAX_ITERATIONS;
        if (r.conservative_rescaling <= 0) r.conservative_rescaling = TightInclusionCCD::DEFAULT_CONSERVATIVE_RESCALING;
    }
    return r;
}

bool narrow_ccd(int kind, const V3* t0, const V3* t1, double min_distance, double tmax, const ipcb_ccd_params& p, double& toi)
{
    if (p.kind == IPCB_CCD_ADDITIVE) {
        AdditiveCCD a;
        a.max_iterations = long(p.max_iterations);
        a.conservative_rescaling = p.conservative_rescaling;
        switch (kind) {
        case IPCB_VV: return a.point_point_ccd(t0, t1, toi, min_distance, tmax);
        case IPCB_EV: return a.point_edge_ccd(t0, t1, toi, min_distance, tmax);
        case IPCB_EE: return a.edge_edge_ccd(t0, t1, toi, min_distance, tmax);
        default: return a.point_triangle_ccd(t0, t1, toi, min_distance, tmax);
        }
    }
    TightInclusionCCD ti;
    ti.tolerance = p.tolerance;
    ti.max_iterations = long(p.max_iterations);
    ti.conservative_rescaling = p.conservative_rescaling;
    switch (kind) {
    case IPCB_VV: return ti.point_point_ccd(t0, t1, toi, min_distance, tmax);
    case IPCB_EV: return ti.point_edge_ccd(t0, t1, toi, min_distance, tmax);
    case IPCB_EE: return ti.edge_edge_ccd(t0, t1, toi, min_distance, tmax);
    default: return ti.point_triangle_ccd(t0, t1, toi, min_distance, tmax);
    }
}

// Candidates::compute_collision_free_stepsize (candidates.cpp:252-292)
double stepsize_from_candidates(const ipcb_ctx* ctx, const std::vector<V3>& V0, const std::vector<V3>& V1, double min_distance,
                                const ipcb_ccd_params& p)
{
    size_t total = 0;
    for (auto& c : ctx->cand) total += c.size();
    if (total == 0) return 1.0;
    std::atomic<double> earliest_toi(1.0);
    for (int k = 0; k < 4; k++) {
        const auto& cs = ctx->cand[k];
#pragma omp parallel for schedule(dynamic, 256)
        for (size_t i = 0; i < cs.size(); i++) {
            const double tmax = earliest_toi.load(std::memory_order_relaxed);
            int32_t ids[4];
            const int n = stencil_ids(ctx, k, cs[i][0], cs[i][1], ids);
            V3 t0[4], t1[4];
            for (int j = 0; j < n; j++) t0[j] = V0[ids[j]], t1[j] = V1[ids[j]];
            double toi = std::numeric_limits<double>::infinity();
            if (narrow_ccd(k, t0, t1, min_distance, tmax, p, toi)) {
                double prev = earliest_toi.load(std::memory_order_relaxed);
                while (toi < prev && !earliest_toi.compare_exchange_weak(prev, toi, std::memory_order_relaxed)) { }
            }
        }
    }
    return earliest_toi.load();
}

} // namespace

// ===========================================================================
// C ABI
extern "C" {

int ipco_ctx_create(int, ipcb_ctx** out)
{
    *out = new ipcb_ctx();
    return 0;
}
void ipco_ctx_destroy(ipcb_ctx* ctx) { delete ctx; }
const char* ipco_last_error(void) { return g_error.c_str(); }
const char* ipco_backend_name(void) { return "oracle-cpu"; }
void* ipco_ctx_stream(ipcb_ctx*) { return nullptr; }

// oracle-only knobs (tests): broad phase method 0 auto / 1 brute force / 2 LBVH
int ipco_set_broad_method(ipcb_ctx* ctx, int method)
{
    ctx->broad_method = method;
    return 0;
}
int ipco_num_threads(void) { return omp_get_max_threads(); }
void ipco_set_num_threads(int n) { omp_set_num_threads(n); }

// CollisionMesh: collision_mesh.cpp:15-127, :145-183, :309-374, :510-543
int ipco_mesh_set(ipcb_ctx* ctx, int32_t nV, const double* rest, int32_t ld_rest, int32_t nE, const int32_t* E, int32_t ldE,
                  int32_t nF, const int32_t* F, int32_t ldF)
{
    ctx->nV = nV, ctx->nE = nE, ctx->nF = nF;
    ctx->patch_ids.clear(), ctx->n_dynamic = -1; // a new mesh accepts all pairs
    ctx->rest = load_vertices(nV, rest, ld_rest);
    ctx->E.resize(2 * size_t(nE));
    ctx->F.resize(3 * size_t(nF));
    for (int i = 0; i < nE; i++)
        for (int k = 0; k < 2; k++) {
            const int v = E[i + size_t(ldE) * k];
            if (v < 0 || v >= nV) return fail("edge vertex id out of range");
            ctx->E[2 * i + k] = v;
        }
    for (int i = 0; i < nF; i++)
        for (int k = 0; k < 3; k++) {
            const int v = F[i + size_t(ldF) * k];
            if (v < 0 || v >= nV) return fail("face vertex id out of range");
            ctx->F[3 * i + k] = v;
        }
    // faces_to_edges (:510-543)
    std::unordered_map<uint64_t, int> edge_map;
    edge_map.reserve(nE * 2);
    auto key = [](int a, int b) { return (uint64_t(uint32_t(std::min(a, b))) << 32) | uint32_t(std::max(a, b)); };
    for (int i = 0; i < nE; i++) edge_map.emplace(key(ctx->E[2 * i], ctx->E[2 * i + 1]), i);
    ctx->F2E.resize(3 * size_t(nF));
    for (int i = 0; i < nF; i++)
        for (int k = 0; k < 3; k++) {
            auto it = edge_map.find(key(ctx->F[3 * i + k], ctx->F[3 * i + (k + 1) % 3]));
            if (it == edge_map.end()) return fail("Unable to find edge!");
            ctx->F2E[3 * i + k] = it->second;
        }
    // codim vertices / edges (:145-183)
    std::vector<char> is_codim_v(nV, 1), is_codim_e(nE, 1);
    for (int v : ctx->E) is_codim_v[v] = 0;
    for (int e : ctx->F2E) is_codim_e[e] = 0;
    ctx->codim_vertices.clear();
    ctx->codim_edges.clear();
    for (int i = 0; i < nV; i++)
        if (is_codim_v[i]) ctx->codim_vertices.push_back(i);
    for (int i = 0; i < nE; i++)
        if (is_codim_e[i]) ctx->codim_edges.push_back(i);
    // areas (:309-374)
    auto edge_len = [&](int i) { return std::sqrt(sqnorm(ctx->rest[ctx->E[2 * i]] - ctx->rest[ctx->E[2 * i + 1]])); };
    std::vector<double> vea(nV, -1), vfa(nV, -1);
    for (int i = 0; i < nE; i++) {
        const double len = edge_len(i);
        for (int k = 0; k < 2; k++) {
            double& a = vea[ctx->E[2 * i + k]];
            a = std::max(a, 0.0);
            a += 0.5 * len;
        }
    }
    ctx->edge_areas.assign(nE, -1);
    for (int i = 0; i < nF; i++) {
        const V3 a = ctx->rest[ctx->F[3 * i]], b = ctx->rest[ctx->F[3 * i + 1]], c = ctx->rest[ctx->F[3 * i + 2]];
        const double face_area = 0.5 * std::sqrt(sqnorm(cross(b - a, c - a))); // geometry/area.cpp triangle_area
        for (int k = 0; k < 3; k++) {
            double& va = vfa[ctx->F[3 * i + k]];
            va = std::max(va, 0.0);
            va += face_area / 3.0;
            double& ea = ctx->edge_areas[ctx->F2E[3 * i + k]];
            ea = std::max(ea, 0.0);
            ea += face_area / 3.0;
        }
    }
    ctx->vertex_areas.resize(nV);
    for (int i = 0; i < nV; i++) ctx->vertex_areas[i] = vfa[i] < 0 ? (vea[i] < 0 ? 1.0 : vea[i]) : vfa[i];
    for (int i = 0; i < nE; i++)
        if (ctx->edge_areas[i] < 0) ctx->edge_areas[i] = edge_len(i);
    // init_adjacencies (collision_mesh.cpp:247-307)
    auto dedupe = [](std::vector<std::vector<int32_t>>& v) {
        for (auto& a : v) {
            std::sort(a.begin(), a.end());
            a.erase(std::unique(a.begin(), a.end()), a.end());
        }
    };
    ctx->vv_adj.assign(nV, {}), ctx->ve_adj.assign(nV, {}), ctx->ev_adj.assign(nE, {});
    for (int i = 0; i < nE; i++) {
        const int32_t a = ctx->E[2 * i], b = ctx->E[2 * i + 1];
        ctx->vv_adj[a].push_back(b), ctx->vv_adj[b].push_back(a);
        ctx->ve_adj[a].push_back(i), ctx->ve_adj[b].push_back(i);
    }
    for (int i = 0; i < nF; i++)
        for (int j = 0; j < 3; j++) ctx->ev_adj[ctx->F2E[3 * i + j]].push_back(ctx->F[3 * i + (j + 2) % 3]);
    dedupe(ctx->vv_adj), dedupe(ctx->ve_adj), dedupe(ctx->ev_adj);
    ctx->on_boundary.assign(nV, 1); // 3D: a vertex of an edge shared by two triangles is not on the boundary (:283-292)
    for (int i = 0; i < nE; i++)
        if (ctx->ev_adj[i].size() >= 2) ctx->on_boundary[ctx->E[2 * i]] = ctx->on_boundary[ctx->E[2 * i + 1]] = 0;
    ctx->built = false;
    for (auto& c : ctx->cand) c.clear();
    for (auto& c : ctx->coll) c.clear();
    return 0;
}
int ipco_mesh_set_collision_filter(ipcb_ctx* ctx, const int32_t* patch_ids, int32_t n_dynamic)
{
    ctx->patch_ids.clear();
    if (patch_ids) ctx->patch_ids.assign(patch_ids, patch_ids + ctx->nV);
    ctx->n_dynamic = n_dynamic;
    return 0;
}
int ipco_mesh_num_codim_vertices(ipcb_ctx* ctx, int32_t* n)
{
    *n = int32_t(ctx->codim_vertices.size());
    return 0;
}
int ipco_mesh_num_codim_edges(ipcb_ctx* ctx, int32_t* n)
{
    *n = int32_t(ctx->codim_edges.size());
    return 0;
}
int ipco_mesh_faces_to_edges(ipcb_ctx* ctx, int32_t* f2e)
{
    for (int i = 0; i < ctx->nF; i++)
        for (int k = 0; k < 3; k++) f2e[i + size_t(ctx->nF) * k] = ctx->F2E[3 * i + k];
    return 0;
}
int ipco_mesh_areas(ipcb_ctx* ctx, double* va, double* ea)
{
    std::copy(ctx->vertex_areas.begin(), ctx->vertex_areas.end(), va);
    std::copy(ctx->edge_areas.begin(), ctx->edge_areas.end(), ea);
    return 0;
}

int ipco_broad_build_static(ipcb_ctx* ctx, const double* V, int32_t ld, double r, int32_t boxes)
{
    build_boxes(ctx, load_vertices(ctx->nV, V, ld), nullptr, r, boxes);
    return 0;
}
int ipco_broad_build_swept(ipcb_ctx* ctx, const double* V0, const double* V1, int32_t ld, double r, int32_t boxes)
{
    const auto v1 = load_vertices(ctx->nV, V1, ld);
    build_boxes(ctx, load_vertices(ctx->nV, V0, ld), &v1, r, boxes);
    return 0;
}
int ipco_broad_detect(ipcb_ctx* ctx, int32_t kind, int64_t* count)
{
    if (!ctx->built) return fail("broad phase not built");
    if (kind < 0 || kind > 5) return fail("bad candidate kind");
    broad_detect_kind(ctx, kind, ctx->detected[kind]);
    *count = int64_t(ctx->detected[kind].size());
    return 0;
}
int ipco_broad_fetch(ipcb_ctx* ctx, int32_t kind, int32_t* pairs)
{
    std::memcpy(pairs, ctx->detected[kind].data(), ctx->detected[kind].size() * sizeof(Pair));
    return 0;
}
int ipco_broad_vertex_boxes(ipcb_ctx* ctx, void* boxes)
{
    for (int i = 0; i < ctx->nV; i++)
        for (int c = 0; c < 3; c++) {
            if (ctx->boxes_mode == IPCB_BOXES_FLOAT) {
                static_cast<float*>(boxes)[6 * i + c] = float(ctx->vbox[i].lo[c]);
                static_cast<float*>(boxes)[6 * i + 3 + c] = float(ctx->vbox[i].hi[c]);
            } else {
                static_cast<double*>(boxes)[6 * i + c] = ctx->vbox[i].lo[c];
                static_cast<double*>(boxes)[6 * i + 3 + c] = ctx->vbox[i].hi[c];
            }
        }
    return 0;
}

static void fill_counts4(const std::vector<Pair>* c, int64_t counts[4])
{
    for (int k = 0; k < 4; k++) counts[k] = int64_t(c[k].size());
}
int ipco_candidates_build_static(ipcb_ctx* ctx, const double* V, int32_t ld, double r, int64_t counts[4])
{
    candidates_build(ctx, load_vertices(ctx->nV, V, ld), nullptr, r);
    fill_counts4(ctx->cand, counts);
    return 0;
}
int ipco_candidates_build_swept(ipcb_ctx* ctx, const double* V0, const double* V1, int32_t ld, double r, int64_t counts[4])
{
    const auto v1 = load_vertices(ctx->nV, V1, ld);
    candidates_build(ctx, load_vertices(ctx->nV, V0, ld), &v1, r);
    fill_counts4(ctx->cand, counts);
    return 0;
}
int ipco_candidates_fetch(ipcb_ctx* ctx, int32_t kind, int32_t* pairs)
{
    if (kind < 0 || kind > 3) return fail("bad candidate kind");
    std::memcpy(pairs, ctx->cand[kind].data(), ctx->cand[kind].size() * sizeof(Pair));
    return 0;
}
int ipco_candidates_set(ipcb_ctx* ctx, int32_t kind, int64_t count, const int32_t* pairs)
{
    if (kind < 0 || kind > 3) return fail("bad candidate kind");
    ctx->cand[kind].resize(count);
    std::memcpy(ctx->cand[kind].data(), pairs, count * sizeof(Pair));
    // unordered kinds are kept as (min, max), like every broad phase emits them (the distance type of an edge-edge
    // collision is relative to the stored order)
    if (kind == IPCB_VV || kind == IPCB_EE)
        for (auto& p : ctx->cand[kind])
            if (p[0] > p[1]) std::swap(p[0], p[1]);
    return 0;
}

static void coll_counts(ipcb_ctx* ctx, int64_t counts[4])
{
    for (int k = 0; k < 4; k++) counts[k] = int64_t(ctx->coll[k].size());
}
int ipco_collisions_build_from_candidates(ipcb_ctx* ctx, const double* V, int32_t ld, double dhat, double dmin, int32_t flags,
                                          int64_t counts[4])
{
    collisions_build(ctx, load_vertices(ctx->nV, V, ld), dhat, dmin, flags);
    coll_counts(ctx, counts);
    return 0;
}
int ipco_collisions_build(ipcb_ctx* ctx, const double* V, int32_t ld, double dhat, double dmin, int32_t flags, int64_t counts[4])
{
    const auto v = load_vertices(ctx->nV, V, ld);
    candidates_build(ctx, v, nullptr, 0.5 * (dhat + dmin)); // normal_collisions.cpp:30
    collisions_build(ctx, v, dhat, dmin, flags);
    coll_counts(ctx, counts);
    return 0;
}
int ipco_collisions_corrections_keys(ipcb_ctx* ctx, int64_t n[4])
{
    if (!ctx->ima_pending) return fail("no deferred IMPROVED_MAX_APPROX build on this context");
    for (int k = 0; k < 4; k++) n[k] = int64_t(ctx->ima_keys[k].size());
    return 0;
}
int ipco_collisions_corrections_pack(ipcb_ctx* ctx, uint64_t* keys)
{
    if (!ctx->ima_pending) return fail("no deferred IMPROVED_MAX_APPROX build on this context");
    size_t o = 0;
    for (int k = 0; k < 4; k++)
        for (const Pair& p : ctx->ima_keys[k]) keys[o++] = (uint64_t(uint32_t(p[0])) << 32) | uint32_t(p[1]);
    return 0;
}
int ipco_collisions_corrections_apply(ipcb_ctx* ctx, const uint64_t* keys, const int64_t n[4], int32_t rank, int32_t world, int64_t counts[4])
{
    if (!ctx->ima_pending) return fail("no deferred IMPROVED_MAX_APPROX build on this context");
    if (world < 1 || rank < 0 || rank >= world) return fail("bad slice");
    std::vector<Pair> lists[4];
    size_t o = 0;
    for (int k = 0; k < 4; k++)
        for (int64_t i = 0; i < n[k]; i++, o++) lists[k].push_back({ int32_t(keys[o] >> 32), int32_t(keys[o] & 0xffffffffu) });
    collisions_corrections_apply(ctx, lists, rank, world);
    coll_counts(ctx, counts);
    return 0;
}
int ipco_collisions_fetch(ipcb_ctx* ctx, int32_t kind, int32_t* ids, double* weight, double* eps_x, uint8_t* dtype)
{
    if (kind < 0 || kind > 3) return fail("bad collision kind");
    const auto& cs = ctx->coll[kind];
    for (size_t i = 0; i < cs.size(); i++) {
        if (ids) ids[2 * i] = cs[i].a, ids[2 * i + 1] = cs[i].b;
        if (weight) weight[i] = cs[i].w;
        if (eps_x) eps_x[i] = cs[i].eps_x;
        if (dtype) dtype[i] = cs[i].dtype;
    }
    return 0;
}
int ipco_collisions_clear(ipcb_ctx* ctx)
{
    for (int k = 0; k < 4; k++) ctx->coll[k].clear(), ctx->appended[k].clear();
    return 0;
}
int ipco_collisions_append(ipcb_ctx* ctx, int32_t kind, int64_t count, const int32_t* ids, const double* weight, const double* eps_x,
                           const uint8_t* dtype)
{
    if (kind < 0 || kind > 3) return fail("bad collision kind");
    if (kind == IPCB_EE && count > 0 && (!eps_x || !dtype)) return fail("edge-edge collision records need eps_x and dtype");
    for (int64_t i = 0; i < count; i++) {
        int32_t a = ids[2 * i], b = ids[2 * i + 1];
        if (kind == IPCB_VV && a > b) std::swap(a, b); // an edge-edge record keeps its orientation: its dtype refers to it
        ctx->appended[kind].push_back({ a, b, weight[i], kind == IPCB_EE ? eps_x[i] : 0.0, kind == IPCB_EE ? dtype[i] : uint8_t(0) });
    }
    return 0;
}
int ipco_collisions_merge(ipcb_ctx* ctx, double dmin, int32_t /* flags: a hint, the set is the same */, int64_t counts[4])
{
    for (int k = 0; k < 4; k++) {
        merge_collisions(ctx, k, ctx->appended[k]);
        ctx->appended[k].clear();
    }
    ctx->dmin = dmin;
    coll_counts(ctx, counts);
    return 0;
}
struct ipcb_collision_set {
    std::vector<Coll> coll[4];
    double dmin = 0;
};
int ipco_collision_set_create(ipcb_ctx*, ipcb_collision_set** out)
{
    *out = new ipcb_collision_set();
    return 0;
}
void ipco_collision_set_destroy(ipcb_collision_set* set) { delete set; }
int ipco_collisions_swap(ipcb_ctx* ctx, ipcb_collision_set* set, int64_t counts[4])
{
    for (int k = 0; k < 4; k++) ctx->coll[k].swap(set->coll[k]);
    std::swap(ctx->dmin, set->dmin);
    coll_counts(ctx, counts);
    return 0;
}
int ipco_ctx_set_collision_range(ipcb_ctx* ctx, int32_t rank, int32_t world)
{
    if (world < 1 || rank < 0 || rank >= world) return fail("bad collision range");
    ctx->coll_rank = rank, ctx->coll_world = world;
    return 0;
}
int ipco_ctx_set_row_block(ipcb_ctx* ctx, int32_t v_begin, int32_t v_end)
{
    if (v_end >= 0 && (v_begin < 0 || v_begin > v_end)) return fail("bad row block");
    ctx->row_lo = v_end < 0 ? 0 : v_begin, ctx->row_hi = v_end < 0 ? -1 : v_end;
    return 0;
}
// same rule as the product: per vertex the number of 3x3 blocks its column receives, boundaries at r * total / world
int ipco_hessian_balanced_row_blocks(ipcb_ctx* ctx, int32_t world, int32_t* bounds)
{
    if (world < 1) return fail("bad world size");
    std::vector<long long> prefix(size_t(ctx->nV) + 1, 0);
    for (int k = 0; k < 4; k++)
        for (const Coll& c : ctx->coll[k]) {
            int32_t ids[4];
            const int n = stencil_ids(ctx, k, c.a, c.b, ids);
            for (int j = 0; j < n; j++) prefix[size_t(ids[j]) + 1] += n;
        }
    for (int v = 0; v < ctx->nV; v++) prefix[v + 1] += prefix[v];
    const long long total = prefix[ctx->nV];
    for (int r = 0; r <= world; r++) {
        const long long want = total * r / world;
        bounds[r] = r == 0 ? 0 : (r == world ? ctx->nV : int32_t(std::lower_bound(prefix.begin(), prefix.begin() + ctx->nV, want) - prefix.begin()));
    }
    for (int r = 1; r <= world; r++) bounds[r] = std::max(bounds[r], bounds[r - 1]);
    return 0;
}
int ipco_collisions_min_distance(ipcb_ctx* ctx, const double* Vp, int32_t ld, double* out)
{
    const auto V = load_vertices(ctx->nV, Vp, ld);
    double best = std::numeric_limits<double>::infinity();
    for (int k = 0; k < 4; k++)
        for (const Coll& c : ctx->coll[k]) {
            int32_t ids[4];
            const int n = stencil_ids(ctx, k, c.a, c.b, ids);
            V3 x[4];
            for (int j = 0; j < n; j++) x[j] = V[ids[j]];
            best = std::min(best, prim_value(collision_embed(k, c), x));
        }
    *out = best;
    return 0;
}

int ipco_barrier_energy(ipcb_ctx* ctx, const double* Vp, int32_t ld, const ipcb_barrier_params* bp, double* energy)
{
    const auto V = load_vertices(ctx->nV, Vp, ld);
    const Barrier B = { bp->dhat, bp->stiffness, bp->use_physical_barrier != 0 };
    double total = 0;
    for (int k = 0; k < 4; k++) {
        const auto& cs = ctx->coll[k];
        double sum = 0;
        const size_t lo = cs.size() * size_t(ctx->coll_rank) / size_t(ctx->coll_world), hi = cs.size() * size_t(ctx->coll_rank + 1) / size_t(ctx->coll_world);
#pragma omp parallel for reduction(+ : sum) schedule(static)
        for (size_t i = lo; i < hi; i++) {
            int32_t ids[4];
            const int n = stencil_ids(ctx, k, cs[i].a, cs[i].b, ids);
            V3 x[4];
            for (int j = 0; j < n; j++) x[j] = V[ids[j]];
            sum += collision_energy(k, cs[i], x, B, ctx->dmin);
        }
        total += sum;
    }
    *energy = total;
    return 0;
}

int ipco_barrier_gradient(ipcb_ctx* ctx, const double* Vp, int32_t ld, const ipcb_barrier_params* bp, double* grad)
{
    const auto V = load_vertices(ctx->nV, Vp, ld);
    const Barrier B = { bp->dhat, bp->stiffness, bp->use_physical_barrier != 0 };
    const size_t ndof = 3 * size_t(ctx->nV);
    const int nt = omp_get_max_threads();
    std::vector<std::vector<double>> loc(nt); // tbb::combinable<VectorXd> (potential.cpp:74-94)
    for (int k = 0; k < 4; k++) {
        const auto& cs = ctx->coll[k];
        const size_t lo = cs.size() * size_t(ctx->coll_rank) / size_t(ctx->coll_world), hi = cs.size() * size_t(ctx->coll_rank + 1) / size_t(ctx->coll_world);
#pragma omp parallel for schedule(static)
        for (size_t i = lo; i < hi; i++) {
            auto& mine = loc[omp_get_thread_num()];
            if (mine.empty()) mine.assign(ndof, 0.0);
            int32_t ids[4];
            const int n = stencil_ids(ctx, k, cs[i].a, cs[i].b, ids);
            V3 x[4];
            for (int j = 0; j < n; j++) x[j] = V[ids[j]];
            double g[12];
            collision_gradient(k, cs[i], x, B, ctx->dmin, g);
            for (int j = 0; j < n; j++) // utils/local_to_global.hpp:21-45 (RowMajor)
                for (int c = 0; c < 3; c++) mine[3 * size_t(ids[j]) + c] += g[3 * j + c];
        }
    }
    std::fill(grad, grad + ndof, 0.0);
    for (auto& l : loc)
        if (!l.empty())
            for (size_t i = 0; i < ndof; i++) grad[i] += l[i];
    return 0;
}

struct Trip {
    int32_t col, row;
    double val;
};
// setFromTriplets (potential.cpp:218) into ctx->outer / inner / vals
static void assemble_triplets(ipcb_ctx* ctx, std::vector<std::vector<Trip>>& loc, int ndof, int64_t* nnz)
{
    std::vector<Trip> all;
    size_t total = 0;
    for (auto& l : loc) total += l.size();
    all.reserve(total);
    for (auto& l : loc) all.insert(all.end(), l.begin(), l.end());
    // setFromTriplets (potential.cpp:218): compressed columns, rows ascending, duplicates summed
    __gnu_parallel::stable_sort(all.begin(), all.end(), [](const Trip& a, const Trip& b) {
        return a.col != b.col ? a.col < b.col : a.row < b.row;
    });
    ctx->outer.assign(ndof + 1, 0);
    ctx->inner.clear();
    ctx->vals.clear();
    for (size_t i = 0; i < all.size();) {
        size_t j = i;
        double s = 0;
        while (j < all.size() && all[j].col == all[i].col && all[j].row == all[i].row) s += all[j++].val;
        ctx->inner.push_back(all[i].row);
        ctx->vals.push_back(s);
        ctx->outer[all[i].col + 1]++;
        i = j;
    }
    for (int c = 0; c < ndof; c++) ctx->outer[c + 1] += ctx->outer[c];
    *nnz = int64_t(ctx->inner.size());
}

int ipco_barrier_hessian(ipcb_ctx* ctx, const double* Vp, int32_t ld, const ipcb_barrier_params* bp, int32_t psd_mode, int64_t* nnz)
{
    const auto V = load_vertices(ctx->nV, Vp, ld);
    const Barrier B = { bp->dhat, bp->stiffness, bp->use_physical_barrier != 0 };
    const int ndof = 3 * ctx->nV;
    const int nt = omp_get_max_threads();
    const int row_lo = ctx->row_hi < 0 ? 0 : ctx->row_lo, row_hi = ctx->row_hi < 0 ? ctx->nV : ctx->row_hi;
    std::vector<std::vector<Trip>> loc(nt);
    for (int k = 0; k < 4; k++) {
        const auto& cs = ctx->coll[k];
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < cs.size(); i++) {
            auto& mine = loc[omp_get_thread_num()];
            int32_t ids[4];
            const int n = stencil_ids(ctx, k, cs[i].a, cs[i].b, ids);
            V3 x[4];
            for (int j = 0; j < n; j++) x[j] = V[ids[j]];
            bool any_owned = false;
            for (int j = 0; j < n; j++) any_owned |= ids[j] >= row_lo && ids[j] < row_hi;
            if (!any_owned) continue; // row block of a sharded Hessian: this collision touches no owned vertex
            double H[144];
            collision_hessian(k, cs[i], x, B, ctx->dmin, psd_mode, H);
            // utils/local_to_global.hpp:263-305: exact zeros are skipped
            for (int a = 0; a < n; a++)
                for (int b = 0; b < n; b++) {
                    if (ids[b] < row_lo || ids[b] >= row_hi) continue; // columns (== rows) of other ranks
                    for (int r = 0; r < 3; r++)
                        for (int c = 0; c < 3; c++) {
                            const double val = H[(3 * a + r) + 12 * (3 * b + c)];
                            if (val != 0) mine.push_back({ 3 * ids[b] + c, 3 * ids[a] + r, val });
                        }
                }
        }
    }
    assemble_triplets(ctx, loc, ndof, nnz);
    return 0;
}
int ipco_barrier_hessian_fetch(ipcb_ctx* ctx, int32_t* outer, int32_t* inner, double* values)
{
    std::copy(ctx->outer.begin(), ctx->outer.end(), outer);
    std::copy(ctx->inner.begin(), ctx->inner.end(), inner);
    std::copy(ctx->vals.begin(), ctx->vals.end(), values);
    return 0;
}

// ---- ipc::has_intersections (ipc.cpp:105-166), 3D ------------------------------------------------------------------
// exact orientation by rational-free expansion arithmetic is overkill for a checker: long double products of the
// determinant are NOT exact, so the checker decides the sign with exact integer arithmetic on the doubles' binary
// representation (every double is m * 2^e; the determinant's monomials are brought to a common exponent and summed in
// arbitrary precision).
namespace {
struct Big { // signed arbitrary-precision integer, little-endian base 2^32 magnitude
    std::vector<uint32_t> mag;
    int sign = 0;
};
Big big_from(uint64_t v, int sign)
{
    Big b;
    if (v == 0 || sign == 0) return b;
    b.sign = sign;
    b.mag = { uint32_t(v), uint32_t(v >> 32) };
    while (!b.mag.empty() && b.mag.back() == 0) b.mag.pop_back();
    return b;
}
Big big_mul(const Big& a, const Big& b)
{
    Big r;
    if (a.sign == 0 || b.sign == 0) return r;
    r.sign = a.sign * b.sign;
    r.mag.assign(a.mag.size() + b.mag.size(), 0);
    for (size_t i = 0; i < a.mag.size(); i++) {
        uint64_t carry = 0;
        for (size_t j = 0; j < b.mag.size() || carry; j++) {
            uint64_t cur = r.mag[i + j] + carry + (j < b.mag.size() ? uint64_t(a.mag[i]) * b.mag[j] : 0);
            r.mag[i + j] = uint32_t(cur);
            carry = cur >> 32;
        }
    }
    while (!r.mag.empty() && r.mag.back() == 0) r.mag.pop_back();
    return r;
}
Big big_shl(const Big& a, int bits)
{
    Big r;
    if (a.sign == 0) return r;
    r.sign = a.sign;
    r.mag.assign(a.mag.size() + size_t(bits / 32) + 1, 0);
    const int w = bits / 32, s = bits % 32;
    for (size_t i = 0; i < a.mag.size(); i++) {
        const uint64_t v = uint64_t(a.mag[i]) << s;
        r.mag[i + w] |= uint32_t(v);
        r.mag[i + w + 1] |= uint32_t(v >> 32);
    }
    while (!r.mag.empty() && r.mag.back() == 0) r.mag.pop_back();
    return r;
}
int big_cmp_mag(const Big& a, const Big& b)
{
    if (a.mag.size() != b.mag.size()) return a.mag.size() < b.mag.size() ? -1 : 1;
    for (size_t i = a.mag.size(); i-- > 0;)
        if (a.mag[i] != b.mag[i]) return a.mag[i] < b.mag[i] ? -1 : 1;
    return 0;
}
Big big_add(const Big& a, const Big& b)
{
    if (a.sign == 0) return b;
    if (b.sign == 0) return a;
    Big r;
    if (a.sign == b.sign) {
        r.sign = a.sign;
        r.mag.assign(std::max(a.mag.size(), b.mag.size()) + 1, 0);
        uint64_t carry = 0;
        for (size_t i = 0; i < r.mag.size(); i++) {
            const uint64_t cur = carry + (i < a.mag.size() ? a.mag[i] : 0) + (i < b.mag.size() ? b.mag[i] : 0);
            r.mag[i] = uint32_t(cur);
            carry = cur >> 32;
        }
    } else {
        const int c = big_cmp_mag(a, b);
        if (c == 0) return r;
        const Big& hi = c > 0 ? a : b;
        const Big& lo = c > 0 ? b : a;
        r.sign = hi.sign;
        r.mag.assign(hi.mag.size(), 0);
        int64_t borrow = 0;
        for (size_t i = 0; i < hi.mag.size(); i++) {
            int64_t cur = int64_t(hi.mag[i]) - borrow - (i < lo.mag.size() ? int64_t(lo.mag[i]) : 0);
            borrow = cur < 0;
            if (cur < 0) cur += (int64_t(1) << 32);
            r.mag[i] = uint32_t(cur);
        }
    }
    while (!r.mag.empty() && r.mag.back() == 0) r.mag.pop_back();
    return r;
}
// exact sign of det [a - d; b - d; c - d]
int orient3d_exact(const V3& a, const V3& b, const V3& c, const V3& d)
{
    // every coordinate as integer mantissa * 2^exp with a common exponent
    const double* pts[4] = { &a.x, &b.x, &c.x, &d.x };
    int emin = 100000;
    for (int p = 0; p < 4; p++)
        for (int k = 0; k < 3; k++) {
            if (pts[p][k] == 0) continue;
            int e;
            std::frexp(pts[p][k], &e);
            emin = std::min(emin, e - 53);
        }
    if (emin == 100000) return 0;
    Big I[4][3];
    for (int p = 0; p < 4; p++)
        for (int k = 0; k < 3; k++) {
            const double v = pts[p][k];
            if (v == 0) continue;
            int e;
            const double m = std::frexp(std::abs(v), &e); // v = m 2^e, m in [0.5, 1)
            const uint64_t mant = uint64_t(std::ldexp(m, 53));
            I[p][k] = big_shl(big_from(mant, v < 0 ? -1 : 1), (e - 53) - emin);
        }
    auto sub = [](const Big& x, const Big& y) {
        Big ny = y;
        ny.sign = -ny.sign;
        return big_add(x, ny);
    };
    Big m[3][3];
    for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) m[r][k] = sub(I[r][k], I[3][k]);
    auto det2 = [&](int r0, int r1, int c0, int c1) { return sub(big_mul(m[r0][c0], m[r1][c1]), big_mul(m[r0][c1], m[r1][c0])); };
    Big det = big_mul(m[0][0], det2(1, 2, 1, 2));
    det = sub(det, big_mul(m[0][1], det2(1, 2, 0, 2)));
    det = big_add(det, big_mul(m[0][2], det2(1, 2, 0, 1)));
    return det.sign;
}
// geometry/intersection.cpp:115-145 (the LU solve in double like the default build of the reference)
bool edge_intersects_triangle(const V3& e0, const V3& e1, const V3& t0, const V3& t1, const V3& t2)
{
    const int o1 = orient3d_exact(t0, t1, t2, e0), o2 = orient3d_exact(t0, t1, t2, e1);
    if (o1 != 0 && o2 != 0 && o1 == o2) return false;
    double M[3][3] = { { t1.x - t0.x, t2.x - t0.x, e0.x - e1.x }, { t1.y - t0.y, t2.y - t0.y, e0.y - e1.y }, { t1.z - t0.z, t2.z - t0.z, e0.z - e1.z } };
    double b[3] = { e0.x - t0.x, e0.y - t0.y, e0.z - t0.z };
    int perm[3] = { 0, 1, 2 }, rank = 3;
    for (int k = 0; k < 3; k++) { // Eigen::FullPivLU
        int pr = k, pc = k;
        double best = 0;
        for (int i = k; i < 3; i++)
            for (int j = k; j < 3; j++)
                if (std::abs(M[i][j]) > best) best = std::abs(M[i][j]), pr = i, pc = j;
        if (best == 0) {
            rank = k;
            break;
        }
        for (int j = 0; j < 3; j++) std::swap(M[k][j], M[pr][j]);
        std::swap(b[k], b[pr]);
        for (int i = 0; i < 3; i++) std::swap(M[i][k], M[i][pc]);
        std::swap(perm[k], perm[pc]);
        for (int i = k + 1; i < 3; i++) {
            const double l = M[i][k] / M[k][k];
            for (int j = k + 1; j < 3; j++) M[i][j] -= l * M[k][j];
            b[i] -= l * b[k];
        }
    }
    double y[3] = { 0, 0, 0 }, uvt[3] = { 0, 0, 0 };
    for (int k = rank - 1; k >= 0; k--) {
        double acc = b[k];
        for (int j = k + 1; j < rank; j++) acc -= M[k][j] * y[j];
        y[k] = acc / M[k][k];
    }
    for (int k = 0; k < 3; k++) uvt[perm[k]] = y[k];
    return uvt[0] >= 0.0 && uvt[1] >= 0.0 && uvt[0] + uvt[1] <= 1.0 && uvt[2] >= 0.0 && uvt[2] <= 1.0;
}
} // namespace
int ipco_has_intersections(ipcb_ctx* ctx, const double* Vp, int32_t ld, int32_t* result)
{
    const auto V = load_vertices(ctx->nV, Vp, ld);
    double ext[3] = { 0, 0, 0 };
    for (int k = 0; k < 3; k++) {
        double lo = INFINITY, hi = -INFINITY;
        for (const V3& v : V) lo = std::min(lo, v[k]), hi = std::max(hi, v[k]);
        ext[k] = V.empty() ? 0.0 : hi - lo;
    }
    const double r = 1e-6 * std::sqrt((ext[0] * ext[0] + ext[1] * ext[1]) + ext[2] * ext[2]); // ipc.cpp:120-121
    build_boxes(ctx, V, nullptr, r, IPCB_BOXES_FLOAT);
    std::vector<Pair> ef;
    broad_detect_kind(ctx, IPCB_EF, ef);
    *result = 0;
    for (const Pair& p : ef) {
        const int32_t* e = &ctx->E[2 * size_t(p[0])];
        const int32_t* f = &ctx->F[3 * size_t(p[1])];
        if (edge_intersects_triangle(V[e[0]], V[e[1]], V[f[0]], V[f[1]], V[f[2]])) {
            *result = 1;
            break;
        }
    }
    return 0;
}

// ---- friction ------------------------------------------------------------------------------
// TangentialCollisions::build(mesh, vertices, collisions, normal_potential, mu_s, mu_k) (tangential_collisions.cpp:62-171)
// from the RESIDENT normal collision set
int ipco_tangential_build(ipcb_ctx* ctx, const double* Vp, int32_t ld, const ipcb_barrier_params* bp, const double* mu_s, const double* mu_k,
                          int64_t counts[4])
{
    using namespace oracle;
    const auto V = load_vertices(ctx->nV, Vp, ld);
    const Barrier B = { bp->dhat, bp->stiffness, bp->use_physical_barrier != 0 };
    const double dmin = ctx->dmin;
    // NormalPotential::force_magnitude (barrier_potential.cpp:33-44, barrier_force_magnitude.cpp:7-15)
    auto force = [&](double d_sqr) {
        const double grad_b = barrier_first_derivative(d_sqr - dmin * dmin, (2 * dmin + B.dhat) * B.dhat);
        double N = -B.kappa * grad_b * 2 * std::sqrt(d_sqr);
        if (B.physical) N *= B.scale(dmin);
        return N;
    };
    auto blend = [](double a, double b) { return (a + b) / 2; }; // default_blend_mu (tangential_collisions.hpp:143-149)
    for (int k = 0; k < 4; k++) {
        ctx->tang[k].clear();
        for (const Coll& c : ctx->coll[k]) {
            int32_t ids[4];
            const int n = stencil_ids(ctx, k, c.a, c.b, ids);
            V3 x[4];
            for (int j = 0; j < n; j++) x[j] = V[ids[j]];
            Tang t {};
            t.a = c.a, t.b = c.b, t.weight = c.w;
            if (k == IPCB_VV) {
                point_point_tangent_basis(x[0], x[1], t.P);
                t.normal_force = force(point_point_distance(x[0], x[1]));
                t.mu_s = blend(mu_s[ids[0]], mu_s[ids[1]]), t.mu_k = blend(mu_k[ids[0]], mu_k[ids[1]]);
            } else if (k == IPCB_EV) {
                t.beta[0] = point_edge_closest_point(x[0], x[1], x[2]);
                point_edge_tangent_basis(x[0], x[1], x[2], t.P);
                t.normal_force = force(point_edge_distance(x[0], x[1], x[2])); // known_dtype() == AUTO (candidates/edge_vertex.hpp:63-66)
                t.mu_s = blend((mu_s[ids[2]] - mu_s[ids[1]]) * t.beta[0] + mu_s[ids[1]], mu_s[ids[0]]);
                t.mu_k = blend((mu_k[ids[2]] - mu_k[ids[1]]) * t.beta[0] + mu_k[ids[1]], mu_k[ids[0]]);
            } else if (k == IPCB_EE) {
                if (edge_edge_cross_squarednorm(x[0], x[1], x[2], x[3]) < c.eps_x) continue; // close to parallel: skipped (:123-126)
                edge_edge_closest_point(x[0], x[1], x[2], x[3], t.beta);
                edge_edge_tangent_basis(x[0], x[1], x[2], x[3], t.P);
                t.normal_force = force(edge_edge_distance(x[0], x[1], x[2], x[3], EE_EA_EB)); // collisions/tangential/edge_edge.hpp:27-31
                t.mu_s = blend((mu_s[ids[1]] - mu_s[ids[0]]) * t.beta[0] + mu_s[ids[0]], (mu_s[ids[3]] - mu_s[ids[2]]) * t.beta[1] + mu_s[ids[2]]);
                t.mu_k = blend((mu_k[ids[1]] - mu_k[ids[0]]) * t.beta[0] + mu_k[ids[0]], (mu_k[ids[3]] - mu_k[ids[2]]) * t.beta[1] + mu_k[ids[2]]);
            } else {
                point_triangle_closest_point(x[0], x[1], x[2], x[3], t.beta);
                point_triangle_tangent_basis(x[0], x[1], x[2], x[3], t.P);
                t.normal_force = force(point_triangle_distance(x[0], x[1], x[2], x[3])); // AUTO
                t.mu_s = blend(mu_s[ids[1]] + t.beta[0] * (mu_s[ids[2]] - mu_s[ids[1]]) + t.beta[1] * (mu_s[ids[3]] - mu_s[ids[1]]), mu_s[ids[0]]);
                t.mu_k = blend(mu_k[ids[1]] + t.beta[0] * (mu_k[ids[2]] - mu_k[ids[1]]) + t.beta[1] * (mu_k[ids[3]] - mu_k[ids[1]]), mu_k[ids[0]]);
            }
            ctx->tang[k].push_back(t);
        }
        counts[k] = int64_t(ctx->tang[k].size());
    }
    return 0;
}
int ipco_tangential_fetch(ipcb_ctx* ctx, int32_t kind, int32_t* ids, double* weight, double* normal_force, double* mu_s, double* mu_k,
                          double* closest_point, double* tangent_basis)
{
    if (kind < 0 || kind > 3) return fail("bad collision kind");
    const auto& ts = ctx->tang[kind];
    for (size_t i = 0; i < ts.size(); i++) {
        const auto& t = ts[i];
        if (ids) ids[2 * i] = t.a, ids[2 * i + 1] = t.b;
        if (weight) weight[i] = t.weight;
        if (normal_force) normal_force[i] = t.normal_force;
        if (mu_s) mu_s[i] = t.mu_s;
        if (mu_k) mu_k[i] = t.mu_k;
        if (closest_point) closest_point[2 * i] = t.beta[0], closest_point[2 * i + 1] = t.beta[1];
        if (tangent_basis) {
            const double P[6] = { t.P[0].x, t.P[0].y, t.P[0].z, t.P[1].x, t.P[1].y, t.P[1].z };
            std::copy(P, P + 6, tangent_basis + 6 * i);
        }
    }
    return 0;
}
// FrictionPotential(eps_v) over the resident tangential set: potential.cpp:36-222 with tangential_potential.cpp:162-325
int ipco_friction_energy(ipcb_ctx* ctx, const double* Up, int32_t ld, double eps_v, double* energy)
{
    const auto U = load_vertices(ctx->nV, Up, ld);
    double total = 0;
    for (int k = 0; k < 4; k++) {
        double sum = 0;
        const auto& ts = ctx->tang[k];
#pragma omp parallel for reduction(+ : sum) schedule(static)
        for (size_t i = 0; i < ts.size(); i++) {
            int32_t ids[4];
            const int n = stencil_ids(ctx, k, ts[i].a, ts[i].b, ids);
            V3 v[4];
            for (int j = 0; j < n; j++) v[j] = U[ids[j]];
            sum += oracle::friction_energy(k, ts[i], v, eps_v);
        }
        total += sum;
    }
    *energy = total;
    return 0;
}
int ipco_friction_gradient(ipcb_ctx* ctx, const double* Up, int32_t ld, double eps_v, double* grad)
{
    const auto U = load_vertices(ctx->nV, Up, ld);
    const size_t ndof = 3 * size_t(ctx->nV);
    std::fill(grad, grad + ndof, 0.0);
    for (int k = 0; k < 4; k++)
        for (const auto& t : ctx->tang[k]) {
            int32_t ids[4];
            const int n = stencil_ids(ctx, k, t.a, t.b, ids);
            V3 v[4];
            for (int j = 0; j < n; j++) v[j] = U[ids[j]];
            double g[12];
            oracle::friction_gradient(k, t, v, eps_v, g);
            for (int j = 0; j < n; j++)
                for (int c = 0; c < 3; c++) grad[3 * size_t(ids[j]) + c] += g[3 * j + c];
        }
    return 0;
}
int ipco_friction_hessian(ipcb_ctx* ctx, const double* Up, int32_t ld, double eps_v, int32_t psd_mode, int64_t* nnz)
{
    const auto U = load_vertices(ctx->nV, Up, ld);
    const int ndof = 3 * ctx->nV;
    std::vector<std::vector<Trip>> loc(1);
    for (int k = 0; k < 4; k++)
        for (const auto& t : ctx->tang[k]) {
            int32_t ids[4];
            const int n = stencil_ids(ctx, k, t.a, t.b, ids);
            V3 v[4];
            for (int j = 0; j < n; j++) v[j] = U[ids[j]];
            double H[144];
            oracle::friction_hessian(k, t, v, eps_v, psd_mode, H);
            for (int a = 0; a < n; a++)
                for (int b = 0; b < n; b++)
                    for (int r = 0; r < 3; r++)
                        for (int c = 0; c < 3; c++) {
                            const double val = H[(3 * a + r) + 12 * (3 * b + c)];
                            if (val != 0) loc[0].push_back({ 3 * ids[b] + c, 3 * ids[a] + r, val }); // local_to_global.hpp:290-291
                        }
        }
    assemble_triplets(ctx, loc, ndof, nnz);
    return 0;
}

double ipco_unit_smooth_mu_f0(double y, double mu_s, double mu_k, double eps_v) { return oracle::smooth_mu_f0(y, mu_s, mu_k, eps_v); }
double ipco_unit_smooth_mu_f1_over_x(double y, double mu_s, double mu_k, double eps_v) { return oracle::smooth_mu_f1_over_x(y, mu_s, mu_k, eps_v); }
double ipco_unit_smooth_mu_f2_x_minus_mu_f1_over_x3(double y, double mu_s, double mu_k, double eps_v)
{
    return oracle::smooth_mu_f2_x_minus_mu_f1_over_x3(y, mu_s, mu_k, eps_v);
}

int ipco_ccd_stepsize_from_candidates(ipcb_ctx* ctx, const double* V0, const double* V1, int32_t ld, double min_distance,
                                      const ipcb_ccd_params* ccd, double* step)
{
    *step = stepsize_from_candidates(ctx, load_vertices(ctx->nV, V0, ld), load_vertices(ctx->nV, V1, ld), min_distance, resolve_ccd(ccd));
    return 0;
}
int ipco_ccd_stepsize(ipcb_ctx* ctx, const double* V0, const double* V1, int32_t ld, double min_distance, const ipcb_ccd_params* ccd,
                      double* step)
{
    const auto v0 = load_vertices(ctx->nV, V0, ld), v1 = load_vertices(ctx->nV, V1, ld);
    candidates_build(ctx, v0, &v1, 0.5 * min_distance); // ipc.cpp:95-96
    *step = stepsize_from_candidates(ctx, v0, v1, min_distance, resolve_ccd(ccd));
    return 0;
}
// Candidates::compute_noncandidate_conservative_stepsize (candidates.cpp:294-338)
static double noncandidate_stepsize(ipcb_ctx* ctx, const std::vector<V3>& disp, double dhat)
{
    size_t total = 0;
    for (auto& c : ctx->cand) total += c.size();
    if (total == 0) return 1.0;
    std::vector<char> is_candidate(ctx->nV, 0);
    for (int k = 0; k < 4; k++)
        for (const Pair& p : ctx->cand[k]) {
            int32_t ids[4];
            const int n = stencil_ids(ctx, k, p[0], p[1], ids);
            for (int j = 0; j < n; j++) is_candidate[ids[j]] = 1;
        }
    double m = 0;
    for (int i = 0; i < ctx->nV; i++)
        if (is_candidate[i]) m = std::max(m, std::sqrt((disp[i][0] * disp[i][0] + disp[i][1] * disp[i][1]) + disp[i][2] * disp[i][2]));
    return 0.5 * dhat / m;
}
int ipco_candidates_noncandidate_stepsize(ipcb_ctx* ctx, const double* displacements, int32_t ld, double dhat, double* step)
{
    *step = noncandidate_stepsize(ctx, load_vertices(ctx->nV, displacements, ld), dhat);
    return 0;
}
// Candidates::compute_cfl_stepsize (candidates.cpp:340-363)
int ipco_candidates_cfl_stepsize(ipcb_ctx* ctx, const double* V0, const double* V1, int32_t ld, double dhat, double min_distance,
                                 const ipcb_ccd_params* ccd, double* step)
{
    const auto v0 = load_vertices(ctx->nV, V0, ld), v1 = load_vertices(ctx->nV, V1, ld);
    const double alpha_c = stepsize_from_candidates(ctx, v0, v1, min_distance, resolve_ccd(ccd));
    std::vector<V3> disp(ctx->nV);
    for (int i = 0; i < ctx->nV; i++) disp[i] = { v1[i][0] - v0[i][0], v1[i][1] - v0[i][1], v1[i][2] - v0[i][2] };
    const double alpha_f = noncandidate_stepsize(ctx, disp, dhat);
    if (alpha_f < 0.5 * alpha_c) return ipco_ccd_stepsize(ctx, V0, V1, ld, min_distance, ccd, step);
    *step = std::min(alpha_c, alpha_f);
    return 0;
}
int ipco_ccd_narrow_phase(ipcb_ctx*, int32_t kind, int64_t n, const double* x_t0, const double* x_t1, double min_distance, double tmax,
                          const ipcb_ccd_params* ccd, uint8_t* hit, double* toi)
{
    if (kind < 0 || kind > 3) return fail("bad candidate kind");
    const ipcb_ccd_params p = resolve_ccd(ccd);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; i++) {
        V3 t0[4], t1[4];
        for (int j = 0; j < 4; j++) {
            t0[j] = { x_t0[12 * i + 3 * j], x_t0[12 * i + 3 * j + 1], x_t0[12 * i + 3 * j + 2] };
            t1[j] = { x_t1[12 * i + 3 * j], x_t1[12 * i + 3 * j + 1], x_t1[12 * i + 3 * j + 2] };
        }
        double t = std::numeric_limits<double>::infinity();
        const bool h = narrow_ccd(kind, t0, t1, min_distance, tmax, p, t);
        hit[i] = h;
        toi[i] = h ? t : std::numeric_limits<double>::infinity();
    }
    return 0;
}

// ---------------------------------------------------------------------------
// Unit-level exports used by tests/ to pin the restatement against the
// reference's known-answer tests (tests/src/tests/distance/*.cpp etc.)
int ipco_unit_distance_type(int kind, const double* x)
{
    const V3* p = reinterpret_cast<const V3*>(x);
    switch (kind) {
    case IPCB_EV: return point_edge_distance_type(p[0], p[1], p[2]);
    case IPCB_EE: return edge_edge_distance_type(p[0], p[1], p[2], p[3]);
    case IPCB_FV: return point_triangle_distance_type(p[0], p[1], p[2], p[3]);
    default: return 0;
    }
}
// value / gradient (12) / hessian (12x12 col-major) of the squared distance of
// a stencil with an explicit distance type (dtype < 0: AUTO)
int ipco_unit_distance(int kind, const double* x, int dtype, double* val, double* grad, double* hess)
{
    const V3* p = reinterpret_cast<const V3*>(x);
    Embed em;
    try {
        switch (kind) {
        case IPCB_VV: em = { PRIM_PP, 2, { 0, 1, 0, 0 } }; break;
        case IPCB_EV: em = embed_point_edge(dtype < 0 ? point_edge_distance_type(p[0], p[1], p[2]) : PE(dtype)); break;
        case IPCB_EE: em = embed_edge_edge(dtype < 0 ? edge_edge_distance_type(p[0], p[1], p[2], p[3]) : EE(dtype)); break;
        default: em = embed_point_triangle(dtype < 0 ? point_triangle_distance_type(p[0], p[1], p[2], p[3]) : PT(dtype)); break;
        }
    } catch (const std::exception& e) {
        return fail(e.what());
    }
    Deriv D;
    embed_deriv(em, p, D);
    if (val) *val = prim_value(em, p);
    if (grad) std::copy(D.g, D.g + 12, grad);
    if (hess) std::copy(D.H, D.H + 144, hess);
    return 0;
}
// mollifier pieces: out[0] = cross sqnorm, out[1] = m(x, eps), grad (12), hess (12x12) of m
int ipco_unit_mollifier(const double* x, double eps_x, double* out, double* grad, double* hess)
{
    const V3* p = reinterpret_cast<const V3*>(x);
    Deriv S;
    edge_edge_cross_squarednorm_deriv(p[0], p[1], p[2], p[3], S);
    const double s = edge_edge_cross_squarednorm(p[0], p[1], p[2], p[3]);
    out[0] = s;
    out[1] = edge_edge_mollifier(s, eps_x);
    const double dm = edge_edge_mollifier_gradient(s, eps_x), ddm = edge_edge_mollifier_hessian(s, eps_x);
    for (int i = 0; i < 12; i++) grad[i] = s < eps_x ? dm * S.g[i] : 0;
    for (int j = 0; j < 12; j++)
        for (int i = 0; i < 12; i++) hess[i + 12 * j] = s < eps_x ? dm * S.h(i, j) + ddm * S.g[i] * S.g[j] : 0;
    return 0;
}
double ipco_unit_mollifier_threshold(const double* rest)
{
    const V3* p = reinterpret_cast<const V3*>(rest);
    return edge_edge_mollifier_threshold(p[0], p[1], p[2], p[3]);
}
void ipco_unit_barrier(double d, double dhat, double out[3])
{
    out[0] = barrier(d, dhat);
    out[1] = barrier_first_derivative(d, dhat);
    out[2] = barrier_second_derivative(d, dhat);
}
int ipco_unit_project_to_psd(int n, double* A, int mode)
{
    try {
        project_to_psd(n, A, n, mode);
    } catch (const std::exception& e) {
        return fail(e.what());
    }
    return 0;
}
uint64_t ipco_unit_morton_3D(double x, double y, double z) { return morton_3D(x, y, z); }

} // extern "C"
