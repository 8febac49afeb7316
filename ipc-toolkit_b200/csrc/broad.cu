// broad.cu — CUDA broad phase behind ipc::BroadPhase.
//
// Replaces (reference src/ipc/): broad_phase/aabb.cpp:35-126 (boxes with
// conservative nextafter inflation), broad_phase/lbvh.cpp:29-41 (outward
// rounding to float), :134-330 (Morton + hierarchy build), :348-797
// (traversal + candidate emission), broad_phase/broad_phase.cpp:12-202 and
// candidates/candidates.cpp:43-222 (Candidates::build orchestration).
//
// Design (B200-first, not the reference's CPU layout):
//  * positions are converted once per call to one 32-byte double4 per vertex;
//  * the tree is a Karras radix tree over Morton keys (sorted with
//    cub::DeviceRadixSort, plumbing).  Node boxes are RANGE UNIONS over the
//    sorted leaf boxes answered by a three-level min / max table (k_rmq_*),
//    so the thread that builds a node also finishes it: no bottom-up pass, no
//    arrival counters (the former k_refit stays behind IPCB_REFIT_BOTTOM_UP
//    as the A/B and test alternative);
//    one internal node = one 64-byte fetch holding BOTH child boxes, so leaves
//    are never fetched during traversal; the face tree is additionally
//    collapsed to 4-wide nodes (k_collapse4 / k_traverse4);
//  * traversal is one thread per Morton-ordered query leaf, warp-synchronous,
//    with hits staged in a per-warp shared-memory queue and flushed with ONE
//    global atomic per ~100 pairs (warp-aggregated compaction) and coalesced
//    int2 stores;
//  * output capacity is learned from the previous step; on overflow the kernel
//    keeps counting and the pass is repeated once with a larger buffer; beyond
//    IPCB_MAX_PAIRS the query leaves are processed in chunks (traverse_chunk);
//  * a sweep-and-prune over axis-sorted boxes is the second method
//    (IPCB_BROAD_SAP), same predicate, same candidate sets.
// The candidate SET is the tree-independent predicate
//   { (i,j) : fbox_i ∩ fbox_j ≠ ∅ (closed) ∧ no shared vertex }
// exactly like the reference's default LBVH (SURVEY §7 hard part 1).
#include "ctx.cuh"
#include "exact_orient3d.hpp"
#include "geom.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cfloat>
#include <climits>
#include <algorithm>
#include <cstdlib>

namespace ipcb {

// ---------------------------------------------------------------------------
// positions
__global__ void k_to_aos(int n, const double* __restrict__ V, int ld, double4* __restrict__ X)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    X[i] = make_double4(V[i], V[i + (size_t)ld], V[i + 2 * (size_t)ld], 0.0);
}

void upload_positions(ipcb_ctx* ctx, const double* hV, int ld, Buf<double>& stage)
{
    stage.reserve(3 * size_t(ctx->nV));
    // copy the three columns into a compact nV x 3 column-major staging buffer (ld may exceed nV)
    for (int k = 0; k < 3; k++)
        IPCB_CUDA(cudaMemcpyAsync(stage.p + size_t(k) * ctx->nV, hV + size_t(k) * ld, sizeof(double) * ctx->nV,
                                  cudaMemcpyHostToDevice, ctx->stream));
}

void convert_positions(ipcb_ctx* ctx, const double* dV, int ld, Buf<double4>& X)
{
    X.reserve(ctx->nV);
    ctx->scene_covers_positions = false; // new positions: the scene box of an earlier build no longer bounds them
    if (ctx->nV == 0) return;
    k_to_aos<<<grid_for(ctx->nV, 256), 256, 0, ctx->stream>>>(ctx->nV, dV, ld, X.p);
    ctx->launches++;
}

// ---------------------------------------------------------------------------
// boxes
__device__ inline float round_down(double v) { return nextafterf(__double2float_rn(v), -INFINITY); }
__device__ inline float round_up(double v) { return nextafterf(__double2float_rn(v), INFINITY); }

__device__ inline void atomic_min_f(float* addr, float v)
{
    if (v >= 0)
        atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else
        atomicMax(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}
__device__ inline void atomic_max_f(float* addr, float v)
{
    if (v >= 0)
        atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else
        atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

__global__ void k_scene_init(float* scene)
{
    if (threadIdx.x < 3) scene[threadIdx.x] = INFINITY;
    else if (threadIdx.x < 6) scene[threadIdx.x] = -INFINITY;
}

// aabb.cpp:35-87 + lbvh.cpp:29-41; also reduces the scene box (broad_phase.cpp:93-125)
__global__ void k_vertex_boxes(int n, const double4* __restrict__ X0, const double4* __restrict__ X1, double r,
                               FBox* __restrict__ box, int4* __restrict__ prim, float* scene)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    if (i < n) {
        const double4 p = X0[i];
        const double c[3] = { p.x, p.y, p.z };
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo[k] = round_down(nextafter(c[k] - r, -INFINITY));
            hi[k] = round_up(nextafter(c[k] + r, INFINITY));
        }
        if (X1) {
            const double4 q = X1[i];
            const double e[3] = { q.x, q.y, q.z };
#pragma unroll
            for (int k = 0; k < 3; k++) {
                lo[k] = fminf(lo[k], round_down(nextafter(e[k] - r, -INFINITY)));
                hi[k] = fmaxf(hi[k], round_up(nextafter(e[k] + r, INFINITY)));
            }
        }
        FBox b;
#pragma unroll
        for (int k = 0; k < 3; k++) b.lo[k] = lo[k], b.hi[k] = hi[k];
        box[i] = b;
        prim[i] = make_int4(i, -1, -1, i);
    }
    // block reduction of the scene bounds
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (lo[k] <= hi[k]) {
                atomic_min_f(scene + k, lo[k]);
                atomic_max_f(scene + 3 + k, hi[k]);
            }
        }
    }
}

// aabb.cpp:89-126: edge / face box = union of its vertex boxes (float rounding
// is monotone, so rounding commutes with min / max)
__global__ void k_edge_boxes(int n, const int2* __restrict__ E, const FBox* __restrict__ vb, FBox* __restrict__ box,
                             int4* __restrict__ prim)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int2 e = E[i];
    const FBox a = vb[e.x], b = vb[e.y];
    FBox o;
#pragma unroll
    for (int k = 0; k < 3; k++) o.lo[k] = fminf(a.lo[k], b.lo[k]), o.hi[k] = fmaxf(a.hi[k], b.hi[k]);
    box[i] = o;
    prim[i] = make_int4(e.x, e.y, -1, i);
}
__global__ void k_face_boxes(int n, const int4* __restrict__ F, const FBox* __restrict__ vb, FBox* __restrict__ box,
                             int4* __restrict__ prim)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 f = F[i];
    const FBox a = vb[f.x], b = vb[f.y], c = vb[f.z];
    FBox o;
#pragma unroll
    for (int k = 0; k < 3; k++)
        o.lo[k] = fminf(fminf(a.lo[k], b.lo[k]), c.lo[k]), o.hi[k] = fmaxf(fmaxf(a.hi[k], b.hi[k]), c.hi[k]);
    box[i] = o;
    prim[i] = make_int4(f.x, f.y, f.z, i);
}
// gather a subset (codimensional vertices / edges) of a PrimSet
// local: the codimensional passes of Candidates::build work on RE-INDEXED vertex subsets, and the collision filter sees
// those ids (candidates.cpp:61,66-77,83-108).  With a filter the vertex ids of the gathered primitives are replaced by
// the local ones (0: keep global ids; 1: a vertex's position in the subset; 2: the table `local_ids`); the primitive id
// (.w, what candidates are made of) stays global.  These passes never test for shared vertices.
__global__ void k_gather_set(int n, const int* __restrict__ ids, const FBox* __restrict__ box, const int4* __restrict__ prim,
                             FBox* __restrict__ obox, int4* __restrict__ oprim, int local, const int2* __restrict__ local_ids)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    obox[i] = box[ids[i]];
    int4 p = prim[ids[i]];
    if (local == 1) p.x = i;
    if (local == 2) p.x = local_ids[i].x, p.y = local_ids[i].y;
    oprim[i] = p;
}

// ---------------------------------------------------------------------------
// Morton keys (math/morton.hpp:23-63, lbvh.cpp:150-168)
__device__ inline unsigned long long expand21(unsigned long long v)
{
    v = (v | v << 32) & 0x1F00000000FFFFull;
    v = (v | v << 16) & 0x1F0000FF0000FFull;
    v = (v | v << 8) & 0x100F00F00F00F00Full;
    v = (v | v << 4) & 0x10C30C30C30C30C3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}
__global__ void k_morton(int n, int bits, int uniform, const FBox* __restrict__ box, const float* __restrict__ scene,
                         unsigned long long* __restrict__ key, int* __restrict__ ord)
{
    const double cells = double(1u << bits);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const FBox b = box[i];
    unsigned long long code = 0;
    // ONE cell size for the three axes (the largest scene extent), unlike the reference's per-axis normalisation
    // (lbvh.cpp:152-166): in a flat scene (stacked cloth) the short axis then only contributes its low bits, the
    // upper levels of the tree split along the long axes and node boxes stay close to cubes.  The candidate set
    // does not depend on the tree; `uniform` = 0 restores the per-axis cells.
    const double wmax = fmax(fmax(double(scene[3]) - double(scene[0]), double(scene[4]) - double(scene[1])), double(scene[5]) - double(scene[2]));
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double w = uniform ? wmax : double(scene[3 + k]) - double(scene[k]);
        double m = w > 0 ? (0.5 * (double(b.lo[k]) + double(b.hi[k])) - double(scene[k])) / w : 0.0;
        m = fmin(fmax(m * cells, 0.0), cells - 1.0);
        code |= expand21((unsigned long long)m) << (2 - k);
    }
    key[i] = code;
    ord[i] = i;
}
__global__ void k_apply_order(int n, const int* __restrict__ ord, const FBox* __restrict__ box, const int4* __restrict__ prim,
                              FBox* __restrict__ sbox, int4* __restrict__ sprim)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = ord[i];
    sbox[i] = box[j];
    sprim[i] = prim[j];
}

// ---------------------------------------------------------------------------
// Karras radix tree over the sorted keys; duplicates are broken by position
// (lbvh.cpp:104-131 uses the same fallback)
__device__ inline int delta(const unsigned long long* __restrict__ key, int n, int i, unsigned long long ki, int j)
{
    if (j < 0 || j >= n) return -1;
    const unsigned long long kj = key[j];
    if (ki == kj) return 64 + __clz(i ^ j);
    return __clzll(ki ^ kj);
}
__global__ void k_karras(int n, const unsigned long long* __restrict__ key, Node* __restrict__ nodes, int* __restrict__ parent,
                         int* __restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const unsigned long long ki = key[i];
    const int d = delta(key, n, i, ki, i + 1) > delta(key, n, i, ki, i - 1) ? 1 : -1;
    const int dmin = delta(key, n, i, ki, i - d);
    int lmax = 2;
    while (delta(key, n, i, ki, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(key, n, i, ki, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(key, n, i, ki, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(key, n, i, ki, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int left = (lo == gamma) ? ~gamma : gamma;
    const int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    nodes[i].child[0] = left;
    nodes[i].child[1] = right;
    nodes[i].split = gamma;
    nodes[i].last = hi;
    parent[left < 0 ? (n - 1 + ~left) : left] = i;
    parent[right < 0 ? (n - 1 + ~right) : right] = i;
    flag[i] = 0;
    if (i == 0) parent[0] = -1;
}

// bottom-up refit: the second thread to arrive at a node computes its child boxes
__global__ void k_refit(int n, const FBox* __restrict__ sbox, Node* nodes, const int* __restrict__ parent, int* flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = parent[n - 1 + i];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(flag + node, 1) == 0) return;
        Node* nd = nodes + node;
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int ch = nd->child[c];
            if (ch < 0) {
                const FBox b = sbox[~ch];
#pragma unroll
                for (int k = 0; k < 3; k++) nd->lo[c][k] = b.lo[k], nd->hi[c][k] = b.hi[k];
            } else {
                const volatile Node* cn = nodes + ch;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    nd->lo[c][k] = fminf(cn->lo[0][k], cn->lo[1][k]);
                    nd->hi[c][k] = fmaxf(cn->hi[0][k], cn->hi[1][k]);
                }
            }
        }
        node = parent[node];
    }
}

// ---------------------------------------------------------------------------
// Node boxes WITHOUT a bottom-up pass.  A Karras node knows the range of sorted leaves below each child, so a child
// box is a range union over the sorted leaf boxes: a three-level range-min/max structure answers it with a handful
// of independent loads, and every node is finished by the thread that built it (no parent links, no arrival
// counters, no dependent chain of tree depth).  Unions are min / max, hence exactly the boxes the bottom-up refit
// (lbvh.cpp:262-330) produces.
//   level A: blocks of 32 leaves        — per leaf the union from its block's start (preA) and to its block's end (sufA)
//   level B: super-blocks of 32 blocks  — the same over the block totals (preB / sufB per block)
//   level C: sparse table over the super-block totals (built by one CTA)
struct Rmq {
    const FBox* sbox;
    FBox *preA, *sufA, *blk, *preB, *sufB, *table;
    int n, nb, ns, levels;
};
__device__ inline FBox box_union(const FBox& a, const FBox& b)
{
    FBox r;
#pragma unroll
    for (int k = 0; k < 3; k++) r.lo[k] = fminf(a.lo[k], b.lo[k]), r.hi[k] = fmaxf(a.hi[k], b.hi[k]);
    return r;
}
__device__ inline FBox box_empty()
{
    FBox r;
#pragma unroll
    for (int k = 0; k < 3; k++) r.lo[k] = INFINITY, r.hi[k] = -INFINITY;
    return r;
}
__device__ inline FBox box_shfl_up(const FBox& b, int o)
{
    FBox r;
#pragma unroll
    for (int k = 0; k < 3; k++) r.lo[k] = __shfl_up_sync(0xffffffffu, b.lo[k], o), r.hi[k] = __shfl_up_sync(0xffffffffu, b.hi[k], o);
    return r;
}
__device__ inline FBox box_shfl_down(const FBox& b, int o)
{
    FBox r;
#pragma unroll
    for (int k = 0; k < 3; k++) r.lo[k] = __shfl_down_sync(0xffffffffu, b.lo[k], o), r.hi[k] = __shfl_down_sync(0xffffffffu, b.hi[k], o);
    return r;
}
// inclusive prefix / suffix unions over the 32 lanes of a warp
__device__ inline void warp_scan_boxes(const FBox& b, int lane, FBox& pre, FBox& suf)
{
    pre = b, suf = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const FBox up = box_shfl_up(pre, o), dn = box_shfl_down(suf, o);
        if (lane >= o) pre = box_union(pre, up);
        if (lane + o < 32) suf = box_union(suf, dn);
    }
}
// apply_order fused with level A: one warp per block of 32 sorted leaves
__global__ void __launch_bounds__(256) k_apply_order_scan(int n, const int* __restrict__ ord, const FBox* __restrict__ box,
                                                          const int4* __restrict__ prim, FBox* __restrict__ sbox, int4* __restrict__ sprim, Rmq q)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    FBox b = box_empty();
    if (i < n) {
        const int j = ord[i];
        b = box[j];
        sbox[i] = b;
        sprim[i] = prim[j];
    }
    FBox pre, suf;
    warp_scan_boxes(b, lane, pre, suf);
    if (i < n) q.preA[i] = pre, q.sufA[i] = suf;
    if (lane == 31 && (i >> 5) < q.nb) q.blk[i >> 5] = pre;
}
// level B: one warp per super-block of 32 blocks
__global__ void __launch_bounds__(256) k_rmq_super(Rmq q)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const FBox v = b < q.nb ? q.blk[b] : box_empty();
    FBox pre, suf;
    warp_scan_boxes(v, lane, pre, suf);
    if (b < q.nb) q.preB[b] = pre, q.sufB[b] = suf;
    if (lane == 31 && (b >> 5) < q.ns) q.table[b >> 5] = pre;
}
// level C: sparse table over the super-block totals, one CTA
__global__ void __launch_bounds__(1024) k_rmq_table(Rmq q)
{
    for (int l = 1; l <= q.levels; l++) {
        __syncthreads();
        const int half = 1 << (l - 1);
        const FBox* src = q.table + size_t(l - 1) * q.ns;
        FBox* dst = q.table + size_t(l) * q.ns;
        for (int s = threadIdx.x; s + 2 * half <= q.ns; s += blockDim.x) dst[s] = box_union(src[s], src[s + half]);
    }
}
// union of the sorted leaf boxes i..j (inclusive)
__device__ inline FBox range_box(const Rmq& q, int i, int j)
{
    const int bi = i >> 5, bj = j >> 5;
    if (bi == bj) {
        if ((i & 31) == 0) return q.preA[j];
        if ((j & 31) == 31 || j == q.n - 1) return q.sufA[i];
        FBox r = q.sbox[i];
        for (int k = i + 1; k <= j; k++) r = box_union(r, q.sbox[k]);
        return r;
    }
    FBox r = box_union(q.sufA[i], q.preA[j]);
    const int b0 = bi + 1, b1 = bj - 1;
    if (b0 > b1) return r;
    const int s0 = b0 >> 5, s1 = b1 >> 5;
    if (s0 == s1) {
        if ((b0 & 31) == 0) return box_union(r, q.preB[b1]);
        if ((b1 & 31) == 31 || b1 == q.nb - 1) return box_union(r, q.sufB[b0]);
        for (int b = b0; b <= b1; b++) r = box_union(r, q.blk[b]);
        return r;
    }
    r = box_union(r, box_union(q.sufB[b0], q.preB[b1]));
    const int t0 = s0 + 1, t1 = s1 - 1;
    if (t0 > t1) return r;
    const int l = 31 - __clz(t1 - t0 + 1);
    const FBox* tab = q.table + size_t(l) * q.ns;
    return box_union(r, box_union(tab[t0], tab[t1 - (1 << l) + 1]));
}
// Karras node + both child boxes in one go
__global__ void k_karras_boxes(int n, const unsigned long long* __restrict__ key, Node* __restrict__ nodes, Rmq q)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const unsigned long long ki = key[i];
    const int d = delta(key, n, i, ki, i + 1) > delta(key, n, i, ki, i - 1) ? 1 : -1;
    const int dmin = delta(key, n, i, ki, i - d);
    int lmax = 2;
    while (delta(key, n, i, ki, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(key, n, i, ki, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(key, n, i, ki, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(key, n, i, ki, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const FBox L = lo == gamma ? q.sbox[lo] : range_box(q, lo, gamma);
    const FBox R = hi == gamma + 1 ? q.sbox[hi] : range_box(q, gamma + 1, hi);
    Node nd;
#pragma unroll
    for (int k = 0; k < 3; k++) nd.lo[0][k] = L.lo[k], nd.hi[0][k] = L.hi[k], nd.lo[1][k] = R.lo[k], nd.hi[1][k] = R.hi[k];
    nd.child[0] = (lo == gamma) ? ~gamma : gamma;
    nd.child[1] = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    nd.split = gamma;
    nd.last = hi;
    nodes[i] = nd;
}

// 4-wide collapse: node i takes the children of its two children (a leaf child stays as it is)
__global__ void k_collapse4(int n, const Node* __restrict__ nodes, Node4* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const Node nd = nodes[i];
    Node4 w;
    int cnt = 0;
    auto put = [&](const float* lo, const float* hi, int child, int last) {
        w.lox[cnt] = lo[0], w.loy[cnt] = lo[1], w.loz[cnt] = lo[2], w.hix[cnt] = hi[0], w.hiy[cnt] = hi[1], w.hiz[cnt] = hi[2];
        w.child[cnt] = child, w.last[cnt] = last;
        cnt++;
    };
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const int ch = nd.child[side];
        if (ch < 0) {
            put(nd.lo[side], nd.hi[side], ch, side == 0 ? nd.split : nd.last);
        } else {
            const Node k = nodes[ch];
            put(k.lo[0], k.hi[0], k.child[0], k.split);
            put(k.lo[1], k.hi[1], k.child[1], k.last);
        }
    }
    for (; cnt < 4; cnt++) {
        w.lox[cnt] = w.loy[cnt] = w.loz[cnt] = INFINITY, w.hix[cnt] = w.hiy[cnt] = w.hiz[cnt] = -INFINITY;
        w.child[cnt] = INT_MIN, w.last[cnt] = -1;
    }
    out[i] = w;
}

// Morton resolution per axis.  The candidate SET does not depend on the tree (SURVEY §7 hard part 1), only the
// traversal cost does, so the keys only need enough cells to separate neighbouring primitives: 10 bits per axis
// (4 radix passes) up to 4M primitives, 13 (5 passes) up to 64M, the reference's 21 (8 passes) beyond.
static int morton_bits(int n) { return n <= (1 << 22) ? 10 : (n <= (1 << 26) ? 13 : 21); }

static void sort_keys(ipcb_ctx* ctx, Tree& t, int n, int end_bit, cudaStream_t s)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, t.key.p, t.key_sorted.p, t.ord.p, t.ord_sorted.p, n, 0, end_bit, s);
    t.tmp.reserve(bytes);
    cub::DeviceRadixSort::SortPairs(t.tmp.p, bytes, t.key.p, t.key_sorted.p, t.ord.p, t.ord_sorted.p, n, 0, end_bit, s);
    ctx->launches += 2 + (end_bit + 7) / 8; // onesweep: histogram + exclusive sum + one pass per 8 bits
}

// Morton-sort a PrimSet; with_nodes additionally builds the hierarchy
static void build_tree(ipcb_ctx* ctx, const PrimSet& ps, Tree& t, bool with_nodes, cudaStream_t s = nullptr)
{
    const int n = ps.n;
    t.n = n;
    t.has_nodes = false;
    if (n == 0) return;
    if (!s) s = ctx->stream;
    t.key.reserve(n), t.key_sorted.reserve(n), t.ord.reserve(n), t.ord_sorted.reserve(n);
    t.sbox.reserve(n), t.sprim.reserve(n);
    const int bits = morton_bits(n);
    k_morton<<<grid_for(n, 256), 256, 0, s>>>(n, bits, getenv("IPCB_MORTON_PER_AXIS") ? 0 : 1, ps.box.p, ctx->scene.p, t.key.p, t.ord.p);
    {
        Stage kt(ctx, "k:radix_sort(morton)", s);
        sort_keys(ctx, t, n, 3 * bits, s);
    }
    Stage kt(ctx, "k:lbvh_nodes", s);
    const bool bottom_up = getenv("IPCB_REFIT_BOTTOM_UP") != nullptr; // A/B + test hook: the arrival-counter refit
    if (!with_nodes || n < 2 || bottom_up) {
        k_apply_order<<<grid_for(n, 256), 256, 0, s>>>(n, t.ord_sorted.p, ps.box.p, ps.prim.p, t.sbox.p, t.sprim.p);
        ctx->launches += 2;
        if (!with_nodes || n < 2) {
            t.has_nodes = with_nodes;
            return;
        }
        t.nodes.reserve(n - 1), t.parent.reserve(2 * size_t(n) - 1), t.flag.reserve(n - 1);
        k_karras<<<grid_for(n - 1, 256), 256, 0, s>>>(n, t.key_sorted.p, t.nodes.p, t.parent.p, t.flag.p);
        k_refit<<<grid_for(n, 256), 256, 0, s>>>(n, t.sbox.p, t.nodes.p, t.parent.p, t.flag.p);
        ctx->launches += 2;
        t.has_nodes = true;
        return;
    }
    // range-query node boxes: apply_order + level A, level B, level C, then every node in one pass
    Rmq q;
    q.n = n, q.nb = (n + 31) / 32, q.ns = (q.nb + 31) / 32;
    q.levels = 0;
    while ((2 << q.levels) <= q.ns) q.levels++;
    const size_t total = 2 * size_t(n) + 3 * size_t(q.nb) + size_t(q.levels + 1) * q.ns;
    t.rmq.reserve(total);
    q.sbox = t.sbox.p;
    q.preA = t.rmq.p, q.sufA = q.preA + n, q.blk = q.sufA + n, q.preB = q.blk + q.nb, q.sufB = q.preB + q.nb, q.table = q.sufB + q.nb;
    t.nodes.reserve(n - 1);
    k_apply_order_scan<<<grid_for(size_t(q.nb) * 32, 256), 256, 0, s>>>(n, t.ord_sorted.p, ps.box.p, ps.prim.p, t.sbox.p, t.sprim.p, q);
    k_rmq_super<<<grid_for(size_t(q.ns) * 32, 256), 256, 0, s>>>(q);
    if (q.levels > 0) k_rmq_table<<<1, 1024, 0, s>>>(q);
    k_karras_boxes<<<grid_for(n - 1, 256), 256, 0, s>>>(n, t.key_sorted.p, t.nodes.p, q);
    ctx->launches += 5;
    t.has_nodes = true;
    // 4-wide collapse (k_collapse4 + k_traverse4).  Measured (C3 / C5, ms): vertex queries against the FACE tree 1.08 -> 0.79 /
    // 7.5 -> 2.3 — a vertex box overlaps few faces, the walk is a chain of dependent fetches and the collapse halves it —,
    // edge self-queries 1.99 -> 2.55 / 6.7 -> 9.1 — the self-query pruning already skips half of every level and four box
    // tests per visit cost more than they save.  Default: the face tree only; IPCB_BVH4=0 none, =1 every tree (A/B, tests).
    static const int wide_mode = getenv("IPCB_BVH4") ? atoi(getenv("IPCB_BVH4")) : -1;
    const bool wide = wide_mode == 1 || (wide_mode == -1 && &t == &ctx->ftree);
    t.has_nodes4 = false;
    if (wide) {
        t.nodes4.reserve(n - 1);
        k_collapse4<<<grid_for(n - 1, 256), 256, 0, s>>>(n, t.nodes.p, t.nodes4.p);
        ctx->launches++;
        t.has_nodes4 = true;
    }
}

__global__ void k_axis_key(int n, int axis, const FBox* __restrict__ box, unsigned long long* __restrict__ key, int* __restrict__ ord);
// sweep-and-prune view of a PrimSet: boxes and primitives sorted by the lower bound along ctx->sap_axis
static void build_sap_view(ipcb_ctx* ctx, const PrimSet& ps, Tree& t, cudaStream_t s)
{
    const int n = ps.n;
    t.n = n;
    t.has_nodes = false;
    if (n == 0) return;
    t.key.reserve(n), t.key_sorted.reserve(n), t.ord.reserve(n), t.ord_sorted.reserve(n);
    t.sbox.reserve(n), t.sprim.reserve(n);
    k_axis_key<<<grid_for(n, 256), 256, 0, s>>>(n, ctx->sap_axis, ps.box.p, t.key.p, t.ord.p);
    sort_keys(ctx, t, n, 32, s);
    k_apply_order<<<grid_for(n, 256), 256, 0, s>>>(n, t.ord_sorted.p, ps.box.p, ps.prim.p, t.sbox.p, t.sprim.p);
    ctx->launches += 2;
}
// the sweep axis: the longest extent of the scene box (one 24-byte read-back per build in this mode)
static void choose_sap_axis(ipcb_ctx* ctx)
{
    float h[6];
    IPCB_CUDA(cudaMemcpyAsync(h, ctx->scene.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
    const float e[3] = { h[3] - h[0], h[4] - h[1], h[5] - h[2] };
    ctx->sap_axis = e[0] >= e[1] && e[0] >= e[2] ? 0 : (e[1] >= e[2] ? 1 : 2);
}
static Tree& ensure_sap_view(ipcb_ctx* ctx, int which)
{
    if (!ctx->sap_axis_ok) choose_sap_axis(ctx), ctx->sap_axis_ok = true;
    Tree* t[3] = { &ctx->sap_v, &ctx->sap_e, &ctx->sap_f };
    const PrimSet* ps[3] = { &ctx->vset, &ctx->eset, &ctx->fset };
    if (!ctx->sap_ok[which]) build_sap_view(ctx, *ps[which], *t[which], ctx->stream), ctx->sap_ok[which] = true;
    return *t[which];
}

void broad_build(ipcb_ctx* ctx, bool swept, double r)
{
    Stage st(ctx, "broad_build");
    cudaStream_t s = ctx->stream;
    const int nV = ctx->nV, nE = ctx->nE, nF = ctx->nF;
    ctx->scene.reserve(8);
    k_scene_init<<<1, 32, 0, s>>>(ctx->scene.p);
    ctx->vset.n = nV, ctx->eset.n = nE, ctx->fset.n = nF;
    ctx->vset.box.reserve(nV), ctx->vset.prim.reserve(nV);
    ctx->eset.box.reserve(nE), ctx->eset.prim.reserve(nE);
    ctx->fset.box.reserve(nF), ctx->fset.prim.reserve(nF);
    if (nV)
        k_vertex_boxes<<<grid_for(nV, 256), 256, 0, s>>>(nV, ctx->X0.p, swept ? ctx->X1.p : nullptr, r, ctx->vset.box.p,
                                                        ctx->vset.prim.p, ctx->scene.p);
    if (nE) k_edge_boxes<<<grid_for(nE, 256), 256, 0, s>>>(nE, ctx->dE.p, ctx->vset.box.p, ctx->eset.box.p, ctx->eset.prim.p);
    if (nF) k_face_boxes<<<grid_for(nF, 256), 256, 0, s>>>(nF, ctx->dF.p, ctx->vset.box.p, ctx->fset.box.p, ctx->fset.prim.p);
    ctx->launches += 4;
    ctx->vtree_ok = ctx->etree_ok = ctx->ftree_ok = false;
    ctx->sap_ok[0] = ctx->sap_ok[1] = ctx->sap_ok[2] = false, ctx->sap_axis_ok = false;
    ctx->vorder_valid = false;
    ctx->vtree.n = ctx->etree.n = ctx->ftree.n = 0;
    ctx->vtree.has_nodes = ctx->etree.has_nodes = ctx->ftree.has_nodes = false;
    ctx->built = true;
    ctx->swept = swept;
    ctx->scene_covers_positions = swept; // reduced from the boxes of X0 and X1
    IPCB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// traversal
constexpr int TRAV_BLOCK = 128;
constexpr int STAGE_CAP = 160; // per-warp staging slots; flushed when > STAGE_CAP - 64

// lbvh.cpp:801-873 with can_vertices_collide == true; QN / TN = vertices of the query / target primitive
template <int QN, int TN> __device__ __forceinline__ bool shares_vertex(int4 a, int4 b)
{
    bool s = a.x == b.x;
    if (TN > 1) s |= a.x == b.y;
    if (TN > 2) s |= a.x == b.z;
    if (QN > 1) {
        s |= a.y == b.x;
        if (TN > 1) s |= a.y == b.y;
        if (TN > 2) s |= a.y == b.z;
    }
    if (QN > 2) {
        s |= a.z == b.x;
        if (TN > 1) s |= a.z == b.y;
        if (TN > 2) s |= a.z == b.z;
    }
    return s;
}

// CollisionFilter descriptor on the device (ipcb_mesh_set_collision_filter)
struct FilterView {
    const int* patch; // per-vertex labels or nullptr
    int n_dynamic;    // < 0: off
};
__device__ __forceinline__ bool can_vertices_collide(const FilterView& f, int vi, int vj)
{
    return (!f.patch || __ldg(f.patch + vi) != __ldg(f.patch + vj)) && (f.n_dynamic < 0 || vi < f.n_dynamic || vj < f.n_dynamic);
}
// broad_phase.cpp:127-202: some pair of vertices of the two primitives can collide
template <int QN, int TN> __device__ __forceinline__ bool any_can_collide(const FilterView& f, int4 a, int4 b)
{
    const int qa[3] = { a.x, a.y, a.z }, tb[3] = { b.x, b.y, b.z };
    bool any = false;
#pragma unroll
    for (int i = 0; i < QN; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) any |= can_vertices_collide(f, qa[i], tb[j]);
    return any;
}

// MODE 0: emit (query, target); 1: emit (target, query); 2: self, emit (min, max)
// flags: bit 0 = reject pairs that share a vertex, bit 1 = apply the collision filter
template <int MODE, int QN, int TN, bool FILTERED>
__global__ void __launch_bounds__(TRAV_BLOCK)
    k_traverse(int q_begin, int q_end, const FBox* __restrict__ qbox, const int4* __restrict__ qprim,
               const Node* __restrict__ nodes, int n_target, const FBox* __restrict__ tbox, const int4* __restrict__ tprim,
               int2* __restrict__ out, unsigned long long* counter, unsigned long long capacity, int flags, FilterView filter)
{
    const bool check_shared = flags & 1;
    constexpr bool filtered = FILTERED; // a template parameter: the unfiltered kernel keeps its 56 registers
    __shared__ int2 stage[TRAV_BLOCK / 32][STAGE_CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qi = q_begin + blockIdx.x * blockDim.x + threadIdx.x;
    bool active = qi < q_end;
    FBox q;
    int4 qp = make_int4(-1, -1, -1, -1);
    if (active) {
        q = qbox[qi];
        qp = qprim[qi];
    }
    int stack[96];
    int sp = 0;
    int node = 0;
    int nstaged = 0;
    int single = -1; // n_target == 1: the only leaf
    if (n_target == 1) {
        if (active && MODE != 2) {
            const FBox b = tbox[0];
            if (q.lo[0] <= b.hi[0] && b.lo[0] <= q.hi[0] && q.lo[1] <= b.hi[1] && b.lo[1] <= q.hi[1] && q.lo[2] <= b.hi[2]
                && b.lo[2] <= q.hi[2])
                single = 0;
        }
    }
    bool walking = active && n_target > 1;
    // Leaf hits need the target primitive's vertex ids (shared-vertex rejection, primitive id): a dependent
    // scattered load.  It is software-pipelined: the ids of the hits found at one node are requested right away
    // and consumed one iteration later, after the NEXT node's fetch has been issued, so the two latencies overlap.
    int pend0 = single, pend1 = -1;
    int4 tp0 = make_int4(0, 0, 0, 0), tp1 = tp0;
    if (pend0 >= 0) tp0 = __ldg(tprim + pend0);
    while (__any_sync(0xffffffffu, walking || pend0 >= 0 || pend1 >= 0)) {
        float4 a = make_float4(0, 0, 0, 0), b = a, c = a;
        int4 d = make_int4(0, 0, 0, 0);
        if (walking) {
            const float4* np = reinterpret_cast<const float4*>(nodes + node);
            a = __ldg(np), b = __ldg(np + 1), c = __ldg(np + 2);
            d = __ldg(reinterpret_cast<const int4*>(np + 3));
        }
        if (__any_sync(0xffffffffu, pend0 >= 0 || pend1 >= 0)) { // most visited nodes have no leaf children that overlap
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int hit = h == 0 ? pend0 : pend1;
                const int4 tp = h == 0 ? tp0 : tp1;
                bool emit = false;
                int2 pr = make_int2(0, 0);
                if (hit >= 0 && (!check_shared || !shares_vertex<QN, TN>(qp, tp)) && (!filtered || any_can_collide<QN, TN>(filter, qp, tp))) {
                    emit = true;
                    if (MODE == 0) pr = make_int2(qp.w, tp.w);
                    else if (MODE == 1) pr = make_int2(tp.w, qp.w);
                    else pr = make_int2(min(qp.w, tp.w), max(qp.w, tp.w));
                }
                const unsigned m = __ballot_sync(0xffffffffu, emit);
                if (m) {
                    if (emit) stage[warp][nstaged + __popc(m & ((1u << lane) - 1))] = pr;
                    nstaged += __popc(m);
                }
            }
            if (nstaged > STAGE_CAP - 64) {
                __syncwarp();
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(counter, (unsigned long long)nstaged);
                base = __shfl_sync(0xffffffffu, base, 0);
                for (int k = lane; k < nstaged; k += 32)
                    if (base + k < capacity) out[base + k] = stage[warp][k];
                nstaged = 0;
                __syncwarp();
            }
        }
        pend0 = pend1 = -1;
        if (walking) {
            // layout: lo[0] = a.xyz, lo[1] = (a.w, b.x, b.y), hi[0] = (b.z, b.w, c.x), hi[1] = c.yzw
            bool ol = q.lo[0] <= b.z && a.x <= q.hi[0] && q.lo[1] <= b.w && a.y <= q.hi[1] && q.lo[2] <= c.x && a.z <= q.hi[2];
            bool orr = q.lo[0] <= c.y && a.w <= q.hi[0] && q.lo[1] <= c.z && b.x <= q.hi[1] && q.lo[2] <= c.w && b.y <= q.hi[2];
            if (MODE == 2) { // only leaves after the query in Morton order (lbvh.cpp:415-424)
                ol = ol && d.z > qi;
                orr = orr && d.w > qi;
            }
            if (ol && d.x < 0) pend0 = ~d.x, tp0 = __ldg(tprim + pend0);
            if (orr && d.y < 0) pend1 = ~d.y, tp1 = __ldg(tprim + pend1);
            const bool tl = ol && d.x >= 0, tr = orr && d.y >= 0;
            if (tl) {
                node = d.x;
                // measured and rejected: an L1 prefetch of the deferred sibling (no effect); lanes refilling from a per-warp chunk of
                // queries when their walk ends (persistent walk: 2.0 -> 3.0 ms — the lanes of a warp then sit at different depths
                // of the tree and the fetches near the root, which all lanes share when they start together, stop coalescing)
                if (tr) stack[sp++] = d.y;
            } else if (tr) {
                node = d.y;
            } else if (sp > 0) {
                node = stack[--sp];
            } else {
                walking = false;
            }
        }
    }
    if (nstaged > 0) {
        __syncwarp();
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)nstaged);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int k = lane; k < nstaged; k += 32)
            if (base + k < capacity) out[base + k] = stage[warp][k];
    }
}

// 4-wide traversal: the same walk over Node4 (collapsed hierarchy): one 128-byte line per visit, up to four box tests, up to
// four leaf hits (pipelined like above), up to three deferred subtrees.
template <int MODE, int QN, int TN, bool FILTERED>
__global__ void __launch_bounds__(TRAV_BLOCK)
    k_traverse4(int q_begin, int q_end, const FBox* __restrict__ qbox, const int4* __restrict__ qprim, const Node4* __restrict__ nodes,
                int n_target, const int4* __restrict__ tprim, int2* __restrict__ out, unsigned long long* counter, unsigned long long capacity,
                int flags, FilterView filter)
{
    __shared__ int2 stage[TRAV_BLOCK / 32][STAGE_CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qi = q_begin + blockIdx.x * blockDim.x + threadIdx.x;
    const bool check_shared = flags & 1;
    const bool active = qi < q_end;
    FBox q;
    int4 qp = make_int4(-1, -1, -1, -1);
    if (active) {
        q = qbox[qi];
        qp = qprim[qi];
    }
    int stack[128];
    int sp = 0, node = 0, nstaged = 0;
    bool walking = active && n_target > 1;
    int pend[4] = { -1, -1, -1, -1 };
    int4 tp[4];
#pragma unroll
    for (int h = 0; h < 4; h++) tp[h] = make_int4(0, 0, 0, 0);
    while (__any_sync(0xffffffffu, walking || pend[0] >= 0 || pend[1] >= 0 || pend[2] >= 0 || pend[3] >= 0)) {
        float4 lx, ly, lz, hx, hy, hz;
        int4 ch = make_int4(INT_MIN, INT_MIN, INT_MIN, INT_MIN), la = make_int4(0, 0, 0, 0);
        lx = ly = lz = hx = hy = hz = make_float4(0, 0, 0, 0);
        if (walking) {
            const float4* np = reinterpret_cast<const float4*>(nodes + node);
            lx = __ldg(np), ly = __ldg(np + 1), lz = __ldg(np + 2), hx = __ldg(np + 3), hy = __ldg(np + 4), hz = __ldg(np + 5);
            ch = __ldg(reinterpret_cast<const int4*>(np + 6)), la = __ldg(reinterpret_cast<const int4*>(np + 7));
        }
        if (__any_sync(0xffffffffu, pend[0] >= 0 || pend[1] >= 0 || pend[2] >= 0 || pend[3] >= 0)) {
#pragma unroll
            for (int h = 0; h < 4; h++) {
                bool emit = false;
                int2 pr = make_int2(0, 0);
                if (pend[h] >= 0 && (!check_shared || !shares_vertex<QN, TN>(qp, tp[h])) && (!FILTERED || any_can_collide<QN, TN>(filter, qp, tp[h]))) {
                    emit = true;
                    if (MODE == 0) pr = make_int2(qp.w, tp[h].w);
                    else if (MODE == 1) pr = make_int2(tp[h].w, qp.w);
                    else pr = make_int2(min(qp.w, tp[h].w), max(qp.w, tp[h].w));
                }
                const unsigned m = __ballot_sync(0xffffffffu, emit);
                if (m) {
                    if (emit) stage[warp][nstaged + __popc(m & ((1u << lane) - 1))] = pr;
                    nstaged += __popc(m);
                }
            }
            if (nstaged > STAGE_CAP - 128) {
                __syncwarp();
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(counter, (unsigned long long)nstaged);
                base = __shfl_sync(0xffffffffu, base, 0);
                for (int k = lane; k < nstaged; k += 32)
                    if (base + k < capacity) out[base + k] = stage[warp][k];
                nstaged = 0;
                __syncwarp();
            }
        }
#pragma unroll
        for (int h = 0; h < 4; h++) pend[h] = -1;
        if (walking) {
            const float clx[4] = { lx.x, lx.y, lx.z, lx.w }, cly[4] = { ly.x, ly.y, ly.z, ly.w }, clz[4] = { lz.x, lz.y, lz.z, lz.w };
            const float chx[4] = { hx.x, hx.y, hx.z, hx.w }, chy[4] = { hy.x, hy.y, hy.z, hy.w }, chz[4] = { hz.x, hz.y, hz.z, hz.w };
            const int cc[4] = { ch.x, ch.y, ch.z, ch.w }, cl[4] = { la.x, la.y, la.z, la.w };
            int next = -1;
#pragma unroll
            for (int h = 0; h < 4; h++) {
                bool ov = cc[h] != INT_MIN && q.lo[0] <= chx[h] && clx[h] <= q.hi[0] && q.lo[1] <= chy[h] && cly[h] <= q.hi[1] && q.lo[2] <= chz[h]
                    && clz[h] <= q.hi[2];
                if (MODE == 2) ov = ov && cl[h] > qi; // only leaves after the query in Morton order
                if (ov && cc[h] < 0) pend[h] = ~cc[h], tp[h] = __ldg(tprim + pend[h]);
                if (ov && cc[h] >= 0) {
                    if (next < 0) next = cc[h];
                    else stack[sp++] = cc[h];
                }
            }
            if (next >= 0) node = next;
            else if (sp > 0) node = stack[--sp];
            else walking = false;
        }
    }
    if (nstaged > 0) {
        __syncwarp();
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)nstaged);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int k = lane; k < nstaged; k += 32)
            if (base + k < capacity) out[base + k] = stage[warp][k];
    }
}

// ---------------------------------------------------------------------------
// Sweep and prune (north_star subsystem 1; reference semantics broad_phase/sweep_and_prune.cpp:106-119, GPU route
// sweep_and_tiniest_queue.cu:121-212): the boxes are sorted by their lower bound along ONE axis (the longest scene
// extent); every query scans the sorted targets forward while their lower bound does not pass its upper bound and
// tests the two other axes.  No hierarchy, no stack: sequential coalesced box reads — the alternative to the LBVH for
// scenes whose swept boxes overlap so much that a tree prunes little (BASELINE config 4).  The predicate (closed
// float-box overlap, no shared vertex, collision filter) and therefore the candidate SET are those of the LBVH.
__device__ __forceinline__ float axis_lo(const FBox& b, int ax) { return ax == 0 ? b.lo[0] : (ax == 1 ? b.lo[1] : b.lo[2]); }
__device__ __forceinline__ float axis_hi(const FBox& b, int ax) { return ax == 0 ? b.hi[0] : (ax == 1 ? b.hi[1] : b.hi[2]); }
// monotone float -> unsigned key
__global__ void k_axis_key(int n, int axis, const FBox* __restrict__ box, unsigned long long* __restrict__ key, int* __restrict__ ord)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned u = __float_as_uint(axis_lo(box[i], axis));
    key[i] = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    ord[i] = i;
}
// start: 0 = self (targets after the query), 1 = first target with lo >= query lo, 2 = first target with lo > query lo
template <int MODE, int QN, int TN, bool FILTERED>
__global__ void __launch_bounds__(TRAV_BLOCK)
    k_sap_sweep(int q_begin, int q_end, const FBox* __restrict__ qbox, const int4* __restrict__ qprim, int n_target,
                const FBox* __restrict__ tbox, const int4* __restrict__ tprim, int axis, int start_rule, int2* __restrict__ out,
                unsigned long long* counter, unsigned long long capacity, int flags, FilterView filter)
{
    __shared__ int2 stage[TRAV_BLOCK / 32][STAGE_CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qi = q_begin + blockIdx.x * blockDim.x + threadIdx.x;
    const bool check_shared = flags & 1;
    bool active = qi < q_end;
    FBox q;
    int4 qp = make_int4(-1, -1, -1, -1);
    int j = n_target;
    float qhi = -INFINITY;
    if (active) {
        q = qbox[qi];
        qp = qprim[qi];
        qhi = axis_hi(q, axis);
        if (start_rule == 0) {
            j = qi + 1;
        } else { // binary search over the sorted lower bounds
            const float qlo = axis_lo(q, axis);
            int lo = 0, hi = n_target;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const float v = axis_lo(tbox[mid], axis);
                if (start_rule == 1 ? v < qlo : v <= qlo) lo = mid + 1;
                else hi = mid;
            }
            j = lo;
        }
    }
    const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
    int nstaged = 0;
    while (__any_sync(0xffffffffu, active && j < n_target)) {
        bool emit = false;
        int2 pr = make_int2(0, 0);
        if (active && j < n_target) {
            const FBox b = tbox[j];
            if (axis_lo(b, axis) > qhi) {
                active = false; // sorted: no later target can overlap along the sweep axis
            } else if (axis_lo(q, a1) <= axis_hi(b, a1) && axis_lo(b, a1) <= axis_hi(q, a1) && axis_lo(q, a2) <= axis_hi(b, a2)
                       && axis_lo(b, a2) <= axis_hi(q, a2)) {
                const int4 tp = __ldg(tprim + j);
                if ((!check_shared || !shares_vertex<QN, TN>(qp, tp)) && (!FILTERED || any_can_collide<QN, TN>(filter, qp, tp))) {
                    emit = true;
                    if (MODE == 0) pr = make_int2(qp.w, tp.w);
                    else if (MODE == 1) pr = make_int2(tp.w, qp.w);
                    else pr = make_int2(min(qp.w, tp.w), max(qp.w, tp.w));
                }
            }
            j++;
        }
        const unsigned m = __ballot_sync(0xffffffffu, emit);
        if (m) {
            if (emit) stage[warp][nstaged + __popc(m & ((1u << lane) - 1))] = pr;
            nstaged += __popc(m);
            if (nstaged > STAGE_CAP - 32) {
                __syncwarp();
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(counter, (unsigned long long)nstaged);
                base = __shfl_sync(0xffffffffu, base, 0);
                for (int k = lane; k < nstaged; k += 32)
                    if (base + k < capacity) out[base + k] = stage[warp][k];
                nstaged = 0;
                __syncwarp();
            }
        }
    }
    if (nstaged > 0) {
        __syncwarp();
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)nstaged);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int k = lane; k < nstaged; k += 32)
            if (base + k < capacity) out[base + k] = stage[warp][k];
    }
}

// Largest candidate list (pairs of one kind) the library materialises: beyond it the collision-free step size streams
// the query leaves in chunks (ccd.cu: ccd_stepsize_streaming) and everything else fails loudly.  Default 2^29 pairs
// (4 GiB per list); IPCB_MAX_PAIRS overrides it (the tests use tiny values to force the chunked path).
size_t max_pairs()
{
    const char* e = getenv("IPCB_MAX_PAIRS"); // read per call: the tests switch it inside one process
    return e ? std::max<size_t>(64, strtoull(e, nullptr, 10)) : (size_t(1) << 29);
}

// one detection: queries (sorted view) against a tree.  launch() enqueues the kernel and the read-back of the
// pair count on the job's stream; finish() waits for it and, if the output buffer was too small, grows it and
// repeats the pass.  Two jobs on different streams / counter slots run concurrently.
struct TraverseJob {
    ipcb_ctx* ctx;
    const Tree* q;
    const Tree* t;
    int mode, qn, tn;
    bool check_shared;
    PairList* out;
    cudaStream_t s;
    int slot; // device counter / pinned slot (0 or 1)
    bool sap = false; // q / t are axis-sorted views and the pass is a sweep (two sweeps for two different sets)
    int q_begin = 0, q_end = 0;
    bool live = false;
    unsigned long long overflow = 0; // pairs found by a pass that exceeded max_pairs() (the list is then invalid)

    // the rank's Morton range of query leaves; false when there is nothing to traverse
    bool full_range(bool shard)
    {
        if (q->n == 0 || t->n == 0 || (mode == 2 && t->n < 2)) return false;
        q_begin = 0, q_end = q->n;
        if (shard && ctx->shard_world > 1) {
            q_begin = int((int64_t(q->n) * ctx->shard_rank) / ctx->shard_world);
            q_end = int((int64_t(q->n) * (ctx->shard_rank + 1)) / ctx->shard_world);
        }
        return q_end - q_begin > 0;
    }
    void launch(bool shard)
    {
        out->count = 0;
        out->sorted = false;
        live = false;
        overflow = 0;
        if (!full_range(shard)) return;
        launch_range(q_begin, q_end);
    }
    // queries [qb, qe) only (chunked emission: SURVEY §7 hard part 7)
    void launch_range(int qb, int qe)
    {
        out->count = 0;
        out->sorted = false;
        overflow = 0;
        q_begin = qb, q_end = qe;
        if (out->pairs.cap == 0) out->pairs.reserve(std::min(size_t(q_end - q_begin) * 8 + 1024, max_pairs()));
        live = true;
        enqueue();
    }
    void enqueue()
    {
        unsigned long long* counter = ctx->dCounters.p + 16 + slot;
        IPCB_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s));
        const unsigned long long cap = out->pairs.cap;
        const unsigned grid = grid_for(q_end - q_begin, TRAV_BLOCK);
        Stage kt(ctx, mode == 2 && qn == 2 ? "k:k_traverse<EE>" : (mode == 1 && tn == 3 ? "k:k_traverse<FV>" : "k:k_traverse<other>"), s);
        const int flags = (check_shared ? 1 : 0) | (ctx->filter_on() ? 2 : 0);
        const FilterView filter { ctx->filter_patches ? ctx->dPatch.p : nullptr, ctx->filter_n_dynamic };
        if (sap) {
            enqueue_sap(counter, cap, flags, filter);
            IPCB_CUDA(cudaMemcpyAsync(ctx->pinned.p + 16 + slot, counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
            return;
        }
#define IPCB_TRAVERSE(M, QN, TN)                                                                                                         \
    if (t->has_nodes4 && t->n > 2) {                                                                                                    \
        if (flags & 2)                                                                                                                  \
            k_traverse4<M, QN, TN, true><<<grid, TRAV_BLOCK, 0, s>>>(q_begin, q_end, q->sbox.p, q->sprim.p, t->nodes4.p, t->n, t->sprim.p, \
                                                                     out->pairs.p, counter, cap, flags, filter);                       \
        else                                                                                                                            \
            k_traverse4<M, QN, TN, false><<<grid, TRAV_BLOCK, 0, s>>>(q_begin, q_end, q->sbox.p, q->sprim.p, t->nodes4.p, t->n, t->sprim.p, \
                                                                      out->pairs.p, counter, cap, flags, filter);                      \
    } else if (flags & 2)                                                                                                                      \
        k_traverse<M, QN, TN, true><<<grid, TRAV_BLOCK, 0, s>>>(q_begin, q_end, q->sbox.p, q->sprim.p, t->nodes.p, t->n, t->sbox.p, t->sprim.p, \
                                                                out->pairs.p, counter, cap, flags, filter);                             \
    else                                                                                                                                \
        k_traverse<M, QN, TN, false><<<grid, TRAV_BLOCK, 0, s>>>(q_begin, q_end, q->sbox.p, q->sprim.p, t->nodes.p, t->n, t->sbox.p, t->sprim.p, \
                                                                 out->pairs.p, counter, cap, flags, filter)
        switch (mode * 100 + qn * 10 + tn) {
        case 211: IPCB_TRAVERSE(2, 1, 1); break; // vertex - vertex
        case 21: IPCB_TRAVERSE(0, 2, 1); break;  // edges walk the vertex tree
        case 222: IPCB_TRAVERSE(2, 2, 2); break; // edge - edge
        case 113: IPCB_TRAVERSE(1, 1, 3); break; // vertices walk the face tree
        case 132: IPCB_TRAVERSE(1, 3, 2); break; // faces walk the edge tree
        case 233: IPCB_TRAVERSE(2, 3, 3); break; // face - face
        default: throw Error("broad phase: unsupported traversal");
        }
#undef IPCB_TRAVERSE
        ctx->launches++;
        IPCB_CUDA(cudaGetLastError());
        IPCB_CUDA(cudaMemcpyAsync(ctx->pinned.p + 16 + slot, counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    }
    // sweep: self kinds one pass; two sets: the queries scan the targets that start inside them, then the targets scan
    // the queries that start strictly inside them (every overlapping pair is found exactly once)
    void enqueue_sap(unsigned long long* counter, unsigned long long cap, int flags, const FilterView& filter)
    {
        Stage kt(ctx, "k:k_sap_sweep", s);
        const int axis = ctx->sap_axis;
#define IPCB_SAP(M, QN, TN, QB, QE, QT, TT, RULE)                                                                                       \
    if (flags & 2)                                                                                                                      \
        k_sap_sweep<M, QN, TN, true><<<grid_for((QE) - (QB), TRAV_BLOCK), TRAV_BLOCK, 0, s>>>(QB, QE, (QT)->sbox.p, (QT)->sprim.p, (TT)->n, (TT)->sbox.p, \
                                                                                             (TT)->sprim.p, axis, RULE, out->pairs.p, counter, cap, flags, filter); \
    else                                                                                                                                \
        k_sap_sweep<M, QN, TN, false><<<grid_for((QE) - (QB), TRAV_BLOCK), TRAV_BLOCK, 0, s>>>(QB, QE, (QT)->sbox.p, (QT)->sprim.p, (TT)->n, (TT)->sbox.p, \
                                                                                              (TT)->sprim.p, axis, RULE, out->pairs.p, counter, cap, flags, filter)
        // second pass of a two-set kind: the rank's share of the TARGET set scans the query set
        int tb = 0, te = t->n;
        if (ctx->shard_world > 1 && (q_begin != 0 || q_end != q->n)) {
            tb = int((int64_t(t->n) * ctx->shard_rank) / ctx->shard_world);
            te = int((int64_t(t->n) * (ctx->shard_rank + 1)) / ctx->shard_world);
        }
        switch (mode * 100 + qn * 10 + tn) {
        case 211: IPCB_SAP(2, 1, 1, q_begin, q_end, q, t, 0); break;
        case 222: IPCB_SAP(2, 2, 2, q_begin, q_end, q, t, 0); break;
        case 233: IPCB_SAP(2, 3, 3, q_begin, q_end, q, t, 0); break;
        case 21: // edges x vertices, emitted (edge, vertex)
            IPCB_SAP(0, 2, 1, q_begin, q_end, q, t, 1);
            if (te > tb) { IPCB_SAP(1, 1, 2, tb, te, t, q, 2); }
            break;
        case 113: // vertices x faces, emitted (face, vertex)
            IPCB_SAP(1, 1, 3, q_begin, q_end, q, t, 1);
            if (te > tb) { IPCB_SAP(0, 3, 1, tb, te, t, q, 2); }
            break;
        case 132: // faces x edges, emitted (edge, face)
            IPCB_SAP(1, 3, 2, q_begin, q_end, q, t, 1);
            if (te > tb) { IPCB_SAP(0, 2, 3, tb, te, t, q, 2); }
            break;
        default: throw Error("broad phase: unsupported sweep");
        }
#undef IPCB_SAP
        ctx->launches += 2;
        IPCB_CUDA(cudaGetLastError());
    }
    void finish()
    {
        if (!live) return;
        for (int attempt = 0; attempt < 3; attempt++) {
            IPCB_CUDA(cudaStreamSynchronize(s));
            const unsigned long long found = (unsigned long long)ctx->pinned.p[16 + slot];
            if (found <= out->pairs.cap) {
                out->count = int64_t(found);
                return;
            }
            if (found > max_pairs()) { // more pairs than one list may hold: the caller chunks the queries (or gives up)
                overflow = found;
                out->count = 0;
                return;
            }
            out->pairs.reserve(std::min(size_t(found) + size_t(found) / 8, max_pairs())); // overflow: grow and repeat the pass
            enqueue();
        }
        throw Error("broad phase: candidate buffer overflow persisted");
    }
};

static void too_many_pairs(unsigned long long found)
{
    throw Error("broad phase: " + std::to_string(found) + " candidate pairs of one kind exceed the list limit of " + std::to_string(max_pairs())
                + " (IPCB_MAX_PAIRS); only compute_collision_free_stepsize streams larger sets");
}
static void run_traverse(ipcb_ctx* ctx, const Tree& q, const Tree& t, int mode, int qn, int tn, bool check_shared, PairList& out, bool shard,
                         bool sap = false)
{
    TraverseJob job { ctx, &q, &t, mode, qn, tn, check_shared, &out, ctx->stream, 0 };
    job.sap = sap;
    job.launch(shard);
    job.finish();
    if (job.overflow) too_many_pairs(job.overflow);
}

// chunked emission for the streaming step size: the queries [qb, qe) of the rank's range of one main kind (edge-edge or
// face-vertex) of the BUILT trees into ctx->cand[kind].  Returns the pairs found; a value above max_pairs() means the
// chunk was too large (the list is invalid).  range: out-parameter with the rank's whole query range when qb < 0.
unsigned long long traverse_chunk(ipcb_ctx* ctx, int kind, int qb, int qe, int* range)
{
    const bool ee = kind == IPCB_EE;
    // the chunks walk the LBVH (also when the full pass was a sweep: its two-pass scheme does not split into query chunks)
    if (ee && !ctx->etree_ok) build_tree(ctx, ctx->eset, ctx->etree, true), ctx->etree_ok = true;
    if (!ee && !ctx->ftree_ok) build_tree(ctx, ctx->fset, ctx->ftree, true), ctx->ftree_ok = true;
    if (!ee && !ctx->vorder_valid) build_tree(ctx, ctx->vset, ctx->vtree, false), ctx->vorder_valid = true;
    TraverseJob job { ctx, ee ? &ctx->etree : &ctx->vtree, ee ? &ctx->etree : &ctx->ftree, ee ? 2 : 1, ee ? 2 : 1, ee ? 2 : 3, true, &ctx->cand[kind],
                      ctx->stream, 0 };
    if (qb < 0) {
        const bool any = job.full_range(true);
        range[0] = any ? job.q_begin : 0, range[1] = any ? job.q_end : 0;
        return 0;
    }
    job.launch_range(qb, qe);
    job.finish();
    return job.overflow ? job.overflow : (unsigned long long)ctx->cand[kind].count;
}

static Tree& ensure_tree(ipcb_ctx* ctx, int which)
{
    if (which == 0) {
        if (!ctx->vtree_ok) build_tree(ctx, ctx->vset, ctx->vtree, true), ctx->vtree_ok = true, ctx->vorder_valid = true;
        return ctx->vtree;
    }
    if (which == 1) {
        if (!ctx->etree_ok) build_tree(ctx, ctx->eset, ctx->etree, true), ctx->etree_ok = true;
        return ctx->etree;
    }
    if (!ctx->ftree_ok) build_tree(ctx, ctx->fset, ctx->ftree, true), ctx->ftree_ok = true;
    return ctx->ftree;
}

// broad_phase.hpp:71-97 — which BVH is walked by which leaves follows lbvh.cpp:692-797
void broad_detect(ipcb_ctx* ctx, int kind, PairList& out)
{
    if (!ctx->built) throw Error("broad phase not built");
    Stage st(ctx, "broad_detect");
    if (ctx->broad_method == IPCB_BROAD_SAP) { // same six kinds, same orientation of the pairs, by sweeps
        static const int spec[6][5] = { { 0, 0, 2, 1, 1 }, { 1, 0, 0, 2, 1 }, { 1, 1, 2, 2, 2 }, { 0, 2, 1, 1, 3 }, { 2, 1, 1, 3, 2 }, { 2, 2, 2, 3, 3 } };
        if (kind < 0 || kind > 5) throw Error("bad candidate kind");
        Tree& q = ensure_sap_view(ctx, spec[kind][0]);
        Tree& t = ensure_sap_view(ctx, spec[kind][1]);
        run_traverse(ctx, q, t, spec[kind][2], spec[kind][3], spec[kind][4], true, out, true, true);
        return;
    }
    switch (kind) {
    case IPCB_VV: {
        Tree& v = ensure_tree(ctx, 0);
        run_traverse(ctx, v, v, 2, 1, 1, true, out, true);
        break;
    }
    case IPCB_EV: { // edges walk the vertex BVH
        Tree& e = ensure_tree(ctx, 1);
        Tree& v = ensure_tree(ctx, 0);
        run_traverse(ctx, e, v, 0, 2, 1, true, out, true);
        break;
    }
    case IPCB_EE: {
        Tree& e = ensure_tree(ctx, 1);
        run_traverse(ctx, e, e, 2, 2, 2, true, out, true);
        break;
    }
    case IPCB_FV: { // vertices walk the face BVH, emitted as (face, vertex)
        Tree& v = ensure_tree(ctx, 0);
        Tree& f = ensure_tree(ctx, 2);
        run_traverse(ctx, v, f, 1, 1, 3, true, out, true);
        break;
    }
    case IPCB_EF: { // faces walk the edge BVH, emitted as (edge, face)
        Tree& f = ensure_tree(ctx, 2);
        Tree& e = ensure_tree(ctx, 1);
        run_traverse(ctx, f, e, 1, 3, 2, true, out, true);
        break;
    }
    case IPCB_FF: {
        Tree& f = ensure_tree(ctx, 2);
        run_traverse(ctx, f, f, 2, 3, 3, true, out, true);
        break;
    }
    default: throw Error("bad candidate kind");
    }
}

// ---------------------------------------------------------------------------
// ipc::has_intersections (ipc.cpp:105-166), 3D: edge-face candidates of a broad phase inflated by 1e-6 of the scene
// diagonal, then is_edge_intersecting_triangle (geometry/intersection.cpp:115-145): both end points strictly on one side
// of the triangle's plane (exact orient3d) -> no; otherwise solve [t1 - t0, t2 - t0, e0 - e1] (u, v, t) = e0 - t0 with a
// full-pivoting LU and test 0 <= u, v, u + v <= 1, 0 <= t <= 1.
// The orientation is evaluated in FP64 with Shewchuk's static error bound; undecided candidates (status 2) get the
// exact predicate on the host (exact_orient3d.hpp).
__host__ __device__ inline bool edge_triangle_solve(d3 e0, d3 e1, d3 t0, d3 t1, d3 t2)
{
    // columns of M; full-pivoting Gaussian elimination like Eigen::FullPivLU (largest |entry| of the remaining block)
    double M[3][3] = { { t1.x - t0.x, t2.x - t0.x, e0.x - e1.x }, { t1.y - t0.y, t2.y - t0.y, e0.y - e1.y }, { t1.z - t0.z, t2.z - t0.z, e0.z - e1.z } };
    double b[3] = { e0.x - t0.x, e0.y - t0.y, e0.z - t0.z };
    int colperm[3] = { 0, 1, 2 };
    int rank = 3;
    for (int k = 0; k < 3; k++) {
        int pr = k, pc = k;
        double best = 0;
        for (int i = k; i < 3; i++)
            for (int j = k; j < 3; j++)
                if (fabs(M[i][j]) > best) best = fabs(M[i][j]), pr = i, pc = j;
        if (best == 0) {
            rank = k;
            break;
        }
        if (pr != k) {
            for (int j = 0; j < 3; j++) {
                const double t = M[k][j];
                M[k][j] = M[pr][j], M[pr][j] = t;
            }
            const double t = b[k];
            b[k] = b[pr], b[pr] = t;
        }
        if (pc != k) {
            for (int i = 0; i < 3; i++) {
                const double t = M[i][k];
                M[i][k] = M[i][pc], M[i][pc] = t;
            }
            const int t = colperm[k];
            colperm[k] = colperm[pc], colperm[pc] = t;
        }
        for (int i = k + 1; i < 3; i++) {
            const double l = M[i][k] / M[k][k];
            for (int j = k + 1; j < 3; j++) M[i][j] -= l * M[k][j];
            b[i] -= l * b[k];
        }
    }
    double y[3] = { 0, 0, 0 };
    for (int k = rank - 1; k >= 0; k--) {
        double acc = b[k];
        for (int j = k + 1; j < rank; j++) acc -= M[k][j] * y[j];
        y[k] = acc / M[k][k];
    }
    double uvt[3] = { 0, 0, 0 };
    for (int k = 0; k < 3; k++) uvt[colperm[k]] = y[k];
    return uvt[0] >= 0.0 && uvt[1] >= 0.0 && uvt[0] + uvt[1] <= 1.0 && uvt[2] >= 0.0 && uvt[2] <= 1.0;
}
// FP64 orientation of d against the plane (a, b, c) with Shewchuk's static filter: +-1 certain, 0 undecided
__device__ inline int orient3d_filtered(d3 a, d3 b, d3 c, d3 d)
{
    const double adx = a.x - d.x, bdx = b.x - d.x, cdx = c.x - d.x, ady = a.y - d.y, bdy = b.y - d.y, cdy = c.y - d.y;
    const double adz = a.z - d.z, bdz = b.z - d.z, cdz = c.z - d.z;
    const double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy, cdxady = cdx * ady, adxcdy = adx * cdy, adxbdy = adx * bdy, bdxady = bdx * ady;
    const double det = adz * (bdxcdy - cdxbdy) + bdz * (cdxady - adxcdy) + cdz * (adxbdy - bdxady);
    const double permanent = (fabs(bdxcdy) + fabs(cdxbdy)) * fabs(adz) + (fabs(cdxady) + fabs(adxcdy)) * fabs(bdz) + (fabs(adxbdy) + fabs(bdxady)) * fabs(cdz);
    const double errbound = 7.771561172376103e-16 * permanent; // (7 + 56 eps) eps, eps = 2^-53
    if (det > errbound) return 1;
    if (-det > errbound) return -1;
    return 0;
}
__global__ void k_edge_face_intersect(int64_t n, const int2* __restrict__ pairs, const int2* __restrict__ E, const int4* __restrict__ F,
                                      const double4* __restrict__ X, unsigned long long* counters, int* __restrict__ undecided)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int2 c = pairs[i]; // (edge, face)
    const int2 e = __ldg(E + c.x);
    const int4 f = __ldg(F + c.y);
    const d3 e0 = load_vertex(X, e.x), e1 = load_vertex(X, e.y), t0 = load_vertex(X, f.x), t1 = load_vertex(X, f.y), t2 = load_vertex(X, f.z);
    const int o1 = orient3d_filtered(t0, t1, t2, e0), o2 = orient3d_filtered(t0, t1, t2, e1);
    if (o1 != 0 && o2 != 0) {
        if (o1 == o2) return; // strictly on one side of the plane
        if (edge_triangle_solve(e0, e1, t0, t1, t2)) atomicAdd(counters, 1ull);
        return;
    }
    undecided[atomicAdd(counters + 1, 1ull)] = int(i); // the exact predicate decides on the host
}

bool has_intersections(ipcb_ctx* ctx, double inflation_radius)
{
    broad_build(ctx, false, inflation_radius);
    PairList& pl = ctx->detected[IPCB_EF];
    broad_detect(ctx, IPCB_EF, pl);
    const int64_t n = pl.count;
    if (n == 0) return false;
    cudaStream_t s = ctx->stream;
    unsigned long long* cnt = ctx->dCounters.p + 20;
    IPCB_CUDA(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned long long), s));
    ctx->hsel.reserve(size_t(n));
    k_edge_face_intersect<<<grid_for(n, 256), 256, 0, s>>>(n, pl.pairs.p, ctx->dE.p, ctx->dF.p, ctx->X0.p, cnt, ctx->hsel.p);
    ctx->launches++;
    IPCB_CUDA(cudaGetLastError());
    unsigned long long h[2];
    IPCB_CUDA(cudaMemcpyAsync(h, cnt, sizeof h, cudaMemcpyDeviceToHost, s));
    IPCB_CUDA(cudaStreamSynchronize(s));
    if (h[0] > 0) return true;
    if (h[1] == 0) return false;
    // undecided orientations: exact predicate on the host
    std::vector<int> idx(h[1]);
    IPCB_CUDA(cudaMemcpyAsync(idx.data(), ctx->hsel.p, sizeof(int) * h[1], cudaMemcpyDeviceToHost, s));
    std::vector<double4> X(size_t(ctx->nV));
    IPCB_CUDA(cudaMemcpyAsync(X.data(), ctx->X0.p, sizeof(double4) * size_t(ctx->nV), cudaMemcpyDeviceToHost, s));
    std::vector<int2> pairs(static_cast<size_t>(n));
    IPCB_CUDA(cudaMemcpyAsync(pairs.data(), pl.pairs.p, sizeof(int2) * size_t(n), cudaMemcpyDeviceToHost, s));
    IPCB_CUDA(cudaStreamSynchronize(s));
    auto P = [&](int v) { return mk3(X[size_t(v)].x, X[size_t(v)].y, X[size_t(v)].z); };
    for (int i : idx) {
        const int2 c = pairs[size_t(i)];
        const int ev[2] = { ctx->hE[2 * size_t(c.x)], ctx->hE[2 * size_t(c.x) + 1] };
        const int fv[3] = { ctx->hF[3 * size_t(c.y)], ctx->hF[3 * size_t(c.y) + 1], ctx->hF[3 * size_t(c.y) + 2] };
        const d3 e0 = P(ev[0]), e1 = P(ev[1]), t0 = P(fv[0]), t1 = P(fv[1]), t2 = P(fv[2]);
        const int o1 = ipcb_exact::orient3d_sign(&t0.x, &t1.x, &t2.x, &e0.x), o2 = ipcb_exact::orient3d_sign(&t0.x, &t1.x, &t2.x, &e1.x);
        if (o1 != 0 && o2 != 0 && o1 == o2) continue;
        if (edge_triangle_solve(e0, e1, t0, t1, t2)) return true;
    }
    return false;
}

// sort pairs lexicographically (canonical order for fetch / parity checks)
__global__ void k_pairs_to_keys(int64_t n, const int2* __restrict__ p, unsigned long long* __restrict__ k)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) k[i] = ((unsigned long long)(unsigned)p[i].x << 32) | (unsigned)p[i].y;
}
__global__ void k_keys_to_pairs(int64_t n, const unsigned long long* __restrict__ k, int2* __restrict__ p)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = make_int2(int(k[i] >> 32), int(k[i] & 0xffffffffu));
}
void sort_pairs(ipcb_ctx* ctx, PairList& pl)
{
    if (pl.sorted || pl.count < 2) {
        pl.sorted = true;
        return;
    }
    const int64_t n = pl.count;
    ctx->hkey.reserve(n), ctx->hkey_sorted.reserve(n);
    cudaStream_t s = ctx->stream;
    k_pairs_to_keys<<<grid_for(n, 256), 256, 0, s>>>(n, pl.pairs.p, ctx->hkey.p);
    size_t bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, ctx->hkey.p, ctx->hkey_sorted.p, n, 0, 64, s);
    ctx->cubtmp.reserve(bytes);
    cub::DeviceRadixSort::SortKeys(ctx->cubtmp.p, bytes, ctx->hkey.p, ctx->hkey_sorted.p, n, 0, 64, s);
    k_keys_to_pairs<<<grid_for(n, 256), 256, 0, s>>>(n, ctx->hkey_sorted.p, pl.pairs.p);
    ctx->launches += 11;
    pl.sorted = true;
}

// Candidates::build, 3D (candidates.cpp:43-222): EE + FV on the whole mesh,
// VV between codim vertices, EV between codim edges and codim vertices
// allow_overflow: a main kind with more than max_pairs() candidates does not throw; the function returns true and leaves
// that kind empty (the trees stay built: the streaming step size traverses them in chunks)
bool candidates_build(ipcb_ctx* ctx, bool swept, double r, bool allow_overflow)
{
    bool overflowed = false;
    broad_build(ctx, swept, r);
    for (auto& c : ctx->cand) c.count = 0, c.sorted = false;
    {
        // The edge tree + edge-edge traversal and the face tree + face-vertex traversal are independent: they run on
        // two auxiliary streams (the small sort / hierarchy kernels of one chain fill the tails of the other); the
        // vertex queries only need the Morton order, not a hierarchy (main stream).
        Stage st(ctx, ctx->broad_method == IPCB_BROAD_SAP ? "sap_sort+sweep" : "lbvh_build+traverse");
        const bool do_ee = ctx->nE >= 2, do_fv = ctx->nF && ctx->nV;
        if (ctx->broad_method == IPCB_BROAD_SAP) {
            TraverseJob ee { ctx, nullptr, nullptr, 2, 2, 2, true, &ctx->cand[IPCB_EE], ctx->stream, 0 };
            TraverseJob fv { ctx, nullptr, nullptr, 1, 1, 3, true, &ctx->cand[IPCB_FV], ctx->stream, 1 };
            ee.sap = fv.sap = true;
            if (do_ee) ee.q = ee.t = &ensure_sap_view(ctx, 1), ee.launch(true);
            if (do_fv) fv.q = &ensure_sap_view(ctx, 0), fv.t = &ensure_sap_view(ctx, 2), fv.launch(true);
            ee.finish();
            fv.finish();
            ctx->cand_overflow[IPCB_EE] = ee.overflow, ctx->cand_overflow[IPCB_FV] = fv.overflow;
            if (ee.overflow || fv.overflow) {
                if (!allow_overflow) too_many_pairs(std::max(ee.overflow, fv.overflow));
                overflowed = true;
            }
        } else {
        ctx->fork();
        TraverseJob ee { ctx, &ctx->etree, &ctx->etree, 2, 2, 2, true, &ctx->cand[IPCB_EE], ctx->aux[0], 0 };
        TraverseJob fv { ctx, &ctx->vtree, &ctx->ftree, 1, 1, 3, true, &ctx->cand[IPCB_FV], ctx->aux[1], 1 };
        if (do_ee) {
            if (!ctx->etree_ok) build_tree(ctx, ctx->eset, ctx->etree, true, ctx->aux[0]), ctx->etree_ok = true;
            ee.launch(true);
        }
        if (do_fv) {
            if (!ctx->ftree_ok) build_tree(ctx, ctx->fset, ctx->ftree, true, ctx->aux[1]), ctx->ftree_ok = true;
            if (!ctx->vtree_ok) build_tree(ctx, ctx->vset, ctx->vtree, false, ctx->aux[2]), ctx->vorder_valid = true;
            ctx->join(2); // not needed by the main stream itself, but keeps every later main-stream consumer ordered
            IPCB_CUDA(cudaEventRecord(ctx->ev_fork, ctx->aux[2]));
            IPCB_CUDA(cudaStreamWaitEvent(ctx->aux[1], ctx->ev_fork, 0)); // the vertex order feeds the face-vertex traversal
            fv.launch(true);
        }
        ee.finish();
        fv.finish();
        ctx->join(0);
        ctx->join(1);
        ctx->cand_overflow[IPCB_EE] = ee.overflow, ctx->cand_overflow[IPCB_FV] = fv.overflow;
        if (ee.overflow || fv.overflow) {
            if (!allow_overflow) too_many_pairs(std::max(ee.overflow, fv.overflow));
            overflowed = true;
        }
        }
    }
    const int ncv = int(ctx->codimV.size()), nce = int(ctx->codimE.size());
    if (ncv) {
        Stage st(ctx, "codim");
        cudaStream_t s = ctx->stream;
        ctx->cvset.n = ncv;
        ctx->cvset.box.reserve(ncv), ctx->cvset.prim.reserve(ncv);
        const bool filtered = ctx->filter_on();
        k_gather_set<<<grid_for(ncv, 256), 256, 0, s>>>(ncv, ctx->dCodimV.p, ctx->vset.box.p, ctx->vset.prim.p, ctx->cvset.box.p,
                                                      ctx->cvset.prim.p, filtered ? 1 : 0, nullptr);
        ctx->launches++;
        build_tree(ctx, ctx->cvset, ctx->cvtree, true);
        if (ncv >= 2 && ctx->shard_rank == 0) // tiny sets are not sharded: rank 0 owns them
            run_traverse(ctx, ctx->cvtree, ctx->cvtree, 2, 1, 1, false, ctx->cand[IPCB_VV], false);
        if (nce) {
            ctx->ceset.n = nce;
            ctx->ceset.box.reserve(nce), ctx->ceset.prim.reserve(nce);
            k_gather_set<<<grid_for(nce, 256), 256, 0, s>>>(nce, ctx->dCodimE.p, ctx->eset.box.p, ctx->eset.prim.p, ctx->ceset.box.p,
                                                          ctx->ceset.prim.p, filtered ? 2 : 0, ctx->dCodimELocal.p);
            ctx->launches++;
            build_tree(ctx, ctx->ceset, ctx->cetree, false);
            if (ctx->shard_rank == 0) run_traverse(ctx, ctx->cetree, ctx->cvtree, 0, 2, 1, false, ctx->cand[IPCB_EV], false);
        }
    }
    return overflowed;
}

} // namespace ipcb
