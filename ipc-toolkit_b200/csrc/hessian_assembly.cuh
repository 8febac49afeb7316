// hessian_assembly.cuh — what the barrier potential (potential.cu) and the friction potential (friction.cu) share:
// barrier scalars, collision / mesh views, the per-collision Hessian record format (upper-triangular 3x3 vertex blocks +
// exact-non-zero masks + (vertex, collision) incidences) and the assembly of those records into compressed columns.
#pragma once
#include "ctx.cuh"
#include "geom.cuh"

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
inline BarrierDev make_barrier(const ipcb_barrier_params& bp, double dmin)
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

// The local matrix is symmetric, so only its UPPER-TRIANGULAR vertex blocks are stored: slot (a, b), a <= b, at
// tri_slot(a, b) — 3 / 6 / 10 blocks of 72 bytes per VV / EV / 4-point collision instead of 4 / 9 / 16.  The column of
// point a reads its row block b < a as the TRANSPOSE of slot (b, a) (k_hess_numeric).  The 9-bit exact-non-zero masks
// are tiny and stay full (16 per collision, the mirrored ones transposed), so the symbolic pass is layout-agnostic.
struct HessOut {
    int4* vid;               // stencil vertex ids, -1 padded
    unsigned short* mask;    // 16 per collision (global collision index): 9-bit exact-non-zero mask per slot a * 4 + b
    double* blk;             // blocks of THIS kind: tri_count(NP) x 9 doubles per record (record index within the kind)
    unsigned long long* inc; // incidences: (vertex << 32) | (gi * 4 + a), NP per collision
    int v_lo, v_hi;          // owned vertex range (row block of a sharded Hessian); incidences of other
    int v_none;              // vertices get the vertex key v_none (= nV: sorted behind every column)
    int records_done;        // the ids / incidences were written by k_write_records already: write_record only reports `own`
    unsigned long long* cnt; // optional: per vertex key, the number of its incidences by record size (3 x 21 bits: 2- / 3- / 4-point)
};
constexpr int HSLOTS = 16;
__host__ __device__ constexpr int tri_count(int np) { return np * (np + 1) / 2; }
__host__ __device__ constexpr int tri_slot(int np, int a, int b) { return a * np - a * (a - 1) / 2 + (b - a); } // a <= b
// 9-bit mask of the transposed 3x3 block: bit 3r + c -> bit 3c + r
__device__ __forceinline__ unsigned mask_transpose(unsigned m)
{
    return (m & 0x111u) | ((m & 0x022u) << 2) | ((m & 0x088u) >> 2) | ((m & 0x004u) << 4) | ((m & 0x040u) >> 4);
}

// returns the mask of the stencil points whose vertex this rank owns
template <int NP> __device__ inline unsigned write_record(const HessOut& out, int64_t gi, int64_t inc_base, const int* vid)
{
    if (!out.records_done) out.vid[gi] = make_int4(vid[0], vid[1], NP > 2 ? vid[2] : -1, NP > 3 ? vid[3] : -1);
    unsigned own = 0;
#pragma unroll
    for (int a = 0; a < NP; a++) {
        const bool mine = vid[a] >= out.v_lo && vid[a] < out.v_hi;
        own |= unsigned(mine) << a;
        if (!out.records_done) {
            const unsigned vkey = unsigned(mine ? vid[a] : out.v_none);
            out.inc[inc_base + a] = ((unsigned long long)vkey << 32) | (unsigned long long)(gi * 4 + a);
            // counting placement of the incidences (hessian_assemble); parked incidences are never read: not counted (they
            // would all hit one address)
            if (out.cnt && mine) atomicAdd(out.cnt + vkey, 1ull << (21 * (NP - 2)));
        }
    }
    return own;
}

inline CollView view(const ipcb_ctx* ctx, int k)
{
    const CollisionSet& cs = ctx->coll[k];
    return { k, cs.count, cs.ids.p, cs.w.p, cs.eps.p, cs.dtype.p };
}
// energy / gradient on a rank of a sharded potential: the slice [rank*n/world, (rank+1)*n/world) of the kind
inline CollView view_slice(const ipcb_ctx* ctx, int k)
{
    CollView c = view(ctx, k);
    if (ctx->coll_world > 1) {
        const int64_t lo = c.n * ctx->coll_rank / ctx->coll_world, hi = c.n * (ctx->coll_rank + 1) / ctx->coll_world;
        c.ids += lo, c.w += lo, c.n = hi - lo;
        if (k == IPCB_EE) c.eps += lo, c.dt += lo;
    }
    return c;
}
inline MeshView mesh_view(const ipcb_ctx* ctx) { return { ctx->dE.p, ctx->dF.p, ctx->X0.p }; }


// first stored block of each kind's records in hblk (3 / 6 / 10 upper-triangular blocks of 72 bytes per VV / EV / 4-point record).
// The vertex-vertex part is padded to an even number of blocks so that every 432- / 720-byte record behind it starts on a
// 16-byte boundary (the TMA bulk stores of k_hessian_fast need it).
inline void hess_block_offsets(const int64_t nk[4], int64_t blk0[4])
{
    const int64_t vv = 3 * nk[0] + ((3 * nk[0]) & 1);
    blk0[0] = 0, blk0[1] = vv, blk0[2] = vv + 6 * nk[1], blk0[3] = vv + 6 * nk[1] + 10 * nk[2];
}

// records of nk[VV..FV] collisions (written by the local kernels at hvid / hmask / hblk / hkey, kinds in that order)
// -> the context's resident compressed columns (outer / inner / vals, nnz); potential.cu
void hessian_assemble(ipcb_ctx* ctx, const int64_t nk[4]);
// the same in two halves: everything that only needs the records' ids (incidence sort, column ranges, active columns) on
// stream `s` — it can run beside the kernels that compute the blocks —, then the symbolic and numeric passes on the context's stream
void hessian_assemble_prepare(ipcb_ctx* ctx, const int64_t nk[4], cudaStream_t s, bool force_radix = false);
// false: a column was too large for the counting placement (seen at the pass's own synchronisation point): prepare again with
// force_radix and finish again
bool hessian_assemble_finish(ipcb_ctx* ctx, const int64_t nk[4]);
// buffers for the records of nk collisions; returns per-kind views of them (potential.cu)
void hessian_records(ipcb_ctx* ctx, const int64_t nk[4], int v_lo, int v_hi, HessOut outs[4]);
// all-zero ndof x ndof matrix
void hessian_empty(ipcb_ctx* ctx);

} // namespace ipcb
