// ctx.cuh — the library context: host mesh tables, device mirrors, resident
// broad-phase trees, candidate and collision sets, Hessian CSR.
//
// Data layout in HBM (all struct-of-arrays unless noted):
//   X0/X1      double4 per vertex (x,y,z,pad): ONE 32-byte sector per gather
//   boxes      6 x float per primitive (outward-rounded, lbvh.cpp:29-41)
//   prims      int4 per primitive: vertex ids (-1 padded) + primitive id
//   tree nodes 64 B per internal node: both child boxes + child links + range
//   pairs      int2 per candidate
//   collisions int2 ids + double weight (+ double eps_x + uint8 dtype for EE)
#pragma once
#include "common.cuh"
#include <array>
#include <chrono>

namespace ipcb {

struct alignas(8) FBox {
    float lo[3], hi[3];
};

// internal LBVH node; children: >= 0 internal node index, < 0 leaf ~sorted_position
struct alignas(64) Node {
    float lo[2][3], hi[2][3]; // child boxes (0 = left, 1 = right)
    int child[2];
    int split; // last sorted leaf of the left child
    int last;  // last sorted leaf of the right child (= of this node)
};
static_assert(sizeof(Node) == 64, "one node = one 64-byte fetch");

// 4-wide node: the (up to) four grandchildren of a binary node in one 128-byte line — half the dependent fetches per
// root-to-leaf path.  Boxes struct-of-arrays (one float4 per bound and axis); child as in Node, INT_MIN = empty slot;
// last = last sorted leaf below the child (self-query pruning).
struct alignas(128) Node4 {
    float lox[4], loy[4], loz[4], hix[4], hiy[4], hiz[4];
    int child[4];
    int last[4];
};
static_assert(sizeof(Node4) == 128, "one 4-wide node = one 128-byte line");

struct PrimSet {
    int n = 0;
    Buf<FBox> box;
    Buf<int4> prim;
};

// Morton-sorted view of a PrimSet (+ hierarchy when built as a tree)
struct Tree {
    int n = 0;
    bool has_nodes = false;
    Buf<unsigned long long> key, key_sorted;
    Buf<int> ord, ord_sorted;
    Buf<FBox> sbox;
    Buf<int4> sprim;
    Buf<Node> nodes;
    Buf<Node4> nodes4; // collapsed 4-wide hierarchy (one per binary node; the traversal only visits every other level)
    bool has_nodes4 = false;
    Buf<FBox> rmq;   // range-union tables over the sorted leaf boxes (node boxes without a bottom-up pass)
    Buf<int> parent; // [0,n-1) internal nodes, [n-1,2n-1) leaves (bottom-up refit only)
    Buf<int> flag;
    Buf<char> tmp;
};

struct PairList {
    Buf<int2> pairs;
    int64_t count = 0;
    bool sorted = false;
};

struct CollisionSet {
    // build streams (unsorted, duplicates)
    Buf<unsigned long long> key_raw, key_sorted;
    Buf<double> w_raw;
    Buf<double> eps_raw;
    Buf<unsigned char> dt_raw;
    Buf<int> idx_raw, idx_sorted;
    // final, sorted by key, merged
    Buf<int2> ids;
    Buf<double> w, eps;
    Buf<unsigned char> dtype;
    Buf<int> head, pos;
    Buf<double> wsum;
    Buf<char> cubtmp; // per stream: the four merges run concurrently
    int64_t count = 0;
    bool sorted = true;    // false after a disjoint merge concatenated the builders' records (collisions_sort restores it)
    int64_t raw_count = 0; // records appended through collisions_append since the last collisions_clear
    // exchange the FINAL records (what a NormalCollisions object is) with another set; the build scratch stays
    void swap_records(CollisionSet& o)
    {
        ids.swap(o.ids), w.swap(o.w), eps.swap(o.eps), dtype.swap(o.dtype);
        std::swap(count, o.count);
        std::swap(sorted, o.sorted);
    }
};

// the lagged tangential collisions of one kind (collisions/tangential/tangential_collision.hpp), struct of arrays
struct TangSet {
    Buf<int2> ids;
    Buf<double> w, N, mus, muk; // weight, normal force magnitude, blended static / kinetic coefficients
    Buf<double2> beta;          // closest point
    Buf<double> P;              // tangent basis: 6 per record (column 0, column 1)
    int64_t count = 0;
};

} // namespace ipcb
// A NormalCollisions object that is NOT the resident set of its context: the reference lets any number of collision sets
// exist per mesh (normal_collisions.hpp: plain containers); here a set is resident while the potential works on it and
// can be parked in / revived from one of these with ipcb_collisions_swap (O(1): buffers are exchanged, not copied).
struct ipcb_collision_set {
    int device = 0;
    ipcb::CollisionSet coll[4];
    double dmin = 0;
    bool valid = false;
};
namespace ipcb {
struct TIWork; // Tight-Inclusion scratch (ccd.cu), owned by the context
void ti_work_free(TIWork* w);

} // namespace ipcb

struct ipcb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    // auxiliary streams for independent work inside one call (tree builds, traversals, per-kind kernels);
    // fork() makes them wait for `stream`, join(k) makes `stream` wait for aux[k]
    static constexpr int NAUX = 3;
    cudaStream_t aux[NAUX] = { nullptr, nullptr, nullptr };
    cudaEvent_t ev_fork = nullptr, ev_join[NAUX] = { nullptr, nullptr, nullptr };
    void fork()
    {
        cudaEventRecord(ev_fork, stream);
        for (int k = 0; k < NAUX; k++) cudaStreamWaitEvent(aux[k], ev_fork, 0);
    }
    void join(int k)
    {
        cudaEventRecord(ev_join[k], aux[k]);
        cudaStreamWaitEvent(stream, ev_join[k], 0);
    }
    int64_t launches = 0;
    int shard_rank = 0, shard_world = 1;
    int coll_rank = 0, coll_world = 1; // energy / gradient: slice of every kind's collisions
    int row_lo = 0, row_hi = -1;       // Hessian: owned vertex range (row_hi < 0: all)
    bool timing = false; // per-stage CUDA-event timing (synchronises at every stage boundary)

    // ---- host mesh (collision_mesh.cpp:15-127)
    int nV = 0, nE = 0, nF = 0;
    std::vector<double> rest; // 3 per vertex
    std::vector<int32_t> hE, hF, hF2E;
    std::vector<int32_t> codimV, codimE;
    std::vector<double> vArea, eArea;
    // ---- device mesh
    ipcb::Buf<int2> dE;
    ipcb::Buf<int4> dF;   // f0,f1,f2,pad
    ipcb::Buf<int4> dF2E; // e0,e1,e2,pad
    ipcb::Buf<double4> dRest;
    ipcb::Buf<double> dVArea, dEArea;
    ipcb::Buf<int> dCodimV, dCodimE;
    // CollisionMesh::can_collide as a descriptor (ipcb_mesh_set_collision_filter): per-vertex patch labels and / or the
    // number of dynamic vertices; dCodimELocal: ids of the codim edges' endpoints in the re-indexed vertex set of the
    // codimensional edge-vertex pass (candidates.cpp:83-108)
    bool filter_patches = false;
    int filter_n_dynamic = -1;
    ipcb::Buf<int> dPatch;
    ipcb::Buf<int2> dCodimELocal;
    bool filter_on() const { return filter_patches || filter_n_dynamic >= 0; }

    // ---- positions
    ipcb::Buf<double> stageA, stageB; // col-major staging for host entry points
    ipcb::Buf<double4> X0, X1;

    // ---- broad phase
    bool built = false;
    bool swept = false;
    ipcb::PrimSet vset, eset, fset, cvset, ceset;
    ipcb::Tree vtree, etree, ftree, cvtree, cetree;
    bool vtree_ok = false, etree_ok = false, ftree_ok = false;
    bool vorder_valid = false; // vtree.ord_sorted holds a permutation of all vertices (Morton order of the last build)
    // sweep and prune (IPCB_BROAD_SAP): axis-sorted views of the three primitive sets
    int broad_method = 0;
    int sap_axis = 0;
    bool sap_axis_ok = false, sap_ok[3] = { false, false, false };
    ipcb::Tree sap_v, sap_e, sap_f;
    ipcb::Buf<float> scene; // 6 floats: min xyz, max xyz
    bool scene_covers_positions = false; // the scene box was reduced from the CURRENT X0 / X1 (swept build of this call)
    ipcb::PairList detected[6];

    // ---- candidates / collisions
    ipcb::PairList cand[4];
    unsigned long long cand_overflow[4] = { 0, 0, 0, 0 }; // pairs a kind would have had when it exceeded max_pairs()
    ipcb::CollisionSet coll[4];
    double dmin = 0;
    bool coll_valid = false;
    // IMPROVED_MAX_APPROX: adjacency tables (collision_mesh.cpp:247-307) as CSR, built on first use;
    // sub-element candidate keys
    bool adj_ready = false;
    // deferred IMPROVED_MAX_APPROX build of a sharded context (collisions.cu: collisions_corrections_*)
    bool ima_pending = false, ima_area = false;
    int64_t ima_raw[4] = { 0, 0, 0, 0 }, ima_nu[4] = { 0, 0, 0, 0 };
    int adj_max_ve = 0; // largest number of edges at a vertex
    ipcb::Buf<int> adjVVoff, adjVV, adjVEoff, adjVE, adjEVoff, adjEV;
    ipcb::Buf<unsigned char> adjBoundary;
    ipcb::Buf<unsigned long long> subkey, subkey_sorted, subuniq[4];

    // ---- potential
    ipcb::Buf<double> dScalar; // small device scalars (energy, toi, ...)
    ipcb::Buf<double> dGrad;
    // hessian assembly (potential.cu): per-collision records, vertex incidences, per-column sorted items
    ipcb::Buf<unsigned long long> hkey, hkey_sorted; // incidence keys (vertex << 32 | collision * 4 + point); also pair sorting
    ipcb::Buf<int4> hvid;                            // stencil vertex ids per collision (-1 padded)
    ipcb::Buf<unsigned short> hmask;                 // 16 x 9-bit non-zero masks per collision (slot = col point * 4 + row point)
    ipcb::Buf<unsigned long long> hcount, hcursor;   // per vertex key: incidences by record size (3 x 21 bits), placement cursors
    ipcb::Buf<int> hactive;                          // columns with anything to assemble, in visiting order
    ipcb::Buf<char> hseltmp;
    ipcb::Buf<double> hblk;                          // upper-triangular vertex blocks: 3 / 6 / 10 x 9 doubles per VV / EV / 4-point record
    ipcb::Buf<unsigned char> hflag;                  // row block: does the collision touch an owned vertex
    ipcb::Buf<int> hsel;                             // row block: per kind, the collisions that do (ascending)
    ipcb::Buf<int> hslow;                            // edge-edge collisions handed to the general kernel
    ipcb::Buf<int2> hcolb;                           // per column vertex: first EV incidence, first 4-point incidence
    ipcb::Buf<int> hcolinc, hcolR, hitemoff;         // per column vertex: first incidence, #items, first item
    ipcb::Buf<unsigned> hsref;                       // per item, grouped by column then by row vertex: block slot
    ipcb::Buf<int2> hudesc;                          // per unique block: (first item within its column, row vertex)
    ipcb::Buf<int> hcolU;                            // unique blocks per column
    ipcb::Buf<int> hcnt;                             // entries per scalar column
    ipcb::Buf<int> hbig;                             // columns too large for one warp's shared memory
    ipcb::Buf<char> hscratch;                        // global sort scratch for huge columns
    size_t hscratch_items = 0;
    bool hfast_attr_set = false, colsort_attr_set = false;
    bool hess_counting_used = false; // the last hessian_assemble_prepare placed the incidences by counting
    bool hess_attr_set = false;
    ipcb::Buf<int> outer, inner;
    ipcb::Buf<double> vals;
    int64_t nnz = 0;
    ipcb::Buf<char> cubtmp;

    // ---- friction
    ipcb::TangSet tang[4], tang_tmp;
    bool tang_valid = false;
    ipcb::Buf<double> dMuS, dMuK; // per-vertex coefficients of the last tangential build

    // ---- ccd
    ipcb::TIWork* ti_work = nullptr;

    // ---- counters read back through pinned memory
    ipcb::Pinned pinned;
    ipcb::Buf<unsigned long long> dCounters; // 16 device counters

    // ---- per-stage timing of the last API call
    std::vector<std::pair<std::string, float>> stage_ms;
    std::vector<std::string> stage_names_keepalive;
};

namespace ipcb {

// RAII stage timer using CUDA events on the context's stream; a no-op unless ctx->timing is set.
// With a stream argument it times ONE kernel (or one group of launches) on the stream it was launched on — the
// per-kernel durations behind bench.py's rooflines; the names of such timers start with "k:".
struct Stage {
    ipcb_ctx* ctx;
    const char* name;
    cudaStream_t s;
    cudaEvent_t a = nullptr, b = nullptr;
    Stage(ipcb_ctx* c, const char* n, cudaStream_t on = nullptr) : ctx(c), name(n), s(on ? on : c->stream)
    {
        if (!ctx->timing) return;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, s);
    }
    ~Stage()
    {
        if (!a) return;
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        ctx->stage_ms.emplace_back(name, ms);
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
};

// positions
void upload_positions(ipcb_ctx* ctx, const double* hV, int ld, Buf<double>& stage);
void convert_positions(ipcb_ctx* ctx, const double* dV, int ld, Buf<double4>& X);

// broad phase (broad.cu)
void broad_build(ipcb_ctx* ctx, bool swept, double inflation_radius);
void broad_detect(ipcb_ctx* ctx, int kind, PairList& out);
bool candidates_build(ipcb_ctx* ctx, bool swept, double inflation_radius, bool allow_overflow = false);
size_t max_pairs();
unsigned long long traverse_chunk(ipcb_ctx* ctx, int kind, int qb, int qe, int* range);
void sort_pairs(ipcb_ctx* ctx, PairList& pl);
bool has_intersections(ipcb_ctx* ctx, double inflation_radius);

// collisions (collisions.cu)
void collisions_build(ipcb_ctx* ctx, double dhat, double dmin, int flags, bool may_defer = false);
void collisions_clear(ipcb_ctx* ctx);
void collisions_append_dev(ipcb_ctx* ctx, int kind, int64_t n, const int32_t* d_ids, const double* d_w, const double* d_eps,
                           const uint8_t* d_dt);
void collisions_append_packed_dev(ipcb_ctx* ctx, const void* d_buffer, const int64_t n[4], const int64_t ids_off[4], const int64_t w_off[4],
                                  int64_t eps_off, int64_t dt_off);
void collisions_merge(ipcb_ctx* ctx, double dmin, int flags);
void collisions_corrections_keys(ipcb_ctx* ctx, int64_t n[4]);
void collisions_corrections_pack(ipcb_ctx* ctx, unsigned long long* d_out);
void collisions_corrections_apply(ipcb_ctx* ctx, const unsigned long long* d_keys, const int64_t n[4], int rank, int world);
void collisions_sort(ipcb_ctx* ctx, int kind);
double collisions_min_distance(ipcb_ctx* ctx);

// potential (potential.cu)
void barrier_energy(ipcb_ctx* ctx, const ipcb_barrier_params& bp, double* d_out);
void barrier_gradient(ipcb_ctx* ctx, const ipcb_barrier_params& bp, double* d_grad);
void barrier_hessian(ipcb_ctx* ctx, const ipcb_barrier_params& bp, int psd_mode);
void hessian_balanced_row_blocks(ipcb_ctx* ctx, int world, int32_t* bounds);

// friction (friction.cu)
void tangential_build(ipcb_ctx* ctx, const ipcb_barrier_params& bp, const double* d_mu_s, const double* d_mu_k);
void friction_energy(ipcb_ctx* ctx, double eps_v, double* d_out);
void friction_gradient(ipcb_ctx* ctx, double eps_v, double* d_grad);
void friction_hessian(ipcb_ctx* ctx, double eps_v, int psd_mode);

// ccd (ccd.cu)
void ccd_stepsize(ipcb_ctx* ctx, double min_distance, const ipcb_ccd_params& p, double* d_out);
void ccd_stepsize_streaming(ipcb_ctx* ctx, double min_distance, const ipcb_ccd_params& p, double* d_out);
double noncandidate_stepsize(ipcb_ctx* ctx, bool difference, double dhat);
void ccd_narrow_phase(ipcb_ctx* ctx, int kind, int64_t n, const double* h_t0, const double* h_t1, double min_distance, double tmax,
                      const ipcb_ccd_params& p, uint8_t* h_hit, double* h_toi);

ipcb_ccd_params resolve_ccd(const ipcb_ccd_params* p);

} // namespace ipcb
