// ipcb200.hpp — C++17 host-side mirror of the reference's operator interface for the per-step
// contact path, header-only on top of the C ABI (include/ipcb200.h).  Same names, argument meaning
// and error behaviour as ipc-toolkit v1.6.0 (exceptions instead of status codes), but Eigen-free:
// matrices are passed as column-major views (pointer, rows, leading dimension) exactly like
// Eigen::Ref<const MatrixXd> (utils/eigen_ext.hpp:17).  ipc_toolkit_adapter.hpp wraps these classes
// into real ipc::BroadPhase / ipc:: types when Eigen and the toolkit headers are available.
//
//   reference (src/ipc/)                                  here
//   collision_mesh.hpp: CollisionMesh                     ipcb200::CollisionMesh
//   broad_phase/broad_phase.hpp: BroadPhase               ipcb200::CudaBroadPhase
//   candidates/candidates.hpp: Candidates                 ipcb200::Candidates
//   collisions/normal/normal_collisions.hpp               ipcb200::NormalCollisions
//   potentials/barrier_potential.hpp: BarrierPotential    ipcb200::BarrierPotential
//   ipc.hpp: compute_collision_free_stepsize, ...         ipcb200::compute_collision_free_stepsize
//   ccd/{tight_inclusion,additive}_ccd.hpp                ipcb200::TightInclusionCCD / AdditiveCCD
#pragma once
#include "../../include/ipcb200.h"

#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace ipcb200 {

using index_t = int32_t; // config.hpp.in:30-34

// column-major matrix view == Eigen::Ref<const Matrix<T, Dynamic, Dynamic>>
template <typename T> struct ConstRef {
    const T* data = nullptr;
    index_t rows = 0, cols = 0, ld = 0;
    ConstRef() = default;
    ConstRef(const T* d, index_t r, index_t c, index_t l = -1) : data(d), rows(r), cols(c), ld(l < 0 ? (r > 0 ? r : 1) : l) { }
};
using MatrixXd = ConstRef<double>;
using MatrixXi = ConstRef<index_t>;

inline void check(int rc)
{
    if (rc != 0) throw std::runtime_error(ipcb_last_error()); // log_and_throw_error, utils/logger.cpp:41-45
}

enum class PSDProjectionMethod { NONE = IPCB_PSD_NONE, CLAMP = IPCB_PSD_CLAMP, ABS = IPCB_PSD_ABS }; // utils/eigen_ext.hpp:202-206

// ccd/narrow_phase_ccd.hpp — the two implementations on the reference path
struct NarrowPhaseCCD {
    ipcb_ccd_params params;
};
struct TightInclusionCCD : NarrowPhaseCCD { // ccd/tight_inclusion_ccd.hpp:11-19
    static constexpr double DEFAULT_TOLERANCE = 1e-6;
    static constexpr long DEFAULT_MAX_ITERATIONS = 10'000'000L;
    static constexpr double DEFAULT_CONSERVATIVE_RESCALING = 0.8;
    explicit TightInclusionCCD(double tolerance = DEFAULT_TOLERANCE, long max_iterations = DEFAULT_MAX_ITERATIONS,
                               double conservative_rescaling = DEFAULT_CONSERVATIVE_RESCALING)
    {
        params = { IPCB_CCD_TIGHT_INCLUSION, tolerance, max_iterations, conservative_rescaling };
    }
};
struct AdditiveCCD : NarrowPhaseCCD { // ccd/additive_ccd.hpp:21-26
    static constexpr long DEFAULT_MAX_ITERATIONS = 10'000'000L;
    static constexpr double DEFAULT_CONSERVATIVE_RESCALING = 0.9;
    explicit AdditiveCCD(long max_iterations = DEFAULT_MAX_ITERATIONS, double conservative_rescaling = DEFAULT_CONSERVATIVE_RESCALING)
    {
        params = { IPCB_CCD_ADDITIVE, 0.0, max_iterations, conservative_rescaling };
    }
};
inline const TightInclusionCCD DEFAULT_NARROW_PHASE_CCD {}; // ccd/default_narrow_phase_ccd.cpp

// collision_mesh.hpp — owns the library context: device mirrors of the mesh, the broad phase and the
// resident candidate / collision sets
class CollisionMesh {
public:
    CollisionMesh(MatrixXd rest_positions, MatrixXi edges = {}, MatrixXi faces = {}, int device = 0)
    {
        if (rest_positions.cols != 3) throw std::invalid_argument("this path supports 3D meshes only");
        ipcb_ctx* raw = nullptr;
        check(ipcb_ctx_create(device, &raw));
        m_ctx.reset(raw, ipcb_ctx_destroy);
        m_nv = rest_positions.rows, m_ne = edges.rows, m_nf = faces.rows;
        check(ipcb_mesh_set(raw, m_nv, rest_positions.data, rest_positions.ld, m_ne, edges.data, edges.ld, m_nf, faces.data, faces.ld));
    }
    size_t num_vertices() const { return m_nv; }
    size_t num_edges() const { return m_ne; }
    size_t num_faces() const { return m_nf; }
    int dim() const { return 3; }
    size_t ndof() const { return 3 * size_t(m_nv); }
    size_t num_codim_vertices() const
    {
        int32_t n;
        check(ipcb_mesh_num_codim_vertices(ctx(), &n));
        return n;
    }
    size_t num_codim_edges() const
    {
        int32_t n;
        check(ipcb_mesh_num_codim_edges(ctx(), &n));
        return n;
    }
    std::vector<index_t> faces_to_edges() const // column-major num_faces x 3
    {
        std::vector<index_t> out(3 * size_t(m_nf));
        check(ipcb_mesh_faces_to_edges(ctx(), out.data()));
        return out;
    }
    /// CollisionMesh::can_collide (collision_mesh.hpp:338) as the descriptor the C ABI carries: the intersection of
    /// make_vertex_patches_filter(patch_ids) and make_static_obstacle_filter(n_dynamic) (collision_filter.hpp:113-143);
    /// an empty vector / a negative count switches that factory off.  Arbitrary callables: see ipc_toolkit_adapter.hpp.
    void set_collision_filter(const std::vector<index_t>& patch_ids, index_t n_dynamic = -1) const
    {
        if (!patch_ids.empty() && patch_ids.size() != size_t(m_nv)) throw std::invalid_argument("patch_ids must hold one label per vertex");
        check(ipcb_mesh_set_collision_filter(ctx(), patch_ids.empty() ? nullptr : patch_ids.data(), n_dynamic));
    }
    /// IPCB_BROAD_LBVH (default) or IPCB_BROAD_SAP for every later broad-phase build on this mesh
    void set_broad_phase_method(int method) const { check(ipcb_ctx_set_broad_phase_method(ctx(), method)); }
    std::vector<double> vertex_areas() const { return areas().first; }
    std::vector<double> edge_areas() const { return areas().second; }
    ipcb_ctx* ctx() const { return m_ctx.get(); }
    /// the NormalCollisions whose records are the context's resident set (several sets may exist per mesh; see NormalCollisions)
    mutable const void* resident_collisions = nullptr;

private:
    std::pair<std::vector<double>, std::vector<double>> areas() const
    {
        std::vector<double> va(m_nv), ea(m_ne);
        check(ipcb_mesh_areas(ctx(), va.data(), ea.data()));
        return { va, ea };
    }
    std::shared_ptr<ipcb_ctx> m_ctx;
    index_t m_nv = 0, m_ne = 0, m_nf = 0;
};

using Pair = std::array<index_t, 2>;

// broad_phase/broad_phase.hpp:19-133 — the CUDA LBVH; candidates come back sorted, unordered kinds as (min, max)
class CudaBroadPhase {
public:
    explicit CudaBroadPhase(const CollisionMesh& mesh) : m_mesh(&mesh) { }
    std::string name() const { return "CudaLBVH"; }
    void build(MatrixXd vertices, double inflation_radius = 0)
    {
        check(ipcb_broad_build_static(m_mesh->ctx(), vertices.data, vertices.ld, inflation_radius, IPCB_BOXES_FLOAT));
    }
    void build(MatrixXd vertices_t0, MatrixXd vertices_t1, double inflation_radius = 0)
    {
        check(ipcb_broad_build_swept(m_mesh->ctx(), vertices_t0.data, vertices_t1.data, vertices_t0.ld, inflation_radius, IPCB_BOXES_FLOAT));
    }
    void clear() { }
    void detect_vertex_vertex_candidates(std::vector<Pair>& c) const { detect(IPCB_VV, c); }
    void detect_edge_vertex_candidates(std::vector<Pair>& c) const { detect(IPCB_EV, c); }
    void detect_edge_edge_candidates(std::vector<Pair>& c) const { detect(IPCB_EE, c); }
    void detect_face_vertex_candidates(std::vector<Pair>& c) const { detect(IPCB_FV, c); }
    void detect_edge_face_candidates(std::vector<Pair>& c) const { detect(IPCB_EF, c); }
    void detect_face_face_candidates(std::vector<Pair>& c) const { detect(IPCB_FF, c); }

private:
    void detect(int kind, std::vector<Pair>& out) const
    {
        int64_t n = 0;
        check(ipcb_broad_detect(m_mesh->ctx(), kind, &n));
        const size_t old = out.size();
        out.resize(old + size_t(n));
        if (n) check(ipcb_broad_fetch(m_mesh->ctx(), kind, out[old].data()));
    }
    const CollisionMesh* m_mesh;
};

// candidates/candidates.hpp — handle to the context's resident candidate set; the host vectors are
// materialised lazily (they are the compatibility path, not the timed one)
class Candidates {
public:
    void build(const CollisionMesh& mesh, MatrixXd vertices, double inflation_radius)
    {
        check(ipcb_candidates_build_static(mesh.ctx(), vertices.data, vertices.ld, inflation_radius, m_counts));
        bind(mesh);
    }
    void build(const CollisionMesh& mesh, MatrixXd vertices_t0, MatrixXd vertices_t1, double inflation_radius)
    {
        check(ipcb_candidates_build_swept(mesh.ctx(), vertices_t0.data, vertices_t1.data, vertices_t0.ld, inflation_radius, m_counts));
        bind(mesh);
    }
    size_t size() const { return size_t(m_counts[0] + m_counts[1] + m_counts[2] + m_counts[3]); }
    bool empty() const { return size() == 0; }
    const std::vector<Pair>& vv_candidates() const { return get(IPCB_VV); }
    const std::vector<Pair>& ev_candidates() const { return get(IPCB_EV); }
    const std::vector<Pair>& ee_candidates() const { return get(IPCB_EE); }
    const std::vector<Pair>& fv_candidates() const { return get(IPCB_FV); }
    double compute_collision_free_stepsize(const CollisionMesh& mesh, MatrixXd vertices_t0, MatrixXd vertices_t1, double min_distance = 0.0,
                                           const NarrowPhaseCCD& ccd = DEFAULT_NARROW_PHASE_CCD) const
    {
        double step = 1.0;
        check(ipcb_ccd_stepsize_from_candidates(mesh.ctx(), vertices_t0.data, vertices_t1.data, vertices_t0.ld, min_distance, &ccd.params, &step));
        return step;
    }
    bool is_step_collision_free(const CollisionMesh& mesh, MatrixXd vertices_t0, MatrixXd vertices_t1, double min_distance = 0.0,
                                const NarrowPhaseCCD& ccd = DEFAULT_NARROW_PHASE_CCD) const
    {
        return compute_collision_free_stepsize(mesh, vertices_t0, vertices_t1, min_distance, ccd) >= 1.0;
    }
    // candidates.cpp:294-338
    double compute_noncandidate_conservative_stepsize(const CollisionMesh& mesh, MatrixXd displacements, double dhat) const
    {
        double step = 1.0;
        check(ipcb_candidates_noncandidate_stepsize(mesh.ctx(), displacements.data, displacements.ld, dhat, &step));
        return step;
    }
    // candidates.cpp:340-363 (the resident candidates are replaced when the full CCD had to run)
    double compute_cfl_stepsize(const CollisionMesh& mesh, MatrixXd vertices_t0, MatrixXd vertices_t1, double dhat, double min_distance = 0.0,
                                const NarrowPhaseCCD& ccd = DEFAULT_NARROW_PHASE_CCD) const
    {
        double step = 1.0;
        check(ipcb_candidates_cfl_stepsize(mesh.ctx(), vertices_t0.data, vertices_t1.data, vertices_t0.ld, dhat, min_distance, &ccd.params, &step));
        return step;
    }

private:
    void bind(const CollisionMesh& mesh)
    {
        m_mesh = &mesh;
        for (auto& h : m_host) h.clear();
        for (bool& f : m_fetched) f = false;
    }
    const std::vector<Pair>& get(int kind) const
    {
        if (!m_mesh) throw std::runtime_error("Candidates not built");
        if (!m_fetched[kind]) {
            m_host[kind].resize(size_t(m_counts[kind]));
            if (m_counts[kind]) check(ipcb_candidates_fetch(m_mesh->ctx(), kind, m_host[kind][0].data()));
            m_fetched[kind] = true;
        }
        return m_host[kind];
    }
    const CollisionMesh* m_mesh = nullptr;
    int64_t m_counts[4] = { 0, 0, 0, 0 };
    mutable std::vector<Pair> m_host[4];
    mutable bool m_fetched[4] = { false, false, false, false };
};

// collisions/normal/normal_collisions.hpp — IPC collision set type.
// The reference's NormalCollisions are plain containers: any number may exist per mesh.  The context works on ONE
// resident set, so every object here owns a collision-set object of the library (ipcb_collision_set) and is swapped in
// (ipcb_collisions_swap, O(1)) when it is used; the previous owner's records are parked in ITS object first.
class NormalCollisions {
public:
    struct Records { // one typed vector of the reference, struct-of-arrays
        std::vector<Pair> ids; // VV (v0,v1); EV (edge,vertex); EE (ea,eb); FV (face,vertex)
        std::vector<double> weight, eps_x;
        std::vector<uint8_t> dtype;
    };
    NormalCollisions() = default;
    NormalCollisions(const NormalCollisions&) = delete;
    NormalCollisions& operator=(const NormalCollisions&) = delete;
    ~NormalCollisions()
    {
        if (m_mesh && m_mesh->resident_collisions == this) m_mesh->resident_collisions = nullptr;
        if (m_set) ipcb_collision_set_destroy(m_set);
    }
    void set_use_area_weighting(bool v) { m_area = v; }
    bool use_area_weighting() const { return m_area; }
    /// normal_collisions.hpp:29-39; OGC is outside this path
    enum class CollisionSetType { IPC, IMPROVED_MAX_APPROX, OGC };
    void set_collision_set_type(CollisionSetType t)
    {
        if (t == CollisionSetType::OGC) throw std::invalid_argument("CollisionSetType::OGC is outside this path");
        m_type = t;
    }
    CollisionSetType collision_set_type() const { return m_type; }
    /// defer_corrections (IMPROVED_MAX_APPROX built by several builders over disjoint candidate shards, IPCB_DEFER_CORRECTIONS): the
    /// build stops after this builder's sub-element pairs; exchange correction_keys() and finish with apply_corrections()
    void build(const CollisionMesh& mesh, MatrixXd vertices, double dhat, double dmin = 0, bool defer_corrections = false)
    {
        claim(mesh);
        check(ipcb_collisions_build(mesh.ctx(), vertices.data, vertices.ld, dhat, dmin, flags(defer_corrections), m_counts));
        m_built = !defer_corrections;
    }
    void build(const Candidates&, const CollisionMesh& mesh, MatrixXd vertices, double dhat, double dmin = 0, bool defer_corrections = false)
    {
        claim(mesh);
        check(ipcb_collisions_build_from_candidates(mesh.ctx(), vertices.data, vertices.ld, dhat, dmin, flags(defer_corrections), m_counts));
        m_built = !defer_corrections;
    }
    /// the four lists of this builder's unique sub-element pairs after a deferred build (opaque 64-bit keys)
    std::array<std::vector<uint64_t>, 4> correction_keys() const
    {
        int64_t n[4];
        check(ipcb_collisions_corrections_keys(m_mesh->ctx(), n));
        std::vector<uint64_t> all(size_t(n[0] + n[1] + n[2] + n[3]) + 1);
        check(ipcb_collisions_corrections_pack(m_mesh->ctx(), all.data()));
        std::array<std::vector<uint64_t>, 4> out;
        size_t o = 0;
        for (int k = 0; k < 4; k++) out[k].assign(all.begin() + o, all.begin() + o + n[k]), o += size_t(n[k]);
        return out;
    }
    /// lists: for each of the four lists the keys of ALL builders (duplicates allowed); adds the corrections of slice `rank` of `world`
    void apply_corrections(const std::array<std::vector<uint64_t>, 4>& lists, int rank, int world)
    {
        std::vector<uint64_t> all;
        int64_t n[4];
        for (int k = 0; k < 4; k++) n[k] = int64_t(lists[k].size()), all.insert(all.end(), lists[k].begin(), lists[k].end());
        all.push_back(0);
        check(ipcb_collisions_corrections_apply(m_mesh->ctx(), all.data(), n, rank, world, m_counts));
        m_built = true;
    }
    size_t size() const { return size_t(m_counts[0] + m_counts[1] + m_counts[2] + m_counts[3]); }
    bool empty() const { return size() == 0; }
    size_t count(int kind) const { return size_t(m_counts[kind]); }
    Records records(int kind) const
    {
        make_resident();
        Records r;
        const size_t n = count(kind);
        r.ids.resize(n), r.weight.resize(n), r.eps_x.resize(n), r.dtype.resize(n);
        if (n) check(ipcb_collisions_fetch(m_mesh->ctx(), kind, r.ids[0].data(), r.weight.data(), r.eps_x.data(), r.dtype.data()));
        return r;
    }
    double compute_minimum_distance(const CollisionMesh& mesh, MatrixXd vertices) const
    {
        make_resident();
        double d;
        check(ipcb_collisions_min_distance(mesh.ctx(), vertices.data, vertices.ld, &d));
        return d;
    }
    /// Fill the set from the records of several builders and merge them like NormalCollisionsBuilder::merge
    /// (normal_collisions_builder.cpp:547-689): equal collisions united, weights added, weight == 0 dropped.
    /// `builders[b][kind]`; disjoint_shards promises builders over disjoint candidate shards (ranks of a sharded build).
    void assign(const CollisionMesh& mesh, const std::vector<std::array<Records, 4>>& builders, double dmin = 0, bool disjoint_shards = false)
    {
        claim(mesh);
        check(ipcb_collisions_clear(mesh.ctx()));
        for (const auto& b : builders)
            for (int kind = 0; kind < 4; kind++) {
                const Records& r = b[kind];
                if (r.ids.empty()) continue;
                check(ipcb_collisions_append(mesh.ctx(), kind, int64_t(r.ids.size()), r.ids[0].data(), r.weight.data(),
                                             kind == IPCB_EE ? r.eps_x.data() : nullptr, kind == IPCB_EE ? r.dtype.data() : nullptr));
            }
        check(ipcb_collisions_merge(mesh.ctx(), dmin, disjoint_shards ? IPCB_MERGE_DISJOINT_SHARDS : 0, m_counts));
        m_built = true;
    }
    /// make this set the resident one of its mesh (the potentials call it before they evaluate)
    void make_resident() const
    {
        if (!m_mesh || !m_built) throw std::runtime_error("NormalCollisions has not been built");
        if (m_mesh->resident_collisions == this) return;
        park_resident(*m_mesh);
        int64_t counts[4];
        check(ipcb_collisions_swap(m_mesh->ctx(), handle(), counts)); // this set's records become resident
        m_mesh->resident_collisions = this;
    }

private:
    ipcb_collision_set* handle() const
    {
        if (!m_set) check(ipcb_collision_set_create(m_mesh->ctx(), &m_set));
        return m_set;
    }
    static void park_resident(const CollisionMesh& mesh)
    {
        if (auto* owner = static_cast<const NormalCollisions*>(mesh.resident_collisions)) {
            int64_t counts[4];
            check(ipcb_collisions_swap(mesh.ctx(), owner->handle(), counts)); // the owner's records move into its object
        }
        mesh.resident_collisions = nullptr;
    }
    // before this object (re)builds on `mesh`: another set's records leave the context, this one becomes the owner
    void claim(const CollisionMesh& mesh)
    {
        if (m_mesh && m_mesh != &mesh) throw std::runtime_error("a NormalCollisions object cannot move to another mesh");
        m_mesh = &mesh;
        if (mesh.resident_collisions != this) park_resident(mesh);
        mesh.resident_collisions = this;
    }
    const CollisionMesh* m_mesh = nullptr;
    mutable ipcb_collision_set* m_set = nullptr;
    int64_t m_counts[4] = { 0, 0, 0, 0 };
    bool m_area = false, m_built = false;
    CollisionSetType m_type = CollisionSetType::IPC;
    int32_t flags(bool defer) const
    {
        return (m_area ? IPCB_USE_AREA_WEIGHTING : 0) | (m_type == CollisionSetType::IMPROVED_MAX_APPROX ? IPCB_SET_IMPROVED_MAX_APPROX : 0)
            | (defer ? IPCB_DEFER_CORRECTIONS : 0);
    }
};

// Eigen::SparseMatrix<double> in compressed-column form (== compressed rows of the symmetric matrix)
struct SparseMatrix {
    index_t rows = 0, cols = 0;
    std::vector<index_t> outer, inner;
    std::vector<double> values;
    size_t nonZeros() const { return values.size(); }
};

// potentials/barrier_potential.hpp
class BarrierPotential {
public:
    explicit BarrierPotential(double dhat, double stiffness = 1.0, bool use_physical_barrier = false)
        : m_bp { dhat, stiffness, use_physical_barrier ? 1 : 0 }
    {
    }
    double dhat() const { return m_bp.dhat; }
    double stiffness() const { return m_bp.stiffness; }
    bool use_physical_barrier() const { return m_bp.use_physical_barrier != 0; }
    double operator()(const NormalCollisions& c, const CollisionMesh& mesh, MatrixXd X) const
    {
        c.make_resident();
        double e;
        check(ipcb_barrier_energy(mesh.ctx(), X.data, X.ld, &m_bp, &e));
        return e;
    }
    std::vector<double> gradient(const NormalCollisions& c, const CollisionMesh& mesh, MatrixXd X) const
    {
        c.make_resident();
        std::vector<double> g(mesh.ndof());
        check(ipcb_barrier_gradient(mesh.ctx(), X.data, X.ld, &m_bp, g.data()));
        return g;
    }
    SparseMatrix hessian(const NormalCollisions& c, const CollisionMesh& mesh, MatrixXd X,
                         PSDProjectionMethod project_hessian_to_psd = PSDProjectionMethod::NONE) const
    {
        c.make_resident();
        int64_t nnz = 0;
        check(ipcb_barrier_hessian(mesh.ctx(), X.data, X.ld, &m_bp, int(project_hessian_to_psd), &nnz));
        SparseMatrix H;
        H.rows = H.cols = index_t(mesh.ndof());
        H.outer.resize(mesh.ndof() + 1), H.inner.resize(size_t(nnz)), H.values.resize(size_t(nnz));
        check(ipcb_barrier_hessian_fetch(mesh.ctx(), H.outer.data(), H.inner.data(), H.values.data()));
        return H;
    }

private:
    ipcb_barrier_params m_bp;
};

// collisions/tangential/tangential_collisions.hpp — the lagged tangential set of the mesh's context (isotropic coefficients)
class TangentialCollisions {
public:
    struct Records {
        std::vector<Pair> ids;
        std::vector<double> weight, normal_force_magnitude, mu_s, mu_k;
        std::vector<std::array<double, 2>> closest_point;
        std::vector<std::array<double, 6>> tangent_basis; // column 0, column 1
    };
    /// build(mesh, vertices, collisions, normal_potential, mu_s, mu_k) — tangential_collisions.cpp:62-171; per-vertex coefficients
    void build(const CollisionMesh& mesh, MatrixXd vertices, const NormalCollisions& collisions, const BarrierPotential& normal_potential,
               const std::vector<double>& mu_s, const std::vector<double>& mu_k)
    {
        collisions.make_resident();
        if (mu_s.size() != mesh.num_vertices() || mu_k.size() != mesh.num_vertices()) throw std::invalid_argument("one coefficient per vertex");
        const ipcb_barrier_params bp { normal_potential.dhat(), normal_potential.stiffness(), normal_potential.use_physical_barrier() ? 1 : 0 };
        check(ipcb_tangential_build(mesh.ctx(), vertices.data, vertices.ld, &bp, mu_s.data(), mu_k.data(), m_counts));
        m_mesh = &mesh;
    }
    size_t size() const { return size_t(m_counts[0] + m_counts[1] + m_counts[2] + m_counts[3]); }
    bool empty() const { return size() == 0; }
    size_t count(int kind) const { return size_t(m_counts[kind]); }
    Records records(int kind) const
    {
        Records r;
        const size_t n = count(kind);
        r.ids.resize(n), r.weight.resize(n), r.normal_force_magnitude.resize(n), r.mu_s.resize(n), r.mu_k.resize(n);
        r.closest_point.resize(n), r.tangent_basis.resize(n);
        if (n)
            check(ipcb_tangential_fetch(m_mesh->ctx(), kind, r.ids[0].data(), r.weight.data(), r.normal_force_magnitude.data(), r.mu_s.data(),
                                        r.mu_k.data(), r.closest_point[0].data(), r.tangent_basis[0].data()));
        return r;
    }

private:
    const CollisionMesh* m_mesh = nullptr;
    int64_t m_counts[4] = { 0, 0, 0, 0 };
};

// potentials/friction_potential.hpp
class FrictionPotential {
public:
    explicit FrictionPotential(double eps_v) : m_eps_v(eps_v) { }
    double eps_v() const { return m_eps_v; }
    double operator()(const TangentialCollisions&, const CollisionMesh& mesh, MatrixXd velocities) const
    {
        double e;
        check(ipcb_friction_energy(mesh.ctx(), velocities.data, velocities.ld, m_eps_v, &e));
        return e;
    }
    std::vector<double> gradient(const TangentialCollisions&, const CollisionMesh& mesh, MatrixXd velocities) const
    {
        std::vector<double> g(mesh.ndof());
        check(ipcb_friction_gradient(mesh.ctx(), velocities.data, velocities.ld, m_eps_v, g.data()));
        return g;
    }
    SparseMatrix hessian(const TangentialCollisions&, const CollisionMesh& mesh, MatrixXd velocities,
                         PSDProjectionMethod project_hessian_to_psd = PSDProjectionMethod::NONE) const
    {
        int64_t nnz = 0;
        check(ipcb_friction_hessian(mesh.ctx(), velocities.data, velocities.ld, m_eps_v, int(project_hessian_to_psd), &nnz));
        SparseMatrix H;
        H.rows = H.cols = index_t(mesh.ndof());
        H.outer.resize(mesh.ndof() + 1), H.inner.resize(size_t(nnz)), H.values.resize(size_t(nnz));
        check(ipcb_barrier_hessian_fetch(mesh.ctx(), H.outer.data(), H.inner.data(), H.values.data()));
        return H;
    }

private:
    double m_eps_v;
};

// ipc.hpp / ipc.cpp:105-166 (3D)
inline bool has_intersections(const CollisionMesh& mesh, MatrixXd vertices)
{
    int32_t r = 0;
    check(ipcb_has_intersections(mesh.ctx(), vertices.data, vertices.ld, &r));
    return r != 0;
}

// ipc.hpp:45-51 / ipc.cpp:45-101
inline double compute_collision_free_stepsize(const CollisionMesh& mesh, MatrixXd vertices_t0, MatrixXd vertices_t1, double min_distance = 0.0,
                                              const NarrowPhaseCCD& ccd = DEFAULT_NARROW_PHASE_CCD)
{
    double step = 1.0;
    check(ipcb_ccd_stepsize(mesh.ctx(), vertices_t0.data, vertices_t1.data, vertices_t0.ld, min_distance, &ccd.params, &step));
    return step;
}
// ipc.hpp / ipc.cpp:20-43
inline bool is_step_collision_free(const CollisionMesh& mesh, MatrixXd vertices_t0, MatrixXd vertices_t1, double min_distance = 0.0,
                                   const NarrowPhaseCCD& ccd = DEFAULT_NARROW_PHASE_CCD)
{
    return compute_collision_free_stepsize(mesh, vertices_t0, vertices_t1, min_distance, ccd) >= 1.0;
}

} // namespace ipcb200
