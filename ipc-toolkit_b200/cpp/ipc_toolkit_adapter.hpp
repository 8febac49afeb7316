// ipc_toolkit_adapter.hpp — drop-in adapters for a build of ipc-toolkit v1.6.0.
//
// NOT compiled in this repository's container (it needs Eigen and the toolkit's headers, which are
// fetched by CPM at the toolkit's configure time).  A maintainer adds this header and
// links libipcb200.so; see INTEGRATION.md.  All logic lives under the C ABI — this file only converts
// Eigen views to (pointer, rows, leading dimension) and fills the toolkit's public containers.
//
// Extension points used (reference src/ipc/):
//   * ipc::BroadPhase (broad_phase/broad_phase.hpp:19-133): CudaBroadPhase subclasses it, so it can be
//     passed wherever a BroadPhase* is accepted (Candidates::build, NormalCollisions::build,
//     ipc::compute_collision_free_stepsize, create_broad_phase).
//   * NormalCollisions::build, BarrierPotential::operator()/gradient/hessian and
//     compute_collision_free_stepsize are concrete (no vtable): ipc::cuda::* below are parallel entry
//     points with IDENTICAL signatures; the reference itself routes by `broad_phase->name()`
//     (ipc.cpp:64), and the same hook routes here (`name() == "CudaLBVH"`).
#pragma once
#if __has_include(<Eigen/Core>) && __has_include(<ipc/broad_phase/broad_phase.hpp>)

#include "ipcb200.hpp"

#include <ipc/broad_phase/broad_phase.hpp>
#include <ipc/candidates/candidates.hpp>
#include <ipc/ccd/additive_ccd.hpp>
#include <ipc/ccd/tight_inclusion_ccd.hpp>
#include <ipc/collision_mesh.hpp>
#include <ipc/collisions/normal/normal_collisions.hpp>
#include <ipc/potentials/barrier_potential.hpp>

#include <Eigen/Sparse>
#include <mutex>
#include <unordered_map>

namespace ipc::cuda {

inline ipcb200::MatrixXd view(Eigen::ConstRef<Eigen::MatrixXd> M) { return { M.data(), index_t(M.rows()), index_t(M.cols()), index_t(M.outerStride()) }; }
inline ipcb200::MatrixXi view(Eigen::ConstRef<Eigen::MatrixXi> M) { return { M.data(), index_t(M.rows()), index_t(M.cols()), index_t(M.outerStride()) }; }

/// One device context per ipc::CollisionMesh (the mesh is immutable after construction, collision_mesh.hpp:13), created
/// on first use.  The cache is guarded by a mutex and every entry remembers what it was built from (sizes and the address
/// of the rest positions): an ipc::CollisionMesh that was destroyed and whose address is reused by ANOTHER mesh gets a
/// fresh context instead of the stale one.  release_device_mesh() evicts an entry (call it from the owner of the mesh).
struct DeviceMeshEntry {
    std::unique_ptr<ipcb200::CollisionMesh> dm;
    std::unique_ptr<ipcb200::NormalCollisions> collisions; // the device-resident set of the last ipc::cuda::build on this mesh
    Eigen::Index nv = 0, ne = 0, nf = 0;
    const double* rest = nullptr;
};
inline std::mutex& device_mesh_mutex()
{
    static std::mutex m;
    return m;
}
inline std::unordered_map<const CollisionMesh*, DeviceMeshEntry>& device_mesh_cache()
{
    static std::unordered_map<const CollisionMesh*, DeviceMeshEntry> cache;
    return cache;
}
inline ipcb200::CollisionMesh& device_mesh(const CollisionMesh& mesh)
{
    std::lock_guard<std::mutex> lock(device_mesh_mutex());
    DeviceMeshEntry& e = device_mesh_cache()[&mesh];
    const bool same = e.dm && e.nv == mesh.rest_positions().rows() && e.ne == mesh.edges().rows() && e.nf == mesh.faces().rows()
        && e.rest == mesh.rest_positions().data();
    if (!same) {
        e.dm = std::make_unique<ipcb200::CollisionMesh>(view(mesh.rest_positions()), view(mesh.edges()), view(mesh.faces()));
        e.nv = mesh.rest_positions().rows(), e.ne = mesh.edges().rows(), e.nf = mesh.faces().rows(), e.rest = mesh.rest_positions().data();
    }
    return *e.dm;
}
inline ipcb200::NormalCollisions& device_collisions(const CollisionMesh& mesh)
{
    device_mesh(mesh);
    std::lock_guard<std::mutex> lock(device_mesh_mutex());
    DeviceMeshEntry& e = device_mesh_cache()[&mesh];
    if (!e.collisions) e.collisions = std::make_unique<ipcb200::NormalCollisions>();
    return *e.collisions;
}
inline void release_device_mesh(const CollisionMesh& mesh)
{
    std::lock_guard<std::mutex> lock(device_mesh_mutex());
    device_mesh_cache().erase(&mesh);
}

inline ipcb200::NarrowPhaseCCD convert(const NarrowPhaseCCD& ccd)
{
    if (auto* ti = dynamic_cast<const TightInclusionCCD*>(&ccd)) return ipcb200::TightInclusionCCD(ti->tolerance, ti->max_iterations, ti->conservative_rescaling);
    if (auto* ac = dynamic_cast<const AdditiveCCD*>(&ccd)) return ipcb200::AdditiveCCD(ac->max_iterations, ac->conservative_rescaling);
    throw std::runtime_error("CudaLBVH: only TightInclusionCCD and AdditiveCCD run on the device");
}

/// The CUDA LBVH behind the toolkit's BroadPhase interface.  build() needs the owning mesh because
/// the device tables are per mesh; Candidates::build passes mesh.edges()/faces(), which identify it.
class CudaBroadPhase : public BroadPhase {
public:
    explicit CudaBroadPhase(const CollisionMesh& mesh) : m_ipc_mesh(&mesh), m_mesh(&device_mesh(mesh)), m_bp(*m_mesh) { }
    std::string name() const override { return "CudaLBVH"; }

    /// The CollisionFilter of the toolkit is an opaque callable (collision_filter.hpp:30-111); what IS data — vertex
    /// patches and / or the number of dynamic vertices (collision_filter.hpp:113-143) — can be handed to the device here,
    /// after which `can_vertices_collide` is not consulted on the host any more.  Without this call the callable is
    /// applied as a host post-filter to every candidate the device returns (fill() below).
    void set_device_filter(const Eigen::VectorXi& patch_ids, const Eigen::Index n_dynamic = -1)
    {
        m_mesh->set_collision_filter(std::vector<ipcb200::index_t>(patch_ids.data(), patch_ids.data() + patch_ids.size()), ipcb200::index_t(n_dynamic));
        m_device_filter = true;
    }

    void build(Eigen::ConstRef<Eigen::MatrixXd> V, Eigen::ConstRef<Eigen::MatrixXi> E, Eigen::ConstRef<Eigen::MatrixXi> F, const double r = 0) override
    {
        check_same_mesh(V, E, F);
        dim = uint8_t(V.cols());
        m_bp.build(view(V), r);
    }
    void build(Eigen::ConstRef<Eigen::MatrixXd> V0, Eigen::ConstRef<Eigen::MatrixXd> V1, Eigen::ConstRef<Eigen::MatrixXi> E,
               Eigen::ConstRef<Eigen::MatrixXi> F, const double r = 0) override
    {
        check_same_mesh(V0, E, F);
        dim = uint8_t(V0.cols());
        m_bp.build(view(V0), view(V1), r);
    }
    void clear() override { BroadPhase::clear(); }

    void detect_vertex_vertex_candidates(std::vector<VertexVertexCandidate>& c) const override { fill(c, &ipcb200::CudaBroadPhase::detect_vertex_vertex_candidates); }
    void detect_edge_vertex_candidates(std::vector<EdgeVertexCandidate>& c) const override { fill(c, &ipcb200::CudaBroadPhase::detect_edge_vertex_candidates); }
    void detect_edge_edge_candidates(std::vector<EdgeEdgeCandidate>& c) const override { fill(c, &ipcb200::CudaBroadPhase::detect_edge_edge_candidates); }
    void detect_face_vertex_candidates(std::vector<FaceVertexCandidate>& c) const override { fill(c, &ipcb200::CudaBroadPhase::detect_face_vertex_candidates); }
    void detect_edge_face_candidates(std::vector<EdgeFaceCandidate>& c) const override { fill(c, &ipcb200::CudaBroadPhase::detect_edge_face_candidates); }
    void detect_face_face_candidates(std::vector<FaceFaceCandidate>& c) const override { fill(c, &ipcb200::CudaBroadPhase::detect_face_face_candidates); }

private:
    // vertices of a primitive of the candidate kinds: 1 (vertex), 2 (edge), 3 (face)
    void prim_vertices(int nv, index_t id, index_t* v) const
    {
        if (nv == 1) v[0] = id;
        else if (nv == 2) v[0] = m_ipc_mesh->edges()(id, 0), v[1] = m_ipc_mesh->edges()(id, 1);
        else v[0] = m_ipc_mesh->faces()(id, 0), v[1] = m_ipc_mesh->faces()(id, 1), v[2] = m_ipc_mesh->faces()(id, 2);
    }
    template <typename C> static constexpr std::pair<int, int> arity()
    {
        if constexpr (std::is_same_v<C, VertexVertexCandidate>) return { 1, 1 };
        else if constexpr (std::is_same_v<C, EdgeVertexCandidate>) return { 2, 1 };
        else if constexpr (std::is_same_v<C, EdgeEdgeCandidate>) return { 2, 2 };
        else if constexpr (std::is_same_v<C, FaceVertexCandidate>) return { 3, 1 };
        else if constexpr (std::is_same_v<C, EdgeFaceCandidate>) return { 2, 3 };
        else return { 3, 3 };
    }
    template <typename C, typename Fn> void fill(std::vector<C>& out, Fn fn) const
    {
        std::vector<ipcb200::Pair> pairs;
        (m_bp.*fn)(pairs);
        out.reserve(out.size() + pairs.size());
        // The device applies the share-a-vertex rule and, after set_device_filter(), the filter descriptor.  Otherwise the
        // toolkit's callable `can_vertices_collide` (broad_phase.hpp: public member, set by Candidates::build from
        // mesh.can_collide, candidates.cpp:61,147) is applied HERE with the rule of broad_phase.cpp:127-202: some pair of
        // vertices of the two primitives can collide.
        constexpr auto ar = arity<C>();
        for (const auto& p : pairs) {
            bool keep = m_device_filter;
            if (!keep) {
                index_t va[3], vb[3];
                prim_vertices(ar.first, p[0], va), prim_vertices(ar.second, p[1], vb);
                for (int i = 0; i < ar.first && !keep; i++)
                    for (int j = 0; j < ar.second && !keep; j++) keep = can_vertices_collide(size_t(va[i]), size_t(vb[j]));
            }
            if (keep) out.emplace_back(p[0], p[1]);
        }
    }
    void check_same_mesh(Eigen::ConstRef<Eigen::MatrixXd> V, Eigen::ConstRef<Eigen::MatrixXi> E, Eigen::ConstRef<Eigen::MatrixXi> F) const
    {
        if (size_t(V.rows()) != m_mesh->num_vertices() || size_t(E.rows()) != m_mesh->num_edges() || size_t(F.rows()) != m_mesh->num_faces())
            throw std::runtime_error("CudaLBVH was created for another CollisionMesh (codimensional sub-builds are done on the device)");
    }
    const CollisionMesh* m_ipc_mesh;
    ipcb200::CollisionMesh* m_mesh;
    mutable ipcb200::CudaBroadPhase m_bp;
    bool m_device_filter = false;
};

/// NormalCollisions::build(mesh, V, dhat, dmin, broad_phase) — normal_collisions.cpp:20-36.
/// The collision set stays resident on the device; the public host vectors of `collisions` are filled
/// from it (compatibility path).
inline void build(NormalCollisions& collisions, const CollisionMesh& mesh, Eigen::ConstRef<Eigen::MatrixXd> V, const double dhat, const double dmin = 0)
{
    auto& dm = device_mesh(mesh);
    auto& dc = device_collisions(mesh);
    dc.set_use_area_weighting(collisions.use_area_weighting());
    dc.build(dm, view(V), dhat, dmin);
    collisions.clear();
    const Eigen::SparseVector<double> no_gradient(V.size());
    auto vv = dc.records(IPCB_VV), ev = dc.records(IPCB_EV), ee = dc.records(IPCB_EE), fv = dc.records(IPCB_FV);
    for (size_t i = 0; i < vv.ids.size(); i++) collisions.vv_collisions.emplace_back(vv.ids[i][0], vv.ids[i][1], vv.weight[i], no_gradient);
    for (size_t i = 0; i < ev.ids.size(); i++) collisions.ev_collisions.emplace_back(ev.ids[i][0], ev.ids[i][1], ev.weight[i], no_gradient);
    for (size_t i = 0; i < ee.ids.size(); i++)
        collisions.ee_collisions.emplace_back(ee.ids[i][0], ee.ids[i][1], ee.eps_x[i], ee.weight[i], no_gradient, EdgeEdgeDistanceType(ee.dtype[i]));
    for (size_t i = 0; i < fv.ids.size(); i++) collisions.fv_collisions.emplace_back(fv.ids[i][0], fv.ids[i][1], fv.weight[i], no_gradient);
    for (size_t i = 0; i < collisions.size(); i++) collisions[i].dmin = dmin; // normal_collisions.cpp:154-157
}

/// BarrierPotential::operator()/gradient/hessian on the device-resident collision set of `mesh`
/// (potentials/potential.cpp:36-222); call ipc::cuda::build first.
inline double barrier_potential(const BarrierPotential& B, const CollisionMesh& mesh, Eigen::ConstRef<Eigen::MatrixXd> X)
{
    return ipcb200::BarrierPotential(B.dhat(), B.stiffness(), B.use_physical_barrier())(device_collisions(mesh), device_mesh(mesh), view(X));
}
inline Eigen::VectorXd barrier_potential_gradient(const BarrierPotential& B, const CollisionMesh& mesh, Eigen::ConstRef<Eigen::MatrixXd> X)
{
    const auto g = ipcb200::BarrierPotential(B.dhat(), B.stiffness(), B.use_physical_barrier()).gradient(device_collisions(mesh), device_mesh(mesh), view(X));
    return Eigen::Map<const Eigen::VectorXd>(g.data(), Eigen::Index(g.size()));
}
inline Eigen::SparseMatrix<double> barrier_potential_hessian(const BarrierPotential& B, const CollisionMesh& mesh, Eigen::ConstRef<Eigen::MatrixXd> X,
                                                             const PSDProjectionMethod project_hessian_to_psd = PSDProjectionMethod::NONE)
{
    const auto H = ipcb200::BarrierPotential(B.dhat(), B.stiffness(), B.use_physical_barrier())
                       .hessian(device_collisions(mesh), device_mesh(mesh), view(X), ipcb200::PSDProjectionMethod(int(project_hessian_to_psd)));
    return Eigen::Map<const Eigen::SparseMatrix<double>>(H.rows, H.cols, Eigen::Index(H.nonZeros()), H.outer.data(), H.inner.data(), H.values.data());
}

/// ipc::compute_collision_free_stepsize(mesh, V0, V1, min_distance, broad_phase, ccd) — ipc.cpp:45-101
inline double compute_collision_free_stepsize(const CollisionMesh& mesh, Eigen::ConstRef<Eigen::MatrixXd> V0, Eigen::ConstRef<Eigen::MatrixXd> V1,
                                              const double min_distance = 0.0, const NarrowPhaseCCD& ccd = DEFAULT_NARROW_PHASE_CCD)
{
    return ipcb200::compute_collision_free_stepsize(device_mesh(mesh), view(V0), view(V1), min_distance, convert(ccd));
}

/// ipc::has_intersections(mesh, vertices) — ipc.cpp:105-166 (3D)
inline bool has_intersections(const CollisionMesh& mesh, Eigen::ConstRef<Eigen::MatrixXd> V)
{
    return ipcb200::has_intersections(device_mesh(mesh), view(V));
}

} // namespace ipc::cuda
#endif
