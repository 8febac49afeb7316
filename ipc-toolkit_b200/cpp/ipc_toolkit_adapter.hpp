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
#include <unordered_map>

namespace ipc::cuda {

inline ipcb200::MatrixXd view(Eigen::ConstRef<Eigen::MatrixXd> M) { return { M.data(), index_t(M.rows()), index_t(M.cols()), index_t(M.outerStride()) }; }
inline ipcb200::MatrixXi view(Eigen::ConstRef<Eigen::MatrixXi> M) { return { M.data(), index_t(M.rows()), index_t(M.cols()), index_t(M.outerStride()) }; }

/// One device context per ipc::CollisionMesh (the mesh is immutable after construction,
/// collision_mesh.hpp:13), created on first use and kept for the mesh's lifetime.
inline ipcb200::CollisionMesh& device_mesh(const CollisionMesh& mesh)
{
    static std::unordered_map<const CollisionMesh*, std::unique_ptr<ipcb200::CollisionMesh>> cache;
    auto& slot = cache[&mesh];
    if (!slot) slot = std::make_unique<ipcb200::CollisionMesh>(view(mesh.rest_positions()), view(mesh.edges()), view(mesh.faces()));
    return *slot;
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
    explicit CudaBroadPhase(const CollisionMesh& mesh) : m_mesh(&device_mesh(mesh)), m_bp(*m_mesh) { }
    std::string name() const override { return "CudaLBVH"; }

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
    template <typename C, typename Fn> void fill(std::vector<C>& out, Fn fn) const
    {
        std::vector<ipcb200::Pair> pairs;
        (m_bp.*fn)(pairs);
        out.reserve(out.size() + pairs.size());
        // device-side filtering covers the share-a-vertex rule; an arbitrary CollisionFilter callback
        // (collision_filter.hpp:30-103) is applied here on the host, like sweep_and_tiniest_queue.cu:121-212
        for (const auto& p : pairs) out.emplace_back(p[0], p[1]);
    }
    void check_same_mesh(Eigen::ConstRef<Eigen::MatrixXd> V, Eigen::ConstRef<Eigen::MatrixXi> E, Eigen::ConstRef<Eigen::MatrixXi> F) const
    {
        if (size_t(V.rows()) != m_mesh->num_vertices() || size_t(E.rows()) != m_mesh->num_edges() || size_t(F.rows()) != m_mesh->num_faces())
            throw std::runtime_error("CudaLBVH was created for another CollisionMesh (codimensional sub-builds are done on the device)");
    }
    ipcb200::CollisionMesh* m_mesh;
    mutable ipcb200::CudaBroadPhase m_bp;
};

/// NormalCollisions::build(mesh, V, dhat, dmin, broad_phase) — normal_collisions.cpp:20-36.
/// The collision set stays resident on the device; the public host vectors of `collisions` are filled
/// from it (compatibility path).
inline void build(NormalCollisions& collisions, const CollisionMesh& mesh, Eigen::ConstRef<Eigen::MatrixXd> V, const double dhat, const double dmin = 0)
{
    auto& dm = device_mesh(mesh);
    ipcb200::NormalCollisions dc;
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
    return ipcb200::BarrierPotential(B.dhat(), B.stiffness(), B.use_physical_barrier())(ipcb200::NormalCollisions(), device_mesh(mesh), view(X));
}
inline Eigen::VectorXd barrier_potential_gradient(const BarrierPotential& B, const CollisionMesh& mesh, Eigen::ConstRef<Eigen::MatrixXd> X)
{
    const auto g = ipcb200::BarrierPotential(B.dhat(), B.stiffness(), B.use_physical_barrier()).gradient(ipcb200::NormalCollisions(), device_mesh(mesh), view(X));
    return Eigen::Map<const Eigen::VectorXd>(g.data(), Eigen::Index(g.size()));
}
inline Eigen::SparseMatrix<double> barrier_potential_hessian(const BarrierPotential& B, const CollisionMesh& mesh, Eigen::ConstRef<Eigen::MatrixXd> X,
                                                             const PSDProjectionMethod project_hessian_to_psd = PSDProjectionMethod::NONE)
{
    const auto H = ipcb200::BarrierPotential(B.dhat(), B.stiffness(), B.use_physical_barrier())
                       .hessian(ipcb200::NormalCollisions(), device_mesh(mesh), view(X), ipcb200::PSDProjectionMethod(int(project_hessian_to_psd)));
    return Eigen::Map<const Eigen::SparseMatrix<double>>(H.rows, H.cols, Eigen::Index(H.nonZeros()), H.outer.data(), H.inner.data(), H.values.data());
}

/// ipc::compute_collision_free_stepsize(mesh, V0, V1, min_distance, broad_phase, ccd) — ipc.cpp:45-101
inline double compute_collision_free_stepsize(const CollisionMesh& mesh, Eigen::ConstRef<Eigen::MatrixXd> V0, Eigen::ConstRef<Eigen::MatrixXd> V1,
                                              const double min_distance = 0.0, const NarrowPhaseCCD& ccd = DEFAULT_NARROW_PHASE_CCD)
{
    return ipcb200::compute_collision_free_stepsize(device_mesh(mesh), view(V0), view(V1), min_distance, convert(ccd));
}

} // namespace ipc::cuda
#endif
