// C++ host-mirror check (compiled with g++, linked against libipcb200.so): the reference's
// "Codim. vertex-vertex collisions" known-answer test (tests/src/tests/collisions/test_normal_collisions.cpp:14-107)
// written against ipcb200.hpp the way it is written against ipc-toolkit.
#include "../../ipc-toolkit_b200/cpp/ipcb200.hpp"
#include <cmath>
#include <cstdio>

using namespace ipcb200;

#define CHECK(c)                                                     \
    do {                                                             \
        if (!(c)) {                                                  \
            std::printf("CHECK failed: %s (line %d)\n", #c, __LINE__); \
            return 1;                                                \
        }                                                            \
    } while (0)

int main()
{
    constexpr double thickness = 0.4, min_distance = 2 * thickness, dhat = 0.25;
    // 8 cube corners centred at the origin, column-major 8 x 3
    double V[24], V1[24];
    for (int i = 0; i < 8; i++) {
        V[i] = ((i >> 2) & 1) - 0.5, V[8 + i] = ((i >> 1) & 1) - 0.5, V[16 + i] = (i & 1) - 0.5;
        V1[i] = V[i], V1[8 + i] = 0.5 * V[8 + i], V1[16 + i] = V[16 + i];
    }
    try {
        CollisionMesh mesh(MatrixXd(V, 8, 3));
        CHECK(mesh.num_vertices() == 8 && mesh.num_codim_vertices() == 8 && mesh.num_edges() == 0 && mesh.num_faces() == 0);

        Candidates candidates;
        candidates.build(mesh, MatrixXd(V, 8, 3), MatrixXd(V1, 8, 3), thickness);
        CHECK(!candidates.empty() && candidates.vv_candidates().size() == candidates.size());
        CHECK(!candidates.is_step_collision_free(mesh, MatrixXd(V, 8, 3), MatrixXd(V1, 8, 3), min_distance));
        const double expected_toi = (1 - (min_distance + 1e-4)) / 2.0 / 0.25;
        const double toi = candidates.compute_collision_free_stepsize(mesh, MatrixXd(V, 8, 3), MatrixXd(V1, 8, 3), min_distance);
        CHECK(std::abs(toi - expected_toi) <= 1.2e-5 * expected_toi);

        NormalCollisions collisions;
        collisions.build(mesh, MatrixXd(V, 8, 3), dhat, min_distance);
        CHECK(collisions.size() == 12 && collisions.count(IPCB_VV) == 12);
        BarrierPotential B(dhat, 1.0);
        CHECK(B(collisions, mesh, MatrixXd(V, 8, 3)) > 0.0);
        const std::vector<double> g = B.gradient(collisions, mesh, MatrixXd(V, 8, 3));
        for (int i = 0; i < 8; i++) { // the force is radial
            const double f[3] = { -g[3 * i], -g[3 * i + 1], -g[3 * i + 2] }, x[3] = { V[i], V[8 + i], V[16 + i] };
            const double nf = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]), nx = std::sqrt(0.75);
            for (int k = 0; k < 3; k++) CHECK(std::abs(f[k] / nf - x[k] / nx) < 1e-9);
        }
        { // CollisionSetType::IMPROVED_MAX_APPROX on the cube of points (test_normal_collisions.cpp:69-107): 12 collisions; and the
          // same set from a deferred build finished with the builder's own pairs (one builder = slice 0 of 1)
            NormalCollisions ima, two;
            ima.set_collision_set_type(NormalCollisions::CollisionSetType::IMPROVED_MAX_APPROX);
            ima.build(mesh, MatrixXd(V, 8, 3), dhat, min_distance);
            CHECK(ima.size() == 12);
            two.set_collision_set_type(NormalCollisions::CollisionSetType::IMPROVED_MAX_APPROX);
            two.build(mesh, MatrixXd(V, 8, 3), dhat, min_distance, true);
            two.apply_corrections(two.correction_keys(), 0, 1);
            CHECK(two.size() == ima.size());
        }
        const SparseMatrix H = B.hessian(collisions, mesh, MatrixXd(V, 8, 3), PSDProjectionMethod::CLAMP);
        CHECK(H.rows == 24 && H.nonZeros() > 0 && H.outer.back() == index_t(H.nonZeros()));
        CHECK(compute_collision_free_stepsize(mesh, MatrixXd(V, 8, 3), MatrixXd(V, 8, 3)) == 1.0);
        // error behaviour: a face whose edge is missing (collision_mesh.cpp:537)
        const index_t F[3] = { 0, 1, 2 }, E[2] = { 0, 1 };
        bool threw = false;
        try {
            CollisionMesh bad(MatrixXd(V, 8, 3), MatrixXi(E, 1, 2), MatrixXi(F, 1, 3));
        } catch (const std::runtime_error& e) {
            threw = std::string(e.what()).find("Unable to find edge!") != std::string::npos;
        }
        CHECK(threw);
    } catch (const std::exception& e) {
        std::printf("exception: %s\n", e.what());
        return 2;
    }
    std::printf("host mirror ok\n");
    return 0;
}
