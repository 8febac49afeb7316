// The C++ multi-GPU contact step (ipc-toolkit_b200/cpp/ipcb200_sharded.hpp: libipcb200 + NCCL) against the single-GPU
// step: one process, one host thread per GPU (ncclCommInitAll), a stack of jittered cloth sheets built here.
//   energy / gradient / step size: equal after the all-reduces;   collision sets: identical counts;
//   Hessian: the ranks' row blocks tile the single-GPU matrix (same entries, values to 1e-12).
// usage: test_sharded_step [world] [two_lanes=1]     prints "sharded step ok" on success
#include "../../ipc-toolkit_b200/cpp/ipcb200_sharded.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>

using namespace ipcb200;

struct Scene {
    std::vector<double> V0, V1; // column-major nV x 3
    std::vector<index_t> E, F;  // column-major
    index_t nV = 0, nE = 0, nF = 0;
    double dhat = 1e-3;
};
// `layers` n x n sheets, 0.5 dhat apart, every other one shifted by a third of a cell; deterministic jitter
static Scene make_scene(int layers, int n)
{
    Scene s;
    const int row = n + 1, per = row * row;
    s.nV = layers * per;
    std::vector<std::array<double, 3>> P(size_t(s.nV)), Q;
    unsigned state = 12345u;
    auto rnd = [&] { state = state * 1664525u + 1013904223u; return double(state >> 8) / double(1 << 24) - 0.5; };
    const double h = 1.0 / n;
    for (int k = 0; k < layers; k++)
        for (int i = 0; i < row; i++)
            for (int j = 0; j < row; j++)
                P[size_t(k * per + i * row + j)] = { i * h + (k % 2) * h / 3 + 1e-3 * h * rnd(), j * h + (k % 3) * h / 5 + 1e-3 * h * rnd(),
                                                     k * 0.5 * s.dhat + 1e-2 * s.dhat * rnd() };
    Q = P;
    for (int k = 0; k < layers; k++)
        for (int v = 0; v < per; v++) Q[size_t(k * per + v)][2] -= (k - 0.5 * (layers - 1)) * 0.65 * s.dhat;
    std::vector<std::array<index_t, 3>> F;
    std::vector<std::array<index_t, 2>> E;
    for (int k = 0; k < layers; k++)
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                const index_t a = k * per + i * row + j, b = a + row, c = b + 1, d = a + 1;
                F.push_back({ a, b, c }), F.push_back({ a, c, d });
                E.push_back({ a, b }), E.push_back({ a, d }), E.push_back({ a, c });
                if (i == n - 1) E.push_back({ b, c });
                if (j == n - 1) E.push_back({ d, c });
            }
    s.nE = index_t(E.size()), s.nF = index_t(F.size());
    s.V0.resize(3 * size_t(s.nV)), s.V1.resize(3 * size_t(s.nV));
    for (index_t v = 0; v < s.nV; v++)
        for (int c = 0; c < 3; c++) s.V0[size_t(c) * s.nV + v] = P[size_t(v)][c], s.V1[size_t(c) * s.nV + v] = Q[size_t(v)][c];
    s.E.resize(2 * size_t(s.nE)), s.F.resize(3 * size_t(s.nF));
    for (index_t e = 0; e < s.nE; e++)
        for (int c = 0; c < 2; c++) s.E[size_t(c) * s.nE + e] = E[size_t(e)][c];
    for (index_t f = 0; f < s.nF; f++)
        for (int c = 0; c < 3; c++) s.F[size_t(c) * s.nF + f] = F[size_t(f)][c];
    return s;
}

struct RankOut {
    ShardedContactStep::Result r;
    double energy = 0, step = 0;
    std::vector<double> grad;
    SparseMatrix H;
    std::string error;
};

static void run_rank(const Scene& s, int device, int rank, int world, ncclComm_t comm, bool two_lanes, RankOut& out)
{
    try {
        cuda_check(cudaSetDevice(device), "cudaSetDevice");
        CollisionMesh mesh(MatrixXd(s.V0.data(), s.nV, 3), MatrixXi(s.E.data(), s.nE, 2), MatrixXi(s.F.data(), s.nF, 3), device);
        std::unique_ptr<CollisionMesh> lane;
        if (two_lanes) lane.reset(new CollisionMesh(MatrixXd(s.V0.data(), s.nV, 3), MatrixXi(s.E.data(), s.nE, 2), MatrixXi(s.F.data(), s.nF, 3), device));
        double *d0, *d1, *dE, *dG, *dS;
        const size_t nb = sizeof(double) * 3 * size_t(s.nV);
        cuda_check(cudaMalloc(reinterpret_cast<void**>(&d0), nb), "cudaMalloc");
        cuda_check(cudaMalloc(reinterpret_cast<void**>(&d1), nb), "cudaMalloc");
        cuda_check(cudaMalloc(reinterpret_cast<void**>(&dG), nb), "cudaMalloc");
        cuda_check(cudaMalloc(reinterpret_cast<void**>(&dE), 16), "cudaMalloc");
        dS = dE + 1;
        cuda_check(cudaMemcpy(d0, s.V0.data(), nb, cudaMemcpyHostToDevice), "cudaMemcpy");
        cuda_check(cudaMemcpy(d1, s.V1.data(), nb, cudaMemcpyHostToDevice), "cudaMemcpy");
        {
            ShardedContactStep stepper(mesh, lane.get(), rank, world, comm);
            const BarrierPotential B(s.dhat, 1.0);
            const AdditiveCCD accd; // the same arithmetic per candidate whatever the shard: the minimum is exact
            for (int it = 0; it < 2; it++) // twice: the exchange buffers are reused
                out.r = stepper.step(d0, d1, s.nV, B, PSDProjectionMethod::CLAMP, dE, dG, dS, 0.0, 0.0, accd);
            out.grad.resize(3 * size_t(s.nV));
            cuda_check(cudaMemcpy(out.grad.data(), dG, nb, cudaMemcpyDeviceToHost), "cudaMemcpy");
            cuda_check(cudaMemcpy(&out.energy, dE, 8, cudaMemcpyDeviceToHost), "cudaMemcpy");
            cuda_check(cudaMemcpy(&out.step, dS, 8, cudaMemcpyDeviceToHost), "cudaMemcpy");
            out.H.rows = out.H.cols = index_t(3 * s.nV);
            out.H.outer.resize(3 * size_t(s.nV) + 1), out.H.inner.resize(size_t(out.r.nnz)), out.H.values.resize(size_t(out.r.nnz));
            check(ipcb_barrier_hessian_fetch(mesh.ctx(), out.H.outer.data(), out.H.inner.data(), out.H.values.data()));
        }
        cudaFree(d0), cudaFree(d1), cudaFree(dG), cudaFree(dE);
    } catch (const std::exception& e) {
        out.error = e.what();
    }
}

int main(int argc, char** argv)
{
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    int world = argc > 1 ? std::atoi(argv[1]) : std::min(ndev, 4);
    const bool two_lanes = argc > 2 ? std::atoi(argv[2]) != 0 : true;
    if (world < 2 || ndev < world) {
        std::printf("needs %d GPUs, found %d\n", std::max(world, 2), ndev);
        return 77;
    }
    const Scene s = make_scene(4, 40);
    // ---- single GPU reference
    RankOut one;
    run_rank(s, 0, 0, 1, nullptr, false, one);
    if (!one.error.empty()) return std::printf("single rank failed: %s\n", one.error.c_str()), 1;
    // ---- one thread per GPU
    std::vector<ncclComm_t> comms(size_t(world), nullptr);
    std::vector<int> devs(size_t(world), 0);
    for (int r = 0; r < world; r++) devs[size_t(r)] = r;
    if (ncclCommInitAll(comms.data(), world, devs.data()) != ncclSuccess) return std::printf("ncclCommInitAll failed\n"), 1;
    std::vector<RankOut> outs;
    outs.resize(size_t(world));
    std::vector<std::thread> threads;
    for (int r = 0; r < world; r++) threads.emplace_back(run_rank, std::cref(s), r, r, world, comms[size_t(r)], two_lanes, std::ref(outs[size_t(r)]));
    for (auto& t : threads) t.join();
    for (auto c : comms) ncclCommDestroy(c);
    int bad = 0;
#define EXPECT(c)                                                             \
    do {                                                                      \
        if (!(c)) std::printf("EXPECT failed: %s (line %d)\n", #c, __LINE__), bad++; \
    } while (0)
    int64_t nnz = 0, next_row = 0;
    double gn = 0;
    for (double g : one.grad) gn += g * g;
    EXPECT(one.r.collisions[2] > 1000 && one.r.nnz > 10000 && one.step > 0 && one.step < 1);
    for (int r = 0; r < world; r++) {
        const RankOut& o = outs[size_t(r)];
        if (!o.error.empty()) {
            std::printf("rank %d failed: %s\n", r, o.error.c_str());
            return 1;
        }
        EXPECT(o.r.collisions == one.r.collisions);
        EXPECT(std::abs(o.energy - one.energy) <= 1e-12 * std::abs(one.energy));
        EXPECT(o.step == one.step);
        double dg = 0;
        for (size_t i = 0; i < o.grad.size(); i++) dg += (o.grad[i] - one.grad[i]) * (o.grad[i] - one.grad[i]);
        EXPECT(dg <= 1e-24 * gn);
        EXPECT(o.r.row_begin == next_row && o.r.row_end > o.r.row_begin);
        next_row = o.r.row_end;
        nnz += o.r.nnz;
        // the rank's matrix == the single-GPU matrix restricted to its columns
        const index_t lo = 3 * o.r.row_begin, hi = 3 * o.r.row_end;
        EXPECT(o.H.outer[size_t(lo)] == 0 && o.H.outer[size_t(hi)] == index_t(o.r.nnz) && o.H.outer.back() == index_t(o.r.nnz));
        double err = 0, ref = 0;
        bool pattern = true;
        for (index_t c = lo; c < hi && pattern; c++) {
            const index_t a = o.H.outer[size_t(c)], b = o.H.outer[size_t(c) + 1], a1 = one.H.outer[size_t(c)], b1 = one.H.outer[size_t(c) + 1];
            pattern = b - a == b1 - a1;
            for (index_t k = 0; k < b - a && pattern; k++) {
                pattern = o.H.inner[size_t(a + k)] == one.H.inner[size_t(a1 + k)];
                const double d = o.H.values[size_t(a + k)] - one.H.values[size_t(a1 + k)];
                err += d * d, ref += one.H.values[size_t(a1 + k)] * one.H.values[size_t(a1 + k)];
            }
        }
        EXPECT(pattern);
        EXPECT(err <= 1e-24 * ref);
    }
    EXPECT(next_row == s.nV && nnz == one.r.nnz);
    if (bad) return 1;
    std::printf("sharded step ok: world %d, lanes %d, collisions [%lld %lld %lld %lld], nnz %lld, step %.6f\n", world, two_lanes ? 2 : 1,
                (long long)one.r.collisions[0], (long long)one.r.collisions[1], (long long)one.r.collisions[2], (long long)one.r.collisions[3],
                (long long)one.r.nnz, one.step);
    return 0;
}
