// ipcb200_sharded.hpp — the multi-GPU contact step in C++: libipcb200 (C ABI) + NCCL, one rank per GPU.
//
// The C++ counterpart of ipc-toolkit_b200/sharded.py (DeviceShardedStep); the partitioning is SURVEY §8e / north_star:
//   1. broad phase + classification sharded by Morton range of query leaves (ipcb_ctx_set_shard): disjoint candidate
//      shards, no exchange;
//   2. the ranks' collision records — the reference's per-thread NormalCollisionsBuilder at rank granularity
//      (normal_collisions_builder.cpp:547-689) — travel in ONE ncclAllGather of [header | packed records] slots on a side
//      stream while the CCD half runs, and are merged on every rank (ipcb_collisions_append_packed_dev + _merge with
//      IPCB_MERGE_DISJOINT_SHARDS);
//   3. energy / gradient by collision range (ncclAllReduce sum), Hessian by balanced row block (no collective:
//      per-rank row-block CSR), step size by ncclAllReduce min.
// With a second context on the same mesh (`ccd_mesh`) the swept broad phase + CCD run on their own streams, issued by
// their own host thread, beside the build + potential half.
//
// Device-resident: positions, energy, gradient and step size are device pointers; nothing but counts crosses PCIe.
// Header-only; needs <cuda_runtime.h> and <nccl.h> (link -lipcb200 -lnccl -lcudart).
#pragma once
#include "ipcb200.hpp"

#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <array>
#include <cstring>
#include <future>

namespace ipcb200 {

inline void cuda_check(cudaError_t e, const char* what)
{
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
inline void nccl_check(ncclResult_t r, const char* what)
{
    if (r != ncclSuccess) throw std::runtime_error(std::string(what) + ": " + ncclGetErrorString(r));
}

class ShardedContactStep {
public:
    struct Result {
        int64_t nnz = 0;                                 // entries of this rank's row block of the Hessian
        std::array<int64_t, 4> collisions { 0, 0, 0, 0 }; // the merged set (identical on every rank)
        std::array<int64_t, 4> shard { 0, 0, 0, 0 };      // what this rank's candidate shard produced
        int32_t row_begin = 0, row_end = 0;               // owned vertices: DOF rows [3 row_begin, 3 row_end)
    };

    /// mesh: the rank's context (already on the rank's device); ccd_mesh: optional second context on the same mesh for the
    /// CCD lane; comm: the rank's NCCL communicator (nullptr for a single rank).
    ShardedContactStep(const CollisionMesh& mesh, const CollisionMesh* ccd_mesh, int rank, int world, ncclComm_t comm)
        : m_mesh(&mesh), m_ccd(ccd_mesh), m_rank(rank), m_world(world), m_comm(comm)
    {
        if (world > 1 && !comm) throw std::invalid_argument("a communicator is needed for more than one rank");
        check(ipcb_ctx_set_shard(mesh.ctx(), rank, world));
        if (ccd_mesh) check(ipcb_ctx_set_shard(ccd_mesh->ctx(), rank, world));
        m_stream = static_cast<cudaStream_t>(ipcb_ctx_stream(mesh.ctx()));
        cuda_check(cudaStreamCreateWithFlags(&m_side, cudaStreamNonBlocking), "cudaStreamCreate");
        for (cudaEvent_t* e : { &m_ev_packed, &m_ev_gathered, &m_ev_start, &m_ev_ccd })
            cuda_check(cudaEventCreateWithFlags(e, cudaEventDisableTiming), "cudaEventCreate");
        cuda_check(cudaMallocHost(reinterpret_cast<void**>(&m_headers), sizeof(int64_t) * 4 * size_t(world + 1)), "cudaMallocHost");
    }
    ~ShardedContactStep()
    {
        cudaStreamSynchronize(m_side);
        cudaFree(m_send), cudaFree(m_recv), cudaFreeHost(m_headers);
        for (cudaEvent_t e : { m_ev_packed, m_ev_gathered, m_ev_start, m_ev_ccd }) cudaEventDestroy(e);
        cudaStreamDestroy(m_side);
    }
    ShardedContactStep(const ShardedContactStep&) = delete;
    ShardedContactStep& operator=(const ShardedContactStep&) = delete;

    /// One contact step.  dV0 / dV1: device, column-major nV x 3 (leading dimension ld); d_energy (1), d_grad (3 nV),
    /// d_step (1): device outputs, all-reduced over the ranks.  The rank's Hessian row block stays resident in the
    /// context (ipcb_barrier_hessian_dev_ptrs / _fetch).
    Result step(const double* dV0, const double* dV1, int32_t ld, const BarrierPotential& B, PSDProjectionMethod psd, double* d_energy,
                double* d_grad, double* d_step, double dmin = 0.0, double min_distance = 0.0,
                const NarrowPhaseCCD& ccd = DEFAULT_NARROW_PHASE_CCD)
    {
        ipcb_ctx* ctx = m_mesh->ctx();
        const int32_t nV = int32_t(m_mesh->num_vertices());
        const ipcb_barrier_params bp { B.dhat(), B.stiffness(), B.use_physical_barrier() ? 1 : 0 };
        Result out;
        // ---- CCD lane (second context, own host thread) or inline after the exchange has started
        std::future<void> lane;
        auto ccd_half = [&](ipcb_ctx* c) { check(ipcb_ccd_stepsize_dev(c, dV0, dV1, ld, min_distance, &ccd.params, d_step)); };
        if (m_ccd) {
            cudaStream_t sb = static_cast<cudaStream_t>(ipcb_ctx_stream(m_ccd->ctx()));
            cuda_check(cudaEventRecord(m_ev_start, m_stream), "cudaEventRecord");
            cuda_check(cudaStreamWaitEvent(sb, m_ev_start, 0), "cudaStreamWaitEvent"); // the caller's inputs are ready
            lane = std::async(std::launch::async, [&, this] { ccd_half(m_ccd->ctx()); });
        }
        int64_t counts[4];
        check(ipcb_collisions_build_dev(ctx, dV0, ld, B.dhat(), dmin, 0, counts));
        std::copy(counts, counts + 4, out.shard.begin());
        if (m_world > 1) start_exchange(ctx, counts);
        if (!m_ccd) ccd_half(ctx);
        if (m_world > 1) {
            finish_exchange(ctx, dmin, counts);
            std::vector<int32_t> bounds(size_t(m_world) + 1);
            check(ipcb_hessian_balanced_row_blocks(ctx, m_world, bounds.data()));
            out.row_begin = bounds[size_t(m_rank)], out.row_end = bounds[size_t(m_rank) + 1];
            check(ipcb_ctx_set_collision_range(ctx, m_rank, m_world));
            check(ipcb_ctx_set_row_block(ctx, out.row_begin, out.row_end));
        } else {
            out.row_begin = 0, out.row_end = nV;
        }
        std::copy(counts, counts + 4, out.collisions.begin());
        check(ipcb_barrier_energy_dev(ctx, dV0, ld, &bp, d_energy));
        check(ipcb_barrier_gradient_dev(ctx, dV0, ld, &bp, d_grad));
        check(ipcb_barrier_hessian_dev(ctx, dV0, ld, &bp, int(psd), &out.nnz));
        if (m_ccd) { // join the lanes: this context's stream continues after the step size has been written
            lane.get();
            cudaStream_t sb = static_cast<cudaStream_t>(ipcb_ctx_stream(m_ccd->ctx()));
            cuda_check(cudaEventRecord(m_ev_ccd, sb), "cudaEventRecord");
            cuda_check(cudaStreamWaitEvent(m_stream, m_ev_ccd, 0), "cudaStreamWaitEvent");
        }
        if (m_world > 1) { // sum / sum / min all-reduces on the context's stream; the Hessian needs no collective
            nccl_check(ncclGroupStart(), "ncclGroupStart");
            nccl_check(ncclAllReduce(d_energy, d_energy, 1, ncclDouble, ncclSum, m_comm, m_stream), "ncclAllReduce");
            nccl_check(ncclAllReduce(d_grad, d_grad, 3 * size_t(nV), ncclDouble, ncclSum, m_comm, m_stream), "ncclAllReduce");
            nccl_check(ncclAllReduce(d_step, d_step, 1, ncclDouble, ncclMin, m_comm, m_stream), "ncclAllReduce");
            nccl_check(ncclGroupEnd(), "ncclGroupEnd");
            check(ipcb_ctx_set_collision_range(ctx, 0, 1)); // the context is back to "whole set" for other callers
        }
        cuda_check(cudaStreamSynchronize(m_stream), "cudaStreamSynchronize");
        return out;
    }
    /// back to the whole matrix for later un-sharded calls on the context
    void reset_row_block() const { check(ipcb_ctx_set_row_block(m_mesh->ctx(), 0, -1)); }

private:
    static constexpr int64_t HEADER = 32; // bytes in front of every rank's slot: its four record counts
    static int64_t packed_bytes(const int64_t n[4]) { return 16 * (n[0] + n[1] + n[3]) + 24 * n[2] + (n[2] + 7) / 8 * 8; }

    void resize(int64_t need)
    {
        cuda_check(cudaStreamSynchronize(m_side), "cudaStreamSynchronize");
        cudaFree(m_send), cudaFree(m_recv);
        m_cap = (need + need / 4 + 4096) / 16 * 16;
        cuda_check(cudaMalloc(&m_send, size_t(m_cap)), "cudaMalloc");
        cuda_check(cudaMalloc(&m_recv, size_t(m_cap) * size_t(m_world)), "cudaMalloc");
    }
    // [header | packed records] of every rank in ONE all-gather of m_cap bytes per rank on the side stream
    void gather_once(ipcb_ctx* ctx, const int64_t counts[4])
    {
        int64_t* mine = m_headers + 4 * size_t(m_world); // pinned staging of this rank's header
        std::copy(counts, counts + 4, mine);
        cuda_check(cudaMemcpyAsync(m_send, mine, HEADER, cudaMemcpyHostToDevice, m_stream), "cudaMemcpyAsync");
        const int64_t need = HEADER + packed_bytes(counts);
        if (need <= m_cap && counts[0] + counts[1] + counts[2] + counts[3] > 0) {
            int64_t bytes = 0;
            check(ipcb_collisions_pack_dev(ctx, static_cast<char*>(m_send) + HEADER, m_cap - HEADER, &bytes));
        }
        cuda_check(cudaEventRecord(m_ev_packed, m_stream), "cudaEventRecord");
        cuda_check(cudaStreamWaitEvent(m_side, m_ev_packed, 0), "cudaStreamWaitEvent");
        nccl_check(ncclAllGather(m_send, m_recv, size_t(m_cap), ncclChar, m_comm, m_side), "ncclAllGather");
        cuda_check(cudaMemcpy2DAsync(m_headers, HEADER, m_recv, size_t(m_cap), HEADER, size_t(m_world), cudaMemcpyDeviceToHost, m_side),
                   "cudaMemcpy2DAsync");
        cuda_check(cudaEventRecord(m_ev_gathered, m_side), "cudaEventRecord");
    }
    void start_exchange(ipcb_ctx* ctx, const int64_t counts[4])
    {
        if (m_cap == 0) { // first step: the ranks agree on a slot size (one max all-reduce of 8 bytes)
            int64_t want = 2 * (HEADER + packed_bytes(counts)) + (1 << 20);
            int64_t* d = nullptr;
            cuda_check(cudaMalloc(reinterpret_cast<void**>(&d), sizeof(int64_t)), "cudaMalloc");
            cuda_check(cudaMemcpyAsync(d, &want, sizeof want, cudaMemcpyHostToDevice, m_side), "cudaMemcpyAsync");
            nccl_check(ncclAllReduce(d, d, 1, ncclInt64, ncclMax, m_comm, m_side), "ncclAllReduce");
            cuda_check(cudaMemcpyAsync(&want, d, sizeof want, cudaMemcpyDeviceToHost, m_side), "cudaMemcpyAsync");
            cuda_check(cudaStreamSynchronize(m_side), "cudaStreamSynchronize");
            cudaFree(d);
            resize(want);
        }
        gather_once(ctx, counts);
    }
    // NormalCollisionsBuilder::merge over the ranks' records
    void finish_exchange(ipcb_ctx* ctx, double dmin, int64_t counts[4])
    {
        const int64_t mine[4] = { counts[0], counts[1], counts[2], counts[3] };
        for (int attempt = 0;; attempt++) {
            cuda_check(cudaEventSynchronize(m_ev_gathered), "cudaEventSynchronize"); // the headers are on the host
            int64_t need = 0;
            for (int r = 0; r < m_world; r++) need = std::max(need, HEADER + packed_bytes(m_headers + 4 * size_t(r)));
            if (need <= m_cap) break;
            if (attempt >= 2) throw std::runtime_error("collision exchange: slot overflow persisted");
            resize(need); // identical decision on every rank: repeat the exchange with room
            gather_once(ctx, mine);
        }
        cuda_check(cudaStreamWaitEvent(m_stream, m_ev_gathered, 0), "cudaStreamWaitEvent");
        check(ipcb_collisions_clear(ctx));
        for (int r = 0; r < m_world; r++) {
            const int64_t* c = m_headers + 4 * size_t(r);
            if (c[0] + c[1] + c[2] + c[3] == 0) continue;
            check(ipcb_collisions_append_packed_dev(ctx, static_cast<char*>(m_recv) + size_t(r) * size_t(m_cap) + HEADER, c));
        }
        check(ipcb_collisions_merge(ctx, dmin, IPCB_MERGE_DISJOINT_SHARDS, counts));
    }

    const CollisionMesh* m_mesh;
    const CollisionMesh* m_ccd;
    int m_rank, m_world;
    ncclComm_t m_comm;
    cudaStream_t m_stream = nullptr, m_side = nullptr;
    cudaEvent_t m_ev_packed = nullptr, m_ev_gathered = nullptr, m_ev_start = nullptr, m_ev_ccd = nullptr;
    void *m_send = nullptr, *m_recv = nullptr;
    int64_t m_cap = 0;
    int64_t* m_headers = nullptr; // pinned: world headers + this rank's staging header
};

} // namespace ipcb200
