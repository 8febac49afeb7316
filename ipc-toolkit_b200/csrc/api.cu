// api.cu — extern "C" entry points of libipcb200.so (include/ipcb200.h).
// Host variants stage caller buffers into the context's device buffers and call
// the device path; there is no CPU implementation of any stage in this library.
#include "ctx.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_map>

using namespace ipcb;

namespace {
thread_local std::string g_error;

template <typename F> int guarded(F&& f)
{
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    } catch (...) {
        g_error = "unknown error";
        return 1;
    }
}

void use_device(ipcb_ctx* ctx) { IPCB_CUDA(cudaSetDevice(ctx->device)); }

// ---- roofline denominators (measure_fp64_peak / measure_copy_bandwidth)
constexpr int PEAK_ITERS = 2048, PEAK_CHAINS = 8;
__global__ void __launch_bounds__(256) k_fp64_peak(double seed, double* out)
{
    double x[PEAK_CHAINS];
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; k++) x[k] = seed + double(threadIdx.x + k);
    const double a = 1.0 + 1e-9 * seed, b = 1e-9;
#pragma unroll 1
    for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
        for (int k = 0; k < PEAK_CHAINS; k++) x[k] = fma(x[k], a, b); // 8 independent FMA chains
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; k++) s += x[k];
    if (s == 12345.678) out[0] = s; // never true: keeps the chains alive
}
__global__ void __launch_bounds__(256) k_copy16(size_t n, const uint4* __restrict__ in, uint4* __restrict__ out)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) out[i] = in[i];
}

template <typename T> void upload(ipcb_ctx* ctx, Buf<T>& dst, const std::vector<T>& src)
{
    dst.reserve(std::max<size_t>(src.size(), 1));
    if (!src.empty()) IPCB_CUDA(cudaMemcpyAsync(dst.p, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice, ctx->stream));
}

void begin_call(ipcb_ctx* ctx)
{
    use_device(ctx);
    ctx->stage_ms.clear();
}

// stage host positions and convert to the AoS device layout
void stage_positions(ipcb_ctx* ctx, const double* V, int ld, bool second)
{
    Stage st(ctx, second ? "h2d_positions_t1" : "h2d_positions");
    Buf<double>& stage = second ? ctx->stageB : ctx->stageA;
    upload_positions(ctx, V, ld, stage);
    convert_positions(ctx, stage.p, ctx->nV, second ? ctx->X1 : ctx->X0);
}

void fill_counts(const PairList* lists, int64_t counts[4])
{
    for (int k = 0; k < 4; k++) counts[k] = lists[k].count;
}
void coll_counts(const ipcb_ctx* ctx, int64_t counts[4])
{
    for (int k = 0; k < 4; k++) counts[k] = ctx->coll[k].count;
}
void require_collisions(const ipcb_ctx* ctx)
{
    if (!ctx->coll_valid) throw Error("no collision set has been built on this context");
}
} // namespace

extern "C" {

const char* ipcb_last_error(void) { return g_error.c_str(); }
const char* ipcb_backend_name(void) { return "cuda-sm100a"; }

int ipcb_ctx_create(int device, ipcb_ctx** out)
{
    return guarded([&] {
        int count = 0;
        const cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            throw Error(std::string("ipcb200 needs a CUDA device and has no CPU fallback (cudaGetDeviceCount: ")
                        + cudaGetErrorString(e) + ")");
        if (device < 0 || device >= count) throw Error("invalid CUDA device index");
        IPCB_CUDA(cudaSetDevice(device));
        ipcb_ctx* ctx = new ipcb_ctx();
        ctx->device = device;
        IPCB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        const bool single_stream = getenv("IPCB_SINGLE_STREAM") != nullptr; // A/B switch: no intra-call concurrency
        for (int k = 0; k < ipcb_ctx::NAUX; k++) {
            if (single_stream) ctx->aux[k] = ctx->stream;
            else IPCB_CUDA(cudaStreamCreateWithFlags(&ctx->aux[k], cudaStreamNonBlocking));
            IPCB_CUDA(cudaEventCreateWithFlags(&ctx->ev_join[k], cudaEventDisableTiming));
        }
        IPCB_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        ctx->pinned.init(32);
        ctx->dCounters.reserve(32);
        IPCB_CUDA(cudaMemsetAsync(ctx->dCounters.p, 0, 32 * sizeof(unsigned long long), ctx->stream));
        ctx->dScalar.reserve(64);
        *out = ctx;
    });
}
void ipcb_ctx_destroy(ipcb_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int k = 0; k < ipcb_ctx::NAUX; k++) {
        cudaStreamSynchronize(ctx->aux[k]);
        if (ctx->aux[k] != ctx->stream) cudaStreamDestroy(ctx->aux[k]);
        cudaEventDestroy(ctx->ev_join[k]);
    }
    cudaEventDestroy(ctx->ev_fork);
    ti_work_free(ctx->ti_work);
    ctx->ti_work = nullptr;
    cudaStream_t s = ctx->stream;
    delete ctx; // frees the device buffers (cudaFree synchronises)
    cudaStreamDestroy(s);
}
void* ipcb_ctx_stream(ipcb_ctx* ctx) { return ctx->stream; }

int ipcb_ctx_set_shard(ipcb_ctx* ctx, int32_t rank, int32_t world)
{
    return guarded([&] {
        if (world < 1 || rank < 0 || rank >= world) throw Error("bad shard");
        ctx->shard_rank = rank;
        ctx->shard_world = world;
    });
}
int ipcb_ctx_set_broad_phase_method(ipcb_ctx* ctx, int32_t method)
{
    return guarded([&] {
        if (method != IPCB_BROAD_LBVH && method != IPCB_BROAD_SAP) throw Error("unknown broad-phase method");
        ctx->broad_method = method;
        ctx->built = false;
    });
}
int ipcb_ctx_set_collision_range(ipcb_ctx* ctx, int32_t rank, int32_t world)
{
    return guarded([&] {
        if (world < 1 || rank < 0 || rank >= world) throw Error("bad collision range");
        ctx->coll_rank = rank;
        ctx->coll_world = world;
    });
}
int ipcb_ctx_set_row_block(ipcb_ctx* ctx, int32_t v_begin, int32_t v_end)
{
    return guarded([&] {
        if (v_end >= 0 && (v_begin < 0 || v_begin > v_end)) throw Error("bad row block");
        ctx->row_lo = v_end < 0 ? 0 : v_begin;
        ctx->row_hi = v_end < 0 ? -1 : v_end;
    });
}
int ipcb_hessian_balanced_row_blocks(ipcb_ctx* ctx, int32_t world, int32_t* bounds)
{
    return guarded([&] {
        begin_call(ctx);
        require_collisions(ctx);
        hessian_balanced_row_blocks(ctx, world, bounds);
    });
}
int ipcb_ctx_launch_count(ipcb_ctx* ctx, int64_t* n)
{
    *n = ctx->launches;
    return 0;
}
int ipcb_ctx_enable_stage_timing(ipcb_ctx* ctx, int32_t on)
{
    ctx->timing = on != 0;
    return 0;
}
int ipcb_ctx_stage_times(ipcb_ctx* ctx, int32_t max_stages, const char** names, float* ms)
{
    ctx->stage_names_keepalive.clear();
    for (auto& s : ctx->stage_ms) ctx->stage_names_keepalive.push_back(s.first);
    int n = 0;
    for (size_t i = 0; i < ctx->stage_ms.size() && n < max_stages; i++, n++) {
        names[n] = ctx->stage_names_keepalive[i].c_str();
        ms[n] = ctx->stage_ms[i].second;
    }
    return n;
}

int ipcb_measure_fp64_peak(ipcb_ctx* ctx, int32_t repeats, double* tflops)
{
    return guarded([&] {
        begin_call(ctx);
        cudaEvent_t a, b;
        IPCB_CUDA(cudaEventCreate(&a));
        IPCB_CUDA(cudaEventCreate(&b));
        const int grid = NUM_SMS * 8, block = 256;
        double best = 0;
        for (int r = 0; r < std::max(repeats, 1) + 1; r++) { // the first launch warms up
            IPCB_CUDA(cudaEventRecord(a, ctx->stream));
            k_fp64_peak<<<grid, block, 0, ctx->stream>>>(1.0, ctx->dScalar.p);
            IPCB_CUDA(cudaEventRecord(b, ctx->stream));
            IPCB_CUDA(cudaEventSynchronize(b));
            float ms = 0;
            IPCB_CUDA(cudaEventElapsedTime(&ms, a, b));
            const double flops = 2.0 * PEAK_ITERS * PEAK_CHAINS * double(grid) * block;
            if (r > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
        }
        cudaEventDestroy(a), cudaEventDestroy(b);
        *tflops = best;
    });
}
int ipcb_measure_copy_bandwidth(ipcb_ctx* ctx, int64_t bytes, int32_t repeats, double* gbs)
{
    return guarded([&] {
        begin_call(ctx);
        if (bytes < 16) throw Error("measure_copy_bandwidth: at least 16 bytes");
        Buf<uint4> src, dst;
        const size_t n = size_t(bytes) / 16;
        src.reserve(n), dst.reserve(n);
        IPCB_CUDA(cudaMemsetAsync(src.p, 1, n * 16, ctx->stream));
        cudaEvent_t a, b;
        IPCB_CUDA(cudaEventCreate(&a));
        IPCB_CUDA(cudaEventCreate(&b));
        double best = 0;
        for (int r = 0; r < std::max(repeats, 1) + 1; r++) {
            IPCB_CUDA(cudaEventRecord(a, ctx->stream));
            k_copy16<<<NUM_SMS * 16, 256, 0, ctx->stream>>>(n, src.p, dst.p);
            IPCB_CUDA(cudaEventRecord(b, ctx->stream));
            IPCB_CUDA(cudaEventSynchronize(b));
            float ms = 0;
            IPCB_CUDA(cudaEventElapsedTime(&ms, a, b));
            if (r > 0) best = std::max(best, 2.0 * double(n) * 16 / (ms * 1e-3) / 1e9);
        }
        cudaEventDestroy(a), cudaEventDestroy(b);
        *gbs = best;
    });
}

// ---- CollisionMesh: collision_mesh.cpp:15-127 (host tables) + device mirrors
int ipcb_mesh_set(ipcb_ctx* ctx, int32_t nV, const double* rest, int32_t ld_rest, int32_t nE, const int32_t* E, int32_t ldE,
                  int32_t nF, const int32_t* F, int32_t ldF)
{
    return guarded([&] {
        begin_call(ctx);
        if (nV < 0 || nE < 0 || nF < 0) throw Error("negative mesh size");
        ctx->nV = nV, ctx->nE = nE, ctx->nF = nF;
        ctx->rest.resize(3 * size_t(nV));
        for (int i = 0; i < nV; i++)
            for (int k = 0; k < 3; k++) ctx->rest[3 * size_t(i) + k] = rest[i + size_t(ld_rest) * k];
        ctx->hE.resize(2 * size_t(nE));
        ctx->hF.resize(3 * size_t(nF));
        for (int i = 0; i < nE; i++)
            for (int k = 0; k < 2; k++) {
                const int v = E[i + size_t(ldE) * k];
                if (v < 0 || v >= nV) throw Error("edge vertex id out of range");
                ctx->hE[2 * size_t(i) + k] = v;
            }
        for (int i = 0; i < nF; i++)
            for (int k = 0; k < 3; k++) {
                const int v = F[i + size_t(ldF) * k];
                if (v < 0 || v >= nV) throw Error("face vertex id out of range");
                ctx->hF[3 * size_t(i) + k] = v;
            }
        // faces_to_edges (:510-543)
        std::unordered_map<uint64_t, int> edge_of;
        edge_of.reserve(size_t(nE) * 2);
        auto ekey = [](int a, int b) { return (uint64_t(uint32_t(std::min(a, b))) << 32) | uint32_t(std::max(a, b)); };
        for (int i = 0; i < nE; i++) edge_of.emplace(ekey(ctx->hE[2 * size_t(i)], ctx->hE[2 * size_t(i) + 1]), i);
        ctx->hF2E.resize(3 * size_t(nF));
        for (int i = 0; i < nF; i++)
            for (int k = 0; k < 3; k++) {
                auto it = edge_of.find(ekey(ctx->hF[3 * size_t(i) + k], ctx->hF[3 * size_t(i) + (k + 1) % 3]));
                if (it == edge_of.end()) throw Error("Unable to find edge!");
                ctx->hF2E[3 * size_t(i) + k] = it->second;
            }
        // codimensional vertices / edges (:145-183)
        std::vector<char> cv(nV, 1), ce(nE, 1);
        for (int v : ctx->hE) cv[v] = 0;
        for (int e : ctx->hF2E) ce[e] = 0;
        ctx->codimV.clear(), ctx->codimE.clear();
        for (int i = 0; i < nV; i++)
            if (cv[i]) ctx->codimV.push_back(i);
        for (int i = 0; i < nE; i++)
            if (ce[i]) ctx->codimE.push_back(i);
        // vertex / edge areas (:309-374)
        auto P = [&](int v, int k) { return ctx->rest[3 * size_t(v) + k]; };
        auto elen = [&](int i) {
            const int a = ctx->hE[2 * size_t(i)], b = ctx->hE[2 * size_t(i) + 1];
            double s = 0;
            for (int k = 0; k < 3; k++) s += (P(a, k) - P(b, k)) * (P(a, k) - P(b, k));
            return std::sqrt(s);
        };
        std::vector<double> vea(nV, -1.0), vfa(nV, -1.0);
        for (int i = 0; i < nE; i++) {
            const double len = elen(i);
            for (int k = 0; k < 2; k++) {
                double& a = vea[ctx->hE[2 * size_t(i) + k]];
                a = std::max(a, 0.0) + 0.5 * len;
            }
        }
        ctx->eArea.assign(nE, -1.0);
        for (int i = 0; i < nF; i++) {
            const int a = ctx->hF[3 * size_t(i)], b = ctx->hF[3 * size_t(i) + 1], c = ctx->hF[3 * size_t(i) + 2];
            const double u[3] = { P(b, 0) - P(a, 0), P(b, 1) - P(a, 1), P(b, 2) - P(a, 2) };
            const double v[3] = { P(c, 0) - P(a, 0), P(c, 1) - P(a, 1), P(c, 2) - P(a, 2) };
            const double n[3] = { u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0] };
            const double area = 0.5 * std::sqrt(n[0] * n[0] + (n[1] * n[1] + n[2] * n[2]));
            for (int k = 0; k < 3; k++) {
                double& va = vfa[ctx->hF[3 * size_t(i) + k]];
                va = std::max(va, 0.0) + area / 3.0;
                double& ea = ctx->eArea[ctx->hF2E[3 * size_t(i) + k]];
                ea = std::max(ea, 0.0) + area / 3.0;
            }
        }
        ctx->vArea.resize(nV);
        for (int i = 0; i < nV; i++) ctx->vArea[i] = vfa[i] < 0 ? (vea[i] < 0 ? 1.0 : vea[i]) : vfa[i];
        for (int i = 0; i < nE; i++)
            if (ctx->eArea[i] < 0) ctx->eArea[i] = elen(i);
        // device mirrors
        std::vector<int2> e2(nE);
        for (int i = 0; i < nE; i++) e2[i] = make_int2(ctx->hE[2 * size_t(i)], ctx->hE[2 * size_t(i) + 1]);
        std::vector<int4> f4(nF), fe4(nF);
        for (int i = 0; i < nF; i++) {
            f4[i] = make_int4(ctx->hF[3 * size_t(i)], ctx->hF[3 * size_t(i) + 1], ctx->hF[3 * size_t(i) + 2], 0);
            fe4[i] = make_int4(ctx->hF2E[3 * size_t(i)], ctx->hF2E[3 * size_t(i) + 1], ctx->hF2E[3 * size_t(i) + 2], 0);
        }
        std::vector<double4> r4(nV);
        for (int i = 0; i < nV; i++) r4[i] = make_double4(P(i, 0), P(i, 1), P(i, 2), 0.0);
        upload(ctx, ctx->dE, e2);
        upload(ctx, ctx->dF, f4);
        upload(ctx, ctx->dF2E, fe4);
        upload(ctx, ctx->dRest, r4);
        upload(ctx, ctx->dVArea, ctx->vArea);
        upload(ctx, ctx->dEArea, ctx->eArea);
        upload(ctx, ctx->dCodimV, ctx->codimV);
        upload(ctx, ctx->dCodimE, ctx->codimE);
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream)); // host vectors go out of scope
        // ids of the codim edges' endpoints in the re-indexed vertex set of the codimensional edge-vertex pass:
        // [codim vertices; referenced vertices of the codim edges, ascending] (candidates.cpp:83-108)
        {
            std::vector<int32_t> ref;
            for (int e : ctx->codimE) ref.push_back(ctx->hE[2 * size_t(e)]), ref.push_back(ctx->hE[2 * size_t(e) + 1]);
            std::sort(ref.begin(), ref.end());
            ref.erase(std::unique(ref.begin(), ref.end()), ref.end());
            const int nCV = int(ctx->codimV.size());
            auto local = [&](int v) { return nCV + int(std::lower_bound(ref.begin(), ref.end(), v) - ref.begin()); };
            std::vector<int2> le(ctx->codimE.size());
            for (size_t i = 0; i < le.size(); i++)
                le[i] = make_int2(local(ctx->hE[2 * size_t(ctx->codimE[i])]), local(ctx->hE[2 * size_t(ctx->codimE[i]) + 1]));
            upload(ctx, ctx->dCodimELocal, le);
            IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        ctx->filter_patches = false, ctx->filter_n_dynamic = -1; // a new mesh accepts all pairs
        ctx->built = false;
        ctx->coll_valid = false;
        ctx->tang_valid = false;
        ctx->adj_ready = false;
        for (auto& c : ctx->cand) c.count = 0;
        for (auto& c : ctx->coll) c.count = 0;
        for (auto& c : ctx->detected) c.count = 0;
    });
}
int ipcb_mesh_set_collision_filter(ipcb_ctx* ctx, const int32_t* patch_ids, int32_t n_dynamic)
{
    return guarded([&] {
        begin_call(ctx);
        ctx->filter_patches = patch_ids != nullptr && ctx->nV > 0;
        if (ctx->filter_patches) {
            ctx->dPatch.reserve(ctx->nV);
            IPCB_CUDA(cudaMemcpyAsync(ctx->dPatch.p, patch_ids, sizeof(int32_t) * ctx->nV, cudaMemcpyHostToDevice, ctx->stream));
            IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        ctx->filter_n_dynamic = n_dynamic;
        ctx->built = false; // resident candidates of the old filter are stale
        for (auto& c : ctx->cand) c.count = 0;
        for (auto& c : ctx->detected) c.count = 0;
    });
}
int ipcb_mesh_num_codim_vertices(ipcb_ctx* ctx, int32_t* n)
{
    *n = int32_t(ctx->codimV.size());
    return 0;
}
int ipcb_mesh_num_codim_edges(ipcb_ctx* ctx, int32_t* n)
{
    *n = int32_t(ctx->codimE.size());
    return 0;
}
int ipcb_mesh_faces_to_edges(ipcb_ctx* ctx, int32_t* f2e)
{
    for (int i = 0; i < ctx->nF; i++)
        for (int k = 0; k < 3; k++) f2e[i + size_t(ctx->nF) * k] = ctx->hF2E[3 * size_t(i) + k];
    return 0;
}
int ipcb_mesh_areas(ipcb_ctx* ctx, double* va, double* ea)
{
    std::copy(ctx->vArea.begin(), ctx->vArea.end(), va);
    std::copy(ctx->eArea.begin(), ctx->eArea.end(), ea);
    return 0;
}

// ---- BroadPhase
int ipcb_broad_build_static(ipcb_ctx* ctx, const double* V, int32_t ld, double r, int32_t boxes)
{
    return guarded([&] {
        begin_call(ctx);
        if (boxes != IPCB_BOXES_FLOAT) throw Error("the CUDA broad phase implements the LBVH float-box predicate only");
        stage_positions(ctx, V, ld, false);
        broad_build(ctx, false, r);
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}
int ipcb_broad_build_swept(ipcb_ctx* ctx, const double* V0, const double* V1, int32_t ld, double r, int32_t boxes)
{
    return guarded([&] {
        begin_call(ctx);
        if (boxes != IPCB_BOXES_FLOAT) throw Error("the CUDA broad phase implements the LBVH float-box predicate only");
        stage_positions(ctx, V0, ld, false);
        stage_positions(ctx, V1, ld, true);
        broad_build(ctx, true, r);
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}
int ipcb_broad_detect(ipcb_ctx* ctx, int32_t kind, int64_t* count)
{
    return guarded([&] {
        begin_call(ctx);
        if (kind < 0 || kind > 5) throw Error("bad candidate kind");
        broad_detect(ctx, kind, ctx->detected[kind]);
        *count = ctx->detected[kind].count;
    });
}
static void fetch_pairs(ipcb_ctx* ctx, PairList& pl, int32_t* pairs)
{
    if (pl.count == 0) return;
    sort_pairs(ctx, pl);
    IPCB_CUDA(cudaMemcpyAsync(pairs, pl.pairs.p, sizeof(int2) * pl.count, cudaMemcpyDeviceToHost, ctx->stream));
    IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
}
int ipcb_broad_fetch(ipcb_ctx* ctx, int32_t kind, int32_t* pairs)
{
    return guarded([&] {
        begin_call(ctx);
        if (kind < 0 || kind > 5) throw Error("bad candidate kind");
        fetch_pairs(ctx, ctx->detected[kind], pairs);
    });
}
int ipcb_broad_vertex_boxes(ipcb_ctx* ctx, void* boxes)
{
    return guarded([&] {
        begin_call(ctx);
        if (!ctx->built) throw Error("broad phase not built");
        IPCB_CUDA(cudaMemcpyAsync(boxes, ctx->vset.box.p, sizeof(FBox) * ctx->nV, cudaMemcpyDeviceToHost, ctx->stream));
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

// ---- Candidates
int ipcb_candidates_build_static(ipcb_ctx* ctx, const double* V, int32_t ld, double r, int64_t counts[4])
{
    return guarded([&] {
        begin_call(ctx);
        stage_positions(ctx, V, ld, false);
        candidates_build(ctx, false, r);
        fill_counts(ctx->cand, counts);
    });
}
int ipcb_candidates_build_swept(ipcb_ctx* ctx, const double* V0, const double* V1, int32_t ld, double r, int64_t counts[4])
{
    return guarded([&] {
        begin_call(ctx);
        stage_positions(ctx, V0, ld, false);
        stage_positions(ctx, V1, ld, true);
        candidates_build(ctx, true, r);
        fill_counts(ctx->cand, counts);
    });
}
int ipcb_candidates_build_swept_dev(ipcb_ctx* ctx, const double* dV0, const double* dV1, int32_t ld, double r, int64_t counts[4])
{
    return guarded([&] {
        begin_call(ctx);
        convert_positions(ctx, dV0, ld, ctx->X0);
        convert_positions(ctx, dV1, ld, ctx->X1);
        candidates_build(ctx, true, r);
        fill_counts(ctx->cand, counts);
    });
}
int ipcb_candidates_fetch(ipcb_ctx* ctx, int32_t kind, int32_t* pairs)
{
    return guarded([&] {
        begin_call(ctx);
        if (kind < 0 || kind > 3) throw Error("bad candidate kind");
        fetch_pairs(ctx, ctx->cand[kind], pairs);
    });
}
int ipcb_candidates_set(ipcb_ctx* ctx, int32_t kind, int64_t count, const int32_t* pairs)
{
    return guarded([&] {
        begin_call(ctx);
        if (kind < 0 || kind > 3) throw Error("bad candidate kind");
        const int limit = kind == IPCB_VV ? ctx->nV : (kind == IPCB_FV ? ctx->nF : ctx->nE);
        const int limit2 = kind == IPCB_EE ? ctx->nE : ctx->nV;
        for (int64_t i = 0; i < count; i++)
            if (pairs[2 * i] < 0 || pairs[2 * i] >= limit || pairs[2 * i + 1] < 0 || pairs[2 * i + 1] >= limit2)
                throw Error("candidate index out of range");
        PairList& pl = ctx->cand[kind];
        pl.pairs.reserve(std::max<int64_t>(count, 1));
        // unordered kinds are kept as (min, max) like the broad phase emits them: a distance type of an edge-edge
        // collision is always relative to that order (vertex_vertex.hpp / edge_edge.hpp compare unordered pairs)
        std::vector<int32_t> canon;
        if ((kind == IPCB_VV || kind == IPCB_EE) && count) {
            canon.assign(pairs, pairs + 2 * count);
            for (int64_t i = 0; i < count; i++)
                if (canon[2 * i] > canon[2 * i + 1]) std::swap(canon[2 * i], canon[2 * i + 1]);
            pairs = canon.data();
        }
        if (count) IPCB_CUDA(cudaMemcpyAsync(pl.pairs.p, pairs, sizeof(int2) * count, cudaMemcpyHostToDevice, ctx->stream));
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
        pl.count = count;
        pl.sorted = false;
    });
}

// ---- NormalCollisions::build
int ipcb_collisions_build_from_candidates_dev(ipcb_ctx* ctx, const double* dV, int32_t ld, double dhat, double dmin, int32_t flags,
                                              int64_t counts[4])
{
    return guarded([&] {
        begin_call(ctx);
        convert_positions(ctx, dV, ld, ctx->X0);
        collisions_build(ctx, dhat, dmin, flags, true);
        coll_counts(ctx, counts);
    });
}
int ipcb_collisions_build_dev(ipcb_ctx* ctx, const double* dV, int32_t ld, double dhat, double dmin, int32_t flags, int64_t counts[4])
{
    return guarded([&] {
        begin_call(ctx);
        convert_positions(ctx, dV, ld, ctx->X0);
        candidates_build(ctx, false, 0.5 * (dhat + dmin)); // normal_collisions.cpp:30
        collisions_build(ctx, dhat, dmin, flags, true);
        coll_counts(ctx, counts);
    });
}
int ipcb_collisions_corrections_keys_dev(ipcb_ctx* ctx, int64_t n[4])
{
    return guarded([&] {
        begin_call(ctx);
        collisions_corrections_keys(ctx, n);
    });
}
int ipcb_collisions_corrections_pack_dev(ipcb_ctx* ctx, void* d_keys)
{
    return guarded([&] {
        begin_call(ctx);
        collisions_corrections_pack(ctx, static_cast<unsigned long long*>(d_keys));
    });
}
int ipcb_collisions_corrections_apply_dev(ipcb_ctx* ctx, const void* d_keys, const int64_t n[4], int64_t counts[4])
{
    return guarded([&] {
        begin_call(ctx);
        collisions_corrections_apply(ctx, static_cast<const unsigned long long*>(d_keys), n, ctx->shard_rank, ctx->shard_world);
        coll_counts(ctx, counts);
    });
}
// host-buffer forms (several builders in one process or over any transport: include/ipcb200.h)
int ipcb_collisions_corrections_keys(ipcb_ctx* ctx, int64_t n[4])
{
    return guarded([&] {
        begin_call(ctx);
        collisions_corrections_keys(ctx, n);
    });
}
int ipcb_collisions_corrections_pack(ipcb_ctx* ctx, uint64_t* keys)
{
    return guarded([&] {
        begin_call(ctx);
        int64_t n[4];
        collisions_corrections_keys(ctx, n);
        const int64_t total = n[0] + n[1] + n[2] + n[3];
        if (total == 0) return;
        ctx->hkey.reserve(total);
        collisions_corrections_pack(ctx, ctx->hkey.p);
        IPCB_CUDA(cudaMemcpyAsync(keys, ctx->hkey.p, sizeof(uint64_t) * total, cudaMemcpyDeviceToHost, ctx->stream));
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}
int ipcb_collisions_corrections_apply(ipcb_ctx* ctx, const uint64_t* keys, const int64_t n[4], int32_t rank, int32_t world, int64_t counts[4])
{
    return guarded([&] {
        begin_call(ctx);
        const int64_t total = n[0] + n[1] + n[2] + n[3];
        ctx->hkey.reserve(std::max<int64_t>(total, 1));
        if (total) IPCB_CUDA(cudaMemcpyAsync(ctx->hkey.p, keys, sizeof(uint64_t) * total, cudaMemcpyHostToDevice, ctx->stream));
        collisions_corrections_apply(ctx, ctx->hkey.p, n, rank, world);
        coll_counts(ctx, counts);
    });
}
int ipcb_collisions_build_from_candidates(ipcb_ctx* ctx, const double* V, int32_t ld, double dhat, double dmin, int32_t flags,
                                          int64_t counts[4])
{
    return guarded([&] {
        begin_call(ctx);
        stage_positions(ctx, V, ld, false);
        collisions_build(ctx, dhat, dmin, flags);
        coll_counts(ctx, counts);
    });
}
int ipcb_collisions_build(ipcb_ctx* ctx, const double* V, int32_t ld, double dhat, double dmin, int32_t flags, int64_t counts[4])
{
    return guarded([&] {
        begin_call(ctx);
        stage_positions(ctx, V, ld, false);
        candidates_build(ctx, false, 0.5 * (dhat + dmin));
        collisions_build(ctx, dhat, dmin, flags);
        coll_counts(ctx, counts);
    });
}
int ipcb_collisions_fetch(ipcb_ctx* ctx, int32_t kind, int32_t* ids, double* weight, double* eps_x, uint8_t* dtype)
{
    return guarded([&] {
        begin_call(ctx);
        if (kind < 0 || kind > 3) throw Error("bad collision kind");
        require_collisions(ctx);
        collisions_sort(ctx, kind);
        const CollisionSet& cs = ctx->coll[kind];
        const size_t n = size_t(cs.count);
        if (n == 0) return;
        cudaStream_t s = ctx->stream;
        if (ids) IPCB_CUDA(cudaMemcpyAsync(ids, cs.ids.p, sizeof(int2) * n, cudaMemcpyDeviceToHost, s));
        if (weight) IPCB_CUDA(cudaMemcpyAsync(weight, cs.w.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
        if (kind == IPCB_EE) {
            if (eps_x) IPCB_CUDA(cudaMemcpyAsync(eps_x, cs.eps.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
            if (dtype) IPCB_CUDA(cudaMemcpyAsync(dtype, cs.dtype.p, n, cudaMemcpyDeviceToHost, s));
        } else {
            if (eps_x) std::fill(eps_x, eps_x + n, 0.0);
            if (dtype) std::fill(dtype, dtype + n, uint8_t(0));
        }
        IPCB_CUDA(cudaStreamSynchronize(s));
    });
}
int ipcb_collisions_clear(ipcb_ctx* ctx)
{
    return guarded([&] {
        begin_call(ctx);
        collisions_clear(ctx);
    });
}
static void check_collision_kind(int32_t kind, int64_t count)
{
    if (kind < 0 || kind > 3) throw Error("bad collision kind");
    if (count < 0) throw Error("negative collision count");
}
int ipcb_collisions_append_dev(ipcb_ctx* ctx, int32_t kind, int64_t count, const int32_t* d_ids, const double* d_weight,
                               const double* d_eps_x, const uint8_t* d_dtype)
{
    return guarded([&] {
        begin_call(ctx);
        check_collision_kind(kind, count);
        collisions_append_dev(ctx, kind, count, d_ids, d_weight, d_eps_x, d_dtype);
    });
}
int ipcb_collisions_append(ipcb_ctx* ctx, int32_t kind, int64_t count, const int32_t* ids, const double* weight, const double* eps_x,
                           const uint8_t* dtype)
{
    return guarded([&] {
        begin_call(ctx);
        check_collision_kind(kind, count);
        if (count == 0) return;
        const int limit = kind == IPCB_VV ? ctx->nV : (kind == IPCB_FV ? ctx->nF : ctx->nE);
        const int limit2 = kind == IPCB_EE ? ctx->nE : ctx->nV;
        for (int64_t i = 0; i < count; i++)
            if (ids[2 * i] < 0 || ids[2 * i] >= limit || ids[2 * i + 1] < 0 || ids[2 * i + 1] >= limit2)
                throw Error("collision index out of range");
        if (kind == IPCB_EE && (!eps_x || !dtype)) throw Error("edge-edge collision records need eps_x and dtype");
        // stage through device scratch (the Hessian key buffers are free between assemblies)
        cudaStream_t s = ctx->stream;
        const size_t n = size_t(count);
        ctx->hkey.reserve(n), ctx->hkey_sorted.reserve(n), ctx->stageA.reserve(n), ctx->stageB.reserve(n);
        IPCB_CUDA(cudaMemcpyAsync(ctx->hkey.p, ids, sizeof(int2) * n, cudaMemcpyHostToDevice, s));
        IPCB_CUDA(cudaMemcpyAsync(ctx->stageA.p, weight, sizeof(double) * n, cudaMemcpyHostToDevice, s));
        if (kind == IPCB_EE) {
            IPCB_CUDA(cudaMemcpyAsync(ctx->stageB.p, eps_x, sizeof(double) * n, cudaMemcpyHostToDevice, s));
            IPCB_CUDA(cudaMemcpyAsync(ctx->hkey_sorted.p, dtype, n, cudaMemcpyHostToDevice, s));
        }
        collisions_append_dev(ctx, kind, count, reinterpret_cast<const int32_t*>(ctx->hkey.p), ctx->stageA.p, ctx->stageB.p,
                              reinterpret_cast<const uint8_t*>(ctx->hkey_sorted.p));
        IPCB_CUDA(cudaStreamSynchronize(s));
    });
}
int ipcb_collisions_merge(ipcb_ctx* ctx, double dmin, int32_t flags, int64_t counts[4])
{
    return guarded([&] {
        begin_call(ctx);
        collisions_merge(ctx, dmin, flags);
        coll_counts(ctx, counts);
    });
}
int ipcb_collision_set_create(ipcb_ctx* ctx, ipcb_collision_set** out)
{
    return guarded([&] {
        *out = new ipcb_collision_set();
        (*out)->device = ctx->device;
    });
}
void ipcb_collision_set_destroy(ipcb_collision_set* set)
{
    if (!set) return;
    cudaSetDevice(set->device);
    delete set;
}
int ipcb_collisions_swap(ipcb_ctx* ctx, ipcb_collision_set* set, int64_t counts[4])
{
    return guarded([&] {
        begin_call(ctx);
        if (!set) throw Error("collisions_swap: null collision set");
        if (set->device != ctx->device) throw Error("collisions_swap: the set lives on another device");
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream)); // nothing in flight may still read the outgoing records
        for (int k = 0; k < 4; k++) ctx->coll[k].swap_records(set->coll[k]);
        std::swap(ctx->dmin, set->dmin);
        std::swap(ctx->coll_valid, set->valid);
        for (int k = 0; k < 4; k++) counts[k] = ctx->coll_valid ? ctx->coll[k].count : 0;
    });
}
int ipcb_collisions_dev_ptrs(ipcb_ctx* ctx, int32_t kind, int64_t* count, const int32_t** d_ids, const double** d_weight,
                             const double** d_eps_x, const uint8_t** d_dtype)
{
    return guarded([&] {
        if (kind < 0 || kind > 3) throw Error("bad collision kind");
        require_collisions(ctx);
        begin_call(ctx);
        collisions_sort(ctx, kind);
        const CollisionSet& cs = ctx->coll[kind];
        *count = cs.count;
        *d_ids = reinterpret_cast<const int32_t*>(cs.ids.p);
        *d_weight = cs.w.p;
        *d_eps_x = kind == IPCB_EE ? cs.eps.p : nullptr;
        *d_dtype = kind == IPCB_EE ? cs.dtype.p : nullptr;
    });
}
namespace {
struct PackedLayout {
    int64_t ids[4], w[4], eps, dt, bytes;
};
PackedLayout packed_layout(const int64_t n[4])
{
    PackedLayout L;
    int64_t off = 0;
    for (int k = 0; k < 4; k++) {
        L.ids[k] = off, off += 8 * n[k];
        L.w[k] = off, off += 8 * n[k];
        if (k == IPCB_EE) L.eps = off, off += 8 * n[k];
    }
    L.dt = off, off += (n[IPCB_EE] + 7) / 8 * 8;
    L.bytes = off;
    return L;
}
} // namespace
int ipcb_collisions_pack_dev(ipcb_ctx* ctx, void* d_buffer, int64_t capacity_bytes, int64_t* bytes)
{
    return guarded([&] {
        begin_call(ctx);
        require_collisions(ctx);
        int64_t n[4];
        coll_counts(ctx, n);
        const PackedLayout L = packed_layout(n);
        *bytes = L.bytes;
        if (L.bytes > capacity_bytes) throw Error("collisions_pack_dev: buffer too small");
        char* out = static_cast<char*>(d_buffer);
        cudaStream_t s = ctx->stream;
        for (int k = 0; k < 4; k++) {
            if (n[k] == 0) continue;
            const CollisionSet& cs = ctx->coll[k];
            IPCB_CUDA(cudaMemcpyAsync(out + L.ids[k], cs.ids.p, 8 * n[k], cudaMemcpyDeviceToDevice, s));
            IPCB_CUDA(cudaMemcpyAsync(out + L.w[k], cs.w.p, 8 * n[k], cudaMemcpyDeviceToDevice, s));
            if (k == IPCB_EE) {
                IPCB_CUDA(cudaMemcpyAsync(out + L.eps, cs.eps.p, 8 * n[k], cudaMemcpyDeviceToDevice, s));
                IPCB_CUDA(cudaMemcpyAsync(out + L.dt, cs.dtype.p, n[k], cudaMemcpyDeviceToDevice, s));
            }
        }
    });
}
int ipcb_collisions_append_packed_dev(ipcb_ctx* ctx, const void* d_buffer, const int64_t counts[4])
{
    return guarded([&] {
        begin_call(ctx);
        for (int k = 0; k < 4; k++) check_collision_kind(k, counts[k]);
        const PackedLayout L = packed_layout(counts);
        collisions_append_packed_dev(ctx, d_buffer, counts, L.ids, L.w, L.eps, L.dt);
    });
}
int ipcb_collisions_min_distance(ipcb_ctx* ctx, const double* V, int32_t ld, double* out)
{
    return guarded([&] {
        begin_call(ctx);
        require_collisions(ctx);
        stage_positions(ctx, V, ld, false);
        *out = collisions_min_distance(ctx);
    });
}

// ---- BarrierPotential
int ipcb_barrier_energy_dev(ipcb_ctx* ctx, const double* dV, int32_t ld, const ipcb_barrier_params* bp, double* d_energy)
{
    return guarded([&] {
        begin_call(ctx);
        require_collisions(ctx);
        convert_positions(ctx, dV, ld, ctx->X0);
        barrier_energy(ctx, *bp, d_energy);
    });
}
int ipcb_barrier_energy(ipcb_ctx* ctx, const double* V, int32_t ld, const ipcb_barrier_params* bp, double* energy)
{
    return guarded([&] {
        begin_call(ctx);
        require_collisions(ctx);
        stage_positions(ctx, V, ld, false);
        double* d_out = reinterpret_cast<double*>(ctx->dCounters.p + 12);
        barrier_energy(ctx, *bp, d_out);
        IPCB_CUDA(cudaMemcpyAsync(energy, d_out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}
int ipcb_barrier_gradient_dev(ipcb_ctx* ctx, const double* dV, int32_t ld, const ipcb_barrier_params* bp, double* d_grad)
{
    return guarded([&] {
        begin_call(ctx);
        require_collisions(ctx);
        convert_positions(ctx, dV, ld, ctx->X0);
        barrier_gradient(ctx, *bp, d_grad);
    });
}
int ipcb_barrier_gradient(ipcb_ctx* ctx, const double* V, int32_t ld, const ipcb_barrier_params* bp, double* grad)
{
    return guarded([&] {
        begin_call(ctx);
        require_collisions(ctx);
        stage_positions(ctx, V, ld, false);
        ctx->dGrad.reserve(3 * size_t(ctx->nV) + 1);
        barrier_gradient(ctx, *bp, ctx->dGrad.p);
        IPCB_CUDA(cudaMemcpyAsync(grad, ctx->dGrad.p, sizeof(double) * 3 * size_t(ctx->nV), cudaMemcpyDeviceToHost, ctx->stream));
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}
int ipcb_barrier_hessian_dev(ipcb_ctx* ctx, const double* dV, int32_t ld, const ipcb_barrier_params* bp, int32_t psd_mode, int64_t* nnz)
{
    return guarded([&] {
        begin_call(ctx);
        require_collisions(ctx);
        if (psd_mode < 0 || psd_mode > 2) throw Error("Invalid type of PSD projection!");
        convert_positions(ctx, dV, ld, ctx->X0);
        barrier_hessian(ctx, *bp, psd_mode);
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
        *nnz = ctx->nnz;
    });
}
int ipcb_barrier_hessian(ipcb_ctx* ctx, const double* V, int32_t ld, const ipcb_barrier_params* bp, int32_t psd_mode, int64_t* nnz)
{
    return guarded([&] {
        begin_call(ctx);
        require_collisions(ctx);
        if (psd_mode < 0 || psd_mode > 2) throw Error("Invalid type of PSD projection!");
        stage_positions(ctx, V, ld, false);
        barrier_hessian(ctx, *bp, psd_mode);
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
        *nnz = ctx->nnz;
    });
}
int ipcb_barrier_hessian_dev_ptrs(ipcb_ctx* ctx, const int32_t** d_outer, const int32_t** d_inner, const double** d_values)
{
    *d_outer = ctx->outer.p;
    *d_inner = ctx->inner.p;
    *d_values = ctx->vals.p;
    return 0;
}
int ipcb_barrier_hessian_fetch(ipcb_ctx* ctx, int32_t* outer, int32_t* inner, double* values)
{
    return guarded([&] {
        begin_call(ctx);
        cudaStream_t s = ctx->stream;
        IPCB_CUDA(cudaMemcpyAsync(outer, ctx->outer.p, sizeof(int) * (3 * size_t(ctx->nV) + 1), cudaMemcpyDeviceToHost, s));
        if (ctx->nnz) {
            IPCB_CUDA(cudaMemcpyAsync(inner, ctx->inner.p, sizeof(int) * ctx->nnz, cudaMemcpyDeviceToHost, s));
            IPCB_CUDA(cudaMemcpyAsync(values, ctx->vals.p, sizeof(double) * ctx->nnz, cudaMemcpyDeviceToHost, s));
        }
        IPCB_CUDA(cudaStreamSynchronize(s));
    });
}

// world_bbox_diagonal_length (utils/world_bbox_diagonal_length.hpp:10-15) of column-major positions
static double bbox_diagonal(const double* V, int n, int ld)
{
    double d2 = 0;
    double ext[3] = { 0, 0, 0 };
    for (int k = 0; k < 3; k++) {
        double lo = INFINITY, hi = -INFINITY;
        for (int i = 0; i < n; i++) lo = std::min(lo, V[i + size_t(ld) * k]), hi = std::max(hi, V[i + size_t(ld) * k]);
        ext[k] = n ? hi - lo : 0.0;
    }
    d2 = (ext[0] * ext[0] + ext[1] * ext[1]) + ext[2] * ext[2];
    return std::sqrt(d2);
}
int ipcb_has_intersections(ipcb_ctx* ctx, const double* V, int32_t ld, int32_t* result)
{
    return guarded([&] {
        begin_call(ctx);
        stage_positions(ctx, V, ld, false);
        *result = has_intersections(ctx, 1e-6 * bbox_diagonal(V, ctx->nV, ld)) ? 1 : 0; // ipc.cpp:120-121
    });
}

// ---- Friction
int ipcb_tangential_build(ipcb_ctx* ctx, const double* V, int32_t ld, const ipcb_barrier_params* normal_potential, const double* mu_s,
                          const double* mu_k, int64_t counts[4])
{
    return guarded([&] {
        begin_call(ctx);
        require_collisions(ctx);
        if (!mu_s || !mu_k) throw Error("tangential_build: mu_s and mu_k are per-vertex arrays");
        stage_positions(ctx, V, ld, false);
        const size_t n = std::max<size_t>(size_t(ctx->nV), 1);
        ctx->dMuS.reserve(n), ctx->dMuK.reserve(n);
        IPCB_CUDA(cudaMemcpyAsync(ctx->dMuS.p, mu_s, sizeof(double) * ctx->nV, cudaMemcpyHostToDevice, ctx->stream));
        IPCB_CUDA(cudaMemcpyAsync(ctx->dMuK.p, mu_k, sizeof(double) * ctx->nV, cudaMemcpyHostToDevice, ctx->stream));
        tangential_build(ctx, *normal_potential, ctx->dMuS.p, ctx->dMuK.p);
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int k = 0; k < 4; k++) counts[k] = ctx->tang[k].count;
    });
}
int ipcb_tangential_fetch(ipcb_ctx* ctx, int32_t kind, int32_t* ids, double* weight, double* normal_force, double* mu_s, double* mu_k,
                          double* closest_point, double* tangent_basis)
{
    return guarded([&] {
        begin_call(ctx);
        if (kind < 0 || kind > 3) throw Error("bad collision kind");
        if (!ctx->tang_valid) throw Error("no tangential collision set has been built on this context");
        const TangSet& t = ctx->tang[kind];
        const size_t n = size_t(t.count);
        if (n == 0) return;
        cudaStream_t s = ctx->stream;
        auto get = [&](void* dst, const void* src, size_t bytes) {
            if (dst) IPCB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
        };
        get(ids, t.ids.p, sizeof(int2) * n), get(weight, t.w.p, 8 * n), get(normal_force, t.N.p, 8 * n), get(mu_s, t.mus.p, 8 * n);
        get(mu_k, t.muk.p, 8 * n), get(closest_point, t.beta.p, 16 * n), get(tangent_basis, t.P.p, 48 * n);
        IPCB_CUDA(cudaStreamSynchronize(s));
    });
}
int ipcb_friction_energy(ipcb_ctx* ctx, const double* velocities, int32_t ld, double eps_v, double* energy)
{
    return guarded([&] {
        begin_call(ctx);
        if (!(eps_v > 0)) throw Error("eps_v must be positive");
        stage_positions(ctx, velocities, ld, false);
        double* d_out = reinterpret_cast<double*>(ctx->dCounters.p + 12);
        friction_energy(ctx, eps_v, d_out);
        IPCB_CUDA(cudaMemcpyAsync(energy, d_out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}
int ipcb_friction_gradient(ipcb_ctx* ctx, const double* velocities, int32_t ld, double eps_v, double* grad)
{
    return guarded([&] {
        begin_call(ctx);
        if (!(eps_v > 0)) throw Error("eps_v must be positive");
        stage_positions(ctx, velocities, ld, false);
        ctx->dGrad.reserve(3 * size_t(ctx->nV) + 1);
        friction_gradient(ctx, eps_v, ctx->dGrad.p);
        IPCB_CUDA(cudaMemcpyAsync(grad, ctx->dGrad.p, sizeof(double) * 3 * size_t(ctx->nV), cudaMemcpyDeviceToHost, ctx->stream));
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}
int ipcb_friction_hessian(ipcb_ctx* ctx, const double* velocities, int32_t ld, double eps_v, int32_t psd_mode, int64_t* nnz)
{
    return guarded([&] {
        begin_call(ctx);
        if (!(eps_v > 0)) throw Error("eps_v must be positive");
        if (psd_mode < 0 || psd_mode > 2) throw Error("Invalid type of PSD projection!");
        stage_positions(ctx, velocities, ld, false);
        friction_hessian(ctx, eps_v, psd_mode);
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
        *nnz = ctx->nnz;
    });
}

// device-resident forms of the friction calls: positions / velocities, coefficients and results stay in HBM
int ipcb_tangential_build_dev(ipcb_ctx* ctx, const double* dV, int32_t ld, const ipcb_barrier_params* normal_potential, const double* d_mu_s,
                              const double* d_mu_k, int64_t counts[4])
{
    return guarded([&] {
        begin_call(ctx);
        require_collisions(ctx);
        if (!d_mu_s || !d_mu_k) throw Error("tangential_build: mu_s and mu_k are per-vertex arrays");
        convert_positions(ctx, dV, ld, ctx->X0);
        tangential_build(ctx, *normal_potential, d_mu_s, d_mu_k);
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int k = 0; k < 4; k++) counts[k] = ctx->tang[k].count;
    });
}
int ipcb_friction_energy_dev(ipcb_ctx* ctx, const double* d_velocities, int32_t ld, double eps_v, double* d_energy)
{
    return guarded([&] {
        begin_call(ctx);
        if (!(eps_v > 0)) throw Error("eps_v must be positive");
        convert_positions(ctx, d_velocities, ld, ctx->X0);
        friction_energy(ctx, eps_v, d_energy);
    });
}
int ipcb_friction_gradient_dev(ipcb_ctx* ctx, const double* d_velocities, int32_t ld, double eps_v, double* d_grad)
{
    return guarded([&] {
        begin_call(ctx);
        if (!(eps_v > 0)) throw Error("eps_v must be positive");
        convert_positions(ctx, d_velocities, ld, ctx->X0);
        friction_gradient(ctx, eps_v, d_grad);
    });
}
int ipcb_friction_hessian_dev(ipcb_ctx* ctx, const double* d_velocities, int32_t ld, double eps_v, int32_t psd_mode, int64_t* nnz)
{
    return guarded([&] {
        begin_call(ctx);
        if (!(eps_v > 0)) throw Error("eps_v must be positive");
        if (psd_mode < 0 || psd_mode > 2) throw Error("Invalid type of PSD projection!");
        convert_positions(ctx, d_velocities, ld, ctx->X0);
        friction_hessian(ctx, eps_v, psd_mode);
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
        *nnz = ctx->nnz;
    });
}

// ---- CCD
int ipcb_ccd_stepsize_from_candidates_dev(ipcb_ctx* ctx, const double* dV0, const double* dV1, int32_t ld, double min_distance,
                                          const ipcb_ccd_params* ccd, double* d_step)
{
    return guarded([&] {
        begin_call(ctx);
        convert_positions(ctx, dV0, ld, ctx->X0);
        convert_positions(ctx, dV1, ld, ctx->X1);
        ccd_stepsize(ctx, min_distance, resolve_ccd(ccd), d_step);
    });
}
int ipcb_ccd_stepsize_dev(ipcb_ctx* ctx, const double* dV0, const double* dV1, int32_t ld, double min_distance, const ipcb_ccd_params* ccd,
                          double* d_step)
{
    return guarded([&] {
        begin_call(ctx);
        convert_positions(ctx, dV0, ld, ctx->X0);
        convert_positions(ctx, dV1, ld, ctx->X1);
        if (candidates_build(ctx, true, 0.5 * min_distance, true)) // ipc.cpp:95-96; too many pairs for one list: stream them
            ccd_stepsize_streaming(ctx, min_distance, resolve_ccd(ccd), d_step);
        else
            ccd_stepsize(ctx, min_distance, resolve_ccd(ccd), d_step);
    });
}
int ipcb_ccd_stepsize_from_candidates(ipcb_ctx* ctx, const double* V0, const double* V1, int32_t ld, double min_distance,
                                      const ipcb_ccd_params* ccd, double* step)
{
    return guarded([&] {
        begin_call(ctx);
        stage_positions(ctx, V0, ld, false);
        stage_positions(ctx, V1, ld, true);
        double* d_out = reinterpret_cast<double*>(ctx->dCounters.p + 13);
        ccd_stepsize(ctx, min_distance, resolve_ccd(ccd), d_out);
        IPCB_CUDA(cudaMemcpyAsync(step, d_out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}
int ipcb_ccd_stepsize(ipcb_ctx* ctx, const double* V0, const double* V1, int32_t ld, double min_distance, const ipcb_ccd_params* ccd,
                      double* step)
{
    return guarded([&] {
        begin_call(ctx);
        stage_positions(ctx, V0, ld, false);
        stage_positions(ctx, V1, ld, true);
        double* d_out = reinterpret_cast<double*>(ctx->dCounters.p + 13);
        if (candidates_build(ctx, true, 0.5 * min_distance, true)) ccd_stepsize_streaming(ctx, min_distance, resolve_ccd(ccd), d_out);
        else ccd_stepsize(ctx, min_distance, resolve_ccd(ccd), d_out);
        IPCB_CUDA(cudaMemcpyAsync(step, d_out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}
int ipcb_candidates_noncandidate_stepsize(ipcb_ctx* ctx, const double* displacements, int32_t ld, double dhat, double* step)
{
    return guarded([&] {
        begin_call(ctx);
        stage_positions(ctx, displacements, ld, true); // X1 <- displacements
        *step = noncandidate_stepsize(ctx, false, dhat);
    });
}
int ipcb_candidates_cfl_stepsize(ipcb_ctx* ctx, const double* V0, const double* V1, int32_t ld, double dhat, double min_distance,
                                 const ipcb_ccd_params* ccd, double* step)
{
    return guarded([&] {
        begin_call(ctx);
        stage_positions(ctx, V0, ld, false);
        stage_positions(ctx, V1, ld, true);
        double* d_out = reinterpret_cast<double*>(ctx->dCounters.p + 13);
        auto fetch = [&] {
            double v = 0;
            IPCB_CUDA(cudaMemcpyAsync(&v, d_out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            IPCB_CUDA(cudaStreamSynchronize(ctx->stream));
            return v;
        };
        ccd_stepsize(ctx, min_distance, resolve_ccd(ccd), d_out);
        const double alpha_c = fetch();
        const double alpha_f = noncandidate_stepsize(ctx, true, dhat);
        if (alpha_f < 0.5 * alpha_c) { // candidates.cpp:356-360: do the full CCD
            if (candidates_build(ctx, true, 0.5 * min_distance, true)) ccd_stepsize_streaming(ctx, min_distance, resolve_ccd(ccd), d_out);
            else ccd_stepsize(ctx, min_distance, resolve_ccd(ccd), d_out);
            *step = fetch();
        } else {
            *step = std::min(alpha_c, alpha_f);
        }
    });
}
int ipcb_ccd_narrow_phase(ipcb_ctx* ctx, int32_t kind, int64_t n, const double* x_t0, const double* x_t1, double min_distance, double tmax,
                          const ipcb_ccd_params* ccd, uint8_t* hit, double* toi)
{
    return guarded([&] {
        begin_call(ctx);
        if (kind < 0 || kind > 3) throw Error("bad candidate kind");
        if (!(tmax >= 0 && tmax <= 1)) throw Error("tmax must be in [0, 1]");
        ccd_narrow_phase(ctx, kind, n, x_t0, x_t1, min_distance, tmax, resolve_ccd(ccd), hit, toi);
    });
}

} // extern "C"
