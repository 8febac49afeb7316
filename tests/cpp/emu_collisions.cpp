// emu_collisions.cpp — TEST INFRASTRUCTURE.  Runs the collision-set kernels of ipc-toolkit_b200/csrc/collisions.cu that
// need no warp of more than one lane (the merge kernels and the whole IMPROVED_MAX_APPROX section) ON THE HOST: the
// kernel source between the [emu-begin]/[emu-end] tags is included verbatim (emu_kernels.inc is cut out of the .cu by
// tests/test_kernel_emulation.py) and compiled with g++ behind a shim that maps the CUDA built-ins onto one-lane warps
// (blockDim = 1, lane 0 is always the leader, atomics are plain adds).  It checks the kernels' LOGIC — index
// arithmetic, adjacency look-ups, weights, distance-type mapping, typed keys, run merging — against the oracle without
// a GPU; races and memory errors are what the -m gpu tests are for.
#include <cuda_runtime.h> // vector types only
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#undef __global__
#undef __device__
#undef __host__
#undef __launch_bounds__
#undef __forceinline__
#define __global__
#define __device__
#define __host__
#define __launch_bounds__(...)
#define __forceinline__ inline
#ifndef __restrict__
#define __restrict__
#endif

namespace emu {
struct Idx {
    unsigned x = 0, y = 0, z = 0;
};
static Idx thread_idx, block_idx, block_dim;
} // namespace emu
#define threadIdx emu::thread_idx
#define blockIdx emu::block_idx
#define blockDim emu::block_dim

static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline int __ffs(unsigned m) { return __builtin_ffs(int(m)); }
static inline int __popc(unsigned m) { return __builtin_popcount(m); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v)
{
    const unsigned long long old = *p;
    *p += v;
    return old;
}
using std::max;
using std::min;

#include "geom.cuh" // with the qualifiers defined away every function in it is a host function

enum { IPCB_VV = 0, IPCB_EV = 1, IPCB_EE = 2, IPCB_FV = 3 };

namespace ipcb {
__device__ inline d3 ld3(const double4* X, int i) { return load_vertex(X, i); }
__device__ inline unsigned long long mkkey(int a, int b) { return ((unsigned long long)(unsigned)a << 32) | (unsigned)b; }
#include "emu_kernels.inc"

// launch with one thread per block: i = blockIdx.x * 1 + 0, lane 0
template <typename F> static void launch(int64_t n, F&& body)
{
    emu::block_dim.x = 1, emu::thread_idx.x = 0;
    for (int64_t i = 0; i < n; i++) {
        emu::block_idx.x = unsigned(i);
        body();
    }
}
} // namespace ipcb

using namespace ipcb;

struct Stream {
    std::vector<unsigned long long> key;
    std::vector<double> w, eps;
    std::vector<unsigned char> dt;
};

// sort by key + run kernels + emit kernels (what merge_stream_enqueue / merge_ee_typed_enqueue enqueue)
static int64_t merge(Stream& st, int kind, bool typed, int32_t* ids, double* w, double* eps, unsigned char* dt)
{
    const int64_t n = int64_t(st.key.size());
    if (n == 0) return 0;
    std::vector<int> idx(n), keep(n, 0), pos(n, 0);
    for (int64_t i = 0; i < n; i++) idx[i] = int(i);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return st.key[a] < st.key[b]; });
    std::vector<unsigned long long> ks(n);
    for (int64_t i = 0; i < n; i++) ks[i] = st.key[idx[i]];
    std::vector<double> wsum(n, 0.0);
    if (typed) launch(n, [&] { k_runs_ee_typed(n, ks.data(), idx.data(), st.w.data(), keep.data(), wsum.data()); });
    else launch(n, [&] { k_runs(n, ks.data(), idx.data(), st.w.data(), kind != IPCB_FV, keep.data(), wsum.data()); });
    int64_t count = 0;
    for (int64_t i = 0; i < n; i++) pos[i] = int(count), count += keep[i];
    int2* out = reinterpret_cast<int2*>(ids);
    if (typed)
        launch(n, [&] { k_emit_ee_typed(n, ks.data(), idx.data(), keep.data(), pos.data(), wsum.data(), st.eps.data(), out, w, eps, dt); });
    else
        launch(n, [&] {
            k_emit_collisions(n, ks.data(), idx.data(), keep.data(), pos.data(), wsum.data(), kind == IPCB_EE ? st.eps.data() : nullptr, st.dt.data(),
                              out, w, eps, dt);
        });
    return count;
}

static int64_t unique_keys(std::vector<unsigned long long>& keys, int64_t n)
{
    keys.resize(n);
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    return int64_t(keys.size());
}

// in: the IPC set's records (what the classification streams hold after merging: merging is associative), candidates,
// mesh and adjacency tables; out: the IMPROVED_MAX_APPROX set.  Returns 0.
extern "C" int emu_improved_build(int nV, int nE, int nF, const double* X4, const double* rest4, const int32_t* E2, const int32_t* F4,
                                  const int32_t* F2E4, const double* vArea, const double* eArea, const int64_t ncand[4],
                                  const int32_t* const cand[4], const int64_t nrec[4], const int32_t* const rec_ids[4],
                                  const double* const rec_w[4], const double* ee_eps, const unsigned char* ee_dt, const int32_t* vv_off,
                                  const int32_t* vv, const int32_t* ve_off, const int32_t* ve, const int32_t* ev_off, const int32_t* ev,
                                  const unsigned char* boundary, int max_ve, double offset_sqr, int area, int64_t out_count[4],
                                  int32_t* const out_ids[4], double* const out_w[4], double* out_eps, unsigned char* out_dt)
{
    (void)nV, (void)nF;
    const double4* X = reinterpret_cast<const double4*>(X4);
    const double4* rest = reinterpret_cast<const double4*>(rest4);
    const int2* E = reinterpret_cast<const int2*>(E2);
    const int4* F = reinterpret_cast<const int4*>(F4);
    const int4* F2E = reinterpret_cast<const int4*>(F2E4);
    const AdjView A { vv_off, vv, ve_off, ve, ev_off, ev, boundary };
    // raw streams = the IPC records
    Stream st[4];
    for (int k = 0; k < 4; k++) {
        st[k].key.resize(nrec[k]), st[k].w.assign(rec_w[k], rec_w[k] + nrec[k]);
        for (int64_t i = 0; i < nrec[k]; i++) st[k].key[i] = mkkey(rec_ids[k][2 * i], rec_ids[k][2 * i + 1]);
    }
    st[IPCB_EE].eps.assign(ee_eps, ee_eps + nrec[IPCB_EE]), st[IPCB_EE].dt.assign(ee_dt, ee_dt + nrec[IPCB_EE]);
    // 1. sub-element candidates
    std::vector<unsigned long long> sub[4];
    unsigned long long cnt[2];
    const int2* cEV = reinterpret_cast<const int2*>(cand[IPCB_EV]);
    const int2* cEE = reinterpret_cast<const int2*>(cand[IPCB_EE]);
    const int2* cFV = reinterpret_cast<const int2*>(cand[IPCB_FV]);
    int64_t nu[4] = { 0, 0, 0, 0 };
    if (ncand[IPCB_EV]) {
        sub[0].assign(2 * ncand[IPCB_EV], 0), cnt[0] = 0;
        launch(ncand[IPCB_EV], [&] { k_sub_from_ev(ncand[IPCB_EV], cEV, E, X, offset_sqr, sub[0].data(), cnt); });
        nu[0] = unique_keys(sub[0], int64_t(cnt[0]));
    }
    if (ncand[IPCB_EE]) {
        sub[1].assign(4 * ncand[IPCB_EE], 0), cnt[0] = 0;
        launch(ncand[IPCB_EE], [&] { k_sub_from_ee(ncand[IPCB_EE], cEE, E, X, offset_sqr, sub[1].data(), cnt); });
        nu[1] = unique_keys(sub[1], int64_t(cnt[0]));
    }
    if (ncand[IPCB_FV]) {
        sub[2].assign(3 * ncand[IPCB_FV], 0), sub[3].assign(3 * ncand[IPCB_FV], 0), cnt[0] = cnt[1] = 0;
        launch(ncand[IPCB_FV], [&] { k_sub_from_fv(ncand[IPCB_FV], cFV, E, F, F2E, X, offset_sqr, sub[2].data(), cnt, sub[3].data(), cnt + 1); });
        nu[2] = unique_keys(sub[2], int64_t(cnt[0]));
        nu[3] = unique_keys(sub[3], int64_t(cnt[1]));
    }
    // 2. room + retype
    unsigned long long c_vv = st[IPCB_VV].key.size(), c_ev = st[IPCB_EV].key.size(), c_ee = st[IPCB_EE].key.size();
    const size_t more_vv = nu[0] + nu[1] + nu[2] + nu[3], more_ev = nu[1] + nu[2], more_ee = size_t(nu[1]) * size_t(std::max(max_ve, 1));
    st[IPCB_VV].key.resize(c_vv + more_vv), st[IPCB_VV].w.resize(c_vv + more_vv);
    st[IPCB_EV].key.resize(c_ev + more_ev), st[IPCB_EV].w.resize(c_ev + more_ev);
    st[IPCB_EE].key.resize(c_ee + more_ee), st[IPCB_EE].w.resize(c_ee + more_ee), st[IPCB_EE].eps.resize(c_ee + more_ee),
        st[IPCB_EE].dt.resize(c_ee + more_ee);
    const int64_t n_ee0 = int64_t(c_ee);
    launch(n_ee0, [&] { k_retype_ee_keys(n_ee0, st[IPCB_EE].key.data(), st[IPCB_EE].dt.data()); });
    // 3. corrections
    const CorrOut o { st[IPCB_VV].key.data(), st[IPCB_EV].key.data(), st[IPCB_EE].key.data(), st[IPCB_VV].w.data(), st[IPCB_EV].w.data(),
                      st[IPCB_EE].w.data(), st[IPCB_EE].eps.data(), st[IPCB_EE].dt.data(), &c_vv, &c_ev, &c_ee };
    launch(nu[0], [&] { k_corr_vv(nu[0], sub[0].data(), A, vArea, area, 0, o); });
    launch(nu[1], [&] { k_corr_ev_from_ee(nu[1], sub[1].data(), A, E, X, rest, eArea, area, o); });
    launch(nu[2], [&] { k_corr_ev_from_fv(nu[2], sub[2].data(), A, E, X, vArea, area, o); });
    launch(nu[3], [&] { k_corr_vv(nu[3], sub[3].data(), A, vArea, area, 1, o); });
    if (c_vv > st[IPCB_VV].key.size() || c_ev > st[IPCB_EV].key.size() || c_ee > st[IPCB_EE].key.size()) return 1; // capacity bound broken
    st[IPCB_VV].key.resize(c_vv), st[IPCB_VV].w.resize(c_vv);
    st[IPCB_EV].key.resize(c_ev), st[IPCB_EV].w.resize(c_ev);
    st[IPCB_EE].key.resize(c_ee), st[IPCB_EE].w.resize(c_ee), st[IPCB_EE].eps.resize(c_ee), st[IPCB_EE].dt.resize(c_ee);
    // 4. merge
    for (int k = 0; k < 4; k++) out_count[k] = merge(st[k], k, k == IPCB_EE, out_ids[k], out_w[k], out_eps, out_dt);
    (void)nE;
    return 0;
}
