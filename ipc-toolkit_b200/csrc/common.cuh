// common.cuh — context, device buffers, launch accounting for libipcb200.so.
// Hand-written sm_100a CUDA; FP64 arithmetic; no tensor cores (nothing on this
// path is a dense contraction), the kernels are HBM- or FP64-pipe bound.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <utility>
#include <string>
#include <vector>
#include "../../include/ipcb200.h"

namespace ipcb {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define IPCB_CUDA(expr)                                                                                               \
    do {                                                                                                              \
        cudaError_t _e = (expr);                                                                                      \
        if (_e != cudaSuccess)                                                                                        \
            throw ::ipcb::Error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":"            \
                                + std::to_string(__LINE__) + ")");                                                    \
    } while (0)

constexpr int NUM_SMS = 148; // B200: 2 dies x 74 SMs

// Growable device array.  Capacity only grows (x1.5) so that a steady-state
// contact step performs no allocation.
template <typename T> struct Buf {
    T* p = nullptr;
    size_t cap = 0;
    ~Buf() { release(); }
    Buf() = default;
    Buf(const Buf&) = delete;
    Buf& operator=(const Buf&) = delete;
    void swap(Buf& o)
    {
        std::swap(p, o.p);
        std::swap(cap, o.cap);
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    // ensure capacity for n elements; contents are NOT preserved on growth
    void reserve(size_t n)
    {
        if (n <= cap) return;
        release();
        size_t want = n + n / 2 + 64;
        IPCB_CUDA(cudaMalloc(&p, want * sizeof(T)));
        cap = want;
    }
    // grow preserving the first `keep` elements
    void reserve_keep(size_t n, size_t keep, cudaStream_t s)
    {
        if (n <= cap) return;
        size_t want = n + n / 2 + 64;
        T* q = nullptr;
        IPCB_CUDA(cudaMalloc(&q, want * sizeof(T)));
        if (p && keep) IPCB_CUDA(cudaMemcpyAsync(q, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, s));
        IPCB_CUDA(cudaStreamSynchronize(s));
        if (p) cudaFree(p);
        p = q;
        cap = want;
    }
};

// pinned host scalar slots for counters read back once per stage
struct Pinned {
    int64_t* p = nullptr;
    void init(size_t n) { IPCB_CUDA(cudaMallocHost(&p, n * sizeof(int64_t))); }
    ~Pinned()
    {
        if (p) cudaFreeHost(p);
    }
};

inline unsigned grid_for(size_t n, int block) { return unsigned((n + block - 1) / block); }

} // namespace ipcb
