// ee_common.cuh -- shared helpers for the sm_100a engine.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ee_b200.h"

namespace ee {

// ---- error plumbing -----------------------------------------------------------------------------------------
extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launch_count;
extern std::atomic<int> g_pair_variant;

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define EE_CUDA(expr)                                                                                         \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess)                                                                                \
            throw ::ee::Error(EE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                                               ":" + std::to_string(__LINE__) + ")");                        \
    } while (0)

#define EE_REQUIRE(cond, msg)                                        \
    do {                                                             \
        if (!(cond)) throw ::ee::Error(EE_ERR_INVALID, std::string(msg)); \
    } while (0)

inline void count_launch(uint64_t n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

// Per-device allocator stream of the stream-ordered memory pool (defined in ee_nbody.cu).  The pool keeps freed blocks
// (release threshold = unlimited), so allocating and freeing device buffers costs microseconds and never synchronises
// the device: a propagator clone at every Planner snapshot (prediction.rs:224-229) would otherwise spend more time in
// cudaMalloc / cudaFree than in stepping.
cudaStream_t pool_stream(int device);

// RAII device buffer.  Pooled by default (cudaMallocAsync on the device's pool); `ipc = true` asks for plain cudaMalloc
// memory, which CUDA-IPC handles (the NVLink peer path) need.  Discipline that makes the pooled free safe: a buffer is
// only released after the streams that used it have been synchronised (every temporary in this code base already is).
template <class T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
    int dev = -1;
    bool pooled = true;
    DBuf() = default;
    explicit DBuf(size_t count) { alloc(count); }
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    DBuf(DBuf&& o) noexcept : p(o.p), n(o.n), dev(o.dev), pooled(o.pooled) { o.p = nullptr; o.n = 0; }
    DBuf& operator=(DBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p;
            n = o.n;
            dev = o.dev;
            pooled = o.pooled;
            o.p = nullptr;
            o.n = 0;
        }
        return *this;
    }
    ~DBuf() { release(); }
    void alloc(size_t count, bool ipc = false) {
        release();
        n = count;
        if (!count) return;
        EE_CUDA(cudaGetDevice(&dev));
        pooled = !ipc;
        if (pooled) {
            // a stream-ordered allocation may only be touched after its stream has reached it (the pool can map new
            // physical memory in stream order): the pool's stream carries nothing but allocations and frees, so waiting
            // for it costs a few microseconds and makes the pointer usable on every stream
            EE_CUDA(cudaMallocAsync((void**)&p, count * sizeof(T), pool_stream(dev)));
            EE_CUDA(cudaStreamSynchronize(pool_stream(dev)));
        } else
            EE_CUDA(cudaMalloc(&p, count * sizeof(T)));
    }
    void release() {
        if (p) {
            if (pooled) {
                int cur = -1;
                cudaGetDevice(&cur);
                if (cur != dev) cudaSetDevice(dev);
                cudaFreeAsync(p, pool_stream(dev));
                if (cur != dev && cur >= 0) cudaSetDevice(cur);
            } else {
                cudaFree(p);
            }
        }
        p = nullptr;
        n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};

// ---- exact (reference-order) arithmetic: never contracted, round-to-nearest-even ----------------------------
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double xsqrt(double a) { return __dsqrt_rn(a); }

struct D3 {
    double x, y, z;
};
__device__ __forceinline__ D3 d3(double x, double y, double z) { return D3{x, y, z}; }
// lane-wise DVec3 arithmetic in reference order
__device__ __forceinline__ D3 xadd3(D3 a, D3 b) { return {xadd(a.x, b.x), xadd(a.y, b.y), xadd(a.z, b.z)}; }
__device__ __forceinline__ D3 xsub3(D3 a, D3 b) { return {xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z)}; }
__device__ __forceinline__ D3 xmul3(D3 a, double s) { return {xmul(a.x, s), xmul(a.y, s), xmul(a.z, s)}; }
__device__ __forceinline__ D3 xdiv3(D3 a, double s) { return {xdiv(a.x, s), xdiv(a.y, s), xdiv(a.z, s)}; }
__device__ __forceinline__ D3 xneg3(D3 a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ double xdot3(D3 a, D3 b) {  // glam: (x*x) + (y*y) + (z*z)
    return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z));
}
__device__ __forceinline__ D3 xcross3(D3 a, D3 b) {  // glam DVec3::cross
    return {xsub(xmul(a.y, b.z), xmul(b.y, a.z)), xsub(xmul(a.z, b.x), xmul(b.z, a.x)),
            xsub(xmul(a.x, b.y), xmul(b.x, a.y))};
}

}  // namespace ee
