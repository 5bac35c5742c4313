// TEST INFRASTRUCTURE — not part of the product, never loaded by pano360_b200.
//
// A minimal stand-in for <cuda_runtime.h> that lets g++ compile the package's
// .cu sources (after tests/emul/build_emul.py has rewritten the <<<...>>>
// launches) into a host library which executes the kernels' own code on the
// CPU: every CUDA thread of a block is a ucontext fiber, __syncthreads /
// __syncwarp / shuffles / ballots are cooperative rendezvous between fibers,
// blocks are spread over a few OS threads.  Purpose: run the real kernel source
// against the oracle in the CPU test tier (no GPU in the build container) and
// catch logic errors before spending GPU time.  It says nothing about speed and
// is not a fallback: the package loads only libpano360_b200.so (sm_100a).
//
// Arithmetic notes: build with -ffp-contract=off so that a * b + c is contracted
// only where the kernels ask for it (__ffma2_rn, fmaf, fma); nvcc additionally
// contracts plain a * b + c expressions, so float paths that are not written
// with explicit-rounding intrinsics may differ from the GPU in the last ulp.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <thread>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __constant__
#define __shared__ static thread_local
#define P360_EMUL_BUILD 1
#define __align__(n) alignas(n)

// ---- vector types ------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct int4 { int x, y, z, w; };
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

// ---- runtime API subset --------------------------------------------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0 };
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = 227 * 1024; return cudaSuccess; }
struct cudaDeviceProp { int multiProcessorCount, major, minor, l2CacheSize; };
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    *p = cudaDeviceProp{(int)std::thread::hardware_concurrency(), 0, 0, 0};
    return cudaSuccess;
}
inline cudaError_t cudaGetDevice(int *dev) { *dev = 0; return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *p, int value, size_t n, cudaStream_t) { memset(p, value, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                                     int, cudaStream_t) {
    for (size_t r = 0; r < height; ++r) memcpy((char *)dst + r * dpitch, (const char *)src + r * spitch, width);
    return cudaSuccess;
}
template <class K>
inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
template <class T>
inline cudaError_t cudaMemcpyToSymbolAsync(T &symbol, const void *src, size_t n, size_t offset, int, cudaStream_t) {
    memcpy(reinterpret_cast<char *>(&symbol) + offset, src, n);
    return cudaSuccess;
}

// ---- fibers ----------------------------------------------------------------------
namespace p360_emul {

constexpr int MAX_THREADS = 1024;
constexpr size_t STACK_BYTES = 256 * 1024;
constexpr size_t DYN_SMEM_BYTES = 256 * 1024;

struct WarpState {
    unsigned alive;              // lanes that have not returned
    unsigned gen;                // rendezvous generation
    int arrived;
    unsigned ballot[2];
    uint64_t slot[2][32];
};

struct Block {
    ucontext_t scheduler;
    ucontext_t ctx[MAX_THREADS];
    char *stacks = nullptr;
    char *dyn_smem = nullptr;
    bool done[MAX_THREADS];
    uint3 tid[MAX_THREADS];
    int n_threads = 0, alive = 0, current = 0;
    int bar_arrived = 0;
    unsigned bar_gen = 0;
    int bar_or[2] = {0, 0};
    unsigned long progress = 0;
    WarpState warps[MAX_THREADS / 32];
    const std::function<void()> *body = nullptr;
};

struct BlockOwner {              // one per OS thread; workers of a launch end with it, so free what they touched
    Block *b = nullptr;
    ~BlockOwner() {
        if (b) {
            free(b->stacks);
            free(b->dyn_smem);
            delete b;
        }
    }
};

inline Block &block() {
    static thread_local BlockOwner owner;
    if (!owner.b) {
        owner.b = new Block();
        owner.b->stacks = static_cast<char *>(aligned_alloc(64, STACK_BYTES * MAX_THREADS));
        owner.b->dyn_smem = static_cast<char *>(aligned_alloc(64, DYN_SMEM_BYTES));
    }
    return *owner.b;
}
inline void *dyn_smem() { return block().dyn_smem; }

inline void yield() {
    Block &b = block();
    swapcontext(&b.ctx[b.current], &b.scheduler);
}
inline void no_tma() { fprintf(stderr, "p360_emul: TMA path reached in the host build\n"); abort(); }
inline int linear_tid() { return (threadIdx.z * blockDim.y + threadIdx.y) * blockDim.x + threadIdx.x; }
inline unsigned popc(unsigned v) { return (unsigned)__builtin_popcount(v); }

inline void release_block_barrier(Block &b) {
    b.bar_arrived = 0;
    ++b.bar_gen;
    b.bar_or[b.bar_gen & 1 ? 1 : 0] = 0;       // the slot of the generation that starts now
    ++b.progress;
}
inline void release_warp(WarpState &w) {
    w.arrived = 0;
    ++w.gen;
    w.ballot[w.gen & 1] = 0;
}

inline int sync_block(int pred) {
    Block &b = block();
    const unsigned my = b.bar_gen;
    const int par = my & 1;
    b.bar_or[par] |= pred;
    if (++b.bar_arrived == b.alive) {
        const int r = b.bar_or[par];
        release_block_barrier(b);
        return r;
    }
    while (b.bar_gen == my) yield();
    return b.bar_or[par];
}

// rendezvous of the lanes in `mask` that are still alive; returns the parity of the buffers
inline void sync_warp(unsigned mask) {
    Block &b = block();
    WarpState &w = b.warps[linear_tid() >> 5];
    const unsigned my = w.gen;
    if (++w.arrived >= (int)popc(mask & w.alive)) {
        release_warp(w);
        ++b.progress;
        return;
    }
    while (w.gen == my) yield();
}

inline void fiber_main() {
    Block &b = block();
    (*b.body)();
    // the thread returns: it no longer takes part in barriers (CUDA counts exited threads as arrived)
    const int t = b.current;
    b.done[t] = true;
    --b.alive;
    ++b.progress;
    WarpState &w = b.warps[t >> 5];
    w.alive &= ~(1u << (t & 31));
    if (w.arrived > 0 && w.arrived >= (int)popc(w.alive)) release_warp(w);
    if (b.alive > 0 && b.bar_arrived == b.alive) release_block_barrier(b);
    swapcontext(&b.ctx[t], &b.scheduler);
}

inline void run_block(const dim3 &grid, const dim3 &bdim, unsigned linear_block, const std::function<void()> &body) {
    Block &b = block();
    gridDim = grid;
    blockDim = bdim;
    blockIdx.x = linear_block % grid.x;
    blockIdx.y = (linear_block / grid.x) % grid.y;
    blockIdx.z = linear_block / (grid.x * grid.y);
    const int n = (int)(bdim.x * bdim.y * bdim.z);
    b.n_threads = b.alive = n;
    b.bar_arrived = 0; b.bar_gen = 0; b.bar_or[0] = b.bar_or[1] = 0;
    b.body = &body;
    for (int wi = 0; wi < (n + 31) / 32; ++wi) {
        WarpState &w = b.warps[wi];
        const int lanes = std::min(32, n - 32 * wi);
        w.alive = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1);
        w.gen = 0; w.arrived = 0; w.ballot[0] = w.ballot[1] = 0;
    }
    for (int t = 0; t < n; ++t) {
        b.done[t] = false;
        b.tid[t].x = t % bdim.x;
        b.tid[t].y = (t / bdim.x) % bdim.y;
        b.tid[t].z = t / (bdim.x * bdim.y);
        getcontext(&b.ctx[t]);
        b.ctx[t].uc_stack.ss_sp = b.stacks + STACK_BYTES * t;
        b.ctx[t].uc_stack.ss_size = STACK_BYTES;
        b.ctx[t].uc_link = nullptr;
        makecontext(&b.ctx[t], fiber_main, 0);
    }
    while (b.alive > 0) {
        const unsigned long before = b.progress;
        for (int t = 0; t < n; ++t) {
            if (b.done[t]) continue;
            b.current = t;
            threadIdx = b.tid[t];
            swapcontext(&b.scheduler, &b.ctx[t]);
        }
        if (b.alive > 0 && b.progress == before) {
            fprintf(stderr, "p360_emul: deadlock in block (%u,%u,%u): %d threads wait at a barrier not all reach\n",
                    blockIdx.x, blockIdx.y, blockIdx.z, b.alive);
            abort();
        }
    }
}

inline int worker_count() {
    const char *env = getenv("P360_EMUL_THREADS");
    int n = env ? atoi(env) : (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(n, 64));
}

inline void launch(dim3 grid, dim3 bdim, size_t smem_bytes, const std::function<void()> &body) {
    const unsigned long long total = (unsigned long long)grid.x * grid.y * grid.z;
    if (total == 0 || bdim.x * bdim.y * bdim.z == 0) return;
    if (bdim.x * bdim.y * bdim.z > (unsigned)MAX_THREADS || smem_bytes > DYN_SMEM_BYTES) {
        fprintf(stderr, "p360_emul: launch configuration out of range\n");
        abort();
    }
    std::atomic<unsigned long long> next{0};
    auto work = [&]() {
        for (;;) {
            const unsigned long long i = next.fetch_add(1);
            if (i >= total) break;
            run_block(grid, bdim, (unsigned)i, body);
        }
    };
    const int nw = (int)std::min<unsigned long long>(worker_count(), total);
    if (nw <= 1) { work(); return; }
    std::vector<std::thread> pool;
    for (int i = 0; i < nw; ++i) pool.emplace_back(work);
    for (auto &t : pool) t.join();
}

}  // namespace p360_emul

// ---- synchronisation and warp collectives -----------------------------------------
inline void __syncthreads() { p360_emul::sync_block(0); }
inline int __syncthreads_or(int pred) { return p360_emul::sync_block(pred != 0); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { p360_emul::sync_warp(mask); }

template <class T>
inline T __shfl_sync(unsigned mask, T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of up to 64 bits");
    p360_emul::Block &b = p360_emul::block();
    const int t = p360_emul::linear_tid();
    p360_emul::WarpState &w = b.warps[t >> 5];
    const int par = w.gen & 1;
    memcpy(&w.slot[par][t & 31], &v, sizeof(T));
    p360_emul::sync_warp(mask);
    T r;
    memcpy(&r, &w.slot[par][src_lane & 31], sizeof(T));
    return r;
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) {
    return __shfl_sync(mask, v, (p360_emul::linear_tid() & 31) ^ lane_mask);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    p360_emul::Block &b = p360_emul::block();
    const int t = p360_emul::linear_tid();
    p360_emul::WarpState &w = b.warps[t >> 5];
    const int par = w.gen & 1;
    if (pred) w.ballot[par] |= 1u << (t & 31);
    p360_emul::sync_warp(mask);
    return w.ballot[par] & mask;
}

// ---- memory ------------------------------------------------------------------------
template <class T>
inline T __ldg(const T *p) { return *p; }

inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline int atomicMax(int *p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline int atomicMin(int *p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

// ---- arithmetic intrinsics (IEEE round-to-nearest on x86-64 SSE/FMA) -----------------
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }

inline int p360_emul_sat_int(double v) {           // CUDA float->int conversions saturate; NaN -> 0
    if (v != v) return 0;
    if (v >= 2147483647.0) return INT32_MAX;
    if (v <= -2147483648.0) return INT32_MIN;
    return (int)v;
}
inline int __float2int_rz(float v) { return p360_emul_sat_int(trunc((double)v)); }
inline int __float2int_rn(float v) { return p360_emul_sat_int(nearbyint((double)v)); }
inline int __float2int_rd(float v) { return p360_emul_sat_int(floor((double)v)); }
inline int __double2int_rn(double v) { return p360_emul_sat_int(nearbyint(v)); }
inline unsigned __float_as_uint(float v) { unsigned u; memcpy(&u, &v, 4); return u; }
inline float __uint_as_float(unsigned u) { float v; memcpy(&v, &u, 4); return v; }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { unsigned long long v = ((unsigned long long)hi << 32) | lo; return (unsigned)(v >> (s & 31)); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }

inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
