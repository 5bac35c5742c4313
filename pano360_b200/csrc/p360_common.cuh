// Shared helpers for the pano360_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pano360_b200.h"

namespace p360 {

// ---- per-thread error record (C ABI: p360_last_error) ---------------------
inline char *err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
inline int fail(int code, const char *where, const char *what) {
    snprintf(err_buf(), 512, "%s: %s", where, what);
    return code;
}
inline int check_launch(const char *where) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, where, cudaGetErrorString(e));
    return 0;
}
#define P360_REQUIRE(cond, where)                                             \
    do {                                                                      \
        if (!(cond)) return p360::fail(P360_EINVAL, where, "invalid argument: " #cond); \
    } while (0)
#define P360_CUDA(call, where)                                                \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return p360::fail((int)e__, where, cudaGetErrorString(e__)); \
    } while (0)

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// ---- device helpers -------------------------------------------------------
// BORDER_REFLECT  (fedcba|abcdefgh|hgfedcb): period 2n
__device__ __forceinline__ int reflect_edge(int p, int n) {
    if (n == 1) return 0;
    int m = 2 * n;
    int q = p % m;
    if (q < 0) q += m;
    return q < n ? q : m - 1 - q;
}
// BORDER_REFLECT_101 (gfedcb|abcdefgh|gfedcba): period 2n-2
__device__ __forceinline__ int reflect_101(int p, int n) {
    if (n == 1) return 0;
    int m = 2 * n - 2;
    int q = p % m;
    if (q < 0) q += m;
    return q < n ? q : m - q;
}
// cvRound(v * 32) with x86 cvtps2dq semantics: NaN / out of int32 -> INT_MIN.
__device__ __forceinline__ int to_fixed5(float v) {
    float s = __fmul_rn(v, 32.0f);
    if (!(fabsf(s) < 2147483648.0f)) return INT32_MIN;
    return __float2int_rn(s);
}
__device__ __forceinline__ int sat16(int v) { return max(-32768, min(32767, v)); }

// ---- owner map (stitcher.py:196-204) as one 64-bit key per mosaic pixel ----
// key = float_bits(alpha) << 32 | (0xFFFFFFFF - patch): atomicMax over the
// patches gives the largest alpha and, among equal alphas, the smallest patch
// number — np.argmax's "first maximum wins" — independent of execution order.
// Only alpha > 0 ever competes, so key == 0 means "no owner" (-1).
__device__ __forceinline__ unsigned long long owner_key(float alpha, int patch) {
    return ((unsigned long long)__float_as_uint(alpha) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)patch);
}
__device__ __forceinline__ bool key_is_owner(unsigned long long key, int patch) {
    return key != 0ull && (unsigned)(key & 0xFFFFFFFFull) == 0xFFFFFFFFu - (unsigned)patch;
}
__device__ __forceinline__ void owner_compete(unsigned long long *keys, size_t mi, float alpha, int patch) {
    if (alpha > 0.0f) atomicMax(keys + mi, owner_key(alpha, patch));
}

// ---- owned boxes -------------------------------------------------------------
// own = {x0, y0, x1, y1}: box (patch pixels) around the pixels a patch owns, written on the
// device by p360_owned_boxes.  A patch has non-zero blend weights only within the reach of
// the widest blur around that box, so every stage restricts itself to a dilation of it.
// own == nullptr: no restriction.
__device__ __forceinline__ bool near_owned(const int *own, int grow, int ax0, int ay0, int ax1, int ay1) {
    if (own == nullptr) return true;
    const int ox0 = __ldg(own), oy0 = __ldg(own + 1), ox1 = __ldg(own + 2), oy1 = __ldg(own + 3);
    if (ox1 <= ox0 || oy1 <= oy0) return false;          // owns nothing
    return ax0 < ox1 + grow && ax1 > ox0 - grow && ay0 < oy1 + grow && ay1 > oy0 - grow;
}

// Streaming (read-once / write-once) 128-bit accesses: keep L1 for the gathers.
__device__ __forceinline__ float4 ld_stream(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}
__device__ __forceinline__ void st_stream(float4 *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace p360
