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

// ---- the per-patch record shared by reduce, blur and the collapse kernels -----
struct BandPatch {                      // == p360_band_patch
    const float4 *rgba;                 // full-res patch; alpha is replaced by (owner == index) for multiband
    const uint8_t *invalid;             // ph x pw mask (linear / paste only)
    float4 *d2, *d4;                    // reduce outputs (f = 2, f = 4)
    const float4 *low[P360_MAX_LEVELS - 1];   // blurred coarse image of level l (level 0: f = 2, others f = 4)
    int x0, y0, pw, ph;                 // box in (window) mosaic pixels
    int w4, h4;                         // size of the f = 4 grid (f = 2 grid is twice that)
    int pad;                            // extension R in full-res pixels (multiple of 4)
    int index;                          // id of this patch in the owner keys
    int own[4];                         // box around the owned pixels (patch px), see p360_owned_boxes
};
static_assert(sizeof(BandPatch) == sizeof(p360_band_patch), "ABI struct mismatch");

// ---- seam-band maps -------------------------------------------------------------
// Bitmaps over the 64 x 32 mosaic tiles, one bit per patch (p360_tile_maps_build):
//   present  patches that own a pixel of the tile
//   cand     patches that own a pixel within the blur reach of the tile: the only ones
//            that can carry weight there (one bit set: the tile is its owner's pixels)
//   need     patches whose coarse levels are read within the blur chain's reach of the
//            tile: everything else of reduce / blur is never consumed
// maps.need == nullptr: no maps, the owned boxes decide.
// The same bitmaps are also built from the geometry alone, before anything is sampled
// (p360_seam_plan_build, p360_warp.cu): supersets of the ones above, plus `wneed`.
constexpr int TILE_X = 64, TILE_Y = 32;
struct TileMaps {                       // == p360_tile_maps
    uint32_t *present, *cand, *need;
    uint8_t *multi;                     // per tile: more than one candidate
    uint2 *work;                        // compacted list of the blocks a pass has to run
    int *work_count;
    uint32_t *wneed;                    // seam plan only (p360_seam_plan_build): where float pixels are wanted;
                                        // non-null also says: solo tiles were written by p360_warp_tiles
    int tiles_x, tiles_y, words;
    int row0;                           // window row of tile row 0 (<= 0; tiles sit on absolute mosaic rows)
    int reach_x, reach_y;               // blur reach in tiles
    int work_cap;
    int h_rows;                         // 4: horizontal blur lists in 64-cell segments (else 256)
};

// Persistent grids over a work list: blocks per SM x SMs (cached per process).
inline int persistent_blocks(int per_sm) {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaDeviceProp prop;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148 * per_sm;
        sms = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    }
    return sms * per_sm;
}
static_assert(sizeof(TileMaps) == sizeof(p360_tile_maps), "ABI struct mismatch");

// Is `patch`'s bit set in any tile of `map` that the window-mosaic box [xa, xb) x [ya, yb)
// touches?  Boxes beyond the mosaic (reflected extension) count for the nearest edge tile.
// Every thread of a block evaluates this on the same arguments: the outcome is block-uniform.
__device__ __forceinline__ bool tiles_test(const TileMaps &m, const uint32_t *map, int patch,
                                           int xa, int ya, int xb, int yb) {
    const int tx0 = min(max(xa >> 6, 0), m.tiles_x - 1), tx1 = min(max((xb - 1) >> 6, 0), m.tiles_x - 1);
    const int ty0 = min(max((ya - m.row0) >> 5, 0), m.tiles_y - 1);
    const int ty1 = min(max((yb - 1 - m.row0) >> 5, 0), m.tiles_y - 1);
    const int word = patch >> 5;
    const uint32_t bit = 1u << (patch & 31);
    for (int ty = ty0; ty <= ty1; ++ty)
        for (int tx = tx0; tx <= tx1; ++tx)
            if (__ldg(map + ((size_t)ty * m.tiles_x + tx) * m.words + word) & bit) return true;
    return false;
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
