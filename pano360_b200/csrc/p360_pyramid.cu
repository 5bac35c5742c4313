// Reduced-resolution evaluation of the reference's wide Gaussians and the
// output-stationary blenders that consume them.
//
// The reference blurs every full-resolution RGBA patch with sigma = 4, 6.9,
// 8.9, 10.6(, 12) (stitcher.py:218, :226).  Those filters are so smooth that
// the result can be evaluated on a coarser grid: area-
// reduce the patch by f = 2 (level 0) or f = 4 (levels >= 1), blur there with
// sigma' = sqrt(sigma^2 - (f^2-1)/12 - f^2/6) / f (the box and the bilinear
// kernels contribute the subtracted variance), and expand bilinearly where the
// band is consumed.  BORDER_REFLECT_101 at the patch edges is honoured by
// reducing the *reflected extension* of the patch: the coarse grids cover
// [-R, n + R) in full-resolution pixels.
//
//   p360_pyramid_reduce_batch  full-res RGBA (+ owner keys) -> D2, D4  ("reduce")
//   p360_gauss_blur_batch      coarse blurs (p360_blur.cu)
//   p360_multiband_collapse    expand + band + weighted accumulate over the
//                              patches covering each mosaic pixel + normalise +
//                              clamp + uint8; nothing accumulated in HBM
//   p360_linear_collapse / p360_paste_collapse   same gather form for the
//                              linear blender and for pasting
#include "p360_common.cuh"

namespace p360 {

__device__ __forceinline__ float4 add4(const float4 &a, const float4 &b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 shfl_xor4(const float4 &v, int m) {
    return make_float4(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m),
                       __shfl_xor_sync(0xffffffffu, v.z, m), __shfl_xor_sync(0xffffffffu, v.w, m));
}
__device__ __forceinline__ float4 scale4(const float4 &v, float s) {
    return make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
}

// ---- reduce ----------------------------------------------------------------
// Block = 8 warps; warp w of block (bx, by) produces coarse row 8*by + w of D4
// (two rows of D2) for 32 full-resolution columns: lanes run along x, so each
// of the four row loads is one coalesced 512-byte request; 2x2 and 4x4 sums
// are finished with xor-shuffles.  grid.z = patch.
// Does anybody read what block (bxi, byi) of this patch's reduce grid produces?
__device__ __forceinline__ bool reduce_block_needed(const BandPatch &bp, int bxi, int byi, bool owners,
                                                    const TileMaps &maps) {
    const int bx0 = bxi * 32 - bp.pad, by0 = byi * 32 - bp.pad;           // block in patch pixels
    if (maps.need != nullptr)    // no seam within the blur chain's reach reads these cells
        return tiles_test(maps, maps.need, bp.index, bx0 + bp.x0, by0 + bp.y0, bx0 + bp.x0 + 32, by0 + bp.y0 + 32);
    // nothing within the blur chain's reach of the owned box reads these cells
    return !owners || near_owned(bp.own, 2 * bp.pad + 4, bx0, by0, bx0 + 32, by0 + 32);
}

__device__ __forceinline__ void reduce_block(const BandPatch &bp, int bxi, int byi,
                                             const unsigned long long *__restrict__ keys, int W) {
    const int w4 = bp.w4, h4 = bp.h4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cy = byi * 8 + warp;                     // D4 row
    const int xe = bxi * 32 + lane;                    // column in the extended frame
    if (cy >= h4 || bxi * 32 >= 4 * w4) return;        // warp-uniform
    const int pw = bp.pw, ph = bp.ph, pad = bp.pad, w2 = 2 * w4;
    const float4 *rgba = bp.rgba;
    const bool live = xe < 4 * w4;
    const int sx = reflect_101(xe - pad, pw);
    float4 s[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int sy = reflect_101(4 * cy + 2 * half + k - pad, ph);
            float4 v = live ? ld_stream(rgba + (size_t)sy * pw + sx) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (keys != nullptr && live)     // stitcher.py:207-208
                v.w = key_is_owner(__ldg(keys + (size_t)(sy + bp.y0) * W + (sx + bp.x0)), bp.index) ? 1.0f : 0.0f;
            acc = add4(acc, v);
        }
        s[half] = add4(acc, shfl_xor4(acc, 1));        // 2x2 sums (both lanes of a pair hold it)
    }
    if (live && !(lane & 1)) {
        const size_t c2 = (size_t)(2 * cy) * w2 + (xe >> 1);
        bp.d2[c2] = scale4(s[0], 0.25f);
        bp.d2[c2 + w2] = scale4(s[1], 0.25f);
    }
    float4 q = add4(s[0], s[1]);                       // 2 columns x 4 rows
    q = add4(q, shfl_xor4(q, 2));                      // 4 x 4
    if (live && !(lane & 3)) bp.d4[(size_t)cy * w4 + (xe >> 2)] = scale4(q, 0.0625f);
}

// dense grid: grid.z = patch, every block decides for itself
__global__ void __launch_bounds__(256)
pyramid_reduce_kernel(const BandPatch *__restrict__ patches,
                      const unsigned long long *__restrict__ keys, int W, TileMaps maps) {
    const BandPatch &bp = patches[blockIdx.z];
    if (!reduce_block_needed(bp, blockIdx.x, blockIdx.y, keys != nullptr, maps)) return;   // block-uniform
    reduce_block(bp, blockIdx.x, blockIdx.y, keys, W);
}

// Seam-band maps leave a few percent of a dense grid busy, and an empty block costs about as
// much as a busy one here.  So: one THREAD per block of the dense grid appends the needed ones
// to a work list, and a persistent grid walks the list.
__global__ void __launch_bounds__(256)
reduce_scan_kernel(const BandPatch *__restrict__ patches, int n_patches, int gx, int gy, TileMaps maps) {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= (long long)gx * gy * n_patches) return;
    const int bxi = (int)(t % gx), byi = (int)((t / gx) % gy), p = (int)(t / ((long long)gx * gy));
    const BandPatch &bp = patches[p];
    if (byi * 8 >= bp.h4 || bxi * 32 >= 4 * bp.w4) return;
    if (!reduce_block_needed(bp, bxi, byi, true, maps)) return;
    const int at = atomicAdd(maps.work_count, 1);
    if (at < maps.work_cap) maps.work[at] = make_uint2((unsigned)p, (unsigned)bxi | ((unsigned)byi << 16));
}

__global__ void __launch_bounds__(256)
pyramid_reduce_list_kernel(const BandPatch *__restrict__ patches,
                           const unsigned long long *__restrict__ keys, int W, TileMaps maps) {
    const int n = min(*maps.work_count, maps.work_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) maps.work_count[1] = n;      // (statistics for the bench)
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const uint2 item = maps.work[i];
        reduce_block(patches[item.x], (int)(item.y & 0xffffu), (int)(item.y >> 16), keys, W);
    }
}

// ---- owned boxes -------------------------------------------------------------
// One block per 64 x 32 mosaic tile: which patches own a pixel here?  (shared bitmap, up to
// 1024 patches.)  Each of them grows its box to this tile with four atomics.
// With `present_out` the bitmap is also kept per tile (seam-band maps); tile rows then start
// at window row `row0` (<= 0) so that they sit on absolute mosaic rows.
__global__ void __launch_bounds__(256)
owned_boxes_kernel(const unsigned long long *__restrict__ keys, BandPatch *patches, int n_patches,
                   int H, int W, int row0, uint32_t *__restrict__ present_out, int words,
                   const uint8_t *__restrict__ covered, uint8_t *__restrict__ unowned_out) {
    __shared__ unsigned present[32];
    __shared__ unsigned unowned;         // a valid pixel nobody owns (alpha == 0 everywhere): the tile is not
                                         // simply its owner's pixels
    const int tid = threadIdx.x;
    if (tid < 32) present[tid] = 0u;
    if (tid == 0) unowned = 0u;
    __syncthreads();
    const int tx0 = blockIdx.x * 64, ty0 = row0 + (int)blockIdx.y * 32;
    // thread -> two adjacent keys (one 128-bit load) on 4 rows; runs of equal owners cost one
    // shared-memory atomic
    const int cx = tx0 + 2 * (tid & 31), ry = ty0 + (tid >> 5);
    unsigned last = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int y = ry + 8 * i;
        if (y >= H) break;
        if (y < 0) continue;
        unsigned long long k2[2] = {0ull, 0ull};
        const unsigned long long *src = keys + (size_t)y * W + cx;
        if (cx + 1 < W && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
            const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(src));
            k2[0] = v.x; k2[1] = v.y;
        } else {
            if (cx < W) k2[0] = __ldg(src);
            if (cx + 1 < W) k2[1] = __ldg(src + 1);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (k2[j] == 0ull) {
                if (covered != nullptr && cx + j < W && covered[(size_t)y * W + cx + j] != 0) unowned = 1u;
                continue;
            }
            const unsigned p = 0xFFFFFFFFu - (unsigned)(k2[j] & 0xFFFFFFFFull);
            if (p != last && p < (unsigned)n_patches && p < 1024u) {
                atomicOr(&present[p >> 5], 1u << (p & 31));
                last = p;
            }
        }
    }
    __syncthreads();
    if (present_out != nullptr && tid < words)
        present_out[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * words + tid] = present[tid];
    if (unowned_out != nullptr && tid == 0) unowned_out[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = (uint8_t)unowned;
    if (tid < 32) {
        unsigned word = present[tid];
        while (word) {
            const int bit = __ffs(word) - 1;
            word &= word - 1;
            BandPatch &bp = patches[tid * 32 + bit];
            atomicMin(&bp.own[0], max(tx0 - bp.x0, 0));
            atomicMin(&bp.own[1], max(max(ty0, 0) - bp.y0, 0));
            atomicMax(&bp.own[2], min(tx0 + 64 - bp.x0, bp.pw));
            atomicMax(&bp.own[3], min(ty0 + 32 - bp.y0, bp.ph));
        }
    }
}

// ---- seam-band maps ------------------------------------------------------------
// cand(T) = OR of present over the tiles within blur reach of T; multi(T) = |cand(T)| > 1.
// One thread per tile.
__global__ void __launch_bounds__(256)
tile_candidates_kernel(TileMaps m) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= m.tiles_x * m.tiles_y) return;
    const int tx = t % m.tiles_x, ty = t / m.tiles_x;
    const int x0 = max(tx - m.reach_x, 0), x1 = min(tx + m.reach_x, m.tiles_x - 1);
    const int y0 = max(ty - m.reach_y, 0), y1 = min(ty + m.reach_y, m.tiles_y - 1);
    int count = 0;
    for (int w = 0; w < m.words; ++w) {
        uint32_t bits = 0u;
        for (int y = y0; y <= y1; ++y)
            for (int x = x0; x <= x1; ++x) bits |= __ldg(m.present + ((size_t)y * m.tiles_x + x) * m.words + w);
        m.cand[(size_t)t * m.words + w] = bits;
        count += __popc(bits);
    }
    // on entry multi[t] = "holds a valid pixel nobody owns": such a tile is blended in full too
    m.multi[t] = (count > 1 || (count == 1 && m.multi[t] != 0)) ? 1 : 0;
}

// need(T) = OR of cand over the multi tiles within blur reach of T.
__global__ void __launch_bounds__(256)
tile_needs_kernel(TileMaps m) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= m.tiles_x * m.tiles_y) return;
    const int tx = t % m.tiles_x, ty = t / m.tiles_x;
    const int x0 = max(tx - m.reach_x, 0), x1 = min(tx + m.reach_x, m.tiles_x - 1);
    const int y0 = max(ty - m.reach_y, 0), y1 = min(ty + m.reach_y, m.tiles_y - 1);
    for (int w = 0; w < m.words; ++w) {
        uint32_t bits = 0u;
        for (int y = y0; y <= y1; ++y)
            for (int x = x0; x <= x1; ++x) {
                const size_t n = (size_t)y * m.tiles_x + x;
                if (m.multi[n]) bits |= m.cand[n * m.words + w];
            }
        m.need[(size_t)t * m.words + w] = bits;
    }
}

// ---- tile lists -------------------------------------------------------------
constexpr int CT_X = 64, CT_Y = 32;     // mosaic tile per block; block = 64 x 4 threads, 8 rows each
constexpr int MAX_TILE_PATCHES = 1024;  // patches that may overlap one tile

// Ordered list (patch order = accumulation order, stitcher.py:223) of the
// patches whose box intersects this block's tile, built cooperatively in
// shared memory with ballots so that no host-side tile lists are needed.
template <bool SUPPORT>
__device__ int build_tile_list(const BandPatch *__restrict__ patches, int n_patches,
                               int tx0, int ty0, int16_t *list) {
    __shared__ int warp_hits[8];
    __shared__ int total;
    const int tid = threadIdx.y * CT_X + threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) total = 0;
    __syncthreads();
    for (int base = 0; base < n_patches; base += 256) {
        const int t = base + tid;
        bool hit = false;
        if (t < n_patches) {
            const BandPatch &bp = patches[t];
            hit = bp.x0 < tx0 + CT_X && bp.x0 + bp.pw > tx0 && bp.y0 < ty0 + CT_Y && bp.y0 + bp.ph > ty0;
            if (SUPPORT && hit)      // weights vanish farther than `pad` from the owned box
                hit = near_owned(bp.own, bp.pad, tx0 - bp.x0, ty0 - bp.y0, tx0 + CT_X - bp.x0, ty0 + CT_Y - bp.y0);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) warp_hits[warp] = __popc(ballot);
        __syncthreads();
        int offset = total;
        for (int w = 0; w < warp; ++w) offset += warp_hits[w];
        offset += __popc(ballot & ((1u << lane) - 1));
        if (hit && offset < MAX_TILE_PATCHES) list[offset] = (int16_t)t;
        __syncthreads();
        if (tid == 0) {
            int sum = total;
            for (int w = 0; w < 8; ++w) sum += warp_hits[w];
            total = min(sum, MAX_TILE_PATCHES);
        }
        __syncthreads();
    }
    return total;
}

// The same list from the tile's candidate bitmap (seam-band maps); ascending bits = patch order.
__device__ int tile_list_from_maps(const BandPatch *__restrict__ patches, int n_patches, const TileMaps &m,
                                   size_t tile, int tx0, int ty0, int16_t *list) {
    __shared__ int total;
    if (threadIdx.y == 0 && threadIdx.x == 0) {
        int n = 0;
        for (int w = 0; w < m.words; ++w) {
            uint32_t bits = __ldg(m.cand + tile * m.words + w);
            while (bits) {
                const int t = 32 * w + __ffs(bits) - 1;
                bits &= bits - 1;
                if (t >= n_patches) break;
                const BandPatch &bp = patches[t];
                if (bp.x0 < tx0 + CT_X && bp.x0 + bp.pw > tx0 && bp.y0 < ty0 + CT_Y && bp.y0 + bp.ph > ty0 &&
                    near_owned(bp.own, bp.pad, tx0 - bp.x0, ty0 - bp.y0, tx0 + CT_X - bp.x0, ty0 + CT_Y - bp.y0))
                    list[n++] = (int16_t)t;
            }
        }
        total = n;
    }
    __syncthreads();
    return total;
}

// Position of full-res patch coordinate p on a coarse grid of factor 2^SHIFT
// anchored `pad` pixels before the patch: u = (p + pad + 0.5) / f - 0.5.
template <int SHIFT>
__device__ __forceinline__ void coarse_coord(int pad, int p, int &i, float &frac) {
    constexpr int F = 1 << SHIFT;
    const int n = 2 * (p + pad) + 1 - F;               // u = n / (2f)
    i = n >> (SHIFT + 1);
    frac = (float)(n & (2 * F - 1)) * (0.5f / F);
}

// Drop from the tile list every patch whose blurred mask is identically zero
// over the tile: its weights vanish at every level (the supports of the
// truncated Gaussians nest and all taps are positive), so it contributes
// exact zeros.  Tested on the coarse alpha of the widest level.
template <int L>
__device__ int cull_tile_list(const BandPatch *__restrict__ patches, int n_hit, int tx0, int ty0,
                              int16_t *list) {
    if (L < 2) return n_hit;
    constexpr int SHIFT = (L == 2) ? 1 : 2;
    const int tid = threadIdx.y * CT_X + threadIdx.x;
    int kept = 0;
    for (int it = 0; it < n_hit; ++it) {
        const int id = list[it];
        const BandPatch &bp = patches[id];
        const int lw = bp.w4 * (SHIFT == 1 ? 2 : 1);
        // tile ∩ patch box ∩ support (owned box grown by pad): only there were the levels computed
        const int px0 = max(max(tx0 - bp.x0, 0), bp.own[0] - bp.pad);
        const int px1 = min(min(tx0 + CT_X - bp.x0, bp.pw), bp.own[2] + bp.pad) - 1;
        const int py0 = max(max(ty0 - bp.y0, 0), bp.own[1] - bp.pad);
        const int py1 = min(min(ty0 + CT_Y - bp.y0, bp.ph), bp.own[3] + bp.pad) - 1;
        int ix0, ix1, iy0, iy1;
        float unused;
        coarse_coord<SHIFT>(bp.pad, px0, ix0, unused); coarse_coord<SHIFT>(bp.pad, px1, ix1, unused);
        coarse_coord<SHIFT>(bp.pad, py0, iy0, unused); coarse_coord<SHIFT>(bp.pad, py1, iy1, unused);
        const int nx = ix1 + 2 - ix0, ny = iy1 + 2 - iy0;
        const float *alpha = reinterpret_cast<const float *>(bp.low[L - 2]) + 3;
        bool any = false;
        for (int i = tid; i < nx * ny; i += 256) {
            const int cx = ix0 + i % nx, cy = iy0 + i / nx;
            any |= __ldg(alpha + 4 * ((size_t)cy * lw + cx)) != 0.0f;
        }
        const bool keep = __syncthreads_or(any);
        if (keep) {
            if (tid == 0) list[kept] = (int16_t)id;     // kept <= it: never overtakes the read position
            ++kept;
        }
    }
    __syncthreads();
    return kept;
}

// ---- multiband collapse (gather form) ---------------------------------------
struct Pair2 { float2 lo, hi; };       // RGBA as two packed halves for fma.rn.f32x2

__device__ __forceinline__ Pair2 to_pair(const float4 &v) {
    Pair2 p; p.lo = make_float2(v.x, v.y); p.hi = make_float2(v.z, v.w); return p;
}
// bilinear sample ("expand") of a coarse image: a*w00 + b*w01 + c*w10 + d*w11
__device__ __forceinline__ Pair2 expand_at(const float4 *__restrict__ p, int lw, float w00, float w01,
                                           float w10, float w11) {
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + lw), d = __ldg(p + lw + 1);
    const float2 k00 = make_float2(w00, w00), k01 = make_float2(w01, w01);
    const float2 k10 = make_float2(w10, w10), k11 = make_float2(w11, w11);
    Pair2 o;
    o.lo = __fmul2_rn(make_float2(a.x, a.y), k00);
    o.hi = __fmul2_rn(make_float2(a.z, a.w), k00);
    o.lo = __ffma2_rn(make_float2(b.x, b.y), k01, o.lo);
    o.hi = __ffma2_rn(make_float2(b.z, b.w), k01, o.hi);
    o.lo = __ffma2_rn(make_float2(c.x, c.y), k10, o.lo);
    o.hi = __ffma2_rn(make_float2(c.z, c.w), k10, o.hi);
    o.lo = __ffma2_rn(make_float2(d.x, d.y), k11, o.lo);
    o.hi = __ffma2_rn(make_float2(d.z, d.w), k11, o.hi);
    return o;
}

constexpr int ROWS_PER_THREAD = CT_Y / 4;

template <int L>
__global__ void __launch_bounds__(256, (L <= 5) ? 4 : 3)
multiband_collapse_kernel(const BandPatch *__restrict__ patches, int n_patches,
                          const unsigned long long *__restrict__ keys,
                          const uint8_t *__restrict__ covered, uint8_t *__restrict__ out, int OW, int first_tile_row,
                          int y_begin, int H, int first_tile_col, int x_end, int W, TileMaps maps) {
    // Tiles are anchored at absolute mosaic rows (first_tile_row may be < y_begin, even < 0):
    // whether a tile takes the single-contributor shortcut must not depend on how the
    // mosaic was cut into strips or row bands.
    __shared__ int16_t list[MAX_TILE_PATCHES];
    const int tile_col = first_tile_col + (int)blockIdx.x;          // (columns [64 * first_tile_col, x_end) are produced)
    const int tx0 = tile_col * CT_X, ty0 = first_tile_row + blockIdx.y * CT_Y;
    int n_hit;
    bool pure = false;      // every valid pixel of the tile belongs to list[0]: the owner keys are not needed
    if (maps.cand != nullptr) {
        // a single candidate: every valid pixel of the tile is its own (p360_tile_maps_build), no
        // coarse level is read and none was computed here — nothing to cull either
        const size_t tile = (size_t)((ty0 - maps.row0) >> 5) * maps.tiles_x + tile_col;
        const bool blended = __ldg(maps.multi + tile) != 0;
        if (maps.wneed != nullptr && !blended) return;   // seam plan: p360_warp_tiles wrote this tile (block-uniform)
        n_hit = tile_list_from_maps(patches, n_patches, maps, tile, tx0, ty0, list);
        if (n_hit > 1 || blended) n_hit = cull_tile_list<L>(patches, n_hit, tx0, ty0, list);
        pure = L > 1 && n_hit == 1 && !blended;
    } else {
        n_hit = build_tile_list<(L > 1)>(patches, n_patches, tx0, ty0, list);
        n_hit = cull_tile_list<L>(patches, n_hit, tx0, ty0, list);
    }
    {   // pull this tile's slice of every contributing patch (and of the owner keys) towards L2
        // now: one 128-byte line per thread and patch, so that the gathers below find their
        // streamed operands on chip instead of paying a DRAM round trip per patch iteration
        const int tid = threadIdx.y * CT_X + threadIdx.x;
        const int trow = ty0 + (tid >> 3), tcol = tx0 + 8 * (tid & 7);
        if (!pure && trow >= 0 && trow < H && tcol < W) prefetch_l2(keys + (size_t)trow * W + tcol);
        for (int it = 0; it < n_hit; ++it) {
            const BandPatch &bp = patches[list[it]];
            const int px = tcol - bp.x0, py = trow - bp.y0;
            if ((unsigned)py < (unsigned)bp.ph && px > -8 && px < bp.pw)
                prefetch_l2(bp.rgba + (size_t)py * bp.pw + max(px, 0));
        }
    }
    const int X = tx0 + threadIdx.x;
    if (X >= x_end) return;
    for (int sub = 0; sub < ROWS_PER_THREAD; ++sub) {
        const int Y = ty0 + threadIdx.y + 4 * sub;
        if (Y >= H) break;
        if (Y < y_begin) continue;
        const size_t mi = (size_t)Y * W + X;
        // per level: lo = (sum band.x*w, sum band.y*w), hi = (sum band.z*w, sum w)
        float2 lo[L], hi[L];
#pragma unroll
        for (int l = 0; l < L; ++l) lo[l] = hi[l] = make_float2(0.f, 0.f);
        if (covered[mi] != 0) {
            const unsigned long long key = pure ? 0ull : __ldg(keys + mi);
            if (n_hit == 1 && L > 1) {
                // Only one patch has non-zero weights anywhere in this tile.  Where it also
                // owns the pixel every level weight is > 0, so sum_l band_l * w_l / w_l
                // telescopes to the warped pixel itself (I - B0 + B0 - B1 + ... + B_{L-2}):
                // skip the pyramid, exact up to float rounding (~1e-7).
                const BandPatch &bp = patches[list[0]];
                const int px = X - bp.x0, py = Y - bp.y0;
                if ((unsigned)px < (unsigned)bp.pw && (unsigned)py < (unsigned)bp.ph &&
                    (pure || key_is_owner(key, bp.index))) {
                    const float4 pix = ld_stream(bp.rgba + (size_t)py * bp.pw + px);
                    uint8_t *o = out + ((size_t)Y * OW + X) * 3;
                    o[0] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, fminf(fmaxf(pix.x, 0.f), 1.f)));
                    o[1] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, fminf(fmaxf(pix.y, 0.f), 1.f)));
                    o[2] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, fminf(fmaxf(pix.z, 0.f), 1.f)));
                    continue;
                }
            }
            // (a valid pixel of a pure tile lies inside its owner's box and was written above)
            for (int it = 0; it < (pure ? 0 : n_hit); ++it) {     // patch order = list order
                const BandPatch &bp = patches[list[it]];
                const int px = X - bp.x0, py = Y - bp.y0;
                if ((unsigned)px >= (unsigned)bp.pw || (unsigned)py >= (unsigned)bp.ph) continue;
                if (L > 1 && (px < bp.own[0] - bp.pad || px >= bp.own[2] + bp.pad ||
                              py < bp.own[1] - bp.pad || py >= bp.own[3] + bp.pad)) continue;   // all weights 0
                const float4 pix = ld_stream(bp.rgba + (size_t)py * bp.pw + px);
                Pair2 prev = to_pair(pix);
                prev.hi.y = key_is_owner(key, bp.index) ? 1.0f : 0.0f;     // stitcher.py:207-208
                const int pad = bp.pad, w4 = bp.w4;
                // bilinear weights on the f = 2 grid (level 0) and the f = 4 grid (levels >= 1)
                int ix2, iy2, ix4, iy4;
                float fx2, fy2, fx4, fy4;
                coarse_coord<1>(pad, px, ix2, fx2); coarse_coord<1>(pad, py, iy2, fy2);
                coarse_coord<2>(pad, px, ix4, fx4); coarse_coord<2>(pad, py, iy4, fy4);
                const size_t off2 = (size_t)iy2 * (2 * w4) + ix2, off4 = (size_t)iy4 * w4 + ix4;
                const float a2 = (1.0f - fy2) * (1.0f - fx2), b2 = (1.0f - fy2) * fx2;
                const float c2 = fy2 * (1.0f - fx2), d2 = fy2 * fx2;
                const float a4 = (1.0f - fy4) * (1.0f - fx4), b4 = (1.0f - fy4) * fx4;
                const float c4 = fy4 * (1.0f - fx4), d4 = fy4 * fx4;
#pragma unroll
                for (int l = 0; l < L - 1; ++l) {             // stitcher.py:224-232
                    const Pair2 cur = (l == 0) ? expand_at(bp.low[0] + off2, 2 * w4, a2, b2, c2, d2)
                                               : expand_at(bp.low[l] + off4, w4, a4, b4, c4, d4);
                    const float2 ww = make_float2(cur.hi.y, cur.hi.y);      // weight = blurred mask
                    const float2 dlo = __fadd2_rn(prev.lo, make_float2(-cur.lo.x, -cur.lo.y));
                    const float2 dhi = make_float2(prev.hi.x - cur.hi.x, 1.0f);
                    lo[l] = __ffma2_rn(dlo, ww, lo[l]);
                    hi[l] = __ffma2_rn(dhi, ww, hi[l]);
                    prev = cur;
                }
                const float2 ww = make_float2(prev.hi.y, prev.hi.y);
                lo[L - 1] = __ffma2_rn(prev.lo, ww, lo[L - 1]);
                hi[L - 1] = __ffma2_rn(make_float2(prev.hi.x, 1.0f), ww, hi[L - 1]);
            }
        }
        float m0 = 0.f, m1 = 0.f, m2 = 0.f;                   // stitcher.py:236-238
#pragma unroll
        for (int l = 0; l < L; ++l) {
            // one reciprocal per level instead of three correctly rounded divisions:
            // the coarse evaluation is ~1e-3 accurate, a 1-ulp quotient is noise
            const float inv = hi[l].y == 0.0f ? 1.0f : __frcp_rn(hi[l].y);
            m0 = fmaf(lo[l].x, inv, m0);
            m1 = fmaf(lo[l].y, inv, m1);
            m2 = fmaf(hi[l].x, inv, m2);
        }
        uint8_t *o = out + ((size_t)Y * OW + X) * 3;          // stitcher.py:240-241
        o[0] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, fminf(fmaxf(m0, 0.f), 1.f)));
        o[1] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, fminf(fmaxf(m1, 0.f), 1.f)));
        o[2] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, fminf(fmaxf(m2, 0.f), 1.f)));
    }
}

// ---- linear blend / paste (gather form) ------------------------------------
// MODE 0: linear_blend (stitcher.py:171-183): sum alpha*rgb / sum alpha in patch order.
// MODE 1: no_blend (stitcher.py:160-168): the last valid writer wins.
template <int MODE>
__global__ void __launch_bounds__(256)
pointwise_collapse_kernel(const BandPatch *__restrict__ patches, int n_patches,
                          uint8_t *__restrict__ out, int OW, int first_tile_row, int y_begin, int H, int W) {
    __shared__ int16_t list[MAX_TILE_PATCHES];
    const int tx0 = blockIdx.x * CT_X, ty0 = first_tile_row + blockIdx.y * CT_Y;
    const int n_hit = build_tile_list<false>(patches, n_patches, tx0, ty0, list);
    const int X = tx0 + threadIdx.x;
    if (X >= W) return;
    for (int sub = 0; sub < ROWS_PER_THREAD; ++sub) {
        const int Y = ty0 + threadIdx.y + 4 * sub;
        if (Y >= H) break;
        if (Y < y_begin) continue;
        const size_t mi = (size_t)Y * W + X;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, wsum = 0.f;
        for (int it = 0; it < n_hit; ++it) {
            const BandPatch &bp = patches[list[it]];
            const int px = X - bp.x0, py = Y - bp.y0;
            if ((unsigned)px >= (unsigned)bp.pw || (unsigned)py >= (unsigned)bp.ph) continue;
            const size_t pi = (size_t)py * bp.pw + px;
            const bool bad = __ldg(bp.invalid + pi) != 0;
            if (MODE == 1) {
                if (bad) continue;
                const float4 p = ld_stream(bp.rgba + pi);
                a0 = p.x; a1 = p.y; a2 = p.z;
            } else {
                const float4 p = ld_stream(bp.rgba + pi);
                a0 = __fadd_rn(a0, __fmul_rn(bad ? 0.f : p.x, p.w));
                a1 = __fadd_rn(a1, __fmul_rn(bad ? 0.f : p.y, p.w));
                a2 = __fadd_rn(a2, __fmul_rn(bad ? 0.f : p.z, p.w));
                wsum = __fadd_rn(wsum, p.w);
            }
        }
        if (MODE == 0) {
            const float w = wsum == 0.0f ? 1.0f : wsum;
            a0 = __fdiv_rn(a0, w); a1 = __fdiv_rn(a1, w); a2 = __fdiv_rn(a2, w);
        }
        // (255 * v).astype(uint8): truncation, no clip in either reference blender
        uint8_t *o = out + ((size_t)Y * OW + X) * 3;
        o[0] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, a0));
        o[1] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, a1));
        o[2] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, a2));
    }
}

inline int first_tile(int y_begin, int row_origin) {     // absolute-row-aligned tile containing y_begin
    int phase = (y_begin + row_origin) % CT_Y;
    if (phase < 0) phase += CT_Y;
    return y_begin - phase;
}

template <int L>
int launch_collapse(const BandPatch *patches, int n_patches, const unsigned long long *keys,
                    const uint8_t *covered, uint8_t *out, int OW, int y0, int y1, int x0, int x1, int row_origin, int W,
                    const TileMaps &maps, cudaStream_t s) {
    const int first = first_tile(y0, row_origin);
    dim3 grid(cdiv(x1 - x0, CT_X), cdiv(y1 - first, CT_Y)), block(CT_X, 4);
    multiband_collapse_kernel<L><<<grid, block, 0, s>>>(patches, n_patches, keys, covered, out, OW, first, y0, y1,
                                                        x0 / CT_X, x1, W, maps);
    return check_launch("p360_multiband_collapse");
}

}  // namespace p360

using namespace p360;

static TileMaps device_maps(const p360_tile_maps *maps_host) {
    TileMaps m;
    memset(&m, 0, sizeof(m));
    if (maps_host != nullptr) memcpy(&m, maps_host, sizeof(m));
    return m;
}
static bool maps_ok(const p360_tile_maps *m) {
    return m == nullptr || (m->present && m->cand && m->need && m->multi && m->work && m->work_count &&
                            (reinterpret_cast<uintptr_t>(m->work) & 7) == 0 &&
                            m->work_cap > 0 && m->tiles_x > 0 && m->tiles_y > 0 &&
                            m->words > 0 && m->words <= 32 && m->row0 <= 0 && m->row0 > -32 &&
                            m->reach_x >= 0 && m->reach_y >= 0);
}

extern "C" int p360_owned_boxes(const uint64_t *owner_keys, p360_band_patch *patches, int n_patches,
                                int H, int W, void *stream) {
    const char *where = "p360_owned_boxes";
    P360_REQUIRE(owner_keys && patches && n_patches >= 0 && n_patches <= MAX_TILE_PATCHES && H > 0 && W > 0, where);
    if (n_patches == 0) return 0;
    dim3 grid(cdiv(W, 64), cdiv(H, 32));
    P360_REQUIRE(grid.y <= 65535, where);
    owned_boxes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const unsigned long long *>(owner_keys), reinterpret_cast<BandPatch *>(patches),
        n_patches, H, W, 0, nullptr, 0, nullptr, nullptr);
    return check_launch(where);
}

extern "C" int p360_tile_maps_build(const uint64_t *owner_keys, const uint8_t *covered, p360_band_patch *patches,
                                    int n_patches, int H, int W, const p360_tile_maps *maps_host, void *stream) {
    const char *where = "p360_tile_maps_build";
    P360_REQUIRE(owner_keys && covered && patches && maps_host && maps_ok(maps_host), where);
    P360_REQUIRE(n_patches > 0 && n_patches <= MAX_TILE_PATCHES && H > 0 && W > 0, where);
    const TileMaps m = device_maps(maps_host);
    P360_REQUIRE(m.words == (n_patches + 31) / 32 && m.tiles_x == (int)cdiv(W, TILE_X) &&
                 m.tiles_y == (int)cdiv(H - m.row0, TILE_Y) && m.tiles_y <= 65535, where);
    cudaStream_t s = (cudaStream_t)stream;
    owned_boxes_kernel<<<dim3(m.tiles_x, m.tiles_y), 256, 0, s>>>(
        reinterpret_cast<const unsigned long long *>(owner_keys), reinterpret_cast<BandPatch *>(patches),
        n_patches, H, W, m.row0, m.present, m.words, covered, m.multi);
    if (int e = check_launch(where)) return e;
    const unsigned blocks = cdiv((long long)m.tiles_x * m.tiles_y, 256);
    tile_candidates_kernel<<<blocks, 256, 0, s>>>(m);
    if (int e = check_launch(where)) return e;
    tile_needs_kernel<<<blocks, 256, 0, s>>>(m);
    return check_launch(where);
}

extern "C" int p360_pyramid_dims(int pw, int ph, int pad, int32_t out_host[4]) {
    const char *where = "p360_pyramid_dims";
    P360_REQUIRE(out_host && pw > 0 && ph > 0 && pad >= 0 && pad % 4 == 0, where);
    int w4 = (pw + 2 * pad + 3) / 4, h4 = (ph + 2 * pad + 3) / 4;
    out_host[0] = 2 * w4; out_host[1] = 2 * h4; out_host[2] = w4; out_host[3] = h4;
    return 0;
}

extern "C" int p360_pyramid_reduce_batch(const p360_band_patch *patches, int n_patches, int max_w4,
                                         int max_h4, const uint64_t *owner_keys, int W,
                                         const p360_tile_maps *maps_host, void *stream) {
    const char *where = "p360_pyramid_reduce_batch";
    P360_REQUIRE(patches && n_patches >= 0 && n_patches <= 65535 && max_w4 >= 0 && max_h4 >= 0, where);
    P360_REQUIRE(owner_keys == nullptr || W > 0, where);
    P360_REQUIRE(maps_ok(maps_host) && (maps_host == nullptr || owner_keys != nullptr), where);
    if (n_patches == 0 || max_w4 == 0 || max_h4 == 0) return 0;
    dim3 grid(cdiv(4 * max_w4, 32), cdiv(max_h4, 8), n_patches);
    P360_REQUIRE(grid.y <= 65535, where);
    auto bp = reinterpret_cast<const BandPatch *>(patches);
    auto keys = reinterpret_cast<const unsigned long long *>(owner_keys);
    cudaStream_t s = (cudaStream_t)stream;
    const TileMaps maps = device_maps(maps_host);
    if (maps_host == nullptr) {
        pyramid_reduce_kernel<<<grid, 256, 0, s>>>(bp, keys, W, maps);
        return check_launch(where);
    }
    const long long blocks = (long long)grid.x * grid.y * grid.z;
    P360_REQUIRE(blocks <= maps.work_cap && grid.x <= 65535, where);
    P360_CUDA(cudaMemsetAsync(maps.work_count, 0, sizeof(int), s), where);
    reduce_scan_kernel<<<cdiv(blocks, 256), 256, 0, s>>>(bp, n_patches, (int)grid.x, (int)grid.y, maps);
    if (int e = check_launch(where)) return e;
    pyramid_reduce_list_kernel<<<persistent_blocks(8), 256, 0, s>>>(bp, keys, W, maps);
    return check_launch(where);
}

extern "C" int p360_multiband_collapse(const p360_band_patch *patches, int n_patches, int n_levels,
                                       const uint64_t *owner_keys, const uint8_t *covered,
                                       uint8_t *out_u8, int out_pitch, int y_begin, int y_end, int x_begin, int x_end,
                                       int row_origin, int W, const p360_tile_maps *maps_host, void *stream) {
    const char *where = "p360_multiband_collapse";
    const int OW = out_pitch ? out_pitch : W;
    P360_REQUIRE(maps_ok(maps_host), where);
    const TileMaps maps = device_maps(maps_host);
    if (maps_host != nullptr) {         // the maps' tile grid must be the collapse's
        int phase = row_origin % 32;
        if (phase < 0) phase += 32;
        P360_REQUIRE(maps.row0 == -phase && maps.tiles_x == (int)cdiv(W, TILE_X) &&
                     y_end <= maps.row0 + 32 * maps.tiles_y, where);
    }
    P360_REQUIRE(patches && owner_keys && covered && out_u8, where);
    P360_REQUIRE(n_patches >= 0 && n_patches <= MAX_TILE_PATCHES, where);
    P360_REQUIRE(n_levels >= 1 && n_levels <= P360_MAX_LEVELS && W > 0, where);
    P360_REQUIRE(y_begin >= 0 && y_end >= y_begin, where);
    P360_REQUIRE(x_begin >= 0 && x_begin <= x_end && x_end <= W && x_begin % CT_X == 0 && OW >= x_end, where);
    if (y_end == y_begin || x_end == x_begin) return 0;
    const int H = y_end;
    auto bp = reinterpret_cast<const BandPatch *>(patches);
    auto keys = reinterpret_cast<const unsigned long long *>(owner_keys);
    cudaStream_t s = (cudaStream_t)stream;
    switch (n_levels) {
        case 1: return launch_collapse<1>(bp, n_patches, keys, covered, out_u8, OW, y_begin, H, x_begin, x_end, row_origin, W, maps, s);
        case 2: return launch_collapse<2>(bp, n_patches, keys, covered, out_u8, OW, y_begin, H, x_begin, x_end, row_origin, W, maps, s);
        case 3: return launch_collapse<3>(bp, n_patches, keys, covered, out_u8, OW, y_begin, H, x_begin, x_end, row_origin, W, maps, s);
        case 4: return launch_collapse<4>(bp, n_patches, keys, covered, out_u8, OW, y_begin, H, x_begin, x_end, row_origin, W, maps, s);
        case 5: return launch_collapse<5>(bp, n_patches, keys, covered, out_u8, OW, y_begin, H, x_begin, x_end, row_origin, W, maps, s);
        case 6: return launch_collapse<6>(bp, n_patches, keys, covered, out_u8, OW, y_begin, H, x_begin, x_end, row_origin, W, maps, s);
        case 7: return launch_collapse<7>(bp, n_patches, keys, covered, out_u8, OW, y_begin, H, x_begin, x_end, row_origin, W, maps, s);
        default: return launch_collapse<8>(bp, n_patches, keys, covered, out_u8, OW, y_begin, H, x_begin, x_end, row_origin, W, maps, s);
    }
}

static int pointwise_collapse(const char *where, int mode, const p360_band_patch *patches, int n_patches,
                              uint8_t *out_u8, int out_pitch, int y_begin, int y_end, int row_origin, int W, void *stream) {
    const int OW = out_pitch ? out_pitch : W;
    P360_REQUIRE(patches && out_u8 && n_patches >= 0 && n_patches <= MAX_TILE_PATCHES && W > 0 && OW >= W, where);
    P360_REQUIRE(y_begin >= 0 && y_end >= y_begin, where);
    if (y_end == y_begin) return 0;
    const int first = first_tile(y_begin, row_origin);
    dim3 grid(cdiv(W, CT_X), cdiv(y_end - first, CT_Y)), block(CT_X, 4);
    auto bp = reinterpret_cast<const BandPatch *>(patches);
    if (mode == 0)
        pointwise_collapse_kernel<0><<<grid, block, 0, (cudaStream_t)stream>>>(bp, n_patches, out_u8, OW, first, y_begin, y_end, W);
    else
        pointwise_collapse_kernel<1><<<grid, block, 0, (cudaStream_t)stream>>>(bp, n_patches, out_u8, OW, first, y_begin, y_end, W);
    return check_launch(where);
}

extern "C" int p360_linear_collapse(const p360_band_patch *patches, int n_patches, uint8_t *out_u8, int out_pitch,
                                    int y_begin, int y_end, int row_origin, int W, void *stream) {
    return pointwise_collapse("p360_linear_collapse", 0, patches, n_patches, out_u8, out_pitch, y_begin, y_end,
                              row_origin, W, stream);
}

extern "C" int p360_paste_collapse(const p360_band_patch *patches, int n_patches, uint8_t *out_u8, int out_pitch,
                                   int y_begin, int y_end, int row_origin, int W, void *stream) {
    return pointwise_collapse("p360_paste_collapse", 1, patches, n_patches, out_u8, out_pitch, y_begin, y_end,
                              row_origin, W, stream);
}
