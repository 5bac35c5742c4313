// Reduced-resolution evaluation of the reference's wide Gaussians and the
// output-stationary band blend that consumes them.
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
//   p360_pyramid_reduce      full-res RGBA (+ owner map) -> D2, D4    ("reduce")
//   p360_gauss_blur          coarse blur (p360_blur.cu)
//   p360_multiband_collapse  expand + band + weighted accumulate over the
//                            patches covering each mosaic pixel + normalise +
//                            clamp + uint8, nothing accumulated in HBM
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
// are finished with xor-shuffles.
__global__ void __launch_bounds__(256)
pyramid_reduce_kernel(const float4 *__restrict__ rgba, int pw, int ph, int x0, int y0, int idx,
                      const int32_t *__restrict__ owner, int W, int pad,
                      float4 *__restrict__ d2, float4 *__restrict__ d4, int w4, int h4) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cy = blockIdx.y * 8 + warp;              // D4 row
    const int xe = blockIdx.x * 32 + lane;             // column in the extended frame
    if (cy >= h4) return;                              // warp-uniform
    const int w2 = 2 * w4;
    const bool live = xe < 4 * w4;
    const int sx = reflect_101(xe - pad, pw);
    float4 s[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            int sy = reflect_101(4 * cy + 2 * half + k - pad, ph);
            float4 v = live ? ld_stream(rgba + (size_t)sy * pw + sx) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (owner != nullptr && live)
                v.w = (__ldg(owner + (size_t)(sy + y0) * W + (sx + x0)) == idx) ? 1.0f : 0.0f;
            acc = add4(acc, v);
        }
        s[half] = add4(acc, shfl_xor4(acc, 1));        // 2x2 sums (both lanes of a pair hold it)
    }
    if (live && !(lane & 1)) {
        size_t c2 = (size_t)(2 * cy) * w2 + (xe >> 1);
        d2[c2] = scale4(s[0], 0.25f);
        d2[c2 + w2] = scale4(s[1], 0.25f);
    }
    float4 q = add4(s[0], s[1]);                       // 2 columns x 4 rows
    q = add4(q, shfl_xor4(q, 2));                      // 4 x 4
    if (live && !(lane & 3)) d4[(size_t)cy * w4 + (xe >> 2)] = scale4(q, 0.0625f);
}

// ---- collapse (gather form) ------------------------------------------------
struct BandPatch {
    const float4 *rgba;                 // full-res patch, alpha ignored (owner map decides)
    const float4 *low[P360_MAX_LEVELS - 1];   // blurred coarse image of level l
    int lw[P360_MAX_LEVELS - 1];        // coarse width of level l
    int shift[P360_MAX_LEVELS - 1];     // log2 of the reduction factor of level l
    int x0, y0, pw, ph;                 // box in (window) mosaic pixels
    int pad;                            // extension R in full-res pixels
    int index;                          // value stored in the owner map for this patch
};
static_assert(sizeof(BandPatch) == sizeof(p360_band_patch), "ABI struct mismatch");

// Position of full-res patch pixel (px, py) on a coarse grid of factor 2^shift
// anchored `pad` pixels before the patch: u = (p + pad + 0.5) / f - 0.5.
struct CoarseTap {
    int ix, iy;
    float fx, fy;
};
__device__ __forceinline__ CoarseTap coarse_tap(int shift, int pad, int px, int py) {
    const int f = 1 << shift;
    const int nx = 2 * (px + pad) + 1 - f, ny = 2 * (py + pad) + 1 - f;   // u = n / (2f)
    const float inv = 0.5f / (float)f;
    CoarseTap t;
    t.ix = nx >> (shift + 1);
    t.iy = ny >> (shift + 1);
    t.fx = (float)(nx & (2 * f - 1)) * inv;
    t.fy = (float)(ny & (2 * f - 1)) * inv;
    return t;
}
// bilinear sample of a coarse level ("expand")
__device__ __forceinline__ float4 expand_at(const float4 *__restrict__ low, int lw, const CoarseTap &t) {
    const float4 *p = low + (size_t)t.iy * lw + t.ix;
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + lw), d = __ldg(p + lw + 1);
    const float gx = 1.0f - t.fx, gy = 1.0f - t.fy;
    float4 o;
    o.x = (a.x * gx + b.x * t.fx) * gy + (c.x * gx + d.x * t.fx) * t.fy;
    o.y = (a.y * gx + b.y * t.fx) * gy + (c.y * gx + d.y * t.fx) * t.fy;
    o.z = (a.z * gx + b.z * t.fx) * gy + (c.z * gx + d.z * t.fx) * t.fy;
    o.w = (a.w * gx + b.w * t.fx) * gy + (c.w * gx + d.w * t.fx) * t.fy;
    return o;
}

constexpr int CT_X = 64, CT_Y = 8;      // mosaic tile per block; block = 64 x 4 threads, 2 rows each
constexpr int MAX_TILE_PATCHES = 1024;  // patches that may overlap one tile

// Ordered list (patch order = accumulation order, stitcher.py:223) of the
// patches whose box intersects this block's tile, built cooperatively in
// shared memory with ballots so that no host-side tile lists are needed.
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

// Drop from the tile list every patch whose blurred mask is identically zero
// over the tile: its weights vanish at every level (the supports of the
// truncated Gaussians nest and all taps are positive), so it contributes
// exact zeros.  Tested on the coarse alpha of the widest level.
template <int L>
__device__ int cull_tile_list(const BandPatch *__restrict__ patches, int n_hit, int tx0, int ty0,
                              int16_t *list) {
    if (L < 2) return n_hit;
    const int tid = threadIdx.y * CT_X + threadIdx.x;
    int kept = 0;
    for (int it = 0; it < n_hit; ++it) {
        const int id = list[it];
        const BandPatch &bp = patches[id];
        const int shift = bp.shift[L - 2], lw = bp.lw[L - 2];
        const int px0 = max(tx0, bp.x0) - bp.x0, px1 = min(tx0 + CT_X, bp.x0 + bp.pw) - 1 - bp.x0;
        const int py0 = max(ty0, bp.y0) - bp.y0, py1 = min(ty0 + CT_Y, bp.y0 + bp.ph) - 1 - bp.y0;
        const CoarseTap lo = coarse_tap(shift, bp.pad, px0, py0), hi = coarse_tap(shift, bp.pad, px1, py1);
        const int nx = hi.ix + 2 - lo.ix, ny = hi.iy + 2 - lo.iy;
        const float *alpha = reinterpret_cast<const float *>(bp.low[L - 2]) + 3;
        bool any = false;
        for (int i = tid; i < nx * ny; i += 256) {
            const int cx = lo.ix + i % nx, cy = lo.iy + i / nx;
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

template <int L>
__global__ void __launch_bounds__(256)
multiband_collapse_kernel(const BandPatch *__restrict__ patches, int n_patches,
                          const int32_t *__restrict__ owner, const uint8_t *__restrict__ covered,
                          uint8_t *__restrict__ out, int H, int W) {
    __shared__ int16_t list[MAX_TILE_PATCHES];
    const int tx0 = blockIdx.x * CT_X, ty0 = blockIdx.y * CT_Y;
    int n_hit = build_tile_list(patches, n_patches, tx0, ty0, list);
    n_hit = cull_tile_list<L>(patches, n_hit, tx0, ty0, list);
    const int X = tx0 + threadIdx.x;
#pragma unroll
    for (int sub = 0; sub < CT_Y / 4; ++sub) {
        const int Y = ty0 + threadIdx.y + 4 * sub;
        if (X >= W || Y >= H) continue;
        const size_t mi = (size_t)Y * W + X;
        float num[L][3], den[L];
#pragma unroll
        for (int l = 0; l < L; ++l) num[l][0] = num[l][1] = num[l][2] = den[l] = 0.f;
        const bool cov = covered[mi] != 0;
        if (cov) {
            const int own = __ldg(owner + mi);
            for (int it = 0; it < n_hit; ++it) {              // patch order = list order
                const BandPatch &bp = patches[list[it]];
                const int px = X - bp.x0, py = Y - bp.y0;
                if (px < 0 || py < 0 || px >= bp.pw || py >= bp.ph) continue;
                float4 prev = ld_stream(bp.rgba + (size_t)py * bp.pw + px);
                prev.w = (own == bp.index) ? 1.0f : 0.0f;     // stitcher.py:207-208
                const int pad = bp.pad;
                const CoarseTap t2 = coarse_tap(1, pad, px, py), t4 = coarse_tap(2, pad, px, py);
#pragma unroll
                for (int l = 0; l < L - 1; ++l) {             // stitcher.py:224-232
                    const float4 cur = expand_at(bp.low[l], bp.lw[l], bp.shift[l] == 1 ? t2 : t4);
                    num[l][0] += (prev.x - cur.x) * cur.w;
                    num[l][1] += (prev.y - cur.y) * cur.w;
                    num[l][2] += (prev.z - cur.z) * cur.w;
                    den[l] += cur.w;
                    prev = cur;
                }
                num[L - 1][0] += prev.x * prev.w;
                num[L - 1][1] += prev.y * prev.w;
                num[L - 1][2] += prev.z * prev.w;
                den[L - 1] += prev.w;
            }
        }
        float m0 = 0.f, m1 = 0.f, m2 = 0.f;                   // stitcher.py:236-238
#pragma unroll
        for (int l = 0; l < L; ++l) {
            float w = den[l] == 0.0f ? 1.0f : den[l];
            m0 = __fadd_rn(m0, __fdiv_rn(num[l][0], w));
            m1 = __fadd_rn(m1, __fdiv_rn(num[l][1], w));
            m2 = __fadd_rn(m2, __fdiv_rn(num[l][2], w));
        }
        uint8_t *o = out + mi * 3;                            // stitcher.py:240-241
        o[0] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, fminf(fmaxf(m0, 0.f), 1.f)));
        o[1] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, fminf(fmaxf(m1, 0.f), 1.f)));
        o[2] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, fminf(fmaxf(m2, 0.f), 1.f)));
    }
}

template <int L>
int launch_collapse(const BandPatch *patches, int n_patches, const int32_t *owner,
                    const uint8_t *covered, uint8_t *out, int H, int W, cudaStream_t s) {
    dim3 grid(cdiv(W, CT_X), cdiv(H, CT_Y)), block(CT_X, 4);
    multiband_collapse_kernel<L><<<grid, block, 0, s>>>(patches, n_patches, owner, covered, out, H, W);
    return check_launch("p360_multiband_collapse");
}

}  // namespace p360

using namespace p360;

extern "C" int p360_pyramid_dims(int pw, int ph, int pad, int32_t out_host[4]) {
    const char *where = "p360_pyramid_dims";
    P360_REQUIRE(out_host && pw > 0 && ph > 0 && pad >= 0 && pad % 4 == 0, where);
    int w4 = (pw + 2 * pad + 3) / 4, h4 = (ph + 2 * pad + 3) / 4;
    out_host[0] = 2 * w4; out_host[1] = 2 * h4; out_host[2] = w4; out_host[3] = h4;
    return 0;
}

extern "C" int p360_pyramid_reduce(const float *rgba, int pw, int ph, int x0, int y0, int idx,
                                   const int32_t *owner, int W, int pad, float *d2, float *d4,
                                   void *stream) {
    const char *where = "p360_pyramid_reduce";
    P360_REQUIRE(rgba && d2 && d4 && aligned16(rgba) && aligned16(d2) && aligned16(d4), where);
    P360_REQUIRE(pw > 0 && ph > 0 && pad >= 0 && pad % 4 == 0, where);
    P360_REQUIRE(owner == nullptr || (W > 0 && x0 >= 0 && y0 >= 0 && x0 + pw <= W), where);
    int w4 = (pw + 2 * pad + 3) / 4, h4 = (ph + 2 * pad + 3) / 4;
    dim3 grid(cdiv(4 * w4, 32), cdiv(h4, 8));
    pyramid_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(rgba), pw, ph, x0, y0, idx, owner, W, pad,
        reinterpret_cast<float4 *>(d2), reinterpret_cast<float4 *>(d4), w4, h4);
    return check_launch(where);
}

extern "C" int p360_multiband_collapse(const p360_band_patch *patches, int n_patches, int n_levels,
                                       const int32_t *owner, const uint8_t *covered,
                                       uint8_t *out_u8, int H, int W, void *stream) {
    const char *where = "p360_multiband_collapse";
    P360_REQUIRE(patches && owner && covered && out_u8, where);
    P360_REQUIRE(n_patches >= 0 && n_patches <= MAX_TILE_PATCHES, where);
    P360_REQUIRE(n_levels >= 1 && n_levels <= P360_MAX_LEVELS && H > 0 && W > 0, where);
    auto bp = reinterpret_cast<const BandPatch *>(patches);
    cudaStream_t s = (cudaStream_t)stream;
    switch (n_levels) {
        case 1: return launch_collapse<1>(bp, n_patches, owner, covered, out_u8, H, W, s);
        case 2: return launch_collapse<2>(bp, n_patches, owner, covered, out_u8, H, W, s);
        case 3: return launch_collapse<3>(bp, n_patches, owner, covered, out_u8, H, W, s);
        case 4: return launch_collapse<4>(bp, n_patches, owner, covered, out_u8, H, W, s);
        case 5: return launch_collapse<5>(bp, n_patches, owner, covered, out_u8, H, W, s);
        case 6: return launch_collapse<6>(bp, n_patches, owner, covered, out_u8, H, W, s);
        case 7: return launch_collapse<7>(bp, n_patches, owner, covered, out_u8, H, W, s);
        default: return launch_collapse<8>(bp, n_patches, owner, covered, out_u8, H, W, s);
    }
}
