// K2, K4, K5, K6, K7: owner map, band accumulate, collapse, linear blend,
// paste.  All are pointwise in mosaic coordinates and HBM-bound: one thread
// per pixel, float4 accesses, patch rows mapped onto mosaic rows.
#include "p360_common.cuh"

namespace p360 {

constexpr int BX = 64, BY = 4;   // 256 threads, 64 px x 4 rows; a warp spans 512 B of a row

#define P360_PATCH_XY()                                        \
    int c = blockIdx.x * BX + threadIdx.x;                     \
    int r = blockIdx.y * BY + threadIdx.y;                     \
    if (c >= pw || r >= ph) return;                            \
    size_t pi = (size_t)r * pw + c;                            \
    size_t mi = (size_t)(r + y0) * W + (c + x0);

// ---- K2 owner (stitcher.py:196-208, :233-234) -----------------------------
__global__ void __launch_bounds__(BX *BY)
owner_update_kernel(const float4 *__restrict__ rgba, const uint8_t *__restrict__ invalid,
                    int pw, int ph, int x0, int y0, int idx, float *__restrict__ best,
                    int32_t *__restrict__ owner, uint8_t *__restrict__ covered, int W) {
    P360_PATCH_XY();
    float a = ld_stream(rgba + pi).w;
    if (a > best[mi]) {          // strict: the first maximum wins (np.argmax)
        best[mi] = a;
        owner[mi] = idx;
    }
    if (!invalid[pi]) covered[mi] = 1;
}

__global__ void __launch_bounds__(BX *BY)
owner_to_alpha_kernel(float4 *__restrict__ rgba, int pw, int ph, int x0, int y0, int idx,
                      const int32_t *__restrict__ owner, int W) {
    P360_PATCH_XY();
    reinterpret_cast<float *>(rgba + pi)[3] = (__ldg(owner + mi) == idx) ? 1.0f : 0.0f;
}

__global__ void __launch_bounds__(BX *BY)
cover_update_kernel(const uint8_t *__restrict__ invalid, int pw, int ph, int x0, int y0,
                    uint8_t *__restrict__ covered, int W) {
    P360_PATCH_XY();
    if (!invalid[pi]) covered[mi] = 1;
}

// ---- K4 band accumulate (stitcher.py:224-232) ------------------------------
template <bool LAST>
__global__ void __launch_bounds__(BX *BY)
band_accumulate_kernel(const float4 *__restrict__ prev, const float4 *__restrict__ cur,
                       int pw, int ph, int x0, int y0, float4 *__restrict__ acc, int W) {
    P360_PATCH_XY();
    float4 p = ld_stream(prev + pi);
    float bx, by, bz, wt;
    if (LAST) {
        bx = p.x; by = p.y; bz = p.z; wt = p.w;
    } else {
        float4 q = ld_stream(cur + pi);
        bx = __fsub_rn(p.x, q.x); by = __fsub_rn(p.y, q.y); bz = __fsub_rn(p.z, q.z);
        wt = q.w;
    }
    float4 a = acc[mi];
    a.x = __fadd_rn(a.x, __fmul_rn(bx, wt));
    a.y = __fadd_rn(a.y, __fmul_rn(by, wt));
    a.z = __fadd_rn(a.z, __fmul_rn(bz, wt));
    a.w = __fadd_rn(a.w, wt);
    acc[mi] = a;
}

__device__ __forceinline__ uint8_t to_u8_trunc(float v) {   // (255*v).astype(np.uint8), v in [0,1]
    return (uint8_t)__float2int_rz(__fmul_rn(255.0f, v));
}

// ---- K5 collapse + normalise + clamp (stitcher.py:236-241) -----------------
// One warp handles 32 consecutive pixels: 96 output bytes are exchanged with
// shuffles so that each of the first 24 lanes stores one aligned 32-bit word.
__global__ void __launch_bounds__(256)
collapse_finalize_kernel(const float4 *__restrict__ acc, int n_levels,
                         const uint8_t *__restrict__ covered, uint8_t *__restrict__ out,
                         long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float mx = 0.f, my = 0.f, mz = 0.f;
    bool live = i < n;
    if (live && covered[i]) {
        for (int l = 0; l < n_levels; ++l) {
            float4 a = ld_stream(acc + (size_t)l * n + i);
            float w = a.w == 0.0f ? 1.0f : a.w;
            mx = __fadd_rn(mx, __fdiv_rn(a.x, w));
            my = __fadd_rn(my, __fdiv_rn(a.y, w));
            mz = __fadd_rn(mz, __fdiv_rn(a.z, w));
        }
    }
    unsigned b0 = to_u8_trunc(fminf(fmaxf(mx, 0.f), 1.f));
    unsigned b1 = to_u8_trunc(fminf(fmaxf(my, 0.f), 1.f));
    unsigned b2 = to_u8_trunc(fminf(fmaxf(mz, 0.f), 1.f));
    unsigned packed = b0 | (b1 << 8) | (b2 << 16);
    long long warp_base = i - (threadIdx.x & 31);
    if (warp_base + 32 <= n) {
        // lane j (< 24) assembles output bytes 4j .. 4j+3 of the warp's 96
        int lane = threadIdx.x & 31;
        unsigned word = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int byte = 4 * lane + k;           // 0..127 (only < 96 meaningful)
            unsigned src = __shfl_sync(0xffffffffu, packed, (byte / 3) & 31);
            word |= ((src >> (8 * (byte % 3))) & 0xffu) << (8 * k);
        }
        if (lane < 24) reinterpret_cast<unsigned *>(out + warp_base * 3)[lane] = word;
    } else if (live) {
        out[i * 3] = b0; out[i * 3 + 1] = b1; out[i * 3 + 2] = b2;
    }
}

// ---- K6 linear blend (stitcher.py:171-183) ---------------------------------
__global__ void __launch_bounds__(BX *BY)
linear_accumulate_kernel(const float4 *__restrict__ rgba, const uint8_t *__restrict__ invalid,
                         int pw, int ph, int x0, int y0, float4 *__restrict__ acc, int W) {
    P360_PATCH_XY();
    float4 p = ld_stream(rgba + pi);
    bool bad = invalid[pi];
    float4 a = acc[mi];
    a.x = __fadd_rn(a.x, __fmul_rn(bad ? 0.f : p.x, p.w));
    a.y = __fadd_rn(a.y, __fmul_rn(bad ? 0.f : p.y, p.w));
    a.z = __fadd_rn(a.z, __fmul_rn(bad ? 0.f : p.z, p.w));
    a.w = __fadd_rn(a.w, p.w);
    acc[mi] = a;
}

__global__ void __launch_bounds__(256)
linear_finalize_kernel(const float4 *__restrict__ acc, uint8_t *__restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = ld_stream(acc + i);
    float w = a.w == 0.0f ? 1.0f : a.w;
    // (255 * (acc / wsum)).astype(uint8): no clip in the reference; values are
    // convex combinations of [0,1] samples so the cast is in range.
    out[i * 3] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, __fdiv_rn(a.x, w)));
    out[i * 3 + 1] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, __fdiv_rn(a.y, w)));
    out[i * 3 + 2] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, __fdiv_rn(a.z, w)));
}

// ---- K7 paste (stitcher.py:160-168) ----------------------------------------
__global__ void __launch_bounds__(BX *BY)
paste_kernel(const float4 *__restrict__ rgba, const uint8_t *__restrict__ invalid,
             int pw, int ph, int x0, int y0, uint8_t *__restrict__ mosaic, int W) {
    P360_PATCH_XY();
    if (invalid[pi]) return;
    float4 p = ld_stream(rgba + pi);
    uint8_t *o = mosaic + mi * 3;
    o[0] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, p.x));
    o[1] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, p.y));
    o[2] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, p.z));
}

inline dim3 patch_grid(int pw, int ph) { return dim3(cdiv(pw, BX), cdiv(ph, BY)); }

}  // namespace p360

using namespace p360;

#define P360_PATCH_ARGS_OK(where)                                                   \
    P360_REQUIRE(pw >= 0 && ph >= 0 && x0 >= 0 && y0 >= 0 && W > 0 && x0 + pw <= W, where); \
    if (pw == 0 || ph == 0) return 0;

extern "C" int p360_owner_update(const float *rgba, const uint8_t *invalid, int pw, int ph,
                                 int x0, int y0, int idx, float *best, int32_t *owner,
                                 uint8_t *covered, int W, void *stream) {
    const char *where = "p360_owner_update";
    P360_REQUIRE(rgba && invalid && best && owner && covered && aligned16(rgba), where);
    P360_PATCH_ARGS_OK(where);
    owner_update_kernel<<<patch_grid(pw, ph), dim3(BX, BY), 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(rgba), invalid, pw, ph, x0, y0, idx, best, owner, covered, W);
    return check_launch(where);
}

extern "C" int p360_owner_to_alpha(float *rgba, int pw, int ph, int x0, int y0, int idx,
                                   const int32_t *owner, int W, void *stream) {
    const char *where = "p360_owner_to_alpha";
    P360_REQUIRE(rgba && owner && aligned16(rgba), where);
    P360_PATCH_ARGS_OK(where);
    owner_to_alpha_kernel<<<patch_grid(pw, ph), dim3(BX, BY), 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4 *>(rgba), pw, ph, x0, y0, idx, owner, W);
    return check_launch(where);
}

extern "C" int p360_cover_update(const uint8_t *invalid, int pw, int ph, int x0, int y0,
                                 uint8_t *covered, int W, void *stream) {
    const char *where = "p360_cover_update";
    P360_REQUIRE(invalid && covered, where);
    P360_PATCH_ARGS_OK(where);
    cover_update_kernel<<<patch_grid(pw, ph), dim3(BX, BY), 0, (cudaStream_t)stream>>>(
        invalid, pw, ph, x0, y0, covered, W);
    return check_launch(where);
}

extern "C" int p360_band_accumulate(const float *prev_rgba, const float *cur_rgba, int pw, int ph,
                                    int x0, int y0, float *acc, int W, void *stream) {
    const char *where = "p360_band_accumulate";
    P360_REQUIRE(prev_rgba && acc && aligned16(prev_rgba) && aligned16(cur_rgba) && aligned16(acc), where);
    P360_PATCH_ARGS_OK(where);
    auto p = reinterpret_cast<const float4 *>(prev_rgba);
    auto c = reinterpret_cast<const float4 *>(cur_rgba);
    auto a = reinterpret_cast<float4 *>(acc);
    if (cur_rgba)
        band_accumulate_kernel<false><<<patch_grid(pw, ph), dim3(BX, BY), 0, (cudaStream_t)stream>>>(
            p, c, pw, ph, x0, y0, a, W);
    else
        band_accumulate_kernel<true><<<patch_grid(pw, ph), dim3(BX, BY), 0, (cudaStream_t)stream>>>(
            p, c, pw, ph, x0, y0, a, W);
    return check_launch(where);
}

extern "C" int p360_collapse_finalize(const float *acc, int n_levels, const uint8_t *covered,
                                      uint8_t *out_u8, int64_t n_pixels, void *stream) {
    const char *where = "p360_collapse_finalize";
    P360_REQUIRE(acc && covered && out_u8 && aligned16(acc) && n_levels >= 1 && n_pixels >= 0, where);
    P360_REQUIRE((reinterpret_cast<uintptr_t>(out_u8) & 3) == 0, where);
    if (n_pixels == 0) return 0;
    collapse_finalize_kernel<<<cdiv(n_pixels, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(acc), n_levels, covered, out_u8, (long long)n_pixels);
    return check_launch(where);
}

extern "C" int p360_linear_accumulate(const float *rgba, const uint8_t *invalid, int pw, int ph,
                                      int x0, int y0, float *acc, int W, void *stream) {
    const char *where = "p360_linear_accumulate";
    P360_REQUIRE(rgba && invalid && acc && aligned16(rgba) && aligned16(acc), where);
    P360_PATCH_ARGS_OK(where);
    linear_accumulate_kernel<<<patch_grid(pw, ph), dim3(BX, BY), 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(rgba), invalid, pw, ph, x0, y0,
        reinterpret_cast<float4 *>(acc), W);
    return check_launch(where);
}

extern "C" int p360_linear_finalize(const float *acc, uint8_t *out_u8, int64_t n_pixels, void *stream) {
    const char *where = "p360_linear_finalize";
    P360_REQUIRE(acc && out_u8 && aligned16(acc) && n_pixels >= 0, where);
    if (n_pixels == 0) return 0;
    linear_finalize_kernel<<<cdiv(n_pixels, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(acc), out_u8, (long long)n_pixels);
    return check_launch(where);
}

extern "C" int p360_paste(const float *rgba, const uint8_t *invalid, int pw, int ph, int x0, int y0,
                          uint8_t *mosaic_u8, int W, void *stream) {
    const char *where = "p360_paste";
    P360_REQUIRE(rgba && invalid && mosaic_u8 && aligned16(rgba), where);
    P360_PATCH_ARGS_OK(where);
    paste_kernel<<<patch_grid(pw, ph), dim3(BX, BY), 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(rgba), invalid, pw, ph, x0, y0, mosaic_u8, W);
    return check_launch(where);
}
