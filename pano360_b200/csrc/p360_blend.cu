// K2 and the crop mask: owner-map competition for externally supplied patches
// (blender API) and the union of valid pixels.  Pointwise in mosaic
// coordinates, HBM-bound: one thread per pixel, patch rows on mosaic rows.
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
                    int pw, int ph, int x0, int y0, int idx, unsigned long long *__restrict__ keys,
                    uint8_t *__restrict__ covered, int W) {
    P360_PATCH_XY();
    owner_compete(keys, mi, ld_stream(rgba + pi).w, idx);
    if (!invalid[pi]) covered[mi] = 1;
}

__global__ void __launch_bounds__(256)
owner_decode_kernel(const unsigned long long *__restrict__ keys, int32_t *__restrict__ owner, long long n) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    owner[i] = k == 0ull ? -1 : (int32_t)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull));
}

__global__ void __launch_bounds__(BX *BY)
cover_update_kernel(const uint8_t *__restrict__ invalid, int pw, int ph, int x0, int y0,
                    uint8_t *__restrict__ covered, int W) {
    P360_PATCH_XY();
    if (!invalid[pi]) covered[mi] = 1;
}

inline dim3 patch_grid(int pw, int ph) { return dim3(cdiv(pw, BX), cdiv(ph, BY)); }

}  // namespace p360

using namespace p360;

#define P360_PATCH_ARGS_OK(where)                                                   \
    P360_REQUIRE(pw >= 0 && ph >= 0 && x0 >= 0 && y0 >= 0 && W > 0 && x0 + pw <= W, where); \
    if (pw == 0 || ph == 0) return 0;

extern "C" int p360_owner_update(const float *rgba, const uint8_t *invalid, int pw, int ph,
                                 int x0, int y0, int idx, uint64_t *owner_keys,
                                 uint8_t *covered, int W, void *stream) {
    const char *where = "p360_owner_update";
    P360_REQUIRE(rgba && invalid && owner_keys && covered && aligned16(rgba), where);
    P360_PATCH_ARGS_OK(where);
    owner_update_kernel<<<patch_grid(pw, ph), dim3(BX, BY), 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(rgba), invalid, pw, ph, x0, y0, idx,
        reinterpret_cast<unsigned long long *>(owner_keys), covered, W);
    return check_launch(where);
}

extern "C" int p360_owner_decode(const uint64_t *owner_keys, int32_t *owner, int64_t n_pixels, void *stream) {
    const char *where = "p360_owner_decode";
    P360_REQUIRE(owner_keys && owner && n_pixels >= 0, where);
    if (n_pixels == 0) return 0;
    owner_decode_kernel<<<cdiv(n_pixels, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const unsigned long long *>(owner_keys), owner, (long long)n_pixels);
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
