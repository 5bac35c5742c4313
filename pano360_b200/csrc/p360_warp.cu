// K1: fused inverse projection + 1/32-px bilinear remap + validity mask.
// Replaces stitcher.py:257-263 and :300-317 (NumPy coordinate maths, BLAS
// 3x3 projection, cv2.remap, alpha masking) with one pass that reads the u8
// source through L1 and writes each RGBA float4 exactly once, coalesced.
#include "p360_common.cuh"

namespace p360 {

struct SrcView {
    const uint8_t *pix;
    const float *lut;        // 256 entries
    const double *hat_y;     // h entries
    const double *hat_x;     // w entries
    int h, w, c;
    float inv_2h, inv_2w;    // 1 / (2h), 1 / (2w): reflection period reciprocals
};

// Optional fused K2 state (stitcher.py:196-204, :233-234): running arg-max of
// alpha and the union of valid pixels, updated in patch order.
struct OwnerState {
    float *best;
    int32_t *owner;
    uint8_t *covered;
    int x0, y0, W, idx;
};

// BORDER_REFLECT for p in the int16 range without an integer division:
// q = p mod 2n through a float reciprocal (|p| <= 2^15, so the quotient is off
// by at most one) and two fix-ups.
__device__ __forceinline__ int reflect_fast(int p, int n, float inv_2n) {
    const int m = 2 * n;
    int q = p - m * __float2int_rd((float)p * inv_2n);
    q += (q < 0) ? m : 0;
    q -= (q >= m) ? m : 0;
    return q < n ? q : m - 1 - q;
}

__device__ __forceinline__ float4 sample(const SrcView &s, const float *lut, int y, int x, double hy,
                                         double hx) {
    const uint8_t *p = s.pix + ((size_t)y * s.w + x) * s.c;
    float4 v;
    if (s.c == 4) {
        const uint32_t u = __ldg(reinterpret_cast<const uint32_t *>(p));
        v.x = lut[u & 0xff]; v.y = lut[(u >> 8) & 0xff]; v.z = lut[(u >> 16) & 0xff];
    } else {
        v.x = lut[__ldg(p)]; v.y = lut[__ldg(p + 1)]; v.z = lut[__ldg(p + 2)];
    }
    v.w = (float)(hy * hx);          // float32(hat_y * hat_x), stitcher.py:261
    return v;
}

// ((s00*w00 + s01*w01) + s10*w10) + s11*w11, separately rounded products and
// sums: the exact evaluation order of OpenCV's remapBilinear float path.
__device__ __forceinline__ float blend4(float a, float b, float c, float d,
                                        float w00, float w01, float w10, float w11) {
    float acc = __fmul_rn(a, w00);
    acc = __fadd_rn(acc, __fmul_rn(b, w01));
    acc = __fadd_rn(acc, __fmul_rn(c, w10));
    acc = __fadd_rn(acc, __fmul_rn(d, w11));
    return acc;
}

constexpr int WARP_BX = 64, WARP_BY = 4;

__global__ void __launch_bounds__(WARP_BX *WARP_BY)
warp_patch_kernel(SrcView s, const double *__restrict__ col_tab, const double *__restrict__ row_tab,
                  int pw, int ph, float half_w, float half_h, float max_x, float max_y,
                  float4 *__restrict__ out, uint8_t *__restrict__ invalid, OwnerState own) {
    __shared__ float lut[256];
    lut[threadIdx.y * WARP_BX + threadIdx.x] = s.lut[threadIdx.y * WARP_BX + threadIdx.x];
    __syncthreads();
    int c = blockIdx.x * WARP_BX + threadIdx.x;
    int r = blockIdx.y * WARP_BY + threadIdx.y;
    if (c >= pw || r >= ph) return;
    // p = K R (rx, ry, rz): column part + row part, float64, then cast
    // (stitcher.py:303-306)
    const double *ct = col_tab + (size_t)c * 3, *rt = row_tab + (size_t)r * 3;
    float px = (float)(__ldg(ct) + __ldg(rt));
    float py = (float)(__ldg(ct + 1) + __ldg(rt + 1));
    float pz = (float)(__ldg(ct + 2) + __ldg(rt + 2));
    bool bad = pz < 0.0f;                                  // stitcher.py:308
    float x = __fadd_rn(__fdiv_rn(px, pz), half_w);        // stitcher.py:310
    float y = __fadd_rn(__fdiv_rn(py, pz), half_h);
    bad |= (x < 0.0f) | (x > max_x) | (y < 0.0f) | (y > max_y);   // :311-312
    const int sx = to_fixed5(x), sy = to_fixed5(y);
    const int ix = sat16(sx >> 5), iy = sat16(sy >> 5);
    int x0 = ix, x1 = ix + 1, y0 = iy, y1 = iy + 1;
    if ((unsigned)ix >= (unsigned)(s.w - 1) || (unsigned)iy >= (unsigned)(s.h - 1)) {
        // a tap falls outside the image: BORDER_REFLECT (cv2.remap at stitcher.py:315-316)
        x0 = reflect_fast(ix, s.w, s.inv_2w); x1 = reflect_fast(ix + 1, s.w, s.inv_2w);
        y0 = reflect_fast(iy, s.h, s.inv_2h); y1 = reflect_fast(iy + 1, s.h, s.inv_2h);
    }
    const float ax = (float)(sx & 31) * 0.03125f, ay = (float)(sy & 31) * 0.03125f;
    const float w00 = __fmul_rn(1.0f - ay, 1.0f - ax), w01 = __fmul_rn(1.0f - ay, ax);
    const float w10 = __fmul_rn(ay, 1.0f - ax), w11 = __fmul_rn(ay, ax);
    const double hy0 = __ldg(s.hat_y + y0), hy1 = __ldg(s.hat_y + y1);
    const double hx0 = __ldg(s.hat_x + x0), hx1 = __ldg(s.hat_x + x1);
    const float4 a = sample(s, lut, y0, x0, hy0, hx0), b = sample(s, lut, y0, x1, hy0, hx1);
    const float4 cc = sample(s, lut, y1, x0, hy1, hx0), d = sample(s, lut, y1, x1, hy1, hx1);
    float4 o;
    o.x = blend4(a.x, b.x, cc.x, d.x, w00, w01, w10, w11);
    o.y = blend4(a.y, b.y, cc.y, d.y, w00, w01, w10, w11);
    o.z = blend4(a.z, b.z, cc.z, d.z, w00, w01, w10, w11);
    o.w = blend4(a.w, b.w, cc.w, d.w, w00, w01, w10, w11);
    if (bad) o.w = 0.0f;                                   // stitcher.py:317
    size_t idx = (size_t)r * pw + c;
    st_stream(out + idx, o);
    invalid[idx] = bad ? 1 : 0;
    if (own.best != nullptr) {
        const size_t mi = (size_t)(r + own.y0) * own.W + (c + own.x0);
        if (o.w > own.best[mi]) {          // strict: the first maximum wins (np.argmax)
            own.best[mi] = o.w;
            own.owner[mi] = own.idx;
        }
        if (!bad) own.covered[mi] = 1;
    }
}

// u8 x 3 -> u8 x 4 (one aligned 32-bit word per source pixel for the gathers)
__global__ void __launch_bounds__(256)
pack_rgbx_kernel(const uint8_t *__restrict__ src, uint32_t *__restrict__ dst, long long n) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint8_t *p = src + i * 3;
    dst[i] = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
}

}  // namespace p360

extern "C" int p360_pack_rgbx(const uint8_t *src_rgb, uint8_t *dst_rgbx, int64_t n_pixels, void *stream) {
    using namespace p360;
    const char *where = "p360_pack_rgbx";
    P360_REQUIRE(src_rgb && dst_rgbx && n_pixels >= 0, where);
    P360_REQUIRE((reinterpret_cast<uintptr_t>(dst_rgbx) & 3) == 0, where);
    if (n_pixels == 0) return 0;
    pack_rgbx_kernel<<<cdiv(n_pixels, 256), 256, 0, (cudaStream_t)stream>>>(
        src_rgb, reinterpret_cast<uint32_t *>(dst_rgbx), (long long)n_pixels);
    return check_launch(where);
}

extern "C" int p360_warp_patch(const uint8_t *src, int src_h, int src_w, int src_c,
                               const float *lut, const double *hat_y, const double *hat_x,
                               const double *col_tab, const double *row_tab,
                               int pw, int ph, float *out_rgba, uint8_t *out_invalid,
                               int x0, int y0, int idx, float *best, int32_t *owner,
                               uint8_t *covered, int W, void *stream) {
    using namespace p360;
    const char *where = "p360_warp_patch";
    P360_REQUIRE(src && lut && hat_y && hat_x && col_tab && row_tab && out_rgba && out_invalid, where);
    P360_REQUIRE(src_h > 0 && src_w > 0 && src_h <= 32767 && src_w <= 32767, where);
    P360_REQUIRE(src_c == 3 || (src_c == 4 && (reinterpret_cast<uintptr_t>(src) & 3) == 0), where);
    P360_REQUIRE(pw >= 0 && ph >= 0, where);
    P360_REQUIRE(aligned16(out_rgba), where);
    P360_REQUIRE(best == nullptr || (owner && covered && W > 0 && x0 >= 0 && y0 >= 0 && x0 + pw <= W), where);
    if (pw == 0 || ph == 0) return 0;
    SrcView s{src, lut, hat_y, hat_x, src_h, src_w, src_c, 1.0f / (2.0f * src_h), 1.0f / (2.0f * src_w)};
    OwnerState own{best, owner, covered, x0, y0, W, idx};
    dim3 block(WARP_BX, WARP_BY), grid(cdiv(pw, WARP_BX), cdiv(ph, WARP_BY));
    warp_patch_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
        s, col_tab, row_tab, pw, ph, (float)(src_w / 2.0), (float)(src_h / 2.0),
        (float)(src_w - 1), (float)(src_h - 1), reinterpret_cast<float4 *>(out_rgba), out_invalid, own);
    return check_launch(where);
}
