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
};

// One source sample as the reference's float RGBA image would hold it:
// rgb = lut[u8] (== u8/255 in float32, gain folded in), a = float(hat_y*hat_x).
__device__ __forceinline__ float4 src_rgba(const SrcView &s, int y, int x) {
    const uint8_t *p = s.pix + ((size_t)y * s.w + x) * s.c;
    float4 v;
    v.x = __ldg(s.lut + __ldg(p));
    v.y = __ldg(s.lut + __ldg(p + 1));
    v.z = __ldg(s.lut + __ldg(p + 2));
    v.w = (float)(__ldg(s.hat_y + y) * __ldg(s.hat_x + x));
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

__device__ __forceinline__ float4 bilinear_q5(const SrcView &s, int y0, int y1, int x0, int x1,
                                              int fx, int fy) {
    float ax = (float)fx * 0.03125f, ay = (float)fy * 0.03125f;
    float w00 = __fmul_rn(1.0f - ay, 1.0f - ax), w01 = __fmul_rn(1.0f - ay, ax);
    float w10 = __fmul_rn(ay, 1.0f - ax), w11 = __fmul_rn(ay, ax);
    float4 a = src_rgba(s, y0, x0), b = src_rgba(s, y0, x1);
    float4 c = src_rgba(s, y1, x0), d = src_rgba(s, y1, x1);
    float4 o;
    o.x = blend4(a.x, b.x, c.x, d.x, w00, w01, w10, w11);
    o.y = blend4(a.y, b.y, c.y, d.y, w00, w01, w10, w11);
    o.z = blend4(a.z, b.z, c.z, d.z, w00, w01, w10, w11);
    o.w = blend4(a.w, b.w, c.w, d.w, w00, w01, w10, w11);
    return o;
}

constexpr int WARP_BX = 64, WARP_BY = 4;

__global__ void __launch_bounds__(WARP_BX *WARP_BY)
warp_patch_kernel(SrcView s, const double *__restrict__ col_tab, const double *__restrict__ row_tab,
                  int pw, int ph, float half_w, float half_h, float max_x, float max_y,
                  float4 *__restrict__ out, uint8_t *__restrict__ invalid) {
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
    int sx = to_fixed5(x), sy = to_fixed5(y);
    int ix = sat16(sx >> 5), iy = sat16(sy >> 5);
    float4 o = bilinear_q5(s, reflect_edge(iy, s.h), reflect_edge(iy + 1, s.h),
                           reflect_edge(ix, s.w), reflect_edge(ix + 1, s.w), sx & 31, sy & 31);
    if (bad) o.w = 0.0f;                                   // stitcher.py:317
    size_t idx = (size_t)r * pw + c;
    st_stream(out + idx, o);
    invalid[idx] = bad ? 1 : 0;
}

}  // namespace p360

extern "C" int p360_warp_patch(const uint8_t *src, int src_h, int src_w, int src_c,
                               const float *lut, const double *hat_y, const double *hat_x,
                               const double *col_tab, const double *row_tab,
                               int pw, int ph, float *out_rgba, uint8_t *out_invalid,
                               void *stream) {
    using namespace p360;
    const char *where = "p360_warp_patch";
    P360_REQUIRE(src && lut && hat_y && hat_x && col_tab && row_tab && out_rgba && out_invalid, where);
    P360_REQUIRE(src_h > 0 && src_w > 0 && (src_c == 3 || src_c == 4), where);
    P360_REQUIRE(pw >= 0 && ph >= 0, where);
    P360_REQUIRE(aligned16(out_rgba), where);
    if (pw == 0 || ph == 0) return 0;
    SrcView s{src, lut, hat_y, hat_x, src_h, src_w, src_c};
    dim3 block(WARP_BX, WARP_BY), grid(cdiv(pw, WARP_BX), cdiv(ph, WARP_BY));
    warp_patch_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
        s, col_tab, row_tab, pw, ph, (float)(src_w / 2.0), (float)(src_h / 2.0),
        (float)(src_w - 1), (float)(src_h - 1), reinterpret_cast<float4 *>(out_rgba), out_invalid);
    return check_launch(where);
}
