// Ingest: cv2.resize(img, None, fx=1/S, fy=1/S) on uint8 images (stitcher.py:418-421, the `-s` flag),
// bit-exact with OpenCV's 8-bit path (modules/imgproc/src/resize.cpp):
//   * INTER_LINEAR in 11-bit fixed point: horizontal pass S = a0 * p[sx] + a1 * p[sx + 1] (int32),
//     vertical pass ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2  (FixedPtCast<22>);
//   * an exact 2x shrink is INTER_AREA's 2 x 2 integer mean (a + b + c + d + 2) >> 2; a block that
//     sticks out of an odd-sized image averages the pixels it has (float division, round-half-even).
// The per-column / per-row tables (first source index, the two int16 weights) are built on the host
// (pano360_b200/geometry.py: resize_tables) exactly as resize.cpp builds xofs / ialpha / yofs / ibeta.
// One thread per destination pixel, all channels; HBM-bound: every source byte is read once for S >= 2.
#include "p360_common.cuh"

namespace p360 {

template <int C>
__global__ void __launch_bounds__(256)
resize_linear_kernel(const uint8_t *__restrict__ src, int h, int w, uint8_t *__restrict__ dst, int dh, int dw,
                     const int *__restrict__ xofs, const short *__restrict__ xw,
                     const int *__restrict__ yofs, const short *__restrict__ yw) {
    const int dx = blockIdx.x * 64 + threadIdx.x, dy = blockIdx.y * 4 + threadIdx.y;
    if (dx >= dw || dy >= dh) return;
    const int sx = __ldg(xofs + dx), x1 = min(sx + 1, w - 1);
    const int a0 = __ldg(xw + 2 * dx), a1 = __ldg(xw + 2 * dx + 1);
    const int sy = __ldg(yofs + dy), y0 = min(max(sy, 0), h - 1), y1 = min(max(sy + 1, 0), h - 1);
    const int b0 = __ldg(yw + 2 * dy), b1 = __ldg(yw + 2 * dy + 1);
    const uint8_t *r0 = src + (size_t)y0 * w * C, *r1 = src + (size_t)y1 * w * C;
    uint8_t *o = dst + ((size_t)dy * dw + dx) * C;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int s0 = (int)__ldg(r0 + sx * C + k) * a0 + (int)__ldg(r0 + x1 * C + k) * a1;
        const int s1 = (int)__ldg(r1 + sx * C + k) * a0 + (int)__ldg(r1 + x1 * C + k) * a1;
        const int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
        o[k] = (uint8_t)min(max(v, 0), 255);
    }
}

template <int C>
__global__ void __launch_bounds__(256)
resize_area2_kernel(const uint8_t *__restrict__ src, int h, int w, uint8_t *__restrict__ dst, int dh, int dw) {
    const int dx = blockIdx.x * 64 + threadIdx.x, dy = blockIdx.y * 4 + threadIdx.y;
    if (dx >= dw || dy >= dh) return;
    const int x0 = 2 * dx, y0 = 2 * dy;
    const int nx = min(2, w - x0), ny = min(2, h - y0);
    uint8_t *o = dst + ((size_t)dy * dw + dx) * C;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        int sum = 0;
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) sum += (int)__ldg(src + ((size_t)(y0 + y) * w + x0 + x) * C + k);
        if (nx == 2 && ny == 2) o[k] = (uint8_t)((sum + 2) >> 2);
        else o[k] = (uint8_t)min(max(__float2int_rn(__fdiv_rn((float)sum, (float)(nx * ny))), 0), 255);
    }
}

}  // namespace p360

extern "C" int p360_resize_u8(const uint8_t *src, int h, int w, int c, uint8_t *dst, int dh, int dw,
                              const int32_t *xofs, const int16_t *xw, const int32_t *yofs, const int16_t *yw,
                              int area2, void *stream) {
    using namespace p360;
    const char *where = "p360_resize_u8";
    P360_REQUIRE(src && dst && h > 0 && w > 0 && dh > 0 && dw > 0 && (c == 1 || c == 3 || c == 4), where);
    P360_REQUIRE(area2 || (xofs && xw && yofs && yw), where);
    P360_REQUIRE(!area2 || (2 * (dh - 1) < h && 2 * (dw - 1) < w), where);
    dim3 grid(cdiv(dw, 64), cdiv(dh, 4)), block(64, 4);
    P360_REQUIRE(grid.y <= 65535, where);
    cudaStream_t s = (cudaStream_t)stream;
    if (area2) {
        if (c == 1) resize_area2_kernel<1><<<grid, block, 0, s>>>(src, h, w, dst, dh, dw);
        else if (c == 3) resize_area2_kernel<3><<<grid, block, 0, s>>>(src, h, w, dst, dh, dw);
        else resize_area2_kernel<4><<<grid, block, 0, s>>>(src, h, w, dst, dh, dw);
    } else {
        if (c == 1) resize_linear_kernel<1><<<grid, block, 0, s>>>(src, h, w, dst, dh, dw, xofs, xw, yofs, yw);
        else if (c == 3) resize_linear_kernel<3><<<grid, block, 0, s>>>(src, h, w, dst, dh, dw, xofs, xw, yofs, yw);
        else resize_linear_kernel<4><<<grid, block, 0, s>>>(src, h, w, dst, dh, dw, xofs, xw, yofs, yw);
    }
    return check_launch(where);
}
