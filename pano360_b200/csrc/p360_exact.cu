// The reference's multiband loop nest, stage by stage at full resolution (stitcher.py:186-241):
// owner mask into alpha (:207-208), cv2.GaussianBlur of every patch at every level (K3,
// p360_gauss_blur), band = prev - blur weighted by the blurred mask and accumulated per level
// (:224-232), per-level normalisation, sum, clip and uint8 (:236-241).  ~2000 FMA per patch pixel:
// FP32-issue-bound and two orders of magnitude slower than the coarse-grid pipeline of
// p360_pyramid.cu — it exists as the device-side ground truth of that pipeline (whole mosaics at
// sizes the CPU oracle cannot hold) and as the path taken for images so small that the coarse
// grids cannot resolve their owner masks.
#include "p360_common.cuh"

namespace p360 {

constexpr int EX = 64, EY = 4;

// alpha := (owner == idx), in place (stitcher.py:207-208)
__global__ void __launch_bounds__(EX *EY)
owner_to_alpha_kernel(float4 *__restrict__ rgba, int pw, int ph, int x0, int y0, int idx,
                      const unsigned long long *__restrict__ keys, int W) {
    const int c = blockIdx.x * EX + threadIdx.x, r = blockIdx.y * EY + threadIdx.y;
    if (c >= pw || r >= ph) return;
    const bool mine = key_is_owner(keys[(size_t)(r + y0) * W + (c + x0)], idx);
    reinterpret_cast<float *>(rgba + (size_t)r * pw + c)[3] = mine ? 1.0f : 0.0f;
}

// acc[pixel] += (band * weight, weight): band = prev.rgb - cur.rgb, weight = cur.a; for the last
// level (cur == nullptr) band = prev.rgb, weight = prev.a.  Products and sums separately rounded,
// like the NumPy expressions at stitcher.py:227-232.
__global__ void __launch_bounds__(EX *EY)
band_accumulate_kernel(const float4 *__restrict__ prev, const float4 *__restrict__ cur, int pw, int ph,
                       int x0, int y0, float4 *__restrict__ acc, int W) {
    const int c = blockIdx.x * EX + threadIdx.x, r = blockIdx.y * EY + threadIdx.y;
    if (c >= pw || r >= ph) return;
    const size_t pi = (size_t)r * pw + c, mi = (size_t)(r + y0) * W + (c + x0);
    float4 p = prev[pi];
    float wgt = p.w;
    if (cur != nullptr) {
        const float4 q = cur[pi];
        p.x = __fadd_rn(p.x, -q.x); p.y = __fadd_rn(p.y, -q.y); p.z = __fadd_rn(p.z, -q.z);
        wgt = q.w;
    }
    float4 a = acc[mi];
    a.x = __fadd_rn(a.x, __fmul_rn(p.x, wgt));
    a.y = __fadd_rn(a.y, __fmul_rn(p.y, wgt));
    a.z = __fadd_rn(a.z, __fmul_rn(p.z, wgt));
    a.w = __fadd_rn(a.w, wgt);
    acc[mi] = a;
}

// mosaic = sum_l where(covered, layer_l, 0) / where(wsum_l == 0, 1, wsum_l); clip; uint8 (:236-241)
__global__ void __launch_bounds__(256)
exact_collapse_kernel(const float4 *__restrict__ acc, long long level_stride, int n_levels,
                      const uint8_t *__restrict__ covered, uint8_t *__restrict__ out, long long n_px) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_px) return;
    float m0 = 0.f, m1 = 0.f, m2 = 0.f;
    const bool valid = covered[i] != 0;
    for (int l = 0; l < n_levels; ++l) {
        const float4 a = acc[(size_t)l * level_stride + i];
        const float w = a.w == 0.0f ? 1.0f : a.w;
        m0 = __fadd_rn(m0, __fdiv_rn(valid ? a.x : 0.0f, w));
        m1 = __fadd_rn(m1, __fdiv_rn(valid ? a.y : 0.0f, w));
        m2 = __fadd_rn(m2, __fdiv_rn(valid ? a.z : 0.0f, w));
    }
    uint8_t *o = out + i * 3;
    o[0] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, fminf(fmaxf(m0, 0.f), 1.f)));
    o[1] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, fminf(fmaxf(m1, 0.f), 1.f)));
    o[2] = (uint8_t)__float2int_rz(__fmul_rn(255.0f, fminf(fmaxf(m2, 0.f), 1.f)));
}

}  // namespace p360

using namespace p360;

extern "C" int p360_owner_to_alpha(float *rgba, int pw, int ph, int x0, int y0, int idx,
                                   const uint64_t *owner_keys, int W, void *stream) {
    const char *where = "p360_owner_to_alpha";
    P360_REQUIRE(rgba && owner_keys && aligned16(rgba) && pw >= 0 && ph >= 0 && x0 >= 0 && y0 >= 0 && x0 + pw <= W, where);
    if (pw == 0 || ph == 0) return 0;
    owner_to_alpha_kernel<<<dim3(cdiv(pw, EX), cdiv(ph, EY)), dim3(EX, EY), 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4 *>(rgba), pw, ph, x0, y0, idx, reinterpret_cast<const unsigned long long *>(owner_keys), W);
    return check_launch(where);
}

extern "C" int p360_band_accumulate(const float *prev_rgba, const float *cur_rgba, int pw, int ph, int x0, int y0,
                                    float *acc_rgbw, int W, void *stream) {
    const char *where = "p360_band_accumulate";
    P360_REQUIRE(prev_rgba && acc_rgbw && aligned16(prev_rgba) && aligned16(acc_rgbw) && aligned16(cur_rgba), where);
    P360_REQUIRE(pw >= 0 && ph >= 0 && x0 >= 0 && y0 >= 0 && x0 + pw <= W, where);
    if (pw == 0 || ph == 0) return 0;
    band_accumulate_kernel<<<dim3(cdiv(pw, EX), cdiv(ph, EY)), dim3(EX, EY), 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(prev_rgba), reinterpret_cast<const float4 *>(cur_rgba), pw, ph, x0, y0,
        reinterpret_cast<float4 *>(acc_rgbw), W);
    return check_launch(where);
}

extern "C" int p360_exact_collapse(const float *acc_rgbw, int n_levels, const uint8_t *covered, uint8_t *out_u8,
                                   int H, int W, void *stream) {
    const char *where = "p360_exact_collapse";
    P360_REQUIRE(acc_rgbw && covered && out_u8 && aligned16(acc_rgbw), where);
    P360_REQUIRE(n_levels >= 1 && n_levels <= P360_MAX_LEVELS && H > 0 && W > 0, where);
    const long long n = (long long)H * W;
    exact_collapse_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(acc_rgbw), n, n_levels, covered, out_u8, n);
    return check_launch(where);
}
