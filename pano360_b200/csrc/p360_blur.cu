// K3: separable float32 Gaussian on RGBA patches, BORDER_REFLECT_101 at the
// patch edges — cv2.GaussianBlur(warped, (0, 0), sigma) at stitcher.py:226.
// Horizontal pass into tmp, vertical pass into out.  Taps arrive as a kernel
// parameter (constant bank, uniform across the warp).
#include "p360_common.cuh"

namespace p360 {

struct Taps {
    float k[P360_MAX_KSIZE];
    int ksize;
};

__device__ __forceinline__ void fma4(float4 &acc, float w, const float4 &v) {
    acc.x = fmaf(w, v.x, acc.x);
    acc.y = fmaf(w, v.y, acc.y);
    acc.z = fmaf(w, v.z, acc.z);
    acc.w = fmaf(w, v.w, acc.w);
}

constexpr int HB = 256;      // output pixels per block (one row segment)

// Horizontal: the row segment plus a halo of r pixels each side is staged in
// shared memory (reflect applied while staging), each thread then slides over
// its ksize neighbours.
__global__ void __launch_bounds__(HB)
blur_h_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, int pw, int ph, Taps t) {
    extern __shared__ float4 tile[];
    const int r = t.ksize >> 1;
    const int nxb = (pw + HB - 1) / HB;
    const int row = blockIdx.x / nxb;
    const int xb = (blockIdx.x % nxb) * HB;
    const float4 *src = in + (size_t)row * pw;
    for (int i = threadIdx.x; i < HB + 2 * r; i += HB) {
        tile[i] = src[reflect_101(xb - r + i, pw)];
    }
    __syncthreads();
    int x = xb + threadIdx.x;
    if (x >= pw) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < t.ksize; ++k) fma4(acc, t.k[k], tile[threadIdx.x + k]);
    out[(size_t)row * pw + x] = acc;
}

constexpr int VBX = 32, VBY = 8;

// Vertical: lanes along x (coalesced 512-byte row segments), every thread
// walks the ksize rows above/below its pixel; reuse between neighbouring rows
// is served by L1/L2.
__global__ void __launch_bounds__(VBX *VBY)
blur_v_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, int pw, int ph, Taps t) {
    int x = blockIdx.x * VBX + threadIdx.x;
    int y = blockIdx.y * VBY + threadIdx.y;
    if (x >= pw || y >= ph) return;
    const int r = t.ksize >> 1;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < t.ksize; ++k) {
        int yy = reflect_101(y - r + k, ph);
        fma4(acc, t.k[k], __ldg(in + (size_t)yy * pw + x));
    }
    out[(size_t)y * pw + x] = acc;
}

}  // namespace p360

extern "C" int p360_gauss_blur(const float *in_rgba, float *out_rgba, float *tmp_rgba,
                               int pw, int ph, const float *taps_host, int ksize, void *stream) {
    using namespace p360;
    const char *where = "p360_gauss_blur";
    P360_REQUIRE(in_rgba && out_rgba && tmp_rgba && taps_host, where);
    P360_REQUIRE(aligned16(in_rgba) && aligned16(out_rgba) && aligned16(tmp_rgba), where);
    P360_REQUIRE(ksize >= 1 && (ksize & 1) && ksize <= P360_MAX_KSIZE, where);
    P360_REQUIRE(pw >= 0 && ph >= 0, where);
    P360_REQUIRE(in_rgba != out_rgba && in_rgba != tmp_rgba && out_rgba != tmp_rgba, where);
    if (pw == 0 || ph == 0) return 0;
    Taps t;
    memset(&t, 0, sizeof(t));
    memcpy(t.k, taps_host, sizeof(float) * ksize);
    t.ksize = ksize;
    auto in = reinterpret_cast<const float4 *>(in_rgba);
    auto tmp = reinterpret_cast<float4 *>(tmp_rgba);
    auto out = reinterpret_cast<float4 *>(out_rgba);
    cudaStream_t s = (cudaStream_t)stream;
    size_t smem = sizeof(float4) * (HB + 2 * (ksize >> 1));
    blur_h_kernel<<<cdiv(pw, HB) * (unsigned)ph, HB, smem, s>>>(in, tmp, pw, ph, t);
    if (int e = check_launch(where)) return e;
    blur_v_kernel<<<dim3(cdiv(pw, VBX), cdiv(ph, VBY)), dim3(VBX, VBY), 0, s>>>(tmp, out, pw, ph, t);
    return check_launch(where);
}
