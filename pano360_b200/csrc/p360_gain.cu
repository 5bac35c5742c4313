// K8: pair overlap statistics for the exposure-gain solve
// (stitcher.py:48-63).  The reference warps image j into image i's frame with
// cv2.warpPerspective (INTER_LINEAR, BORDER_TRANSPARENT, zero destination —
// SURVEY.md F7), then takes the overlap count and two mean intensities.  Here
// nothing is materialised: every thread maps one pixel of i, samples j, and
// the three sums are reduced with warp shuffles -> one partial per block ->
// a fixed-order final pass (deterministic).  All pairs of a panorama run in one
// launch (grid.y = pair).
#include "p360_common.cuh"

namespace p360 {

struct PairJob {                     // == p360_pair_job
    const uint8_t *src_i;            // image i (destination frame)
    const uint8_t *src_j;            // image j (warped into i's frame)
    double inv[9];                   // INVERSE of the i <- j homography, un-centred pixel coordinates
};
static_assert(sizeof(PairJob) == sizeof(p360_pair_job), "ABI struct mismatch");

constexpr int GB = 256;
constexpr int GAIN_PIX_PER_THREAD = 4;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int sat_round_int(double v) {     // saturate_cast<int>(double)
    v = fmin(fmax(v, -2147483648.0), 2147483647.0);
    return __double2int_rn(v);
}

// grid = (blocks per image, pairs): every pair of the panorama in one launch.
__global__ void __launch_bounds__(GB)
pair_stats_kernel(const PairJob *__restrict__ jobs, int h, int w, int sc, const float *__restrict__ lut,
                  const double *__restrict__ hat_y, const double *__restrict__ hat_x,
                  double *__restrict__ partial) {
    __shared__ double red[3][GB / 32];
    const PairJob &job = jobs[blockIdx.y];
    const uint8_t *src_i = job.src_i, *src_j = job.src_j;
    double m[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = job.inv[k];
    double cnt = 0.0, sum_i = 0.0, sum_j = 0.0;
    long long n = (long long)h * w;
    long long base = ((long long)blockIdx.x * GB) * GAIN_PIX_PER_THREAD + threadIdx.x;
#pragma unroll
    for (int it = 0; it < GAIN_PIX_PER_THREAD; ++it) {
        long long p = base + (long long)it * GB;
        if (p >= n) break;
        int y = (int)(p / w), x = (int)(p - (long long)y * w);
        // WarpPerspectiveInvoker: X = saturate<int>(rint(32 * X0 / W0)), double, products and
        // sums separately rounded like the compiled C++ (no FMA contraction)
        double xd = __dadd_rn(__dadd_rn(__dmul_rn(m[0], x), __dmul_rn(m[1], y)), m[2]);
        double yd = __dadd_rn(__dadd_rn(__dmul_rn(m[3], x), __dmul_rn(m[4], y)), m[5]);
        double wd = __dadd_rn(__dadd_rn(__dmul_rn(m[6], x), __dmul_rn(m[7], y)), m[8]);
        wd = wd != 0.0 ? 32.0 / wd : 0.0;
        int fx = sat_round_int(__dmul_rn(xd, wd)), fy = sat_round_int(__dmul_rn(yd, wd));
        int ix = sat16(fx >> 5), iy = sat16(fy >> 5);
        if (ix < 0 || ix > w - 1 || iy < 0 || iy > h - 1) continue;   // destination untouched (zero)
        int ix1 = min(ix + 1, w - 1), iy1 = min(iy + 1, h - 1);        // right/bottom taps replicate
        float ax = (float)(fx & 31) * 0.03125f, ay = (float)(fy & 31) * 0.03125f;
        float w00 = __fmul_rn(1.0f - ay, 1.0f - ax), w01 = __fmul_rn(1.0f - ay, ax);
        float w10 = __fmul_rn(ay, 1.0f - ax), w11 = __fmul_rn(ay, ax);
        double hy0 = __ldg(hat_y + iy), hy1 = __ldg(hat_y + iy1);
        double hx0 = __ldg(hat_x + ix), hx1 = __ldg(hat_x + ix1);
        float a = __fmul_rn((float)(hy0 * hx0), w00);
        a = __fadd_rn(a, __fmul_rn((float)(hy0 * hx1), w01));
        a = __fadd_rn(a, __fmul_rn((float)(hy1 * hx0), w10));
        a = __fadd_rn(a, __fmul_rn((float)(hy1 * hx1), w11));
        if (a == 0.0f) continue;                                       // stitcher.py:58
        const uint8_t *p00 = src_j + ((size_t)iy * w + ix) * sc, *p01 = src_j + ((size_t)iy * w + ix1) * sc;
        const uint8_t *p10 = src_j + ((size_t)iy1 * w + ix) * sc, *p11 = src_j + ((size_t)iy1 * w + ix1) * sc;
        const uint8_t *pi = src_i + (size_t)p * sc;
        float sj = 0.f, si = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float v = __fmul_rn(__ldg(lut + __ldg(p00 + ch)), w00);
            v = __fadd_rn(v, __fmul_rn(__ldg(lut + __ldg(p01 + ch)), w01));
            v = __fadd_rn(v, __fmul_rn(__ldg(lut + __ldg(p10 + ch)), w10));
            v = __fadd_rn(v, __fmul_rn(__ldg(lut + __ldg(p11 + ch)), w11));
            sj += v;
            si += __ldg(lut + __ldg(pi + ch));
        }
        cnt += 1.0; sum_i += (double)si; sum_j += (double)sj;
    }
    cnt = warp_sum(cnt); sum_i = warp_sum(sum_i); sum_j = warp_sum(sum_j);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { red[0][wid] = cnt; red[1][wid] = sum_i; red[2][wid] = sum_j; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int k = 0; k < GB / 32; ++k) s += red[threadIdx.x][k];
        partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = s;
    }
}

__global__ void pair_stats_final_kernel(const double *__restrict__ partial, int nblocks,
                                        double *__restrict__ out) {
    // one block of 3 warps per pair, one warp per statistic; lane-strided partial sums then a
    // shuffle tree: the order is fixed by nblocks alone, so the result is reproducible.
    int stat = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double *mine = partial + (size_t)blockIdx.x * nblocks * 3;
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += mine[(size_t)b * 3 + stat];
    s = warp_sum(s);
    if (lane == 0) out[(size_t)blockIdx.x * 3 + stat] = s;
}

}  // namespace p360

extern "C" int p360_pair_stats_blocks(int h, int w) {
    using namespace p360;
    if (h <= 0 || w <= 0) return 0;
    return (int)cdiv((long long)h * w, (long long)GB * GAIN_PIX_PER_THREAD);
}

extern "C" int p360_pair_overlap_stats(const p360_pair_job *jobs, int n_pairs, int h, int w, int src_c,
                                       const float *lut, const double *hat_y, const double *hat_x,
                                       double *partial, double *out, void *stream) {
    using namespace p360;
    const char *where = "p360_pair_overlap_stats";
    P360_REQUIRE(jobs && lut && hat_y && hat_x && partial && out, where);
    P360_REQUIRE(n_pairs >= 0 && n_pairs <= 65535, where);
    P360_REQUIRE(h > 0 && w > 0 && (src_c == 3 || src_c == 4 || src_c == 8), where);
    if (n_pairs == 0) return 0;
    int nblocks = p360_pair_stats_blocks(h, w);
    cudaStream_t s = (cudaStream_t)stream;
    pair_stats_kernel<<<dim3(nblocks, n_pairs), GB, 0, s>>>(reinterpret_cast<const PairJob *>(jobs), h, w, src_c,
                                                            lut, hat_y, hat_x, partial);
    if (int e = check_launch(where)) return e;
    pair_stats_final_kernel<<<n_pairs, 96, 0, s>>>(partial, nblocks, out);
    return check_launch(where);
}
