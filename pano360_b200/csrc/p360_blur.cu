// K3: separable float32 Gaussian on RGBA images, BORDER_REFLECT_101 at the
// image edges — the arithmetic of cv2.GaussianBlur(img, (0, 0), sigma) at
// stitcher.py:226.  Used on the coarse grids of the band pipeline and, as a
// general primitive, on full-resolution patches.
//
// Both passes stage their input tile (with halo, reflection applied while
// staging) in shared memory and give every thread R = 8 consecutive outputs
// along the filter axis: an input sample is read from shared memory once and
// scattered into the 8 accumulators it contributes to, so shared-memory
// traffic is (R + ksize - 1) / R loads per output instead of ksize, and the
// inner loop is pure packed FFMA2 (fma.rn.f32x2, two RGBA halves per tap).
#include "p360_common.cuh"
#include "p360_tma.cuh"

namespace p360 {

constexpr int R = 8;                              // outputs per thread along the filter axis
constexpr int TAP_SLOTS = P360_MAX_KSIZE + 3 * R; // taps, zero-padded by R-1 in front and 2R behind

struct Taps {
    float k[TAP_SLOTS];     // k[t + R - 1] = tap t
    int ksize;
};

__device__ __forceinline__ void fma_rgba(float2 &lo, float2 &hi, float w, const float4 &v) {
    const float2 ww = make_float2(w, w);
    lo = __ffma2_rn(ww, make_float2(v.x, v.y), lo);
    hi = __ffma2_rn(ww, make_float2(v.z, v.w), hi);
}

// out[i] = sum_t k[t] * in[i + t], i = 0..R-1, with in[j] = load(j), j = 0 .. R + ksize - 2.
template <class Load>
__device__ __forceinline__ void convolve_r(float2 (&lo)[R], float2 (&hi)[R], const Taps &t, Load load) {
    const int nj = R + t.ksize - 1;
    for (int jc = 0; jc < nj; jc += R) {
        float w[2 * R - 1];
#pragma unroll
        for (int c = 0; c < 2 * R - 1; ++c) w[c] = t.k[jc + c];
#pragma unroll
        for (int jj = 0; jj < R; ++jj) {
            const float4 v = load(jc + jj, jc + jj < nj);
#pragma unroll
            for (int i = 0; i < R; ++i) fma_rgba(lo[i], hi[i], w[jj - i + R - 1], v);
        }
    }
}

// The batched kernels skip blocks nobody reads, so a block that does run may stage cells that no
// kernel of this composite wrote.  Those cells only ever meet zero taps or outputs that are
// themselves unread, which is harmless as long as they hold FINITE values (0 * NaN != 0): the
// caller keeps the coarse pools in buffers that were zeroed once and are only ever written by
// these kernels (Compositor._coarse_pool).

// ---- horizontal ------------------------------------------------------------
constexpr int H_WARPS = 4;             // rows per block (one warp per row)
constexpr int H_SEG = 32 * R;          // outputs per warp
__host__ __device__ __forceinline__ int h_phys(int q) { return q + (q >> 3); }   // 1 pad slot per 8: lane stride 9

// ROWS rows per warp, H_SEG / ROWS outputs on each (ROWS = 1: one 256-wide segment per warp;
// ROWS = 4: four 64-wide ones, for the block lists of the seam-band maps, where a 1024-pixel
// segment across a 200-pixel seam band is mostly wasted work).  pitch = h_phys(segment + ksize - 1) + 1.
template <int ROWS>
__device__ __forceinline__ void blur_h_body(const float4 *__restrict__ in, float4 *__restrict__ out,
                                            int pw, int ph, int pitch, const Taps &t, int block) {
    constexpr int LPR = 32 / ROWS, SEG = H_SEG / ROWS;   // lanes, outputs per row
    extern __shared__ float4 smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPR, lr = lane % LPR;
    const int nxb = (pw + SEG - 1) / SEG;
    const int row = ((block / nxb) * H_WARPS + warp) * ROWS + sub;
    const int xb = (block % nxb) * SEG;
    if (ROWS == 1 && row >= ph) return;                  // warp-uniform, no block barrier below
    const bool live = row < ph;                          // ROWS > 1: rows of a warp end separately
    float4 *tile = smem + ((size_t)warp * ROWS + sub) * pitch;
    const int r = t.ksize >> 1;
    const float4 *src = in + (size_t)(live ? row : 0) * pw;
    const int nq = SEG + t.ksize - 1;
    if (live)
        for (int q = lr; q < nq; q += LPR) tile[h_phys(q)] = __ldg(src + reflect_101(xb - r + q, pw));
    __syncwarp();
    float2 lo[R], hi[R];
#pragma unroll
    for (int i = 0; i < R; ++i) lo[i] = hi[i] = make_float2(0.f, 0.f);
    const int base = R * lr;
    if (live)
        convolve_r(lo, hi, t, [&](int j, bool ok) {
            return ok ? tile[h_phys(base + j)] : make_float4(0.f, 0.f, 0.f, 0.f);
        });
    __syncwarp();
    if (live) {
#pragma unroll
        for (int i = 0; i < R; ++i)
            tile[h_phys(base + i)] = make_float4(lo[i].x, lo[i].y, hi[i].x, hi[i].y);
    }
    __syncwarp();
    float4 *dst = out + (size_t)(live ? row : 0) * pw + xb;
#pragma unroll
    for (int c = 0; c < R; ++c) {
        const int x = c * LPR + lr;
        if (live && xb + x < pw) dst[x] = tile[h_phys(x)];
    }
}

struct BlurJob {                       // == p360_blur_job
    const float4 *in;
    float4 *out;
    float4 *tmp;
    int w, h, slot;
    int shift;                         // log2 coarse factor
    const BandPatch *patch;            // the patch this coarse image belongs to (or nullptr: no skipping)
    int pad, grow;
};

// Does anybody read the coarse block [cx0, cx1) x [cy0, cy1) of this job?  With seam-band maps:
// a tile under it carries the patch's `need` bit; else: it lies within reach of the owned box.
__device__ __forceinline__ bool job_block_needed(const BlurJob &job, const TileMaps &maps,
                                                 int cx0, int cy0, int cx1, int cy1) {
    if (job.patch == nullptr) return true;
    const int s = job.shift, p = job.pad;
    const int xa = (cx0 << s) - p, ya = (cy0 << s) - p, xb = (cx1 << s) - p, yb = (cy1 << s) - p;   // patch px
    if (maps.need != nullptr) {
        const int x0 = __ldg(&job.patch->x0), y0 = __ldg(&job.patch->y0);
        return tiles_test(maps, maps.need, __ldg(&job.patch->index), xa + x0, ya + y0, xb + x0, yb + y0);
    }
    return near_owned(job.patch->own, job.grow, xa, ya, xb, yb);
}
static_assert(sizeof(BlurJob) == sizeof(p360_blur_job), "ABI struct mismatch");

__constant__ Taps c_taps[P360_MAX_LEVELS];     // tap sets of the batched blurs (p360_blur_set_taps)

__global__ void __launch_bounds__(32 * H_WARPS)
blur_h_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, int pw, int ph, int pitch, Taps t) {
    blur_h_body<1>(in, out, pw, ph, pitch, t, blockIdx.x);
}

// block `b` of a job's horizontal grid (H_SEG / ROWS cells x H_WARPS * ROWS rows): exists and is
// read by somebody?
template <int ROWS>
__device__ __forceinline__ bool h_block_needed(const BlurJob &job, const TileMaps &maps, int b) {
    constexpr int SEG = H_SEG / ROWS, BROWS = H_WARPS * ROWS;
    const int nxb = (job.w + SEG - 1) / SEG;
    if (b >= nxb * ((job.h + BROWS - 1) / BROWS)) return false;
    const int cx0 = (b % nxb) * SEG, cy0 = (b / nxb) * BROWS;
    return job_block_needed(job, maps, cx0, cy0, cx0 + SEG, cy0 + BROWS);
}

__global__ void __launch_bounds__(32 * H_WARPS)
blur_h_batch_kernel(const BlurJob *__restrict__ jobs, int pitch, TileMaps maps) {
    const BlurJob &job = jobs[blockIdx.y];
    if (!h_block_needed<1>(job, maps, blockIdx.x)) return;                            // block-uniform
    blur_h_body<1>(job.in, job.tmp, job.w, job.h, pitch, c_taps[job.slot], blockIdx.x);
}

// With seam-band maps: the needed blocks of the dense grids are compacted into a work list by
// one thread per block, and persistent grids walk the list (see reduce_scan_kernel).
template <int ROWS>
__global__ void __launch_bounds__(256)
blur_h_scan_kernel(const BlurJob *__restrict__ jobs, int n_jobs, int gx, TileMaps maps) {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= (long long)gx * n_jobs) return;
    const int b = (int)(t % gx), j = (int)(t / gx);
    if (!h_block_needed<ROWS>(jobs[j], maps, b)) return;
    const int at = atomicAdd(maps.work_count, 1);
    if (at < maps.work_cap) maps.work[at] = make_uint2((unsigned)j, (unsigned)b);
}

template <int ROWS>
__global__ void __launch_bounds__(32 * H_WARPS)
blur_h_list_kernel(const BlurJob *__restrict__ jobs, int pitch, TileMaps maps) {
    const int n = min(*maps.work_count, maps.work_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) maps.work_count[2] = n;      // (statistics for the bench)
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const uint2 item = maps.work[i];
        const BlurJob &job = jobs[item.x];
        blur_h_body<ROWS>(job.in, job.tmp, job.w, job.h, pitch, c_taps[job.slot], (int)item.y);
        __syncwarp();                   // the warp's tiles are restaged by the next item
    }
}

// ---- vertical --------------------------------------------------------------
constexpr int V_WARPS = 8;
constexpr int V_ROWS = V_WARPS * R;    // output rows per block, 32 columns wide

// One 32-column x V_ROWS-row tile.  Interior tiles (32 whole columns, no reflection above or
// below) are staged by the TMA engine: one 512-byte bulk copy per row, issued by the lanes of warp
// 0, completing on `bar` (phase parity in `phase`); edge tiles are staged by the threads with the
// reflection applied.  `bar` is initialised by the caller; every thread of the block calls this.
__device__ __forceinline__ void blur_v_body(const float4 *__restrict__ in, float4 *__restrict__ out,
                                            int pw, int ph, const Taps &t, int bxi, int byi,
                                            uint64_t *bar, uint32_t &phase) {
    extern __shared__ float4 smem[];   // [V_ROWS + ksize - 1][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = bxi * 32 + lane;
    const int yb = byi * V_ROWS;
    const int r = t.ksize >> 1;
    const int nq = V_ROWS + t.ksize - 1;
    const bool interior = kHaveTma && bxi * 32 + 32 <= pw && yb - r >= 0 && yb - r + nq <= ph;   // block-uniform
    if (interior) {
        if (warp == 0) {
            fence_async_smem();        // the tile's previous readers / writers are behind a block barrier
            if (lane == 0) mbar_expect_tx(bar, (uint32_t)nq * 512u);
            const float4 *src = in + (size_t)(yb - r) * pw + bxi * 32;
            for (int q = lane; q < nq; q += 32) bulk_g2s(smem + q * 32, src + (size_t)q * pw, 512u, bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
    } else {
        const int xs = min(x, pw - 1);
        for (int q = warp; q < nq; q += V_WARPS)
            smem[q * 32 + lane] = __ldg(in + (size_t)reflect_101(yb - r + q, ph) * pw + xs);
        __syncthreads();
    }
    float2 lo[R], hi[R];
#pragma unroll
    for (int i = 0; i < R; ++i) lo[i] = hi[i] = make_float2(0.f, 0.f);
    const int base = R * warp;
    convolve_r(lo, hi, t, [&](int j, bool ok) {
        return ok ? smem[(base + j) * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    });
    if (x >= pw) return;
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int y = yb + base + i;
        if (y < ph) out[(size_t)y * pw + x] = make_float4(lo[i].x, lo[i].y, hi[i].x, hi[i].y);
    }
}

// the block's mbarrier for the bulk copies: one arrival (the thread that announces the bytes)
__device__ __forceinline__ void v_barrier_init(uint64_t *bar) {
    if (kHaveTma && threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_async_smem();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(32 * V_WARPS)
blur_v_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, int pw, int ph, Taps t) {
    __shared__ uint64_t bar;
    uint32_t phase = 0u;
    v_barrier_init(&bar);
    blur_v_body(in, out, pw, ph, t, blockIdx.x, blockIdx.y, &bar, phase);
}

__device__ __forceinline__ bool v_block_needed(const BlurJob &job, const TileMaps &maps, int bxi, int byi) {
    if (bxi * 32 >= job.w || byi * V_ROWS >= job.h) return false;
    return job_block_needed(job, maps, bxi * 32, byi * V_ROWS, bxi * 32 + 32, byi * V_ROWS + V_ROWS);
}

__global__ void __launch_bounds__(32 * V_WARPS)
blur_v_batch_kernel(const BlurJob *__restrict__ jobs, TileMaps maps) {
    __shared__ uint64_t bar;
    const BlurJob &job = jobs[blockIdx.z];
    if (!v_block_needed(job, maps, blockIdx.x, blockIdx.y)) return;                   // block-uniform
    uint32_t phase = 0u;
    v_barrier_init(&bar);
    blur_v_body(job.tmp, job.out, job.w, job.h, c_taps[job.slot], blockIdx.x, blockIdx.y, &bar, phase);
}

__global__ void __launch_bounds__(256)
blur_v_scan_kernel(const BlurJob *__restrict__ jobs, int n_jobs, int gx, int gy, TileMaps maps) {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= (long long)gx * gy * n_jobs) return;
    const int bxi = (int)(t % gx), byi = (int)((t / gx) % gy), j = (int)(t / ((long long)gx * gy));
    if (!v_block_needed(jobs[j], maps, bxi, byi)) return;
    const int at = atomicAdd(maps.work_count, 1);
    if (at < maps.work_cap) maps.work[at] = make_uint2((unsigned)j, (unsigned)bxi | ((unsigned)byi << 16));
}

__global__ void __launch_bounds__(32 * V_WARPS)
blur_v_list_kernel(const BlurJob *__restrict__ jobs, TileMaps maps) {
    __shared__ uint64_t bar;
    const int n = min(*maps.work_count, maps.work_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) maps.work_count[3] = n;      // (statistics for the bench)
    uint32_t phase = 0u;
    v_barrier_init(&bar);
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const uint2 item = maps.work[i];
        const BlurJob &job = jobs[item.x];
        blur_v_body(job.tmp, job.out, job.w, job.h, c_taps[job.slot], (int)(item.y & 0xffffu),
                    (int)(item.y >> 16), &bar, phase);
        __syncthreads();                // the staged tile is replaced by the next item
    }
}

inline void fill_taps(Taps &t, const float *taps_host, int ksize) {
    memset(&t, 0, sizeof(t));
    memcpy(t.k + R - 1, taps_host, sizeof(float) * ksize);
    t.ksize = ksize;
}

// dynamic shared memory opt-in above 48 KB, remembered per kernel AND per device (the attribute
// is per device; one process may drive several GPUs)
struct SmemLimit {
    size_t bytes[64];
    SmemLimit() { for (size_t &b : bytes) b = 48 * 1024 - 256; }      // (static shared memory counts too: barriers)
};
template <class K>
int ensure_smem(K kernel, size_t bytes, SmemLimit &limit, const char *where) {
    int dev = 0;
    P360_CUDA(cudaGetDevice(&dev), where);
    size_t &have = limit.bytes[dev & 63];
    if (bytes > have) {
        P360_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), where);
        have = bytes;
    }
    return 0;
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
    fill_taps(t, taps_host, ksize);
    auto in = reinterpret_cast<const float4 *>(in_rgba);
    auto tmp = reinterpret_cast<float4 *>(tmp_rgba);
    auto out = reinterpret_cast<float4 *>(out_rgba);
    cudaStream_t s = (cudaStream_t)stream;
    const int pitch = h_phys(H_SEG + ksize - 1) + 1;
    const size_t smem_h = sizeof(float4) * pitch * H_WARPS;
    const size_t smem_v = sizeof(float4) * 32 * (V_ROWS + ksize - 1);
    static SmemLimit h_limit, v_limit;
    if (int e = ensure_smem(blur_h_kernel, smem_h, h_limit, where)) return e;
    if (int e = ensure_smem(blur_v_kernel, smem_v, v_limit, where)) return e;
    blur_h_kernel<<<cdiv(pw, H_SEG) * cdiv(ph, H_WARPS), 32 * H_WARPS, smem_h, s>>>(in, tmp, pw, ph, pitch, t);
    if (int e = check_launch(where)) return e;
    blur_v_kernel<<<dim3(cdiv(pw, 32), cdiv(ph, V_ROWS)), 32 * V_WARPS, smem_v, s>>>(tmp, out, pw, ph, t);
    return check_launch(where);
}

static int g_slot_ksize[P360_MAX_LEVELS] = {0};

extern "C" int p360_blur_set_taps(int slot, const float *taps_host, int ksize, void *stream) {
    using namespace p360;
    const char *where = "p360_blur_set_taps";
    P360_REQUIRE(slot >= 0 && slot < P360_MAX_LEVELS && taps_host, where);
    P360_REQUIRE(ksize >= 1 && (ksize & 1) && ksize <= P360_MAX_KSIZE, where);
    Taps t;
    fill_taps(t, taps_host, ksize);
    P360_CUDA(cudaMemcpyToSymbolAsync(c_taps, &t, sizeof(Taps), (size_t)slot * sizeof(Taps),
                                      cudaMemcpyHostToDevice, (cudaStream_t)stream), where);
    g_slot_ksize[slot] = ksize;
    return 0;
}

extern "C" int p360_gauss_blur_batch(const p360_blur_job *jobs, int n_jobs, int max_w, int max_h,
                                     const p360_tile_maps *maps_host, void *stream) {
    using namespace p360;
    const char *where = "p360_gauss_blur_batch";
    TileMaps maps;
    memset(&maps, 0, sizeof(maps));
    if (maps_host != nullptr) {
        memcpy(&maps, maps_host, sizeof(maps));
        P360_REQUIRE(maps.need && maps.work && maps.work_count && maps.work_cap > 0, where);
        P360_REQUIRE((reinterpret_cast<uintptr_t>(maps.work) & 7) == 0, where);
        P360_REQUIRE(maps.tiles_x > 0 && maps.tiles_y > 0 && maps.words > 0, where);
    }
    P360_REQUIRE(jobs && n_jobs >= 0 && n_jobs <= 65535 && max_w >= 0 && max_h >= 0, where);
    if (n_jobs == 0 || max_w == 0 || max_h == 0) return 0;
    int ksize = 1;
    for (int i = 0; i < P360_MAX_LEVELS; ++i) ksize = g_slot_ksize[i] > ksize ? g_slot_ksize[i] : ksize;
    cudaStream_t s = (cudaStream_t)stream;
    auto bj = reinterpret_cast<const BlurJob *>(jobs);
    const int pitch = h_phys(H_SEG + ksize - 1) + 1;
    const size_t smem_h = sizeof(float4) * pitch * H_WARPS;
    const size_t smem_v = sizeof(float4) * 32 * (V_ROWS + ksize - 1);
    static SmemLimit h_limit, v_limit;
    if (int e = ensure_smem(blur_h_batch_kernel, smem_h, h_limit, where)) return e;
    if (int e = ensure_smem(blur_v_batch_kernel, smem_v, v_limit, where)) return e;
    const unsigned h_blocks = cdiv(max_w, H_SEG) * cdiv(max_h, H_WARPS);
    dim3 grid_v(cdiv(max_w, 32), cdiv(max_h, V_ROWS), n_jobs);
    P360_REQUIRE(grid_v.y <= 65535, where);
    if (maps_host == nullptr) {
        blur_h_batch_kernel<<<dim3(h_blocks, n_jobs), 32 * H_WARPS, smem_h, s>>>(bj, pitch, maps);
        if (int e = check_launch(where)) return e;
        blur_v_batch_kernel<<<grid_v, 32 * V_WARPS, smem_v, s>>>(bj, maps);
        return check_launch(where);
    }
    // compacted work lists, persistent grids; the horizontal pass in 64-cell segments if asked for
    const bool narrow = maps.h_rows == 4;
    const int seg = narrow ? H_SEG / 4 : H_SEG, brows = narrow ? H_WARPS * 4 : H_WARPS;
    const int pitch_l = h_phys(seg + ksize - 1) + 1;
    const size_t smem_l = sizeof(float4) * pitch_l * (narrow ? 4 : 1) * H_WARPS;
    static SmemLimit hl_limit, hn_limit, vl_limit;
    if (int e = narrow ? ensure_smem(blur_h_list_kernel<4>, smem_l, hn_limit, where)
                       : ensure_smem(blur_h_list_kernel<1>, smem_l, hl_limit, where)) return e;
    if (int e = ensure_smem(blur_v_list_kernel, smem_v, vl_limit, where)) return e;
    const unsigned hl_blocks = cdiv(max_w, seg) * cdiv(max_h, brows);
    const long long h_cand = (long long)hl_blocks * n_jobs, v_cand = (long long)grid_v.x * grid_v.y * n_jobs;
    P360_REQUIRE(h_cand <= maps.work_cap && v_cand <= maps.work_cap && grid_v.x <= 65535, where);
    P360_CUDA(cudaMemsetAsync(maps.work_count, 0, sizeof(int), s), where);
    if (narrow) {
        blur_h_scan_kernel<4><<<cdiv(h_cand, 256), 256, 0, s>>>(bj, n_jobs, (int)hl_blocks, maps);
        if (int e = check_launch(where)) return e;
        blur_h_list_kernel<4><<<persistent_blocks(8), 32 * H_WARPS, smem_l, s>>>(bj, pitch_l, maps);
    } else {
        blur_h_scan_kernel<1><<<cdiv(h_cand, 256), 256, 0, s>>>(bj, n_jobs, (int)hl_blocks, maps);
        if (int e = check_launch(where)) return e;
        blur_h_list_kernel<1><<<persistent_blocks(8), 32 * H_WARPS, smem_l, s>>>(bj, pitch_l, maps);
    }
    if (int e = check_launch(where)) return e;
    P360_CUDA(cudaMemsetAsync(maps.work_count, 0, sizeof(int), s), where);
    blur_v_scan_kernel<<<cdiv(v_cand, 256), 256, 0, s>>>(bj, n_jobs, (int)grid_v.x, (int)grid_v.y, maps);
    if (int e = check_launch(where)) return e;
    blur_v_list_kernel<<<persistent_blocks(4), 32 * V_WARPS, smem_v, s>>>(bj, maps);
    return check_launch(where);
}
