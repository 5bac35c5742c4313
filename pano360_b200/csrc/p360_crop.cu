// Crop stage: the largest all-valid axis-aligned rectangle of the mosaic (stitcher.py:340-369,
// `crop_mosaic`, Numba-jitted in the reference), from the union of valid pixels the composite
// leaves in HBM.  The reference scans the rows top to bottom, keeps per column the height of the
// valid run that ends in the current row, and for every column j takes the widest span
// [lefts[j], rights[j]] over which no column is lower than column j; the first strictly larger
// area in (row, column) order wins.  Here:
//   crop_heights_kernel   per column: running heights of every row (one thread per column)
//   crop_row_kernel       per row (one block): nearest-lower column to the left and to the right of
//                         every column through a two-level min hierarchy over 32-column chunks
//                         (instead of the reference's serial pointer chase), best area of the row
//                         with the reference's tie rule (lowest column)
//   crop_pick_kernel      best row (lowest row among equal areas)
// The reference's quirk that rights[0] is never written (its loop stops at j = 1, :359) — column 0
// never extends to the right — is reproduced.
#include "p360_common.cuh"

namespace p360 {

struct RowBest {                 // best rectangle whose bottom edge lies in one row
    long long area;
    int col, left, right, height;
};

__global__ void __launch_bounds__(256)
crop_heights_kernel(const uint8_t *__restrict__ covered, int H, int W, int32_t *__restrict__ heights) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= W) return;
    int run = 0;
    for (int y = 0; y < H; ++y) {
        run = covered[(size_t)y * W + x] ? run + 1 : 0;      // stitcher.py:351-352
        heights[(size_t)y * W + x] = run;
    }
}

constexpr int CROP_THREADS = 1024;

// index of the nearest column k < j with h[k] < v, or -1
__device__ __forceinline__ int lower_to_the_left(const int32_t *h, const int32_t *m1, const int32_t *m2, int j, int v) {
    int k = j - 1;
    const int chunk0 = j & ~31;
    while (k >= chunk0 && h[k] >= v) --k;
    if (k >= chunk0) return k;
    int c = (j >> 5) - 1;                                // whole chunks to the left
    const int super0 = (j >> 10) << 5;                   // first chunk of j's group of 32 chunks
    while (c >= super0 && m1[c] >= v) --c;
    if (c < super0) {
        int s = (j >> 10) - 1;
        while (s >= 0 && m2[s] >= v) --s;
        if (s < 0) return -1;
        c = 32 * s + 31;
        while (m1[c] >= v) --c;                          // the group holds a lower chunk
    }
    k = 32 * c + 31;
    while (h[k] >= v) --k;                               // the chunk holds a lower column
    return k;
}

// index of the nearest column k > j with h[k] < v, or W
__device__ __forceinline__ int lower_to_the_right(const int32_t *h, const int32_t *m1, const int32_t *m2, int j, int v,
                                                  int W, int n1, int n2) {
    int k = j + 1;
    const int chunk1 = min((j | 31) + 1, W);
    while (k < chunk1 && h[k] >= v) ++k;
    if (k < chunk1) return k;
    int c = (j >> 5) + 1;
    const int super1 = min((((j >> 10) + 1) << 5), n1);
    while (c < super1 && m1[c] >= v) ++c;
    if (c >= super1) {
        int s = (j >> 10) + 1;
        while (s < n2 && m2[s] >= v) ++s;
        if (s >= n2) return W;
        c = 32 * s;
        while (m1[c] >= v) ++c;
    }
    k = 32 * c;
    while (h[k] >= v) ++k;
    return k;
}

// one block per mosaic row; dynamic shared memory: the row's heights + chunk minima (if `staged`),
// else the searches read the row from global memory (rows too wide for shared memory)
__global__ void __launch_bounds__(CROP_THREADS)
crop_row_kernel(const int32_t *__restrict__ heights, int W, int32_t *__restrict__ mins_global, int staged,
                RowBest *__restrict__ best_rows) {
    extern __shared__ int32_t crop_smem[];
    __shared__ long long red_area[CROP_THREADS / 32];
    __shared__ int red_col[CROP_THREADS / 32];
    const int row = blockIdx.x, tid = threadIdx.x;
    const int n1 = (W + 31) >> 5, n2 = (n1 + 31) >> 5;
    const int32_t *src = heights + (size_t)row * W;
    const int32_t *h;
    int32_t *m1, *m2;
    if (staged) {
        int32_t *hs = crop_smem;
        for (int i = tid; i < W; i += CROP_THREADS) hs[i] = src[i];
        h = hs; m1 = crop_smem + W; m2 = m1 + n1;
    } else {
        h = src; m1 = mins_global + (size_t)row * (n1 + n2); m2 = m1 + n1;
    }
    __syncthreads();
    for (int c = tid; c < n1; c += CROP_THREADS) {
        int v = INT32_MAX;
        for (int i = 32 * c; i < min(32 * c + 32, W); ++i) v = min(v, h[i]);
        m1[c] = v;
    }
    __syncthreads();
    for (int s = tid; s < n2; s += CROP_THREADS) {
        int v = INT32_MAX;
        for (int c = 32 * s; c < min(32 * s + 32, n1); ++c) v = min(v, m1[c]);
        m2[s] = v;
    }
    __syncthreads();
    long long area = 0;
    int col = INT32_MAX;
    for (int j = tid; j < W; j += CROP_THREADS) {         // ascending columns per thread: first maximum kept
        const int v = h[j];
        if (v == 0) continue;
        const int left = lower_to_the_left(h, m1, m2, j, v) + 1;                       // stitcher.py:353-356
        const int right = j == 0 ? 0 : lower_to_the_right(h, m1, m2, j, v, W, n1, n2) - 1;   // :357-360 (rights[0] stays 0)
        const long long a = (long long)(right - left + 1) * v;
        if (a > area) { area = a; col = j; }
    }
    // block maximum of (area, lowest column)
    for (int o = 16; o > 0; o >>= 1) {
        const long long oa = __shfl_xor_sync(0xffffffffu, area, o);
        const int oc = __shfl_xor_sync(0xffffffffu, col, o);
        if (oa > area || (oa == area && oc < col)) { area = oa; col = oc; }
    }
    if ((tid & 31) == 0) { red_area[tid >> 5] = area; red_col[tid >> 5] = col; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < CROP_THREADS / 32; ++w)
            if (red_area[w] > area || (red_area[w] == area && red_col[w] < col)) { area = red_area[w]; col = red_col[w]; }
        RowBest b{0, 0, 0, 0, 0};
        if (area > 0) {
            const int v = h[col];
            b.area = area; b.col = col; b.height = v;
            b.left = lower_to_the_left(h, m1, m2, col, v) + 1;
            b.right = col == 0 ? 0 : lower_to_the_right(h, m1, m2, col, v, W, n1, n2) - 1;
        }
        best_rows[row] = b;
    }
}

// the first row (top to bottom) that holds the largest area: strict '>' in the reference's scan (:363)
__global__ void __launch_bounds__(1024)
crop_pick_kernel(const RowBest *__restrict__ best_rows, int H, int32_t *__restrict__ out) {
    __shared__ long long red_area[32];
    __shared__ int red_row[32];
    long long area = 0;
    int row = INT32_MAX;
    for (int y = threadIdx.x; y < H; y += 1024) {
        const long long a = best_rows[y].area;
        if (a > area) { area = a; row = y; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const long long oa = __shfl_xor_sync(0xffffffffu, area, o);
        const int orow = __shfl_xor_sync(0xffffffffu, row, o);
        if (oa > area || (oa == area && orow < row)) { area = oa; row = orow; }
    }
    if ((threadIdx.x & 31) == 0) { red_area[threadIdx.x >> 5] = area; red_row[threadIdx.x >> 5] = row; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w)
            if (red_area[w] > area || (red_area[w] == area && red_row[w] < row)) { area = red_area[w]; row = red_row[w]; }
        if (area == 0) { out[0] = out[1] = out[2] = out[3] = 0; return; }     // nothing valid: an empty crop
        const RowBest b = best_rows[row];
        out[0] = row - b.height + 1; out[1] = row + 1;           // rows [y0, y1)
        out[2] = b.left; out[3] = b.right + 1;                   // columns [x0, x1)
    }
}

}  // namespace p360

extern "C" int64_t p360_crop_scratch_bytes(int H, int W) {
    if (H <= 0 || W <= 0) return 0;
    const int64_t n1 = (W + 31) / 32, n2 = (n1 + 31) / 32;
    // heights + per-row chunk minima (only used when a row does not fit shared memory) + per-row results
    return (int64_t)H * W * 4 + (int64_t)H * (n1 + n2) * 4 + (int64_t)H * (int64_t)sizeof(p360::RowBest) + 64;
}

extern "C" int p360_crop_rect(const uint8_t *covered, int H, int W, void *scratch, int32_t *rect_dev, void *stream) {
    using namespace p360;
    const char *where = "p360_crop_rect";
    P360_REQUIRE(covered && scratch && rect_dev && H > 0 && W > 0, where);
    P360_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 15) == 0, where);
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n1 = (W + 31) / 32, n2 = (n1 + 31) / 32;
    int32_t *heights = static_cast<int32_t *>(scratch);
    int32_t *mins = heights + (size_t)H * W;
    const size_t after = ((size_t)H * W + (size_t)H * (n1 + n2)) * 4;
    RowBest *rows = reinterpret_cast<RowBest *>(static_cast<char *>(scratch) + ((after + 15) & ~(size_t)15));
    crop_heights_kernel<<<cdiv(W, 256), 256, 0, s>>>(covered, H, W, heights);
    if (int e = check_launch(where)) return e;
    const size_t smem = (size_t)(W + n1 + n2) * 4;
    int dev = 0, limit = 0;
    P360_CUDA(cudaGetDevice(&dev), where);
    P360_CUDA(cudaDeviceGetAttribute(&limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev), where);
    const bool staged = smem + 1024 <= (size_t)limit;
    if (staged && smem > 48 * 1024)
        P360_CUDA(cudaFuncSetAttribute(crop_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), where);
    crop_row_kernel<<<H, CROP_THREADS, staged ? smem : 0, s>>>(heights, W, mins, staged ? 1 : 0, rows);
    if (int e = check_launch(where)) return e;
    crop_pick_kernel<<<1, 1024, 0, s>>>(rows, H, rect_dev);
    return check_launch(where);
}
