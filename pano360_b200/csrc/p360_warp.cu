// K1: fused inverse projection + 1/32-px bilinear remap + validity mask
// (+ owner-map competition).  Replaces stitcher.py:257-263 and :300-317 (NumPy
// coordinate maths, BLAS 3x3 projection, cv2.remap, alpha masking) and the
// weights tensor / argmax of :196-204 with one pass that reads the u8 source
// through L1 and writes each RGBA float4 exactly once, coalesced.
// One launch serves every patch of a composite (grid.z = job).
#include "p360_common.cuh"

namespace p360 {

struct WarpJob {                 // == p360_warp_job
    const uint8_t *src;          // u8 [h][w][c]
    const float *lut;            // 256 entries
    const double *hat_y;         // h entries
    const double *hat_x;         // w entries
    const double *ray_x;         // per mosaic column: x component of proj2hom (sin theta)
    const double *ray_z;         // per mosaic column: z component (cos theta)
    const double *ray_y;         // per mosaic row: y component (tan phi / phi)
    float4 *out;                 // ph x pw RGBA
    uint8_t *invalid;            // ph x pw
    double kr[9];                // K * R, row-major (bundle_adj.py:31-33)
    int h, w, c;
    int pw, ph;
    int x0, y0;                  // position in the (window) mosaic
    int col0, row0;              // absolute mosaic column / row of the patch origin (ray tables)
    int patch;                   // id in the owner map
    float half_w, half_h;        // float32(w / 2), float32(h / 2)          (stitcher.py:310)
    float max_x, max_y;          // float32(w - 1), float32(h - 1)          (stitcher.py:311-312)
    float inv_2w, inv_2h;        // 1 / (2w), 1 / (2h): reflection period reciprocals
    int ty0, ty1;                // rows / columns of the image's TRUE box (for a seam-split image: of this column
    int tx0, tx1;                // run) in window coordinates: a window crops x0 / y0 / pw / ph, and the seam plan
                                 // must not depend on where the window was cut
};
static_assert(sizeof(WarpJob) == sizeof(p360_warp_job), "ABI struct mismatch");

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

__device__ __forceinline__ float lut_at(const float *lut, uint32_t byte_offset) {
    return *reinterpret_cast<const float *>(reinterpret_cast<const char *>(lut) + byte_offset);
}

// ---- the warp of one pixel in three phases, so that the gathers of all the pixels a thread
// handles can be in flight together (the kernel is bound by their latency otherwise) ---------

struct TapPlan {            // phase 1: where to sample and with which weights
    int off00, off01, off10, off11;       // pixel offsets (y * w + x) of the four taps
    float w00, w01, w10, w11;             // bilinear weights, products of 1/32 fractions (exact)
    bool bad;
};
struct TapAlpha { float a00, a01, a10, a11; };   // alpha of the four taps: float32(hat_y * hat_x), stitcher.py:261

// What a thread needs of its mosaic column for every row it visits: the z component of the ray and
// the first products of p = K R (rx, ry, rz) — float64, k = 0, 1, 2 in order (stitcher.py:303-306).
struct ColumnTerms { double rz, x0, x1, x2; };
__device__ __forceinline__ ColumnTerms column_terms(const WarpJob &s, int c) {
    const double rx = __ldg(s.ray_x + s.col0 + c);
    return ColumnTerms{__ldg(s.ray_z + s.col0 + c), s.kr[0] * rx, s.kr[3] * rx, s.kr[6] * rx};
}

template <bool ALPHA>
__device__ __forceinline__ TapPlan plan_taps(const WarpJob &s, const ColumnTerms &col, double ry, TapAlpha &alpha) {
    const float px = (float)fma(s.kr[2], col.rz, fma(s.kr[1], ry, col.x0));     // then cast to float32 (stitcher.py:306)
    const float py = (float)fma(s.kr[5], col.rz, fma(s.kr[4], ry, col.x1));
    const float pz = (float)fma(s.kr[8], col.rz, fma(s.kr[7], ry, col.x2));
    TapPlan t;
    t.bad = pz < 0.0f;                                           // stitcher.py:308
    const float x = __fadd_rn(__fdiv_rn(px, pz), s.half_w);      // stitcher.py:310
    const float y = __fadd_rn(__fdiv_rn(py, pz), s.half_h);
    t.bad |= (x < 0.0f) | (x > s.max_x) | (y < 0.0f) | (y > s.max_y);   // :311-312
    const int sx = to_fixed5(x), sy = to_fixed5(y);
    // integer parts; inside the image (w, h <= 32767) the int16 saturation of OpenCV's map is the identity
    int x0 = sx >> 5, y0 = sy >> 5;
    int x1 = x0 + 1, y1 = y0 + 1;
    if ((unsigned)x0 >= (unsigned)(s.w - 1) || (unsigned)y0 >= (unsigned)(s.h - 1)) {
        // a tap falls outside the image: BORDER_REFLECT (cv2.remap at stitcher.py:315-316)
        const int ix = sat16(x0), iy = sat16(y0);
        x0 = reflect_fast(ix, s.w, s.inv_2w); x1 = reflect_fast(ix + 1, s.w, s.inv_2w);
        y0 = reflect_fast(iy, s.h, s.inv_2h); y1 = reflect_fast(iy + 1, s.h, s.inv_2h);
    }
    t.off00 = y0 * s.w + x0; t.off01 = y0 * s.w + x1;
    t.off10 = y1 * s.w + x0; t.off11 = y1 * s.w + x1;
    const float ax = (float)(sx & 31) * 0.03125f, ay = (float)(sy & 31) * 0.03125f;
    t.w00 = __fmul_rn(1.0f - ay, 1.0f - ax); t.w01 = __fmul_rn(1.0f - ay, ax);
    t.w10 = __fmul_rn(ay, 1.0f - ax); t.w11 = __fmul_rn(ay, ax);
    if (ALPHA) {            // the weight image of _add_weights, evaluated at the taps (float64 product, then cast)
        const double hy0 = __ldg(s.hat_y + y0), hy1 = __ldg(s.hat_y + y1);
        const double hx0 = __ldg(s.hat_x + x0), hx1 = __ldg(s.hat_x + x1);
        alpha.a00 = (float)(hy0 * hx0); alpha.a01 = (float)(hy0 * hx1);
        alpha.a10 = (float)(hy1 * hx0); alpha.a11 = (float)(hy1 * hx1);
    }
    return t;
}

// One tap: the three u8 samples of a source pixel in the low 24 bits.  RGBX: the source is the
// 4-byte-per-pixel layout of p360_pack_rgbx (or a 4-channel upload): one aligned 32-bit load.
template <bool RGBX>
__device__ __forceinline__ uint32_t load_tap(const WarpJob &s, int off) {
    if (RGBX) return __ldg(reinterpret_cast<const uint32_t *>(s.src) + off);
    const uint8_t *p = s.src + (size_t)off * 3;
    return (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
}

// The two taps of one source row.  RGBX: two aligned 32-bit loads.  Packed u8 x 3 (as uploaded):
// the usual case — the second tap is the next pixel — reads the six bytes through three aligned
// 32-bit words and two funnel shifts; anything else (reflection at the border, the last pixels of
// the image) takes byte loads.  `n_px` = pixels in the image.
template <bool RGBX>
__device__ __forceinline__ void load_row_taps(const WarpJob &s, int off_a, int off_b, uint32_t &qa, uint32_t &qb) {
    if (RGBX) {
        qa = __ldg(reinterpret_cast<const uint32_t *>(s.src) + off_a);
        qb = __ldg(reinterpret_cast<const uint32_t *>(s.src) + off_b);
        return;
    }
    if (off_b == off_a + 1 && off_a + 4 <= s.h * s.w) {
        const unsigned b0 = 3u * (unsigned)off_a;
        const uint32_t *w = reinterpret_cast<const uint32_t *>(s.src) + (b0 >> 2);
        const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
        const unsigned sh = (b0 & 3u) * 8u;
        qa = __funnelshift_r(w0, w1, sh) & 0xffffffu;
        qb = (sh ? __funnelshift_r(w1, w2, sh - 8u) : __funnelshift_r(w0, w1, 24u)) & 0xffffffu;
        return;
    }
    qa = load_tap<false>(s, off_a);
    qb = load_tap<false>(s, off_b);
}

// ((s00*w00 + s01*w01) + s10*w10) + s11*w11, separately rounded products and
// sums: the exact evaluation order of OpenCV's remapBilinear float path.
__device__ __forceinline__ float blend4(float a, float b, float c, float d, const TapPlan &t) {
    float acc = __fmul_rn(a, t.w00);
    acc = __fadd_rn(acc, __fmul_rn(b, t.w01));
    acc = __fadd_rn(acc, __fmul_rn(c, t.w10));
    acc = __fadd_rn(acc, __fmul_rn(d, t.w11));
    return acc;
}

// phase 3: LUT (u8 -> float exactly as the reference's float image holds it) + bilinear blend
__device__ __forceinline__ float3 finish_rgb(const float *lut, const TapPlan &t, const uint32_t (&q)[4]) {
    float3 o;
    o.x = blend4(lut_at(lut, (q[0] << 2) & 0x3fc), lut_at(lut, (q[1] << 2) & 0x3fc),
                 lut_at(lut, (q[2] << 2) & 0x3fc), lut_at(lut, (q[3] << 2) & 0x3fc), t);
    o.y = blend4(lut_at(lut, (q[0] >> 6) & 0x3fc), lut_at(lut, (q[1] >> 6) & 0x3fc),
                 lut_at(lut, (q[2] >> 6) & 0x3fc), lut_at(lut, (q[3] >> 6) & 0x3fc), t);
    o.z = blend4(lut_at(lut, (q[0] >> 14) & 0x3fc), lut_at(lut, (q[1] >> 14) & 0x3fc),
                 lut_at(lut, (q[2] >> 14) & 0x3fc), lut_at(lut, (q[3] >> 14) & 0x3fc), t);
    return o;
}
__device__ __forceinline__ float4 finish_pixel(const float *lut, const TapPlan &t, const TapAlpha &a,
                                               const uint32_t (&q)[4]) {
    const float3 c = finish_rgb(lut, t, q);
    float4 o = make_float4(c.x, c.y, c.z, blend4(a.a00, a.a01, a.a10, a.a11, t));
    if (t.bad) o.w = 0.0f;                                       // stitcher.py:317
    return o;
}

constexpr int WARP_BX = 64, WARP_BY = 4;
constexpr int WARP_ROWS = 4;            // rows per thread: 16 gathers in flight before the first use

template <bool OWNER>
__device__ __forceinline__ void warp_commit(const WarpJob &s, int c, int r, const float4 &o, bool bad,
                                            unsigned long long *keys, uint8_t *covered, int W) {
    const size_t idx = (size_t)r * s.pw + c;
    st_stream(s.out + idx, o);
    s.invalid[idx] = bad ? 1 : 0;
    if (OWNER) {
        const size_t mi = (size_t)(r + s.y0) * W + (c + s.x0);
        owner_compete(keys, mi, o.w, s.patch);                   // stitcher.py:196-204
        if (!bad) covered[mi] = 1;                               // stitcher.py:233-234
    }
}

constexpr int WARP_JOBS_PER_LAUNCH = 128;
constexpr int TILE_JOBS_MAX = 256;                 // p360_warp_tiles: the whole job table sits in constant memory
__constant__ WarpJob c_warp_jobs[TILE_JOBS_MAX];   // block-uniform reads: no LSU traffic per pixel, and the compiler
                                                   // may re-read a field instead of holding it in a register

// RGBX: every job's source has 4 bytes per pixel; OWNER: owner keys wanted.
template <bool RGBX, bool OWNER>
__device__ __forceinline__ void warp_block(const WarpJob &job, float *lut, unsigned long long *__restrict__ keys,
                                           uint8_t *__restrict__ covered, int W) {
    const int r0 = blockIdx.y * (WARP_BY * WARP_ROWS);
    const int tid = threadIdx.y * WARP_BX + threadIdx.x;
    lut[tid] = __ldg(job.lut + tid);
    __syncthreads();
    const int c = blockIdx.x * WARP_BX + threadIdx.x;
    if (c >= job.pw) return;
    // coordinates of every row first, then all gathers back to back, then LUT + blend,
    // then the stores and the owner competition
    TapPlan plan[WARP_ROWS];
    TapAlpha alpha[WARP_ROWS];
    uint32_t taps[WARP_ROWS][4];
    const ColumnTerms col = column_terms(job, c);
#pragma unroll
    for (int k = 0; k < WARP_ROWS; ++k)
        plan[k] = plan_taps<true>(job, col, __ldg(job.ray_y + job.row0 + min(r0 + (int)threadIdx.y + k * WARP_BY, job.ph - 1)),
                                  alpha[k]);
#pragma unroll
    for (int k = 0; k < WARP_ROWS; ++k) {
        load_row_taps<RGBX>(job, plan[k].off00, plan[k].off01, taps[k][0], taps[k][1]);
        load_row_taps<RGBX>(job, plan[k].off10, plan[k].off11, taps[k][2], taps[k][3]);
    }
#pragma unroll
    for (int k = 0; k < WARP_ROWS; ++k) {
        const int r = r0 + threadIdx.y + k * WARP_BY;
        if (r < job.ph)
            warp_commit<OWNER>(job, c, r, finish_pixel(lut, plan[k], alpha[k], taps[k]), plan[k].bad, keys, covered, W);
    }
}

template <bool RGBX, bool OWNER>
__global__ void __launch_bounds__(WARP_BX *WARP_BY)
warp_batch_kernel(unsigned long long *__restrict__ keys, uint8_t *__restrict__ covered, int W) {
    __shared__ float lut[256];
    const WarpJob &job = c_warp_jobs[blockIdx.z];
    const int r0 = blockIdx.y * (WARP_BY * WARP_ROWS);
    if ((int)(blockIdx.x * WARP_BX) >= job.pw || r0 >= job.ph) return;   // block-uniform
    warp_block<RGBX, OWNER>(job, lut, keys, covered, W);
}

// ---- K0: the seam plan — who can own a pixel of a tile?  (geometry only, before anything is
// sampled) ----------------------------------------------------------------------------------
// Ownership is arg-max of alpha = hat_y(v) * hat_x(u).  Interval arithmetic over a 64 x 32 tile —
// ray tables -> K R ray -> (u, v) -> alpha — bounds alpha of every patch on the tile; a patch whose
// upper bound lies below another patch's lower bound can never win there.  Every true owner is a
// candidate (tools/gate_bounds.py checks it against the owner keys on random rigs); at cfg4 92 % of
// the tiles keep a single candidate.  Everything downstream is planned from these bitmaps:
//   present  candidates for owning a pixel of the tile
//   cand     candidates within the blur reach whose box meets the tile: the only patches that can
//            carry weight there.  One bit set ("solo"): the multiband sum telescopes to that
//            patch's warped pixel on the whole tile -> p360_warp_tiles writes uint8 straight into
//            the mosaic and nothing else ever touches the tile.  More ("multi"): full blend.
//   need     candidates of the multi tiles within reach: where coarse levels are consumed
//   wneed    where float pixels / owner keys are consumed: candidates of the multi tiles within
//            reach + one tile (block overhang; mirror images at patch edges are no farther from
//            the consuming tile than the position they stand for), plus — in every such tile —
//            all its own `present` patches, so that the owner keys are right wherever they are read
// Tiles are evaluated on their full extent in absolute mosaic rows against the images' TRUE boxes,
// so the plan of a row window agrees with the plan of the whole mosaic on every tile that lies
// (with its reach) inside the window.
struct Interval { double lo, hi; };
__device__ __forceinline__ Interval scaled(double k, Interval v) {
    const double a = k * v.lo, b = k * v.hi;
    return Interval{fmin(a, b), fmax(a, b)};
}
__device__ __forceinline__ Interval table_range(const double *__restrict__ t, int a, int b) {   // t[a .. b)
    Interval r{t[a], t[a]};
    for (int i = a + 1; i < b; ++i) { r.lo = fmin(r.lo, t[i]); r.hi = fmax(r.hi, t[i]); }
    return r;
}
// hat over source coordinates [lo, hi] clipped to [0, size - 1]: lower bound of the interpolated
// table (concave: the smaller end, at the surrounding integers), upper bound of its envelope
__device__ __forceinline__ void hat_range(double lo, double hi, int size, double &mn, double &mx,
                                          bool &any_valid, bool &all_valid) {
    any_valid = hi >= 0.0 && lo <= size - 1.0;
    all_valid = lo >= 0.0 && hi <= size - 1.0;
    const double a = fmin(fmax(lo, 0.0), size - 1.0), b = fmin(fmax(hi, 0.0), size - 1.0), c = 0.5 * size;
    mx = (a <= c && b >= c) ? 0.5 : fmax(0.5 - fabs(a - c) / size, 0.5 - fabs(b - c) / size);
    const double la = fmin(0.5 - fabs(floor(a) - c) / size, 0.5 - fabs(ceil(a) - c) / size);
    const double lb = fmin(0.5 - fabs(floor(b) - c) / size, 0.5 - fabs(ceil(b) - c) / size);
    mn = fmin(la, lb);
}
// alpha bounds of one patch on one tile; false if the patch has no valid pixel there
__device__ __forceinline__ bool alpha_range(const WarpJob &j, Interval rx, Interval ry, Interval rz,
                                            double &a_min, double &a_max) {
    Interval p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const Interval x = scaled(j.kr[3 * k], rx), y = scaled(j.kr[3 * k + 1], ry), z = scaled(j.kr[3 * k + 2], rz);
        p[k] = Interval{x.lo + y.lo + z.lo, x.hi + y.hi + z.hi};
    }
    if (p[2].hi <= 0.0) return false;                      // behind the camera
    if (p[2].lo <= 1e-9) { a_min = 0.0; a_max = 0.25; return true; }   // partly behind: anything goes
    const double slack = 1.0 / 32 + 1e-3;                  // fixed-point sampling + float32 quotient
    double u0 = fmin(fmin(p[0].lo / p[2].lo, p[0].lo / p[2].hi), fmin(p[0].hi / p[2].lo, p[0].hi / p[2].hi));
    double u1 = fmax(fmax(p[0].lo / p[2].lo, p[0].lo / p[2].hi), fmax(p[0].hi / p[2].lo, p[0].hi / p[2].hi));
    double v0 = fmin(fmin(p[1].lo / p[2].lo, p[1].lo / p[2].hi), fmin(p[1].hi / p[2].lo, p[1].hi / p[2].hi));
    double v1 = fmax(fmax(p[1].lo / p[2].lo, p[1].lo / p[2].hi), fmax(p[1].hi / p[2].lo, p[1].hi / p[2].hi));
    double xn, xx, yn, yx;
    bool x_any, x_all, y_any, y_all;
    hat_range(u0 + 0.5 * j.w - slack, u1 + 0.5 * j.w + slack, j.w, xn, xx, x_any, x_all);
    hat_range(v0 + 0.5 * j.h - slack, v1 + 0.5 * j.h + slack, j.h, yn, yx, y_any, y_all);
    if (!(x_any && y_any)) return false;
    a_max = xx * yx * (1.0 + 1e-5);
    a_min = (x_all && y_all) ? xn * yn * (1.0 - 1e-5) : 0.0;
    return true;
}

// one thread per tile; jobs = DEVICE copy of the warp jobs (patch id = position).  [0, H) are the
// rows of the window buffer, [-abs_row0, mosaic_h - abs_row0) those of the whole mosaic in the
// same coordinates.  Also grows the `own` boxes of the band patches (box around the tiles a patch
// may own a pixel of, clipped to its cropped box).
__global__ void __launch_bounds__(128)
seam_candidates_kernel(const WarpJob *__restrict__ jobs, int n_jobs, BandPatch *patches, int H, int W,
                       int abs_row0, int mosaic_h, TileMaps m) {
    const int t = blockIdx.x * 128 + threadIdx.x;
    if (t >= m.tiles_x * m.tiles_y) return;
    const int tx = t % m.tiles_x, ty = t / m.tiles_x;
    const int xa = tx * TILE_X, xb = min(xa + TILE_X, W);
    const int ta = m.row0 + ty * TILE_Y;                                     // tile rows (window coordinates)
    const int ya = max(ta, -abs_row0), yb = min(ta + TILE_Y, mosaic_h - abs_row0);   // ... inside the mosaic
    for (int w = 0; w < m.words; ++w) m.present[(size_t)t * m.words + w] = 0u;
    if (yb <= ya || n_jobs == 0) return;
    // the ray tables are shared by all jobs: absolute mosaic column / row = col0 - x0 + x, row0 - y0 + y
    const WarpJob &j0 = jobs[0];
    const int dc = j0.col0 - j0.x0, dr = j0.row0 - j0.y0;
    const Interval rx = table_range(j0.ray_x, xa + dc, xb + dc), rz = table_range(j0.ray_z, xa + dc, xb + dc);
    const Interval ry = table_range(j0.ray_y, ya + dr, yb + dr);
    double best_min = 0.0;
    for (int k = 0; k < n_jobs; ++k) {
        const WarpJob &j = jobs[k];
        if (j.tx0 >= xb || j.tx1 <= xa || j.ty0 >= yb || j.ty1 <= ya) continue;
        // a patch only dominates a tile it covers completely: beyond its box it has no pixels,
        // however large alpha would be there (boxes end where the reference's ranges end)
        if (j.tx0 > xa || j.tx1 < xb || j.ty0 > ya || j.ty1 < yb) continue;
        double a_min, a_max;
        if (alpha_range(j, rx, ry, rz, a_min, a_max)) best_min = fmax(best_min, a_min);
    }
    for (int k = 0; k < n_jobs; ++k) {
        const WarpJob &j = jobs[k];
        if (j.tx0 >= xb || j.tx1 <= xa || j.ty0 >= yb || j.ty1 <= ya) continue;
        double a_min, a_max;
        if (alpha_range(j, rx, ry, rz, a_min, a_max) && a_max >= best_min) {
            m.present[(size_t)t * m.words + (j.patch >> 5)] |= 1u << (j.patch & 31);
            if (patches != nullptr) {
                BandPatch &bp = patches[j.patch];
                atomicMin(&bp.own[0], max(xa - bp.x0, 0));
                atomicMin(&bp.own[1], max(max(ta, 0) - bp.y0, 0));
                atomicMax(&bp.own[2], min(xa + TILE_X - bp.x0, bp.pw));
                atomicMax(&bp.own[3], min(ta + TILE_Y - bp.y0, bp.ph));
            }
        }
    }
}

// cand(T) = present dilated by the blur reach, restricted to the patches whose (true) box meets T;
// multi(T) = |cand(T)| > 1.  One thread per tile.
__global__ void __launch_bounds__(256)
seam_cand_kernel(const WarpJob *__restrict__ jobs, int n_jobs, int W, TileMaps m) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= m.tiles_x * m.tiles_y) return;
    const int tx = t % m.tiles_x, ty = t / m.tiles_x;
    const int x0 = max(tx - m.reach_x, 0), x1 = min(tx + m.reach_x, m.tiles_x - 1);
    const int y0 = max(ty - m.reach_y, 0), y1 = min(ty + m.reach_y, m.tiles_y - 1);
    const int xa = tx * TILE_X, xb = min(xa + TILE_X, W), ta = m.row0 + ty * TILE_Y;
    int count = 0;
    for (int w = 0; w < m.words; ++w) {
        uint32_t bits = 0u;
        for (int y = y0; y <= y1; ++y)
            for (int x = x0; x <= x1; ++x) bits |= __ldg(m.present + ((size_t)y * m.tiles_x + x) * m.words + w);
        uint32_t keep = 0u;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int k = 32 * w + b;
            if (k >= n_jobs) break;
            const WarpJob &j = jobs[k];
            if (j.tx0 < xb && j.tx1 > xa && j.ty0 < ta + TILE_Y && j.ty1 > ta) keep |= 1u << b;
        }
        m.cand[(size_t)t * m.words + w] = keep;
        count += __popc(keep);
    }
    m.multi[t] = count > 1 ? 1 : 0;
}

// need(T) = OR of cand over the multi tiles within the blur reach of T;  wneed(T) = the same over
// reach + 1 tile, plus present(T) itself if there is any multi tile that close.
__global__ void __launch_bounds__(256)
seam_need_kernel(TileMaps m) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= m.tiles_x * m.tiles_y) return;
    const int tx = t % m.tiles_x, ty = t / m.tiles_x;
    const int x0 = max(tx - m.reach_x - 1, 0), x1 = min(tx + m.reach_x + 1, m.tiles_x - 1);
    const int y0 = max(ty - m.reach_y - 1, 0), y1 = min(ty + m.reach_y + 1, m.tiles_y - 1);
    bool zone = false;
    for (int y = y0; y <= y1 && !zone; ++y)
        for (int x = x0; x <= x1; ++x)
            if (m.multi[(size_t)y * m.tiles_x + x]) { zone = true; break; }
    for (int w = 0; w < m.words; ++w) {
        uint32_t near = 0u, wide = 0u;
        if (zone) {
            for (int y = y0; y <= y1; ++y)
                for (int x = x0; x <= x1; ++x) {
                    const size_t n = (size_t)y * m.tiles_x + x;
                    if (!m.multi[n]) continue;
                    const uint32_t bits = __ldg(m.cand + n * m.words + w);
                    wide |= bits;
                    if (x - tx <= m.reach_x && tx - x <= m.reach_x && y - ty <= m.reach_y && ty - y <= m.reach_y) near |= bits;
                }
            wide |= __ldg(m.present + (size_t)t * m.words + w);
        }
        m.need[(size_t)t * m.words + w] = near;
        m.wneed[(size_t)t * m.words + w] = wide;
    }
}

// ---- K1t: the tile warp --------------------------------------------------------------------
// One block per 64 x 32 mosaic tile, driven by the seam plan.
//  * A solo tile (one candidate) is that patch's warped pixels wherever it is valid, truncated to
//    uint8 like the blender's last line (stitcher.py:240-241), and zero elsewhere — exactly what
//    the multiband sum telescopes to.  Outside the seam zone such a tile never exists as float
//    RGBA, owner keys or coarse levels: source pixels in, mosaic bytes out.
//  * In the seam zone every patch with a `wneed` bit is warped to float RGBA + mask over the tile
//    (what reduce and collapse read).  The block visits the patches in order, so the owner
//    competition of stitcher.py:196-204 is a running maximum in registers — first maximum wins,
//    like np.argmax — and keys / covered are written once, without atomics or clearing.
// Rows of uint8 output are staged in shared memory at the byte phase of their destination, so
// that the mosaic is written with aligned 128-bit stores whatever W is.
constexpr int DT_PITCH = 3 * TILE_X + 16 + 16;          // 192 bytes of pixels + alignment phase (+ bank skew)

__device__ __forceinline__ uint8_t to_u8(float v) {      // (255 * clip(v, 0, 1)).astype(uint8)
    return (uint8_t)__float2int_rz(__fmul_rn(255.0f, fminf(fmaxf(v, 0.f), 1.f)));
}

// store `n` bytes staged at row[phase ...] to dst (dst & 15 == phase): bytes up to the first
// 16-byte boundary, aligned 128-bit words, bytes after the last one.  `lanes` threads cooperate.
__device__ __forceinline__ void store_row_bytes(uint8_t *dst, const uint8_t *row, int n, int lane, int lanes) {
    const int phase = (int)(reinterpret_cast<uintptr_t>(dst) & 15);
    const int head = min((16 - phase) & 15, n);
    const int body = (n - head) >> 4, tail = n - head - (body << 4);
    const uint8_t *src = row + phase;
    for (int i = lane; i < head; i += lanes) dst[i] = src[i];
    for (int i = lane; i < body; i += lanes)
        *reinterpret_cast<uint4 *>(dst + head + 16 * i) = *reinterpret_cast<const uint4 *>(src + head + 16 * i);
    for (int i = lane; i < tail; i += lanes) dst[head + 16 * body + i] = src[head + 16 * body + i];
}

constexpr int TW_ROWS = TILE_Y / 4;     // rows per thread (block = 64 x 4 threads)

// A warped pixel lies in [0, 1] (a convex combination of table values in [0, 1], the weights are
// exact and sum to one, rounding is monotonic): the blender's clip (stitcher.py:240) is a no-op.
__device__ __forceinline__ uint8_t to_u8_unit(float v) { return (uint8_t)__float2int_rz(__fmul_rn(255.0f, v)); }

// One patch over one tile.  FLOAT: write RGBA + mask to the patch and compete for the pixels;
// to_bytes (block-uniform): stage the truncated pixel for the mosaic (rows [y_lo, y_hi) of the window).
template <bool RGBX, bool FLOAT>
__device__ __forceinline__ void warp_tile_patch(const WarpJob &job, const float *lut, int tx0, int ty0, bool to_bytes,
                                                uint8_t (*rows)[DT_PITCH], const uint8_t *out, int W,
                                                int y_lo, int y_hi, float (*best_a)[TILE_X],
                                                int16_t (*best_p)[TILE_X], unsigned &valid_bits) {
    const int X = tx0 + threadIdx.x, c = X - job.x0;
    if ((unsigned)c >= (unsigned)job.pw) return;
    const ColumnTerms col = column_terms(job, c);
    const double *ray_rows = job.ray_y + job.row0;            // y component of the ray per patch row
    // byte phase of a staged row at its destination: (out + 3 (Y W + tx0)) & 15, in 32-bit arithmetic
    const unsigned phase0 = (unsigned)(reinterpret_cast<uintptr_t>(out) + 3u * (unsigned)tx0) & 15u;
    const unsigned phase_step = (3u * (unsigned)W) & 15u;
    constexpr int RP = FLOAT ? 2 : 4;       // rows per pass (the float path carries alpha and more addresses)
#pragma unroll 1
    for (int half = 0; half < TW_ROWS / RP; ++half) {
        // RP rows per pass, branch-free up to the stores: all coordinates (rows clamped into the
        // patch), then all gathers, then LUT + blend; only the commit looks at what is live
        TapPlan plan[RP];
        TapAlpha alpha[RP];
        uint32_t taps[RP][4];
#pragma unroll
        for (int k = 0; k < RP; ++k) {
            const int r = ty0 + (int)threadIdx.y + 4 * (RP * half + k) - job.y0;
            plan[k] = plan_taps<FLOAT>(job, col, __ldg(ray_rows + min(max(r, 0), job.ph - 1)), alpha[k]);
        }
#pragma unroll
        for (int k = 0; k < RP; ++k) {
            load_row_taps<RGBX>(job, plan[k].off00, plan[k].off01, taps[k][0], taps[k][1]);
            load_row_taps<RGBX>(job, plan[k].off10, plan[k].off11, taps[k][2], taps[k][3]);
        }
#pragma unroll
        for (int k = 0; k < RP; ++k) {
            const int slot = RP * half + k;
            const int ry = threadIdx.y + 4 * slot, Y = ty0 + ry, r = Y - job.y0;
            const float3 rgb = finish_rgb(lut, plan[k], taps[k]);
            const bool bad = plan[k].bad;
            if ((unsigned)r >= (unsigned)job.ph) continue;            // the tile row lies outside the patch
            if (FLOAT) {
                float4 o = make_float4(rgb.x, rgb.y, rgb.z, 0.0f);
                if (!bad) o.w = blend4(alpha[k].a00, alpha[k].a01, alpha[k].a10, alpha[k].a11, plan[k]);   // stitcher.py:317
                const size_t idx = (size_t)r * job.pw + c;
                st_stream(job.out + idx, o);
                job.invalid[idx] = bad ? 1 : 0;
                if (o.w > best_a[ry][threadIdx.x]) {                  // first maximum wins (patches come in order)
                    best_a[ry][threadIdx.x] = o.w;
                    best_p[ry][threadIdx.x] = (int16_t)job.patch;
                }
            }
            if (!bad) valid_bits |= 1u << slot;
            if (to_bytes && !bad && Y >= y_lo && Y < y_hi) {
                uint8_t *stage = rows[ry] + ((phase0 + (unsigned)Y * phase_step) & 15u) + 3 * threadIdx.x;
                stage[0] = to_u8_unit(rgb.x); stage[1] = to_u8_unit(rgb.y); stage[2] = to_u8_unit(rgb.z);
            }
        }
    }
}

#ifndef P360_TILE_BLOCKS
#define P360_TILE_BLOCKS 4      // resident blocks per SM the tile warp is compiled for (64 registers; B200, cfg4: 3.92 -> 3.69 ms vs 3)
#endif
template <bool RGBX>
__global__ void __launch_bounds__(256, P360_TILE_BLOCKS)
warp_tiles_kernel(int n_jobs, unsigned long long *__restrict__ keys,
                  uint8_t *__restrict__ covered, uint8_t *__restrict__ out, int OW, int y_begin, int y_end,
                  int x_begin, int x_end, int H, int W, int want_covered, TileMaps m) {
    __align__(16) __shared__ uint8_t rows[TILE_Y][DT_PITCH];
    __shared__ float lut[256];
    __shared__ float best_a[TILE_Y][TILE_X];             // running owner of every pixel: each thread only
    __shared__ int16_t best_p[TILE_Y][TILE_X];           // ever touches its own eight slots
    const int tid = threadIdx.y * TILE_X + threadIdx.x;
    const int tx0 = blockIdx.x * TILE_X, ty0 = m.row0 + (int)blockIdx.y * TILE_Y;
    const size_t tile = (size_t)blockIdx.y * m.tiles_x + blockIdx.x;
    const bool multi = __ldg(m.multi + tile) != 0;
    const uint32_t *cand = m.cand + tile * m.words, *wneed = m.wneed + tile * m.words;
    bool zone = false;
    int solo = -1;                                       // the candidate of a non-multi tile
    for (int w = 0; w < m.words; ++w) {
        zone |= __ldg(wneed + w) != 0u;
        const uint32_t c = __ldg(cand + w);
        if (c && solo < 0) solo = 32 * w + __ffs(c) - 1;
    }
    if (multi) solo = -1;
    // this block writes mosaic bytes (x_begin / x_end sit on tile edges)
    const bool bytes = !multi && ty0 < y_end && ty0 + TILE_Y > y_begin && tx0 >= x_begin && tx0 < x_end;
    if (!zone && !bytes) return;                         // (block-uniform)
    if (bytes) {                                         // zeros wherever the candidate has no valid pixel
        for (int i = tid; i < TILE_Y * DT_PITCH / 16; i += 256)
            reinterpret_cast<uint4 *>(&rows[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    unsigned valid_bits = 0u;
    if (zone) {
#pragma unroll
        for (int k = 0; k < TW_ROWS; ++k) {
            best_a[threadIdx.y + 4 * k][threadIdx.x] = 0.0f;
            best_p[threadIdx.y + 4 * k][threadIdx.x] = -1;
        }
    }
    const float *lut_of = nullptr;
    // the patches wanted as float here (ascending = patch order), then the solo candidate if it
    // was not among them
    bool solo_done = false;
    for (int w = 0; w <= m.words; ++w) {
        uint32_t bits;
        if (w < m.words) bits = zone ? __ldg(wneed + w) : 0u;
        else bits = (solo >= 0 && !solo_done && bytes) ? 1u : 0u;
        while (bits) {
            int id;
            bool as_float;
            if (w < m.words) { id = 32 * w + __ffs(bits) - 1; as_float = true; } else { id = solo; as_float = false; }
            bits &= bits - 1;
            if (id >= n_jobs) break;
            const WarpJob &job = c_warp_jobs[id];
            if (job.lut != lut_of) {                     // (block-uniform) all images without gains share one table
                __syncthreads();
                lut[tid] = __ldg(job.lut + tid);
                lut_of = job.lut;
                __syncthreads();
            }
            if (as_float)
                warp_tile_patch<RGBX, true>(job, lut, tx0, ty0, bytes && id == solo, rows, out, OW, y_begin, y_end,
                                            best_a, best_p, valid_bits);
            else
                warp_tile_patch<RGBX, false>(job, lut, tx0, ty0, true, rows, out, OW, y_begin, y_end,
                                             best_a, best_p, valid_bits);
            solo_done |= id == solo;
        }
    }
    const int X = tx0 + threadIdx.x;
    if (X < W && (zone || want_covered)) {
#pragma unroll
        for (int k = 0; k < TW_ROWS; ++k) {
            const int Y = ty0 + threadIdx.y + 4 * k;
            if (Y < 0 || Y >= H) continue;
            const size_t mi = (size_t)Y * W + X;
            if (zone) {
                const int p = best_p[threadIdx.y + 4 * k][threadIdx.x];
                keys[mi] = p < 0 ? 0ull : owner_key(best_a[threadIdx.y + 4 * k][threadIdx.x], p);
            }
            covered[mi] = (valid_bits >> k) & 1u;
        }
    }
    if (!bytes) return;
    __syncthreads();
    // 8 threads per row: aligned 128-bit stores of the staged bytes
    const int ry = tid >> 3, Y = ty0 + ry;
    if (Y >= y_begin && Y < y_end)
        store_row_bytes(out + ((size_t)Y * OW + tx0) * 3, rows[ry], 3 * min(TILE_X, W - tx0), tid & 7, 8);
}

// u8 x 3 -> u8 x 4 (RGBX): one aligned 32-bit word per source pixel, so that a bilinear tap of the
// warp is a single load.  Four pixels per thread: three 32-bit loads, one 128-bit store.
__global__ void __launch_bounds__(256)
pack_rgbx_kernel(const uint8_t *__restrict__ src, long long n_px, uint32_t *__restrict__ dst) {
    const long long g = (long long)blockIdx.x * 256 + threadIdx.x;     // group of 4 pixels
    const long long first = 4 * g;
    if (first >= n_px) return;
    if (first + 4 <= n_px) {
        const uint32_t *in = reinterpret_cast<const uint32_t *>(src) + 3 * g;
        const uint32_t a = __ldg(in), b = __ldg(in + 1), c = __ldg(in + 2);
        uint4 o;
        o.x = a & 0xffffffu;
        o.y = (a >> 24) | ((b & 0xffffu) << 8);
        o.z = (b >> 16) | ((c & 0xffu) << 16);
        o.w = c >> 8;
        *reinterpret_cast<uint4 *>(dst + first) = o;
        return;
    }
    for (long long i = first; i < n_px; ++i) {
        const uint8_t *p = src + 3 * i;
        dst[i] = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
    }
}


// The same for the pixels [c0, c1) of rows [r0, r1) of an h x w image (c0 % 4 == 0): what a column
// window of the compositor uploads of an image.  grid.y = row.
__global__ void __launch_bounds__(256)
pack_rgbx_rect_kernel(const uint8_t *__restrict__ src, int w, int r0, int c0, int c1, uint32_t *__restrict__ dst) {
    const int row = r0 + (int)blockIdx.y;
    const int first = c0 + 4 * ((int)blockIdx.x * 256 + (int)threadIdx.x);
    if (first >= c1) return;
    const size_t px = (size_t)row * w + first;                 // pixel index in the image
    if (first + 4 <= c1 && ((3 * px) & 3) == 0) {
        const uint32_t *in = reinterpret_cast<const uint32_t *>(src + 3 * px);
        const uint32_t a = __ldg(in), b = __ldg(in + 1), c = __ldg(in + 2);
        uint4 o;
        o.x = a & 0xffffffu;
        o.y = (a >> 24) | ((b & 0xffffu) << 8);
        o.z = (b >> 16) | ((c & 0xffu) << 16);
        o.w = c >> 8;
        if (((px & 3) == 0)) { *reinterpret_cast<uint4 *>(dst + px) = o; return; }
        dst[px] = o.x; dst[px + 1] = o.y; dst[px + 2] = o.z; dst[px + 3] = o.w;
        return;
    }
    for (int i = first; i < min(first + 4, c1); ++i) {
        const uint8_t *q = src + 3 * ((size_t)row * w + i);
        dst[(size_t)row * w + i] = (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16);
    }
}

// All images of a composite in one launch: grid.z = job, every job a rectangle of one image.
struct PackJob {                        // == p360_pack_job
    const uint8_t *src;
    uint32_t *dst;
    int h, w, r0, r1, c0, c1;
};
static_assert(sizeof(PackJob) == sizeof(p360_pack_job), "ABI struct mismatch");

__global__ void __launch_bounds__(256)
pack_rgbx_batch_kernel(const PackJob *__restrict__ jobs) {
    const PackJob j = jobs[blockIdx.z];
    const int row = j.r0 + (int)blockIdx.y;
    const int first = j.c0 + 4 * ((int)blockIdx.x * 256 + (int)threadIdx.x);
    if (row >= j.r1 || first >= j.c1) return;
    const size_t px = (size_t)row * j.w + first;
    if (first + 4 <= j.c1 && (px & 3) == 0) {
        const uint32_t *in = reinterpret_cast<const uint32_t *>(j.src + 3 * px);
        const uint32_t a = __ldg(in), b = __ldg(in + 1), c = __ldg(in + 2);
        uint4 o;
        o.x = a & 0xffffffu;
        o.y = (a >> 24) | ((b & 0xffffu) << 8);
        o.z = (b >> 16) | ((c & 0xffu) << 16);
        o.w = c >> 8;
        *reinterpret_cast<uint4 *>(j.dst + px) = o;
        return;
    }
    for (int i = first; i < min(first + 4, j.c1); ++i) {
        const uint8_t *q = j.src + 3 * ((size_t)row * j.w + i);
        j.dst[(size_t)row * j.w + i] = (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16);
    }
}

// ---- K0s: which source pixels does the plan read? ---------------------------------------------
// For every patch, the box (source pixels) around everything its tiles can sample: the tiles
// where it is warped to float (`wneed`) and the solo tiles it writes directly, each pushed through
// the same interval arithmetic as the alpha bounds (ray tables -> K R ray -> source position),
// grown by the bilinear footprint and the fixed-point slack.  Uploading just that rectangle of
// every image is exact: nothing else is ever loaded.  rects[8 * patch] = {u0, v0, u1, v1} (source
// pixels) and {x0, y0, x1, y1} (the buffer pixels of the tiles the patch is read for): half-open,
// clipped to the image / the patch; initialise to {MAX, MAX, MIN, MIN} twice; positions beyond the image
// fold back by BORDER_REFLECT; a patch partly behind the camera on a tile takes the whole image.
// One thread per tile.
// taps of positions in [a, b] along an axis of n pixels: floor(p) and floor(p) + 1, indices outside
// the image folded back by BORDER_REFLECT (p < 0 -> -p - 1, p >= n -> 2n - 1 - p).  [lo, hi) covers them.
__device__ __forceinline__ void axis_taps(double a, double b, int n, int &lo, int &hi) {
    if (!(a > -(double)n && b < 2.0 * n - 2.0)) { lo = 0; hi = n; return; }     // farther than one period (or NaN)
    lo = a < 0.0 ? 0 : (int)floor(a);
    hi = b > n - 2.0 ? n : (int)floor(b) + 2;
    if (a < 0.0) hi = max(hi, min(n, (int)ceil(-a) + 1));
    if (b > n - 2.0) lo = min(lo, max(0, 2 * n - 3 - (int)floor(b)));
}

__device__ __forceinline__ bool source_range(const WarpJob &j, Interval rx, Interval ry, Interval rz, int *box) {
    Interval p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const Interval x = scaled(j.kr[3 * k], rx), y = scaled(j.kr[3 * k + 1], ry), z = scaled(j.kr[3 * k + 2], rz);
        p[k] = Interval{x.lo + y.lo + z.lo, x.hi + y.hi + z.hi};
    }
    if (p[2].lo <= 1e-9) return false;                     // (partly) behind the camera: anything may be sampled
    const double u0 = fmin(fmin(p[0].lo / p[2].lo, p[0].lo / p[2].hi), fmin(p[0].hi / p[2].lo, p[0].hi / p[2].hi));
    const double u1 = fmax(fmax(p[0].lo / p[2].lo, p[0].lo / p[2].hi), fmax(p[0].hi / p[2].lo, p[0].hi / p[2].hi));
    const double v0 = fmin(fmin(p[1].lo / p[2].lo, p[1].lo / p[2].hi), fmin(p[1].hi / p[2].lo, p[1].hi / p[2].hi));
    const double v1 = fmax(fmax(p[1].lo / p[2].lo, p[1].lo / p[2].hi), fmax(p[1].hi / p[2].lo, p[1].hi / p[2].hi));
    const double slack = 1.0 / 32 + 1e-3 + 1e-6 * (fabs(u0) + fabs(u1) + fabs(v0) + fabs(v1));   // fixed point + float32 quotient
    axis_taps(u0 + 0.5 * j.w - slack, u1 + 0.5 * j.w + slack, j.w, box[0], box[2]);
    axis_taps(v0 + 0.5 * j.h - slack, v1 + 0.5 * j.h + slack, j.h, box[1], box[3]);
    return true;
}

__global__ void __launch_bounds__(128)
source_rects_kernel(const WarpJob *__restrict__ jobs, int n_jobs, int H, int W, int abs_row0, int mosaic_h,
                    TileMaps m, int *__restrict__ rects) {
    const int t = blockIdx.x * 128 + threadIdx.x;
    if (t >= m.tiles_x * m.tiles_y) return;
    const int tx = t % m.tiles_x, ty = t / m.tiles_x;
    const int xa = tx * TILE_X, xb = min(xa + TILE_X, W);
    const int ta = m.row0 + ty * TILE_Y;
    const int ya = max(ta, 0), yb = min(ta + TILE_Y, H);   // only pixels of the buffer are ever warped
    if (yb <= ya) return;
    const bool multi = m.multi[t] != 0;
    const WarpJob &j0 = jobs[0];
    const int dc = j0.col0 - j0.x0, dr = j0.row0 - j0.y0;
    Interval rx{0.0, 0.0}, ry{0.0, 0.0}, rz{0.0, 0.0};
    bool have = false;
    bool solo_seen = false;
    for (int w = 0; w < m.words; ++w) {
        uint32_t bits = __ldg(m.wneed + (size_t)t * m.words + w);
        if (!multi && !solo_seen) {                        // the solo candidate: the first cand bit
            const uint32_t c = __ldg(m.cand + (size_t)t * m.words + w);
            if (c) { bits |= c & (0u - c); solo_seen = true; }
        }
        while (bits) {
            const int k = 32 * w + __ffs(bits) - 1;
            bits &= bits - 1;
            if (k >= n_jobs) break;
            const WarpJob &j = jobs[k];
            // the pixels of the tile the (cropped) patch covers
            const int px0 = max(xa, j.x0), px1 = min(xb, j.x0 + j.pw), py0 = max(ya, j.y0), py1 = min(yb, j.y0 + j.ph);
            if (px1 <= px0 || py1 <= py0) continue;
            if (!have) {
                rx = table_range(j0.ray_x, xa + dc, xb + dc); rz = table_range(j0.ray_z, xa + dc, xb + dc);
                ry = table_range(j0.ray_y, ya + dr, yb + dr);
                have = true;
            }
            int box[4];
            if (!source_range(j, rx, ry, rz, box)) { box[0] = 0; box[1] = 0; box[2] = j.w; box[3] = j.h; }
            atomicMin(rects + 8 * k, box[0]); atomicMin(rects + 8 * k + 1, box[1]);
            atomicMax(rects + 8 * k + 2, box[2]); atomicMax(rects + 8 * k + 3, box[3]);
            atomicMin(rects + 8 * k + 4, px0); atomicMin(rects + 8 * k + 5, py0);      // where in the buffer it is read
            atomicMax(rects + 8 * k + 6, px1); atomicMax(rects + 8 * k + 7, py1);
        }
    }
}

}  // namespace p360

extern "C" int p360_pack_rgbx(const uint8_t *src_rgb, int h, int w, uint8_t *dst_rgbx, void *stream) {
    using namespace p360;
    const char *where = "p360_pack_rgbx";
    P360_REQUIRE(src_rgb && dst_rgbx && h > 0 && w > 0, where);
    P360_REQUIRE((reinterpret_cast<uintptr_t>(src_rgb) & 3) == 0 && aligned16(dst_rgbx), where);
    const long long n = (long long)h * w;
    pack_rgbx_kernel<<<cdiv((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(src_rgb, n, reinterpret_cast<uint32_t *>(dst_rgbx));
    return check_launch(where);
}

extern "C" int p360_pack_rgbx_rect(const uint8_t *src_rgb, int h, int w, int r0, int r1, int c0, int c1,
                                   uint8_t *dst_rgbx, void *stream) {
    using namespace p360;
    const char *where = "p360_pack_rgbx_rect";
    P360_REQUIRE(src_rgb && dst_rgbx && h > 0 && w > 0, where);
    P360_REQUIRE(0 <= r0 && r0 <= r1 && r1 <= h && 0 <= c0 && c0 <= c1 && c1 <= w && c0 % 4 == 0 && r1 - r0 <= 65535, where);
    P360_REQUIRE((reinterpret_cast<uintptr_t>(src_rgb) & 3) == 0 && aligned16(dst_rgbx), where);
    if (r1 == r0 || c1 == c0) return 0;
    dim3 grid(cdiv((c1 - c0 + 3) / 4, 256), r1 - r0);
    pack_rgbx_rect_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src_rgb, w, r0, c0, c1, reinterpret_cast<uint32_t *>(dst_rgbx));
    return check_launch(where);
}

extern "C" int p360_pack_rgbx_batch(const p360_pack_job *jobs_dev, int n_jobs, int max_rows, int max_cols, void *stream) {
    using namespace p360;
    const char *where = "p360_pack_rgbx_batch";
    P360_REQUIRE(jobs_dev && n_jobs >= 0 && n_jobs <= 65535 && max_rows >= 0 && max_rows <= 65535 && max_cols >= 0, where);
    if (n_jobs == 0 || max_rows == 0 || max_cols == 0) return 0;
    dim3 grid(cdiv((max_cols + 3) / 4, 256), max_rows, n_jobs);
    pack_rgbx_batch_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const PackJob *>(jobs_dev));
    return check_launch(where);
}

static int seam_maps_ok(const p360::TileMaps &m, int n_jobs, int H, int W, const char *where) {
    using namespace p360;
    P360_REQUIRE(m.present && m.cand && m.need && m.wneed && m.multi, where);
    P360_REQUIRE(m.words == (n_jobs + 31) / 32 && m.row0 <= 0 && m.row0 > -TILE_Y, where);
    P360_REQUIRE(m.tiles_x == (int)cdiv(W, TILE_X) && m.tiles_y == (int)cdiv(H - m.row0, TILE_Y), where);
    P360_REQUIRE(m.reach_x >= 0 && m.reach_y >= 0, where);
    return 0;
}

extern "C" int p360_seam_plan_build(const p360_warp_job *jobs_dev, int n_jobs, p360_band_patch *patches_dev,
                                    int H, int W, int abs_row0, int mosaic_h,
                                    const p360_tile_maps *maps_host, void *stream) {
    using namespace p360;
    const char *where = "p360_seam_plan_build";
    P360_REQUIRE(jobs_dev && maps_host && n_jobs > 0 && n_jobs <= 1024 && H > 0 && W > 0, where);
    P360_REQUIRE(abs_row0 >= 0 && mosaic_h >= abs_row0 + H, where);
    TileMaps m;
    memcpy(&m, maps_host, sizeof(m));
    if (int e = seam_maps_ok(m, n_jobs, H, W, where)) return e;
    cudaStream_t s = (cudaStream_t)stream;
    const long long tiles = (long long)m.tiles_x * m.tiles_y;
    auto jobs = reinterpret_cast<const WarpJob *>(jobs_dev);
    seam_candidates_kernel<<<cdiv(tiles, 128), 128, 0, s>>>(jobs, n_jobs, reinterpret_cast<BandPatch *>(patches_dev),
                                                            H, W, abs_row0, mosaic_h, m);
    if (int e = check_launch(where)) return e;
    seam_cand_kernel<<<cdiv(tiles, 256), 256, 0, s>>>(jobs, n_jobs, W, m);
    if (int e = check_launch(where)) return e;
    seam_need_kernel<<<cdiv(tiles, 256), 256, 0, s>>>(m);
    return check_launch(where);
}

extern "C" int p360_source_rects(const p360_warp_job *jobs_dev, int n_jobs, int H, int W, int abs_row0, int mosaic_h,
                                 const p360_tile_maps *maps_host, int32_t *rects_dev, void *stream) {
    using namespace p360;
    const char *where = "p360_source_rects";
    P360_REQUIRE(jobs_dev && maps_host && rects_dev && n_jobs > 0 && n_jobs <= 1024 && H > 0 && W > 0, where);
    TileMaps m;
    memcpy(&m, maps_host, sizeof(m));
    if (int e = seam_maps_ok(m, n_jobs, H, W, where)) return e;
    const long long tiles = (long long)m.tiles_x * m.tiles_y;
    source_rects_kernel<<<cdiv(tiles, 128), 128, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const WarpJob *>(jobs_dev), n_jobs, H, W, abs_row0, mosaic_h, m, rects_dev);
    return check_launch(where);
}

extern "C" int p360_warp_tiles(const p360_warp_job *jobs_host, const p360_warp_job *jobs_dev, int n_jobs, uint64_t *owner_keys,
                               uint8_t *covered, uint8_t *out_u8, int out_pitch, int y_begin, int y_end, int x_begin,
                               int x_end, int H, int W, int want_covered, const p360_tile_maps *maps_host, void *stream) {
    using namespace p360;
    const char *where = "p360_warp_tiles";
    P360_REQUIRE(jobs_host && maps_host && owner_keys && covered && out_u8, where);
    P360_REQUIRE(n_jobs > 0 && n_jobs <= TILE_JOBS_MAX && H > 0 && W > 0 && y_begin >= 0 && y_end <= H, where);
    P360_REQUIRE(x_begin >= 0 && x_begin <= x_end && x_end <= W && x_begin % TILE_X == 0 &&
                 (x_end % TILE_X == 0 || x_end == W), where);
    P360_REQUIRE(out_pitch == 0 || out_pitch >= x_end, where);
    const int OW = out_pitch ? out_pitch : W;
    TileMaps m;
    memcpy(&m, maps_host, sizeof(m));
    if (int e = seam_maps_ok(m, n_jobs, H, W, where)) return e;
    dim3 grid(m.tiles_x, m.tiles_y), block(TILE_X, 4);
    P360_REQUIRE(grid.y <= 65535, where);
    cudaStream_t s = (cudaStream_t)stream;
    // Stream-ordered: waits for the previous launch that still reads the table.  From the DEVICE copy
    // of the table if there is one: a copy from pageable host memory makes the runtime synchronise
    // the stream first, which would put the host in lockstep with the GPU.
    if (jobs_dev != nullptr)
        P360_CUDA(cudaMemcpyToSymbolAsync(c_warp_jobs, jobs_dev, sizeof(WarpJob) * n_jobs, 0, cudaMemcpyDeviceToDevice, s), where);
    else
        P360_CUDA(cudaMemcpyToSymbolAsync(c_warp_jobs, jobs_host, sizeof(WarpJob) * n_jobs, 0, cudaMemcpyHostToDevice, s), where);
    bool rgbx = true;
    for (int k = 0; k < n_jobs; ++k) {
        const p360_warp_job &j = jobs_host[k];
        P360_REQUIRE(j.src && j.lut && j.hat_y && j.hat_x && j.ray_x && j.ray_z && j.ray_y && j.out && j.invalid, where);
        P360_REQUIRE(j.h > 0 && j.w > 0 && j.h <= 32767 && j.w <= 32767 && j.pw > 0 && j.ph > 0 && aligned16(j.out), where);
        P360_REQUIRE(j.c == 3 || (j.c == 4 && (reinterpret_cast<uintptr_t>(j.src) & 3) == 0), where);
        P360_REQUIRE(j.x0 >= 0 && j.y0 >= 0 && j.x0 + j.pw <= W && j.y0 + j.ph <= H && j.patch == k, where);
        rgbx = rgbx && j.c == 4;
    }
    auto keys = reinterpret_cast<unsigned long long *>(owner_keys);
    if (rgbx)
        warp_tiles_kernel<true><<<grid, block, 0, s>>>(n_jobs, keys, covered, out_u8, OW, y_begin, y_end, x_begin, x_end,
                                                       H, W, want_covered, m);
    else
        warp_tiles_kernel<false><<<grid, block, 0, s>>>(n_jobs, keys, covered, out_u8, OW, y_begin, y_end, x_begin, x_end,
                                                        H, W, want_covered, m);
    return check_launch(where);
}

extern "C" int p360_warp_batch(const p360_warp_job *jobs_host, const p360_warp_job *jobs_dev, int n_jobs,
                               uint64_t *owner_keys, uint8_t *covered, int W, void *stream) {
    using namespace p360;
    const char *where = "p360_warp_batch";
    P360_REQUIRE(jobs_host && n_jobs >= 0, where);
    P360_REQUIRE(owner_keys == nullptr || (covered != nullptr && W > 0), where);
    cudaStream_t s = (cudaStream_t)stream;
    for (int first = 0; first < n_jobs; first += WARP_JOBS_PER_LAUNCH) {
        const int count = n_jobs - first < WARP_JOBS_PER_LAUNCH ? n_jobs - first : WARP_JOBS_PER_LAUNCH;
        int max_pw = 0, max_ph = 0;
        bool rgbx = true;
        for (int k = first; k < first + count; ++k) {
            const p360_warp_job &j = jobs_host[k];
            P360_REQUIRE(j.src && j.lut && j.hat_y && j.hat_x && j.ray_x && j.ray_z && j.ray_y && j.out && j.invalid, where);
            P360_REQUIRE(j.h > 0 && j.w > 0 && j.h <= 32767 && j.w <= 32767 && j.pw >= 0 && j.ph >= 0, where);
            P360_REQUIRE(j.c == 3 || (j.c == 4 && (reinterpret_cast<uintptr_t>(j.src) & 3) == 0), where);
            P360_REQUIRE(aligned16(j.out), where);
            P360_REQUIRE(owner_keys == nullptr || (j.x0 >= 0 && j.y0 >= 0 && j.x0 + j.pw <= W), where);
            rgbx = rgbx && j.c == 4;
            max_pw = j.pw > max_pw ? j.pw : max_pw;
            max_ph = j.ph > max_ph ? j.ph : max_ph;
        }
        if (max_pw == 0 || max_ph == 0) continue;
        // stream-ordered: waits for the previous launch that still reads the table (see p360_warp_tiles)
        if (jobs_dev != nullptr)
            P360_CUDA(cudaMemcpyToSymbolAsync(c_warp_jobs, jobs_dev + first, sizeof(WarpJob) * count, 0,
                                              cudaMemcpyDeviceToDevice, s), where);
        else
            P360_CUDA(cudaMemcpyToSymbolAsync(c_warp_jobs, jobs_host + first, sizeof(WarpJob) * count, 0,
                                              cudaMemcpyHostToDevice, s), where);
        dim3 block(WARP_BX, WARP_BY), grid(cdiv(max_pw, WARP_BX), cdiv(max_ph, WARP_BY * WARP_ROWS), count);
        P360_REQUIRE(grid.y <= 65535, where);
        auto keys = reinterpret_cast<unsigned long long *>(owner_keys);
        if (rgbx && keys) warp_batch_kernel<true, true><<<grid, block, 0, s>>>(keys, covered, W);
        else if (rgbx) warp_batch_kernel<true, false><<<grid, block, 0, s>>>(keys, covered, W);
        else if (keys) warp_batch_kernel<false, true><<<grid, block, 0, s>>>(keys, covered, W);
        else warp_batch_kernel<false, false><<<grid, block, 0, s>>>(keys, covered, W);
        if (int e = check_launch(where)) return e;
    }
    return 0;
}
