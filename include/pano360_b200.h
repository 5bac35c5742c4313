/*
 * pano360_b200 — C ABI of the B200-native compositing path.
 *
 * The reference (Banus/pano360) is pure Python and has no FFI of its own; the
 * entry points below are what a binding for its compositing hot path
 * (stitcher.py:274-327 `stitch`, :160-241 blenders, :24-66 exposure gains)
 * binds instead of the NumPy/OpenCV calls at the cited lines.  They are
 * called from Python through ctypes (pano360_b200/_lib.py); INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host;
 *  - images are row-major, pixel-interleaved; "rgba" is float32 x 4 per pixel
 *    (channel order = the caller's, the reference feeds BGR), 16-byte aligned;
 *  - a patch is the ph x pw bounding box of one warped image placed at
 *    (x0, y0) in a mosaic of width W (stitcher.py:318 `irange`);
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - return value: 0 on success, a cudaError_t (>0) or P360_EINVAL (<0)
 *    otherwise; p360_last_error() returns the message for the calling thread.
 *    Nothing aborts the process and there is no CPU fallback.
 */
#ifndef PANO360_B200_H
#define PANO360_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P360_VERSION 100          /* 0.1.0 */
#define P360_EINVAL  (-22)
#define P360_MAX_KSIZE 129        /* widest separable Gaussian supported */
#define P360_MAX_LEVELS 8         /* most bands of the multiband blender */

int  p360_version(void);
/* Copies the calling thread's last error message into buf (NUL-terminated). */
int  p360_last_error(char *buf, int n);
/* cudaGetDeviceProperties subset: {sm_count, cc_major, cc_minor, l2_bytes}. */
int  p360_device_info(int device, int32_t out_host[4]);

/* ---- K1: inverse projection + 1/32-px bilinear remap + validity mask ------
 * Replaces stitcher.py:257-263 (_add_weights) + :300-317 (coordinates, mask,
 * cv2.remap INTER_LINEAR/BORDER_REFLECT, alpha *= ~mask), fused; optionally
 * also the weights tensor / argmax of :196-204 and `allmask` of :233-234.
 * One launch warps every patch of a composite (up to 128 per launch; the job
 * records are staged in constant memory): `jobs_host` is a HOST array whose
 * pointer members are device pointers.
 *   src         u8, h x w x c bytes per pixel: c = 3 (as uploaded) or c = 4 (RGBX:
 *               p360_pack_rgbx or a 4-channel upload, the 4th byte is ignored);
 *               alpha = float32(hat_y * hat_x) is evaluated at the four taps
 *   lut         256 float32: value of a u8 sample (u8/255, optionally
 *               gain-scaled and clipped, stitcher.py:65-66)
 *   hat_y/hat_x float64 tables of `_hat(h)` / `_hat(w)` (stitcher.py:251-254)
 *   ray_x/ray_z float64 per MOSAIC column: x and z components of proj2hom
 *   ray_y       float64 per MOSAIC row: y component of proj2hom (the projection
 *               is separable: stitcher.py:84-87, :101-104); the patch origin
 *               sits at (col0, row0) in these tables
 *   kr          K*R row-major float64 (bundle_adj.py:31-33); p = kr . ray is
 *               evaluated in float64 per pixel, then cast (stitcher.py:306)
 *   out         ph x pw x 4 float32; invalid: ph x pw u8 (1 = masked)
 *   x0, y0      position of the patch in the (window) mosaic of width W
 *   patch       id of the patch in the owner keys (its list position)
 * owner_keys (H x W uint64, zero-initialised) / covered (H x W u8, zero-
 * initialised) may both be NULL to skip the owner-map competition.
 * Owner key = float_bits(alpha) << 32 | (0xFFFFFFFF - patch), combined with
 * atomicMax: largest alpha wins, ties go to the smallest patch id (np.argmax's
 * first maximum), key 0 = no owner; only alpha > 0 competes.
 */
typedef struct p360_warp_job {
    const uint8_t *src;
    const float *lut;
    const double *hat_y, *hat_x;
    const double *ray_x, *ray_z, *ray_y;
    float *out;
    uint8_t *invalid;
    double kr[9];
    int32_t h, w, c;
    int32_t pw, ph;
    int32_t x0, y0;
    int32_t col0, row0;
    int32_t patch;
    float half_w, half_h;         /* float32(w / 2.0), float32(h / 2.0)                   */
    float max_x, max_y;           /* float32(w - 1), float32(h - 1)                       */
    float inv_2w, inv_2h;         /* 1 / (2 w), 1 / (2 h)                                 */
    int32_t ty0, ty1;             /* rows / columns of the image's TRUE box (of this column run, for a */
    int32_t tx0, tx1;             /* seam-split image) in window coordinates: x0 / y0 / pw / ph may be   */
                                  /* cropped to a window; the seam plan must not depend on the cut       */
} p360_warp_job;

/* u8 x 3 -> u8 x 4 (RGBX): the source layout in which a bilinear tap is one aligned 32-bit load.
 * _rect: only pixels [c0, c1) of rows [r0, r1) of the h x w image (c0 % 4 == 0) — what was uploaded
 * of an image when only part of it is read (p360_source_rects). */
int p360_pack_rgbx(const uint8_t *src_rgb, int h, int w, uint8_t *dst_rgbx, void *stream);
int p360_pack_rgbx_rect(const uint8_t *src_rgb, int h, int w, int r0, int r1, int c0, int c1,
                        uint8_t *dst_rgbx, void *stream);
/* One launch for all the images of a composite.  jobs_dev: DEVICE array; every job converts pixels
 * [c0, c1) of rows [r0, r1) of an h x w image (c0 % 4 == 0; w % 4 == 0 for the vector path, any w
 * is correct); max_rows / max_cols: the largest r1 - r0 / c1 - c0 among the jobs. */
typedef struct p360_pack_job {
    const uint8_t *src;
    uint8_t *dst;
    int32_t h, w, r0, r1, c0, c1;
} p360_pack_job;
int p360_pack_rgbx_batch(const p360_pack_job *jobs_dev, int n_jobs, int max_rows, int max_cols, void *stream);
/* Rectangle copy (cudaMemcpy2DAsync, direction inferred): `rows` runs of width_bytes bytes, pitches
 * in bytes; device, peer-device and page-locked host addresses alike.  Uploads of image
 * sub-rectangles, downloads / NVLink pushes of mosaic column windows. */
int p360_copy_rect(void *dst, int64_t dst_pitch, const void *src, int64_t src_pitch,
                   int64_t width_bytes, int64_t rows, void *stream);
struct p360_tile_maps;
struct p360_band_patch;
/* jobs_dev (optional): a DEVICE copy of the same table — the constant-memory staging then is a
 * device-to-device copy and the call never synchronises the stream (a copy from pageable host
 * memory does). */
int p360_warp_batch(const p360_warp_job *jobs_host, const p360_warp_job *jobs_dev, int n_jobs,
                    uint64_t *owner_keys, uint8_t *covered, int W, void *stream);

/* ---- K0 + K1d: the seam plan and the direct tiles (multiband, >= 2 bands) -----------------
 * Ownership (stitcher.py:196-204) is arg-max of alpha = hat_y(v) * hat_x(u): a function of the
 * geometry alone.  p360_seam_plan_build bounds alpha of every patch on every 64 x 32 mosaic tile
 * by interval arithmetic (ray tables -> K R ray -> source position -> alpha) and fills the tile
 * bitmaps of p360_tile_maps BEFORE anything is sampled:
 *   present  patches no other patch dominates on the whole tile: every true owner is among them
 *            (also grows p360_band_patch.own of patches_dev, may be NULL)
 *   cand     `present` within reach_x / reach_y tiles, restricted to patches whose box meets the
 *            tile: the only patches with non-zero blend weights there.  multi = more than one.
 *            A tile with ONE candidate is that patch's warped pixels wherever it is valid (the
 *            sum over the bands of stitcher.py:224-238 telescopes) and zero elsewhere
 *   need     candidates of the multi tiles within reach: where coarse levels are consumed
 *   wneed    candidates of the multi tiles within reach + 1 tile, plus `present` of every tile
 *            that close to a multi tile: where float pixels and owner keys are consumed
 * Tiles are evaluated on their full extent against the TRUE boxes (ty0 / ty1) in absolute rows
 * (buffer row 0 = mosaic row abs_row0, mosaic height mosaic_h), so that a row window plans every
 * tile that lies with its reach inside the window exactly like the whole mosaic does.
 * jobs_dev = DEVICE copy of the job table (patch = position), at most 1024 jobs.
 *
 * p360_warp_tiles (jobs_host: the same table in HOST memory, at most 256 jobs: it is staged in
 * constant memory) then runs one block per tile (K1t).  A non-multi tile of rows [y_begin, y_end)
 * is written to the uint8 mosaic straight from the source images (rows staged in shared memory,
 * stored as aligned 128-bit words).  In every tile with `wneed` bits those patches are warped
 * to float RGBA + mask (job.out / job.invalid, tile by tile: pixels outside such tiles stay
 * unwritten and are never read) and compete for the pixels in patch order in registers:
 * owner_keys / covered are written once per pixel, without atomics, and need no initialising.
 * If want_covered, covered also records the valid mask of the other tiles (stitcher.py:266-271).
 * p360_multiband_collapse with the same record writes only the multi tiles.
 * Mosaic bytes are produced for rows [y_begin, y_end) x columns [x_begin, x_end) of the buffer
 * (x_begin % 64 == 0; x_end % 64 == 0 or x_end == W): a window of the mosaic, computed in a buffer
 * that also holds the halo the blurs need.  out_u8 is addressed as out_u8[(y * out_pitch + x) * 3]
 * for buffer pixel (x, y) — out_pitch in pixels, 0 = W — so that a window can be written straight
 * into its place in a larger image: the whole mosaic, or rank 0's mosaic mapped into this rank's
 * address space over NVLink (the strip gather of the multi-GPU path then costs no extra pass: the
 * tile warp and the collapse store their bytes in the peer's memory as they produce them; pass
 * out_u8 = that image's address of buffer pixel (0, 0)).  The buffer's column 0 must sit on a multiple of 64 of
 * the whole mosaic and its width must be a multiple of 64 unless it ends at the mosaic's right
 * edge, so that its tiles are tiles of the whole mosaic.
 *
 * p360_source_rects (K0s): after the plan, per patch the rectangle {u0, v0, u1, v1} (source pixels,
 * half-open) that covers every tap p360_warp_tiles can load for it — the same interval arithmetic,
 * over the tiles where the patch is wanted as float or is the solo candidate, BORDER_REFLECT folded
 * in — followed by {x0, y0, x1, y1}: the box (buffer pixels) around those tiles.  rects_dev: 8 int32
 * per job, initialised to {INT_MAX, INT_MAX, INT_MIN, INT_MIN} twice.  Pixels outside the first
 * rectangle are never read: only it has to be uploaded; a window of the mosaic that does not meet
 * the second one does not need the image at all.
 */
int p360_seam_plan_build(const p360_warp_job *jobs_dev, int n_jobs, struct p360_band_patch *patches_dev,
                         int H, int W, int abs_row0, int mosaic_h,
                         const struct p360_tile_maps *maps_host, void *stream);
int p360_warp_tiles(const p360_warp_job *jobs_host, const p360_warp_job *jobs_dev, int n_jobs, uint64_t *owner_keys,
                    uint8_t *covered, uint8_t *out_u8, int out_pitch, int y_begin, int y_end, int x_begin, int x_end,
                    int H, int W, int want_covered, const struct p360_tile_maps *maps_host, void *stream);
int p360_source_rects(const p360_warp_job *jobs_dev, int n_jobs, int H, int W, int abs_row0, int mosaic_h,
                      const struct p360_tile_maps *maps_host, int32_t *rects_dev, void *stream);

/* ---- K2: owner map for externally supplied patches (stitcher.py:196-208) ---
 * p360_owner_update: the same competition for one already-warped patch (the
 * blender API receives NumPy patches); also ORs `!invalid` into covered.
 * p360_owner_decode: keys -> int32 owner index (-1 = none), for inspection.
 */
int p360_owner_update(const float *rgba, const uint8_t *invalid, int pw, int ph,
                      int x0, int y0, int idx, uint64_t *owner_keys,
                      uint8_t *covered, int W, void *stream);
int p360_owner_decode(const uint64_t *owner_keys, int32_t *owner, int64_t n_pixels, void *stream);

/* ---- K3: cv2.GaussianBlur(rgba, (0,0), sigma) (stitcher.py:226) -----------
 * Separable float32 convolution with BORDER_REFLECT_101 at the image edges.
 * taps: ksize float32 (host computes cv2.getGaussianKernel semantics).
 * tmp: scratch of the same size as in/out.  in may alias neither out nor tmp.
 * Batched form: tap sets live in P360_MAX_LEVELS constant-memory slots
 * (p360_blur_set_taps, stream-ordered); every job names its slot.
 */
typedef struct p360_blur_job {
    const float *in;
    float *out;
    float *tmp;
    int32_t w, h, slot;
    int32_t shift;                /* log2 of the coarse factor of this image (patch != NULL)  */
    const struct p360_band_patch *patch;   /* DEVICE record of the patch this image belongs to, */
                                  /* or NULL: blocks nobody reads are skipped (seam-band maps  */
    int32_t pad, grow;            /* if given, else farther than `grow` px from patch->own)    */
} p360_blur_job;

int p360_gauss_blur(const float *in_rgba, float *out_rgba, float *tmp_rgba,
                    int pw, int ph, const float *taps_host, int ksize,
                    void *stream);
int p360_blur_set_taps(int slot, const float *taps_host, int ksize, void *stream);
struct p360_tile_maps;
int p360_gauss_blur_batch(const p360_blur_job *jobs, int n_jobs, int max_w, int max_h,
                          const struct p360_tile_maps *maps_host, void *stream);

/* ---- reduced-resolution band pipeline + output-stationary blenders ----------
 * The blurs of stitcher.py:226 are evaluated on coarse grids (f = 2 for level
 * 0, f = 4 above) of the BORDER_REFLECT_101 extension of the patch by `pad`
 * full-resolution pixels (pad % 4 == 0):
 *   p360_pyramid_dims          -> {w2, h2, w4, h4}: sizes of the coarse images
 *   p360_pyramid_reduce_batch  area-reduce every patch (alpha := owner == index
 *                              when owner_keys != NULL, stitcher.py:207-208)
 *                              into its d2 (f = 2) and d4 (f = 4)
 *   p360_multiband_collapse    for every mosaic pixel, over the patches covering
 *                              it in list order: expand the blurred coarse levels
 *                              bilinearly, form bands and weights
 *                              (stitcher.py:224-232), normalise per level, sum,
 *                              clamp, truncate to uint8 (stitcher.py:236-241)
 *   p360_linear_collapse       stitcher.py:171-183 in the same gather form
 *   p360_paste_collapse        stitcher.py:160-168 in the same gather form
 * Nothing mosaic-sized is accumulated in HBM.  `patches` is a DEVICE array.
 * The collapse kernels produce rows [y_begin, y_end) of the output buffer (0 and
 * H for the whole mosaic; the multiband one also takes columns [x_begin, x_end),
 * x_begin % 64 == 0) so that callers can overlap the download or the NVLink
 * send of finished row bands with the computation of the next ones.  row_origin
 * is the absolute mosaic row of buffer row 0 (0 unless the buffer is a strip):
 * work tiles are anchored at absolute rows, which makes the result independent
 * of how the mosaic is cut into strips and bands.
 */
typedef struct p360_band_patch {
    const float *rgba;                        /* full-res patch                         */
    const uint8_t *invalid;                   /* ph x pw mask (linear / paste)          */
    float *d2, *d4;                           /* reduce outputs                         */
    const float *low[P360_MAX_LEVELS - 1];    /* blurred coarse image of level l        */
    int32_t x0, y0, pw, ph;                   /* box in (window) mosaic pixels          */
    int32_t w4, h4;                           /* f = 4 grid size (f = 2 grid: twice)    */
    int32_t pad;                              /* extension in full-res pixels           */
    int32_t index;                            /* id of this patch in the owner keys     */
    int32_t own[4];                           /* box around the owned pixels (patch px), */
                                              /* filled on the device by p360_owned_boxes; */
                                              /* initialise to {MAX, MAX, MIN, MIN}       */
} p360_band_patch;

 /*  p360_owned_boxes        per patch, the (tile-granular) box around the pixels it owns,
 *                           from the owner keys: weights vanish beyond the blur reach of
 *                           that box, so reduce / blur / collapse skip everything farther
 *                           away (exact: skipped values only ever meet zero weights)      */
int p360_owned_boxes(const uint64_t *owner_keys, p360_band_patch *patches, int n_patches,
                     int H, int W, void *stream);

/* Seam-band maps: bitmaps over the 64 x 32 mosaic tiles, one bit per patch, `words` =
 * ceil(n_patches / 32) uint32 per tile, tile rows anchored at absolute mosaic rows
 * (row0 = -((row_origin mod 32 + 32) mod 32), tiles_y = ceil((H - row0) / 32),
 * tiles_x = ceil(W / 64)).  The weights of stitcher.py:207-232 are non-zero only within
 * the blur reach of an owner seam; everywhere else the multiband sum telescopes to the
 * owner's warped pixel.  p360_tile_maps_build fills, from the owner keys:
 *   present  patches that own a pixel of the tile (also grows p360_band_patch.own)
 *   cand     patches that own a pixel within reach_x / reach_y tiles: the candidates for
 *            non-zero weight in the tile;  multi = more than one candidate, or one and a
 *            valid pixel nobody owns (alpha == 0): the tile is not just its owner's pixels
 *   need     candidates of the multi tiles within reach: where coarse levels are consumed
 * reduce / blur run only the blocks whose tiles carry the patch's `need` bit (a scan compacts
 * them into `work`, a persistent grid consumes the list: an empty block of a dense grid costs
 * as much as a small busy one), the collapse takes its patch list from `cand`.  Results are identical with and without maps (maps_host == NULL).
 */
typedef struct p360_tile_maps {
    uint32_t *present, *cand, *need;          /* DEVICE [tiles_y][tiles_x][words]       */
    uint8_t *multi;                           /* DEVICE [tiles_y][tiles_x]              */
    uint32_t *work;                           /* DEVICE scratch, 2 * work_cap uint32: the list  */
    int32_t *work_count;                      /* of blocks a reduce / blur pass has to run; work_count: */
                                              /* 4 int32 — [0] its length (live), [1..3] blocks run by   */
                                              /* the last reduce / horizontal / vertical blur pass;      */
                                              /* work_cap >= blocks of the largest grid                  */
    uint32_t *wneed;                          /* seam plan only (else NULL): where float pixels and  */
                                              /* owner keys are wanted, see p360_seam_plan_build     */
    int32_t tiles_x, tiles_y, words;
    int32_t row0;
    int32_t reach_x, reach_y;                 /* ceil(pad / 64), ceil(pad / 32)          */
    int32_t work_cap;
    int32_t h_rows;                           /* 4: horizontal blur in 64-cell segments x 16 rows */
                                              /* per block instead of 256 x 4 (0 / 1)             */
} p360_tile_maps;
int p360_tile_maps_build(const uint64_t *owner_keys, const uint8_t *covered,
                         p360_band_patch *patches, int n_patches, int H, int W,
                         const p360_tile_maps *maps_host, void *stream);
int p360_pyramid_dims(int pw, int ph, int pad, int32_t out_host[4]);
int p360_pyramid_reduce_batch(const p360_band_patch *patches, int n_patches, int max_w4,
                              int max_h4, const uint64_t *owner_keys, int W,
                              const p360_tile_maps *maps_host, void *stream);
int p360_multiband_collapse(const p360_band_patch *patches, int n_patches, int n_levels,
                            const uint64_t *owner_keys, const uint8_t *covered,
                            uint8_t *out_u8, int out_pitch, int y_begin, int y_end, int x_begin, int x_end,
                            int row_origin, int W, const p360_tile_maps *maps_host, void *stream);
int p360_linear_collapse(const p360_band_patch *patches, int n_patches, uint8_t *out_u8, int out_pitch,
                         int y_begin, int y_end, int row_origin, int W, void *stream);
int p360_paste_collapse(const p360_band_patch *patches, int n_patches, uint8_t *out_u8, int out_pitch,
                        int y_begin, int y_end, int row_origin, int W, void *stream);

/* ---- K8: pair overlap statistics for exposure gains (stitcher.py:48-63) ---
 * For every pair (i, j) and every pixel of image i: fixed-point perspective map
 * into image j (cv2.warpPerspective semantics, zero destination), overlap =
 * warped alpha != 0.  out[3*p + 0] = overlap count of pair p, [3*p + 1] = sum of
 * image-i rgb over the overlap, [3*p + 2] = sum of warped image-j rgb.  All
 * images share h, w, src_c (as in the reference).  `jobs` is a DEVICE array;
 * inv = INVERSE of the i <- j homography in un-centred pixel coordinates.
 * partial: scratch of at least 3 * n_pairs * p360_pair_stats_blocks(h, w) float64.
 */
typedef struct p360_pair_job {
    const uint8_t *src_i;
    const uint8_t *src_j;
    double inv[9];
} p360_pair_job;

int p360_pair_stats_blocks(int h, int w);
int p360_pair_overlap_stats(const p360_pair_job *jobs, int n_pairs, int h, int w, int src_c,
                            const float *lut, const double *hat_y, const double *hat_x,
                            double *partial, double *out, void *stream);

/* ---- valid-area mask for the crop stage (stitcher.py:266-271) -------------*/
int p360_cover_update(const uint8_t *invalid, int pw, int ph, int x0, int y0,
                      uint8_t *covered, int W, void *stream);

/* ---- the reference's multiband loop nest at full resolution (stitcher.py:186-241) ---------
 * Stage by stage, as SURVEY.md §8(b) lists them: the owner mask into alpha (:207-208), one
 * p360_gauss_blur per patch and level (:226), band x weight accumulated per level into a mosaic-
 * sized {sum r*w, sum g*w, sum b*w, sum w} float4 image (:224-232; cur_rgba == NULL: the last
 * level), then per-level normalisation, sum, clip, uint8 (:236-241) over acc_rgbw =
 * [n_levels][H][W] float4.  FP32-issue-bound; used as the device-side ground truth of the
 * coarse-grid pipeline and for images too small for it (pano360_b200/compositor.py). */
int p360_owner_to_alpha(float *rgba, int pw, int ph, int x0, int y0, int idx,
                        const uint64_t *owner_keys, int W, void *stream);
int p360_band_accumulate(const float *prev_rgba, const float *cur_rgba, int pw, int ph, int x0, int y0,
                         float *acc_rgbw, int W, void *stream);
int p360_exact_collapse(const float *acc_rgbw, int n_levels, const uint8_t *covered, uint8_t *out_u8,
                        int H, int W, void *stream);

/* ---- crop stage: largest all-valid rectangle (stitcher.py:340-369, crop_mosaic) ------------
 * covered: H x W u8 union of valid pixels (stitcher.py:266-271).  rect_dev receives
 * {y0, y1, x0, x1} (int32, device): the crop is mosaic[y0:y1, x0:x1]; all zero if nothing is
 * valid.  Same scan order and tie rules as the reference (first strictly larger area in
 * (row, column) order; column 0 never extends to the right, its :359 loop stops at j = 1).
 * scratch: p360_crop_scratch_bytes(H, W) bytes, 16-byte aligned. */
int64_t p360_crop_scratch_bytes(int H, int W);
int p360_crop_rect(const uint8_t *covered, int H, int W, void *scratch, int32_t *rect_dev, void *stream);

/* ---- ingest: cv2.resize(img, None, fx=1/S, fy=1/S) on uint8 (stitcher.py:418-421, `-s`) ------
 * Bit-exact with OpenCV's 8-bit path (resize.cpp).  src: h x w x c, dst: dh x dw x c (c = 1, 3, 4).
 * area2 != 0: the exact 2x shrink, which OpenCV reroutes from INTER_LINEAR to INTER_AREA's 2 x 2
 * integer mean (tables unused).  Else INTER_LINEAR in 11-bit fixed point from DEVICE tables built as
 * resize.cpp builds them (pano360_b200/geometry.py: resize_tables): xofs[dw] / yofs[dh] = first
 * source index, xw[2 dw] / yw[2 dh] = the two int16 weights (sum 2048).  Column indices are already
 * clamped (fraction 0 at the edges), row indices are clipped by the kernel. */
int p360_resize_u8(const uint8_t *src, int h, int w, int c, uint8_t *dst, int dh, int dw,
                   const int32_t *xofs, const int16_t *xw, const int32_t *yofs, const int16_t *yw,
                   int area2, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PANO360_B200_H */
