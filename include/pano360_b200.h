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
 * cv2.remap INTER_LINEAR/BORDER_REFLECT, alpha *= ~mask), fused.
 *   src         u8, src_h x src_w x src_c (src_c = 3 or 4; 4th channel ignored)
 *   lut         256 float32: value of a u8 sample (u8/255, optionally
 *               gain-scaled and clipped, stitcher.py:65-66)
 *   hat_y/hat_x float64 tables of `_hat(h)` / `_hat(w)` (stitcher.py:251-254)
 *   col_tab     pw x 3 float64: K*R[:,0]*rx(c) + K*R[:,2]*rz(c) per patch column
 *   row_tab     ph x 3 float64: K*R[:,1]*ry(r) per patch row
 *               (proj2hom is separable: stitcher.py:84-87, :101-104)
 *   out_rgba    ph x pw x 4 float32, out_invalid ph x pw u8 (1 = masked)
 *   best/owner/covered  optional (all NULL to skip): the K2 owner-map update of
 *               p360_owner_update fused into the same pass, for a patch placed
 *               at (x0, y0) in a mosaic of width W and known as `idx`.
 * p360_pack_rgbx widens u8 x 3 pixels to one aligned 32-bit word each so that
 * every bilinear tap is a single load (src_c = 4).
 */
int p360_pack_rgbx(const uint8_t *src_rgb, uint8_t *dst_rgbx, int64_t n_pixels, void *stream);
int p360_warp_patch(const uint8_t *src, int src_h, int src_w, int src_c,
                    const float *lut, const double *hat_y, const double *hat_x,
                    const double *col_tab, const double *row_tab,
                    int pw, int ph, float *out_rgba, uint8_t *out_invalid,
                    int x0, int y0, int idx, float *best, int32_t *owner,
                    uint8_t *covered, int W, void *stream);

/* ---- K2: owner map (stitcher.py:196-208) -----------------------------------
 * p360_owner_update: running arg-max of alpha over patches visited in index
 * order; strict '>' keeps the first maximum like np.argmax.  best must start
 * at 0 and owner at -1.  Also ORs `!invalid` into covered (stitcher.py:233-234).
 * p360_owner_to_alpha: alpha := (owner == idx) in place (stitcher.py:207-208).
 */
int p360_owner_update(const float *rgba, const uint8_t *invalid, int pw, int ph,
                      int x0, int y0, int idx, float *best, int32_t *owner,
                      uint8_t *covered, int W, void *stream);
int p360_owner_to_alpha(float *rgba, int pw, int ph, int x0, int y0, int idx,
                        const int32_t *owner, int W, void *stream);

/* ---- K3: cv2.GaussianBlur(rgba, (0,0), sigma) (stitcher.py:226) -----------
 * Separable float32 convolution with BORDER_REFLECT_101 at the patch edges.
 * taps: ksize float32 (host computes cv2.getGaussianKernel semantics).
 * tmp: scratch of the same size as in/out.  in may alias neither out nor tmp.
 */
int p360_gauss_blur(const float *in_rgba, float *out_rgba, float *tmp_rgba,
                    int pw, int ph, const float *taps_host, int ksize,
                    void *stream);

/* ---- K4: band weighted accumulate (stitcher.py:224-232) -------------------
 * acc is H x W float4 {sum band*wgt (3), sum wgt}.
 * cur != NULL : band = prev.rgb - cur.rgb, wgt = cur.a   (levels 0 .. L-2)
 * cur == NULL : band = prev.rgb,           wgt = prev.a  (last level)
 */
int p360_band_accumulate(const float *prev_rgba, const float *cur_rgba,
                         int pw, int ph, int x0, int y0, float *acc, int W,
                         void *stream);

/* ---- K5: collapse + normalise + clamp (stitcher.py:236-241) ---------------
 * mosaic = sum_l covered ? acc_l.rgb / (acc_l.w == 0 ? 1 : acc_l.w) : 0, then
 * out = trunc(255 * clip(mosaic, 0, 1)).  acc holds n_levels planes of H*W
 * float4, level-major.  out_u8 is H x W x 3.
 */
int p360_collapse_finalize(const float *acc, int n_levels, const uint8_t *covered,
                           uint8_t *out_u8, int64_t n_pixels, void *stream);

/* ---- K6: linear blend (stitcher.py:171-183) --------------------------------*/
int p360_linear_accumulate(const float *rgba, const uint8_t *invalid, int pw, int ph,
                           int x0, int y0, float *acc, int W, void *stream);
int p360_linear_finalize(const float *acc, uint8_t *out_u8, int64_t n_pixels,
                         void *stream);

/* ---- K7: paste without blending (stitcher.py:160-168) ----------------------*/
int p360_paste(const float *rgba, const uint8_t *invalid, int pw, int ph,
               int x0, int y0, uint8_t *mosaic_u8, int W, void *stream);

/* ---- K8: pair overlap statistics for exposure gains (stitcher.py:48-63) ---
 * For every pixel of image i: fixed-point perspective map into image j
 * (cv2.warpPerspective semantics, zero destination), overlap = warped alpha
 * != 0.  out[0] = overlap count, out[1] = sum of image-i rgb over the overlap,
 * out[2] = sum of warped image-j rgb.  inv_hom_host: 9 float64, the INVERSE of
 * the i<-j homography in un-centred pixel coordinates.  partial: scratch of
 * at least 3 * p360_pair_stats_blocks(h, w) float64.
 */
int p360_pair_stats_blocks(int h, int w);
int p360_pair_overlap_stats(const uint8_t *src_i, const uint8_t *src_j,
                            int h, int w, int src_c, const float *lut,
                            const double *hat_y, const double *hat_x,
                            const double *inv_hom_host, double *partial,
                            double *out3, void *stream);

/* ---- reduced-resolution band pipeline (multiband hot path) ------------------
 * The blurs of stitcher.py:226 are evaluated on coarse grids (f = 2 for level
 * 0, f = 4 above) of the BORDER_REFLECT_101 extension of the patch by `pad`
 * full-resolution pixels (pad % 4 == 0):
 *   p360_pyramid_dims    -> {w2, h2, w4, h4}: sizes of the two coarse images
 *   p360_pyramid_reduce  area-reduce rgba (alpha := owner == idx when owner is
 *                        not NULL, stitcher.py:207-208) into d2 (f = 2) and d4
 *                        (f = 4); blur them with p360_gauss_blur afterwards
 *   p360_multiband_collapse  for every mosaic pixel, in patch order: expand
 *                        the coarse levels bilinearly, form the bands and
 *                        weights (stitcher.py:224-232), normalise per level,
 *                        sum, clamp, truncate to uint8 (stitcher.py:236-241).
 *                        Nothing is accumulated in HBM.
 */
typedef struct p360_band_patch {
    const float *rgba;                        /* full-res patch (alpha ignored)        */
    const float *low[P360_MAX_LEVELS - 1];    /* blurred coarse image of level l       */
    int32_t lw[P360_MAX_LEVELS - 1];          /* its width in coarse pixels            */
    int32_t shift[P360_MAX_LEVELS - 1];       /* log2 of its reduction factor (1 or 2) */
    int32_t x0, y0, pw, ph;                   /* box in (window) mosaic pixels         */
    int32_t pad;                              /* extension in full-res pixels          */
    int32_t index;                            /* id of this patch in the owner map     */
} p360_band_patch;

int p360_pyramid_dims(int pw, int ph, int pad, int32_t out_host[4]);
int p360_pyramid_reduce(const float *rgba, int pw, int ph, int x0, int y0, int idx,
                        const int32_t *owner, int W, int pad, float *d2, float *d4,
                        void *stream);
int p360_multiband_collapse(const p360_band_patch *patches, int n_patches, int n_levels,
                            const int32_t *owner, const uint8_t *covered,
                            uint8_t *out_u8, int H, int W, void *stream);

/* ---- valid-area mask for the crop stage (stitcher.py:266-271) -------------*/
int p360_cover_update(const uint8_t *invalid, int pw, int ph, int x0, int y0,
                      uint8_t *covered, int W, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PANO360_B200_H */
