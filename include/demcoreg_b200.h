/*
 * demcoreg_b200 -- C-ABI of the B200-native co-registration hot path of dshean/demcoreg.
 *
 * Every entry point takes plain pointers and sizes (no torch types).  Device
 * pointers are CALLER-OWNED (e.g. torch tensor .data_ptr()); the library never
 * allocates result buffers.  Scratch comes from a caller-provided workspace whose
 * size is queried with dcr_*_workspace_bytes().  `stream` is a cudaStream_t passed
 * as void* (torch.cuda.current_stream().cuda_stream).  All functions return
 * 0 on success, <0 for an invalid argument, >0 for a CUDA error code; the message
 * is available (thread-local) from dcr_last_error().  Nothing here calls exit().
 *
 * Rasters are float32, row-major, rows north->south, `stride` in ELEMENTS.
 * Masks are uint8 with 1 = masked, matching numpy.ma.
 *
 * Each function cites the reference interface it replaces (file:line in
 * dshean/demcoreg, or the un-vendored pygeotools/GDAL routine that file calls).
 */
#ifndef DEMCOREG_B200_H
#define DEMCOREG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCR_NBINS 360

/* bits of the per-pixel mask byte written by dcr_warp_diff_horn / updated by the filters */
#define DCR_M_BASE   0x01u /* ref or src nodata, or static (stable-surface) mask: dem_align.py:79-86 */
#define DCR_M_MAXDZ  0x02u /* |dh| > max_dz: dem_align.py:39 */
#define DCR_M_MAD    0x04u /* outside med +- 3*NMAD: dem_align.py:46 (filtlib.mad_fltr) */
#define DCR_M_SLOPE  0x08u /* slope nodata or outside slope_lim: dem_align.py:51-60, 107 */
#define DCR_M_ASPECT 0x10u /* aspect nodata (flat / border / nodata window): dem_align.py:100 */
#define DCR_M_DIFF   (DCR_M_BASE | DCR_M_MAXDZ | DCR_M_MAD | DCR_M_SLOPE) /* final static_mask, dem_align.py:110 */

const char* dcr_last_error(void);
int dcr_version(void);
/* sizeof() of the ABI structs, for binding self-checks: 0 raster_desc, 1 warp_geom, 2 grid, 3 front_args,
 * 4 nuth_stats, 5 iter_result, 6 align_iter, 7 align_state, 8 poly2d; -1 for an unknown index */
int dcr_struct_size(int which);

/* ---- geometry (host only) -------------------------------------------------------------
 * One raster of a warp: geotransform, size, nodata.  Replaces the bookkeeping of
 * pygeotools warplib.memwarp_multi(extent='intersection') <- dem_align.py:66-67, 298-299. */
typedef struct dcr_raster_desc {
    double gt[6];      /* GDAL geotransform (x0, res, 0, y0, 0, -res) */
    int32_t width;     /* RasterXSize */
    int32_t height;    /* RasterYSize */
    double ndv;        /* band nodata value */
    int32_t has_ndv;
    int32_t _pad;
} dcr_raster_desc;

/* Per-raster constants of the translation-only cubic warp. */
typedef struct dcr_warp_geom {
    int32_t ix0, iy0;      /* floor(ox), floor(oy): first tap of output (i,j) is source (i+iy0-1, j+ix0-1) */
    int32_t nx0, ny0;      /* floor(ox+.5), floor(oy+.5): the "nearest" source pixel GDAL tests first */
    double fx, fy;         /* fractional offsets (bilinear fallback weights) */
    double wx[4], wy[4];   /* Catmull-Rom (a=-0.5) weights, GDAL GWKCubicComputeWeights */
    int32_t identity;      /* 1: memwarp_multi returns this dataset unchanged (no resample, ndv kept) */
    int32_t _pad;
} dcr_warp_geom;

typedef struct dcr_grid {
    double gt[6];          /* geotransform of the common grid */
    int32_t width, height; /* int(round(extent/res)), Python round (half to even) */
} dcr_grid;

/* memwarp_multi([a, b], res, extent='intersection'): common grid + per-raster geometry.
 * `res` <= 0 selects res='max' of the two inputs.  Returns -2 when the extents do not intersect,
 * -3 when the resolutions differ (the down-sampling GWKResample path is not part of this library). */
int dcr_plan_intersection(const dcr_raster_desc* a, const dcr_raster_desc* b, double res,
                          dcr_grid* grid, dcr_warp_geom* ga, dcr_warp_geom* gb);
/* geometry of one raster onto a given grid (used for the world-fixed stable-surface mask) */
int dcr_plan_onto_grid(const dcr_raster_desc* a, const dcr_grid* grid, dcr_warp_geom* ga);

/* ---- K1: cubic warp -------------------------------------------------------------------
 * GDAL ReprojectImage(GRA_Cubic) for a pure translation, as reached through
 * warplib.memwarp_multi <- dem_align.py:66-67 (also 298-299, 368-369).  dst nodata `dst_ndv`. */
int dcr_warp_cubic(const float* src, int src_h, int src_w, int64_t src_stride, double src_ndv, int has_ndv,
                   const dcr_warp_geom* geom, float* dst, int h, int w, int64_t dst_stride, float dst_ndv,
                   void* stream);

/* GDAL cubic warp between north-up grids of the same SRS but different pixel size: the one-time
 * warplib.memwarp_multi(ds_list, res=args.res, extent='intersection') of dem_align.py:298-299 / compute_diff.py:97 when the
 * inputs do not share a resolution (-res min|max|mean, -tr).  scale = source res / grid res: >= 0.95 -> the 4x4 Catmull-Rom
 * at per-pixel fractions (bilinear near nodata / edges), < 0.95 -> GDAL's scaled-footprint GWKResample (radius ceil(2/scale),
 * weight-normalised, nodata pixels skipped).  dst: grid->height x grid->width, nodata `dst_ndv`. */
int dcr_resample_cubic(const float* src, const dcr_raster_desc* src_desc, int64_t src_stride, const dcr_grid* grid,
                       float* dst, int64_t dst_stride, double dst_ndv, void* stream);

/* ---- K2: Horn slope / aspect ----------------------------------------------------------
 * gdal.DEMProcessing('slope'|'aspect', computeEdges=False) as reached through
 * geolib.gdaldem_mem_ds <- dem_align.py:54, 100.  Outputs use nodata -9999.  Either output may be NULL. */
int dcr_horn_slope_aspect(const float* dem, int h, int w, int64_t stride, double ndv, int has_ndv,
                          double ewres, double nsres, float* slope, float* aspect, void* stream);

/* ---- K1+K2 fused front end of one iteration ---------------------------------------------
 * dem_align.py:66-107: warp ref and src onto the common grid, dh = src - ref, nodata / static
 * mask / |dh|<=max_dz, Horn slope (range-filtered) and aspect of the WARPED src.  One kernel. */
typedef struct dcr_front_args {
    const float* ref; int64_t ref_stride; dcr_raster_desc ref_desc; dcr_warp_geom ref_geom;
    const float* src; int64_t src_stride; dcr_raster_desc src_desc; dcr_warp_geom src_geom;
    const uint8_t* smask; int64_t smask_stride; int32_t smask_h, smask_w; /* world-fixed static mask or NULL */
    int32_t smask_nx0, smask_ny0;                                        /* nearest-neighbour offsets onto the grid */
    int32_t h, w;                      /* common grid */
    double ewres, nsres;               /* gt[1], gt[5] of the common grid */
    double dst_ndv;                    /* memwarp_multi dst_ndv (0) for resampled rasters */
    float max_dz; int32_t use_max_dz;  /* dem_align.py:39 */
    float slope_lo, slope_hi;          /* dem_align.py:58 */
    /* outputs (device, contiguous h*w): */
    float* dh; float* slope; float* aspect; uint8_t* maskbits;
    float* ref_w; float* src_w;        /* optional warped rasters (NULL to skip) */
    /* row-band mode (all zero = whole rasters): compute output rows [row_begin, row_end) only; `ref`, `src`, `smask`
     * then point at the first row HELD IN MEMORY, global row index *_row0, *_nrows rows (band + halo); outputs are
     * band-local ((row_end - row_begin) x w).  -5 is returned when a halo is too small for the current shift. */
    int32_t row_begin, row_end;
    int32_t ref_row0, ref_nrows, src_row0, src_nrows, smask_row0, smask_nrows;
    /* deferred apply_z_shift (coreglib.py:42-55): up to 16 scalar shifts of the SOURCE band that have not been written
     * back yet; the kernel applies them, in order and in float32, to every source pixel it loads (pixels within numpy's
     * masked_values tolerance of ndv become ndv) -- bit-identical to running dcr_apply_z_shift first, without the extra
     * read-modify-write pass over the raster. */
    int32_t src_ndz, _pad2;
    float src_dz[16];
} dcr_front_args;
int dcr_warp_diff_horn(const dcr_front_args* args, void* stream);

/* ---- K3: exact order statistics ---------------------------------------------------------
 * numpy.percentile (linear interpolation) over the unmasked elements, as used by
 * malib.fast_median / mad / calcperc <- coreglib.py:88, 214-215; filtlib.mad_fltr <- dem_align.py:46.
 * Element i takes part iff mask == NULL or (mask[i] & reject) == 0, and the value is not NaN.
 * transform 0: x ; 1: |x - center| evaluated in float32 (malib.mad).
 * Results (host): out[q] = interpolated value, count = number of participating elements.
 * Exact integer rank (float64 virtual index) + numpy's _lerp in float32.  */
size_t dcr_select_workspace_bytes(void);          /* enough for the radix variant and for n <= 393216 */
size_t dcr_select_workspace_bytes_n(int64_t n);   /* dcr_select_quantiles on n elements */
int dcr_select_quantiles(const float* x, const uint8_t* mask, uint32_t reject, int64_t n,
                         int transform, float center, const double* q_percent, int nq,
                         float* out, int64_t* count, void* workspace, size_t workspace_bytes, void* stream);

/* Same result through the 3-pass radix select only (dcr_select_quantiles tries the one-pass
 * sample-bracket select first and falls back to this when its on-device verification fails). */
int dcr_select_quantiles_radix(const float* x, const uint8_t* mask, uint32_t reject, int64_t n,
                               int transform, float center, const double* q_percent, int nq,
                               float* out, int64_t* count, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K4: moments ------------------------------------------------------------------------
 * count / min / max / mean / std (float64 accumulation) <- malib.get_stats (dem_align.py:114, 384, 392),
 * robust_stats.py:55-65.  out = {count, min, max, mean, std(ddof=0), sum_sq} */
int dcr_moments(const float* x, const uint8_t* mask, uint32_t reject, int64_t n, double out[6],
                void* workspace, size_t workspace_bytes, void* stream);

/* ---- K5: strided subsample of the compressed array ---------------------------------------
 * malib.get_stats: compressed()[::stride] in row-major order of the unmasked pixels
 * (<- dem_align.py:114).  Writes ceil(count/stride) values to dense; returns their number. */
int dcr_compress_strided(const float* x, const uint8_t* mask, uint32_t reject, int64_t n, int64_t stride,
                         float* dense, int64_t dense_capacity, int64_t* n_out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- K6: Nuth-Kaab aspect-bin statistics ---------------------------------------------------
 * coreglib.py:221-263: common mask, y = dh/tan(deg2rad(slope)) (float32), 3-sigma trim on y,
 * 360 one-degree bins: count and exact median (scipy.stats.binned_statistic).  */
typedef struct dcr_nuth_stats {
    int64_t n_common;            /* samples after the common mask: coreglib.py:221-228 */
    int64_t n_final;             /* samples after the 3-sigma trim: coreglib.py:241-245 */
    float mean_y, std_y;         /* float32 statistics of y: coreglib.py:238 */
    float rmin, rmax;            /* coreglib.py:240-241 */
    int64_t bin_count[DCR_NBINS];
    float bin_med[DCR_NBINS];    /* NaN for empty bins */
} dcr_nuth_stats;
size_t dcr_nuth_workspace_bytes(int64_t n);
int dcr_nuth_bin_stats(const float* dh, const float* slope, const float* aspect, const uint8_t* maskbits,
                       uint32_t reject_dh, uint32_t reject_slope, uint32_t reject_aspect, int64_t n,
                       int remove_outliers, dcr_nuth_stats* out /* host */,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- K7a: NCC offset search -------------------------------------------------------------------
 * coreglib.compute_offset_ncc <- coreglib.py:118-198: z-score ref and kernel = src[p:-p, p:-p] over their
 * valid pixels (float64 mean/std), masked pixels <- noise (device arrays of the ref / kernel shapes, or NULL
 * for zeros; the reference draws unseeded np.random.randn), m = correlate2d(ref, kernel, 'valid').
 * ref/src: h x w float32 with the same element stride; masks uint8 (1 = masked) or NULL.
 * m_out (host): (2p+1) x (2p+1) float64, row-major.  stats_out (host): ref mean, std, kernel mean, std. */
size_t dcr_ncc_workspace_bytes(int h, int w, int pad);
int dcr_ncc_corr(const float* ref, const uint8_t* ref_mask, const float* src, const uint8_t* src_mask, int h, int w,
                 int64_t stride, int pad, const float* ref_noise, const float* kernel_noise, double* m_out,
                 double stats_out[4], void* workspace, size_t workspace_bytes, void* stream);

/* ---- K7b: SAD offset search -------------------------------------------------------------------
 * coreglib.compute_offset_sad <- coreglib.py:65-115: for every offset, diff = ref window - kernel (float32),
 * trim outside its [P25, P75] (exact percentiles, malib.calcperc), m = sum|diff| / count.
 * m_out (host): (2p+1)^2 float64 of the mean absolute difference (the caller negates it, coreglib.py:100).
 * All (2p+1)^2 offsets go through the same few passes over the rasters (batched; see dcr_sad.cu); an offset whose
 * on-device verification fails (heavy ties, degenerate distribution) is redone by the exact per-offset path. */
size_t dcr_sad_workspace_bytes(int h, int w, int pad);
int dcr_sad_search(const float* ref, const uint8_t* ref_mask, const float* src, const uint8_t* src_mask, int h, int w,
                   int64_t stride, int pad, double* m_out, void* workspace, size_t workspace_bytes, void* stream);
/* The same with flags (DCR_SAD_PER_OFFSET: every offset through the exact per-offset path -- materialised diff + two K3
 * selections, the fallback of the batched path), the number of offsets that took the per-offset path (n_per_offset, may be
 * NULL) and, when stage_ms != NULL, the CUDA-event times of the 6 batched stages: prep, sample, pass 1, find, pass 2, final. */
#define DCR_SAD_PER_OFFSET 1
int dcr_sad_search_ex(const float* ref, const uint8_t* ref_mask, const float* src, const uint8_t* src_mask, int h, int w,
                      int64_t stride, int pad, int flags, double* m_out, int* n_per_offset, float* stage_ms /* [6] or NULL */,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- fit (host) -----------------------------------------------------------------------------
 * scipy.optimize.curve_fit(nuth_func, bin_centers, bin_med, x0) <- coreglib.py:305 with
 * nuth_func(x,a,b,c) = a*cos(deg2rad(b-x)) + c (coreglib.py:58-63).  The model is linear in
 * (a cos b, a sin b, c); the unique least-squares minimum is solved in closed form (float64 QR). */
int dcr_fit_nuth(const double* bin_centers, const double* bin_med, int n, double fit[3]);

/* ---- K8: apply_z_shift ------------------------------------------------------------------------
 * coreglib.py:42-55: a = b_getma(band); a = a + dz; WriteArray(a.filled()).  In place; pixels within
 * numpy masked_values tolerance of ndv become exactly ndv.  dz_grid (device, h*w) may replace the scalar. */
int dcr_apply_z_shift(float* band, int h, int w, int64_t stride, double ndv, float dz, const float* dz_grid,
                      void* stream);
/* n (<= 16) successive scalar apply_z_shift calls in one pass: the same ordered float32 adds (dz: host array). */
int dcr_apply_z_shift_seq(float* band, int h, int w, int64_t stride, double ndv, const float* dz, int n, void* stream);

/* ---- one full Nuth-Kaab iteration (device-resident inputs -> dx,dy,dz on the host) --------------
 * dem_align.compute_offset(mode='nuth') <- dem_align.py:62-172 and coreglib.compute_offset_nuth
 * <- coreglib.py:201-315.  All kernels are chained on `stream`; ONE device->host copy of the result. */
typedef struct dcr_iter_result {
    int32_t status;              /* 0 ok; 1 no overlapping unmasked pixels (dem_align.py:88-89);
                                    2 not enough dh samples; 3 not enough slope samples (coreglib.py:206-209);
                                    4 fit skipped: too few bins / aspect spread (coreglib.py:301) */
    int32_t n_valid_bins;
    int64_t count_base;          /* diff.count() after nodata + static mask: dem_align.py:86 */
    int64_t count_maxdz;         /* after |dh| <= max_dz: dem_align.py:39-40 */
    int64_t count_mad;           /* after mad_fltr: dem_align.py:46-48 */
    int64_t count_final;         /* after the slope mask: dem_align.py:107 */
    int64_t count_slope;         /* slope.count(): dem_align.py:59 */
    int64_t stats_stride;        /* malib.get_stats subsample stride (1 = none) */
    float med1, nmad1, mad_lo, mad_hi; /* filtlib.mad_fltr intermediates */
    float dz_med;                /* diff_stats[5]: dem_align.py:115 */
    float med_dh, med_slope;     /* coreglib.py:214-215 */
    float c_seed;                /* coreglib.py:216 */
    dcr_nuth_stats nuth;
    double fit[3];               /* a, b (deg), c */
    double dx, dy, dz;           /* RETURNED shift = (-a sin b, -a cos b, -dz_med): dem_align.py:155-156, 172 */
    float stage_ms[8];           /* DCR_ITER_PROFILE only: CUDA-event time of 0 front kernel, 1 median+MAD selects,
                                    2 mask update+scan+subsample, 3 med_dh/dz/med_slope selects, 4 y/bin prepare+moments,
                                    5 bin count+median passes, 6 device->host copy */
    int32_t select_engine;       /* 1: one-pass sample-bracket selects verified ok; 0: 3-pass radix selects (fallback) */
    int32_t z_applied;           /* DCR_ITER_APPLY_Z: the source band has already been shifted by `dz` on the device */
    int32_t bin_engine;          /* 1: two-pass shared-memory bin statistics; 0: 3-pass radix bins (heavy ties / degenerate range) */
    int32_t _pad;
} dcr_iter_result;
#define DCR_ITER_REMOVE_OUTLIERS 1 /* flags of dcr_nuth_iteration: dem_align.py:91 remove_outliers */
#define DCR_ITER_PROFILE 2         /* record per-stage CUDA events into stage_ms */
#define DCR_ITER_RADIX_ONLY 4      /* skip the one-pass sample-bracket selects (tests / profiling) */
#define DCR_ITER_APPLY_Z 8         /* also run coreglib.apply_z_shift(src, dz) on the device (front->src must be writable,
                                      whole raster); the host must then NOT apply dz again (see z_applied) */
#define DCR_ITER_TAN_SLOPE 16      /* front->slope receives tan(slope) = |gradient|/8 instead of degrees (the consumers of the
                                    * iteration want the tangent: y = dh / tan(slope), coreglib.py:224-225; the slope median,
                                    * coreglib.py:215, is taken on the tangent and converted).  Masks are still decided on the
                                    * float32 slope in degrees.  Used by dcr_align_nuth, whose slope raster is scratch. */
#define DCR_ITER_NO_BINS 128       /* skip the aspect-bin statistics and the fit (status 4, dx = dy = 0): the masks, counts and
                                      medians are all the caller wants -- the -mode ncc / sad branch, dem_align.py:126-143 */
#define DCR_ITER_RADIX_BINS 64     /* row-band mode: radix bin statistics (the fast bins declined for this raster), one-pass
                                      selections kept */
size_t dcr_iteration_workspace_bytes(int64_t n);
int dcr_nuth_iteration(const dcr_front_args* front, int flags, dcr_iter_result* out /* host */,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- the dem_align iteration loop, mode 'nuth' (dem_align.py:291-357) ---------------------------------------
 * while True: compute_offset -> totals -> apply_xy_shift (geotransform) -> apply_z_shift (band) -> stop when the
 * incremental shift magnitude < tol or after max_iter iterations.  The state lives in a caller-owned struct so that
 * one call can run the whole loop (max_steps <= 0) or a bounded number of iterations (max_steps > 0), with no host
 * work between iterations other than the fit.  The source raster is modified IN PLACE (pass a copy to keep the
 * original, as the reference does with mem_drv.CreateCopy, dem_align.py:294): its geotransform in src_desc.gt; its band
 * through deferred z shifts (src_ndz/src_dz, applied by the front kernel while loading, folded into the band every
 * DCR_ALIGN_MAX_PENDING iterations and by dcr_align_flush_z). */
#define DCR_ALIGN_MAX_PENDING 4
typedef struct dcr_align_iter {
    double dx, dy, dz;           /* incremental shift returned by compute_offset */
    int32_t status;              /* dcr_iter_result.status */
    int32_t select_engine;
    int32_t h, w;                /* common grid of the iteration */
    int32_t bin_engine, _pad;    /* dcr_iter_result.bin_engine */
    double fit[3];               /* a, b (deg), c of the iteration (NaN when status == 4) */
    float dz_med, med_slope;     /* for the reference's per-iteration 'Median dz / Nuth dz' message (dem_align.py:157-159) */
} dcr_align_iter;
typedef struct dcr_align_state {
    /* inputs (device pointers, caller-owned) */
    const float* ref; int64_t ref_stride; dcr_raster_desc ref_desc;
    float* src; int64_t src_stride; dcr_raster_desc src_desc;                 /* shifted in place */
    const uint8_t* smask; int64_t smask_stride; dcr_raster_desc smask_desc;   /* world-fixed static mask or NULL */
    float* dh; float* slope; float* aspect; uint8_t* maskbits;                /* iteration outputs, out_capacity pixels each;
                                                                                 aspect may be NULL (nothing in the loop reads it) */
    int64_t out_capacity;
    double tol, max_offset;      /* dem_align.py -tol, -max_offset */
    float max_dz, slope_lo, slope_hi;
    int32_t max_iter;
    int32_t remove_outliers;
    /* loop state: zero before the first call, carried across calls */
    int32_t n_done;              /* iterations finished */
    int32_t finished;            /* 1: converged (dm < tol) or max_iter reached */
    int32_t exit_code;           /* 0 ok; 1-3: compute_offset sys.exit (dcr_iter_result.status); 5: total offset > max_offset */
    int32_t src_ndz;             /* pending z shifts of the band */
    float src_dz[16];
    double dx_total, dy_total;
    float dz_total;              /* accumulated in float32 like the reference (dz is np.float32) */
    float _pad;
    dcr_iter_result last;        /* full result of the last iteration */
} dcr_align_state;
/* Runs iterations until the loop finishes or max_steps (> 0) iterations were done in this call; appends one record per
 * iteration to iters[0..iters_cap) (may be NULL); *n_iters = iterations run by this call.  flags: DCR_ITER_PROFILE,
 * DCR_ITER_RADIX_ONLY, DCR_ITER_TAN_SLOPE (st->slope is then scratch holding tan(slope)).  Workspace: dcr_iteration_workspace_bytes(out_capacity). */
int dcr_align_nuth(dcr_align_state* st, int max_steps, int flags, dcr_align_iter* iters, int iters_cap, int* n_iters,
                   void* workspace, size_t workspace_bytes, void* stream);
/* Writes the pending z shifts into the source band (call before reading the aligned band). */
int dcr_align_flush_z(dcr_align_state* st, void* stream);

/* ---- -tiltcorr: 2-D polynomial fit of the filtered residuals and its removal --------------------------------
 * dem_align.py:396-447, 478-483: geolib.ma_fitpoly(diff_align_filt, order, gt, perc=(0,100), origmask=False) ->
 * geolib.polyfit2d (least squares on x^i y^j, (i, j) in product(range(order+1), repeat=2), x / y = map coordinates of the
 * pixel centres), geolib.polyval2d, coreglib.apply_z_shift(ds, -valgrid), diff_align -= vals.
 * The fit runs in centred / scaled coordinates u = (x - x0)/xs, v = (y - y0)/ys (same surface: the tensor-product basis is
 * closed under affine maps; well-conditioned normal equations); `coeff` is that surface expanded in the reference's raw
 * basis (informational: in raw coordinates the coefficients are ill-determined, parity is on the surface). */
typedef struct dcr_poly2d {
    int32_t order, _pad;
    double x0, y0, xs, ys;
    double coeff_norm[16];       /* c[i*(order+1)+j] of u^i v^j */
    double coeff[16];            /* m[i*(order+1)+j] of x^i y^j (geolib.polyfit2d ordering) */
} dcr_poly2d;
size_t dcr_polyfit2d_workspace_bytes(void);
/* z: device h x w float32 (element stride `stride`); mask: device h*w uint8 (contiguous) or NULL; a pixel takes part iff
 * (mask & reject) == 0 and z is not NaN.  1 <= order <= 3.  model / count: host.  -6: too few samples / singular. */
int dcr_polyfit2d(const float* z, const uint8_t* mask, uint32_t reject, int h, int w, int64_t stride, const double gt[6],
                  int order, dcr_poly2d* model, int64_t* count, void* workspace, size_t workspace_bytes, void* stream);
/* vals (device h x w float32) = polynomial at the pixel centres of the grid with geotransform gt */
int dcr_polyval2d(const dcr_poly2d* model, const double gt[6], int h, int w, float* vals, int64_t vals_stride, void* stream);
/* band <- float32(float64(band) + sign * polynomial), nodata pixels (numpy masked_values tolerance) left at ndv:
 * coreglib.apply_z_shift(ds, sign * valgrid) for a float64 valgrid, and `diff_align -= vals` (has_ndv = 0, sign = -1) */
int dcr_polyval2d_apply(float* band, int h, int w, int64_t stride, const double gt[6], double ndv, int has_ndv,
                        const dcr_poly2d* model, double sign, void* stream);

/* ---- raster pieces next to the hot path ------------------------------------------------------
 * dcr_axis_median: np.ma.median(a, axis) and np.ma.count(a, axis) of coreglib.successive_med (coreglib.py:537-538,
 *   562-563): exact median of the unmasked, non-NaN elements of every row (axis = 1) or column (axis = 0); even counts give
 *   the float32 mean of the two middle values; a fully masked line gives NaN and count 0.  mask: uint8, != 0 = masked, or NULL.
 *   Workspace only for lines longer than 40960 elements (dcr_axis_median_workspace_bytes).
 * dcr_axis_subtract: a -= expand_dims(corr, axis) (coreglib.py:553, 580).
 * dcr_binary_dilate: scipy.ndimage.binary_dilation(mask, iterations = n), default cross structure, border 0
 *   (dem_mask.py:561-565); mask is 0/1 uint8, updated in place; scratch h*w bytes.
 * dcr_mask_from_classes: out = lut[classes] (the NLCD class filters of dem_mask.py:111-139);
 * dcr_mask_threshold: out = x > thresh (greater = 1; bare ground, dem_mask.py:141-157) or x <= thresh (greater = 0; TOA
 *   reflectance, dem_mask.py:403-409). */
size_t dcr_axis_median_workspace_bytes(int h, int w, int axis);
int dcr_axis_median(const float* a, const uint8_t* mask, int h, int w, int64_t stride, int64_t mask_stride, int axis,
                    float* med, int64_t* count, void* workspace, size_t workspace_bytes, void* stream);
int dcr_axis_subtract(float* a, int h, int w, int64_t stride, const float* corr, int axis, void* stream);
int dcr_binary_dilate(uint8_t* mask, int h, int w, int iterations, uint8_t* scratch, void* stream);
int dcr_mask_from_classes(const uint8_t* classes, int64_t n, const uint8_t lut[256], uint8_t* out, void* stream);
int dcr_mask_threshold(const float* x, int64_t n, float thresh, int greater, uint8_t* out, void* stream);

/* ---- row-band mode: one very large raster split into bands of output rows, one band per GPU ----------------
 * (BASELINE.json configs[4]; the reference does not tile -- it loads whole arrays, dem_align.py:79-80.)
 * The iteration of dcr_nuth_iteration is cut into dcr_band_nphases() phases.  Each rank calls dcr_band_phase for
 * its band (front->row_begin/row_end, rasters held as band + halo, see dcr_front_args), then all-reduces (sum)
 * the (up to 8) device buffers the phase declares -- kind 0 = int32 histograms, 1 = int64 counters, 2 = float64 sums -- over
 * all ranks (ncclAllReduce / torch.distributed.all_reduce on `stream`) and calls the next phase.  Halo rows of the
 * rasters are exchanged by the caller (ncclSend/Recv).  After the last phase dcr_band_result() returns the same
 * dcr_iter_result on every rank.  Workspace: dcr_iteration_workspace_bytes(rows_in_band * w). */
#define DCR_STATUS_RERUN_RADIX 100   /* dcr_iter_result.status of dcr_band_result: a one-pass selection (select_engine = 0) or the
                                      * fast bins (bin_engine = 0) declined, on every rank alike; run the iteration again with
                                      * DCR_ITER_RADIX_ONLY resp. DCR_ITER_RADIX_BINS (dcr_band_iteration does, and reports the
                                      * engines that finally ran so that the caller can keep the flag for the next iterations) */
int dcr_band_nphases(void);
int dcr_band_phase(const dcr_front_args* front, int flags, int rank, int nranks, int phase, void* workspace,
                   size_t workspace_bytes, void** reduce_ptr, int64_t* reduce_count, int* reduce_kind, int* nreduce,
                   void* stream);
int dcr_band_result(const dcr_front_args* front, void* workspace, size_t workspace_bytes, dcr_iter_result* out,
                    void* stream);

/* The collective side in the library (one communicator per handle, injected or created here).  NCCL is resolved
 * with dlopen("libnccl.so.2") on first use -- inside a PyTorch process that is the copy torch has loaded.
 *   dcr_comm_unique_id: rank 0 makes the 128-byte ncclUniqueId and hands it to the other ranks by any means;
 *   dcr_comm_init:      ncclCommInitRank (collective over the nranks callers);  dcr_comm_set: wrap an ncclComm_t the
 *                       caller owns;  dcr_comm_destroy: frees the handle (and the communicator it created).
 *   dcr_band_exchange_halos: grouped ncclSend/ncclRecv of the halo rows of one band raster with rank-1 / rank+1
 *                       (row indices are stored-row indices of `band`, row_bytes its pitch in bytes).
 *   dcr_band_iteration: ONE row-band iteration -- every phase of dcr_band_phase followed by the ncclAllReduce (sum) of the
 *                       buffers it declares, back to back on `stream` (no host work or synchronisation in between), then
 *                       dcr_band_result.  Collective.  Returns 100 + ncclResult_t on NCCL errors, -7 without libnccl. */
typedef struct dcr_comm dcr_comm;
int dcr_comm_unique_id(void* id, size_t id_bytes);
int dcr_comm_init(const void* id, size_t id_bytes, int rank, int nranks, dcr_comm** comm);
int dcr_comm_set(void* nccl_comm, int rank, int nranks, dcr_comm** comm);
int dcr_comm_destroy(dcr_comm* comm);
int dcr_comm_rank(const dcr_comm* comm, int* rank, int* nranks);
int dcr_band_exchange_halos(dcr_comm* comm, void* band, int64_t row_bytes, int send_up_row, int n_send_up, int recv_up_row,
                            int n_recv_up, int send_dn_row, int n_send_dn, int recv_dn_row, int n_recv_dn, void* stream);
int dcr_band_iteration(const dcr_front_args* front, int flags, dcr_comm* comm, dcr_iter_result* out, void* workspace,
                       size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEMCOREG_B200_H */
