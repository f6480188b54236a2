// K1 (cubic warp), K2 (Horn slope/aspect) and the fused K1+K2 front end of one iteration.
//
// Replaces, for the translation-only case of the dem_align loop:
//   warplib.memwarp_multi -> gdal.ReprojectImage(GRA_Cubic)   (reference dem_align.py:66-67)
//   iolib.ds_getma, diff = src - ref, static mask               (reference dem_align.py:79-86)
//   np.ma.masked_outside(diff, -max_dz, max_dz)                 (reference dem_align.py:39)
//   geolib.gdaldem_mem_ds('slope'|'aspect') + range_fltr        (reference dem_align.py:51-60, 100, 107)
//
// Arithmetic follows SURVEY.md Appendix A.1/A.2: Catmull-Rom rows-then-columns in FP64 with
// left-to-right sums and NO fused multiply-add (file compiled with -fmad=false), one rounding to
// FP32; Horn 3x3 sums in FP32 in GDAL's order, FP64 atan/atan2, one rounding to FP32.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda.h>
#include "dcr_common.cuh"
#include "dcr_internal.cuh"

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void dcr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* dcr_last_error(void) { return g_err; }
extern "C" int dcr_version(void) { return 100; }
extern "C" int dcr_struct_size(int which) {
    switch (which) {
        case 0: return (int)sizeof(dcr_raster_desc);
        case 1: return (int)sizeof(dcr_warp_geom);
        case 2: return (int)sizeof(dcr_grid);
        case 3: return (int)sizeof(dcr_front_args);
        case 4: return (int)sizeof(dcr_nuth_stats);
        case 5: return (int)sizeof(dcr_iter_result);
        case 6: return (int)sizeof(dcr_align_iter);
        case 7: return (int)sizeof(dcr_align_state);
        case 8: return (int)sizeof(dcr_poly2d);
        default: return -1;
    }
}

DcrCtx* dcr_ctx() {
    static thread_local DcrCtx slots[DCR_MAX_DEVICES];
    static thread_local bool init[DCR_MAX_DEVICES];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { dcr_set_error("cudaGetDevice failed: %s", cudaGetErrorString(e)); return nullptr; }
    if (dev < 0 || dev >= DCR_MAX_DEVICES) { dcr_set_error("device ordinal %d out of range", dev); return nullptr; }
    DcrCtx* c = &slots[dev];
    if (!init[dev]) {
        memset(c, 0, sizeof(*c));
        c->dev = dev;
        int v = 0;
        c->sms = (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) ? v : DCR_NSM_DEFAULT;
        init[dev] = true;
    }
    return c;
}

void* dcr_ctx_pinned(DcrCtx* c, size_t bytes) {
    if (c->pinned_bytes < bytes) {
        if (c->pinned) cudaFreeHost(c->pinned);
        c->pinned = nullptr;
        c->pinned_bytes = 0;
        if (cudaHostAlloc(&c->pinned, bytes, cudaHostAllocDefault) != cudaSuccess) { c->pinned = nullptr; return nullptr; }
        c->pinned_bytes = bytes;
    }
    return c->pinned;
}

int dcr_num_sms() {
    DcrCtx* c = dcr_ctx();
    return c ? c->sms : DCR_NSM_DEFAULT;
}

// ------------------------------------------------------------------------------------------
// geometry (host): SURVEY Appendix D
// ------------------------------------------------------------------------------------------
static void cubic_weights(double d, double c[4]) {
    // GDAL GWKCubicComputeWeights (Catmull-Rom, a = -0.5)
    double h = 0.5 * d, t = 3.0 * d, h2 = h * d;
    c[0] = h * (-1.0 + d * (2.0 - d));
    c[1] = 1.0 + h2 * (-5.0 + t);
    c[2] = h * (1.0 + d * (4.0 - t));
    c[3] = h2 * (-1.0 + d);
}

static void extent_of(const dcr_raster_desc* r, double e[4]) {
    e[0] = r->gt[0];
    e[3] = r->gt[3];
    e[2] = r->gt[0] + r->gt[1] * r->width;
    e[1] = r->gt[3] + r->gt[5] * r->height;
}

static double square_res(const dcr_raster_desc* r) { return (r->gt[1] + fabs(r->gt[5])) / 2.0; }

extern "C" int dcr_plan_onto_grid(const dcr_raster_desc* a, const dcr_grid* grid, dcr_warp_geom* g) {
    DCR_ARG(a && grid && g, "null pointer");
    const double res = grid->gt[1];
    const double xmin = grid->gt[0], ymax = grid->gt[3];
    double ox = (xmin - a->gt[0]) / res;
    double oy = (a->gt[3] - ymax) / res;
    double fox = floor(ox), foy = floor(oy);
    memset(g, 0, sizeof(*g));
    g->ix0 = (int32_t)fox;
    g->iy0 = (int32_t)foy;
    g->fx = ox - fox;
    g->fy = oy - foy;
    g->nx0 = (int32_t)floor(ox + 0.5);
    g->ny0 = (int32_t)floor(oy + 0.5);
    cubic_weights(g->fx, g->wx);
    cubic_weights(g->fy, g->wy);
    double e[4];
    extent_of(a, e);
    const double te[4] = {xmin, ymax - res * grid->height, xmin + res * grid->width, ymax};
    // pygeotools compares extents/resolution at 1e-3 precision; equal -> dataset returned unchanged
    bool same = fabs(square_res(a) - res) < 1e-3;
    // the target extent of memwarp_multi is the un-rounded intersection; the caller passes it through
    // dcr_plan_intersection, which sets `identity` itself.  Here: exact-lattice test only.
    for (int k = 0; k < 4; ++k) same = same && fabs(e[k] - te[k]) < 1e-3;
    g->identity = same ? 1 : 0;
    if (same) {
        g->ix0 = g->iy0 = g->nx0 = g->ny0 = 0;
        g->fx = g->fy = 0.0;
        cubic_weights(0.0, g->wx);
        cubic_weights(0.0, g->wy);
    }
    return 0;
}

extern "C" int dcr_plan_intersection(const dcr_raster_desc* a, const dcr_raster_desc* b, double res,
                                     dcr_grid* grid, dcr_warp_geom* ga, dcr_warp_geom* gb) {
    DCR_ARG(a && b && grid && ga && gb, "null pointer");
    double ra = square_res(a), rb = square_res(b);
    double tres = res > 0 ? res : (ra > rb ? ra : rb);
    if (fabs(ra - tres) > 1e-3 * tres || fabs(rb - tres) > 1e-3 * tres) {
        dcr_set_error("input resolutions differ (%.6f, %.6f, target %.6f): the down-sampling warp is out of scope",
                      ra, rb, tres);
        return -3;
    }
    double ea[4], eb[4];
    extent_of(a, ea);
    extent_of(b, eb);
    double xmin = ea[0] > eb[0] ? ea[0] : eb[0];
    double ymin = ea[1] > eb[1] ? ea[1] : eb[1];
    double xmax = ea[2] < eb[2] ? ea[2] : eb[2];
    double ymax = ea[3] < eb[3] ? ea[3] : eb[3];
    if (!(xmax > xmin) || !(ymax > ymin)) {
        dcr_set_error("no intersection between input extents");
        return -2;
    }
    grid->gt[0] = xmin;
    grid->gt[1] = tres;
    grid->gt[2] = 0.0;
    grid->gt[3] = ymax;
    grid->gt[4] = 0.0;
    grid->gt[5] = -tres;
    // Python round(): half to even == nearbyint in the default rounding mode
    grid->width = (int32_t)nearbyint((xmax - xmin) / tres);
    grid->height = (int32_t)nearbyint((ymax - ymin) / tres);
    const double te[4] = {xmin, ymin, xmax, ymax};
    const dcr_raster_desc* rs[2] = {a, b};
    dcr_warp_geom* gs[2] = {ga, gb};
    for (int k = 0; k < 2; ++k) {
        int rc = dcr_plan_onto_grid(rs[k], grid, gs[k]);
        if (rc) return rc;
        double e[4];
        extent_of(rs[k], e);
        bool same = fabs(square_res(rs[k]) - tres) < 1e-3;
        for (int m = 0; m < 4; ++m) same = same && fabs(e[m] - te[m]) < 1e-3;
        gs[k]->identity = same ? 1 : 0;
        if (same) {
            gs[k]->ix0 = gs[k]->iy0 = gs[k]->nx0 = gs[k]->ny0 = 0;
            gs[k]->fx = gs[k]->fy = 0.0;
            cubic_weights(0.0, gs[k]->wx);
            cubic_weights(0.0, gs[k]->wy);
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// device pieces shared by the kernels
// ------------------------------------------------------------------------------------------
struct DevGeom {
    int ix0, iy0, nx0, ny0;
    double fx, fy;
    double wx[4], wy[4];
    int hs, ws;
    long long stride;
    float ndv;
    int has_ndv;
    int row0;  // global index of the first row held in memory (row-band mode; 0 = whole raster)
    int row1;  // one past the last row held in memory
    int ndz;          // deferred apply_z_shift steps (source raster only)
    float dz[16];
    float zlo, zhi;   // float32 interval of numpy masked_values(|x - ndv| <= tol) for the z-shift masking
};

static DevGeom make_devgeom(const dcr_warp_geom& g, int hs, int ws, long long stride, double ndv, int has_ndv) {
    DevGeom d;
    d.ix0 = g.ix0; d.iy0 = g.iy0; d.nx0 = g.nx0; d.ny0 = g.ny0;
    d.fx = g.fx; d.fy = g.fy;
    for (int k = 0; k < 4; ++k) { d.wx[k] = g.wx[k]; d.wy[k] = g.wy[k]; }
    d.hs = hs; d.ws = ws; d.stride = stride; d.ndv = (float)ndv; d.has_ndv = has_ndv;
    d.row0 = 0;
    d.row1 = hs;
    d.ndz = 0;
    d.zlo = d.zhi = 0.f;
    for (int k = 0; k < 16; ++k) d.dz[k] = 0.f;
    return d;
}

// source pixel -> value, NaN when out of bounds / nodata / NaN
__device__ __forceinline__ float load_px(const float* __restrict__ p, const DevGeom& g, int r, int c) {
    if (r < g.row0 || c < 0 || r >= g.row1 || c >= g.ws) return __int_as_float(0x7fc00000);
    float v = __ldg(p + (long long)(r - g.row0) * g.stride + c);
    if (g.has_ndv && v == g.ndv) return __int_as_float(0x7fc00000);
    return v;
}

// x pass of the separable cubic: left-to-right FP64 sum of 4 taps
__device__ __forceinline__ double xpass(const double w[4], float p0, float p1, float p2, float p3) {
    return w[0] * (double)p0 + w[1] * (double)p1 + w[2] * (double)p2 + w[3] * (double)p3;
}
__device__ __forceinline__ double ypass(const double w[4], double v0, double v1, double v2, double v3) {
    return w[0] * v0 + w[1] * v1 + w[2] * v2 + w[3] * v3;
}

// GDAL GWKBilinearResample4Sample over the valid pixels of the 2x2 (UL, UR, LR, LL); NaN = fail.
// near = the nearest source pixel GDAL tests first (NaN -> dst nodata).
__device__ __noinline__ float bilinear_fallback(float ul, float ur, float lr, float ll, float nearest,
                                                double fx, double fy) {
    if (isnan(nearest)) return nearest;
    double acc = 0.0, div = 0.0;
    const double w_ul = (1.0 - fx) * (1.0 - fy), w_ur = fx * (1.0 - fy), w_lr = fx * fy, w_ll = (1.0 - fx) * fy;
    if (!isnan(ul)) { div = div + w_ul; acc = acc + (double)ul * w_ul; }
    if (!isnan(ur)) { div = div + w_ur; acc = acc + (double)ur * w_ur; }
    if (!isnan(lr)) { div = div + w_lr; acc = acc + (double)lr * w_lr; }
    if (!isnan(ll)) { div = div + w_ll; acc = acc + (double)ll * w_ll; }
    if (!(div >= 0.00001)) return __int_as_float(0x7fc00000);
    double r = (div == 1.0) ? acc : acc / div;
    return (float)r;
}

// Horn slope + aspect from a 3x3 float32 window (GDAL GDALSlopeHornAlg / GDALAspectAlg), A.2.
// Returns false when any input is NaN (nodata).  aspect = NaN for flat.
__device__ __noinline__ bool horn(float w0, float w1, float w2, float w3, float w4, float w5, float w6,
                                     float w7, float w8, double ewres, double nsres, bool want_slope,
                                     bool want_aspect, float* slope, float* aspect, float* tan_slope = nullptr) {
    (void)w4;
    float s = ((w0 + w1) + (w2 + w3)) + ((w4 + w5) + (w6 + w7)) + w8;  // NaN probe only
    if (isnan(s)) return false;
    const float a = (w0 + w3 + w3 + w6);
    const float b = (w2 + w5 + w5 + w8);
    const float c = (w6 + w7 + w7 + w8);
    const float d = (w0 + w1 + w1 + w2);
    const float dyf = c - d;
    if (want_slope) {
        const float dxf = a - b;
        double dx = (double)dxf / ewres;
        double dy = (double)dyf / nsres;
        double key = dx * dx + dy * dy;
        *slope = (float)(atan(sqrt(key) / 8.0) * (180.0 / M_PI));
        if (tan_slope) *tan_slope = (float)(sqrt(key) / 8.0);
    }
    if (want_aspect) {
        const float dxf = b - a;
        double dx = (double)dxf, dy = (double)dyf;
        if (dx == 0.0 && dy == 0.0) {
            *aspect = __int_as_float(0x7fc00000);
        } else {
            float asp = (float)(atan2(dy, -dx) / (M_PI / 180.0));
            asp = (asp > 90.0f) ? __fsub_rn(450.0f, asp) : __fsub_rn(90.0f, asp);
            if (asp == 360.0f) asp = 0.0f;
            *aspect = asp;
        }
    }
    return true;
}

// atan(t) for 0 <= t <= 1: odd minimax polynomial (<= 1 ulp), t + t*(t^2*P(t^2))
__device__ __forceinline__ float atan01(float t) {
    const float z = t * t;
    float p = 0.0024500289000570774f;
    p = fmaf(p, z, -0.014396979473531246f);
    p = fmaf(p, z, 0.039849750697612762f);
    p = fmaf(p, z, -0.072529748082160950f);
    p = fmaf(p, z, 0.10518480092287064f);
    p = fmaf(p, z, -0.14171802997589111f);
    p = fmaf(p, z, 0.19988775253295898f);
    p = fmaf(p, z, -0.33332940936088562f);
    return fmaf(t, z * p, t);
}

// FP32 fast path of the Horn slope/aspect with certified guard bands (DESIGN.md "front kernel"):
// the 3x3 sums are the exact float32 sums of GDAL; atan/atan2 are evaluated in FP32 (<= ~4 ulp), and every
// pixel whose result could change a mask-deciding quantity -- slope within 2e-6 (relative) of a slope
// limit, aspect within 3e-4 degrees of an integer degree (the 1-degree bin edges) -- is recomputed with
// the FP64 formulation above.  Returns false when any input is NaN; *exact = true -> caller must call horn().
__device__ __forceinline__ bool horn_fast(float w0, float w1, float w2, float w3, float w4, float w5, float w6,
                                          float w7, float w8, float c, float d, float inv8ew, float inv8ns, float slope_mid,
                                          float slope_half, float slope_tol, float* slope, float* aspect, bool* exact,
                                          float* tan_slope = nullptr) {
    // c = (w6 + w7 + w7 + w8), d = (w0 + w1 + w1 + w2): the row sums slide down the column with the window (the top row
    // of this window was the bottom row two pixels ago), so the caller carries them instead of re-adding them
    (void)w1; (void)w7;
    const float a = (w0 + w3 + w3 + w6);
    const float b = (w2 + w5 + w5 + w8);
    const float dxs = a - b;  // slope numerator; aspect uses b - a = -dxs exactly
    const float dyf = c - d;
    // any NaN (nodata) among the 8 neighbours poisons dxs or dyf; the centre is tested separately
    if (isnan(dxs + dyf) || isnan(w4)) return false;
    const float gx = dxs * inv8ew, gy = dyf * inv8ns;
    // MUFU.SQRT (sqrt.approx, <= 1 ulp-class) instead of the IEEE sqrtf sequence: the guard bands below absorb it
    float g;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(gx * gx + gy * gy));
    if (tan_slope) *tan_slope = g;
    const bool big = g > 1.0f;
    float sa = atan01(big ? __fdividef(1.0f, g) : g);
    if (big) sa = 1.5707963267948966f - sa;
    const float s = sa * 57.29577951308232f;
    // slope within guard of either limit <=> | |s - mid| - half | small (one band of width 2e-6 * slope_hi for both)
    bool ex = !(fabsf(fabsf(s - slope_mid) - slope_half) > slope_tol);
    *slope = s;
    if (dxs == 0.0f && dyf == 0.0f) {
        *aspect = __int_as_float(0x7fc00000);
    } else {
        // atan2(dy, -dx_aspect) with dx_aspect = -dxs, i.e. atan2(dyf, dxs), built from the [0,1] polynomial
        const float ax = fabsf(dxs), ay = fabsf(dyf);
        const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
        float a = atan01(__fdividef(mn, mx));
        if (ay > ax) a = 1.5707963267948966f - a;
        if (dxs < 0.0f) a = 3.141592653589793f - a;
        a = copysignf(a, dyf);
        float af = a * 57.29577951308232f;
        af = (af > 90.0f) ? (450.0f - af) : (90.0f - af);
        const float fr = af - floorf(af);
        // within 3e-4 deg of an integer degree (bin edge, incl. af == 0 / 360 and the exact multiples of 90 that an
        // infinite gradient produces) or NaN -> exact path.  af lies in [0, 360] by construction.
        ex = ex || !(fabsf(fr - 0.5f) <= 0.5f - 3e-4f);
        *aspect = af;
    }
    *exact = ex;
    return true;
}

#define DCR_GDALDEM_NDV (-9999.0f)

// ------------------------------------------------------------------------------------------
// K1 standalone: one thread per output pixel, taps through the read-only L1 path
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_px_global(const float* __restrict__ p, const DevGeom& g, int i, int j) {
    const int r0 = i + g.iy0 - 1, c0 = j + g.ix0 - 1;
    double v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
        v[r] = xpass(g.wx, load_px(p, g, r0 + r, c0), load_px(p, g, r0 + r, c0 + 1), load_px(p, g, r0 + r, c0 + 2),
                     load_px(p, g, r0 + r, c0 + 3));
    double o = ypass(g.wy, v[0], v[1], v[2], v[3]);
    if (isnan(o)) {
        return bilinear_fallback(load_px(p, g, r0 + 1, c0 + 1), load_px(p, g, r0 + 1, c0 + 2),
                                 load_px(p, g, r0 + 2, c0 + 2), load_px(p, g, r0 + 2, c0 + 1),
                                 load_px(p, g, i + g.ny0, j + g.nx0), g.fx, g.fy);
    }
    // nearest-pixel test precedes everything in GDAL
    if (isnan(load_px(p, g, i + g.ny0, j + g.nx0))) return __int_as_float(0x7fc00000);
    return (float)o;
}

__global__ void __launch_bounds__(256) warp_cubic_kernel(const float* __restrict__ src, DevGeom g, float* __restrict__ dst,
                                                         int h, int w, long long dstride, float dst_ndv) {
    const int j = blockIdx.x * 64 + (threadIdx.x & 63);
    const int i = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (i >= h || j >= w) return;
    float v = warp_px_global(src, g, i, j);
    dst[(long long)i * dstride + j] = isnan(v) ? dst_ndv : v;
}

extern "C" int dcr_warp_cubic(const float* src, int src_h, int src_w, int64_t src_stride, double src_ndv, int has_ndv,
                              const dcr_warp_geom* geom, float* dst, int h, int w, int64_t dst_stride, float dst_ndv,
                              void* stream) {
    DCR_ARG(src && dst && geom, "null pointer");
    DCR_ARG(h > 0 && w > 0 && src_h > 0 && src_w > 0, "empty raster");
    DevGeom g = make_devgeom(*geom, src_h, src_w, src_stride, src_ndv, has_ndv);
    dim3 grid((w + 63) / 64, (h + 3) / 4);
    warp_cubic_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, g, dst, h, w, dst_stride, dst_ndv);
    DCR_LAUNCH_OK("warp_cubic_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------------
// K2 standalone: Horn slope/aspect, tile + halo staged in shared memory
// ------------------------------------------------------------------------------------------
#define K2_TW 64
#define K2_TH 16
__global__ void __launch_bounds__(256) horn_kernel(const float* __restrict__ dem, int h, int w, long long stride, float ndv,
                                                   int has_ndv, double ewres, double nsres, float* __restrict__ slope,
                                                   float* __restrict__ aspect) {
    __shared__ float t[K2_TH + 2][K2_TW + 2];
    const int tj0 = blockIdx.x * K2_TW, ti0 = blockIdx.y * K2_TH;
    for (int k = threadIdx.x; k < (K2_TH + 2) * (K2_TW + 2); k += 256) {
        int r = k / (K2_TW + 2), c = k % (K2_TW + 2);
        int gi = ti0 - 1 + r, gj = tj0 - 1 + c;
        float v = __int_as_float(0x7fc00000);
        if (gi >= 0 && gj >= 0 && gi < h && gj < w) {
            v = __ldg(dem + (long long)gi * stride + gj);
            if (has_ndv && v == ndv) v = __int_as_float(0x7fc00000);
        }
        t[r][c] = v;
    }
    __syncthreads();
    const int c = threadIdx.x & 63;
    for (int r = threadIdx.x >> 6; r < K2_TH; r += 4) {
        const int i = ti0 + r, j = tj0 + c;
        if (i >= h || j >= w) continue;
        float s = DCR_GDALDEM_NDV, a = DCR_GDALDEM_NDV;
        if (i > 0 && j > 0 && i < h - 1 && j < w - 1) {
            float sv, av;
            if (horn(t[r][c], t[r][c + 1], t[r][c + 2], t[r + 1][c], t[r + 1][c + 1], t[r + 1][c + 2], t[r + 2][c],
                     t[r + 2][c + 1], t[r + 2][c + 2], ewres, nsres, slope != nullptr, aspect != nullptr, &sv, &av)) {
                if (slope) s = sv;
                if (aspect) a = isnan(av) ? DCR_GDALDEM_NDV : av;
            }
        }
        if (slope) slope[(long long)i * w + j] = s;
        if (aspect) aspect[(long long)i * w + j] = a;
    }
}

extern "C" int dcr_horn_slope_aspect(const float* dem, int h, int w, int64_t stride, double ndv, int has_ndv,
                                     double ewres, double nsres, float* slope, float* aspect, void* stream) {
    DCR_ARG(dem && (slope || aspect), "null pointer");
    DCR_ARG(h > 0 && w > 0, "empty raster");
    dim3 grid((w + K2_TW - 1) / K2_TW, (h + K2_TH - 1) / K2_TH);
    horn_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dem, h, w, stride, (float)ndv, has_ndv, ewres, nsres, slope, aspect);
    DCR_LAUNCH_OK("horn_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------------
// Fused front end: warp(ref), warp(src), dh, masks, Horn(slope, aspect of warped src)
// ------------------------------------------------------------------------------------------
#define F_TW 128
#define F_TH 28
#define F_SRC_R (F_TH + 5)
#define F_SRC_C (F_TW + 8)   // 133 columns needed; 136 = multiple of 4 floats (TMA box rows are multiples of 16 bytes)
#define F_REF_R (F_TH + 3)
#define F_REF_C (F_TW + 8)   // 131 columns needed (+ up to 3 of TMA start alignment); 136
#define F_MSK_C (F_TW + 16)  // mask tile pitch: 128 columns + up to 15 bytes of TMA start alignment
#define F_SW_R (F_TH + 2)
#define F_SW_C (F_TW + 2)
#define F_ROWS_PER_THREAD (F_TH / 2)
#define F_SRC_WORDS (((F_SRC_R * F_SRC_C + 31) / 32) * 32)
#define F_REF_WORDS (((F_REF_R * F_REF_C + 31) / 32) * 32)

struct FrontParams {
    alignas(64) CUtensorMap tm_src;   // TMA descriptors of the two rasters (valid iff use_tma)
    alignas(64) CUtensorMap tm_ref;
    alignas(64) CUtensorMap tm_msk;   // uint8 static mask (valid iff use_tma_mask)
    int use_tma;
    int tma_dx_src, tma_dx_ref;       // TMA boxes must start on a 16-byte column: extra leading columns (0..3)
    int use_tma_mask, tma_dx_msk;     // same for the mask tile (0..15 leading bytes)
    DevGeom rg, sg;
    const float* ref;
    const float* src;
    const unsigned char* smask;
    long long smask_stride;
    int smask_h, smask_w, smask_nx0, smask_ny0;
    int h, w;
    int row_begin, row_end;  // output rows computed by this launch (row-band mode; whole grid otherwise)
    int smask_row0, smask_row1;  // rows of the static mask held in memory
    double ewres, nsres;
    float inv8ew, inv8ns;
    float ref_gm_lo, ref_gm_hi, src_gm_lo, src_gm_hi;  // ds_getma masked_values: masked iff lo <= x <= hi (float32 bounds of |x-ndv| <= tol)
    float ref_fill, src_fill;                               // value written to ref_w/src_w for invalid pixels
    float max_dz;
    int use_max_dz;
    float slope_lo, slope_hi;
    float slope_mid, slope_half, slope_tol;   // guard band of the FP32 Horn path around both slope limits
    float* dh;
    float* slope;
    float* aspect;
    unsigned char* maskbits;
    float* ref_w;
    float* src_w;
    unsigned short* bin;       // optional aspect-bin code per pixel (see the Horn loop), or nullptr
    unsigned long long* counters;  // [0]=count_base [1]=count_maxdz [2]=count_slope
    // loop mode (the native dem_align loop): the slope raster holds tan(slope) = |gradient| / 8 instead of degrees (the
    // consumers -- y = dh / tan(slope), the slope median -- want the tangent; saves an atan here and a tan there)
    int loopmode;
    float g_lo, g_hi, g_mid, g_half, g_tol;   // slope limits as tangents + guard band of the FP32 path around them
    // fast path: front_fast_kernel handles the interior tiles without nodata and appends every other tile to `list`
    // ([0] = count, then (ty << 16 | tx)); the general kernel then runs over that list
    unsigned* list;
    int list_mode;             // general kernel: 1 = tiles come from `list`
    float max_dz_eff;          // max_dz, or +inf when the filter is off (branch-free test)
};

// 16-tap warp of one pixel out of a shared-memory footprint tile (taps start at (r0, c0))
template <int PITCH>
__device__ __forceinline__ float warp_px_tile(const float* t, int r0, int c0, const DevGeom& g) {
    double v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float* q = t + (r0 + r) * PITCH + c0;
        v[r] = xpass(g.wx, q[0], q[1], q[2], q[3]);
    }
    double o = ypass(g.wy, v[0], v[1], v[2], v[3]);
    const float nearest = t[(r0 + 1 + (g.ny0 - g.iy0)) * PITCH + (c0 + 1 + (g.nx0 - g.ix0))];
    if (isnan(o)) {
        return bilinear_fallback(t[(r0 + 1) * PITCH + c0 + 1], t[(r0 + 1) * PITCH + c0 + 2], t[(r0 + 2) * PITCH + c0 + 2],
                                 t[(r0 + 2) * PITCH + c0 + 1], nearest, g.fx, g.fy);
    }
    if (isnan(nearest)) return nearest;
    return (float)o;
}

// ---- TMA (cp.async.bulk.tensor) staging of a 2-D footprint: one elected thread issues the copy, the bytes land in
// shared memory asynchronously and complete on an mbarrier; out-of-bounds elements are filled with NaN by the copy
// engine (CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA), which is exactly the "outside the raster" encoding.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int x, int y, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// Stage an R x C footprint whose top-left source pixel is (r_base, c_base) into shared memory: one warp per
// tile row, two rows (10 loads) in flight per lane, column predicates hoisted out of the row loop.
template <int R, int C, bool HAS_DZ>
__device__ __forceinline__ void stage_tile(const float* __restrict__ src, const DevGeom& g, int r_base, int c_base,
                                           float* __restrict__ tile, int tid) {
    const int lane = tid & 31, wrp = tid >> 5;
    const float qnan = __int_as_float(0x7fc00000);
    constexpr int NQ = (C + 31) / 32;
    unsigned cmask = 0;   // bit q: column chunk q of this lane is inside the tile and the raster
    unsigned smask = 0;   // bit q: inside the tile (needs a store)
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int cc = q * 32 + lane, gc = c_base + cc;
        if (cc < C) {
            smask |= 1u << q;
            if (gc >= 0 && gc < g.ws) cmask |= 1u << q;
        }
    }
    const float* colp = src + (c_base + lane);
    const bool has_ndv = g.has_ndv != 0;
    const float ndv = g.ndv;
    const int ndz = g.ndz;
    for (int r = wrp; r < R; r += 16) {
        float v[2][NQ];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int rr = r + 8 * u, gr = r_base + rr;
            const bool rok = rr < R && gr >= g.row0 && gr < g.row1;
            const float* rowp = colp + (long long)(gr - g.row0) * g.stride;
#pragma unroll
            for (int q = 0; q < NQ; ++q) v[u][q] = (rok && (cmask >> q & 1u)) ? __ldg(rowp + q * 32) : qnan;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int rr = r + 8 * u;
            if (rr < R) {
                float* trow = tile + rr * C + lane;
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    float x = v[u][q];
                    // deferred z shifts: a = b_getma(band); a = a + dz; filled() -- one float32 add per pending step
                    if (HAS_DZ)
                        for (int k = 0; k < ndz; ++k) x = (x >= g.zlo && x <= g.zhi) ? ndv : x + g.dz[k];
                    if (smask >> q & 1u) trow[q * 32] = (has_ndv && x == ndv) ? qnan : x;
                }
            }
        }
    }
}

// `it`: ordinal of the tile within this CTA (the mbarrier is initialised once and its phase parity alternates)
__device__ __forceinline__ void front_tile(const FrontParams& p, float* smem, const int bx, const int by, const unsigned it) {
    float* s_src_base = smem;                             // F_SRC_R x F_SRC_C
    float* s_ref_base = s_src_base + F_SRC_WORDS;         // F_REF_R x F_REF_C (128-byte aligned: TMA destination)
    float* s_sw = s_ref_base + F_REF_WORDS;               // F_SW_R x F_SW_C : warped src (NaN = nodata)
    // tile views: with TMA the box starts on the 16-byte column at or before the first needed one
    float* s_src = s_src_base + (p.use_tma ? p.tma_dx_src : 0);
    float* s_ref = s_ref_base + (p.use_tma ? p.tma_dx_ref : 0);
    __shared__ unsigned long long s_cnt[3];
    __shared__ __align__(128) unsigned char s_mask_base[F_TH * F_MSK_C];   // static (stable-surface) mask tile, 1 = masked
    unsigned char* s_mask = s_mask_base + (p.use_tma_mask ? p.tma_dx_msk : 0);

    const int tj0 = bx * F_TW, ti0 = p.row_begin + by * F_TH;
    const int tid = threadIdx.x;
    if (tid < 3) s_cnt[tid] = 0ull;

    // ---- stage the two source footprints (nodata / out-of-bounds -> NaN)
    __shared__ __align__(8) unsigned long long s_bar;
    if (p.use_tma) {
        // both footprints by TMA: thread 0 arms the mbarrier with the byte count and issues the two tensor copies
        if (tid == 0 && it == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&s_bar, (unsigned)((F_SRC_R * F_SRC_C + F_REF_R * F_REF_C) * sizeof(float)) +
                                       (p.use_tma_mask ? (unsigned)(F_TH * F_MSK_C) : 0u));
            if (p.use_tma_mask)
                tma_load_2d(s_mask_base, &p.tm_msk, tj0 + p.smask_nx0 - p.tma_dx_msk, ti0 + p.smask_ny0 - p.smask_row0, &s_bar);
            tma_load_2d(s_src_base, &p.tm_src, tj0 - 2 + p.sg.ix0 - p.tma_dx_src, ti0 - 2 + p.sg.iy0 - p.sg.row0, &s_bar);
            tma_load_2d(s_ref_base, &p.tm_ref, tj0 - 1 + p.rg.ix0 - p.tma_dx_ref, ti0 - 1 + p.rg.iy0 - p.rg.row0, &s_bar);
        }
    } else {
        if (p.sg.ndz) stage_tile<F_SRC_R, F_SRC_C, true>(p.src, p.sg, ti0 - 2 + p.sg.iy0, tj0 - 2 + p.sg.ix0, s_src, tid);
        else stage_tile<F_SRC_R, F_SRC_C, false>(p.src, p.sg, ti0 - 2 + p.sg.iy0, tj0 - 2 + p.sg.ix0, s_src, tid);
        stage_tile<F_REF_R, F_REF_C, false>(p.ref, p.rg, ti0 - 1 + p.rg.iy0, tj0 - 1 + p.rg.ix0, s_ref, tid);
    }
    if (p.smask && !p.use_tma_mask) {
        // nearest-neighbour sample of the world-fixed mask; all 16 byte loads of a thread are issued before their
        // first use (a dependent load per pixel inside the main pass cost ~600 cycles of latency per row)
        unsigned char mv[(F_TH * F_TW) / 256];
#pragma unroll
        for (int k = 0; k < (F_TH * F_TW) / 256; ++k) {
            const int e = k * 256 + tid;
            const int mi = ti0 + (e >> 7) + p.smask_ny0, mj = tj0 + (e & (F_TW - 1)) + p.smask_nx0;
            mv[k] = 1;
            if (mi >= p.smask_row0 && mj >= 0 && mi < p.smask_row1 && mj < p.smask_w)
                mv[k] = __ldg(p.smask + (long long)(mi - p.smask_row0) * p.smask_stride + mj);
        }
#pragma unroll
        for (int k = 0; k < (F_TH * F_TW) / 256; ++k) s_mask[((k * 256 + tid) >> 7) * F_MSK_C + ((k * 256 + tid) & (F_TW - 1))] = mv[k];
    }
    if (p.use_tma) {
        // wait for the bytes (bounded spin: a mis-programmed copy must fault, not hang the GPU), then nodata -> NaN
        unsigned spins = 0;
        while (!mbar_try_wait(&s_bar, it & 1u)) {
            if (++spins > (1u << 26)) __trap();
        }
        if (p.use_tma_mask) {
            // the copy engine fills bytes outside the mask raster with 0; outside = masked (1): edge tiles only
            const int mi0 = ti0 + p.smask_ny0, mj0 = tj0 + p.smask_nx0;
            if (mi0 < p.smask_row0 || mj0 < 0 || mi0 + F_TH > p.smask_row1 || mj0 + F_TW > p.smask_w) {
                for (int e = tid; e < F_TH * F_TW; e += 256) {
                    const int mi = mi0 + (e >> 7), mj = mj0 + (e & (F_TW - 1));
                    if (mi < p.smask_row0 || mj < 0 || mi >= p.smask_row1 || mj >= p.smask_w)
                        s_mask[(e >> 7) * F_MSK_C + (e & (F_TW - 1))] = 1;
                }
            }
        }
        // fix-up pass over the landed tiles, 4 elements per access (tiles are 128-byte aligned, sizes multiples of 4)
        const float qnan = __int_as_float(0x7fc00000);
        static_assert((F_SRC_R * F_SRC_C) % 4 == 0 && (F_REF_R * F_REF_C) % 4 == 0, "tile sizes must be float4 multiples");
        if (p.sg.ndz) {
            // deferred z shifts (coreglib.apply_z_shift not yet written back): a = b_getma(band); a = a + dz; filled()
            // -- one float32 add per pending step (<= 4 on this path) on every loaded pixel, then nodata -> NaN
            const float nd = p.sg.ndv, zlo = p.sg.zlo, zhi = p.sg.zhi;
            const bool hn = p.sg.has_ndv != 0;
            const int ndz = p.sg.ndz;
            const float z0 = p.sg.dz[0], z1 = p.sg.dz[1], z2 = p.sg.dz[2], z3 = p.sg.dz[3];
            auto fix = [&](float x) {
                x = (x >= zlo && x <= zhi) ? nd : x + z0;
                if (ndz > 1) x = (x >= zlo && x <= zhi) ? nd : x + z1;
                if (ndz > 2) x = (x >= zlo && x <= zhi) ? nd : x + z2;
                if (ndz > 3) x = (x >= zlo && x <= zhi) ? nd : x + z3;
                return (hn && x == nd) ? qnan : x;
            };
            float4* t4 = reinterpret_cast<float4*>(s_src_base);
            for (int k = tid; k < (F_SRC_R * F_SRC_C) / 4; k += 256) {
                float4 v = t4[k];
                v.x = fix(v.x); v.y = fix(v.y); v.z = fix(v.z); v.w = fix(v.w);
                t4[k] = v;
            }
        } else if (p.sg.has_ndv) {
            const float nd = p.sg.ndv;
            float4* t4 = reinterpret_cast<float4*>(s_src_base);
            for (int k = tid; k < (F_SRC_R * F_SRC_C) / 4; k += 256) {
                float4 v = t4[k];
                if (v.x == nd || v.y == nd || v.z == nd || v.w == nd) {   // rare: nodata in this quad
                    v.x = v.x == nd ? qnan : v.x; v.y = v.y == nd ? qnan : v.y;
                    v.z = v.z == nd ? qnan : v.z; v.w = v.w == nd ? qnan : v.w;
                    t4[k] = v;
                }
            }
        }
        if (p.rg.has_ndv) {
            const float nd = p.rg.ndv;
            float4* t4 = reinterpret_cast<float4*>(s_ref_base);
            for (int k = tid; k < (F_REF_R * F_REF_C) / 4; k += 256) {
                float4 v = t4[k];
                if (v.x == nd || v.y == nd || v.z == nd || v.w == nd) {
                    v.x = v.x == nd ? qnan : v.x; v.y = v.y == nd ? qnan : v.y;
                    v.z = v.z == nd ? qnan : v.z; v.w = v.w == nd ? qnan : v.w;
                    t4[k] = v;
                }
            }
        }
    }
    __syncthreads();

    // ---- halo ring of the warped-src tile (needed by the 3x3 stencil only)
    for (int k = tid; k < 2 * F_SW_C + 2 * F_TH; k += 256) {
        int rr, cc;
        if (k < F_SW_C) { rr = 0; cc = k; }
        else if (k < 2 * F_SW_C) { rr = F_SW_R - 1; cc = k - F_SW_C; }
        else if (k < 2 * F_SW_C + F_TH) { rr = 1 + (k - 2 * F_SW_C); cc = 0; }
        else { rr = 1 + (k - 2 * F_SW_C - F_TH); cc = F_SW_C - 1; }
        const int gi = ti0 - 1 + rr, gj = tj0 - 1 + cc;
        float v = __int_as_float(0x7fc00000);
        if (gi >= 0 && gj >= 0 && gi < p.h && gj < p.w) v = warp_px_tile<F_SRC_C>(s_src, rr, cc, p.sg);
        s_sw[rr * F_SW_C + cc] = v;
    }

    // ---- main pass: thread = (column c, half hf), 16 consecutive rows, sliding x-pass window.
    //      Partial unrolling only: the fully unrolled kernel was 137 KB of SASS and stalled on
    //      instruction fetch (profiles/front_kernel_r1b: stall_no_instruction 4.5 per issue).
    const int c = tid & (F_TW - 1);
    const int hf = tid >> 7;
    const int rbeg = hf * F_ROWS_PER_THREAD;
    const int j = tj0 + c;
    unsigned valid_bits = 0;  // bit rr: base-valid (ref & src unmasked, static mask clear)
    unsigned maxdz_bits = 0;  // bit rr: |dh| > max_dz
    unsigned n_base = 0, n_maxdz = 0, n_slope = 0;
    {
        double xs0, xs1, xs2, xr0, xr1, xr2;
        {
            const float* q = s_src + (rbeg + 1) * F_SRC_C + (c + 1);
            xs0 = xpass(p.sg.wx, q[0], q[1], q[2], q[3]);
            q += F_SRC_C;
            xs1 = xpass(p.sg.wx, q[0], q[1], q[2], q[3]);
            q += F_SRC_C;
            xs2 = xpass(p.sg.wx, q[0], q[1], q[2], q[3]);
            const float* u = s_ref + rbeg * F_REF_C + c;
            xr0 = xpass(p.rg.wx, u[0], u[1], u[2], u[3]);
            u += F_REF_C;
            xr1 = xpass(p.rg.wx, u[0], u[1], u[2], u[3]);
            u += F_REF_C;
            xr2 = xpass(p.rg.wx, u[0], u[1], u[2], u[3]);
        }
        const int s_near = (p.sg.ny0 - p.sg.iy0) * F_SRC_C + (p.sg.nx0 - p.sg.ix0);
        const int r_near = (p.rg.ny0 - p.rg.iy0) * F_REF_C + (p.rg.nx0 - p.rg.ix0);
        // per-thread invariants hoisted out of the row loop: column validity, row limits, output offset
        const bool col_ok = j < p.w;
        const int i0 = ti0 + rbeg;
        const int n_store = col_ok ? min(F_ROWS_PER_THREAD, p.row_end - i0) : 0;   // rows rr < n_store are stored
        const int n_grid = col_ok ? min(F_ROWS_PER_THREAD, p.h - i0) : 0;          // rows rr < n_grid exist in the grid
        long long o = (long long)(i0 - p.row_begin) * p.w + j;
        const float* qs = s_src + (rbeg + 4) * F_SRC_C + (c + 1);
        const float* qr = s_ref + (rbeg + 3) * F_REF_C + c;
        float* swp = s_sw + (rbeg + 1) * F_SW_C + (c + 1);
        const unsigned char* mk = s_mask + rbeg * F_MSK_C + c;
        const bool has_mask = p.smask != nullptr;
        const bool has_warped = (p.ref_w != nullptr) || (p.src_w != nullptr);
        const float qnan = __int_as_float(0x7fc00000);
#pragma unroll 4
        for (int rr = 0; rr < F_ROWS_PER_THREAD; ++rr) {
            const double xs3 = xpass(p.sg.wx, qs[0], qs[1], qs[2], qs[3]);
            const double xr3 = xpass(p.rg.wx, qr[0], qr[1], qr[2], qr[3]);
            const double so = ypass(p.sg.wy, xs0, xs1, xs2, xs3);
            const double ro = ypass(p.rg.wy, xr0, xr1, xr2, xr3);
            xs0 = xs1; xs1 = xs2; xs2 = xs3;
            xr0 = xr1; xr1 = xr2; xr2 = xr3;
            // common case: every tap valid -> the cubic value; anything else (nodata / edge) takes ONE rare branch
            float sv = (float)so, rv = (float)ro;
            if (isnan(sv + rv)) {
                const float* ts = qs - 2 * F_SRC_C + 1;  // UL pixel of the 2x2
                const float* tr = qr - 2 * F_REF_C + 1;
                const float ns = ts[s_near], nr = tr[r_near];
                sv = isnan(so) ? bilinear_fallback(ts[0], ts[1], ts[F_SRC_C + 1], ts[F_SRC_C], ns, p.sg.fx, p.sg.fy)
                               : (isnan(ns) ? ns : sv);
                rv = isnan(ro) ? bilinear_fallback(tr[0], tr[1], tr[F_REF_C + 1], tr[F_REF_C], nr, p.rg.fx, p.rg.fy)
                               : (isnan(nr) ? nr : rv);
            }
            // rows past this launch's band still feed the 3x3 stencil of its last row (row-band mode)
            if (rr >= n_grid) sv = qnan;
            *swp = sv;
            if (rr < n_store) {
                // ds_getma masks as float32 intervals of numpy masked_values' |x - ndv| <= tol (NaN -> masked)
                bool base_ok = (sv < p.src_gm_lo || sv > p.src_gm_hi) && (rv < p.ref_gm_lo || rv > p.ref_gm_hi);
                if (has_mask && *mk) base_ok = false;
                const float dh = base_ok ? (sv - rv) : 0.0f;
                if (base_ok) {
                    valid_bits |= (1u << rr);
                    ++n_base;
                    if (p.use_max_dz && (dh < -p.max_dz || dh > p.max_dz)) maxdz_bits |= (1u << rr);
                    else ++n_maxdz;
                }
                p.dh[o] = dh;
                if (has_warped) {
                    if (p.ref_w) p.ref_w[o] = isnan(rv) ? p.ref_fill : rv;
                    if (p.src_w) p.src_w[o] = isnan(sv) ? p.src_fill : sv;
                }
            }
            o += p.w;
            qs += F_SRC_C;
            qr += F_REF_C;
            swp += F_SW_C;
            mk += F_MSK_C;
        }
    }
    __syncthreads();

    // ---- Horn stencil on the warped src + mask byte
    {
        float w0, w1, w2, w3, w4, w5, w6, w7, w8;
        const float* q = s_sw + rbeg * F_SW_C + c;
        w0 = q[0]; w1 = q[1]; w2 = q[2];
        q += F_SW_C;
        w3 = q[0]; w4 = q[1]; w5 = q[2];
        q += F_SW_C;
        float rs0 = (w0 + w1 + w1 + w2), rs1 = (w3 + w4 + w4 + w5);   // row sums of the window's top / middle row
        // hoisted invariants: stored rows, interior rows [r_lo, r_hi) of this thread (borders are nodata), offset
        const bool col_ok = j < p.w;
        const bool col_in = j > 0 && j < p.w - 1;
        const int i0 = ti0 + rbeg;
        const int n_store = col_ok ? min(F_ROWS_PER_THREAD, p.row_end - i0) : 0;
        const int r_lo = col_in ? max(0, 1 - i0) : F_ROWS_PER_THREAD;
        const int r_hi = col_in ? min(F_ROWS_PER_THREAD, p.h - 1 - i0) : 0;
        long long o = (long long)(i0 - p.row_begin) * p.w + j;
#pragma unroll 2
        for (int rr = 0; rr < F_ROWS_PER_THREAD; ++rr) {
            w6 = q[0]; w7 = q[1]; w8 = q[2];
            q += F_SW_C;
            const float rs2 = (w6 + w7 + w7 + w8);
            if (rr < n_store) {
                float sl = DCR_GDALDEM_NDV, as = DCR_GDALDEM_NDV, sg = DCR_GDALDEM_NDV;
                unsigned bits = DCR_M_SLOPE | DCR_M_ASPECT;
                if (rr >= r_lo && rr < r_hi) {
                    float sv, av, gv;
                    bool ex;
                    if (horn_fast(w0, w1, w2, w3, w4, w5, w6, w7, w8, rs2, rs0, p.inv8ew, p.inv8ns, p.slope_mid, p.slope_half,
                                  p.slope_tol, &sv, &av, &ex, &gv)) {
                        if (ex) horn(w0, w1, w2, w3, w4, w5, w6, w7, w8, p.ewres, p.nsres, true, true, &sv, &av, &gv);
                        sl = sv;
                        sg = gv;
                        // range_fltr(slope, slope_lim): masked_outside keeps lo <= s <= hi
                        if (!(sl < p.slope_lo) && !(sl > p.slope_hi)) { bits &= ~DCR_M_SLOPE; ++n_slope; }
                        if (!isnan(av)) { as = av; bits &= ~DCR_M_ASPECT; }
                    }
                }
                if (!(valid_bits & (1u << rr))) bits |= DCR_M_BASE;
                else if (maxdz_bits & (1u << rr)) bits |= DCR_M_MAXDZ;
                p.slope[o] = p.loopmode ? sg : sl;
                if (p.aspect) p.aspect[o] = as;
                p.maskbits[o] = (unsigned char)bits;
                if (p.bin) {
                    // aspect bin of the Nuth-Kaab sample (coreglib.py:246-262: floor(aspect), 360 joins the last bin) for the
                    // pixels inside every mask known here; only the MAD filter (decided later) can still remove them
                    int bi = (int)as;
                    p.bin[o] = (unsigned short)(bits == 0u ? (bi > DCR_NBINS - 1 ? DCR_NBINS - 1 : bi) : DCR_BIN_INVALID);
                }
            }
            o += p.w;
            w0 = w3; w1 = w4; w2 = w5;
            w3 = w6; w4 = w7; w5 = w8;
            rs0 = rs1; rs1 = rs2;
        }
    }
    // block-level counter reduction -> 3 global integer atomics per CTA
    n_base = __reduce_add_sync(0xffffffffu, n_base);
    n_maxdz = __reduce_add_sync(0xffffffffu, n_maxdz);
    n_slope = __reduce_add_sync(0xffffffffu, n_slope);
    if ((tid & 31) == 0) {
        atomicAdd(&s_cnt[0], (unsigned long long)n_base);
        atomicAdd(&s_cnt[1], (unsigned long long)n_maxdz);
        atomicAdd(&s_cnt[2], (unsigned long long)n_slope);
    }
    __syncthreads();
    if (tid < 3 && p.counters) atomicAdd(&p.counters[tid], s_cnt[tid]);
}

__global__ void __launch_bounds__(256, 4) front_kernel(const __grid_constant__ FrontParams p) {
    extern __shared__ __align__(128) float smem[];
    if (!p.list_mode) {
        front_tile(p, smem, blockIdx.x, blockIdx.y, 0u);
        return;
    }
    // list mode: the tiles the fast kernel declined (grid border, nodata in a footprint), a few per CTA
    const unsigned n = p.list[0];
    unsigned it = 0;
    for (unsigned t = blockIdx.x; t < n; t += gridDim.x, ++it) {
        const unsigned id = p.list[1 + t];
        front_tile(p, smem, (int)(id & 0xffffu), (int)(id >> 16), it);
        __syncthreads();   // shared memory (tiles, mbarrier, counters) is reused by the next tile
    }
}

// ------------------------------------------------------------------------------------------
// Fast path of the fused front end: interior tiles whose footprints hold no nodata.
//
// The general kernel above is issue-bound at ~345 thread-instructions per pixel, a third of them control flow and
// integer bookkeeping for cases that almost never occur inside a raster (grid borders, nodata taps and the bilinear
// fallback, row bands, partial tiles, optional outputs).  This kernel handles ONLY the common tile -- every output pixel of
// the tile and of its halo ring inside the grid, every staged source pixel valid -- with straight-line code:
//   * the axes of the translation on which a raster is pixel-aligned are known at launch (in the dem_align loop each axis
//     has exactly one resampled raster): AX selects the 4-tap FP64 pass only where the offset has a fraction (bit 0/1: source
//     x/y, bit 2/3: reference x/y); an aligned axis contributes its single tap, which is what GDAL's weights (0, 1, 0, 0)
//     return for finite taps;
//   * no NaN probes, no bilinear fallback, no row / column range checks; masks and counters are bit operations;
//   * LOOPMODE: the slope raster receives tan(slope) (no atan), no aspect raster.
// Every other tile (grid border, nodata or out-of-raster taps in a footprint) is appended to p.list and left to the general
// kernel.  Results are bit-identical to the general kernel's (same arithmetic, same order).
// ------------------------------------------------------------------------------------------
template <bool FX>
__device__ __forceinline__ double fast_xrow(const double* __restrict__ w, const float* __restrict__ q) {
    if (FX) return w[0] * (double)q[0] + w[1] * (double)q[1] + w[2] * (double)q[2] + w[3] * (double)q[3];
    return (double)q[1];
}

// warped value of the pixel whose 4x4 taps start at (r0, c0) of the source footprint tile, all taps finite
template <bool FX, bool FY>
__device__ __forceinline__ float fast_px(const float* __restrict__ t, int r0, int c0, const DevGeom& g) {
    const float* q = t + r0 * F_SRC_C + c0;
    if (FY) {
        const double v0 = fast_xrow<FX>(g.wx, q), v1 = fast_xrow<FX>(g.wx, q + F_SRC_C),
                     v2 = fast_xrow<FX>(g.wx, q + 2 * F_SRC_C), v3 = fast_xrow<FX>(g.wx, q + 3 * F_SRC_C);
        return (float)ypass(g.wy, v0, v1, v2, v3);
    }
    return (float)fast_xrow<FX>(g.wx, q + F_SRC_C);
}

// exact Horn slope / aspect of one window (GDAL's FP64 formulation) for the guard-band pixels of the fast path
__device__ __noinline__ void horn_exact(float w0, float w1, float w2, float w3, float w5, float w6, float w7, float w8,
                                        double ewres, double nsres, float* slope_deg, float* tan_slope, float* aspect) {
    const float a = (w0 + w3 + w3 + w6);
    const float b = (w2 + w5 + w5 + w8);
    const float c = (w6 + w7 + w7 + w8);
    const float d = (w0 + w1 + w1 + w2);
    const float dyf = c - d;
    {
        const float dxf = a - b;
        const double dx = (double)dxf / ewres, dy = (double)dyf / nsres;
        const double key = dx * dx + dy * dy;
        *slope_deg = (float)(atan(sqrt(key) / 8.0) * (180.0 / M_PI));
        *tan_slope = (float)(sqrt(key) / 8.0);
    }
    {
        const float dxf = b - a;
        const double dx = (double)dxf, dy = (double)dyf;
        if (dx == 0.0 && dy == 0.0) {
            *aspect = __int_as_float(0x7fc00000);
        } else {
            float asp = (float)(atan2(dy, -dx) / (M_PI / 180.0));
            asp = (asp > 90.0f) ? __fsub_rn(450.0f, asp) : __fsub_rn(90.0f, asp);
            if (asp == 360.0f) asp = 0.0f;
            *aspect = asp;
        }
    }
}

// Unroll factors of the two row loops: fully unrolled the kernel is 61 KB of SASS, twice the 32 KB instruction cache of an
// SM, and 36 % of the warp-stall samples were instruction fetches (profiles/r2/front_fast_r2c_sass.txt)
#ifndef DCR_FF_UNROLL_MAIN
#define DCR_FF_UNROLL_MAIN 14
#endif
#ifndef DCR_FF_UNROLL_HORN
#define DCR_FF_UNROLL_HORN 7
#endif
constexpr int FF_UNROLL_MAIN = DCR_FF_UNROLL_MAIN, FF_UNROLL_HORN = DCR_FF_UNROLL_HORN;
// MODE: 0 = slope in degrees + aspect raster, 1 = slope in degrees, no aspect raster, 2 = tan(slope), no aspect raster
template <int AX, int MODE>
__global__ void __launch_bounds__(256, 4) front_fast_kernel(const __grid_constant__ FrontParams p) {
    constexpr bool SX = (AX & 1) != 0, SY = (AX & 2) != 0, RX = (AX & 4) != 0, RY = (AX & 8) != 0;
    extern __shared__ __align__(128) float smem[];
    float* s_src_base = smem;                             // F_SRC_R x F_SRC_C
    float* s_ref_base = s_src_base + F_SRC_WORDS;         // F_REF_R x F_REF_C
    float* s_sw = s_ref_base + F_REF_WORDS;               // F_SW_R x F_SW_C : warped src
    const float* s_src = s_src_base + p.tma_dx_src;
    const float* s_ref = s_ref_base + p.tma_dx_ref;
    __shared__ __align__(128) unsigned char s_mask_base[F_TH * F_MSK_C];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ unsigned s_cnt[3];
    unsigned char* s_mask = s_mask_base + (p.use_tma_mask ? p.tma_dx_msk : 0);
    const int tid = threadIdx.x;
    const int tj0 = blockIdx.x * F_TW, ti0 = p.row_begin + blockIdx.y * F_TH;   // (row band: rows [row_begin, row_end) of the grid)
    // a tile whose pixels or halo ring touch the border of the grid, or that is cut by the end of the row band, is not ours
    const bool interior = blockIdx.x > 0 && ti0 > 0 && tj0 + F_TW < p.w && ti0 + F_TH < p.h && ti0 + F_TH <= p.row_end;
    if (!interior) {
        if (tid == 0) p.list[1u + atomicAdd(p.list, 1u)] = (blockIdx.y << 16) | blockIdx.x;
        return;
    }
    if (tid < 3) s_cnt[tid] = 0u;
    if (tid == 0) mbar_init(&s_bar, 1);
    if (!p.use_tma_mask)   // no static mask: an all-clear tile (the main pass reads it unconditionally)
        for (int k = tid; k < (F_TH * F_MSK_C) / 16; k += 256) reinterpret_cast<uint4*>(s_mask_base)[k] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&s_bar, (unsigned)((F_SRC_R * F_SRC_C + F_REF_R * F_REF_C) * sizeof(float)) +
                                   (p.use_tma_mask ? (unsigned)(F_TH * F_MSK_C) : 0u));
        if (p.use_tma_mask)
            tma_load_2d(s_mask_base, &p.tm_msk, tj0 + p.smask_nx0 - p.tma_dx_msk, ti0 + p.smask_ny0 - p.smask_row0, &s_bar);
        tma_load_2d(s_src_base, &p.tm_src, tj0 - 2 + p.sg.ix0 - p.tma_dx_src, ti0 - 2 + p.sg.iy0 - p.sg.row0, &s_bar);
        tma_load_2d(s_ref_base, &p.tm_ref, tj0 - 1 + p.rg.ix0 - p.tma_dx_ref, ti0 - 1 + p.rg.iy0 - p.rg.row0, &s_bar);
    }
    {
        unsigned spins = 0;
        while (!mbar_try_wait(&s_bar, 0)) {
            if (++spins > (1u << 26)) __trap();
        }
    }
    if (p.use_tma_mask) {
        // the copy engine fills bytes outside the mask raster with 0; outside = masked (1): tiles at the mask's edge only
        const int mi0 = ti0 + p.smask_ny0, mj0 = tj0 + p.smask_nx0;
        if (mi0 < p.smask_row0 || mj0 < 0 || mi0 + F_TH > p.smask_row1 || mj0 + F_TW > p.smask_w) {
            unsigned char* sm = s_mask_base + p.tma_dx_msk;
            for (int e = tid; e < F_TH * F_TW; e += 256) {
                const int mi = mi0 + (e >> 7), mj = mj0 + (e & (F_TW - 1));
                if (mi < p.smask_row0 || mj < 0 || mi >= p.smask_row1 || mj >= p.smask_w) sm[(e >> 7) * F_MSK_C + (e & (F_TW - 1))] = 1;
            }
        }
    }
    // pending z shifts of the source (coreglib.apply_z_shift deferred by the loop) + the scan for nodata / out-of-raster
    // (NaN-filled by the copy engine) pixels: any of them sends the whole tile to the general kernel
    bool bad = false;
    {
        const float snd = p.sg.has_ndv ? p.sg.ndv : __int_as_float(0x7fc00000);   // (x == NaN) is never true
        const float rnd = p.rg.has_ndv ? p.rg.ndv : __int_as_float(0x7fc00000);
        float4* t4 = reinterpret_cast<float4*>(s_src_base);
        if (p.sg.ndz) {
            const float nd = p.sg.ndv, zlo = p.sg.zlo, zhi = p.sg.zhi;
            const int ndz = p.sg.ndz;
            const float z0 = p.sg.dz[0], z1 = p.sg.dz[1], z2 = p.sg.dz[2], z3 = p.sg.dz[3];
            auto fix = [&](float x) {
                x = (x >= zlo && x <= zhi) ? nd : x + z0;
                if (ndz > 1) x = (x >= zlo && x <= zhi) ? nd : x + z1;
                if (ndz > 2) x = (x >= zlo && x <= zhi) ? nd : x + z2;
                if (ndz > 3) x = (x >= zlo && x <= zhi) ? nd : x + z3;
                return x;
            };
            for (int k = tid; k < (F_SRC_R * F_SRC_C) / 4; k += 256) {
                float4 v = t4[k];
                v.x = fix(v.x); v.y = fix(v.y); v.z = fix(v.z); v.w = fix(v.w);
                t4[k] = v;
                const float sum = (v.x + v.y) + (v.z + v.w);
                bad = bad || (sum != sum) || v.x == snd || v.y == snd || v.z == snd || v.w == snd;
            }
        } else {
            for (int k = tid; k < (F_SRC_R * F_SRC_C) / 4; k += 256) {
                const float4 v = t4[k];
                const float sum = (v.x + v.y) + (v.z + v.w);
                bad = bad || (sum != sum) || v.x == snd || v.y == snd || v.z == snd || v.w == snd;
            }
        }
        const float4* r4 = reinterpret_cast<const float4*>(s_ref_base);
        for (int k = tid; k < (F_REF_R * F_REF_C) / 4; k += 256) {
            const float4 v = r4[k];
            const float sum = (v.x + v.y) + (v.z + v.w);
            bad = bad || (sum != sum) || v.x == rnd || v.y == rnd || v.z == rnd || v.w == rnd;
        }
    }
    if (__syncthreads_or(bad ? 1 : 0)) {
        if (tid == 0) p.list[1u + atomicAdd(p.list, 1u)] = (blockIdx.y << 16) | blockIdx.x;
        return;
    }

    // ---- halo ring of the warped-src tile (3x3 stencil); every tap is finite here, so the lean warp applies
    for (int k = tid; k < 2 * F_SW_C + 2 * F_TH; k += 256) {
        int rr, cc;
        if (k < F_SW_C) { rr = 0; cc = k; }
        else if (k < 2 * F_SW_C) { rr = F_SW_R - 1; cc = k - F_SW_C; }
        else if (k < 2 * F_SW_C + F_TH) { rr = 1 + (k - 2 * F_SW_C); cc = 0; }
        else { rr = 1 + (k - 2 * F_SW_C - F_TH); cc = F_SW_C - 1; }
        s_sw[rr * F_SW_C + cc] = fast_px<SX, SY>(s_src, rr, cc, p.sg);
    }

    // ---- main pass: thread = (column c, half hf), F_ROWS_PER_THREAD consecutive rows
    const int c = tid & (F_TW - 1);
    const int hf = tid >> 7;
    const int rbeg = hf * F_ROWS_PER_THREAD;
    const long long o0 = (long long)(ti0 - p.row_begin + rbeg) * p.w + (tj0 + c);   // outputs hold the band's rows only
    const int w = p.w;
    unsigned n_base = 0, n_big = 0, n_slope = 0;
    // the thread's column of the mask tile: static-mask byte in, DCR_M_BASE / DCR_M_MAXDZ byte out (read back by the same
    // thread in the Horn pass: no bit bookkeeping in registers)
    unsigned char* const mk0 = s_mask + rbeg * F_MSK_C + c;
    {
        // tile row tr of s_src holds source row (output row - 2 + tr): the 4 taps of output row r are tr = r + 1 .. r + 4,
        // the aligned-axis tap is tr = r + 2; columns alike (tap 0 at c + 1).  s_ref: taps tr = r .. r + 3 (aligned: r + 1).
        const float* qs = s_src + (rbeg + 1) * F_SRC_C + (c + 1);
        const float* qr = s_ref + rbeg * F_REF_C + c;
        double xs0 = 0, xs1 = 0, xs2 = 0, xr0 = 0, xr1 = 0, xr2 = 0;
        if (SY) {
            xs0 = fast_xrow<SX>(p.sg.wx, qs);
            xs1 = fast_xrow<SX>(p.sg.wx, qs + F_SRC_C);
            xs2 = fast_xrow<SX>(p.sg.wx, qs + 2 * F_SRC_C);
        }
        if (RY) {
            xr0 = fast_xrow<RX>(p.rg.wx, qr);
            xr1 = fast_xrow<RX>(p.rg.wx, qr + F_REF_C);
            xr2 = fast_xrow<RX>(p.rg.wx, qr + 2 * F_REF_C);
        }
        qs += (SY ? 3 : 1) * F_SRC_C;
        qr += (RY ? 3 : 1) * F_REF_C;
        float* swp = s_sw + (rbeg + 1) * F_SW_C + (c + 1);
        unsigned char* mk = mk0;
        const float slo = p.src_gm_lo, shi = p.src_gm_hi, rlo = p.ref_gm_lo, rhi = p.ref_gm_hi, mdz = p.max_dz_eff;
        float* dhp = p.dh + o0;
#pragma unroll FF_UNROLL_MAIN
        for (int rr = 0; rr < F_ROWS_PER_THREAD; ++rr) {
            float sv, rv;
            if (SY) {
                const double xs3 = fast_xrow<SX>(p.sg.wx, qs);
                sv = (float)ypass(p.sg.wy, xs0, xs1, xs2, xs3);
                xs0 = xs1; xs1 = xs2; xs2 = xs3;
            } else {
                sv = (float)fast_xrow<SX>(p.sg.wx, qs);
            }
            if (RY) {
                const double xr3 = fast_xrow<RX>(p.rg.wx, qr);
                rv = (float)ypass(p.rg.wy, xr0, xr1, xr2, xr3);
                xr0 = xr1; xr1 = xr2; xr2 = xr3;
            } else {
                rv = (float)fast_xrow<RX>(p.rg.wx, qr);
            }
            *swp = sv;
            // ds_getma masks as float32 intervals of numpy masked_values' |x - ndv| <= tol; static mask byte
            const bool s_in = (sv >= slo) & (sv <= shi);
            const bool r_in = (rv >= rlo) & (rv <= rhi);
            const bool ok = !(s_in | r_in | (*mk != 0));
            const float dh = ok ? (sv - rv) : 0.0f;
            const bool big = fabsf(dh) > mdz;          // (dh = 0 where not ok)
            n_base += ok ? 1u : 0u;
            n_big += big ? 1u : 0u;
            *mk = (unsigned char)(ok ? (big ? DCR_M_MAXDZ : 0u) : DCR_M_BASE);
            *dhp = dh;
            qs += F_SRC_C; qr += F_REF_C; swp += F_SW_C; mk += F_MSK_C; dhp += w;
        }
    }
    __syncthreads();

    // ---- Horn stencil on the warped src + mask byte + aspect-bin code
    {
        const float* q = s_sw + rbeg * F_SW_C + c;
        float w0 = q[0], w1 = q[1], w2 = q[2];
        float w3 = q[F_SW_C], w4 = q[F_SW_C + 1], w5 = q[F_SW_C + 2];
        float rs0 = (w0 + w1 + w1 + w2), rs1 = (w3 + w4 + w4 + w5);
        q += 2 * F_SW_C;
        const float inv8ew = p.inv8ew, inv8ns = p.inv8ns;
        const float lim_mid = MODE == 2 ? p.g_mid : p.slope_mid, lim_half = MODE == 2 ? p.g_half : p.slope_half,
                    lim_tol = MODE == 2 ? p.g_tol : p.slope_tol, lim_lo = MODE == 2 ? p.g_lo : p.slope_lo,
                    lim_hi = MODE == 2 ? p.g_hi : p.slope_hi;
        float* slp = p.slope + o0;
        unsigned char* mbp = p.maskbits + o0;
        unsigned short* bnp = p.bin + o0;
        float* asp = MODE == 0 ? p.aspect + o0 : nullptr;
        const unsigned char* mk = mk0;
#pragma unroll FF_UNROLL_HORN
        for (int rr = 0; rr < F_ROWS_PER_THREAD; ++rr) {
            const float w6 = q[0], w7 = q[1], w8 = q[2];
            const float rs2 = (w6 + w7 + w7 + w8);
            const float a = (w0 + w3 + w3 + w6);
            const float b = (w2 + w5 + w5 + w8);
            const float dxs = a - b;      // slope numerator; the aspect uses b - a = -dxs exactly
            const float dyf = rs2 - rs0;
            const float gx = dxs * inv8ew, gy = dyf * inv8ns;
            float g;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(gx * gx + gy * gy));
            // sl: the slope as this mode stores it -- tan(slope) (MODE 2) or degrees; the limits are in the same unit and
            // within the guard band of either one the FP64 path decides
            float sl = g;
            if (MODE != 2) {
                const bool bigg = g > 1.0f;
                float sa = atan01(bigg ? __fdividef(1.0f, g) : g);
                if (bigg) sa = 1.5707963267948966f - sa;
                sl = sa * 57.29577951308232f;
            }
            bool ex = !(fabsf(fabsf(sl - lim_mid) - lim_half) > lim_tol);
            bool s_ok = !(sl < lim_lo) & !(sl > lim_hi);
            // aspect: atan2(dyf, dxs) from the [0, 1] polynomial, GDAL's azimuth convention
            const float ax = fabsf(dxs), ay = fabsf(dyf);
            float an = atan01(__fdividef(fminf(ax, ay), fmaxf(ax, ay)));
            if (ay > ax) an = 1.5707963267948966f - an;
            if (dxs < 0.0f) an = 3.141592653589793f - an;
            an = copysignf(an, dyf);
            float af = an * 57.29577951308232f;
            af = (af > 90.0f) ? (450.0f - af) : (90.0f - af);
            const bool flat = (dxs == 0.0f) & (dyf == 0.0f);
            // within 3e-4 deg of an integer degree (a bin edge) -> exact path; a flat window has no aspect at all
            ex = (ex | !(fabsf((af - floorf(af)) - 0.5f) <= 0.5f - 3e-4f)) & !flat;
            if (ex) {   // ~0.06 % of the pixels
                float sle, ge, ae;
                horn_exact(w0, w1, w2, w3, w5, w6, w7, w8, p.ewres, p.nsres, &sle, &ge, &ae);
                s_ok = !(sle < p.slope_lo) & !(sle > p.slope_hi);
                sl = MODE == 2 ? ge : sle;
                af = ae;
            }
            const unsigned bits = (unsigned)*mk | (s_ok ? 0u : (unsigned)DCR_M_SLOPE) | (flat ? (unsigned)DCR_M_ASPECT : 0u);
            n_slope += s_ok ? 1u : 0u;
            const int bi = min((int)af, DCR_NBINS - 1);   // (NaN / flat never reaches a valid code: bits != 0)
            *slp = sl;
            *mbp = (unsigned char)bits;
            *bnp = (unsigned short)(bits == 0u ? bi : (int)DCR_BIN_INVALID);
            if (MODE == 0) { *asp = flat ? DCR_GDALDEM_NDV : af; asp += w; }
            slp += w; mbp += w; bnp += w; mk += F_MSK_C; q += F_SW_C;
            w0 = w3; w1 = w4; w2 = w5;
            w3 = w6; w4 = w7; w5 = w8;
            rs0 = rs1; rs1 = rs2;
        }
    }
    // counters: base-valid, after max_dz, slope-valid
    n_base = __reduce_add_sync(0xffffffffu, n_base);
    n_big = __reduce_add_sync(0xffffffffu, n_big);
    n_slope = __reduce_add_sync(0xffffffffu, n_slope);
    if ((tid & 31) == 0) {
        atomicAdd(&s_cnt[0], n_base);
        atomicAdd(&s_cnt[1], n_base - n_big);
        atomicAdd(&s_cnt[2], n_slope);
    }
    __syncthreads();
    if (tid < 3 && p.counters) atomicAdd(&p.counters[tid], (unsigned long long)s_cnt[tid]);
}

// float32 interval equivalent to numpy masked_values' |x - ndv| <= 1e-8 + 1e-5*|ndv| evaluated in float64
static void gm_bounds(double ndv, float* lo, float* hi) {
    const double tol = 1e-8 + 1e-5 * fabs(ndv);
    const double dl = ndv - tol, dh = ndv + tol;
    float fl = (float)dl, fh = (float)dh;
    if ((double)fl < dl) fl = nextafterf(fl, INFINITY);
    if ((double)fh > dh) fh = nextafterf(fh, -INFINITY);
    *lo = fl;
    *hi = fh;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
    }
    return fn;
}
// 2-D tensor map (float32, OOB -> NaN; or uint8, OOB -> 0) over the rows held in memory; box = (box_w x box_h)
// (the descriptors of the last few rasters are kept per thread: the iterations of a loop stage the same three arrays, and
// an encode costs microseconds of host time between two iterations)
struct TmapKey { const void* base; int rows, cols, box_w, box_h, u8; long long stride; };
struct TmapSlot { TmapKey k; CUtensorMap tm; bool used; };
static bool make_tmap(CUtensorMap* tm, const void* base, int rows, int cols, long long stride_elems, int box_w, int box_h,
                      bool u8 = false) {
    static thread_local TmapSlot cache[8];
    static thread_local unsigned next = 0;
    const TmapKey key = {base, rows, cols, box_w, box_h, u8 ? 1 : 0, stride_elems};
    for (int q = 0; q < 8; ++q)
        if (cache[q].used && memcmp(&cache[q].k, &key, sizeof(key)) == 0) { *tm = cache[q].tm; return true; }
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return false;
    const size_t esz = u8 ? 1 : sizeof(float);
    if ((((uintptr_t)base) & 15) != 0 || ((stride_elems * esz) & 15) != 0 || rows <= 0 || cols <= 0) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)stride_elems * esz};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(tm, u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box,
            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_NONE,  // (L2_128B promotion faults with this 544-byte box)
            u8 ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE : CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA) != CUDA_SUCCESS)
        return false;
    TmapSlot& sl = cache[next++ & 7u];
    memset(&sl.k, 0, sizeof(sl.k));
    sl.k = key; sl.tm = *tm; sl.used = true;
    return true;
}

// the instantiated axis combinations of the fast kernel (bit 0/1: source x/y fractional, bit 2/3: reference x/y):
// pixel-aligned pair, the four "one resampled raster per axis" cases of the dem_align loop, everything fractional
typedef void (*FastKernelFn)(const FrontParams);
static const int kFastAx[6] = {0, 3, 9, 6, 12, 15};
#define DCR_FAST_ROW(AX) {front_fast_kernel<AX, 0>, front_fast_kernel<AX, 1>, front_fast_kernel<AX, 2>}
static FastKernelFn const kFastFn[6][3] = {DCR_FAST_ROW(0), DCR_FAST_ROW(3), DCR_FAST_ROW(9),
                                           DCR_FAST_ROW(6), DCR_FAST_ROW(12), DCR_FAST_ROW(15)};
static FastKernelFn fast_kernel_fn(int k, int mode) { return kFastFn[k][mode]; }

int dcr_front_launch(const dcr_front_args* a, unsigned long long* counters, cudaStream_t stream, unsigned short* bin,
                     int loopmode, unsigned* tile_list, long long tile_list_cap) {
    DCR_ARG(a, "null args");
    DCR_ARG(a->ref && a->src && a->dh && a->slope && a->maskbits, "null raster pointer");
    DCR_ARG(a->h > 0 && a->w > 0, "empty grid");
    FrontParams p;
    memset(&p, 0, sizeof(p));
    p.rg = make_devgeom(a->ref_geom, a->ref_desc.height, a->ref_desc.width, a->ref_stride, a->ref_desc.ndv, a->ref_desc.has_ndv);
    p.sg = make_devgeom(a->src_geom, a->src_desc.height, a->src_desc.width, a->src_stride, a->src_desc.ndv, a->src_desc.has_ndv);
    p.ref = a->ref;
    p.src = a->src;
    p.smask = a->smask;
    p.smask_stride = a->smask_stride;
    p.smask_h = a->smask_h;
    p.smask_w = a->smask_w;
    p.smask_nx0 = a->smask_nx0;
    p.smask_ny0 = a->smask_ny0;
    p.h = a->h;
    p.w = a->w;
    p.ewres = a->ewres;
    p.nsres = a->nsres;
    p.inv8ew = (float)(1.0 / (8.0 * a->ewres));
    p.inv8ns = (float)(1.0 / (8.0 * a->nsres));
    // memwarp_multi: identity -> dataset unchanged (own ndv); resampled -> dst_ndv
    const double rn = a->ref_geom.identity ? a->ref_desc.ndv : a->dst_ndv;
    const double sn = a->src_geom.identity ? a->src_desc.ndv : a->dst_ndv;
    gm_bounds(rn, &p.ref_gm_lo, &p.ref_gm_hi);
    gm_bounds(sn, &p.src_gm_lo, &p.src_gm_hi);
    p.ref_fill = (float)rn;
    p.src_fill = (float)sn;
    p.max_dz = a->max_dz;
    p.use_max_dz = a->use_max_dz;
    p.slope_lo = a->slope_lo;
    p.slope_hi = a->slope_hi;
    p.slope_mid = 0.5f * (a->slope_lo + a->slope_hi);
    p.slope_half = 0.5f * (a->slope_hi - a->slope_lo);
    p.slope_tol = 2e-6f * fmaxf(fabsf(a->slope_lo), fabsf(a->slope_hi)) + 4e-7f * fabsf(p.slope_mid);
    p.dh = a->dh;
    p.slope = a->slope;
    p.aspect = a->aspect;
    p.maskbits = a->maskbits;
    p.ref_w = a->ref_w;
    p.src_w = a->src_w;
    p.counters = counters;
    p.bin = bin;
    DCR_ARG(a->src_ndz >= 0 && a->src_ndz <= 16, "0 <= src_ndz <= 16");
    p.sg.ndz = a->src_ndz;
    for (int k = 0; k < a->src_ndz; ++k) p.sg.dz[k] = a->src_dz[k];
    gm_bounds(a->src_desc.ndv, &p.sg.zlo, &p.sg.zhi);
    // row-band mode: this launch computes output rows [row_begin, row_end) only, from rasters of which only
    // the rows [row0, row0 + nrows) are in memory (the band plus its halo); outputs are band-local.
    p.row_begin = 0;
    p.row_end = a->h;
    if (a->row_end > 0) {
        DCR_ARG(a->row_begin >= 0 && a->row_end <= a->h && a->row_begin < a->row_end, "bad row band");
        p.row_begin = a->row_begin;
        p.row_end = a->row_end;
    }
    if (a->ref_nrows > 0) { p.rg.row0 = a->ref_row0; p.rg.row1 = a->ref_row0 + a->ref_nrows; }
    if (a->src_nrows > 0) { p.sg.row0 = a->src_row0; p.sg.row1 = a->src_row0 + a->src_nrows; }
    DCR_ARG(p.rg.row0 >= 0 && p.rg.row1 <= p.rg.hs && p.sg.row0 >= 0 && p.sg.row1 <= p.sg.hs, "stored rows outside the raster");
    p.smask_row0 = a->smask_nrows > 0 ? a->smask_row0 : 0;
    p.smask_row1 = a->smask_nrows > 0 ? a->smask_row0 + a->smask_nrows : a->smask_h;
    {
        // every source row the launch touches must be in memory: 4 taps (+1 row each side for the Horn halo)
        const int lo = p.row_begin - 1, hi = p.row_end + 1;  // warped-src rows needed (clipped below)
        struct { const DevGeom* g; int nrows; const char* name; int extra; } chk[2] = {
            {&p.rg, a->ref_nrows, "ref", 0}, {&p.sg, a->src_nrows, "src", 1}};
        for (int k = 0; k < 2; ++k) {
            if (chk[k].nrows <= 0) continue;
            int first = (chk[k].extra ? lo : p.row_begin) + chk[k].g->iy0 - 1;
            int last = (chk[k].extra ? hi : p.row_end) + chk[k].g->iy0 + 2;  // exclusive
            if (first < 0) first = 0;
            if (last > chk[k].g->hs) last = chk[k].g->hs;
            if (first < chk[k].g->row0 || last > chk[k].g->row0 + chk[k].nrows) {
                dcr_set_error("row band of %s too small: rows [%d, %d) needed, [%d, %d) in memory -- re-exchange halos",
                              chk[k].name, first, last, chk[k].g->row0, chk[k].g->row0 + chk[k].nrows);
                return -5;
            }
        }
        if (a->smask && a->smask_nrows > 0) {
            int first = p.row_begin + a->smask_ny0, last = p.row_end + a->smask_ny0;
            if (first < 0) first = 0;
            if (last > a->smask_h) last = a->smask_h;
            if (first < p.smask_row0 || last > p.smask_row0 + a->smask_nrows) {
                dcr_set_error("row band of the static mask too small");
                return -5;
            }
        }
    }
    // TMA staging of the two footprints when the layout allows it (16-byte aligned base and row pitch); otherwise
    // the kernel stages them with coalesced row loads (stage_tile)
    p.use_tma = 0;
    if (p.sg.ndz <= 4 && !getenv("DCR_NO_TMA")) {   // (the fix-up pass of the TMA path unrolls at most 4 deferred shifts)
        const bool ok = make_tmap(&p.tm_src, a->src, p.sg.row1 - p.sg.row0, p.sg.ws, p.sg.stride, F_SRC_C, F_SRC_R) &&
                        make_tmap(&p.tm_ref, a->ref, p.rg.row1 - p.rg.row0, p.rg.ws, p.rg.stride, F_REF_C, F_REF_R);
        p.use_tma = ok ? 1 : 0;
        // the copy engine wants every box to start on a 16-byte boundary of the row (measured: other starts raise
        // "illegal instruction"): start up to 3 columns early (the 136-wide boxes have the room) and shift the views
        p.tma_dx_src = ((p.sg.ix0 - 2) % 4 + 4) % 4;
        p.tma_dx_ref = ((p.rg.ix0 - 1) % 4 + 4) % 4;
        if (ok && a->smask &&
            make_tmap(&p.tm_msk, a->smask, p.smask_row1 - p.smask_row0, a->smask_w, a->smask_stride, F_MSK_C, F_TH, true)) {
            p.use_tma_mask = 1;
            p.tma_dx_msk = (a->smask_nx0 % 16 + 16) % 16;
        }
    }
    const size_t smem = (size_t)(F_SRC_WORDS + F_REF_WORDS + F_SW_R * F_SW_C) * sizeof(float);
    DcrCtx* ctx = dcr_ctx();
    if (!ctx) return -1;
    if (!(ctx->attrs & DCR_ATTR_FRONT)) {   // per device
        DCR_CUDA_OK(cudaFuncSetAttribute(front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int k = 0; k < 18; ++k)
            DCR_CUDA_OK(cudaFuncSetAttribute(fast_kernel_fn(k / 3, k % 3), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->attrs |= DCR_ATTR_FRONT;
    }
    p.loopmode = loopmode ? 1 : 0;
    p.max_dz_eff = a->use_max_dz ? a->max_dz : INFINITY;
    {
        const double lo = tan((double)a->slope_lo * (M_PI / 180.0)), hi = tan((double)a->slope_hi * (M_PI / 180.0));
        p.g_lo = (float)lo;
        p.g_hi = (float)hi;
        p.g_mid = (float)(0.5 * (lo + hi));
        p.g_half = (float)(0.5 * (hi - lo));
        p.g_tol = 4e-6f * fmaxf(fabsf(p.g_lo), fabsf(p.g_hi)) + 4e-7f * fabsf(p.g_mid);
    }
    const int ntx = (a->w + F_TW - 1) / F_TW, nty = (p.row_end - p.row_begin + F_TH - 1) / F_TH;
    dim3 grid(ntx, nty);
    // Fast path: whole-grid launches (no row band) with TMA staging, the aspect-bin codes wanted and no warped rasters:
    // front_fast_kernel takes the interior tiles without nodata, the general kernel the rest (from the list).
    const bool slopes_ok = a->slope_lo >= 0.0f && a->slope_hi < 89.0f && a->slope_lo <= a->slope_hi;
    const bool fast = tile_list && bin && p.use_tma && (!a->smask || p.use_tma_mask) && !a->ref_w && !a->src_w &&
                      (long long)ntx * nty + 1 <= tile_list_cap && ntx < 65536 && nty < 65536 && (!loopmode || slopes_ok) &&
                      !getenv("DCR_NO_FAST_FRONT");
    if (fast) {
        int need = (a->src_geom.fx != 0.0 ? 1 : 0) | (a->src_geom.fy != 0.0 ? 2 : 0) | (a->ref_geom.fx != 0.0 ? 4 : 0) |
                   (a->ref_geom.fy != 0.0 ? 8 : 0);
        int k = 5;
        for (int q = 0; q < 6; ++q)
            if ((need & ~kFastAx[q]) == 0) { k = q; break; }
        p.list = tile_list;
        DCR_CUDA_OK(cudaMemsetAsync(tile_list, 0, sizeof(unsigned), stream));
        p.list_mode = 0;
        fast_kernel_fn(k, loopmode ? 2 : (a->aspect ? 0 : 1))<<<grid, 256, smem, stream>>>(p);
        DCR_LAUNCH_OK("front_fast_kernel");
        p.list_mode = 1;
        long long g1 = (long long)ntx * nty;
        const long long gmax = (long long)ctx->sms * 4;
        if (g1 > gmax) g1 = gmax;
        front_kernel<<<(unsigned)g1, 256, smem, stream>>>(p);
        DCR_LAUNCH_OK("front_kernel (listed tiles)");
        return 0;
    }
    front_kernel<<<grid, 256, smem, stream>>>(p);
    DCR_LAUNCH_OK("front_kernel");
    return 0;
}

extern "C" int dcr_warp_diff_horn(const dcr_front_args* args, void* stream) {
    return dcr_front_launch(args, nullptr, (cudaStream_t)stream, nullptr, 0, nullptr, 0);
}
