// GDAL cubic warp between north-up grids of the same SRS but DIFFERENT resolution: the one-time
// warplib.memwarp_multi(..., res=args.res) of dem_align.py:298-299 / compute_diff.py:97 when the two inputs do not share a
// pixel size (-res min | max | mean, -tr).  Not on the per-iteration path (the loop itself only ever translates), so one
// thread per output pixel and taps through the read-only cache are enough.
//
// Restates GWKGeneralCase of gdalwarpkernel.cpp as oracle/pygeotools_restated.py:gdal_cubic_resample does [RECALLED]:
//   source position of the output pixel centre sx = (x - X0) / sres, sy = (Y0 - y) / sres, scale = sres / tres;
//   nearest source pixel (floor(sy), floor(sx)) outside the raster or nodata -> dst nodata;
//   scale >= 0.95: GWKCubicResample4Sample -- 4x4 Catmull-Rom at the per-pixel fractions (rows first, FP64, left to right),
//                  bilinear over the valid pixels of the 2x2 when any of the 16 taps is missing;
//   scale <  0.95: GWKResample -- taps i, j in [1 - R, R], R = ceil(2 / scale), around (floor(sx - .5), floor(sy - .5)),
//                  weights GWKCubic((i - dx) * scale), invalid / outside pixels skipped, per-row accumulators, result
//                  acc / wsum unless wsum is within 1e-5 of 1, nodata when wsum < 1e-6.
// Same FP64 operation order as the oracle (file compiled with -fmad=false): results are bit-identical.
#include <math.h>
#include "dcr_common.cuh"
#include "dcr_internal.cuh"

struct ResampleParams {
    const float* src;
    long long sstride;
    int hs, ws;
    float ndv;
    int has_ndv;
    double sx0, sy0, sres;     // source geotransform origin / pixel size
    double xmin, ymax, tres;   // target grid
    double scale;
    int R;
    float* dst;
    long long dstride;
    int h, w;
    float dst_ndv;
};

__device__ __forceinline__ double gwk_cubic(double x) {
    const double ax = fabs(x), x2 = x * x;
    if (ax <= 1.0) return x2 * (1.5 * ax - 2.5) + 1.0;
    if (ax <= 2.0) return x2 * (-0.5 * ax + 2.5) - 4.0 * ax + 2.0;
    return 0.0;
}

__device__ __forceinline__ void catmull_rom(double d, double c[4]) {   // GWKCubicComputeWeights
    const double h = 0.5 * d, t = 3.0 * d, h2 = h * d;
    c[0] = h * (-1.0 + d * (2.0 - d));
    c[1] = 1.0 + h2 * (-5.0 + t);
    c[2] = h * (1.0 + d * (4.0 - t));
    c[3] = h2 * (-1.0 + d);
}

// value + validity of source pixel (r, c); invalid: outside, nodata, NaN
__device__ __forceinline__ bool rs_px(const ResampleParams& p, long long r, long long c, double* v) {
    if (r < 0 || c < 0 || r >= p.hs || c >= p.ws) return false;
    const float x = __ldg(p.src + r * p.sstride + c);
    if (x != x || (p.has_ndv && x == p.ndv)) return false;
    *v = (double)x;
    return true;
}

__global__ void __launch_bounds__(256) resample_cubic_kernel(const ResampleParams p) {
    const int j = blockIdx.x * 32 + (threadIdx.x & 31);
    const int i = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (i >= p.h || j >= p.w) return;
    const double sx = ((p.xmin + ((double)j + 0.5) * p.tres) - p.sx0) / p.sres;
    const double sy = (p.sy0 - (p.ymax - ((double)i + 0.5) * p.tres)) / p.sres;
    const long long nx = (long long)floor(sx), ny = (long long)floor(sy);
    const long long ix = (long long)floor(sx - 0.5), iy = (long long)floor(sy - 0.5);
    const double dx = sx - 0.5 - (double)ix, dy = sy - 0.5 - (double)iy;
    float out = p.dst_ndv;
    double tmp;
    if (rs_px(p, ny, nx, &tmp)) {
        if (p.scale >= 0.95) {
            double cx[4], cy[4], v[4];
            catmull_rom(dx, cx);
            catmull_rom(dy, cy);
            bool all16 = true;
            for (int r = 0; r < 4; ++r) {
                double q[4];
                for (int c = 0; c < 4; ++c) {
                    q[c] = 0.0;
                    all16 = rs_px(p, iy - 1 + r, ix - 1 + c, &q[c]) && all16;
                }
                v[r] = cx[0] * q[0] + cx[1] * q[1] + cx[2] * q[2] + cx[3] * q[3];
            }
            if (all16) {
                out = (float)(cy[0] * v[0] + cy[1] * v[1] + cy[2] * v[2] + cy[3] * v[3]);
            } else {
                // bilinear over the valid pixels of the 2x2, order UL, UR, LR, LL
                double acc = 0.0, div = 0.0, q;
                const double w_ul = (1.0 - dx) * (1.0 - dy), w_ur = dx * (1.0 - dy), w_lr = dx * dy, w_ll = (1.0 - dx) * dy;
                if (rs_px(p, iy, ix, &q)) { div = div + w_ul; acc = acc + q * w_ul; }
                if (rs_px(p, iy, ix + 1, &q)) { div = div + w_ur; acc = acc + q * w_ur; }
                if (rs_px(p, iy + 1, ix + 1, &q)) { div = div + w_lr; acc = acc + q * w_lr; }
                if (rs_px(p, iy + 1, ix, &q)) { div = div + w_ll; acc = acc + q * w_ll; }
                if (div >= 0.00001) out = (float)((div == 1.0) ? acc : acc / div);
            }
        } else {
            double acc = 0.0, wsum = 0.0;
            for (int jj = 1 - p.R; jj <= p.R; ++jj) {
                const double w1 = gwk_cubic(((double)jj - dy) * p.scale);
                double accl = 0.0, wl = 0.0;
                for (int ii = 1 - p.R; ii <= p.R; ++ii) {
                    double q;
                    if (rs_px(p, iy + jj, ix + ii, &q)) {
                        const double w2 = gwk_cubic(((double)ii - dx) * p.scale);
                        accl = accl + q * w2;
                        wl = wl + w2;
                    }
                }
                acc = acc + accl * w1;
                wsum = wsum + wl * w1;
            }
            if (!(wsum < 0.000001)) out = (float)((wsum < 0.99999 || wsum > 1.00001) ? acc / wsum : acc);
        }
    }
    p.dst[(long long)i * p.dstride + j] = out;
}

extern "C" int dcr_resample_cubic(const float* src, const dcr_raster_desc* src_desc, int64_t src_stride, const dcr_grid* grid,
                                  float* dst, int64_t dst_stride, double dst_ndv, void* stream) {
    DCR_ARG(src && src_desc && grid && dst, "null pointer");
    DCR_ARG(src_desc->width > 0 && src_desc->height > 0 && grid->width > 0 && grid->height > 0, "empty raster / grid");
    DCR_ARG(src_desc->gt[2] == 0.0 && src_desc->gt[4] == 0.0 && src_desc->gt[5] < 0.0 && grid->gt[5] < 0.0 &&
            grid->gt[2] == 0.0 && grid->gt[4] == 0.0, "north-up rasters only");
    ResampleParams p;
    p.src = src; p.sstride = src_stride; p.hs = src_desc->height; p.ws = src_desc->width;
    p.ndv = (float)src_desc->ndv; p.has_ndv = src_desc->has_ndv;
    p.sx0 = src_desc->gt[0]; p.sy0 = src_desc->gt[3]; p.sres = src_desc->gt[1];
    p.xmin = grid->gt[0]; p.ymax = grid->gt[3]; p.tres = grid->gt[1];
    p.scale = p.sres / p.tres;
    DCR_ARG(p.scale > 1e-3 && p.scale < 1e3, "resolution ratio out of range");
    p.R = (int)ceil(2.0 / p.scale);
    p.dst = dst; p.dstride = dst_stride; p.h = grid->height; p.w = grid->width; p.dst_ndv = (float)dst_ndv;
    dim3 g((p.w + 31) / 32, (p.h + 7) / 8);
    resample_cubic_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(p);
    DCR_LAUNCH_OK("resample_cubic_kernel");
    return 0;
}
