// -tiltcorr: 2-D polynomial least squares over the filtered residuals and its evaluation / removal.
//
// Replaces, for reference dem_align.py:396-447 and 478-483:
//   geolib.ma_fitpoly(diff_align_filt, order, gt, perc=(0,100), origmask=False)  -> polyfit2d (np.linalg.lstsq on the
//       columns x^i y^j, (i, j) in product(range(order+1), repeat=2), x / y = map coordinates of the pixel centres)
//   geolib.polyval2d(xgrid, ygrid, coeff); coreglib.apply_z_shift(ds, -valgrid); diff_align -= vals
//
// The reference fits in raw map coordinates (x ~ 5e5, y ~ 4e6: the columns are nearly collinear and lstsq's answer is
// the SVD's regularised one).  The tensor-product basis is closed under affine changes of x and y, so the fitted SURFACE
// is the same in any centred / scaled coordinates; here u = (x - x0) / xs, v = (y - y0) / ys in [-1, 1], where the normal
// equations are well conditioned: one FP64 reduction kernel for the moments, a (order+1)^2 solve on the host, and an
// evaluation kernel.  Parity is therefore on the surface (values at every pixel), not on the raw coefficients.
#include <string.h>
#include <stdio.h>
#include "dcr_common.cuh"
#include "dcr_internal.cuh"

#define TILT_MAX_ORDER 3
#define TILT_THREADS 256

struct TiltGeo {
    double gt[6];
    double x0, y0, ixs, iys;   // u = (x - x0) * ixs
};

__device__ __forceinline__ void tilt_uv(const TiltGeo& g, int i, int j, double* u, double* v) {
    const double px = (double)j + 0.5, py = (double)i + 0.5;   // geolib.pixelToMap: pixel centres
    const double x = g.gt[0] + px * g.gt[1] + py * g.gt[2];
    const double y = g.gt[3] + px * g.gt[4] + py * g.gt[5];
    *u = (x - g.x0) * g.ixs;
    *v = (y - g.y0) * g.iys;
}

// per-CTA partial sums: [ (2O+1)^2 moments sum u^p v^q | (O+1)^2 right-hand sides sum u^i v^j z | count ]
template <int O>
__global__ void __launch_bounds__(TILT_THREADS) tilt_sums_kernel(const float* __restrict__ z, const unsigned char* __restrict__ m,
                                                                 unsigned reject, int h, int w, long long zstride, TiltGeo g,
                                                                 double* __restrict__ part) {
    constexpr int NM = (2 * O + 1) * (2 * O + 1), NR = (O + 1) * (O + 1), NT = NM + NR + 1;
    double acc[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) acc[k] = 0.0;
    const long long n = (long long)h * w;
    for (long long k = (long long)blockIdx.x * TILT_THREADS + threadIdx.x; k < n; k += (long long)gridDim.x * TILT_THREADS) {
        const int i = (int)(k / w), j = (int)(k - (long long)i * w);
        if (m && (m[k] & reject)) continue;
        const float zf = z[(long long)i * zstride + j];
        if (isnan(zf)) continue;
        double u, v;
        tilt_uv(g, i, j, &u, &v);
        double pu[2 * O + 1], pv[2 * O + 1];
        pu[0] = 1.0; pv[0] = 1.0;
#pragma unroll
        for (int p = 1; p <= 2 * O; ++p) { pu[p] = pu[p - 1] * u; pv[p] = pv[p - 1] * v; }
#pragma unroll
        for (int p = 0; p <= 2 * O; ++p)
#pragma unroll
            for (int q = 0; q <= 2 * O; ++q) acc[p * (2 * O + 1) + q] += pu[p] * pv[q];
        const double zd = (double)zf;
#pragma unroll
        for (int a = 0; a <= O; ++a)
#pragma unroll
            for (int b = 0; b <= O; ++b) acc[NM + a * (O + 1) + b] += pu[a] * pv[b] * zd;
        acc[NT - 1] += 1.0;
    }
    // fixed-order block reduction: warp shuffles, then warp leaders through shared memory
    __shared__ double sred[TILT_THREADS / 32];
    for (int k = 0; k < NT; ++k) {
        double t = dcr_warp_sum(acc[k]);
        if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int q = 0; q < TILT_THREADS / 32; ++q) s += sred[q];
            part[(size_t)blockIdx.x * NT + k] = s;
        }
        __syncthreads();
    }
}

struct TiltCoef { int order; double c[16]; };

__device__ __forceinline__ double tilt_eval(const TiltCoef& c, double u, double v) {
    // sum_{a,b} c[a*(O+1)+b] u^a v^b, Horner in v inside Horner in u
    const int n1 = c.order + 1;
    double r = 0.0;
    for (int a = c.order; a >= 0; --a) {
        double s = 0.0;
        for (int b = c.order; b >= 0; --b) s = s * v + c.c[a * n1 + b];
        r = r * u + s;
    }
    return r;
}

// vals (optional, float32) = polynomial at every pixel of the h x w grid
__global__ void __launch_bounds__(256) tilt_vals_kernel(float* __restrict__ vals, int h, int w, long long vstride, TiltGeo g,
                                                        TiltCoef c) {
    const long long n = (long long)h * w;
    for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < n; k += (long long)gridDim.x * 256) {
        const int i = (int)(k / w), j = (int)(k - (long long)i * w);
        double u, v;
        tilt_uv(g, i, j, &u, &v);
        vals[(long long)i * vstride + j] = (float)tilt_eval(c, u, v);
    }
}

// band = band + sign * polynomial, numpy semantics of `a = b_getma(band); a = a + dz (float64 array); filled()` written to
// a float32 band: (float)((double)a + dz); pixels within masked_values tolerance of ndv become exactly ndv
__global__ void __launch_bounds__(256) tilt_apply_kernel(float* __restrict__ band, int h, int w, long long stride, TiltGeo g,
                                                         TiltCoef c, double sign, int has_ndv, double ndv, double tol,
                                                         float ndv32) {
    const long long n = (long long)h * w;
    for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < n; k += (long long)gridDim.x * 256) {
        const int i = (int)(k / w), j = (int)(k - (long long)i * w);
        float* p = band + (long long)i * stride + j;
        const float x = *p;
        if (has_ndv && fabs((double)x - ndv) <= tol) { *p = ndv32; continue; }
        double u, v;
        tilt_uv(g, i, j, &u, &v);
        *p = (float)((double)x + sign * tilt_eval(c, u, v));
    }
}

static TiltGeo make_geo(const double gt[6], const dcr_poly2d* mdl) {
    TiltGeo g;
    for (int k = 0; k < 6; ++k) g.gt[k] = gt[k];
    g.x0 = mdl->x0; g.y0 = mdl->y0; g.ixs = 1.0 / mdl->xs; g.iys = 1.0 / mdl->ys;
    return g;
}
static TiltCoef make_coef(const dcr_poly2d* mdl) {
    TiltCoef c;
    c.order = mdl->order;
    for (int k = 0; k < 16; ++k) c.c[k] = mdl->coeff_norm[k];
    return c;
}

static int tilt_grid() { return dcr_num_sms() * 4; }
static int tilt_nt(int order) { return (2 * order + 1) * (2 * order + 1) + (order + 1) * (order + 1) + 1; }

extern "C" size_t dcr_polyfit2d_workspace_bytes(void) { return (size_t)tilt_grid() * tilt_nt(TILT_MAX_ORDER) * sizeof(double) + 256; }

// dense solve with partial pivoting (n <= 16)
static bool solve_dense(int n, long double* A, long double* b, double* x) {
    for (int k = 0; k < n; ++k) {
        int piv = k;
        for (int r = k + 1; r < n; ++r)
            if (fabsl(A[r * n + k]) > fabsl(A[piv * n + k])) piv = r;
        if (fabsl(A[piv * n + k]) < 1e-300L) return false;
        if (piv != k) {
            for (int c = 0; c < n; ++c) { long double t = A[k * n + c]; A[k * n + c] = A[piv * n + c]; A[piv * n + c] = t; }
            long double t = b[k]; b[k] = b[piv]; b[piv] = t;
        }
        for (int r = k + 1; r < n; ++r) {
            const long double f = A[r * n + k] / A[k * n + k];
            for (int c = k; c < n; ++c) A[r * n + c] -= f * A[k * n + c];
            b[r] -= f * b[k];
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        long double s = b[k];
        for (int c = k + 1; c < n; ++c) s -= A[k * n + c] * (long double)x[c];
        x[k] = (double)(s / A[k * n + k]);
    }
    return true;
}

static long double binom(int n, int k) {
    long double r = 1.0L;
    for (int q = 1; q <= k; ++q) r = r * (long double)(n - k + q) / (long double)q;
    return r;
}

extern "C" int dcr_polyfit2d(const float* z, const uint8_t* mask, uint32_t reject, int h, int w, int64_t stride,
                             const double gt[6], int order, dcr_poly2d* model, int64_t* count, void* workspace,
                             size_t workspace_bytes, void* stream) {
    DCR_ARG(z && gt && model && workspace, "null pointer");
    DCR_ARG(h > 0 && w > 0, "empty raster");
    DCR_ARG(order >= 1 && order <= TILT_MAX_ORDER, "1 <= polynomial order <= 3");
    DCR_ARG(workspace_bytes >= dcr_polyfit2d_workspace_bytes(), "workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    memset(model, 0, sizeof(*model));
    model->order = order;
    // centre / half-extent of the grid in map coordinates (pixel centres)
    const double xa = gt[0] + 0.5 * gt[1] + 0.5 * gt[2], ya = gt[3] + 0.5 * gt[4] + 0.5 * gt[5];
    const double xb = gt[0] + (w - 0.5) * gt[1] + (h - 0.5) * gt[2], yb = gt[3] + (w - 0.5) * gt[4] + (h - 0.5) * gt[5];
    model->x0 = 0.5 * (xa + xb);
    model->y0 = 0.5 * (ya + yb);
    model->xs = fmax(0.5 * fabs(xb - xa), fabs(gt[1]));
    model->ys = fmax(0.5 * fabs(yb - ya), fabs(gt[5]));
    TiltGeo g = make_geo(gt, model);
    const int grid = tilt_grid(), nt = tilt_nt(order);
    double* part = (double*)workspace;
    if (order == 1) tilt_sums_kernel<1><<<grid, TILT_THREADS, 0, s>>>(z, mask, reject, h, w, stride, g, part);
    else if (order == 2) tilt_sums_kernel<2><<<grid, TILT_THREADS, 0, s>>>(z, mask, reject, h, w, stride, g, part);
    else tilt_sums_kernel<3><<<grid, TILT_THREADS, 0, s>>>(z, mask, reject, h, w, stride, g, part);
    DCR_LAUNCH_OK("tilt_sums_kernel");
    static thread_local double* hp = nullptr;
    static thread_local size_t hp_n = 0;
    const size_t need = (size_t)grid * nt;
    if (hp_n < need) {
        if (hp) cudaFreeHost(hp);
        DCR_CUDA_OK(cudaMallocHost(&hp, need * sizeof(double)));
        hp_n = need;
    }
    DCR_CUDA_OK(cudaMemcpyAsync(hp, part, need * sizeof(double), cudaMemcpyDeviceToHost, s));
    DCR_CUDA_OK(cudaStreamSynchronize(s));
    long double sums[96];
    for (int k = 0; k < nt; ++k) {
        long double t = 0.0L;
        for (int b = 0; b < grid; ++b) t += (long double)hp[(size_t)b * nt + k];   // fixed order
        sums[k] = t;
    }
    const int nm1 = 2 * order + 1, n1 = order + 1, nb = n1 * n1, NM = nm1 * nm1;
    const long long cnt = (long long)llroundl(sums[nt - 1]);
    if (count) *count = cnt;
    if (cnt < nb) { dcr_set_error("polynomial fit: %lld samples for %d coefficients", cnt, nb); return -6; }
    long double A[256], bb[16];
    for (int k = 0; k < nb; ++k) {
        const int ik = k / n1, jk = k % n1;
        for (int l = 0; l < nb; ++l) {
            const int il = l / n1, jl = l % n1;
            A[k * nb + l] = sums[(ik + il) * nm1 + (jk + jl)];
        }
        bb[k] = sums[NM + k];
    }
    if (!solve_dense(nb, A, bb, model->coeff_norm)) { dcr_set_error("polynomial fit: singular normal equations"); return -6; }
    // the same surface in the reference's raw basis: sum m[a*(O+1)+b] x^a y^b (binomial expansion of (x-x0)/xs, (y-y0)/ys)
    long double ex[4][4], ey[4][4];   // ex[i][a]: coefficient of x^a in ((x - x0)/xs)^i
    for (int i = 0; i <= order; ++i)
        for (int a = 0; a <= order; ++a) {
            ex[i][a] = ey[i][a] = 0.0L;
            if (a <= i) {
                ex[i][a] = binom(i, a) * powl(-(long double)model->x0, i - a) / powl((long double)model->xs, i);
                ey[i][a] = binom(i, a) * powl(-(long double)model->y0, i - a) / powl((long double)model->ys, i);
            }
        }
    for (int a = 0; a <= order; ++a)
        for (int b = 0; b <= order; ++b) {
            long double t = 0.0L;
            for (int i = a; i <= order; ++i)
                for (int j = b; j <= order; ++j) t += (long double)model->coeff_norm[i * n1 + j] * ex[i][a] * ey[j][b];
            model->coeff[a * n1 + b] = (double)t;
        }
    return 0;
}

extern "C" int dcr_polyval2d(const dcr_poly2d* model, const double gt[6], int h, int w, float* vals, int64_t vals_stride,
                             void* stream) {
    DCR_ARG(model && gt && vals, "null pointer");
    DCR_ARG(h > 0 && w > 0, "empty raster");
    DCR_ARG(model->order >= 1 && model->order <= TILT_MAX_ORDER, "bad model");
    tilt_vals_kernel<<<dcr_num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(vals, h, w, vals_stride, make_geo(gt, model),
                                                                       make_coef(model));
    DCR_LAUNCH_OK("tilt_vals_kernel");
    return 0;
}

extern "C" int dcr_polyval2d_apply(float* band, int h, int w, int64_t stride, const double gt[6], double ndv, int has_ndv,
                                   const dcr_poly2d* model, double sign, void* stream) {
    DCR_ARG(band && model && gt, "null pointer");
    DCR_ARG(h > 0 && w > 0, "empty raster");
    DCR_ARG(model->order >= 1 && model->order <= TILT_MAX_ORDER, "bad model");
    const double tol = 1e-8 + 1e-5 * fabs(ndv);
    tilt_apply_kernel<<<dcr_num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(band, h, w, stride, make_geo(gt, model),
                                                                        make_coef(model), sign, has_ndv, ndv, tol, (float)ndv);
    DCR_LAUNCH_OK("tilt_apply_kernel");
    return 0;
}
