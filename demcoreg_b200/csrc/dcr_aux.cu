// Raster pieces next to the hot path: exact masked medians along one axis of a raster
// (coreglib.successive_med, coreglib.py:490-584: np.ma.median(a, axis) + np.ma.count(a, axis)), iterated binary dilation of
// a mask (dem_mask.py:561-565: scipy.ndimage.binary_dilation, default cross structure) and the threshold / class-table
// masks of dem_mask.py:111-161, 403-409.
#include <math.h>
#include "dcr_common.cuh"
#include "dcr_internal.cuh"

// ---- exact masked median of every row (axis = 1) or column (axis = 0) --------------------------------------------------
// One CTA per line.  The valid keys of the line are staged in shared memory (monotone uint32 keys of the float32 values),
// then a 4 x 8-bit radix select finds the element of rank floor((n - 1) / 2); for even n the next one is either the same
// value (duplicates) or the smallest key above it.  np.ma.median: mean of the two middle values in float32.
#define AM_THREADS 256
__global__ void __launch_bounds__(AM_THREADS) axis_median_kernel(const float* __restrict__ a, const unsigned char* __restrict__ m,
                                                                int len, long long step_a, long long inc_a,
                                                                long long step_m, long long inc_m, float* __restrict__ med,
                                                                long long* __restrict__ count, unsigned* __restrict__ spill,
                                                                int smem_keys) {
    extern __shared__ unsigned keys_s[];
    __shared__ unsigned hist[256];
    __shared__ unsigned s_n, s_bin, s_below;
    const long long line = blockIdx.x;
    // element k of line `line`: a[line * step_a + k * inc_a]  (axis 1: step = pitch, inc = 1; axis 0: step = 1, inc = pitch)
    const float* base = a + line * step_a;
    const unsigned char* mb = m ? m + line * step_m : nullptr;
    unsigned* keys = (len <= smem_keys) ? keys_s : spill + line * (long long)len;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    // compact the valid keys (order does not matter)
    for (int k0 = 0; k0 < len; k0 += AM_THREADS) {
        const int k = k0 + threadIdx.x;
        bool ok = false;
        unsigned key = 0;
        if (k < len) {
            const float v = base[(long long)k * inc_a];
            ok = !(mb && mb[(long long)k * inc_m]) && v == v;
            key = dcr_f2key(v + 0.0f);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        unsigned wbase = 0;
        if ((threadIdx.x & 31) == 0 && bal) wbase = atomicAdd(&s_n, (unsigned)__popc(bal));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (ok) keys[wbase + __popc(bal & ((1u << (threadIdx.x & 31)) - 1u))] = key;
    }
    __syncthreads();
    const unsigned n = s_n;
    if (threadIdx.x == 0 && count) count[line] = (long long)n;
    if (n == 0) {
        if (threadIdx.x == 0) med[line] = __int_as_float(0x7fc00000);   // fully masked line (np.ma.masked)
        return;
    }
    unsigned rank = (n - 1) >> 1;     // lower middle element
    unsigned prefix = 0, pmask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[threadIdx.x] = 0u;
        __syncthreads();
        for (unsigned k = threadIdx.x; k < n; k += AM_THREADS) {
            const unsigned key = keys[k];
            if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned cum = 0, b = 0;
            for (; b < 256; ++b) {
                if (cum + hist[b] > rank) break;
                cum += hist[b];
            }
            s_bin = b; s_below = cum;
        }
        __syncthreads();
        prefix |= s_bin << shift;
        pmask |= 255u << shift;
        rank -= s_below;
        __syncthreads();
    }
    // prefix = key of the lower middle element; rank = its index among its duplicates
    const unsigned klo = prefix;
    unsigned khi = klo;
    if ((n & 1u) == 0u) {
        // upper middle element: the same value when enough duplicates remain, else the smallest key above
        hist[threadIdx.x] = 0xFFFFFFFFu;
        __shared__ unsigned s_dup;
        if (threadIdx.x == 0) s_dup = 0;
        __syncthreads();
        unsigned dup = 0, mn = 0xFFFFFFFFu;
        for (unsigned k = threadIdx.x; k < n; k += AM_THREADS) {
            const unsigned key = keys[k];
            dup += key == klo;
            if (key > klo) mn = min(mn, key);
        }
        dup = __reduce_add_sync(0xffffffffu, dup);
        mn = __reduce_min_sync(0xffffffffu, mn);
        if ((threadIdx.x & 31) == 0) { atomicAdd(&s_dup, dup); atomicMin(&hist[0], mn); }
        __syncthreads();
        khi = (rank + 1 < s_dup) ? klo : hist[0];
    }
    if (threadIdx.x == 0) {
        const float lo = dcr_key2f(klo), hi = dcr_key2f(khi);
        med[line] = (n & 1u) ? lo : (lo + hi) / 2.0f;     // np.ma.median: float32 mean of the two middle values
    }
}

extern "C" size_t dcr_axis_median_workspace_bytes(int h, int w, int axis) {
    const long long len = axis == 1 ? w : h, lines = axis == 1 ? h : w;
    return len > 40960 ? (size_t)(len * lines) * sizeof(unsigned) : 256;
}

extern "C" int dcr_axis_median(const float* a, const uint8_t* mask, int h, int w, int64_t stride, int64_t mask_stride, int axis,
                               float* med, int64_t* count, void* workspace, size_t workspace_bytes, void* stream) {
    DCR_ARG(a && med, "null pointer");
    DCR_ARG(h > 0 && w > 0 && (axis == 0 || axis == 1), "empty raster / bad axis");
    const int len = axis == 1 ? w : h, lines = axis == 1 ? h : w;
    const int smem_keys = 40960;                        // 160 KB of keys per CTA
    const bool spill = len > smem_keys;
    DCR_ARG(!spill || (workspace && workspace_bytes >= dcr_axis_median_workspace_bytes(h, w, axis)), "workspace too small");
    const size_t smem = (size_t)(spill ? 1 : len) * sizeof(unsigned);
    DCR_CUDA_OK(cudaFuncSetAttribute(axis_median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_keys * 4));
    axis_median_kernel<<<lines, AM_THREADS, smem, (cudaStream_t)stream>>>(
        a, mask, len, axis == 1 ? stride : 1, axis == 1 ? 1 : stride, axis == 1 ? mask_stride : 1, axis == 1 ? 1 : mask_stride,
        med, (long long*)count, (unsigned*)workspace, smem_keys);
    DCR_LAUNCH_OK("axis_median_kernel");
    return 0;
}

// ---- b = a - expand_dims(corr, axis) on the unmasked pixels (successive_med's correction step) -------------------------
__global__ void __launch_bounds__(256) axis_subtract_kernel(float* __restrict__ a, int h, int w, long long stride,
                                                            const float* __restrict__ corr, int axis) {
    const long long n = (long long)h * w;
    for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < n; k += (long long)gridDim.x * 256) {
        const int i = (int)(k / w), j = (int)(k - (long long)i * w);
        a[(long long)i * stride + j] = a[(long long)i * stride + j] - corr[axis == 1 ? i : j];
    }
}
extern "C" int dcr_axis_subtract(float* a, int h, int w, int64_t stride, const float* corr, int axis, void* stream) {
    DCR_ARG(a && corr && h > 0 && w > 0 && (axis == 0 || axis == 1), "bad argument");
    axis_subtract_kernel<<<dcr_num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(a, h, w, stride, corr, axis);
    DCR_LAUNCH_OK("axis_subtract_kernel");
    return 0;
}

// ---- scipy.ndimage.binary_dilation(mask, iterations = n): cross-shaped structure, border value 0 -----------------------
__global__ void __launch_bounds__(256) dilate_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out,
                                                     int h, int w) {
    const int j = blockIdx.x * 64 + (threadIdx.x & 63), i = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (i >= h || j >= w) return;
    const long long o = (long long)i * w + j;
    unsigned v = in[o];
    if (i > 0) v |= in[o - w];
    if (i + 1 < h) v |= in[o + w];
    if (j > 0) v |= in[o - 1];
    if (j + 1 < w) v |= in[o + 1];
    out[o] = v ? 1 : 0;
}
// mask (h*w uint8, 0/1) is dilated `iterations` times in place; scratch: h*w bytes
extern "C" int dcr_binary_dilate(uint8_t* mask, int h, int w, int iterations, uint8_t* scratch, void* stream) {
    DCR_ARG(mask && scratch && h > 0 && w > 0 && iterations >= 0, "bad argument");
    dim3 g((w + 63) / 64, (h + 3) / 4);
    uint8_t* src = mask;
    uint8_t* dst = scratch;
    for (int k = 0; k < iterations; ++k) {
        dilate_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(src, dst, h, w);
        uint8_t* t = src; src = dst; dst = t;
    }
    DCR_LAUNCH_OK("dilate_kernel");
    if (src != mask) DCR_CUDA_OK(cudaMemcpyAsync(mask, src, (size_t)h * w, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

// ---- class-table and threshold masks (dem_mask.py:111-161, 403-409) ------------------------------------------------------
struct Lut256 { unsigned char v[256]; };
__global__ void __launch_bounds__(256) lut_mask_kernel(const unsigned char* __restrict__ l, long long n, Lut256 lut,
                                                       unsigned char* __restrict__ out) {
    for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < n; k += (long long)gridDim.x * 256) out[k] = lut.v[l[k]];
}
extern "C" int dcr_mask_from_classes(const uint8_t* classes, int64_t n, const uint8_t lut[256], uint8_t* out, void* stream) {
    DCR_ARG(classes && lut && out && n > 0, "bad argument");
    Lut256 t;
    for (int k = 0; k < 256; ++k) t.v[k] = lut[k];
    lut_mask_kernel<<<dcr_num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(classes, n, t, out);
    DCR_LAUNCH_OK("lut_mask_kernel");
    return 0;
}
// out = 1 where x > thresh (greater = 1) or x <= thresh (greater = 0); NaN compares false like numpy
__global__ void __launch_bounds__(256) thresh_mask_kernel(const float* __restrict__ x, long long n, float t, int greater,
                                                          unsigned char* __restrict__ out) {
    for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < n; k += (long long)gridDim.x * 256)
        out[k] = greater ? (x[k] > t) : (x[k] <= t);
}
extern "C" int dcr_mask_threshold(const float* x, int64_t n, float thresh, int greater, uint8_t* out, void* stream) {
    DCR_ARG(x && out && n > 0, "bad argument");
    thresh_mask_kernel<<<dcr_num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(x, n, thresh, greater, out);
    DCR_LAUNCH_OK("thresh_mask_kernel");
    return 0;
}
