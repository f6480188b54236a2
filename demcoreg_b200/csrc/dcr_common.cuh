// Shared device/host helpers for the demcoreg_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/demcoreg_b200.h"

#define DCR_NSM_DEFAULT 148

void dcr_set_error(const char* fmt, ...);
int dcr_num_sms();

#define DCR_CUDA_OK(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            dcr_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int)_e;                                                                 \
        }                                                                                   \
    } while (0)

#define DCR_LAUNCH_OK(name)                                                                 \
    do {                                                                                    \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess) {                                                            \
            dcr_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));         \
            return (int)_e;                                                                 \
        }                                                                                   \
    } while (0)

#define DCR_ARG(cond, msg)                                                                  \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            dcr_set_error("invalid argument: %s", msg);                                     \
            return -1;                                                                      \
        }                                                                                   \
    } while (0)

// order-preserving float32 -> uint32 key (radix selection)
__host__ __device__ __forceinline__ uint32_t dcr_f2key(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float dcr_key2f(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// numpy _lerp in float32: a + (b-a)*g, or b - (b-a)*(1-g) when g >= 0.5  (numpy/lib/_function_base_impl.py)
__host__ __device__ __forceinline__ float dcr_lerp32(float a, float b, double g) {
    float d = b - a;
    if (g >= 0.5) return b - d * (float)(1.0 - g);
    return a + d * (float)g;
}

// workspace bump allocator (256-byte aligned)
struct DcrBump {
    char* base;
    size_t off;
    size_t cap;
    __host__ DcrBump(void* p, size_t c) : base((char*)p), off(0), cap(c) {}
    template <typename T>
    __host__ T* take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        if (off + bytes > cap) return nullptr;
        T* r = (T*)(base + off);
        off += bytes;
        return r;
    }
};
static inline size_t dcr_al256(size_t b) { return (b + 255) & ~(size_t)255; }

// ---- block-level helpers ----------------------------------------------------------------
__device__ __forceinline__ unsigned dcr_warp_incl_scan(unsigned v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (unsigned)o) v += t;
    }
    return v;
}

// exclusive scan over the block (blockDim.x <= 1024, multiple of 32); returns exclusive prefix, total in *total
__device__ __forceinline__ unsigned dcr_block_excl_scan(unsigned v, unsigned* smem_warp /*[33]*/, unsigned* total) {
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned inc = dcr_warp_incl_scan(v);
    if (lane == 31) smem_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned w = lane < nw ? smem_warp[lane] : 0u;
        unsigned winc = dcr_warp_incl_scan(w);
        smem_warp[lane] = winc - w;
        if (lane == 31) smem_warp[32] = winc;
    }
    __syncthreads();
    unsigned r = inc - v + smem_warp[wid];
    *total = smem_warp[32];
    __syncthreads();
    return r;
}

__device__ __forceinline__ double dcr_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long dcr_warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#define DCR_BIN_INVALID 0xFFFFu   // aspect-bin code of a pixel outside the common mask

// y = dh / tan(deg2rad(slope)) in float32, like numpy (deg2rad = x * f32(pi/180); np.tan on float32 is a
// few-ulp libm/SVML routine, so FP32 tanf -- also a few ulp -- is as close as any implementation can get).
__device__ __forceinline__ float nuth_y(float dh, float slope) {
    const float rad = slope * 0.017453292519943295f;
    float t;
    if (fabsf(rad) <= 0.78539816f) {
        // |x| <= pi/4 (slopes up to 45 deg; the default filter keeps 0.1..40): odd minimax polynomial (Cephes tanf),
        // max 1.3 ulp / 14 % not correctly rounded -- numpy's own float32 tan: 1.5 ulp / 24 % (measured on 2e6 samples);
        // avoids the inlined Payne-Hanek slow path of tanf (code size) and its range reduction
        const float z = rad * rad;
        float p = 9.38540185543e-3f;
        p = fmaf(p, z, 3.11992232697e-3f);
        p = fmaf(p, z, 2.44301354525e-2f);
        p = fmaf(p, z, 5.34112807005e-2f);
        p = fmaf(p, z, 1.33387994085e-1f);
        p = fmaf(p, z, 3.33331568548e-1f);
        t = fmaf(p * z, rad, rad);
    } else {
        t = tanf(rad);
    }
    return dh / t;
}

// ---- vectorised streaming ------------------------------------------------------------------
// Each warp iteration covers 512 consecutive elements: lane l takes float4 #(k*32 + l), k = 0..3,
// so every load instruction is a fully coalesced 512-byte request (128 bytes for the mask) and a
// thread keeps 8+ independent loads in flight (the scalar grid-stride loops were latency-bound at
// ~1.2 TB/s: profiles/sel_hist_r1a_raw.csv).  The body f sees (value..., mask byte, element index).
// Falls back to scalar accesses for the tail and for unaligned pointers.
#define DCR_AL16(p) ((((uintptr_t)(p)) & 15) == 0)
#define DCR_AL4(p) ((((uintptr_t)(p)) & 3) == 0)
#define DCR_AL8(p) ((((uintptr_t)(p)) & 7) == 0)

template <typename F>
__device__ __forceinline__ void dcr_stream1(const float* __restrict__ x, const unsigned char* __restrict__ m,
                                            long long n, F&& f) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long nblk = (DCR_AL16(x) && DCR_AL4(m)) ? (n >> 9) : 0;
    for (long long w = warp; w < nblk; w += nwarps) {
        const float4* x4 = reinterpret_cast<const float4*>(x) + (w << 7);
        const unsigned* m4 = m ? reinterpret_cast<const unsigned*>(m) + (w << 7) : nullptr;
        float4 t[4];
        unsigned mb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            t[k] = __ldg(x4 + k * 32 + lane);
            mb[k] = m4 ? __ldg(m4 + k * 32 + lane) : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long b = (w << 9) + ((long long)(k * 32 + lane) << 2);
            f(t[k].x, mb[k] & 0xffu, b);
            f(t[k].y, (mb[k] >> 8) & 0xffu, b + 1);
            f(t[k].z, (mb[k] >> 16) & 0xffu, b + 2);
            f(t[k].w, mb[k] >> 24, b + 3);
        }
    }
    for (long long i = (nblk << 9) + warp * 32 + lane; i < n; i += nwarps * 32) f(x[i], m ? (unsigned)m[i] : 0u, i);
}

// Warp-uniform variant: every lane of a warp calls f the same number of times (so f may use warp
// collectives); out-of-range lanes of the tail get mb = 0xFFFFFFFF (callers must skip mb > 0xFF).
template <typename F>
__device__ __forceinline__ void dcr_stream1_uniform(const float* __restrict__ x, const unsigned char* __restrict__ m,
                                                    long long n, F&& f) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long nblk = (DCR_AL16(x) && DCR_AL4(m)) ? (n >> 9) : 0;
    for (long long w = warp; w < nblk; w += nwarps) {
        const float4* x4 = reinterpret_cast<const float4*>(x) + (w << 7);
        const unsigned* m4 = m ? reinterpret_cast<const unsigned*>(m) + (w << 7) : nullptr;
        float4 t[4];
        unsigned mb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            t[k] = __ldg(x4 + k * 32 + lane);
            mb[k] = m4 ? __ldg(m4 + k * 32 + lane) : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long b = (w << 9) + ((long long)(k * 32 + lane) << 2);
            f(t[k].x, mb[k] & 0xffu, b);
            f(t[k].y, (mb[k] >> 8) & 0xffu, b + 1);
            f(t[k].z, (mb[k] >> 16) & 0xffu, b + 2);
            f(t[k].w, mb[k] >> 24, b + 3);
        }
    }
    for (long long base = (nblk << 9) + warp * 32; base < n; base += nwarps * 32) {
        const long long i = base + lane;
        const bool in = i < n;
        f(in ? x[i] : 0.f, in ? (m ? (unsigned)m[i] : 0u) : 0xFFFFFFFFu, i);
    }
}

// Group variant: f(value, mask byte, index, slot) is called for the 16 elements of a lane's group
// (slot = 0..15, a compile-time constant after unrolling), then g() once per group -- warp-uniform,
// so g may use warp collectives to aggregate what the 16 f calls gathered in registers.
template <typename F, typename G>
__device__ __forceinline__ void dcr_stream1_groups(const float* __restrict__ x, const unsigned char* __restrict__ m,
                                                   long long n, F&& f, G&& g) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long nblk = (DCR_AL16(x) && DCR_AL4(m)) ? (n >> 9) : 0;
    for (long long w = warp; w < nblk; w += nwarps) {
        const float4* x4 = reinterpret_cast<const float4*>(x) + (w << 7);
        const unsigned* m4 = m ? reinterpret_cast<const unsigned*>(m) + (w << 7) : nullptr;
        float4 t[4];
        unsigned mb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            t[k] = __ldg(x4 + k * 32 + lane);
            mb[k] = m4 ? __ldg(m4 + k * 32 + lane) : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long b = (w << 9) + ((long long)(k * 32 + lane) << 2);
            f(t[k].x, mb[k] & 0xffu, b, 4 * k);
            f(t[k].y, (mb[k] >> 8) & 0xffu, b + 1, 4 * k + 1);
            f(t[k].z, (mb[k] >> 16) & 0xffu, b + 2, 4 * k + 2);
            f(t[k].w, mb[k] >> 24, b + 3, 4 * k + 3);
        }
        g();
    }
    for (long long base = (nblk << 9) + warp * 32; base < n; base += nwarps * 32) {
        const long long i = base + lane;
        const bool in = i < n;
        f(in ? x[i] : __int_as_float(0x7fc00000), in ? (m ? (unsigned)m[i] : 0u) : 0xFFFFFFFFu, i, 0);  // NaN: never counted
        g();
    }
}

template <typename F>
__device__ __forceinline__ void dcr_stream3(const float* __restrict__ a, const float* __restrict__ b,
                                            const float* __restrict__ c, const unsigned char* __restrict__ m,
                                            long long n, F&& f) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long nblk = (DCR_AL16(a) && DCR_AL16(b) && DCR_AL16(c) && DCR_AL4(m)) ? (n >> 8) : 0;
    for (long long w = warp; w < nblk; w += nwarps) {  // 256 elements per warp iteration
        const long long o = (w << 6);
        const unsigned* m4 = m ? reinterpret_cast<const unsigned*>(m) + o : nullptr;
        float4 ta[2], tb[2], tc[2];
        unsigned mb[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            ta[k] = __ldg(reinterpret_cast<const float4*>(a) + o + k * 32 + lane);
            tb[k] = __ldg(reinterpret_cast<const float4*>(b) + o + k * 32 + lane);
            tc[k] = __ldg(reinterpret_cast<const float4*>(c) + o + k * 32 + lane);
            mb[k] = m4 ? __ldg(m4 + k * 32 + lane) : 0u;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const long long e = (w << 8) + ((long long)(k * 32 + lane) << 2);
            f(ta[k].x, tb[k].x, tc[k].x, mb[k] & 0xffu, e);
            f(ta[k].y, tb[k].y, tc[k].y, (mb[k] >> 8) & 0xffu, e + 1);
            f(ta[k].z, tb[k].z, tc[k].z, (mb[k] >> 16) & 0xffu, e + 2);
            f(ta[k].w, tb[k].w, tc[k].w, mb[k] >> 24, e + 3);
        }
    }
    for (long long i = (nblk << 8) + warp * 32 + lane; i < n; i += nwarps * 32)
        f(a[i], b[i], c[i], m ? (unsigned)m[i] : 0u, i);
}

template <typename F>
__device__ __forceinline__ void dcr_stream_f_u16(const float* __restrict__ y, const unsigned short* __restrict__ q,
                                                 long long n, F&& f) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long nblk = (DCR_AL16(y) && DCR_AL8(q)) ? (n >> 9) : 0;
    for (long long w = warp; w < nblk; w += nwarps) {
        const float4* y4 = reinterpret_cast<const float4*>(y) + (w << 7);
        const uint2* q4 = reinterpret_cast<const uint2*>(q) + (w << 7);
        float4 t[4];
        uint2 qq[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            t[k] = __ldg(y4 + k * 32 + lane);
            qq[k] = __ldg(q4 + k * 32 + lane);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long b = (w << 9) + ((long long)(k * 32 + lane) << 2);
            f(t[k].x, qq[k].x & 0xffffu, b);
            f(t[k].y, qq[k].x >> 16, b + 1);
            f(t[k].z, qq[k].y & 0xffffu, b + 2);
            f(t[k].w, qq[k].y >> 16, b + 3);
        }
    }
    for (long long i = (nblk << 9) + warp * 32 + lane; i < n; i += nwarps * 32) f(y[i], (unsigned)q[i], i);
}

// 8 consecutive elements starting at `base` (order-sensitive kernels): vector loads when the chunk is
// fully inside the array and the pointers are aligned, scalar otherwise.  mb[k] = 0xFFFFFFFF marks "past the end".
__device__ __forceinline__ void dcr_load8(const float* __restrict__ x, const unsigned char* __restrict__ m,
                                          long long base, long long n, float v[8], unsigned mb[8]) {
    if (base + 8 <= n && DCR_AL16(x) && (m == nullptr || DCR_AL8(m)) && (base & 7) == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x + base));
        const float4 b = __ldg(reinterpret_cast<const float4*>(x + base) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        uint2 q = make_uint2(0u, 0u);
        if (m) q = __ldg(reinterpret_cast<const uint2*>(m + base));
#pragma unroll
        for (int k = 0; k < 4; ++k) { mb[k] = (q.x >> (8 * k)) & 0xffu; mb[4 + k] = (q.y >> (8 * k)) & 0xffu; }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const long long i = base + k;
            if (i < n) { v[k] = x[i]; mb[k] = m ? (unsigned)m[i] : 0u; }
            else { v[k] = 0.f; mb[k] = 0xffffffffu; }
        }
    }
}
