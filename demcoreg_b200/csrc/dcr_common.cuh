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
