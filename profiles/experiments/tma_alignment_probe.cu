#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <math.h>
#ifndef BW
#define BW 136
#endif
#ifndef BH
#define BH 33
#endif
#ifndef X0
#define X0 -3
#endif
#ifndef Y0
#define Y0 80
#endif
#ifndef OOB
#define OOB CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA
#endif
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
struct P { int x, y; float* out; };
__global__ void k(const __grid_constant__ CUtensorMap tm, const P p) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((unsigned)(BW*BH*4)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(sm)), "l"(&tm), "r"(p.x), "r"(p.y), "r"(smem_u32(&bar)) : "memory");
    }
    unsigned spins = 0, ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        if (++spins > (1u<<22)) { if (threadIdx.x==0) printf("timeout smem=%u bar=%u\n", smem_u32(sm), smem_u32(&bar)); return; }
    }
    for (int i = threadIdx.x; i < BW*BH; i += blockDim.x) p.out[i] = sm[i];
}
int main() {
    const int H = 100, W = 200;
    float* h = (float*)malloc(H*W*4);
    for (int i = 0; i < H*W; ++i) h[i] = (float)i;
    float *d, *o; cudaMalloc(&d, H*W*4); cudaMalloc(&o, BW*BH*4);
    cudaMemcpy(d, h, H*W*4, cudaMemcpyHostToDevice);
    void* f = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr);
    printf("entry %d %d %p\n", (int)e, (int)qr, f);
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    P p; memset(&p, 0, sizeof(p)); CUtensorMap tmh;
    cuuint64_t dims[2] = {W, H}; cuuint64_t strides[1] = {W*4}; cuuint32_t box[2] = {BW, BH}; cuuint32_t es[2] = {1,1};
    CUresult r = ((Fn)f)(&tmh, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, OOB);
    printf("encode %d\n", (int)r);
    p.x = X0; p.y = Y0; p.out = o;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    k<<<1, 256, BW*BH*4 + 1024>>>(tmh, p);
    e = cudaDeviceSynchronize(); printf("sync %s\n", cudaGetErrorString(e));
    float* ho = (float*)malloc(BW*BH*4); cudaMemcpy(ho, o, BW*BH*4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < BH; ++r2) for (int c = 0; c < BW; ++c) {
        int gr = Y0 + r2, gc = X0 + c; float v = ho[r2*BW+c];
        bool oob = gr < 0 || gr >= H || gc < 0 || gc >= W;
        if (oob ? !isnan(v) : v != (float)(gr*W+gc)) { if (bad < 5) printf("mismatch r%d c%d v=%f\n", r2, c, v); ++bad; }
    }
    printf("bad=%d first=%f %f\n", bad, ho[0], ho[3]);
    return 0;
}
