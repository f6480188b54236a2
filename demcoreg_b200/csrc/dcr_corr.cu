// K7a NCC offset search (K7b SAD: dcr_sad.cu).
//
// Replaces reference coreglib.compute_offset_ncc (coreglib.py:118-198: z-score both rasters over their
// valid pixels, fill masked pixels with noise, scipy.signal.correlate2d(ref, kernel, 'valid')).  The argmax /
// parabolic sub-pixel step (coreglib.py:387-485) stays on the host.
//
// NCC is a direct shared-memory correlation (2*(2p+1)^2 flop per 8 bytes: FP32-issue bound, not HBM
// bound): persistent CTAs walk over 16 x 128 kernel tiles; a CTA stages the tile and the matching reference
// tile (+ halo); a thread owns one offset row and NCC_CJ consecutive offset columns and slides a register window
// along the tile row, so every shared-memory load feeds NCC_CJ FMAs (the kernel values arrive four at a time
// by one broadcast 16-byte load).  Products are accumulated in FP32 over ONE tile row (128 terms) and from
// there in FP64 registers across all the tiles of the CTA; the per-CTA partial maps are summed in FP64 in a
// fixed order (deterministic; ~1e-9 relative on the map against scipy's float64 correlate2d).  Search radii
// whose offset rows do not fit one CTA (p > 19) are split into blocks of offset rows (grid.y).
#include <string.h>
#include "dcr_common.cuh"
#include "dcr_internal.cuh"

// ------------------------------------------------------------------------------------------
// windowed moments + z-score (float64 statistics, like the reference's float64 arithmetic)
// ------------------------------------------------------------------------------------------
struct WinPartial { unsigned long long n; double s, ss; };

// rows r0.., columns c0.. of x (and of its mask): warp = one row at a time, lanes stride the columns (no per-element
// integer division; 4 independent loads in flight per lane)
__global__ void __launch_bounds__(256) win_moments_kernel(const float* __restrict__ x, const unsigned char* __restrict__ m,
                                                          long long xstride, long long mstride, int r0, int c0, int h, int w,
                                                          WinPartial* __restrict__ part) {
    __shared__ WinPartial sp[8];
    unsigned long long cnt = 0;
    double s = 0, ss = 0;
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * 8;
    for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < h; i += nwarps) {
        const float* xr = x + (long long)(r0 + i) * xstride + c0;
        const unsigned char* mr = m ? m + (long long)(r0 + i) * mstride + c0 : nullptr;
        for (int j0 = lane; j0 < w; j0 += 128) {
            float v[4];
            unsigned char mk[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + 32 * u;
                v[u] = j < w ? xr[j] : 0.0f;
                mk[u] = j < w ? (mr ? mr[j] : (unsigned char)0) : (unsigned char)1;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (mk[u]) continue;
                const double d = (double)v[u];
                ++cnt; s += d; ss += d * d;
            }
        }
    }
    cnt = dcr_warp_sum_u64(cnt); s = dcr_warp_sum(s); ss = dcr_warp_sum(ss);
    if (lane == 0) { WinPartial q; q.n = cnt; q.s = s; q.ss = ss; sp[threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        WinPartial q = sp[0];
        for (int k = 1; k < 8; ++k) { q.n += sp[k].n; q.s += sp[k].s; q.ss += sp[k].ss; }
        part[blockIdx.x] = q;
    }
}

// one warp, fixed order: lane l sums partials l, l + 32, ...; then the shuffle tree.  stats[0] = mean, [1] = std (ddof = 0),
// [2] = count
__global__ void win_moments_final_kernel(const WinPartial* __restrict__ part, int nparts, double* stats) {
    unsigned long long n = 0; double s = 0, ss = 0;
    for (int k = threadIdx.x; k < nparts; k += 32) { n += part[k].n; s += part[k].s; ss += part[k].ss; }
    n = dcr_warp_sum_u64(n); s = dcr_warp_sum(s); ss = dcr_warp_sum(ss);
    if (threadIdx.x) return;
    const double mean = n ? s / (double)n : 0.0;
    double var = n ? ss / (double)n - mean * mean : 0.0;
    if (var < 0) var = 0;
    stats[0] = mean; stats[1] = sqrt(var); stats[2] = (double)n;
}

// z = valid ? (x - mean) / std : noise (0 when no noise array): contiguous h x w output
__global__ void __launch_bounds__(256) zscore_kernel(const float* __restrict__ x, const unsigned char* __restrict__ m,
                                                     long long xstride, long long mstride, int r0, int c0, int h, int w,
                                                     const double* __restrict__ stats, const float* __restrict__ noise,
                                                     float* __restrict__ z) {
    const double mean = stats[0], sd = stats[1];
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * 8;
    for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < h; i += nwarps) {
        const float* xr = x + (long long)(r0 + i) * xstride + c0;
        const unsigned char* mr = m ? m + (long long)(r0 + i) * mstride + c0 : nullptr;
        const float* nr = noise ? noise + (long long)i * w : nullptr;
        float* zr = z + (long long)i * w;
        // 4 independent column groups per trip: all their loads are in flight before the first divide
        for (int j0 = lane; j0 < w; j0 += 128) {
            float xv[4], nv[4];
            unsigned char mk[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + 32 * u;
                const bool in = j < w;
                xv[u] = in ? xr[j] : 0.0f;
                mk[u] = (in && mr) ? mr[j] : (unsigned char)0;
                nv[u] = (in && nr) ? nr[j] : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + 32 * u;
                if (j < w) zr[j] = mk[u] ? nv[u] : (float)(((double)xv[u] - mean) / sd);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// NCC: direct correlation
// ------------------------------------------------------------------------------------------
#define NCC_TU 16    // kernel-tile rows
#define NCC_TV 128   // kernel-tile columns
#define NCC_CJ 8     // offset columns per thread (register window)
#define NCC_THREADS 512
#define NCC_MAXP 127 // largest supported pad (search radius in pixels)

// zr: [H x W] z-scored reference; zk: [kh x kw] z-scored kernel (kh = H - 2p, kw = W - 2p).
// CTA (x, y): tiles x, x + gridDim.x, ...; offset rows [y * IB, min(J, (y + 1) * IB)).
// part: [gridDim.x][J*J] float64 partial maps, J = 2p + 1.
__global__ void __launch_bounds__(NCC_THREADS) ncc_tile_kernel(const float* __restrict__ zr, const float* __restrict__ zk, int H,
                                                               int W, int kh, int kw, int p, int IB,
                                                               double* __restrict__ part) {
    extern __shared__ __align__(16) float sm[];
    const int J = 2 * p + 1;
    const int nchunk = (J + NCC_CJ - 1) / NCC_CJ;
    const int i0 = blockIdx.y * IB;
    const int ni = min(IB, J - i0);
    const int RW = NCC_TV + 2 * p;             // reference tile width actually used
    const int RA = NCC_TV + nchunk * NCC_CJ;   // allocated width: the last column chunk may run past J (zero-filled)
    const int RP = RA | 1;                     // odd pitch: lanes on different rows hit different banks
    const int RR = NCC_TU + ni - 1;            // reference tile rows
    float* s_k = sm;                           // NCC_TU x NCC_TV
    float* s_r = sm + NCC_TU * NCC_TV;         // RR x RP
    // work items: (offset row i, column chunk jc, tile-row phase ph); item -> lane mapping keeps i fastest
    const int nitems = ni * nchunk;
    const int nphase = NCC_THREADS / nitems > 0 ? NCC_THREADS / nitems : 1;
    const int item = threadIdx.x % nitems, ph = threadIdx.x / nitems;
    const bool active = ph < nphase && threadIdx.x < nitems * nphase;
    const int i = item % ni, jc = item / ni;
    const int j0 = jc * NCC_CJ;
    double accd[NCC_CJ];
#pragma unroll
    for (int q = 0; q < NCC_CJ; ++q) accd[q] = 0.0;
    const int ntx = (kw + NCC_TV - 1) / NCC_TV, nty = (kh + NCC_TU - 1) / NCC_TU;
    for (int t = blockIdx.x; t < ntx * nty; t += gridDim.x) {
        const int ty = t / ntx, tx = t - ty * ntx;
        const int u0 = ty * NCC_TU, v0 = tx * NCC_TV;
        __syncthreads();
        for (int k = threadIdx.x; k < NCC_TU * NCC_TV; k += NCC_THREADS) {
            const int u = k / NCC_TV, v = k - u * NCC_TV;
            s_k[k] = (u0 + u < kh && v0 + v < kw) ? __ldg(zk + (long long)(u0 + u) * kw + v0 + v) : 0.0f;
        }
        for (int k = threadIdx.x; k < RR * RP; k += NCC_THREADS) {
            const int r = k / RP, c = k - r * RP;
            s_r[k] = (c < RW && u0 + i0 + r < H && v0 + c < W) ? __ldg(zr + (long long)(u0 + i0 + r) * W + v0 + c) : 0.0f;
        }
        __syncthreads();
        if (active) {
            for (int u = ph; u < NCC_TU; u += nphase) {
                const float* rrow = s_r + (u + i) * RP + j0;
                const float4* krow = reinterpret_cast<const float4*>(s_k + u * NCC_TV);
                float acc[NCC_CJ], win[NCC_CJ];
#pragma unroll
                for (int q = 0; q < NCC_CJ; ++q) acc[q] = 0.0f;
#pragma unroll
                for (int q = 0; q < NCC_CJ - 1; ++q) win[q + 1] = rrow[q];
                // slide along the tile row; unrolled by NCC_CJ so the register window rotates for free
                for (int v = 0; v < NCC_TV; v += NCC_CJ) {
                    const float4 ka = krow[v >> 2], kb = krow[(v >> 2) + 1];
                    const float kk[NCC_CJ] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};
#pragma unroll
                    for (int s = 0; s < NCC_CJ; ++s) {
                        // window holds rrow[v+s .. v+s+CJ-1]; shift in the newest element
#pragma unroll
                        for (int q = 0; q < NCC_CJ - 1; ++q) win[q] = win[q + 1];
                        win[NCC_CJ - 1] = rrow[v + s + NCC_CJ - 1];
#pragma unroll
                        for (int q = 0; q < NCC_CJ; ++q) acc[q] = fmaf(win[q], kk[s], acc[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < NCC_CJ; ++q) accd[q] += (double)acc[q];   // FP32 over one tile row, FP64 beyond
            }
        }
    }
    // combine the phases of one item through shared memory (fixed order)
    __syncthreads();
    double* sd = reinterpret_cast<double*>(sm);
    if (active) {
#pragma unroll
        for (int q = 0; q < NCC_CJ; ++q) sd[(ph * nitems + item) * NCC_CJ + q] = accd[q];
    }
    __syncthreads();
    if (threadIdx.x < nitems) {
        double* out = part + (long long)blockIdx.x * (J * J);
        for (int q = 0; q < NCC_CJ; ++q) {
            const int j = j0 + q;
            if (j >= J) break;
            double t = 0.0;
            for (int f = 0; f < nphase; ++f) t += sd[(f * nitems + item) * NCC_CJ + q];
            out[(i0 + i) * J + j] = t;
        }
    }
}

__global__ void ncc_reduce_kernel(const double* __restrict__ part, int nparts, int JJ, double* __restrict__ m) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per map element, partials in CTA order
    if (o >= JJ) return;
    double t = 0.0;
    for (int k = 0; k < nparts; ++k) t += part[(long long)k * JJ + o];
    m[o] = t;
}

// rows of offsets per CTA: every (offset row, column chunk) item of a block needs its own thread
static int ncc_rows_per_block(int pad) {
    const int J = 2 * pad + 1, nchunk = (J + NCC_CJ - 1) / NCC_CJ;
    int ib = NCC_THREADS / nchunk;
    return ib > J ? J : ib;
}
static int ncc_grid_x(int pad) {
    const int J = 2 * pad + 1, ib = ncc_rows_per_block(pad);
    const int nib = (J + ib - 1) / ib;
    int gx = (dcr_num_sms() * 2 + nib - 1) / nib;
    return gx < 1 ? 1 : gx;
}

extern "C" size_t dcr_ncc_workspace_bytes(int h, int w, int pad) {
    const long long kh = h - 2 * pad, kw = w - 2 * pad;
    if (kh <= 0 || kw <= 0 || pad < 1 || pad > NCC_MAXP) return 0;
    const long long J = 2 * pad + 1;
    return dcr_al256((size_t)h * w * 4) + dcr_al256((size_t)kh * kw * 4) + dcr_al256((size_t)ncc_grid_x(pad) * J * J * 8) +
           dcr_al256((size_t)J * J * 8) + dcr_al256(2048 * sizeof(WinPartial)) + dcr_al256(64) + 4096;
}

extern "C" int dcr_ncc_corr(const float* ref, const uint8_t* ref_mask, const float* src, const uint8_t* src_mask, int h, int w,
                            int64_t stride, int pad, const float* ref_noise, const float* kernel_noise, double* m_out /* host */,
                            double stats_out[4] /* host: ref mean,std, kernel mean,std */, void* workspace,
                            size_t workspace_bytes, void* stream) {
    DCR_ARG(ref && src && m_out && workspace, "null pointer");
    DCR_ARG(pad >= 1 && pad <= NCC_MAXP, "1 <= pad <= 127");
    const int kh = h - 2 * pad, kw = w - 2 * pad;
    DCR_ARG(kh > 0 && kw > 0, "raster smaller than the search window");
    DCR_ARG(workspace_bytes >= dcr_ncc_workspace_bytes(h, w, pad), "workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    const int J = 2 * pad + 1;
    const int ib = ncc_rows_per_block(pad), nib = (J + ib - 1) / ib;
    int gx = ncc_grid_x(pad);
    {
        const long long ntiles = (long long)((kw + NCC_TV - 1) / NCC_TV) * ((kh + NCC_TU - 1) / NCC_TU);
        if (gx > ntiles) gx = (int)ntiles;
    }
    DcrBump ws(workspace, workspace_bytes);
    float* zr = ws.take<float>((size_t)h * w);
    float* zk = ws.take<float>((size_t)kh * kw);
    double* part = ws.take<double>((size_t)ncc_grid_x(pad) * J * J);
    double* dm = ws.take<double>((size_t)J * J);
    WinPartial* wp = ws.take<WinPartial>(2048);
    double* stats = ws.take<double>(8);
    const int g = dcr_num_sms() * 4;
    // reference: z-score over its valid pixels (coreglib.py:142-143)
    win_moments_kernel<<<g, 256, 0, s>>>(ref, ref_mask, stride, stride, 0, 0, h, w, wp);
    win_moments_final_kernel<<<1, 32, 0, s>>>(wp, g, stats);
    zscore_kernel<<<g, 256, 0, s>>>(ref, ref_mask, stride, stride, 0, 0, h, w, stats, ref_noise, zr);
    // kernel = src[p:-p, p:-p]
    win_moments_kernel<<<g, 256, 0, s>>>(src, src_mask, stride, stride, pad, pad, kh, kw, wp + 1024);
    win_moments_final_kernel<<<1, 32, 0, s>>>(wp + 1024, g, stats + 4);
    zscore_kernel<<<g, 256, 0, s>>>(src, src_mask, stride, stride, pad, pad, kh, kw, stats + 4, kernel_noise, zk);
    const int nchunk = (J + NCC_CJ - 1) / NCC_CJ;
    const size_t smem_tiles = (size_t)(NCC_TU * NCC_TV + (NCC_TU + ib - 1) * ((NCC_TV + nchunk * NCC_CJ) | 1)) * sizeof(float);
    const size_t smem_red = (size_t)NCC_THREADS * NCC_CJ * sizeof(double);
    const size_t smem = smem_tiles > smem_red ? smem_tiles : smem_red;
    if (smem > 227 * 1024) { dcr_set_error("pad too large for the NCC kernel"); return -1; }
    DCR_CUDA_OK(cudaFuncSetAttribute(ncc_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ncc_tile_kernel<<<dim3(gx, nib), NCC_THREADS, smem, s>>>(zr, zk, h, w, kh, kw, pad, ib, part);
    ncc_reduce_kernel<<<(J * J + 255) / 256, 256, 0, s>>>(part, gx, J * J, dm);
    DCR_LAUNCH_OK("ncc kernels");
    double hs[8];
    DCR_CUDA_OK(cudaMemcpyAsync(m_out, dm, (size_t)J * J * sizeof(double), cudaMemcpyDeviceToHost, s));
    DCR_CUDA_OK(cudaMemcpyAsync(hs, stats, sizeof(hs), cudaMemcpyDeviceToHost, s));
    DCR_CUDA_OK(cudaStreamSynchronize(s));
    if (stats_out) { stats_out[0] = hs[0]; stats_out[1] = hs[1]; stats_out[2] = hs[4]; stats_out[3] = hs[5]; }
    return 0;
}
