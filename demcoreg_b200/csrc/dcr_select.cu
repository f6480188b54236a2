// K3 exact order statistics (radix select, 11+11+10 bits), K4 moments, K5 strided compression.
//
// Replaces numpy.percentile / partition as reached through pygeotools malib.fast_median, mad,
// calcperc (reference coreglib.py:88, 214-215; filtlib.mad_fltr <- dem_align.py:46), the
// float64 mean/std/min/max of malib.get_stats (reference dem_align.py:114) and the
// `compressed()[::stride]` subsample of malib.get_stats.
//
// Integer histograms only: results are exact and independent of launch geometry.
#include <string.h>
#include "dcr_common.cuh"
#include "dcr_internal.cuh"

#define SEL_THREADS 256
#define CHUNK 2048  // elements per block in the order-sensitive (compression) kernels

__device__ __forceinline__ bool sel_get(const SelView& v, long long i, float center, float* out) {
    if (v.m && (v.m[i] & v.reject)) return false;
    float t = v.x[i];
    if (v.transform == 1) t = fabsf(t - center);
    if (isnan(t)) return false;
    *out = t;
    return true;
}

template <int PASS>
__global__ void __launch_bounds__(SEL_THREADS) sel_hist_kernel(const SelView v, const SelState* __restrict__ st,
                                                               unsigned* __restrict__ ghist) {
    __shared__ unsigned h[2][2048];
    for (int k = threadIdx.x; k < 2 * 2048; k += SEL_THREADS) (&h[0][0])[k] = 0u;
    unsigned plo = 0, phi = 0;
    int same = 1;
    if (PASS > 0) {
        if (st->empty) return;
        plo = st->prefix_lo;
        phi = st->prefix_hi;
        same = st->same;
    }
    __syncthreads();
    const long long n = v.n_ptr ? *v.n_ptr : v.n;
    const float center = v.center_ptr ? *v.center_ptr : v.center_imm;
    constexpr int shift_prev = (PASS == 1) ? 21 : 10;
    constexpr int dshift = (PASS == 0) ? 21 : (PASS == 1 ? 10 : 0);
    constexpr unsigned dmask = (PASS == 2) ? 1023u : 2047u;
    const unsigned reject = v.m ? v.reject : 0u;
    const int transform = v.transform;
    dcr_stream1(v.x, v.m, n, [&](float t, unsigned mb, long long) {
        if (mb & reject) return;
        if (transform == 1) t = fabsf(t - center);
        if (isnan(t)) return;
        const unsigned key = dcr_f2key(t);
        if (PASS == 0) {
            atomicAdd(&h[0][key >> 21], 1u);
        } else {
            if (((key ^ plo) >> shift_prev) == 0u) atomicAdd(&h[0][(key >> dshift) & dmask], 1u);
            if (!same && ((key ^ phi) >> shift_prev) == 0u) atomicAdd(&h[1][(key >> dshift) & dmask], 1u);
        }
    });
    __syncthreads();
    for (int k = threadIdx.x; k < 2 * 2048; k += SEL_THREADS) {
        unsigned c = (&h[0][0])[k];
        if (c) atomicAdd(&ghist[k], c);
    }
}

// whole block: find bucket b with cum_excl(b) <= k < cum_incl(b) in hist[nb]
__device__ void find_bucket(const unsigned* __restrict__ hist, int nb, long long k, unsigned* s_scratch,
                            int* s_bucket, long long* s_before) {
    // 1024 threads, nb <= 2048: two bins per thread
    const int t = threadIdx.x;
    unsigned a = (2 * t < nb) ? hist[2 * t] : 0u;
    unsigned b = (2 * t + 1 < nb) ? hist[2 * t + 1] : 0u;
    unsigned total;
    unsigned ex = dcr_block_excl_scan(a + b, s_scratch, &total);
    if ((long long)ex <= k && k < (long long)ex + a) { *s_bucket = 2 * t; *s_before = ex; }
    else if ((long long)ex + a <= k && k < (long long)ex + a + b) { *s_bucket = 2 * t + 1; *s_before = (long long)ex + a; }
    __syncthreads();
}

template <int PASS>
__global__ void __launch_bounds__(1024) sel_find_kernel(SelState* st, unsigned* ghist, double q_percent) {
    __shared__ unsigned s_scratch[33];
    __shared__ int s_b[2];
    __shared__ long long s_before[2];
    constexpr int nb = (PASS == 2) ? 1024 : 2048;
    constexpr int dshift = (PASS == 0) ? 21 : (PASS == 1 ? 10 : 0);
    if (PASS > 0 && st->empty) return;
    long long klo, khi;
    int same;
    if (PASS == 0) {
        // count = sum of the first histogram
        unsigned a = ghist[2 * threadIdx.x], b = ghist[2 * threadIdx.x + 1];
        unsigned total;
        (void)dcr_block_excl_scan(a + b, s_scratch, &total);
        // total may exceed 2^32 only for n > 4.29e9 elements; rasters here are < 2^31 pixels
        const long long count = (long long)total;
        if (count == 0) {
            if (threadIdx.x == 0) {
                st->empty = 1; st->count = 0;
                st->v_lo = st->v_hi = st->result = __int_as_float(0x7fc00000);
            }
            for (int k = threadIdx.x; k < DCR_SEL_HIST_WORDS; k += 1024) ghist[k] = 0u;
            return;
        }
        // exact rank: virtual index (n-1)*(q/100) in float64
        const double vi = (double)(count - 1) * (q_percent / 100.0);
        klo = (long long)floor(vi);
        khi = klo + 1 < count ? klo + 1 : count - 1;
        if (threadIdx.x == 0) {
            st->empty = 0; st->count = count; st->gamma = vi - (double)klo;
        }
        same = 1;
    } else {
        klo = st->k_lo; khi = st->k_hi; same = st->same;
    }
    __syncthreads();
    find_bucket(ghist, nb, klo, s_scratch, &s_b[0], &s_before[0]);
    find_bucket(same ? ghist : ghist + 2048, nb, khi, s_scratch, &s_b[1], &s_before[1]);
    if (threadIdx.x == 0) {
        unsigned plo = (PASS == 0 ? 0u : st->prefix_lo) | ((unsigned)s_b[0] << dshift);
        unsigned phi = (PASS == 0 ? 0u : st->prefix_hi) | ((unsigned)s_b[1] << dshift);
        st->prefix_lo = plo; st->prefix_hi = phi;
        st->k_lo = klo - s_before[0];
        st->k_hi = khi - s_before[1];
        st->same = (plo == phi) ? 1 : 0;
        if (PASS == 2) {
            const float a = dcr_key2f(plo), b = dcr_key2f(phi);
            st->v_lo = a; st->v_hi = b;
            st->result = dcr_lerp32(a, b, st->gamma);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < DCR_SEL_HIST_WORDS; k += 1024) ghist[k] = 0u;
}

int dcr_sel_enqueue(const SelView& v, double q_percent, SelState* st, unsigned* hist, cudaStream_t stream) {
    const int grid = dcr_num_sms() * 8;
    DCR_CUDA_OK(cudaMemsetAsync(hist, 0, DCR_SEL_HIST_WORDS * sizeof(unsigned), stream));
    sel_hist_kernel<0><<<grid, SEL_THREADS, 0, stream>>>(v, st, hist);
    sel_find_kernel<0><<<1, 1024, 0, stream>>>(st, hist, q_percent);
    sel_hist_kernel<1><<<grid, SEL_THREADS, 0, stream>>>(v, st, hist);
    sel_find_kernel<1><<<1, 1024, 0, stream>>>(st, hist, q_percent);
    sel_hist_kernel<2><<<grid, SEL_THREADS, 0, stream>>>(v, st, hist);
    sel_find_kernel<2><<<1, 1024, 0, stream>>>(st, hist, q_percent);
    DCR_LAUNCH_OK("sel kernels");
    return 0;
}

// single passes of the radix select (row-band mode all-reduces `hist` between the hist and find steps)
int dcr_sel_hist_pass(const SelView& v, SelState* st, unsigned* hist, int pass, cudaStream_t s) {
    const int grid = dcr_num_sms() * 8;
    if (pass == 0) sel_hist_kernel<0><<<grid, SEL_THREADS, 0, s>>>(v, st, hist);
    else if (pass == 1) sel_hist_kernel<1><<<grid, SEL_THREADS, 0, s>>>(v, st, hist);
    else sel_hist_kernel<2><<<grid, SEL_THREADS, 0, s>>>(v, st, hist);
    DCR_LAUNCH_OK("sel_hist_kernel");
    return 0;
}
int dcr_sel_find_pass(SelState* st, unsigned* hist, double q_percent, int pass, cudaStream_t s) {
    if (pass == 0) sel_find_kernel<0><<<1, 1024, 0, s>>>(st, hist, q_percent);
    else if (pass == 1) sel_find_kernel<1><<<1, 1024, 0, s>>>(st, hist, q_percent);
    else sel_find_kernel<2><<<1, 1024, 0, s>>>(st, hist, q_percent);
    DCR_LAUNCH_OK("sel_find_kernel");
    return 0;
}

extern "C" size_t dcr_select_workspace_bytes(void) {
    // worst case used by engine.py for any n: query with dcr_select_workspace_bytes_n for large inputs
    return dcr_al256(DCR_SEL_HIST_WORDS * sizeof(unsigned)) + dcr_al256(16 * sizeof(SelState)) +
           dcr_al256(16 * sizeof(BselState)) + dcr_bsel_scratch_bytes(0) + 1024;
}
extern "C" size_t dcr_select_workspace_bytes_n(int64_t n) {
    return dcr_al256(DCR_SEL_HIST_WORDS * sizeof(unsigned)) + dcr_al256(16 * sizeof(SelState)) +
           dcr_al256(16 * sizeof(BselState)) + dcr_bsel_scratch_bytes(n) + 1024;
}

extern "C" int dcr_select_quantiles(const float* x, const uint8_t* mask, uint32_t reject, int64_t n, int transform,
                                    float center, const double* q_percent, int nq, float* out, int64_t* count,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    DCR_ARG(x && q_percent && out && workspace, "null pointer");
    DCR_ARG(nq >= 1 && nq <= 16, "1 <= nq <= 16");
    DCR_ARG(n >= 0, "n >= 0");
    DCR_ARG(workspace_bytes >= dcr_select_workspace_bytes_n(n), "workspace too small (dcr_select_workspace_bytes_n)");
    cudaStream_t s = (cudaStream_t)stream;
    DcrBump ws(workspace, workspace_bytes);
    unsigned* hist = ws.take<unsigned>(DCR_SEL_HIST_WORDS);
    SelState* st = ws.take<SelState>(16);
    BselState* bst = ws.take<BselState>(16);
    void* scratch = ws.take<char>(dcr_bsel_scratch_bytes(n));
    SelView v;
    memset(&v, 0, sizeof(v));
    v.x = x; v.m = mask; v.reject = reject; v.transform = transform; v.center_ptr = nullptr; v.center_imm = center;
    v.n = n; v.n_ptr = nullptr;
    // fast path: one streaming pass per quantile (dcr_bsel.cu) ...
    for (int k = 0; k < nq; ++k) {
        int rc = dcr_bsel_enqueue(v, q_percent[k], bst + k, scratch, n, s);
        if (rc) return rc;
    }
    BselState hb[16];
    DCR_CUDA_OK(cudaMemcpyAsync(hb, bst, nq * sizeof(BselState), cudaMemcpyDeviceToHost, s));
    DCR_CUDA_OK(cudaStreamSynchronize(s));
    // ... verified on the device; a quantile whose bracket missed is redone with the 3-pass radix select
    bool any = false;
    for (int k = 0; k < nq; ++k) {
        out[k] = hb[k].result;
        if (hb[k].fallback) {
            any = true;
            int rc = dcr_sel_enqueue(v, q_percent[k], st + k, hist, s);
            if (rc) return rc;
        }
    }
    if (count) *count = hb[0].count;
    if (any) {
        SelState hst[16];
        DCR_CUDA_OK(cudaMemcpyAsync(hst, st, nq * sizeof(SelState), cudaMemcpyDeviceToHost, s));
        DCR_CUDA_OK(cudaStreamSynchronize(s));
        for (int k = 0; k < nq; ++k)
            if (hb[k].fallback) { out[k] = hst[k].result; if (count && k == 0) *count = hst[0].count; }
    }
    return 0;
}

// radix-only variant (tests, and the reference point for profiles)
extern "C" int dcr_select_quantiles_radix(const float* x, const uint8_t* mask, uint32_t reject, int64_t n, int transform,
                                          float center, const double* q_percent, int nq, float* out, int64_t* count,
                                          void* workspace, size_t workspace_bytes, void* stream) {
    DCR_ARG(x && q_percent && out && workspace, "null pointer");
    DCR_ARG(nq >= 1 && nq <= 16, "1 <= nq <= 16");
    DCR_ARG(n >= 0, "n >= 0");
    DCR_ARG(workspace_bytes >= dcr_select_workspace_bytes(), "workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    DcrBump ws(workspace, workspace_bytes);
    unsigned* hist = ws.take<unsigned>(DCR_SEL_HIST_WORDS);
    SelState* st = ws.take<SelState>(16);
    SelView v;
    memset(&v, 0, sizeof(v));
    v.x = x; v.m = mask; v.reject = reject; v.transform = transform; v.center_ptr = nullptr; v.center_imm = center;
    v.n = n; v.n_ptr = nullptr;
    for (int k = 0; k < nq; ++k) {
        int rc = dcr_sel_enqueue(v, q_percent[k], st + k, hist, s);
        if (rc) return rc;
    }
    SelState hst[16];
    DCR_CUDA_OK(cudaMemcpyAsync(hst, st, nq * sizeof(SelState), cudaMemcpyDeviceToHost, s));
    DCR_CUDA_OK(cudaStreamSynchronize(s));
    for (int k = 0; k < nq; ++k) out[k] = hst[k].result;
    if (count) *count = hst[0].count;
    return 0;
}

// ------------------------------------------------------------------------------------------
// K4 moments: count, min, max, sum, sum of squares (FP64), fixed-order two-level reduction
// ------------------------------------------------------------------------------------------
struct MomPartial {
    unsigned long long n;
    double s, ss;
    float mn, mx;
};

__global__ void __launch_bounds__(256) moments_kernel(const float* __restrict__ x, const unsigned char* __restrict__ m,
                                                      unsigned reject, long long n, MomPartial* __restrict__ part) {
    __shared__ MomPartial sp[8];
    unsigned long long cnt = 0;
    double s = 0.0, ss = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    const unsigned rej = m ? reject : 0u;
    dcr_stream1(x, m, n, [&](float v, unsigned mb, long long) {
        if ((mb & rej) || isnan(v)) return;
        ++cnt;
        s += (double)v;
        ss += (double)v * (double)v;
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    });
    cnt = dcr_warp_sum_u64(cnt);
    s = dcr_warp_sum(s);
    ss = dcr_warp_sum(ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        MomPartial q; q.n = cnt; q.s = s; q.ss = ss; q.mn = mn; q.mx = mx;
        sp[threadIdx.x >> 5] = q;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        MomPartial q = sp[0];
        for (int k = 1; k < 8; ++k) {
            q.n += sp[k].n; q.s += sp[k].s; q.ss += sp[k].ss;
            q.mn = fminf(q.mn, sp[k].mn); q.mx = fmaxf(q.mx, sp[k].mx);
        }
        part[blockIdx.x] = q;
    }
}

extern "C" int dcr_moments(const float* x, const uint8_t* mask, uint32_t reject, int64_t n, double out[6],
                           void* workspace, size_t workspace_bytes, void* stream) {
    DCR_ARG(x && out && workspace, "null pointer");
    const int grid = dcr_num_sms() * 8;
    DCR_ARG(workspace_bytes >= grid * sizeof(MomPartial), "workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    MomPartial* part = (MomPartial*)workspace;
    moments_kernel<<<grid, 256, 0, s>>>(x, mask, reject, n, part);
    DCR_LAUNCH_OK("moments_kernel");
    MomPartial* h = new MomPartial[grid];
    cudaError_t e = cudaMemcpyAsync(h, part, grid * sizeof(MomPartial), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { delete[] h; dcr_set_error("moments copy failed: %s", cudaGetErrorString(e)); return (int)e; }
    unsigned long long cnt = 0; double sum = 0, ss = 0; float mn = INFINITY, mx = -INFINITY;
    for (int k = 0; k < grid; ++k) {
        cnt += h[k].n; sum += h[k].s; ss += h[k].ss; mn = fminf(mn, h[k].mn); mx = fmaxf(mx, h[k].mx);
    }
    delete[] h;
    out[0] = (double)cnt;
    out[1] = cnt ? mn : NAN;
    out[2] = cnt ? mx : NAN;
    const double mean = cnt ? sum / (double)cnt : NAN;
    out[3] = mean;
    double var = cnt ? ss / (double)cnt - mean * mean : NAN;
    if (var < 0) var = 0;
    out[4] = sqrt(var);
    out[5] = ss;
    return 0;
}

// ------------------------------------------------------------------------------------------
// K5 strided subsample of the compressed (row-major, unmasked) array
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) chunk_count_kernel(const unsigned char* __restrict__ m, unsigned reject,
                                                          const float* __restrict__ x, long long n,
                                                          unsigned* __restrict__ chunk_cnt) {
    __shared__ unsigned sw[8];
    const long long base = (long long)blockIdx.x * CHUNK + threadIdx.x * 8;
    unsigned c = 0;
    float v[8];
    unsigned mb[8];
    dcr_load8(x, m, base, n, v, mb);
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (!(mb[k] & (reject | 0xffffff00u)) && !isnan(v[k])) ++c;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int k = 0; k < 8; ++k) t += sw[k];
        chunk_cnt[blockIdx.x] = t;
    }
}

// single block: exclusive scan of the chunk counts; total -> *total_out
__global__ void __launch_bounds__(1024) chunk_scan_kernel(const unsigned* __restrict__ cnt, long long nchunks,
                                                          long long* __restrict__ off, long long* total_out) {
    __shared__ unsigned s_scratch[33];
    __shared__ long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    // 16 chunk counts per thread and round (16384 per round): one block scan + two barriers per round instead of per 1024
    for (long long b = 0; b < nchunks; b += 16384) {
        const long long i0 = b + (long long)threadIdx.x * 16;
        unsigned v[16], sum = 0;
#pragma unroll
        for (int q = 0; q < 16; ++q) { v[q] = (i0 + q < nchunks) ? __ldg(cnt + i0 + q) : 0u; sum += v[q]; }
        unsigned total;
        unsigned ex = dcr_block_excl_scan(sum, s_scratch, &total);
        long long run = s_carry + ex;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            if (i0 + q < nchunks) off[i0 + q] = run;
            run += v[q];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = s_carry;
}

// stride from device (stride_ptr) or immediate; writes element with compressed rank r to dense[r/stride] when r % stride == 0.
// A CTA handles one chunk of 2048 * G elements (off[] holds the exclusive prefix of the per-chunk counts): thread t owns the
// 8 * G consecutive elements starting at t * 8 * G, so one block scan orders the whole chunk.
template <int G>
__global__ void __launch_bounds__(256) compress_strided_kernel(const float* __restrict__ x, const unsigned char* __restrict__ m,
                                                               unsigned reject, long long n,
                                                               const long long* __restrict__ off, long long stride_imm,
                                                               const long long* stride_ptr, float* __restrict__ dense,
                                                               long long cap, const long long* base_ptr) {
    __shared__ unsigned s_scratch[33];
    const long long stride = stride_ptr ? *stride_ptr : stride_imm;
    if (stride_ptr && stride <= 1) return;  // iteration driver: no subsample needed
    const long long base = (long long)blockIdx.x * (CHUNK * G) + threadIdx.x * (8 * G);
    float vals[8 * G];
    unsigned flags = 0, c = 0;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        unsigned mb[8];
        dcr_load8(x, m, base + 8 * g, n, vals + 8 * g, mb);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (!(mb[k] & (reject | 0xffffff00u)) && !isnan(vals[8 * g + k])) { flags |= 1u << (8 * g + k); ++c; }
    }
    unsigned total;
    unsigned ex = dcr_block_excl_scan(c, s_scratch, &total);
    // row-band mode: this band's elements follow `base` unmasked elements of the bands above it
    const long long gbase = base_ptr ? *base_ptr : 0;
    const long long dfirst = (gbase + stride - 1) / stride;  // global subsample index of this band's first pick
    // one 64-bit division per thread; the elements then advance (quotient, remainder) incrementally
    const long long rank0 = gbase + off[blockIdx.x] + ex;
    if (c == 0) return;
    // one 64-bit division per thread; then only the PICKED elements are visited: the j-th unmasked element of this thread
    // (j = 0 .. c-1) has rank rank0 + j, so the picks are j0, j0 + stride, ... with j0 = (-rank0) mod stride, and __fns
    // maps j to its bit in `flags` (a per-element walk cost 35 instructions per element for ~1 pick in 11)
    long long quo = rank0 / stride;
    const long long rem = rank0 - quo * stride;
    quo -= dfirst;
    long long j = 0;
    if (rem != 0) { j = stride - rem; ++quo; }
    for (; j < (long long)c; j += stride, ++quo) {
        const int k = (int)__fns(flags, 0u, (int)j + 1);
        if (quo < cap) dense[quo] = __ldg(x + base + k);   // (a cached re-load: indexing vals[] dynamically would spill it)
    }
}

// groups: 1 -> chunks of 2048 elements (dcr_chunk_count_scan_enqueue), 4 -> chunks of 8192 (the iteration's mask update)
int dcr_compress_enqueue(const float* x, const unsigned char* m, unsigned reject, long long n, const long long* chunk_off,
                         long long stride_imm, const long long* stride_ptr, float* dense, long long cap,
                         cudaStream_t s, const long long* base_ptr, int groups) {
    if (groups == 4) {
        const long long nchunks = (n + 4 * CHUNK - 1) / (4 * CHUNK);
        compress_strided_kernel<4><<<(unsigned)nchunks, 256, 0, s>>>(x, m, reject, n, chunk_off, stride_imm, stride_ptr, dense,
                                                                     cap, base_ptr);
    } else {
        const long long nchunks = (n + CHUNK - 1) / CHUNK;
        compress_strided_kernel<1><<<(unsigned)nchunks, 256, 0, s>>>(x, m, reject, n, chunk_off, stride_imm, stride_ptr, dense,
                                                                     cap, base_ptr);
    }
    DCR_LAUNCH_OK("compress_strided_kernel");
    return 0;
}
int dcr_chunk_count_scan_enqueue(const float* x, const unsigned char* m, unsigned reject, long long n, unsigned* chunk_cnt,
                                 long long* chunk_off, long long* total, cudaStream_t s) {
    const long long nchunks = (n + CHUNK - 1) / CHUNK;
    chunk_count_kernel<<<(unsigned)nchunks, 256, 0, s>>>(m, reject, x, n, chunk_cnt);
    chunk_scan_kernel<<<1, 1024, 0, s>>>(chunk_cnt, nchunks, chunk_off, total);
    DCR_LAUNCH_OK("chunk count/scan");
    return 0;
}
int dcr_chunk_scan_enqueue(const unsigned* cnt, long long nchunks, long long* off, long long* total, cudaStream_t s) {
    chunk_scan_kernel<<<1, 1024, 0, s>>>(cnt, nchunks, off, total);
    DCR_LAUNCH_OK("chunk_scan_kernel");
    return 0;
}
long long dcr_nchunks(long long n) { return (n + CHUNK - 1) / CHUNK; }

extern "C" int dcr_compress_strided(const float* x, const uint8_t* mask, uint32_t reject, int64_t n, int64_t stride,
                                    float* dense, int64_t dense_capacity, int64_t* n_out, void* workspace,
                                    size_t workspace_bytes, void* stream) {
    DCR_ARG(x && dense && workspace && n_out, "null pointer");
    DCR_ARG(stride >= 1, "stride >= 1");
    const long long nchunks = dcr_nchunks(n);
    const size_t need = dcr_al256(nchunks * sizeof(unsigned)) + dcr_al256(nchunks * sizeof(long long)) + 256;
    DCR_ARG(workspace_bytes >= need, "workspace too small");
    if (n == 0) { *n_out = 0; return 0; }
    cudaStream_t s = (cudaStream_t)stream;
    DcrBump ws(workspace, workspace_bytes);
    unsigned* cnt = ws.take<unsigned>(nchunks);
    long long* off = ws.take<long long>(nchunks);
    long long* total = ws.take<long long>(1);
    int rc = dcr_chunk_count_scan_enqueue(x, mask, reject, n, cnt, off, total, s);
    if (rc) return rc;
    rc = dcr_compress_enqueue(x, mask, reject, n, off, stride, nullptr, dense, dense_capacity, s);
    if (rc) return rc;
    long long t = 0;
    DCR_CUDA_OK(cudaMemcpyAsync(&t, total, sizeof(t), cudaMemcpyDeviceToHost, s));
    DCR_CUDA_OK(cudaStreamSynchronize(s));
    *n_out = (t + stride - 1) / stride;
    return 0;
}
