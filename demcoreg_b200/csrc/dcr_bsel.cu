// K3 fast path: exact order statistics by sample-bracketing ("bsel").
//
// The radix select of dcr_select.cu needs three histogram passes over all the data, each limited by
// shared-memory atomics.  Here ONE streaming pass suffices:
//   1. bsel_sample_kernel  - one CTA gathers S = 8192 stratified, hash-jittered samples and localises the keys at the
//                            sample ranks r -+ 5 sigma by three rounds of a 1024-bin shared-memory histogram over a
//                            shrinking key range: a bracket [klo, khi] that contains the wanted order statistic(s)
//                            with overwhelming probability (n <= 8192: the sample is the input, sorted, exact).
//   2. bsel_main_kernel    - the streaming pass (BselCollector, dcr_bsel_collect.cuh): counts the elements below klo
//                            in registers, parks the few per cent inside the bracket in thread-private shared-memory
//                            slots, drains them per warp into a CTA staging buffer + bracket histogram (2048 linear
//                            bins), one global reservation per CTA.  Its LAST CTA locates the histogram bin(s) holding
//                            the exact global rank(s) and VERIFIES that the rank really lies inside the bracket
//                            (exactness does not depend on luck: a miss, an overflow or heavy ties raise `fallback`,
//                            and the caller reruns the 3-pass radix select).
//   3. bsel_refine_kernel  - radix-refines the remaining <= 21 key bits over the candidate buffer only (11 bits per
//                            launch, last-CTA-done finalisation), then applies numpy's linear interpolation.
// Same semantics as dcr_sel_enqueue (numpy.percentile, linear, exact float64 rank).
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include "dcr_common.cuh"
#include "dcr_internal.cuh"
#include "dcr_bsel_collect.cuh"

#define BS_S 8192          // samples
#define BS_THREADS 256

__device__ __forceinline__ void bitonic_sort_smem(unsigned* s, int n /* power of two */) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // j is a power of two
                const int l = i + j;
                const bool up = ((i & k) == 0);
                const unsigned a = s[i], b = s[l];
                if ((a > b) == up) { s[i] = b; s[l] = a; }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ bool bsel_key(const SelView& v, float center, float x, unsigned mb, unsigned* key) {
    if (mb & v.reject) return false;
    if (v.transform == 1) x = fabsf(x - center);
    if (isnan(x)) return false;
    *key = dcr_f2key(x + 0.0f);  // -0 -> +0: float comparisons (main pass) and key order must agree
    return true;
}

// (defined with the row-band code below: the gather half and the localisation half of the sampling step as two kernels)
__global__ void bsel_band_sample_kernel(const SelView v, BselState* st, float* publish, unsigned* __restrict__ keys, int rank,
                                        int nranks, int S);
__global__ void bsel_band_bracket_kernel(const unsigned* __restrict__ keys, double q_percent, BselState* st);

__global__ void __launch_bounds__(1024) bsel_sample_kernel(const SelView v, double q_percent, BselState* st, float* publish) {
    __shared__ unsigned s[BS_S];
    __shared__ unsigned s_valid;
    const long long n = v.n_ptr ? *v.n_ptr : v.n;
    const float center = v.center_ptr ? *v.center_ptr : v.center_imm;
    if (threadIdx.x == 0) s_valid = 0;
    __syncthreads();
    unsigned nv = 0;
    const bool all = n <= BS_S;  // small input: the "sample" is the whole array and the answer is exact
    {
        // all 8 (value, mask) gathers of a thread are issued before their first use: the positions are scattered
        // over the whole array (TLB + DRAM latency ~2 us each), so they must overlap
        float xv[BS_S / 1024];
        unsigned mv[BS_S / 1024];
#pragma unroll
        for (int q = 0; q < BS_S / 1024; ++q) {
            const int k = q * 1024 + threadIdx.x;
            // stratified positions with a hashed offset inside each stratum: a plain stride of n/S aliases with the
            // raster width (8192 x 8192: every sample would sit in column 0, where the Horn border is masked)
            long long i = k;
            if (!all) {
                const unsigned long long lo = ((unsigned long long)k * (unsigned long long)n) / BS_S;
                const unsigned long long hi = ((unsigned long long)(k + 1) * (unsigned long long)n) / BS_S;
                const unsigned long long hsh = ((unsigned long long)k * 0x9E3779B97F4A7C15ull) >> 20;
                i = (long long)(lo + hsh % (hi - lo));
            }
            xv[q] = 0.f;
            mv[q] = 0xFFFFFFFFu;
            if (i < n) {
                xv[q] = v.x[i];
                mv[q] = v.m ? (unsigned)v.m[i] : 0u;
            }
        }
#pragma unroll
        for (int q = 0; q < BS_S / 1024; ++q) {
            unsigned key = 0xFFFFFFFFu, kk;
            if (mv[q] <= 0xffu && bsel_key(v, center, xv[q], mv[q], &kk)) { key = kk; ++nv; }
            s[q * 1024 + threadIdx.x] = key;
        }
    }
    nv = __reduce_add_sync(0xffffffffu, nv);
    if ((threadIdx.x & 31) == 0 && nv) atomicAdd(&s_valid, nv);
    __syncthreads();
    if (threadIdx.x == 0) {
        st->fallback = 0;
        st->done = 0;
        st->empty = 0;
        st->below = 0; st->total = 0; st->ncand = 0; st->nsmall = 0; st->ticket = 0;
        st->publish = publish;
    }
    const long long m = s_valid;
    if (all) {
        // small input: sort everything, the answer is exact
        bitonic_sort_smem(s, BS_S);
        if (threadIdx.x == 0) {
            st->done = 1;
            st->count = m;
            if (m == 0) {
                st->empty = 1;
                st->v_lo = st->v_hi = st->result = __int_as_float(0x7fc00000);
            } else {
                long long klo, khi; double g;
                bsel_ranks(m, q_percent, &klo, &khi, &g);
                st->gamma = g;
                st->v_lo = dcr_key2f(s[klo]); st->v_hi = dcr_key2f(s[khi]);
                st->result = dcr_lerp32(st->v_lo, st->v_hi, g);
            }
            bsel_publish(st);
        }
        return;
    }
    // Bracket = keys around the sample ranks r -+ 5 sigma.  Instead of sorting 8192 keys on one SM (~45 us,
    // instruction bound) the two ranks are localised by three rounds of a 1024-bin histogram over a shrinking
    // key range (shared-memory atomics + one block scan per round, ~2 us each).  The bracket edges are bin
    // edges, marginally wider than the sample values themselves -- exactness is verified later either way.
    __shared__ unsigned hist[1024];
    __shared__ unsigned s_scratch[33];
    __shared__ unsigned s_kmin, s_kmax;
    __shared__ unsigned s_bin[2], s_ex[2];
    if (threadIdx.x == 0) { s_kmin = 0xFFFFFFFFu; s_kmax = 0u; }
    __syncthreads();
    {
        unsigned kmn = 0xFFFFFFFFu, kmx = 0u;
#pragma unroll
        for (int q = 0; q < BS_S / 1024; ++q) {
            const unsigned k = s[q * 1024 + threadIdx.x];
            if (k != 0xFFFFFFFFu) { kmn = min(kmn, k); kmx = max(kmx, k); }
        }
        kmn = __reduce_min_sync(0xffffffffu, kmn);
        kmx = __reduce_max_sync(0xffffffffu, kmx);
        if ((threadIdx.x & 31) == 0) { atomicMin(&s_kmin, kmn); atomicMax(&s_kmax, kmx); }
    }
    __syncthreads();
    bool open_lo = true, open_hi = true;
    long long ra = 0, rb = 0;
    if (m > 0) {
        const double f = q_percent / 100.0;
        const double r = f * (double)(m - 1);
        const double d = 5.0 * sqrt((double)m * f * (1.0 - f)) + 8.0;
        const long long a0 = (long long)floor(r - d), b0 = (long long)ceil(r + d);
        open_lo = a0 < 0;
        open_hi = b0 >= m;
        ra = open_lo ? 0 : a0;
        rb = open_hi ? m - 1 : b0;
    }
    unsigned base = s_kmin;
    unsigned long long span = m > 0 ? (unsigned long long)s_kmax - s_kmin + 1ull : 1ull;
    for (int round = 0; round < 3 && m > 0; ++round) {
        int shift = 0;
        while ((span >> shift) > 1024ull) ++shift;
        hist[threadIdx.x] = 0u;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < BS_S / 1024; ++q) {
            const unsigned k = s[q * 1024 + threadIdx.x];
            if (k != 0xFFFFFFFFu && k >= base && (unsigned long long)(k - base) < span) atomicAdd(&hist[(k - base) >> shift], 1u);
        }
        __syncthreads();
        const unsigned hv = hist[threadIdx.x];
        unsigned tot;
        const unsigned ex = dcr_block_excl_scan(hv, s_scratch, &tot);
        if ((long long)ex <= ra && ra < (long long)ex + hv) { s_bin[0] = threadIdx.x; s_ex[0] = ex; }
        if ((long long)ex <= rb && rb < (long long)ex + hv) { s_bin[1] = threadIdx.x; s_ex[1] = ex; }
        __syncthreads();
        const unsigned long long nb = (unsigned long long)base + ((unsigned long long)s_bin[0] << shift);
        unsigned long long nt = (unsigned long long)base + (((unsigned long long)s_bin[1] + 1ull) << shift);  // exclusive
        if (nt > (unsigned long long)base + span) nt = (unsigned long long)base + span;
        ra -= s_ex[0];
        rb -= s_ex[0];
        base = (unsigned)nb;
        span = nt - nb;
        __syncthreads();
        if (shift == 0) break;
    }
    if (threadIdx.x == 0) {
        unsigned klo = dcr_f2key(-INFINITY), khi = dcr_f2key(INFINITY);
        if (m > 0) {
            if (!open_lo) klo = base;
            if (!open_hi) khi = (unsigned)((unsigned long long)base + span - 1ull);
        }
        st->klo = klo; st->khi = khi;
        st->flo = dcr_key2f(klo); st->fhi = dcr_key2f(khi);
        // linear bins of the bracket: (key - klo) >> shift < 2048
        const unsigned long long bspan = (unsigned long long)khi - klo + 1ull;
        int shift = 0;
        while ((bspan >> shift) > 2048ull) ++shift;
        st->shift = shift;
    }
}

// One streaming pass: the collector (dcr_bsel_collect.cuh) does the counting / candidate gathering / level-0 find.
typedef BselCollector<16, 4096> MainCollector;
template <int TRANSFORM>
__global__ void __launch_bounds__(BS_THREADS) bsel_main_kernel(const SelView v, BselState* st, unsigned* ghist,
                                                               unsigned* __restrict__ cand, unsigned cap, double q_percent,
                                                               int band) {
    extern __shared__ __align__(16) unsigned char bsel_smem[];
    MainCollector::Shared* sh = reinterpret_cast<MainCollector::Shared*>(bsel_smem);
    __shared__ unsigned s_scr[33];
    if (st->done) return;
    MainCollector col;
    col.init(sh, st, cand, cap);
    __syncthreads();
    const long long n = v.n_ptr ? *v.n_ptr : v.n;
    const float center = v.center_ptr ? *v.center_ptr : v.center_imm;
    const unsigned reject = v.m ? v.reject : 0u;
    dcr_stream1_groups(v.x, v.m, n,
        [&](float x, unsigned mb, long long, int) {
            if (TRANSFORM == 1) x = fabsf(x - center);
            col.add(x, !(mb & reject));
        },
        [&]() { col.drain(); });
    col.finish(ghist, q_percent, gridDim.x, s_scr, band);
}

// Refinement over the candidate buffer only (a few per cent of the data): histogram the next <= 11 bits of
// the candidates sharing the resolved prefix of each rank; the last CTA to finish locates the sub-bins.
// Ties of any multiplicity are handled (the final level has one bin per distinct key).
// mode 0: histogram + find by the last CTA (single GPU); row-band mode splits the two around the all-reduce of ghist:
// mode 1 = histogram of this rank's candidates only, mode 2 = find only (one CTA, on the summed histogram)
__global__ void __launch_bounds__(BS_THREADS) bsel_refine_kernel(BselState* st, const unsigned* __restrict__ cand,
                                                                 unsigned* ghist /* [2][2048] */, int mode) {
    __shared__ bool s_last;
    if (st->done) return;
    const unsigned ncand = st->ncand;
    const unsigned klo = st->klo;
    const int cs = st->cur_shift;
    const int ns = cs > 11 ? cs - 11 : 0;
    const unsigned smask = (1u << (cs - ns)) - 1u;
    const unsigned p0 = st->prefix[0] >> cs, p1 = st->prefix[1] >> cs;
    if (mode != 2) {
        const unsigned step = gridDim.x * BS_THREADS;
        for (unsigned i0 = blockIdx.x * BS_THREADS + threadIdx.x; i0 < ncand; i0 += 4 * step) {
            unsigned kv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) kv[u] = (i0 + u * step < ncand) ? __ldg(cand + i0 + u * step) : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i0 + u * step >= ncand) break;
                const unsigned off = kv[u] - klo;
                const unsigned pre = off >> cs;
                if (pre == p0) atomicAdd(&ghist[(off >> ns) & smask], 1u);
                if (p1 != p0 && pre == p1) atomicAdd(&ghist[2048 + ((off >> ns) & smask)], 1u);
            }
        }
    }
    if (mode == 1) return;
    if (mode == 0) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
        __syncthreads();
        if (!s_last) return;
        __threadfence();
    }
    {
        __shared__ unsigned s_scratch[33];
        __shared__ unsigned s_sel[2];
        __shared__ long long s_cumsel[2];
        for (int r = 0; r < 2; ++r) {
            const unsigned* h = ghist + ((r == 1 && p1 != p0) ? 2048 : 0);
            const long long k = st->rem[r];
            unsigned vv[8], sum = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { vv[q] = __ldcg(h + threadIdx.x * 8 + q); sum += vv[q]; }
            unsigned tot;
            long long ex = dcr_block_excl_scan(sum, s_scratch, &tot);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (ex <= k && k < ex + vv[q]) { s_sel[r] = threadIdx.x * 8 + q; s_cumsel[r] = ex; }
                ex += vv[q];
            }
            __syncthreads();
        }
        if (threadIdx.x < 2) {
            const int r = threadIdx.x;
            st->prefix[r] |= s_sel[r] << ns;
            st->rem[r] = st->rem[r] - s_cumsel[r];
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 4096; k += BS_THREADS) ghist[k] = 0u;
    if (threadIdx.x == 0) {
        st->cur_shift = ns;
        st->ticket = 0;
        if (ns == 0) {
            const float x0 = dcr_key2f(klo + st->prefix[0]), x1 = dcr_key2f(klo + st->prefix[1]);
            st->v_lo = x0; st->v_hi = x1;
            st->result = dcr_lerp32(x0, x1, st->gamma);
            st->done = 1;
            bsel_publish(st);
        }
    }
}

size_t dcr_bsel_scratch_bytes(long long n) {
    return dcr_al256(4096 * sizeof(unsigned)) + dcr_al256(dcr_bsel_cap(n) * sizeof(unsigned));
}
unsigned dcr_bsel_cap(long long n) {
    long long c = n / 6;
    if (c < 65536) c = 65536;
    if (c > 0x7fffffffLL) c = 0x7fffffffLL;
    return (unsigned)c;
}

// scratch layout: [4096 u32 histogram][cap u32 candidate keys]
void dcr_bsel_buffers(void* scratch, long long n_max, unsigned** hist, unsigned** cand, unsigned* cap) {
    char* p = (char*)scratch;
    *hist = (unsigned*)p;
    *cand = (unsigned*)(p + dcr_al256(4096 * sizeof(unsigned)));
    *cap = dcr_bsel_cap(n_max);
}

// step 1 of a selection whose streaming pass is fused into another kernel (BselCollector): clear + sample + bracket
int dcr_bsel_sample_enqueue(const SelView& v, double q_percent, BselState* st, void* scratch, long long n_max, cudaStream_t s,
                            float* publish) {
    unsigned *hist, *cand, cap;
    dcr_bsel_buffers(scratch, n_max, &hist, &cand, &cap);
    DCR_CUDA_OK(cudaMemsetAsync(hist, 0, 4096 * sizeof(unsigned), s));
    if (n_max > BS_S) {
        // gather spread over 8 CTAs (one scattered load per thread) + the localisation kernel: 4 + 11 us instead of 23 us for
        // the fused single-CTA kernel, which stays for views that may fit the sample buffer whole (exact by sorting)
        unsigned* keys = cand + (65536 - BS_S);     // (cap >= 65536; candidates are appended from the front, later)
        bsel_band_sample_kernel<<<BS_S / 1024, 1024, 0, s>>>(v, st, publish, keys, 0, 1, BS_S);
        bsel_band_bracket_kernel<<<1, 1024, 0, s>>>(keys, q_percent, st);
        DCR_LAUNCH_OK("bsel sample kernels");
        return 0;
    }
    bsel_sample_kernel<<<1, 1024, 0, s>>>(v, q_percent, st, publish);
    DCR_LAUNCH_OK("bsel_sample_kernel");
    return 0;
}
// step 3: radix refinement over the candidates gathered by the collector
int dcr_bsel_refine_enqueue(BselState* st, void* scratch, long long n_max, cudaStream_t s) {
    unsigned *hist, *cand, cap;
    dcr_bsel_buffers(scratch, n_max, &hist, &cand, &cap);
    bsel_refine_kernel<<<dcr_num_sms() * 4, BS_THREADS, 0, s>>>(st, cand, hist, 0);  // shift -> max(shift-11, 0)
    bsel_refine_kernel<<<dcr_num_sms() * 4, BS_THREADS, 0, s>>>(st, cand, hist, 0);  // -> 0 (returns at once if done)
    DCR_LAUNCH_OK("bsel_refine_kernel");
    return 0;
}

int dcr_bsel_enqueue(const SelView& v, double q_percent, BselState* st, void* scratch, long long n_max, cudaStream_t s,
                     float* publish) {
    unsigned *hist, *cand, cap;
    dcr_bsel_buffers(scratch, n_max, &hist, &cand, &cap);
    int rc = dcr_bsel_sample_enqueue(v, q_percent, st, scratch, n_max, s, publish);
    if (rc) return rc;
    DcrCtx* ctx = dcr_ctx();
    if (!ctx) return -1;
    if (!(ctx->attrs & DCR_ATTR_BSEL)) {   // per device
        DCR_CUDA_OK(cudaFuncSetAttribute(bsel_main_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(MainCollector::Shared)));
        DCR_CUDA_OK(cudaFuncSetAttribute(bsel_main_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(MainCollector::Shared)));
        ctx->attrs |= DCR_ATTR_BSEL;
    }
    const int grid0 = ctx->sms * 8;   // SMs x 8, like the other streaming kernels
    // a CTA stages up to 4096 in-bracket keys (~6 % of what it streams): keep its share below 57k elements
    long long grid = (n_max + 57343) / 57344;
    if (grid < grid0) grid = grid0;
    if (grid > (1 << 20)) grid = 1 << 20;
    if (v.transform == 1)
        bsel_main_kernel<1><<<(unsigned)grid, BS_THREADS, sizeof(MainCollector::Shared), s>>>(v, st, hist, cand, cap, q_percent, 0);
    else
        bsel_main_kernel<0><<<(unsigned)grid, BS_THREADS, sizeof(MainCollector::Shared), s>>>(v, st, hist, cand, cap, q_percent, 0);
    DCR_LAUNCH_OK("bsel_main_kernel");
    return dcr_bsel_refine_enqueue(st, scratch, n_max, s);
}

// ------------------------------------------------------------------------------------------
// Row-band mode: the same selection over a view that is split across ranks.  The one streaming pass stays; what the single
// GPU does with last-CTA hand-offs is cut at the points where the ranks must agree, and the small buffers in between are
// summed over the ranks by the caller (ncclAllReduce / in-process sum):
//   sub 0  sample     this rank gathers its share of the S = 8192 stratified sample keys (strata k in [S r / R, S (r+1) / R)
//                     over its own elements) into a zeroed key array                              -> all-reduce keys[8192]
//   sub 1  bracket    every rank localises the SAME bracket from the complete sample; main pass over its own elements:
//                     below / total counts, bracket histogram, candidates (kept local)   -> all-reduce {below, total}, ghist[2049]
//   sub 2  find       ranks located + bracket verified on the summed histogram (identical on every rank), then the histogram
//                     of the local candidates under the resolved prefix                           -> all-reduce ghist[4096]
//   sub 3  refine     sub-bins located; next 11 bits histogrammed                                 -> all-reduce ghist[4096]
//   sub 4  refine     final value (published); a declined selection leaves st->fallback = 1 on every rank.
// Exactness: the counts, histograms and candidate keys are integers, so every rank computes the same order statistic as the
// single-GPU selection does.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) bsel_band_sample_kernel(const SelView v, BselState* st, float* publish,
                                                                unsigned* __restrict__ keys, int rank, int nranks, int S) {
    const long long n = v.n_ptr ? *v.n_ptr : v.n;
    const float center = v.center_ptr ? *v.center_ptr : v.center_imm;
    const int k0 = (int)(((long long)S * rank) / nranks), k1 = (int)(((long long)S * (rank + 1)) / nranks);
    const int ns = k1 - k0;   // this rank's strata
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->fallback = 0; st->done = 0; st->empty = 0;
        st->below = 0; st->total = 0; st->ncand = 0; st->nsmall = 0; st->ticket = 0; st->ncollect = 0;
        st->publish = publish;
    }
    // one gather per thread, spread over several CTAs: the positions are scattered over the whole array (TLB + DRAM latency
    // ~2 us each), and a single CTA walking its strata eight deep took 23 us
    for (int j = blockIdx.x * 1024 + threadIdx.x; j < ns; j += gridDim.x * 1024) {
        unsigned key = 0xFFFFFFFFu;
        if (n > 0) {
            // hashed position inside stratum j of this rank's elements (a plain stride aliases with the raster width)
            const unsigned long long lo = ((unsigned long long)j * (unsigned long long)n) / (unsigned long long)ns;
            unsigned long long hi = ((unsigned long long)(j + 1) * (unsigned long long)n) / (unsigned long long)ns;
            if (hi <= lo) hi = lo + 1;
            const unsigned long long hsh = ((unsigned long long)(k0 + j) * 0x9E3779B97F4A7C15ull) >> 20;
            const long long i = (long long)(lo + hsh % (hi - lo));
            if (i < n) {
                const unsigned mb = v.m ? (unsigned)v.m[i] : 0u;
                unsigned kk;
                if (bsel_key(v, center, v.x[i], mb, &kk)) key = kk;
            }
        }
        keys[k0 + j] = key;   // (the other ranks' slots stay 0: the sum over ranks assembles the sample)
    }
}

// the localisation half of bsel_sample_kernel on a complete key array (identical input and result on every rank)
__global__ void __launch_bounds__(1024) bsel_band_bracket_kernel(const unsigned* __restrict__ keys, double q_percent, BselState* st) {
    __shared__ unsigned s[BS_S];
    __shared__ unsigned hist[1024];
    __shared__ unsigned s_scratch[33];
    __shared__ unsigned s_kmin, s_kmax, s_valid;
    __shared__ unsigned s_bin[2], s_ex[2];
    if (threadIdx.x == 0) { s_kmin = 0xFFFFFFFFu; s_kmax = 0u; s_valid = 0u; }
    __syncthreads();
    {
        unsigned kmn = 0xFFFFFFFFu, kmx = 0u, nv = 0;
#pragma unroll
        for (int q = 0; q < BS_S / 1024; ++q) {
            const unsigned k = keys[q * 1024 + threadIdx.x];
            s[q * 1024 + threadIdx.x] = k;
            if (k != 0xFFFFFFFFu) { kmn = min(kmn, k); kmx = max(kmx, k); ++nv; }
        }
        kmn = __reduce_min_sync(0xffffffffu, kmn);
        kmx = __reduce_max_sync(0xffffffffu, kmx);
        nv = __reduce_add_sync(0xffffffffu, nv);
        if ((threadIdx.x & 31) == 0) { atomicMin(&s_kmin, kmn); atomicMax(&s_kmax, kmx); atomicAdd(&s_valid, nv); }
    }
    __syncthreads();
    const long long m = s_valid;
    bool open_lo = true, open_hi = true;
    long long ra = 0, rb = 0;
    if (m > 0) {
        const double f = q_percent / 100.0;
        const double r = f * (double)(m - 1);
        const double d = 5.0 * sqrt((double)m * f * (1.0 - f)) + 8.0;
        const long long a0 = (long long)floor(r - d), b0 = (long long)ceil(r + d);
        open_lo = a0 < 0;
        open_hi = b0 >= m;
        ra = open_lo ? 0 : a0;
        rb = open_hi ? m - 1 : b0;
    }
    unsigned base = s_kmin;
    unsigned long long span = m > 0 ? (unsigned long long)s_kmax - s_kmin + 1ull : 1ull;
    for (int round = 0; round < 3 && m > 0; ++round) {
        int shift = 0;
        while ((span >> shift) > 1024ull) ++shift;
        hist[threadIdx.x] = 0u;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < BS_S / 1024; ++q) {
            const unsigned k = s[q * 1024 + threadIdx.x];
            if (k != 0xFFFFFFFFu && k >= base && (unsigned long long)(k - base) < span) atomicAdd(&hist[(k - base) >> shift], 1u);
        }
        __syncthreads();
        const unsigned hv = hist[threadIdx.x];
        unsigned tot;
        const unsigned ex = dcr_block_excl_scan(hv, s_scratch, &tot);
        if ((long long)ex <= ra && ra < (long long)ex + hv) { s_bin[0] = threadIdx.x; s_ex[0] = ex; }
        if ((long long)ex <= rb && rb < (long long)ex + hv) { s_bin[1] = threadIdx.x; s_ex[1] = ex; }
        __syncthreads();
        const unsigned long long nb = (unsigned long long)base + ((unsigned long long)s_bin[0] << shift);
        unsigned long long nt = (unsigned long long)base + (((unsigned long long)s_bin[1] + 1ull) << shift);  // exclusive
        if (nt > (unsigned long long)base + span) nt = (unsigned long long)base + span;
        ra -= s_ex[0];
        rb -= s_ex[0];
        base = (unsigned)nb;
        span = nt - nb;
        __syncthreads();
        if (shift == 0) break;
    }
    if (threadIdx.x == 0) {
        unsigned klo = dcr_f2key(-INFINITY), khi = dcr_f2key(INFINITY);
        if (m > 0) {
            if (!open_lo) klo = base;
            if (!open_hi) khi = (unsigned)((unsigned long long)base + span - 1ull);
        }
        st->klo = klo; st->khi = khi;
        st->flo = dcr_key2f(klo); st->fhi = dcr_key2f(khi);
        const unsigned long long bspan = (unsigned long long)khi - klo + 1ull;
        int shift = 0;
        while ((bspan >> shift) > 2048ull) ++shift;
        st->shift = shift;
    }
}

__global__ void __launch_bounds__(BS_THREADS) bsel_band_find_kernel(BselState* st, unsigned* ghist, double q_percent, unsigned cap) {
    __shared__ unsigned s_cum[2049];
    __shared__ unsigned s_scr[33];
    if (st->done) return;
    bsel_find_block(st, ghist, q_percent, cap, s_cum, s_scr, 1);
}

int dcr_bsel_band_nsub(void) { return 5; }

// Sub-phase `sub` of a band selection; red_ptr / red_count / red_kind / *nred describe what the caller sums over the ranks
// before the next sub-phase (kind 0 = uint32, 1 = int64).
int dcr_bsel_band_phase(const SelView& v, double q_percent, BselState* st, void* scratch, long long n_max, int rank, int nranks,
                        int sub, cudaStream_t s, float* publish, void** red_ptr, int64_t* red_count, int* red_kind, int* nred) {
    unsigned *hist, *cand, cap;
    dcr_bsel_buffers(scratch, n_max, &hist, &cand, &cap);
    DcrCtx* ctx = dcr_ctx();
    if (!ctx) return -1;
    // the sample keys live at the END of the candidate buffer's first 64 K words (cap >= 65536): candidates are appended
    // from the front only after the bracket kernel has consumed the sample
    unsigned* keys = cand + (65536 - BS_S);
    auto want = [&](void* p, int64_t cnt, int kind) { red_ptr[*nred] = p; red_count[*nred] = cnt; red_kind[*nred] = kind; ++*nred; };
    if (sub == 0) {
        DCR_CUDA_OK(cudaMemsetAsync(keys, 0, BS_S * sizeof(unsigned), s));
        DCR_CUDA_OK(cudaMemsetAsync(hist, 0, 4096 * sizeof(unsigned), s));
        bsel_band_sample_kernel<<<BS_S / 1024, 1024, 0, s>>>(v, st, publish, keys, rank, nranks, BS_S);
        want(keys, BS_S, 0);
    } else if (sub == 1) {
        bsel_band_bracket_kernel<<<1, 1024, 0, s>>>(keys, q_percent, st);
        if (!(ctx->attrs & DCR_ATTR_BSEL)) {   // per device
            DCR_CUDA_OK(cudaFuncSetAttribute(bsel_main_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(MainCollector::Shared)));
            DCR_CUDA_OK(cudaFuncSetAttribute(bsel_main_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(MainCollector::Shared)));
            ctx->attrs |= DCR_ATTR_BSEL;
        }
        long long grid = (n_max + 57343) / 57344;
        if (grid < ctx->sms * 8) grid = ctx->sms * 8;
        if (grid > (1 << 20)) grid = 1 << 20;
        if (v.transform == 1)
            bsel_main_kernel<1><<<(unsigned)grid, BS_THREADS, sizeof(MainCollector::Shared), s>>>(v, st, hist, cand, cap, q_percent, 1);
        else
            bsel_main_kernel<0><<<(unsigned)grid, BS_THREADS, sizeof(MainCollector::Shared), s>>>(v, st, hist, cand, cap, q_percent, 1);
        want(&st->below, 2, 1);
        want(hist, 2049, 0);
    } else if (sub == 2) {
        bsel_band_find_kernel<<<1, BS_THREADS, 0, s>>>(st, hist, q_percent, cap);
        bsel_refine_kernel<<<dcr_num_sms() * 4, BS_THREADS, 0, s>>>(st, cand, hist, 1);
        want(hist, 4096, 0);
    } else if (sub == 3) {
        bsel_refine_kernel<<<1, BS_THREADS, 0, s>>>(st, cand, hist, 2);
        bsel_refine_kernel<<<dcr_num_sms() * 4, BS_THREADS, 0, s>>>(st, cand, hist, 1);
        want(hist, 4096, 0);
    } else {
        bsel_refine_kernel<<<1, BS_THREADS, 0, s>>>(st, cand, hist, 2);
    }
    DCR_LAUNCH_OK("band selection kernels");
    return 0;
}

// ------------------------------------------------------------------------------------------
// Median AND MAD in one streaming pass ("msel"): filtlib.mad_fltr needs med = P50(x) and P50(|x - med|) -- two dependent
// selections, i.e. two passes over the data with the one-pass select above.  Here the second bracket is taken around a
// centre ESTIMATE chat and widened by the uncertainty of the median, so that one pass can count / collect for both:
//   sample    65536 stratified keys (the usual gather, 64 CTAs)
//   bracket   bracket 0 = [flo0, fhi0] around P50 of the sample (ranks -+ 5 sigma);  chat = its midpoint,
//             w = max(fhi0 - chat, chat - flo0) >= |med - chat| for every median inside the bracket;
//             bracket 1 = [dlo, dhi] around P50 of |s - chat|;  outer bracket [dlo - w, dhi + w] (with a relative slack of
//             1e-6 against float32 rounding).  For ANY centre c inside bracket 0:  D = |x - chat| < dlo - w  implies
//             |x - c| < dlo, and D > dhi + w implies |x - c| > dhi -- so the pass may count the former as "below" and drop
//             the latter without knowing the median
//   main      one pass: the usual collector for bracket 0; a raw collector keeps x itself for dlo - w <= D <= dhi + w
//   refine 0  -> med (exact), published
//   rekey     the raw candidates get their true keys |x - med|: below dlo -> counted, inside [dlo, dhi] -> histogram +
//             compacted key list, above -> dropped
//   find + refine 1  -> the exact MAD, verified like every bracket (rank inside [dlo, dhi], nothing dropped)
// The 65536-key sample keeps the candidates of both brackets together at ~7 % of the data (an 8192-key sample would make
// the widened MAD bracket alone 15 %).
// ------------------------------------------------------------------------------------------
#define MS_S 65536
#define MS_STAGE 1536    // staged candidates per CTA and bracket: 55.4 KB of shared memory per CTA, 4 CTAs per SM

// localise the keys of ranks ra / rb among the valid keys key_of(i), i < S, inside [kmin, kmax]: three rounds of a 1024-bin
// histogram over a shrinking range -> [base, base + span)
template <typename KeyFn>
__device__ __forceinline__ void msel_localise(KeyFn key_of, int S, long long m, unsigned kmin, unsigned kmax, long long ra,
                                              long long rb, unsigned* hist, unsigned* s_scratch, unsigned* s_bin, unsigned* s_ex,
                                              unsigned* base_out, unsigned long long* span_out) {
    unsigned base = kmin;
    unsigned long long span = m > 0 ? (unsigned long long)kmax - kmin + 1ull : 1ull;
    for (int round = 0; round < 3 && m > 0; ++round) {
        int shift = 0;
        while ((span >> shift) > 1024ull) ++shift;
        hist[threadIdx.x] = 0u;
        __syncthreads();
        // (the keys live in global memory / L2: 16 loads in flight per thread -- one at a time made this kernel 85 us)
        for (int i0 = threadIdx.x; i0 < S; i0 += 16 * 1024) {
            unsigned kk[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) kk[u] = (i0 + u * 1024 < S) ? key_of(i0 + u * 1024) : 0xFFFFFFFFu;
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                // (warp-aggregating these atomics with __match_any_sync made the kernel 60 us SLOWER, although the first
                //  round puts most of the sample into a handful of bins: profiles/experiments/README.md)
                const unsigned k = kk[u];
                if (k != 0xFFFFFFFFu && k >= base && (unsigned long long)(k - base) < span) atomicAdd(&hist[(k - base) >> shift], 1u);
            }
        }
        __syncthreads();
        const unsigned hv = hist[threadIdx.x];
        unsigned tot;
        const unsigned ex = dcr_block_excl_scan(hv, s_scratch, &tot);
        if ((long long)ex <= ra && ra < (long long)ex + hv) { s_bin[0] = threadIdx.x; s_ex[0] = ex; }
        if ((long long)ex <= rb && rb < (long long)ex + hv) { s_bin[1] = threadIdx.x; s_ex[1] = ex; }
        __syncthreads();
        const unsigned long long nb = (unsigned long long)base + ((unsigned long long)s_bin[0] << shift);
        unsigned long long nt = (unsigned long long)base + (((unsigned long long)s_bin[1] + 1ull) << shift);  // exclusive
        if (nt > (unsigned long long)base + span) nt = (unsigned long long)base + span;
        ra -= s_ex[0];
        rb -= s_ex[0];
        base = (unsigned)nb;
        span = nt - nb;
        __syncthreads();
        if (shift == 0) break;
    }
    *base_out = base;
    *span_out = span;
}

__global__ void __launch_bounds__(1024) msel_bracket_kernel(const unsigned* __restrict__ keys, BselState* st0, BselState* st1) {
    __shared__ unsigned hist[1024];
    __shared__ unsigned s_scratch[33];
    __shared__ unsigned s_kmin, s_kmax, s_valid;
    __shared__ unsigned s_bin[2], s_ex[2];
    __shared__ float s_chat;
    auto minmax = [&](auto key_of, long long* m_out, unsigned* kmin_out, unsigned* kmax_out) {
        if (threadIdx.x == 0) { s_kmin = 0xFFFFFFFFu; s_kmax = 0u; s_valid = 0u; }
        __syncthreads();
        unsigned kmn = 0xFFFFFFFFu, kmx = 0u, nv = 0;
        for (int i0 = threadIdx.x; i0 < MS_S; i0 += 16 * 1024) {
            unsigned kk[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) kk[u] = (i0 + u * 1024 < MS_S) ? key_of(i0 + u * 1024) : 0xFFFFFFFFu;
#pragma unroll
            for (int u = 0; u < 16; ++u)
                if (kk[u] != 0xFFFFFFFFu) { kmn = min(kmn, kk[u]); kmx = max(kmx, kk[u]); ++nv; }
        }
        kmn = __reduce_min_sync(0xffffffffu, kmn);
        kmx = __reduce_max_sync(0xffffffffu, kmx);
        nv = __reduce_add_sync(0xffffffffu, nv);
        if ((threadIdx.x & 31) == 0) { atomicMin(&s_kmin, kmn); atomicMax(&s_kmax, kmx); atomicAdd(&s_valid, nv); }
        __syncthreads();
        *m_out = s_valid; *kmin_out = s_kmin; *kmax_out = s_kmax;
        __syncthreads();
    };
    auto ranks = [&](long long m, long long* ra, long long* rb, bool* open) {
        const double r = 0.5 * (double)(m - 1);
        const double d = 5.0 * sqrt((double)m * 0.25) + 8.0;
        const long long a0 = (long long)floor(r - d), b0 = (long long)ceil(r + d);
        *open = (a0 < 0) || (b0 >= m);
        *ra = a0 < 0 ? 0 : a0;
        *rb = b0 >= m ? m - 1 : b0;
    };
    // ---- bracket 0: the median of x
    auto key0 = [&](int i) { return __ldg(keys + i); };
    long long m; unsigned kmin, kmax;
    minmax(key0, &m, &kmin, &kmax);
    bool open0 = true;
    long long ra = 0, rb = 0;
    if (m > 0) ranks(m, &ra, &rb, &open0);
    unsigned base; unsigned long long span;
    msel_localise(key0, MS_S, m, kmin, kmax, ra, rb, hist, s_scratch, s_bin, s_ex, &base, &span);
    const unsigned klo0 = base, khi0 = (unsigned)((unsigned long long)base + span - 1ull);
    if (threadIdx.x == 0) {
        // (a sample too small to bracket -- tiny or almost fully masked input -- declines: the caller's radix path is exact)
        const bool bad = (m < 1024) || open0;
        st0->klo = klo0; st0->khi = khi0;
        st0->flo = dcr_key2f(klo0); st0->fhi = dcr_key2f(khi0);
        const unsigned long long bspan = (unsigned long long)khi0 - klo0 + 1ull;
        int shift = 0;
        while ((bspan >> shift) > 2048ull) ++shift;
        st0->shift = shift;
        if (bad) { st0->fallback = 1; st0->done = 1; st1->fallback = 1; st1->done = 1; }
        const float flo = dcr_key2f(klo0), fhi = dcr_key2f(khi0);
        s_chat = 0.5f * flo + 0.5f * fhi;
    }
    __syncthreads();
    if (m < 1024 || open0) return;
    // ---- bracket 1: the median of |x - chat|
    const float chat = s_chat;
    auto key1 = [&](int i) {
        const unsigned k = __ldg(keys + i);
        if (k == 0xFFFFFFFFu) return k;
        const float dv = fabsf(dcr_key2f(k) - chat);
        return (dv == dv) ? dcr_f2key(dv + 0.0f) : 0xFFFFFFFFu;
    };
    long long m1; unsigned k1min, k1max;
    minmax(key1, &m1, &k1min, &k1max);
    bool open1 = true;
    long long r1a = 0, r1b = 0;
    if (m1 > 0) ranks(m1, &r1a, &r1b, &open1);
    unsigned base1; unsigned long long span1;
    msel_localise(key1, MS_S, m1, k1min, k1max, r1a, r1b, hist, s_scratch, s_bin, s_ex, &base1, &span1);
    if (threadIdx.x == 0) {
        const unsigned klo1 = base1, khi1 = (unsigned)((unsigned long long)base1 + span1 - 1ull);
        const float dlo = dcr_key2f(klo1), dhi = dcr_key2f(khi1);
        const float flo = dcr_key2f(klo0), fhi = dcr_key2f(khi0);
        float w = fmaxf(fhi - chat, chat - flo);
        w = w * 1.000001f + 1e-30f;
        st1->klo = klo1; st1->khi = khi1;
        // outer bracket in D = |x - chat| (what the streaming pass compares), with slack against float32 rounding
        // (margins relative to the bracket edges themselves: D and the true key are each rounded once, 6e-8 relative)
        float olo = (dlo - w) - 1e-6f * fabsf(dlo) - 1e-30f;
        float ohi = (dhi + w) + 1e-6f * (dhi + w) + 1e-30f;
        st1->flo = olo;      // (olo <= 0: every D >= olo, nothing is counted below by the pass)
        st1->fhi = ohi;
        st1->chat = chat;
        const unsigned long long bspan = (unsigned long long)khi1 - klo1 + 1ull;
        int shift = 0;
        while ((bspan >> shift) > 2048ull) ++shift;
        st1->shift = shift;
        if (open1 || m1 != m || !(w == w) || isinf(w) || isinf(ohi)) { st1->fallback = 1; st1->done = 1; st0->fallback = 1; st0->done = 1; }
    }
}

// The raw half of the pass: elements whose D = |x - chat| lies in the outer bracket are kept AS x (their true key needs the
// exact median); counts like BselCollector (below = D < flo, total = all valid).  One global reservation per CTA.
struct MselRaw {
    struct Shared {
        float priv[16 * BSC_THREADS];
        float buf[MS_STAGE];
        unsigned long long below, total;
        unsigned n, gbase;
    };
    Shared* sh;
    BselState* st;
    float* cand;
    unsigned cap;
    float* mine;
    unsigned below, total, c;
    float flo, fhi;
    __device__ __forceinline__ void init(Shared* s, BselState* state, float* cand_buf, unsigned cand_cap) {
        sh = s; st = state; cand = cand_buf; cap = cand_cap;
        mine = s->priv + threadIdx.x;
        if (threadIdx.x == 0) { s->below = 0ull; s->total = 0ull; s->n = 0u; }
        below = total = c = 0;
        flo = state->flo; fhi = state->fhi;
        if (state->done) { flo = INFINITY; fhi = INFINITY; }
    }
    __device__ __forceinline__ void add(float x, float dval, bool ok) {
        const bool lo = ok && (dval < flo), ge = ok && (dval >= flo), in = ge && (dval <= fhi);
        below += lo;
        total += ge;
        if (in) { mine[c * BSC_THREADS] = x; ++c; }
    }
    __device__ __forceinline__ void drain() {
        const unsigned cmax = __reduce_max_sync(0xffffffffu, c);
        if (cmax == 0) return;
        const int lane = threadIdx.x & 31;
        const unsigned inc = dcr_warp_incl_scan(c);
        const unsigned wt = __shfl_sync(0xffffffffu, inc, 31);
        unsigned base = 0, gb = 0;
        if (lane == 31) {
            base = atomicAdd(&sh->n, wt);
            if (base + wt > MS_STAGE) gb = atomicAdd(&st->ncand, base + wt - (base > MS_STAGE ? base : MS_STAGE));
        }
        base = __shfl_sync(0xffffffffu, base, 31);
        gb = __shfl_sync(0xffffffffu, gb, 31);
        unsigned pos = base + inc - c;
        const unsigned ov0 = base > MS_STAGE ? base : MS_STAGE;
        for (unsigned k = 0; k < cmax; ++k) {
            if (k < c) {
                const float xv = mine[k * BSC_THREADS];
                if (pos < MS_STAGE) sh->buf[pos] = xv;
                else if (gb + (pos - ov0) < cap) cand[gb + (pos - ov0)] = xv;
                ++pos;
            }
        }
        c = 0;
    }
    __device__ __forceinline__ void finish() {
        if (st->done) return;
        const int lane = threadIdx.x & 31;
        drain();
        below = __reduce_add_sync(0xffffffffu, below);
        total = __reduce_add_sync(0xffffffffu, total) + below;
        if (lane == 0) {
            if (below) atomicAdd(&sh->below, (unsigned long long)below);
            if (total) atomicAdd(&sh->total, (unsigned long long)total);
        }
        __syncthreads();
        const unsigned nst = sh->n < MS_STAGE ? sh->n : MS_STAGE;
        if (threadIdx.x == 0) {
            sh->gbase = nst ? atomicAdd(&st->ncand, nst) : 0u;
            if (sh->below) atomicAdd(&st->below, sh->below);
            if (sh->total) atomicAdd(&st->total, sh->total);
        }
        __syncthreads();
        const unsigned gbs = sh->gbase;
        for (unsigned k = threadIdx.x; k < nst; k += BSC_THREADS)
            if (gbs + k < cap) cand[gbs + k] = sh->buf[k];
    }
};

typedef BselCollector<16, MS_STAGE> MselCollector;
struct MselShared { MselCollector::Shared a; MselRaw::Shared b; };

__global__ void __launch_bounds__(BS_THREADS) msel_main_kernel(const SelView v, BselState* st0, BselState* st1, unsigned* ghist0,
                                                               unsigned* __restrict__ cand0, unsigned cap0,
                                                               float* __restrict__ cand1, unsigned cap1) {
    extern __shared__ __align__(16) unsigned char msel_smem[];
    MselShared* sh = reinterpret_cast<MselShared*>(msel_smem);
    __shared__ unsigned s_scr[33];
    if (st0->done) return;
    MselCollector c0;
    MselRaw c1;
    c0.init(&sh->a, st0, cand0, cap0);
    c1.init(&sh->b, st1, cand1, cap1);
    __syncthreads();
    const long long n = v.n_ptr ? *v.n_ptr : v.n;
    const unsigned reject = v.m ? v.reject : 0u;
    const float chat = st1->chat;
    dcr_stream1_groups(v.x, v.m, n,
        [&](float x, unsigned mb, long long, int) {
            const bool ok = !(mb & reject);
            c0.add(x, ok);
            c1.add(x, fabsf(x - chat), ok);
        },
        [&]() { c0.drain(); c1.drain(); });
    c1.finish();
    c0.finish(ghist0, 50.0, gridDim.x, s_scr, 0);
}

// raw candidates -> true keys |x - med|: below the inner bracket counted, inside it histogrammed + compacted, above dropped
__global__ void __launch_bounds__(BS_THREADS) msel_rekey_kernel(BselState* st1, const BselState* st0, const float* __restrict__ raw,
                                                                unsigned cap1, unsigned* __restrict__ keys_out, unsigned cap0,
                                                                unsigned* __restrict__ ghist) {
    __shared__ unsigned h[2048];
    __shared__ unsigned s_scr[33];
    __shared__ unsigned s_base;
    __shared__ unsigned long long s_below;
    if (st1->done || st0->fallback) return;
    const unsigned nraw = st1->ncand < cap1 ? st1->ncand : cap1;
    const float med = st0->result;
    const unsigned klo = st1->klo, khi = st1->khi;
    const int shift = st1->shift;
    for (int k = threadIdx.x; k < 2048; k += BS_THREADS) h[k] = 0u;
    if (threadIdx.x == 0) s_below = 0ull;
    __syncthreads();
    unsigned below = 0;
    const unsigned tile = BS_THREADS * 8;
    for (unsigned t0 = blockIdx.x * tile; t0 < nraw; t0 += gridDim.x * tile) {
        unsigned key[8];
        unsigned cnt = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned i = t0 + u * BS_THREADS + threadIdx.x;
            key[u] = 0xFFFFFFFFu;
            if (i < nraw) {
                const float dv = fabsf(raw[i] - med);          // float32, like np.abs(x - med) on a float32 array
                const unsigned kk = dcr_f2key(dv + 0.0f);
                if (kk < klo) ++below;
                else if (kk <= khi) { key[u] = kk; ++cnt; }
            }
        }
        unsigned tot;
        const unsigned ex = dcr_block_excl_scan(cnt, s_scr, &tot);
        if (threadIdx.x == 0) s_base = tot ? atomicAdd(&st1->ncollect, tot) : 0u;
        __syncthreads();
        unsigned pos = s_base + ex;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (key[u] != 0xFFFFFFFFu) {
                if (pos < cap0) keys_out[pos] = key[u];
                ++pos;
                atomicAdd(&h[(key[u] - klo) >> shift], 1u);
            }
        }
        __syncthreads();
    }
    below = __reduce_add_sync(0xffffffffu, below);
    if ((threadIdx.x & 31) == 0 && below) atomicAdd(&s_below, (unsigned long long)below);
    __syncthreads();
    if (threadIdx.x == 0 && s_below) atomicAdd(&st1->below, s_below);
    for (int k = threadIdx.x; k < 2048; k += BS_THREADS)
        if (h[k]) atomicAdd(&ghist[k], h[k]);
}

__global__ void __launch_bounds__(BS_THREADS) msel_find_kernel(BselState* st1, const BselState* st0, unsigned* ghist, unsigned cap0,
                                                               unsigned cap1) {
    __shared__ unsigned s_cum[2049];
    __shared__ unsigned s_scr[33];
    if (st1->done) return;
    if (st0->fallback) {   // no median: no MAD either
        if (threadIdx.x == 0) { st1->fallback = 1; st1->done = 1; }
        return;
    }
    if (threadIdx.x == 0) {
        const unsigned overflow = (st1->ncand > cap1 || st1->ncollect > cap0) ? 1u : 0u;   // dropped raw candidates / keys
        ghist[2048] = overflow;
        st1->ncand = st1->ncollect < cap0 ? st1->ncollect : cap0;   // the refinement runs over the compacted keys
    }
    __syncthreads();
    __threadfence_block();
    bsel_find_block(st1, ghist, 50.0, cap0, s_cum, s_scr, 1);
}

int dcr_msel_enqueue(const SelView& v, BselState* st0, BselState* st1, void* scratch0, void* scratch1, long long n_max,
                     cudaStream_t s, float* publish0, float* publish1) {
    unsigned *hist0, *cand0, cap0, *hist1, *cand1, cap1;
    dcr_bsel_buffers(scratch0, n_max, &hist0, &cand0, &cap0);
    dcr_bsel_buffers(scratch1, n_max, &hist1, &cand1, &cap1);
    DcrCtx* ctx = dcr_ctx();
    if (!ctx) return -1;
    DCR_CUDA_OK(cudaMemsetAsync(hist0, 0, 4096 * sizeof(unsigned), s));
    DCR_CUDA_OK(cudaMemsetAsync(hist1, 0, 4096 * sizeof(unsigned), s));
    // sample keys in the (still unused) candidate buffer of the MAD: cap >= 65536
    unsigned* keys = cand1;
    bsel_band_sample_kernel<<<MS_S / 1024, 1024, 0, s>>>(v, st0, publish0, keys, 0, 1, MS_S);
    bsel_band_sample_kernel<<<1, 32, 0, s>>>(v, st1, publish1, keys + MS_S, 0, 1, 0);   // (S = 0: initialises st1 only)
    msel_bracket_kernel<<<1, 1024, 0, s>>>(keys, st0, st1);
    static bool attr_done[64];
    if (!attr_done[ctx->dev & 63]) {
        DCR_CUDA_OK(cudaFuncSetAttribute(msel_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MselShared)));
        attr_done[ctx->dev & 63] = true;
    }
    long long grid = (n_max + 20479) / 20480;      // a CTA stages up to 1536 candidates per bracket (~2 % + ~5 % of what it streams)
    if (grid < ctx->sms * 8) grid = ctx->sms * 8;
    if (grid > (1 << 20)) grid = 1 << 20;
    msel_main_kernel<<<(unsigned)grid, BS_THREADS, sizeof(MselShared), s>>>(v, st0, st1, hist0, cand0, cap0,
                                                                            reinterpret_cast<float*>(cand1), cap1);
    DCR_LAUNCH_OK("msel_main_kernel");
    int rc = dcr_bsel_refine_enqueue(st0, scratch0, n_max, s);          // -> the median, published
    if (rc) return rc;
    // the median's candidates are spent: their buffer receives the MAD's compacted keys
    msel_rekey_kernel<<<ctx->sms * 4, BS_THREADS, 0, s>>>(st1, st0, reinterpret_cast<const float*>(cand1), cap1, cand0, cap0, hist1);
    msel_find_kernel<<<1, BS_THREADS, 0, s>>>(st1, st0, hist1, cap0, cap1);
    bsel_refine_kernel<<<dcr_num_sms() * 4, BS_THREADS, 0, s>>>(st1, cand0, hist1, 0);
    bsel_refine_kernel<<<dcr_num_sms() * 4, BS_THREADS, 0, s>>>(st1, cand0, hist1, 0);
    DCR_LAUNCH_OK("msel kernels");
    return 0;
}
