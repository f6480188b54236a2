// K6 Nuth-Kaab aspect-bin statistics, the host fit, K8 apply_z_shift, and the chained
// one-iteration driver.
//
// Replaces reference coreglib.compute_offset_nuth (coreglib.py:201-315: common mask,
// y = dh/tan(deg2rad(slope)), 3-sigma trim, malib.bin_stats count + median over 360 one-degree
// aspect bins, curve_fit of nuth_func), coreglib.apply_z_shift (coreglib.py:42-55) and the
// Nuth branch of dem_align.compute_offset (dem_align.py:62-172).
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <nvtx3/nvToolsExt.h>
#include "dcr_common.cuh"
#include "dcr_internal.cuh"

#define NB DCR_NBINS
#define BIN_INVALID 0xFFFFu

struct NuthDev {
    unsigned long long n_common;
    unsigned long long n_final;
    float mean_y, std_y, rmin, rmax;
    unsigned long long bin_count[NB];
    unsigned prefix_lo[NB], prefix_hi[NB];
    long long k_lo[NB], k_hi[NB];
    int same[NB];
    float bin_med[NB];
    // fast path (linear two-level bracket): fine cell f = (y - rmin) * 65536/(rmax - rmin) in [0, 65535], coarse = f >> 10
    unsigned short fa_cell[2][NB];   // coarse cell of the lower / upper median element
    unsigned short fb_cell[2][NB];   // fine cell of the lower / upper median element
    unsigned fa_rank[2][NB];         // rank of the element inside its coarse cell, then inside its fine cell
    unsigned fast_ncand;             // candidates appended by pass B
    int fast_fallback;               // 1: fast path not applicable / overflow / heavy ties -> 3-pass radix bins
    float fast_scale;
    int fast_pad;
};

struct NuthPartial {
    unsigned long long n;
    double s, ss;
};

__global__ void __launch_bounds__(256) nuth_prepare_kernel(const float* __restrict__ dh, const float* __restrict__ slope,
                                                           const float* __restrict__ aspect,
                                                           const unsigned char* __restrict__ m, unsigned reject, long long n,
                                                           float* __restrict__ y, unsigned short* __restrict__ bin,
                                                           NuthPartial* __restrict__ part) {
    __shared__ NuthPartial sp[8];
    unsigned long long cnt = 0;
    double s = 0.0, ss = 0.0;
    const unsigned rej = m ? reject : 0u;
    auto one = [&](float d, float sl, float a, unsigned mb, float* yo, unsigned* bo) {
        unsigned b = BIN_INVALID;
        float yv = 0.f;
        if (!(mb & rej)) {
            yv = nuth_y(d, sl);
            ++cnt;
            s += (double)yv;
            ss += (double)yv * (double)yv;
            // np.digitize on float32 integer edges 0..360: bin = floor(x); x == 360 joins the last bin
            if (a >= 0.0f && a <= 360.0f) {
                int bi = (int)a;  // a >= 0: truncation == floor
                if (bi > NB - 1) bi = NB - 1;
                b = (unsigned)bi;
            } else {
                b = BIN_INVALID - 1;  // in the common mask (counts for mean/std) but outside the bin range
            }
        }
        *yo = yv;
        *bo = b;
    };
    // 4 consecutive pixels per thread and iteration: 16-byte loads of dh/slope/aspect, 16-byte store of y and
    // 8-byte store of bin (scalar stores throttled the LSU: profiles/sel_r1h, stall_lg_throttle 2.0)
    const bool vec = DCR_AL16(dh) && DCR_AL16(slope) && DCR_AL16(aspect) && DCR_AL4(m) && DCR_AL16(y) && DCR_AL8(bin);
    const long long n4 = vec ? (n >> 2) : 0;
    const long long step = (long long)gridDim.x * 256;
    for (long long q0 = (long long)blockIdx.x * 256 + threadIdx.x; q0 < n4; q0 += 2 * step) {
        float4 vd[2], vs[2], va[2];
        unsigned vm[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long q = q0 + u * step;
            if (q < n4) {
                vd[u] = __ldg(reinterpret_cast<const float4*>(dh) + q);
                vs[u] = __ldg(reinterpret_cast<const float4*>(slope) + q);
                va[u] = __ldg(reinterpret_cast<const float4*>(aspect) + q);
                vm[u] = m ? __ldg(reinterpret_cast<const unsigned*>(m) + q) : 0u;
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long q = q0 + u * step;
            if (q < n4) {
                float4 yo;
                unsigned b0, b1, b2, b3;
                one(vd[u].x, vs[u].x, va[u].x, vm[u] & 0xffu, &yo.x, &b0);
                one(vd[u].y, vs[u].y, va[u].y, (vm[u] >> 8) & 0xffu, &yo.y, &b1);
                one(vd[u].z, vs[u].z, va[u].z, (vm[u] >> 16) & 0xffu, &yo.z, &b2);
                one(vd[u].w, vs[u].w, va[u].w, vm[u] >> 24, &yo.w, &b3);
                reinterpret_cast<float4*>(y)[q] = yo;
                reinterpret_cast<uint2*>(bin)[q] = make_uint2(b0 | (b1 << 16), b2 | (b3 << 16));
            }
        }
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += step) {
        float yo;
        unsigned bo;
        one(dh[i], slope[i], aspect[i], m ? (unsigned)m[i] : 0u, &yo, &bo);
        y[i] = yo;
        bin[i] = (unsigned short)bo;
    }
    cnt = dcr_warp_sum_u64(cnt);
    s = dcr_warp_sum(s);
    ss = dcr_warp_sum(ss);
    if ((threadIdx.x & 31) == 0) { NuthPartial q; q.n = cnt; q.s = s; q.ss = ss; sp[threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        NuthPartial q = sp[0];
        for (int k = 1; k < 8; ++k) { q.n += sp[k].n; q.s += sp[k].s; q.ss += sp[k].ss; }
        part[blockIdx.x] = q;
    }
}

// first level for many partials (one per 2048-element chunk): CTA b reduces a contiguous range, fixed order
#define NUTH_RED_CTAS 64
__global__ void __launch_bounds__(256) nuth_moments_reduce_kernel(const NuthPartial* __restrict__ part, int nparts,
                                                                  NuthPartial* __restrict__ out) {
    __shared__ NuthPartial sp[256];
    const int len = (nparts + NUTH_RED_CTAS - 1) / NUTH_RED_CTAS;
    const int lo = blockIdx.x * len, hi = min(nparts, lo + len);
    NuthPartial q; q.n = 0; q.s = 0; q.ss = 0;
    for (int k0 = lo + threadIdx.x; k0 < hi; k0 += 256 * 4) {
        NuthPartial t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k0 + u * 256;
            if (k < hi) t[u] = part[k];
            else { t[u].n = 0; t[u].s = 0; t[u].ss = 0; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { q.n += t[u].n; q.s += t[u].s; q.ss += t[u].ss; }
    }
    sp[threadIdx.x] = q;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sp[threadIdx.x].n += sp[threadIdx.x + o].n;
            sp[threadIdx.x].s += sp[threadIdx.x + o].s;
            sp[threadIdx.x].ss += sp[threadIdx.x + o].ss;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sp[0];
}

// single thread block: fixed-order reduction of the partials -> mean/std (float32) -> rmin/rmax
__global__ void __launch_bounds__(256) nuth_moments_final_kernel(const NuthPartial* __restrict__ part, int nparts,
                                                                 int remove_outliers, NuthDev* nd) {
    __shared__ NuthPartial sp[256];
    NuthPartial q; q.n = 0; q.s = 0; q.ss = 0;
    // (8 independent loads in flight per thread; the additions keep their order: one partial per chunk = 32k of them
    //  for an 8192^2 raster when the fused mask update produced them)
    for (int k0 = threadIdx.x; k0 < nparts; k0 += 256 * 8) {
        NuthPartial t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = k0 + u * 256;
            if (k < nparts) t[u] = part[k];
            else { t[u].n = 0; t[u].s = 0; t[u].ss = 0; }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { q.n += t[u].n; q.s += t[u].s; q.ss += t[u].ss; }
    }
    sp[threadIdx.x] = q;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sp[threadIdx.x].n += sp[threadIdx.x + o].n;
            sp[threadIdx.x].s += sp[threadIdx.x + o].s;
            sp[threadIdx.x].ss += sp[threadIdx.x + o].ss;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const unsigned long long cnt = sp[0].n;
        nd->n_common = cnt;
        nd->n_final = 0;
        float mean = __int_as_float(0x7fc00000), sd = mean;
        if (cnt) {
            const double mu = sp[0].s / (double)cnt;
            double var = sp[0].ss / (double)cnt - mu * mu;
            if (var < 0) var = 0;
            mean = (float)mu;
            sd = (float)sqrt(var);
        }
        nd->mean_y = mean;
        nd->std_y = sd;
        if (remove_outliers) {
            // coreglib.py:237-241 in float32 (NEP 50: python int 3 is weak)
            const float fs = 3.0f * sd;
            nd->rmin = mean - fs;
            nd->rmax = mean + fs;
        } else {
            nd->rmin = -INFINITY;
            nd->rmax = INFINITY;
        }
    }
    for (int k = threadIdx.x; k < NB; k += 256) {
        nd->bin_count[k] = 0ull;
        nd->prefix_lo[k] = nd->prefix_hi[k] = 0u;
        nd->same[k] = 1;
        nd->bin_med[k] = __int_as_float(0x7fc00000);
    }
    if (threadIdx.x == 0) {
        // set-up of the fast bin path (linear cell map over [rmin, rmax]); degenerate ranges take the radix bins
        const float span = nd->rmax - nd->rmin;
        nd->fast_ncand = 0;
        nd->fast_fallback = 0;
        nd->fast_scale = 0.f;
        if (!(span > 0.0f) || isinf(span) || isnan(span)) nd->fast_fallback = 1;
        else nd->fast_scale = 65536.0f / span;
    }
}

// ghist layout: [2][NB][2048] u32 (second plane = the upper-median state when it left the lower one's bucket)
template <int PASS>
__global__ void __launch_bounds__(256) bin_hist_kernel(const float* __restrict__ y, const unsigned short* __restrict__ bin,
                                                       long long n, NuthDev* nd, unsigned* __restrict__ ghist) {
    __shared__ unsigned s_cnt[NB];
    __shared__ unsigned s_plo[NB], s_phi[NB];
    __shared__ unsigned char s_same[NB];
    for (int k = threadIdx.x; k < NB; k += 256) {
        s_cnt[k] = 0u;
        if (PASS > 0) { s_plo[k] = nd->prefix_lo[k]; s_phi[k] = nd->prefix_hi[k]; s_same[k] = (unsigned char)nd->same[k]; }
    }
    __syncthreads();
    const float rmin = nd->rmin, rmax = nd->rmax;
    constexpr int shift_prev = (PASS == 1) ? 21 : 10;
    constexpr int dshift = (PASS == 0) ? 21 : (PASS == 1 ? 10 : 0);
    constexpr unsigned dmask = (PASS == 2) ? 1023u : 2047u;
    dcr_stream_f_u16(y, bin, n, [&](float yv, unsigned b, long long) {
        if (b >= NB) return;
        if (!(yv >= rmin && yv <= rmax)) return;
        const unsigned key = dcr_f2key(yv);
        if (PASS == 0) {
            atomicAdd(&s_cnt[b], 1u);
            atomicAdd(&ghist[b * 2048u + (key >> 21)], 1u);
        } else {
            if (((key ^ s_plo[b]) >> shift_prev) == 0u) atomicAdd(&ghist[b * 2048u + ((key >> dshift) & dmask)], 1u);
            if (!s_same[b] && ((key ^ s_phi[b]) >> shift_prev) == 0u)
                atomicAdd(&ghist[(NB + b) * 2048u + ((key >> dshift) & dmask)], 1u);
        }
    });
    if (PASS == 0) {
        __syncthreads();
        for (int k = threadIdx.x; k < NB; k += 256)
            if (s_cnt[k]) atomicAdd(&nd->bin_count[k], (unsigned long long)s_cnt[k]);
    }
}

// one block (256 threads) per bin
__device__ void bin_find_bucket(const unsigned* __restrict__ hist, int nb, long long k, unsigned* s_scratch,
                                int* s_bucket, long long* s_before) {
    // 8 bins per thread
    unsigned v[8], sum = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int idx = threadIdx.x * 8 + q;
        v[q] = idx < nb ? hist[idx] : 0u;
        sum += v[q];
    }
    unsigned total;
    long long ex = dcr_block_excl_scan(sum, s_scratch, &total);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (ex <= k && k < ex + v[q]) { *s_bucket = threadIdx.x * 8 + q; *s_before = ex; }
        ex += v[q];
    }
    __syncthreads();
}

template <int PASS>
__global__ void __launch_bounds__(256) bin_find_kernel(NuthDev* nd, unsigned* __restrict__ ghist) {
    __shared__ unsigned s_scratch[33];
    __shared__ int s_b[2];
    __shared__ long long s_before[2];
    const int b = blockIdx.x;
    constexpr int nb = (PASS == 2) ? 1024 : 2048;
    constexpr int dshift = (PASS == 0) ? 21 : (PASS == 1 ? 10 : 0);
    unsigned* h0 = ghist + (size_t)b * 2048u;
    unsigned* h1 = ghist + (size_t)(NB + b) * 2048u;
    const long long cnt = (long long)nd->bin_count[b];
    if (cnt == 0) return;  // histograms of an empty bin are all zero already
    long long klo, khi;
    int same;
    if (PASS == 0) {
        // scipy binned_statistic 'median': elements floor/ceil(j + (n-1)/2) of the sorted bin
        klo = (cnt - 1) >> 1;
        khi = cnt >> 1;
        same = 1;
    } else {
        klo = nd->k_lo[b]; khi = nd->k_hi[b]; same = nd->same[b];
    }
    bin_find_bucket(h0, nb, klo, s_scratch, &s_b[0], &s_before[0]);
    bin_find_bucket(same ? h0 : h1, nb, khi, s_scratch, &s_b[1], &s_before[1]);
    if (threadIdx.x == 0) {
        const unsigned plo = (PASS == 0 ? 0u : nd->prefix_lo[b]) | ((unsigned)s_b[0] << dshift);
        const unsigned phi = (PASS == 0 ? 0u : nd->prefix_hi[b]) | ((unsigned)s_b[1] << dshift);
        nd->prefix_lo[b] = plo; nd->prefix_hi[b] = phi;
        nd->k_lo[b] = klo - s_before[0];
        nd->k_hi[b] = khi - s_before[1];
        nd->same[b] = (plo == phi) ? 1 : 0;
        if (PASS == 2) {
            const float a = dcr_key2f(plo), c = dcr_key2f(phi);
            nd->bin_med[b] = (a + c) / 2.0f;  // (mid_a + mid_b) / 2 in the values' dtype (float32)
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 2048; k += 256) { h0[k] = 0u; h1[k] = 0u; }
}

__global__ void nuth_count_final_kernel(NuthDev* nd) {
    // n_final = sum of the bin counts + in-range samples outside [0,360] (none for gdaldem aspect); one warp
    unsigned long long t = 0;
    for (int k = threadIdx.x; k < NB; k += 32) t += nd->bin_count[k];
    t = dcr_warp_sum_u64(t);
    if (threadIdx.x == 0) nd->n_final = t;
}

// ------------------------------------------------------------------------------------------
// K6 fast path: per-bin count + exact per-bin median in TWO streaming passes with shared-memory atomics.
// All samples of the bin statistics lie in [rmin, rmax] (the 3-sigma trim), so a monotone linear map
//   f = min(65535, (int)((y - rmin) * 65536/(rmax - rmin)))        (float32: monotone non-decreasing in y)
// brackets ranks exactly: pass A histograms the coarse cell f >> 10 (64 cells x 360 bins, 92 KB of shared
// memory per CTA) -> exact per-bin counts and the coarse cell of each bin's two median elements; pass B
// histograms the 1024 fine cells of those coarse cells only (~4 % of the samples, global atomics) and stages the
// same samples as candidates; pass C / D pick the candidates of the final fine cells (a handful per bin) and
// select exactly among their float keys.  Heavy ties / degenerate ranges raise fast_fallback and the caller
// runs the 3-pass radix bins instead.
// ------------------------------------------------------------------------------------------
#define FB_COARSE 64
#define FB_FINE 1024
#define FB_SLOT 512    // keys of one final fine cell that pass D ranks exactly (more -> heavy ties -> radix bins)
#define NUTH_GHIST_WORDS (2 * NB * 2048 + 65536)   // radix histogram [2][NB][2048]; the fast path's scratch lives in it too
#define FB_HB_WORDS (2 * NB * FB_FINE)              // histB [2][NB][1024]
#define FB_HA_OFF FB_HB_WORDS                       // histA [NB][64]
#define FB_SLOT_OFF (FB_HA_OFF + NB * FB_COARSE)    // slots [NB][2][FB_SLOT]
#define FB_SLOTN_OFF (FB_SLOT_OFF + NB * 2 * FB_SLOT)  // slot counters [NB][2]
#define FB_WORDS (FB_SLOTN_OFF + NB * 2)
// row-band mode (the candidates of a fine cell are spread over the ranks): one overflow word, this rank's own slots and slot
// counts, and the table of every rank's slot counts -- all inside the same [2][NB][2048] histogram allocation
#define FB_FLAG_OFF FB_WORDS
#define FB_LSLOT_OFF (FB_WORDS + 16)
#define FB_LSLOTN_OFF (FB_LSLOT_OFF + NB * 2 * FB_SLOT)
#define FB_RCNT_OFF (FB_LSLOTN_OFF + NB * 2)
#define FB_BAND_WORDS (FB_RCNT_OFF + 8 * NB * 2)
static_assert(FB_BAND_WORDS <= NUTH_GHIST_WORDS, "band-mode scratch must fit the histogram allocation");

__device__ __forceinline__ unsigned fb_fine(float y, float rmin, float scale) {
    const float t = (y - rmin) * scale;
    int f = (int)t;
    f = f < 0 ? 0 : (f > 65535 ? 65535 : f);
    return (unsigned)f;
}

__global__ void __launch_bounds__(512) fb_passA_kernel(const float* __restrict__ y, const unsigned short* __restrict__ bin,
                                                       long long n, const NuthDev* nd, unsigned* __restrict__ ghist) {
    extern __shared__ unsigned sh[];  // [NB][64]
    if (nd->fast_fallback) return;
    for (int k = threadIdx.x; k < NB * FB_COARSE; k += blockDim.x) sh[k] = 0u;
    __syncthreads();
    const float rmin = nd->rmin, rmax = nd->rmax, scale = nd->fast_scale;
    dcr_stream_f_u16(y, bin, n, [&](float yv, unsigned b, long long) {
        // one predicate, one (predicated) shared-memory atomic: early returns cost a divergent branch each
        const bool ok = (b < NB) && (yv >= rmin) && (yv <= rmax);
        if (ok) atomicAdd(&sh[b * FB_COARSE + (fb_fine(yv, rmin, scale) >> 10)], 1u);
    });
    __syncthreads();
    unsigned* ha = ghist + FB_HA_OFF;
    for (int k = threadIdx.x; k < NB * FB_COARSE; k += blockDim.x)
        if (sh[k]) atomicAdd(&ha[k], sh[k]);
}

// one thread per bin: exact count, coarse cells of the two median elements
__global__ void fb_findA_kernel(NuthDev* nd, const unsigned* __restrict__ ghist) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= NB || nd->fast_fallback) return;
    const unsigned* ha = ghist + FB_HA_OFF + b * FB_COARSE;
    unsigned long long cnt = 0;
    for (int c = 0; c < FB_COARSE; ++c) cnt += ha[c];
    nd->bin_count[b] = cnt;
    if (cnt == 0) { nd->fa_cell[0][b] = nd->fa_cell[1][b] = 0xFFFF; return; }
    const unsigned long long k[2] = {(cnt - 1) >> 1, cnt >> 1};  // scipy 'median': floor/ceil((n-1)/2)
    for (int r = 0; r < 2; ++r) {
        unsigned long long cum = 0;
        for (int c = 0; c < FB_COARSE; ++c) {
            const unsigned v = ha[c];
            if (k[r] < cum + v) { nd->fa_cell[r][b] = (unsigned short)c; nd->fa_rank[r][b] = (unsigned)(k[r] - cum); break; }
            cum += v;
        }
    }
}

#define FB_STAGE 3072
#define FB_PRIV 8
// Pass B: ~4 % of the samples fall in the median coarse cell(s) of their bin.  They are first parked in
// thread-private shared-memory slots; each warp then drains its slots with a uniform trip count (max per-lane
// count, ~2) -- fine-cell histogram (global atomics, spread over 360 x 1024 counters) + CTA staging of the record.
__global__ void __launch_bounds__(256) fb_passB_kernel(const float* __restrict__ y, const unsigned short* __restrict__ bin,
                                                       long long n, NuthDev* nd, unsigned* __restrict__ ghist,
                                                       unsigned long long* __restrict__ cand, unsigned cap) {
    __shared__ unsigned s_c01[NB];   // coarse cells of the two median elements of each bin, packed (one shared load per sample)
    __shared__ unsigned long long s_buf[FB_STAGE];
    __shared__ unsigned long long s_priv[FB_PRIV * 256];   // thread-private slots: slot k of thread t at [k * 256 + t]
    __shared__ unsigned s_n, s_gbase;
    if (nd->fast_fallback) return;
    for (int k = threadIdx.x; k < NB; k += 256) s_c01[k] = (unsigned)nd->fa_cell[0][k] | ((unsigned)nd->fa_cell[1][k] << 16);
    if (threadIdx.x == 0) s_n = 0u;
    __syncthreads();
    const float rmin = nd->rmin, rmax = nd->rmax, scale = nd->fast_scale;
    const int lane = threadIdx.x & 31;
    unsigned long long* mine = s_priv + threadIdx.x;
    unsigned c = 0;
    auto elem = [&](float yv, unsigned b) {
        // straight-line classification (one rare branch): early returns cost a divergent branch each
        const bool ok = (b < NB) && (yv >= rmin) && (yv <= rmax);
        const unsigned f = fb_fine(yv, rmin, scale);
        const unsigned cc = f >> 10;
        const unsigned c01 = s_c01[b < NB ? b : 0u];
        const unsigned c0 = c01 & 0xffffu, c1 = c01 >> 16;
        if (!(ok && (cc == c0 || cc == c1))) return;
        const unsigned plane = (cc == c0) ? 0u : 1u;
        const unsigned long long rec = ((unsigned long long)dcr_f2key(yv) << 32) | ((unsigned long long)b << 16) | f;
        if (c < FB_PRIV) {
            mine[c * 256] = rec | ((unsigned long long)plane << 31);
            ++c;
        } else {  // > 8 hits among 16 consecutive samples of one thread: unaggregated path
            atomicAdd(&ghist[(plane * NB + b) * FB_FINE + (f & (FB_FINE - 1))], 1u);
            const unsigned g = atomicAdd(&nd->fast_ncand, 1u);
            if (g < cap) cand[g] = rec;
        }
    };
    auto drain = [&]() {
        const unsigned cmax = __reduce_max_sync(0xffffffffu, c);
        if (cmax) {
            const unsigned inc = dcr_warp_incl_scan(c);
            unsigned base = 0;
            if (lane == 31) base = atomicAdd(&s_n, inc);
            base = __shfl_sync(0xffffffffu, base, 31);
            unsigned pos = base + inc - c;
            for (unsigned k = 0; k < cmax; ++k) {
                if (k < c) {
                    unsigned long long rec = mine[k * 256];
                    const unsigned plane = (unsigned)(rec >> 31) & 1u;
                    rec &= ~(1ull << 31);
                    const unsigned f = (unsigned)(rec & 0xFFFFu), b = (unsigned)(rec >> 16) & 0x7FFFu;
                    atomicAdd(&ghist[(plane * NB + b) * FB_FINE + (f & (FB_FINE - 1))], 1u);
                    if (pos < FB_STAGE) s_buf[pos] = rec;
                    else {
                        const unsigned g = atomicAdd(&nd->fast_ncand, 1u);
                        if (g < cap) cand[g] = rec;
                    }
                    ++pos;
                }
            }
        }
        c = 0;
    };
    // same 512-element warp groups as dcr_stream_f_u16, with a warp-uniform drain after every group
    {
        const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
        const long long nblk = (DCR_AL16(y) && DCR_AL8(bin)) ? (n >> 9) : 0;
        for (long long w = warp; w < nblk; w += nwarps) {
            const float4* y4 = reinterpret_cast<const float4*>(y) + (w << 7);
            const uint2* q4 = reinterpret_cast<const uint2*>(bin) + (w << 7);
            float4 t[4];
            uint2 qq[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { t[k] = __ldg(y4 + k * 32 + lane); qq[k] = __ldg(q4 + k * 32 + lane); }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                elem(t[k].x, qq[k].x & 0xffffu);
                elem(t[k].y, qq[k].x >> 16);
                elem(t[k].z, qq[k].y & 0xffffu);
                elem(t[k].w, qq[k].y >> 16);
            }
            drain();
        }
        for (long long base = (nblk << 9) + warp * 32; base < n; base += nwarps * 32) {
            const long long i = base + lane;
            if (i < n) elem(y[i], (unsigned)bin[i]);
            drain();
        }
    }
    __syncthreads();
    const unsigned nst = s_n < FB_STAGE ? s_n : FB_STAGE;
    if (threadIdx.x == 0) s_gbase = nst ? atomicAdd(&nd->fast_ncand, nst) : 0u;
    __syncthreads();
    const unsigned gb = s_gbase;
    for (unsigned k = threadIdx.x; k < nst; k += 256)
        if (gb + k < cap) cand[gb + k] = s_buf[k];
}

// one warp per bin: fine cells of the two median elements (scan of 1024 fine cells)
// band_flag (row-band mode): all-reduced "some rank ran out of candidate space" word instead of this rank's own count
__global__ void __launch_bounds__(256) fb_findB_kernel(NuthDev* nd, unsigned* __restrict__ ghist, unsigned cap,
                                                       const unsigned* __restrict__ band_flag) {
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= NB || nd->fast_fallback) return;
    if (b == 0 && lane == 0 && (band_flag ? (*band_flag != 0u) : (nd->fast_ncand > cap))) nd->fast_fallback = 1;
    if (nd->bin_count[b] == 0) return;
    const unsigned c0 = nd->fa_cell[0][b], c1 = nd->fa_cell[1][b];
    for (int r = 0; r < 2; ++r) {
        const int plane = (r == 1 && c1 != c0) ? 1 : 0;
        const unsigned* h = ghist + (plane * NB + b) * FB_FINE;
        const unsigned k = nd->fa_rank[r][b];
        // lane l owns fine cells [32 l, 32 l + 32)
        unsigned v[32], sum = 0;
#pragma unroll
        for (int q = 0; q < 32; ++q) { v[q] = h[lane * 32 + q]; sum += v[q]; }
        const unsigned inc = dcr_warp_incl_scan(sum);
        unsigned ex = inc - sum;
        if (k >= ex && k < inc) {
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                if (k >= ex && k < ex + v[q]) {
                    nd->fb_cell[r][b] = (unsigned short)(((r == 1 ? c1 : c0) << 10) | (lane * 32 + q));
                    nd->fa_rank[r][b] = k - ex;   // now the rank inside the fine cell
                    if (v[q] > FB_SLOT) nd->fast_fallback = 1;  // heavy ties: too many samples in one fine cell
                }
                ex += v[q];
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) fb_passC_kernel(NuthDev* nd, const unsigned long long* __restrict__ cand,
                                                       unsigned* __restrict__ slots, unsigned* __restrict__ slotn) {
    __shared__ unsigned short s_f0[NB], s_f1[NB];
    if (nd->fast_fallback) return;
    for (int k = threadIdx.x; k < NB; k += 256) { s_f0[k] = nd->fb_cell[0][k]; s_f1[k] = nd->fb_cell[1][k]; }
    __syncthreads();
    const unsigned ncand = nd->fast_ncand;
    const unsigned step = gridDim.x * 256;
    for (unsigned i0 = blockIdx.x * 256 + threadIdx.x; i0 < ncand; i0 += 4 * step) {
        unsigned long long rec[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) rec[u] = (i0 + u * step < ncand) ? __ldg(cand + i0 + u * step) : ~0ull;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + u * step >= ncand) break;
            const unsigned f = (unsigned)(rec[u] & 0xFFFFu), b = (unsigned)((rec[u] >> 16) & 0xFFFFu), key = (unsigned)(rec[u] >> 32);
            const unsigned f0 = s_f0[b], f1 = s_f1[b];
            if (f == f0) {
                const unsigned pos = atomicAdd(&slotn[b * 2], 1u);
                if (pos < FB_SLOT) slots[(b * 2) * FB_SLOT + pos] = key;
            } else if (f == f1) {
                const unsigned pos = atomicAdd(&slotn[b * 2 + 1], 1u);
                if (pos < FB_SLOT) slots[(b * 2 + 1) * FB_SLOT + pos] = key;
            }
        }
    }
}

// one warp per bin: exact selection among the <= 128 keys of the final fine cell(s)
__global__ void __launch_bounds__(256) fb_passD_kernel(NuthDev* nd, const unsigned* __restrict__ ghist) {
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= NB || nd->fast_fallback) return;
    if (nd->bin_count[b] == 0) return;
    const unsigned* slots = ghist + FB_SLOT_OFF;
    const unsigned* slotn = ghist + FB_SLOTN_OFF;
    const unsigned f0 = nd->fb_cell[0][b], f1 = nd->fb_cell[1][b];
    float val[2];
    bool bad = false;
    for (int r = 0; r < 2; ++r) {
        const int sl = (r == 1 && f1 != f0) ? 1 : 0;
        const unsigned m = slotn[b * 2 + sl];
        const unsigned k = nd->fa_rank[r][b];
        if (m > FB_SLOT || k >= m) { bad = true; val[r] = 0.f; continue; }
        const unsigned* q = slots + (b * 2 + sl) * FB_SLOT;
        unsigned found = 0;
        bool have = false;
        for (unsigned i = lane; i < m; i += 32) {
            const unsigned e = q[i];
            unsigned rank = 0;
            for (unsigned j = 0; j < m; ++j) { const unsigned o = q[j]; rank += (o < e) || (o == e && j < i); }
            if (rank == k) { found = e; have = true; }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, have);
        if (!bal) { bad = true; val[r] = 0.f; continue; }
        found = __shfl_sync(0xffffffffu, found, __ffs(bal) - 1);
        val[r] = dcr_key2f(found);
    }
    if (lane == 0) {
        if (bad) nd->fast_fallback = 1;
        else nd->bin_med[b] = (val[0] + val[1]) / 2.0f;
    }
}

// ---- row-band mode glue of the fast bins -----------------------------------------------------------------------------
__global__ void fb_band_flag_kernel(const NuthDev* nd, unsigned cap, unsigned* flag) {
    if (threadIdx.x == 0) *flag = (nd->fast_ncand > cap) ? 1u : 0u;
}
// this rank's row of the slot-count table (the other rows stay zero: the sum over ranks assembles the table)
__global__ void fb_band_counts_kernel(const unsigned* __restrict__ lslotn, unsigned* __restrict__ rcnt, int rank) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < NB * 2) rcnt[rank * NB * 2 + k] = lslotn[k];
}
// this rank's keys of every final fine cell go behind those of the lower ranks in the (zeroed) common slot array; the total
// per cell is known to every rank from the table
__global__ void __launch_bounds__(256) fb_band_scatter_kernel(const unsigned* __restrict__ lslots, const unsigned* __restrict__ rcnt,
                                                              int rank, int nranks, unsigned* __restrict__ slots,
                                                              unsigned* __restrict__ slotn, NuthDev* nd) {
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5);   // one warp per (bin, cell)
    const int lane = threadIdx.x & 31;
    if (k >= NB * 2 || nd->fast_fallback) return;
    unsigned off = 0, tot = 0;
    for (int r = 0; r < nranks; ++r) {
        const unsigned c = rcnt[r * NB * 2 + k];
        if (r < rank) off += c;
        tot += c;
    }
    const unsigned mine = rcnt[rank * NB * 2 + k];
    if (lane == 0) slotn[k] = tot;
    if (tot > FB_SLOT) return;                        // (pass D declines: heavy ties)
    for (unsigned i = lane; i < mine; i += 32) slots[k * FB_SLOT + off + i] = lslots[k * FB_SLOT + i];
}

struct NuthWs {
    float* y;
    unsigned short* bin;
    NuthPartial* part;
    unsigned* ghist;
    NuthDev* nd;
    unsigned long long* cand;   // fast bin path: (key << 32 | bin << 16 | fine cell) of the elements in the median coarse cells
    unsigned cand_cap;
};
// partial moments: one per CTA of the prepare kernel, or one per 2048-element chunk of the fused mask update
static long long nuth_nparts(long long n, int grid) {
    const long long nch = (n + 2047) / 2048;
    return (nch > grid ? nch : grid) + NUTH_RED_CTAS;   // + the first-level results of nuth_moments_reduce_kernel
}
static unsigned nuth_cand_cap(long long n) {
    long long c = n / 8;
    if (c < 65536) c = 65536;
    if (c > 0x7fffffffLL) c = 0x7fffffffLL;
    return (unsigned)c;
}
static size_t nuth_ws_bytes(long long n, int grid) {
    return dcr_al256(n * sizeof(float)) + dcr_al256(n * sizeof(unsigned short)) + dcr_al256((size_t)nuth_nparts(n, grid) * sizeof(NuthPartial)) +
           dcr_al256((size_t)NUTH_GHIST_WORDS * sizeof(unsigned)) + dcr_al256(sizeof(NuthDev)) +
           dcr_al256((size_t)nuth_cand_cap(n) * sizeof(unsigned long long));
}
static int nuth_ws_carve(DcrBump& ws, long long n, int grid, NuthWs* w) {
    w->y = ws.take<float>(n);
    w->bin = ws.take<unsigned short>(n);
    w->part = ws.take<NuthPartial>(nuth_nparts(n, grid));
    w->ghist = ws.take<unsigned>((size_t)NUTH_GHIST_WORDS);
    w->nd = ws.take<NuthDev>(1);
    w->cand_cap = nuth_cand_cap(n);
    w->cand = ws.take<unsigned long long>(w->cand_cap);
    if (!w->y || !w->bin || !w->part || !w->ghist || !w->nd || !w->cand) { dcr_set_error("workspace too small"); return -1; }
    return 0;
}

// radix bins (3 passes of global-atomic histograms): the exact fallback of the fast path
static int nuth_bins_radix_enqueue(long long n, const NuthWs& w, int grid, cudaStream_t s) {
    DCR_CUDA_OK(cudaMemsetAsync(w.ghist, 0, (size_t)2 * NB * 2048 * sizeof(unsigned), s));
    DCR_CUDA_OK(cudaMemsetAsync(w.nd->bin_count, 0, sizeof(unsigned long long) * NB, s));
    bin_hist_kernel<0><<<grid, 256, 0, s>>>(w.y, w.bin, n, w.nd, w.ghist);
    bin_find_kernel<0><<<NB, 256, 0, s>>>(w.nd, w.ghist);
    bin_hist_kernel<1><<<grid, 256, 0, s>>>(w.y, w.bin, n, w.nd, w.ghist);
    bin_find_kernel<1><<<NB, 256, 0, s>>>(w.nd, w.ghist);
    bin_hist_kernel<2><<<grid, 256, 0, s>>>(w.y, w.bin, n, w.nd, w.ghist);
    bin_find_kernel<2><<<NB, 256, 0, s>>>(w.nd, w.ghist);
    nuth_count_final_kernel<<<1, 32, 0, s>>>(w.nd);
    DCR_LAUNCH_OK("radix bin kernels");
    return 0;
}

static int nuth_bins_fast_enqueue(long long n, const NuthWs& w, cudaStream_t s) {
    const size_t smemA = (size_t)NB * FB_COARSE * sizeof(unsigned);
    DcrCtx* ctx = dcr_ctx();
    if (!ctx) return -1;
    if (!(ctx->attrs & DCR_ATTR_FBA)) {   // per device
        DCR_CUDA_OK(cudaFuncSetAttribute(fb_passA_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA));
        ctx->attrs |= DCR_ATTR_FBA;
    }
    const int nsm = dcr_num_sms();
    DCR_CUDA_OK(cudaMemsetAsync(w.ghist, 0, (size_t)FB_WORDS * sizeof(unsigned), s));
    fb_passA_kernel<<<nsm * 2, 512, smemA, s>>>(w.y, w.bin, n, w.nd, w.ghist);
    fb_findA_kernel<<<(NB + 127) / 128, 128, 0, s>>>(w.nd, w.ghist);
    fb_passB_kernel<<<nsm * 16, 256, 0, s>>>(w.y, w.bin, n, w.nd, w.ghist, w.cand, w.cand_cap);  // small CTAs: the 4 % candidates of a CTA must fit its staging buffer
    fb_findB_kernel<<<(NB + 7) / 8, 256, 0, s>>>(w.nd, w.ghist, w.cand_cap, nullptr);
    fb_passC_kernel<<<nsm * 8, 256, 0, s>>>(w.nd, w.cand, w.ghist + FB_SLOT_OFF, w.ghist + FB_SLOTN_OFF);
    fb_passD_kernel<<<(NB + 7) / 8, 256, 0, s>>>(w.nd, w.ghist);
    DCR_LAUNCH_OK("fast bin kernels");   // (n_final = sum of the bin counts: nuth_dev_to_stats, on the host)
    return 0;
}

static int nuth_enqueue(const float* dh, const float* slope, const float* aspect, const unsigned char* m, unsigned reject,
                        long long n, int remove_outliers, const NuthWs& w, int grid, cudaStream_t s,
                        cudaEvent_t after_prepare = nullptr, int fast_bins = 1) {
    nuth_prepare_kernel<<<grid, 256, 0, s>>>(dh, slope, aspect, m, reject, n, w.y, w.bin, w.part);
    nuth_moments_final_kernel<<<1, 256, 0, s>>>(w.part, grid, remove_outliers, w.nd);
    if (after_prepare) DCR_CUDA_OK(cudaEventRecord(after_prepare, s));
    DCR_LAUNCH_OK("nuth prepare kernels");
    if (fast_bins) return nuth_bins_fast_enqueue(n, w, s);
    return nuth_bins_radix_enqueue(n, w, grid, s);
}
// the same when y / bin / the partial moments already exist (front kernel + fused mask update of the iteration driver)
static int nuth_enqueue_prepared(long long n, long long nparts, int remove_outliers, const NuthWs& w, int grid, cudaStream_t s,
                                 cudaEvent_t after_prepare, int fast_bins) {
    if (nparts > 4096) {
        nuth_moments_reduce_kernel<<<NUTH_RED_CTAS, 256, 0, s>>>(w.part, (int)nparts, w.part + nparts);
        nuth_moments_final_kernel<<<1, 256, 0, s>>>(w.part + nparts, NUTH_RED_CTAS, remove_outliers, w.nd);
    } else {
        nuth_moments_final_kernel<<<1, 256, 0, s>>>(w.part, (int)nparts, remove_outliers, w.nd);
    }
    if (after_prepare) DCR_CUDA_OK(cudaEventRecord(after_prepare, s));
    DCR_LAUNCH_OK("nuth_moments_final_kernel");
    if (fast_bins) return nuth_bins_fast_enqueue(n, w, s);
    return nuth_bins_radix_enqueue(n, w, grid, s);
}

static void nuth_dev_to_stats(const NuthDev& d, dcr_nuth_stats* o) {
    o->n_common = (int64_t)d.n_common;
    // n_final = sum of the bin counts (+ in-range samples outside [0,360]: none for gdaldem aspect)
    unsigned long long nf = 0;
    for (int k = 0; k < NB; ++k) nf += d.bin_count[k];
    o->n_final = (int64_t)nf;
    o->mean_y = d.mean_y; o->std_y = d.std_y; o->rmin = d.rmin; o->rmax = d.rmax;
    for (int k = 0; k < NB; ++k) { o->bin_count[k] = (int64_t)d.bin_count[k]; o->bin_med[k] = d.bin_med[k]; }
}

extern "C" size_t dcr_nuth_workspace_bytes(int64_t n) { return nuth_ws_bytes(n, dcr_num_sms() * 8) + 4096; }

extern "C" int dcr_nuth_bin_stats(const float* dh, const float* slope, const float* aspect, const uint8_t* maskbits,
                                  uint32_t reject_dh, uint32_t reject_slope, uint32_t reject_aspect, int64_t n,
                                  int remove_outliers, dcr_nuth_stats* out, void* workspace, size_t workspace_bytes,
                                  void* stream) {
    DCR_ARG(dh && slope && aspect && out && workspace, "null pointer");
    DCR_ARG(n > 0, "n > 0");
    const int grid = dcr_num_sms() * 8;
    DcrBump ws(workspace, workspace_bytes);
    NuthWs w;
    if (nuth_ws_carve(ws, n, grid, &w)) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = nuth_enqueue(dh, slope, aspect, maskbits, reject_dh | reject_slope | reject_aspect, n, remove_outliers, w, grid, s);
    if (rc) return rc;
    static thread_local NuthDev host;
    DCR_CUDA_OK(cudaMemcpyAsync(&host, w.nd, sizeof(NuthDev), cudaMemcpyDeviceToHost, s));
    DCR_CUDA_OK(cudaStreamSynchronize(s));
    if (host.fast_fallback) {  // heavy ties / degenerate range: exact 3-pass radix bins
        if ((rc = nuth_bins_radix_enqueue(n, w, grid, s))) return rc;
        DCR_CUDA_OK(cudaMemcpyAsync(&host, w.nd, sizeof(NuthDev), cudaMemcpyDeviceToHost, s));
        DCR_CUDA_OK(cudaStreamSynchronize(s));
    }
    nuth_dev_to_stats(host, out);
    return 0;
}

// ------------------------------------------------------------------------------------------
// host fit: least squares of a*cos(deg2rad(b - x)) + c == A cos x + B sin x + c  (Householder QR)
// ------------------------------------------------------------------------------------------
// sum of a[r] * b[r], r in [r0, n): four independent accumulators (a single one is a 360-long chain of dependent adds, and
// fifteen of those chains were most of the fit's time)
static inline double fit_dot(const double* a, const double* b, int r0, int n) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int r = r0;
    for (; r + 4 <= n; r += 4) {
        s0 += a[r] * b[r]; s1 += a[r + 1] * b[r + 1]; s2 += a[r + 2] * b[r + 2]; s3 += a[r + 3] * b[r + 3];
    }
    for (; r < n; ++r) s0 += a[r] * b[r];
    return (s0 + s1) + (s2 + s3);
}

extern "C" int dcr_fit_nuth(const double* xc, const double* ym, int n, double fit[3]) {
    DCR_ARG(xc && ym && fit, "null pointer");
    DCR_ARG(n >= 3 && n <= 4096, "3 <= n <= 4096");
    // column-major working copy [cos x | sin x | 1 | y] + the Householder vector; plain pointers into the thread's buffers
    // (the loop over 360 rows sits between two iterations of the dem_align loop: indexing a thread_local array through its
    // TLS address in every access made this 37 us, a tenth of it now)
    static thread_local double buf[5 * 4096];
    static thread_local double tab_c[NB], tab_s[NB];
    static thread_local bool tab_ok = false;
    double* const tc = tab_c;
    double* const ts = tab_s;
    if (!tab_ok) {
        // cos / sin of the 360 bin centres k + 0.5 (the only abscissae the iteration ever passes): computed once
        for (int k = 0; k < NB; ++k) {
            const double xr = ((double)k + 0.5) * (M_PI / 180.0);
            tc[k] = cos(xr);
            ts[k] = sin(xr);
        }
        tab_ok = true;
    }
    double* const base = buf;
    double* col[4] = {base, base + 4096, base + 2 * 4096, base + 3 * 4096};
    double* const v = base + 4 * 4096;
    for (int r = 0; r < n; ++r) {
        const int k = (int)xc[r];
        if (k >= 0 && k < NB && xc[r] == (double)k + 0.5) {
            col[0][r] = tc[k];
            col[1][r] = ts[k];
        } else {
            const double xr = xc[r] * (M_PI / 180.0);
            col[0][r] = cos(xr);
            col[1][r] = sin(xr);
        }
        col[2][r] = 1.0;
        col[3][r] = ym[r];
    }
    for (int k = 0; k < 3; ++k) {
        double* const ck = col[k];
        const double norm = sqrt(fit_dot(ck, ck, k, n));
        if (norm == 0) { dcr_set_error("rank-deficient fit"); return -4; }
        const double alpha = ck[k] > 0 ? -norm : norm;
        // v = x - alpha e1
        for (int r = k; r < n; ++r) v[r] = ck[r];
        v[k] -= alpha;
        const double vv = fit_dot(v, v, k, n);
        if (vv == 0) continue;
        for (int c = k; c < 4; ++c) {
            double* const cc = col[c];
            const double f = 2.0 * fit_dot(v, cc, k, n) / vv;
            for (int r = k; r < n; ++r) cc[r] -= f * v[r];
        }
    }
    double sol[3];
    for (int k = 2; k >= 0; --k) {
        double s = col[3][k];
        for (int c = k + 1; c < 3; ++c) s -= col[c][k] * sol[c];
        if (col[k][k] == 0) { dcr_set_error("rank-deficient fit"); return -4; }
        sol[k] = s / col[k][k];
    }
    const double A = sol[0], B = sol[1];
    fit[0] = hypot(A, B);
    double b = atan2(B, A) * (180.0 / M_PI);
    fit[1] = b;
    fit[2] = sol[2];
    return 0;
}

// ------------------------------------------------------------------------------------------
// K8 apply_z_shift
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) zshift_kernel(float* __restrict__ band, int h, int w, long long stride, double ndv,
                                                     double tol, float ndv32, float dz, const float* __restrict__ dzg) {
    const long long n = (long long)h * w;
    const long long step = (long long)gridDim.x * 256;
    if (stride == w && DCR_AL16(band) && (dzg == nullptr || DCR_AL16(dzg))) {
        // contiguous band: float4 read-modify-write
        const long long n4 = n >> 2;
        float4* b4 = reinterpret_cast<float4*>(band);
        for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < n4; k += step) {
            float4 v = b4[k];
            float4 z = make_float4(dz, dz, dz, dz);
            if (dzg) z = __ldg(reinterpret_cast<const float4*>(dzg) + k);
            v.x = (fabs((double)v.x - ndv) <= tol) ? ndv32 : v.x + z.x;
            v.y = (fabs((double)v.y - ndv) <= tol) ? ndv32 : v.y + z.y;
            v.z = (fabs((double)v.z - ndv) <= tol) ? ndv32 : v.z + z.z;
            v.w = (fabs((double)v.w - ndv) <= tol) ? ndv32 : v.w + z.w;
            b4[k] = v;
        }
        for (long long k = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; k < n; k += step) {
            const float v = band[k];
            band[k] = (fabs((double)v - ndv) <= tol) ? ndv32 : v + (dzg ? dzg[k] : dz);
        }
        return;
    }
    for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < n; k += step) {
        const int i = (int)(k / w), j = (int)(k - (long long)i * w);
        float* p = band + (long long)i * stride + j;
        const float v = *p;
        // b_getma: masked_values(|x - ndv| <= 1e-8 + 1e-5|ndv|); a + dz; filled() -> masked pixels become ndv
        if (fabs((double)v - ndv) <= tol) *p = ndv32;
        else *p = v + (dzg ? dzg[k] : dz);
    }
}

extern "C" int dcr_apply_z_shift(float* band, int h, int w, int64_t stride, double ndv, float dz, const float* dz_grid,
                                 void* stream) {
    DCR_ARG(band, "null pointer");
    DCR_ARG(h > 0 && w > 0, "empty raster");
    const double tol = 1e-8 + 1e-5 * fabs(ndv);
    zshift_kernel<<<dcr_num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(band, h, w, stride, ndv, tol, (float)ndv, dz, dz_grid);
    DCR_LAUNCH_OK("zshift_kernel");
    return 0;
}

// several deferred scalar shifts in ONE read-modify-write pass (same ordered float32 adds as separate calls)
struct ZSeq { float dz[16]; int n; };
__global__ void __launch_bounds__(256) zshift_seq_kernel(float* __restrict__ band, int h, int w, long long stride, float zlo,
                                                         float zhi, float ndv32, ZSeq z) {
    // masked_values(|x - ndv| <= tol) as the equivalent float32 interval [zlo, zhi] (see gm_bounds in dcr_front.cu)
    const long long n = (long long)h * w;
    const long long step = (long long)gridDim.x * 256;
    auto shift = [&](float v) {
        for (int q = 0; q < z.n; ++q) v = (v >= zlo && v <= zhi) ? ndv32 : v + z.dz[q];
        return v;
    };
    if (stride == w && DCR_AL16(band)) {   // contiguous band: float4 read-modify-write
        const long long n4 = n >> 2;
        float4* b4 = reinterpret_cast<float4*>(band);
        for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < n4; k += step) {
            float4 v = b4[k];   // (4 loads in flight per thread measured slower: 185 vs 156 us)
            v.x = shift(v.x); v.y = shift(v.y); v.z = shift(v.z); v.w = shift(v.w);
            b4[k] = v;
        }
        for (long long k = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; k < n; k += step) band[k] = shift(band[k]);
        return;
    }
    for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < n; k += step) {
        const int i = (int)(k / w), j = (int)(k - (long long)i * w);
        float* p = band + (long long)i * stride + j;
        *p = shift(*p);
    }
}

extern "C" int dcr_apply_z_shift_seq(float* band, int h, int w, int64_t stride, double ndv, const float* dz, int n,
                                     void* stream) {
    DCR_ARG(band && dz, "null pointer");
    DCR_ARG(h > 0 && w > 0, "empty raster");
    DCR_ARG(n >= 1 && n <= 16, "1 <= n <= 16");
    ZSeq z;
    z.n = n;
    for (int k = 0; k < 16; ++k) z.dz[k] = k < n ? dz[k] : 0.f;
    // float32 interval equivalent to |x - ndv| <= 1e-8 + 1e-5|ndv| evaluated in float64
    const double tol = 1e-8 + 1e-5 * fabs(ndv);
    const double dl = ndv - tol, dh = ndv + tol;
    float zlo = (float)dl, zhi = (float)dh;
    if ((double)zlo < dl) zlo = nextafterf(zlo, INFINITY);
    if ((double)zhi > dh) zhi = nextafterf(zhi, -INFINITY);
    zshift_seq_kernel<<<dcr_num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(band, h, w, stride, zlo, zhi, (float)ndv, z);
    DCR_LAUNCH_OK("zshift_seq_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------------
// one chained iteration
// ------------------------------------------------------------------------------------------
struct IterDev {
    unsigned long long counters[4];  // base, after max_dz, slope-valid, (pad)
    unsigned long long count_mad, count_final;
    long long rank_counts[8];        // row-band mode: unmasked-pixel count of every band (own slot written, all-reduced)
    long long band_base;             // row-band mode: unmasked pixels in the bands above this one
    double mom3[4];                  // row-band mode: {n, sum y, sum y^2, pad} of the common-mask samples (all-reduced)
    SelState sel[5];                 // 0 med(dh|A)  1 med|dh-med|  2 med(dh|F)  3 med(subsample)  4 med(slope)
    BselState bsel[5];               // the same five selections through the one-pass sample-bracket select
    float selres[5];                 // results of whichever engine ran (device-side consumers read these)
    int selpad;
    float nmad, mad_lo, mad_hi, pad0;
    long long total_final, stride, dense_n;
    NuthDev nuth;
};

__global__ void mad_thresholds_kernel(IterDev* d, int remove_outliers) {
    if (threadIdx.x) return;
    if (!remove_outliers || isnan(d->selres[0])) {
        d->nmad = 0.f; d->mad_lo = -INFINITY; d->mad_hi = INFINITY;
        return;
    }
    // filtlib.mad_fltr: mad = 1.4826 * median(|x - med|) ; keep med - 3*mad <= x <= med + 3*mad  (float32)
    const float med = d->selres[0];
    const float nmad = d->selres[1] * 1.4826f;
    const float t = 3.0f * nmad;
    d->nmad = nmad;
    d->mad_lo = med - t;
    d->mad_hi = med + t;
}

#define MU_CHUNK 2048
// MAD filter (filtlib.mad_fltr <- dem_align.py:46) applied to the mask bytes + per-chunk counts of the final mask.
// NUTH: this pass decides the last mask, so it also produces the Nuth-Kaab samples of compute_offset_nuth right away
// (no separate "prepare" pass over dh/slope/aspect/mask): y = dh / tan(slope) for the pixels of the final common mask
// (coreglib.py:229-233), count / sum / sum of squares of y (coreglib.py:238; one partial per chunk, reduced in a fixed
// order), and the bin code -- emitted by the front kernel from the aspect -- invalidated where the MAD filter removes
// a pixel.  The arithmetic of a bandwidth-bound pass is free; the front kernel is issue-bound and could not host it.
#define MU_KCH 4   // consecutive chunks per CTA: one partial-moment reduction and one pair of global atomics per CTA
template <bool NUTH>
__global__ void __launch_bounds__(256) mask_update_kernel(const float* __restrict__ dh, unsigned char* m,
                                                          long long n, IterDev* d, unsigned* __restrict__ chunk_cnt,
                                                          const float* __restrict__ slope, float* __restrict__ y,
                                                          unsigned short* __restrict__ bin, NuthPartial* __restrict__ part,
                                                          int thr_mode /* 0: thresholds in d; 1: compute; 2: no filter */,
                                                          int tan_slope /* `slope` holds tan(slope), not degrees */) {
    __shared__ unsigned s_a[8], s_b[8];
    __shared__ NuthPartial sp[8];
    float lo, hi;
    if (thr_mode == 0) {
        lo = d->mad_lo; hi = d->mad_hi;   // set by mad_thresholds_kernel (row-band mode: after the all-reduced selects)
    } else {
        // filtlib.mad_fltr: mad = 1.4826 * median(|x - med|) ; keep med - 3*mad <= x <= med + 3*mad  (float32)
        float nmad = 0.f;
        lo = -INFINITY; hi = INFINITY;
        const float med = d->selres[0];
        if (thr_mode == 1 && !isnan(med)) {
            nmad = d->selres[1] * 1.4826f;
            const float t = 3.0f * nmad;
            lo = med - t;
            hi = med + t;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) { d->nmad = nmad; d->mad_lo = lo; d->mad_hi = hi; }
    }
    const long long nch = (n + MU_CHUNK - 1) / MU_CHUNK;
    unsigned long long cnt = 0;
    double s = 0.0, ss = 0.0;
    unsigned n_mad = 0, n_fin = 0;
    for (int kc = 0; kc < MU_KCH; ++kc) {
        const long long chunk = (long long)blockIdx.x * MU_KCH + kc;
        if (chunk >= nch) break;     // CTA-uniform
        const long long base = chunk * MU_CHUNK + threadIdx.x * 8;
        float v[8];
        unsigned mb[8];
        dcr_load8(dh, m, base, n, v, mb);
        float sl[8];
        const bool full = base + 8 <= n;
        if (NUTH) {
            if (full && DCR_AL16(slope)) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(slope + base));
                const float4 b = __ldg(reinterpret_cast<const float4*>(slope + base) + 1);
                sl[0] = a.x; sl[1] = a.y; sl[2] = a.z; sl[3] = a.w; sl[4] = b.x; sl[5] = b.y; sl[6] = b.z; sl[7] = b.w;
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) sl[k] = (base + k < n) ? slope[base + k] : 0.f;
            }
        }
        unsigned changed = 0;
        float yv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            yv[k] = 0.f;
            if (mb[k] > 0xffu) continue;  // past the end
            unsigned bits = mb[k];
            if (!(bits & (DCR_M_BASE | DCR_M_MAXDZ))) {
                if (v[k] < lo || v[k] > hi) {
                    bits |= DCR_M_MAD; changed |= 1u << k;
                    // the pixel leaves the common mask: its bin code was valid iff no other mask bit is set
                    if (NUTH && mb[k] == 0u) bin[base + k] = (unsigned short)BIN_INVALID;
                }
                else ++n_mad;
            }
            mb[k] = bits;
            if (!(bits & DCR_M_DIFF)) ++n_fin;
            if (NUTH && bits == 0u) {   // common mask of compute_offset_nuth: dh, slope and aspect all unmasked
                const float t = tan_slope ? __fdiv_rn(v[k], sl[k]) : nuth_y(v[k], sl[k]);
                yv[k] = t;
                ++cnt;
                s += (double)t;
                ss += (double)t * (double)t;
            }
        }
        if (NUTH) {
            if (full && DCR_AL16(y)) {
                reinterpret_cast<float4*>(y + base)[0] = make_float4(yv[0], yv[1], yv[2], yv[3]);
                reinterpret_cast<float4*>(y + base)[1] = make_float4(yv[4], yv[5], yv[6], yv[7]);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (base + k < n) y[base + k] = yv[k];
            }
        }
        if (changed) {
            if (full && DCR_AL8(m)) {
                uint2 q;
                q.x = mb[0] | (mb[1] << 8) | (mb[2] << 16) | (mb[3] << 24);
                q.y = mb[4] | (mb[5] << 8) | (mb[6] << 16) | (mb[7] << 24);
                *reinterpret_cast<uint2*>(m + base) = q;
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (changed & (1u << k)) m[base + k] = (unsigned char)mb[k];
            }
        }
    }
    // one block reduction per CTA: MU_KCH chunks = one 8192-element group of the compress / scan kernels
    n_mad = __reduce_add_sync(0xffffffffu, n_mad);
    n_fin = __reduce_add_sync(0xffffffffu, n_fin);
    if (NUTH) {
        cnt = dcr_warp_sum_u64(cnt);
        s = dcr_warp_sum(s);
        ss = dcr_warp_sum(ss);
    }
    if ((threadIdx.x & 31) == 0) {
        s_a[threadIdx.x >> 5] = n_mad; s_b[threadIdx.x >> 5] = n_fin;
        if (NUTH) { NuthPartial q; q.n = cnt; q.s = s; q.ss = ss; sp[threadIdx.x >> 5] = q; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned a = 0, b = 0;
        for (int k = 0; k < 8; ++k) { a += s_a[k]; b += s_b[k]; }
        chunk_cnt[blockIdx.x] = b;
        if (a) atomicAdd(&d->count_mad, (unsigned long long)a);
        if (b) atomicAdd(&d->count_final, (unsigned long long)b);
        if (NUTH) {
            NuthPartial q = sp[0];
            for (int k = 1; k < 8; ++k) { q.n += sp[k].n; q.s += sp[k].s; q.ss += sp[k].ss; }
            part[blockIdx.x] = q;
        }
    }
}

__global__ void stride_kernel(IterDev* d) {
    if (threadIdx.x) return;
    // malib.get_stats: thresh = 4e6; stride = int(np.around(count / thresh)) (round half to even)
    const long long total = d->total_final;
    long long stride = 1;
    if ((double)total >= 4e6) stride = (long long)rint((double)total / 4e6);
    if (stride < 1) stride = 1;
    d->stride = stride;
    d->dense_n = stride > 1 ? (total + stride - 1) / stride : 0;
}

static size_t iter_ws_bytes(long long n, int grid) {
    const long long nch = (n + MU_CHUNK - 1) / MU_CHUNK;
    return nuth_ws_bytes(n, grid) + dcr_al256(sizeof(IterDev)) + dcr_al256(DCR_SEL_HIST_WORDS * sizeof(unsigned)) +
           dcr_al256(nch * sizeof(unsigned)) + dcr_al256(nch * sizeof(long long)) +
           dcr_al256((size_t)(n < 6000016 ? n : 6000016) * sizeof(float)) + 3 * dcr_bsel_scratch_bytes(n) +
           dcr_al256((size_t)dcr_front_list_cap(n) * sizeof(unsigned)) + 4096;
}
extern "C" size_t dcr_iteration_workspace_bytes(int64_t n) { return iter_ws_bytes(n, dcr_num_sms() * 8); }


// apply_z_shift (coreglib.py:42-55) with dz taken from the device-resident result of THIS iteration
// (dz = -median, dem_align.py:115, 172), enqueued before the host even sees the result: the 0.1 ms
// read-modify-write pass overlaps the host's fit / bookkeeping instead of sitting between two iterations.
// Skipped when a one-pass selection asked for the radix fallback (the rerun applies it).
__global__ void __launch_bounds__(256) zshift_dev_kernel(float* __restrict__ band, int h, int w, long long stride, double ndv,
                                                         double tol, float ndv32, const IterDev* d, int use_bsel) {
    if (use_bsel) {
        bool fb = false;
        for (int k = 0; k < 5; ++k) fb = fb || (d->bsel[k].fallback != 0);
        if (fb) return;
    }
    const float med = d->stride > 1 ? d->selres[3] : d->selres[2];
    if (isnan(med)) return;
    const float dz = -med;
    const long long n = (long long)h * w;
    const long long step = (long long)gridDim.x * 256;
    if (stride == w && DCR_AL16(band)) {
        const long long n4 = n >> 2;
        float4* b4 = reinterpret_cast<float4*>(band);
        for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < n4; k += step) {
            float4 v = b4[k];
            v.x = (fabs((double)v.x - ndv) <= tol) ? ndv32 : v.x + dz;
            v.y = (fabs((double)v.y - ndv) <= tol) ? ndv32 : v.y + dz;
            v.z = (fabs((double)v.z - ndv) <= tol) ? ndv32 : v.z + dz;
            v.w = (fabs((double)v.w - ndv) <= tol) ? ndv32 : v.w + dz;
            b4[k] = v;
        }
        for (long long k = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; k < n; k += step) {
            const float v = band[k];
            band[k] = (fabs((double)v - ndv) <= tol) ? ndv32 : v + dz;
        }
        return;
    }
    for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < n; k += step) {
        const int i = (int)(k / w), j = (int)(k - (long long)i * w);
        float* p = band + (long long)i * stride + j;
        const float v = *p;
        *p = (fabs((double)v - ndv) <= tol) ? ndv32 : v + dz;
    }
}

__global__ void sel_publish_kernel(IterDev* d, int k, int use_bsel) {
    if (threadIdx.x == 0) d->selres[k] = use_bsel ? d->bsel[k].result : d->sel[k].result;
}

struct IterWs {
    IterDev* d;
    unsigned* hist;
    unsigned* chunk_cnt;
    long long* chunk_off;
    float* dense;
    long long dense_cap;
    void* bsel_scratch;
    void* bsel_scratch2;   // selections running on the side stream
    void* bsel_scratch3;   // med_dh on its own side stream
    unsigned* tile_list = nullptr;   // front end: tiles left to the general kernel by the fast interior-tile kernel
    long long tile_cap = 0;
    NuthWs nw;
};

// one exact median (q = 50) of view v into d->selres[k], through either engine
static int iter_select(const SelView& v, int k, const IterWs& w, long long n, int use_bsel, cudaStream_t s) {
    int rc;
    if (use_bsel) return dcr_bsel_enqueue(v, 50.0, &w.d->bsel[k], w.bsel_scratch, n, s, &w.d->selres[k]);
    rc = dcr_sel_enqueue(v, 50.0, &w.d->sel[k], w.hist, s);
    if (rc) return rc;
    sel_publish_kernel<<<1, 32, 0, s>>>(w.d, k, 0);
    return 0;
}

static int iteration_enqueue(const dcr_front_args* front, int remove_outliers, int use_bsel, const IterWs& w, int grid,
                             cudaStream_t s, cudaEvent_t* ev, int fast_bins, int tan_slope, int no_bins = 0) {
    const long long n = (long long)front->h * front->w;
    const long long nch = (n + MU_CHUNK - 1) / MU_CHUNK;
    IterDev* d = w.d;
    int rc;
    // stage boundaries: CUDA events for DCR_ITER_PROFILE and an NVTX range per stage (header-only nvtx3: free unless a tool
    // is attached) -- 0 front, 1 median + MAD, 2 mask update, 3 subsample chain, 4 med_dh / med_slope, 5-6 bin statistics
    static const char* const stage_name[7] = {"dcr:front (warp+dh+masks+Horn+bins)", "dcr:median+MAD selects",
                                              "dcr:mask update (+y, moments)", "dcr:subsample scan/compress/dz",
                                              "dcr:med_dh/med_slope selects + bin statistics", "dcr:bin statistics", "dcr:join+copy"};
    int nvtx_open = 0;
#define STAGE_MARK(k) do { if (ev) DCR_CUDA_OK(cudaEventRecord(ev[k], s)); \
                           if (nvtx_open) { nvtxRangePop(); nvtx_open = 0; } \
                           if ((k) < 6) { nvtxRangePushA(stage_name[(k) == 5 ? 5 : (k)]); nvtx_open = 1; } } while (0)
    STAGE_MARK(0);
    DCR_CUDA_OK(cudaMemsetAsync(d, 0, sizeof(IterDev), s));
    dcr_front_args fa = *front;
    if (!remove_outliers) fa.use_max_dz = 0;
    if ((rc = dcr_front_launch(&fa, d->counters, s, w.nw.bin, tan_slope, w.tile_list, w.tile_cap))) return rc;   // + bin codes

    // Side stream (one-pass-select path): work that does not sit on the dependency chain of the MAD filter runs
    // concurrently -- the slope median right after the front kernel, and the strided-subsample chain (scan, stride,
    // compress, dz median) right after the mask update -- so that their latency-bound small kernels hide behind the
    // bandwidth-bound passes of the main stream.  Joined before the result copy.
    DcrCtx* ctx = dcr_ctx();   // side streams / events belong to the current device of this thread
    if (!ctx) return -1;
    const bool side = use_bsel != 0;
    if (side && !ctx->s2) {
        DCR_CUDA_OK(cudaStreamCreateWithFlags(&ctx->s2, cudaStreamNonBlocking));
        DCR_CUDA_OK(cudaEventCreateWithFlags(&ctx->e_front, cudaEventDisableTiming));
        DCR_CUDA_OK(cudaEventCreateWithFlags(&ctx->e_mask, cudaEventDisableTiming));
        DCR_CUDA_OK(cudaEventCreateWithFlags(&ctx->e_join, cudaEventDisableTiming));
        DCR_CUDA_OK(cudaStreamCreateWithFlags(&ctx->s3, cudaStreamNonBlocking));
        DCR_CUDA_OK(cudaEventCreateWithFlags(&ctx->e_join3, cudaEventDisableTiming));
    }
    const cudaStream_t s2 = ctx->s2, s3 = ctx->s3;
    const cudaEvent_t e_front = ctx->e_front, e_mask = ctx->e_mask, e_join = ctx->e_join, e_join3 = ctx->e_join3;
    cudaStream_t sb = side ? s2 : s;      // stream of the side work
    IterWs wb = w;                        // side selections use their own scratch
    if (side) wb.bsel_scratch = w.bsel_scratch2;
    if (side) {
        DCR_CUDA_OK(cudaEventRecord(e_front, s));
        DCR_CUDA_OK(cudaStreamWaitEvent(s2, e_front, 0));
        SelView vs;
        memset(&vs, 0, sizeof(vs));
        vs.x = front->slope; vs.m = front->maskbits; vs.reject = DCR_M_SLOPE; vs.n = n;
        if ((rc = iter_select(vs, 4, wb, n, use_bsel, sb))) return rc;   // med_slope (coreglib.py:215)
    }
    STAGE_MARK(1);
    SelView v;
    memset(&v, 0, sizeof(v));
    v.x = front->dh; v.m = front->maskbits; v.n = n;
    if (remove_outliers) {
        v.reject = DCR_M_BASE | DCR_M_MAXDZ; v.transform = 0;
        static const bool msel_on = getenv("DCR_NO_MSEL") == nullptr;   // (DCR_NO_MSEL: the two dependent selections, for A/B)
        if (use_bsel && msel_on && n > 262144) {
            // median and MAD of dh (filtlib.mad_fltr) with ONE pass over the data: dcr_msel_enqueue
            if ((rc = dcr_msel_enqueue(v, &d->bsel[0], &d->bsel[1], w.bsel_scratch, w.bsel_scratch3, n, s, &d->selres[0],
                                       &d->selres[1])))
                return rc;
        } else {
            if ((rc = iter_select(v, 0, w, n, use_bsel, s))) return rc;
            v.transform = 1; v.center_ptr = &d->selres[0];
            if ((rc = iter_select(v, 1, w, n, use_bsel, s))) return rc;
        }
    }
    STAGE_MARK(2);
    mask_update_kernel<true><<<(unsigned)((nch + MU_KCH - 1) / MU_KCH), 256, 0, s>>>(
        front->dh, front->maskbits, n, d, w.chunk_cnt, front->slope, w.nw.y, w.nw.bin, w.nw.part, remove_outliers ? 1 : 2,
        tan_slope);
    if (side) {
        DCR_CUDA_OK(cudaEventRecord(e_mask, s));
        DCR_CUDA_OK(cudaStreamWaitEvent(s2, e_mask, 0));
        DCR_CUDA_OK(cudaStreamWaitEvent(s3, e_mask, 0));
    }
    // strided subsample of the final mask + dz on it (malib.get_stats <- dem_align.py:114-115)
    const long long nch4 = (nch + MU_KCH - 1) / MU_KCH;   // the mask update counts per CTA = 8192 elements
    if ((rc = dcr_chunk_scan_enqueue(w.chunk_cnt, nch4, w.chunk_off, &d->total_final, sb))) return rc;
    stride_kernel<<<1, 32, 0, sb>>>(d);
    if ((rc = dcr_compress_enqueue(front->dh, front->maskbits, DCR_M_DIFF, n, w.chunk_off, 1, &d->stride, w.dense,
                                   w.dense_cap, sb, nullptr, MU_KCH)))
        return rc;
    {
        SelView vd;
        memset(&vd, 0, sizeof(vd));
        vd.x = w.dense; vd.m = nullptr; vd.n = 0; vd.n_ptr = &d->dense_n;
        if ((rc = iter_select(vd, 3, wb, w.dense_cap, use_bsel, sb))) return rc;
    }
    STAGE_MARK(3);
    // med_dh over the final mask (coreglib.py:214): also off the critical path (the bin statistics below only need
    // the mask); its own side stream, concurrent with the subsample chain
    v.reject = DCR_M_DIFF; v.transform = 0; v.center_ptr = nullptr;
    {
        IterWs w3 = w;
        if (side) w3.bsel_scratch = w.bsel_scratch3;
        if ((rc = iter_select(v, 2, w3, n, use_bsel, side ? s3 : s))) return rc;
    }
    if (side) {
        DCR_CUDA_OK(cudaEventRecord(e_join, s2));
        DCR_CUDA_OK(cudaEventRecord(e_join3, s3));
    }
    if (!side) {
        SelView vs;
        memset(&vs, 0, sizeof(vs));
        vs.x = front->slope; vs.m = front->maskbits; vs.reject = DCR_M_SLOPE; vs.n = n;
        if ((rc = iter_select(vs, 4, w, n, use_bsel, s))) return rc;
    }
    STAGE_MARK(4);
    // (dem_align.py:147 calls coreglib.compute_offset_nuth WITHOUT forwarding remove_outliers: the 3-sigma trim of y is
    //  always on -- its default -- whatever dem_align's own flag says)
    if (no_bins) {
        // DCR_ITER_NO_BINS: the caller only wants the masks, the counts and the medians (-mode ncc / sad, dem_align.py:126-143)
        if (ev) DCR_CUDA_OK(cudaEventRecord(ev[5], s));
    } else if ((rc = nuth_enqueue_prepared(n, (nch + MU_KCH - 1) / MU_KCH, 1, w.nw, grid, s, ev ? ev[5] : nullptr,
                                           fast_bins)))
        return rc;
    if (side) {
        DCR_CUDA_OK(cudaStreamWaitEvent(s, e_join, 0));
        DCR_CUDA_OK(cudaStreamWaitEvent(s, e_join3, 0));
    }
    DCR_LAUNCH_OK("iteration kernels");
    STAGE_MARK(6);
#undef STAGE_MARK
    return 0;
}

// host part of an iteration: counts, guards, bin selection, fit, shift (coreglib.py:206-305, dem_align.py:150-172)
static int iter_finalize(const IterDev* hd, int engine_used, dcr_iter_result* out, int bin_engine = 0, int tan_slope = 0) {
    out->select_engine = engine_used;
    out->bin_engine = bin_engine;
    float selres[5];
    for (int k = 0; k < 5; ++k) selres[k] = engine_used ? hd->bsel[k].result : hd->sel[k].result;
    out->count_base = (int64_t)hd->counters[0];
    out->count_maxdz = (int64_t)hd->counters[1];
    out->count_slope = (int64_t)hd->counters[2];
    out->count_mad = (int64_t)hd->count_mad;
    out->count_final = (int64_t)hd->count_final;
    out->stats_stride = hd->stride;
    out->med1 = selres[0];
    out->nmad1 = hd->nmad;
    out->mad_lo = hd->mad_lo;
    out->mad_hi = hd->mad_hi;
    out->med_dh = selres[2];
    out->dz_med = hd->stride > 1 ? selres[3] : selres[2];
    // DCR_ITER_TAN_SLOPE: the selection ran on tan(slope), a monotone map of the slope
    out->med_slope = tan_slope ? (float)(atan((double)selres[4]) * (180.0 / M_PI)) : selres[4];
    nuth_dev_to_stats(hd->nuth, &out->nuth);
    out->dz = -(double)out->dz_med;
    if (out->count_base == 0) { out->status = 1; return 0; }
    if (out->count_final < 100) { out->status = 2; return 0; }
    if (out->count_slope < 100) { out->status = 3; return 0; }
    // c_seed (coreglib.py:216), float32
    out->c_seed = out->med_dh / tanf(out->med_slope * 0.017453292519943295f);
    // keep bins with >= 9 samples; need >= 90 bins and ptp(cos(centres)) >= 1 (coreglib.py:280-301)
    double xc[NB], ym[NB];
    int nb = 0;
    double cmin = INFINITY, cmax = -INFINITY;
    static thread_local double cos_center[NB];
    static thread_local bool cos_ok = false;
    if (!cos_ok) {
        for (int k = 0; k < NB; ++k) cos_center[k] = cos(((double)k + 0.5) * (M_PI / 180.0));
        cos_ok = true;
    }
    for (int k = 0; k < NB; ++k) {
        if (out->nuth.bin_count[k] >= 9) {
            xc[nb] = k + 0.5;
            ym[nb] = (double)out->nuth.bin_med[k];
            const double c = cos_center[k];
            cmin = fmin(cmin, c); cmax = fmax(cmax, c);
            ++nb;
        }
    }
    out->n_valid_bins = nb;
    if (nb >= 90 && (cmax - cmin) >= 1.0) {
        int rc = dcr_fit_nuth(xc, ym, nb, out->fit);
        if (rc) return rc;
        const double br = out->fit[1] * (M_PI / 180.0);
        out->dx = -(out->fit[0] * sin(br));
        out->dy = -(out->fit[0] * cos(br));
    } else {
        out->status = 4;
        out->fit[0] = out->fit[1] = out->fit[2] = NAN;
        out->dx = -0.0; out->dy = -0.0;
    }
    return 0;
}

extern "C" int dcr_nuth_iteration(const dcr_front_args* front, int flags, dcr_iter_result* out, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    DCR_ARG(front && out && workspace, "null pointer");
    const int remove_outliers = (flags & DCR_ITER_REMOVE_OUTLIERS) ? 1 : 0;
    const bool prof = (flags & DCR_ITER_PROFILE) != 0;
    const int tan_slope = (flags & DCR_ITER_TAN_SLOPE) ? 1 : 0;
    DcrCtx* ctx = dcr_ctx();
    if (!ctx) return -1;
    cudaEvent_t* ev = ctx->prof;
    if (prof && !ev[0])
        for (int k = 0; k < 8; ++k) DCR_CUDA_OK(cudaEventCreate(&ev[k]));
    const long long n = (long long)front->h * front->w;
    DCR_ARG(n > 0, "empty grid");
    const int grid = dcr_num_sms() * 8;
    DCR_ARG(workspace_bytes >= iter_ws_bytes(n, grid), "workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    DcrBump ws(workspace, workspace_bytes);
    IterWs w;
    w.d = ws.take<IterDev>(1);
    w.hist = ws.take<unsigned>(DCR_SEL_HIST_WORDS);
    const long long nch = (n + MU_CHUNK - 1) / MU_CHUNK;
    w.chunk_cnt = ws.take<unsigned>(nch);
    w.chunk_off = ws.take<long long>(nch);
    w.dense_cap = n < 6000016 ? n : 6000016;
    w.dense = ws.take<float>(w.dense_cap);
    w.bsel_scratch = ws.take<char>(dcr_bsel_scratch_bytes(n));
    w.bsel_scratch2 = ws.take<char>(dcr_bsel_scratch_bytes(n));
    w.bsel_scratch3 = ws.take<char>(dcr_bsel_scratch_bytes(n));
    w.tile_cap = dcr_front_list_cap(n);
    w.tile_list = ws.take<unsigned>(w.tile_cap);
    if (nuth_ws_carve(ws, n, grid, &w.nw)) return -1;
    w.nw.nd = &w.d->nuth;
    IterDev* d = w.d;

    IterDev* hd = (IterDev*)dcr_ctx_pinned(ctx, sizeof(IterDev));
    if (!hd) { dcr_set_error("pinned host allocation failed"); return -1; }
    int engine_used = (flags & DCR_ITER_RADIX_ONLY) ? 0 : 1;
    int bin_engine = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        int rc = iteration_enqueue(front, remove_outliers, engine_used, w, grid, s, prof ? ev : nullptr,
                                   (flags & DCR_ITER_RADIX_ONLY) ? 0 : 1, tan_slope, (flags & DCR_ITER_NO_BINS) ? 1 : 0);
        if (rc) return rc;
        DCR_CUDA_OK(cudaMemcpyAsync(hd, d, sizeof(IterDev), cudaMemcpyDeviceToHost, s));
        if (prof) DCR_CUDA_OK(cudaEventRecord(ev[7], s));
        if (flags & DCR_ITER_APPLY_Z) {
            // the result copy is already queued: wait for IT (event), not for the z-shift pass enqueued behind it --
            // the 0.1 ms read-modify-write of the source band then overlaps the host's fit and bookkeeping
            if (!ctx->e_copy) DCR_CUDA_OK(cudaEventCreateWithFlags(&ctx->e_copy, cudaEventDisableTiming));
            const cudaEvent_t ev_copy = ctx->e_copy;
            DCR_CUDA_OK(cudaEventRecord(ev_copy, s));
            const double zn = front->src_desc.ndv;
            zshift_dev_kernel<<<dcr_num_sms() * 8, 256, 0, s>>>(const_cast<float*>(front->src), front->src_desc.height,
                                                                 front->src_desc.width, front->src_stride, zn,
                                                                 1e-8 + 1e-5 * fabs(zn), (float)zn, d, engine_used);
            DCR_LAUNCH_OK("zshift_dev_kernel");
            DCR_CUDA_OK(cudaEventSynchronize(ev_copy));
        } else {
            DCR_CUDA_OK(cudaStreamSynchronize(s));
        }
        if (hd->nuth.fast_fallback && getenv("DCR_DEBUG"))
            fprintf(stderr, "[dcr] fast bin path declined: ncand=%u scale=%g rmin=%g rmax=%g\n", hd->nuth.fast_ncand,
                    hd->nuth.fast_scale, hd->nuth.rmin, hd->nuth.rmax);
        bin_engine = ((flags & DCR_ITER_RADIX_ONLY) || hd->nuth.fast_fallback) ? 0 : 1;
        if (hd->nuth.fast_fallback) {
            // fast bin path declined (heavy ties / degenerate 3-sigma range): exact 3-pass radix bins, same inputs
            if ((rc = nuth_bins_radix_enqueue(n, w.nw, grid, s))) return rc;
            DCR_CUDA_OK(cudaMemcpyAsync(hd, d, sizeof(IterDev), cudaMemcpyDeviceToHost, s));
            DCR_CUDA_OK(cudaStreamSynchronize(s));
        }
        if (!engine_used) break;
        // the one-pass selects verify themselves on the device; if any bracket missed (or hit heavy ties)
        // the whole iteration is redone with the 3-pass radix select (exact by construction)
        bool fb = false;
        for (int k = 0; k < 5; ++k) fb = fb || (hd->bsel[k].fallback != 0);
        if (!fb) break;
        if (getenv("DCR_DEBUG")) {
            for (int k = 0; k < 5; ++k) {
                const BselState& b = hd->bsel[k];
                fprintf(stderr, "[dcr] bsel[%d] fallback=%d klo=%08x khi=%08x shift=%d below=%llu total=%llu ncand=%u "
                        "prefix=[%u,%u] rem=[%lld,%lld] count=%lld done=%d\n", k, b.fallback, b.klo, b.khi,
                        b.shift, b.below, b.total, b.ncand, b.prefix[0], b.prefix[1], b.rem[0], b.rem[1], b.count, b.done);
            }
        }
        engine_used = 0;
        // the front kernel's mask bytes were updated in place by mask_update: rerun from scratch
    }

    memset(out, 0, sizeof(*out));
    if (prof)
        for (int k = 0; k < 7; ++k) DCR_CUDA_OK(cudaEventElapsedTime(&out->stage_ms[k], ev[k], ev[k + 1]));
    {
        const int rcf = iter_finalize(hd, engine_used, out, bin_engine, tan_slope);
        out->z_applied = ((flags & DCR_ITER_APPLY_Z) && !isnan(out->dz_med)) ? 1 : 0;
        return rcf;
    }
}

// ------------------------------------------------------------------------------------------
// The dem_align loop (dem_align.py:291-357, mode 'nuth') as one native call: no interpreter work between iterations
// ------------------------------------------------------------------------------------------
extern "C" int dcr_align_flush_z(dcr_align_state* st, void* stream) {
    DCR_ARG(st && st->src, "null pointer");
    if (st->src_ndz > 0) {
        int rc = dcr_apply_z_shift_seq(st->src, st->src_desc.height, st->src_desc.width, st->src_stride,
                                       st->src_desc.has_ndv ? st->src_desc.ndv : 0.0, st->src_dz, st->src_ndz, stream);
        if (rc) return rc;
        st->src_ndz = 0;
    }
    return 0;
}

extern "C" int dcr_align_nuth(dcr_align_state* st, int max_steps, int flags, dcr_align_iter* iters, int iters_cap,
                              int* n_iters, void* workspace, size_t workspace_bytes, void* stream) {
    DCR_ARG(st && workspace, "null pointer");
    DCR_ARG(st->ref && st->src && st->dh && st->slope && st->maskbits, "null raster pointer");
    DCR_ARG(st->src_ndz >= 0 && st->src_ndz <= 16, "0 <= src_ndz <= 16");
    int done_here = 0;
    if (n_iters) *n_iters = 0;
    while (!st->finished && st->exit_code == 0 && (max_steps <= 0 || done_here < max_steps)) {
        // compute_offset: memwarp_multi bookkeeping for the CURRENT source geotransform
        dcr_front_args a;
        memset(&a, 0, sizeof(a));
        dcr_grid grid;
        int rc = dcr_plan_intersection(&st->ref_desc, &st->src_desc, 0.0, &grid, &a.ref_geom, &a.src_geom);
        if (rc) return rc;
        if ((long long)grid.height * grid.width > st->out_capacity) {
            dcr_set_error("iteration outputs too small: %d x %d grid, capacity %lld pixels", grid.height, grid.width,
                          (long long)st->out_capacity);
            return -1;
        }
        a.ref = st->ref; a.ref_stride = st->ref_stride; a.ref_desc = st->ref_desc;
        a.src = st->src; a.src_stride = st->src_stride; a.src_desc = st->src_desc;
        if (st->smask) {
            dcr_warp_geom gm;
            if ((rc = dcr_plan_onto_grid(&st->smask_desc, &grid, &gm))) return rc;
            a.smask = st->smask; a.smask_stride = st->smask_stride;
            a.smask_h = st->smask_desc.height; a.smask_w = st->smask_desc.width;
            a.smask_nx0 = gm.nx0; a.smask_ny0 = gm.ny0;
        }
        a.h = grid.height; a.w = grid.width;
        a.ewres = grid.gt[1]; a.nsres = grid.gt[5];
        a.dst_ndv = 0.0;
        a.max_dz = st->max_dz; a.use_max_dz = st->remove_outliers ? 1 : 0;
        a.slope_lo = st->slope_lo; a.slope_hi = st->slope_hi;
        a.dh = st->dh; a.slope = st->slope; a.aspect = st->aspect; a.maskbits = st->maskbits;
        a.src_ndz = st->src_ndz;
        for (int k = 0; k < st->src_ndz; ++k) a.src_dz[k] = st->src_dz[k];
        const int f = (st->remove_outliers ? DCR_ITER_REMOVE_OUTLIERS : 0) |
                      (flags & (DCR_ITER_PROFILE | DCR_ITER_RADIX_ONLY | DCR_ITER_TAN_SLOPE));
        if ((rc = dcr_nuth_iteration(&a, f, &st->last, workspace, workspace_bytes, stream))) return rc;
        const dcr_iter_result& r = st->last;
        if (iters && done_here < iters_cap) {
            dcr_align_iter& it = iters[done_here];
            it.dx = r.dx; it.dy = r.dy; it.dz = r.dz; it.status = r.status; it.select_engine = r.select_engine;
            it.h = grid.height; it.w = grid.width;
            it.bin_engine = r.bin_engine; it._pad = 0;
            it.fit[0] = r.fit[0]; it.fit[1] = r.fit[1]; it.fit[2] = r.fit[2];
            it.dz_med = r.dz_med; it.med_slope = r.med_slope;
        }
        ++done_here;
        if (n_iters) *n_iters = done_here;
        if (r.status >= 1 && r.status <= 3) { st->exit_code = r.status; break; }   // compute_offset's sys.exit paths
        // totals (dz is np.float32 in the reference: float32 accumulation and float32 squares)
        const float dz32 = -r.dz_med;
        st->dx_total += r.dx;
        st->dy_total += r.dy;
        st->dz_total = st->dz_total + dz32;
        // apply_xy_shift: geotransform origin; apply_z_shift: deferred, folded into the band every few iterations
        st->src_desc.gt[0] += r.dx;
        st->src_desc.gt[3] += r.dy;
        st->src_dz[st->src_ndz++] = dz32;
        if (st->src_ndz >= DCR_ALIGN_MAX_PENDING && (rc = dcr_align_flush_z(st, stream))) return rc;
        ++st->n_done;
        const double dm = sqrt(r.dx * r.dx + r.dy * r.dy + (double)(dz32 * dz32));
        const double dm_total = sqrt(st->dx_total * st->dx_total + st->dy_total * st->dy_total +
                                     (double)(st->dz_total * st->dz_total));
        if (dm_total > st->max_offset) { st->exit_code = 5; break; }
        if (st->n_done >= st->max_iter || dm < st->tol) st->finished = 1;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// Row-band mode: one very large raster split into bands of output rows, one band per GPU.
// The iteration is cut into phases; after a phase the caller all-reduces (sum) the buffers the phase
// declares -- integer histograms / counters and three FP64 sums -- and calls the next phase.  All selections
// use the 3-pass radix select (its histograms are what gets all-reduced), so the integer results are
// bit-identical to the single-GPU path.  (BASELINE.json configs[4].)
// ------------------------------------------------------------------------------------------
enum {
    BP_FRONT = 0,
    // 5 phases per selection.  One-pass sample-bracket select (dcr_bsel_band_phase): sample | bracket + main pass | find +
    // refine | refine | final.  Radix select (DCR_ITER_RADIX_ONLY, and the rerun after a declined bracket): H0 | F0+H1 |
    // F1+H2 | F2+publish | -.
    BP_SEL0 = 1,
    BP_SEL1 = 6,
    BP_MASK = 11,
    BP_COMPRESS = 12,
    BP_SEL2 = 13,
    BP_SEL3 = 18,
    BP_SEL4 = 23,
    BP_PREPARE = 28,
    // bins.  Fast path (shared-memory coarse cells, then the fine cells of the median cells): A | findA + B | findB + C |
    // scatter of the per-rank slots | D.  Radix bins (DCR_ITER_RADIX_ONLY / rerun): H0 | F0+H1 | F1+H2 | F2 + counts | -.
    BP_BIN0 = 29,
    BP_BIN1 = 30,
    BP_BIN2 = 31,
    BP_BIN3 = 32,
    BP_BIN4 = 33,
    BP_COUNT = 34
};

__global__ void band_set_engine_kernel(IterDev* d, int use_bsel) {
    if (threadIdx.x == 0) d->selpad = use_bsel;
}

__global__ void band_rank_count_kernel(IterDev* d, int rank) {
    if (threadIdx.x == 0) d->rank_counts[rank] = (long long)d->count_final;
}
__global__ void band_base_kernel(IterDev* d, int rank, int nranks) {
    if (threadIdx.x) return;
    long long base = 0, tot = 0;
    for (int r = 0; r < nranks; ++r) { if (r < rank) base += d->rank_counts[r]; tot += d->rank_counts[r]; }
    d->band_base = base;
    d->total_final = tot;
}
__global__ void band_dense_n_kernel(IterDev* d, int rank) {
    if (threadIdx.x) return;
    const long long stride = d->stride, base = d->band_base, cnt = d->rank_counts[rank];
    d->dense_n = stride > 1 ? ((base + cnt + stride - 1) / stride - (base + stride - 1) / stride) : 0;
}
__global__ void __launch_bounds__(256) band_mom_reduce_kernel(const NuthPartial* __restrict__ part, int nparts, IterDev* d) {
    __shared__ NuthPartial sp[256];
    NuthPartial q; q.n = 0; q.s = 0; q.ss = 0;
    for (int k = threadIdx.x; k < nparts; k += 256) { q.n += part[k].n; q.s += part[k].s; q.ss += part[k].ss; }
    sp[threadIdx.x] = q;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sp[threadIdx.x].n += sp[threadIdx.x + o].n; sp[threadIdx.x].s += sp[threadIdx.x + o].s; sp[threadIdx.x].ss += sp[threadIdx.x + o].ss; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { d->mom3[0] = (double)sp[0].n; d->mom3[1] = sp[0].s; d->mom3[2] = sp[0].ss; d->mom3[3] = 0.0; }
}
// moments_final from the all-reduced sums (same arithmetic as nuth_moments_final_kernel)
__global__ void band_moments_final_kernel(IterDev* d, int remove_outliers) {
    NuthDev* nd = &d->nuth;
    if (threadIdx.x == 0) {
        const unsigned long long cnt = (unsigned long long)d->mom3[0];
        nd->n_common = cnt;
        nd->n_final = 0;
        float mean = __int_as_float(0x7fc00000), sd = mean;
        if (cnt) {
            const double mu = d->mom3[1] / (double)cnt;
            double var = d->mom3[2] / (double)cnt - mu * mu;
            if (var < 0) var = 0;
            mean = (float)mu;
            sd = (float)sqrt(var);
        }
        nd->mean_y = mean; nd->std_y = sd;
        if (remove_outliers) { const float fs = 3.0f * sd; nd->rmin = mean - fs; nd->rmax = mean + fs; }
        else { nd->rmin = -INFINITY; nd->rmax = INFINITY; }
        // set-up of the fast bin path, as nuth_moments_final_kernel does
        const float span = nd->rmax - nd->rmin;
        nd->fast_ncand = 0;
        nd->fast_fallback = 0;
        nd->fast_scale = 0.f;
        if (!(span > 0.0f) || isinf(span) || isnan(span)) nd->fast_fallback = 1;
        else nd->fast_scale = 65536.0f / span;
    }
    for (int k = threadIdx.x; k < NB; k += blockDim.x) {
        nd->bin_count[k] = 0ull;
        nd->prefix_lo[k] = nd->prefix_hi[k] = 0u;
        nd->same[k] = 1;
        nd->bin_med[k] = __int_as_float(0x7fc00000);
    }
}

static int band_carve(void* workspace, size_t workspace_bytes, long long n, int grid, IterWs* w) {
    DcrBump ws(workspace, workspace_bytes);
    w->d = ws.take<IterDev>(1);
    w->hist = ws.take<unsigned>(DCR_SEL_HIST_WORDS);
    const long long nch = (n + MU_CHUNK - 1) / MU_CHUNK;
    w->chunk_cnt = ws.take<unsigned>(nch);
    w->chunk_off = ws.take<long long>(nch);
    w->dense_cap = n < 6000016 ? n : 6000016;
    w->dense = ws.take<float>(w->dense_cap);
    w->bsel_scratch = ws.take<char>(dcr_bsel_scratch_bytes(n));
    w->bsel_scratch2 = ws.take<char>(dcr_bsel_scratch_bytes(n));
    w->bsel_scratch3 = ws.take<char>(dcr_bsel_scratch_bytes(n));
    w->tile_cap = dcr_front_list_cap(n);
    w->tile_list = ws.take<unsigned>(w->tile_cap);
    if (nuth_ws_carve(ws, n, grid, &w->nw)) return -1;
    w->nw.nd = &w->d->nuth;
    return 0;
}

extern "C" int dcr_band_nphases(void) { return BP_COUNT; }

// One elementary phase (BP_*); appends what must be summed over the ranks to reduce_ptr / reduce_count / reduce_kind.
static int band_phase_one(const dcr_front_args* front, int flags, int rank, int nranks, int phase, void* workspace,
                          size_t workspace_bytes, void** reduce_ptr, int64_t* reduce_count, int* reduce_kind,
                          int* nreduce, void* stream) {
    const int remove_outliers = (flags & DCR_ITER_REMOVE_OUTLIERS) ? 1 : 0;
    const int rb = front->row_end > 0 ? front->row_begin : 0, re = front->row_end > 0 ? front->row_end : front->h;
    const long long n = (long long)(re - rb) * front->w;
    DCR_ARG(n > 0, "empty band");
    const int grid = dcr_num_sms() * 8;
    DCR_ARG(workspace_bytes >= iter_ws_bytes(n, grid), "workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    IterWs w;
    if (band_carve(workspace, workspace_bytes, n, grid, &w)) return -1;
    IterDev* d = w.d;
    const long long nch = (n + MU_CHUNK - 1) / MU_CHUNK;
    auto want = [&](void* p, int64_t cnt, int kind) { reduce_ptr[*nreduce] = p; reduce_count[*nreduce] = cnt; reduce_kind[*nreduce] = kind; ++*nreduce; };
    int rc = 0;

    // selections: view k
    auto view_of = [&](int k) {
        SelView v;
        memset(&v, 0, sizeof(v));
        v.x = front->dh; v.m = front->maskbits; v.n = n;
        if (k == 0) { v.reject = DCR_M_BASE | DCR_M_MAXDZ; }
        else if (k == 1) { v.reject = DCR_M_BASE | DCR_M_MAXDZ; v.transform = 1; v.center_ptr = &d->selres[0]; }
        else if (k == 2) { v.reject = DCR_M_DIFF; }
        else if (k == 3) { v.x = w.dense; v.m = nullptr; v.n = 0; v.n_ptr = &d->dense_n; }
        else { v.x = front->slope; v.reject = DCR_M_SLOPE; }
        return v;
    };
    const int use_bsel = (flags & DCR_ITER_RADIX_ONLY) ? 0 : 1;
    const int use_fast_bins = (use_bsel && !(flags & DCR_ITER_RADIX_BINS)) ? 1 : 0;
    auto sel_phase = [&](int k, int sub) -> int {
        if ((k == 0 || k == 1) && !remove_outliers) return 0;
        SelView v = view_of(k);
        if (use_bsel) {
            // (selections that run in lockstep -- 0 with 4, 2 with 3 -- need scratch of their own)
            void* scr = (k == 4) ? w.bsel_scratch2 : (k == 3) ? w.bsel_scratch3 : w.bsel_scratch;
            return dcr_bsel_band_phase(v, 50.0, &d->bsel[k], scr, k == 3 ? w.dense_cap : n, rank, nranks, sub, s,
                                       &d->selres[k], reduce_ptr, reduce_count, reduce_kind, nreduce);
        }
        // radix: sub 0: H0 | 1: F0+H1 | 2: F1+H2 | 3: F2 + publish | 4: -
        if (sub > 3) return 0;
        int r2 = 0;
        if (sub == 0) DCR_CUDA_OK(cudaMemsetAsync(w.hist, 0, DCR_SEL_HIST_WORDS * sizeof(unsigned), s));
        if (sub > 0) r2 = dcr_sel_find_pass(&d->sel[k], w.hist, 50.0, sub - 1, s);
        if (r2) return r2;
        if (sub < 3) {
            r2 = dcr_sel_hist_pass(v, &d->sel[k], w.hist, sub, s);
            if (r2) return r2;
            want(w.hist, DCR_SEL_HIST_WORDS, 0);
        } else {
            sel_publish_kernel<<<1, 32, 0, s>>>(d, k, 0);
        }
        return 0;
    };

    if (phase == BP_FRONT) {
        DCR_CUDA_OK(cudaMemsetAsync(d, 0, sizeof(IterDev), s));
        band_set_engine_kernel<<<1, 32, 0, s>>>(d, use_bsel | (use_fast_bins << 1));
        dcr_front_args fa = *front;
        if (!remove_outliers) fa.use_max_dz = 0;
        // like the single-GPU iteration: the front kernel emits the aspect-bin codes (so the aspect raster is optional)
        // and may use the fast interior-tile kernel on the band
        if ((rc = dcr_front_launch(&fa, d->counters, s, w.nw.bin, 0, w.tile_list, w.tile_cap))) return rc;
        want(d->counters, 4, 1);
    } else if (phase >= BP_SEL0 && phase < BP_SEL1) {
        // one-pass engine: the slope median (independent of the MAD chain) advances in lockstep with the first dh median --
        // the same sub-phases, their all-reduces issued together (5 collective latencies less per iteration)
        rc = sel_phase(0, phase - BP_SEL0);
        if (!rc && use_bsel) rc = sel_phase(4, phase - BP_SEL0);
    } else if (phase >= BP_SEL1 && phase < BP_MASK) {
        rc = sel_phase(1, phase - BP_SEL1);
    } else if (phase == BP_MASK) {
        mad_thresholds_kernel<<<1, 32, 0, s>>>(d, remove_outliers);
        // MAD filter + the Nuth-Kaab samples (y, partial moments, bin codes of the removed pixels) in one pass
        mask_update_kernel<true><<<(unsigned)((nch + MU_KCH - 1) / MU_KCH), 256, 0, s>>>(
            front->dh, front->maskbits, n, d, w.chunk_cnt, front->slope, w.nw.y, w.nw.bin, w.nw.part, 0, 0);
        band_rank_count_kernel<<<1, 32, 0, s>>>(d, rank);
        want(&d->count_mad, 2 + 8, 1);  // count_mad, count_final, rank_counts[8] are contiguous
    } else if (phase == BP_COMPRESS) {
        // after the all-reduce count_final is global; the per-band counts give this band's rank offset
        band_base_kernel<<<1, 32, 0, s>>>(d, rank, nranks);
        if ((rc = dcr_chunk_scan_enqueue(w.chunk_cnt, (nch + MU_KCH - 1) / MU_KCH, w.chunk_off, &d->dense_n /* scratch */, s)))
            return rc;
        stride_kernel<<<1, 32, 0, s>>>(d);
        band_dense_n_kernel<<<1, 32, 0, s>>>(d, rank);
        if ((rc = dcr_compress_enqueue(front->dh, front->maskbits, DCR_M_DIFF, n, w.chunk_off, 1, &d->stride, w.dense,
                                       w.dense_cap, s, &d->band_base, MU_KCH)))
            return rc;
    } else if (phase >= BP_SEL2 && phase < BP_SEL3) {
        rc = sel_phase(2, phase - BP_SEL2);                          // med_dh over the final mask ...
        if (!rc && use_bsel) rc = sel_phase(3, phase - BP_SEL2);     // ... in lockstep with the subsample median
    } else if (phase >= BP_SEL3 && phase < BP_SEL4) {
        if (!use_bsel) rc = sel_phase(3, phase - BP_SEL3);
    } else if (phase >= BP_SEL4 && phase < BP_PREPARE) {
        if (!use_bsel) rc = sel_phase(4, phase - BP_SEL4);
    } else if (phase == BP_PREPARE) {
        band_mom_reduce_kernel<<<1, 256, 0, s>>>(w.nw.part, (int)((nch + MU_KCH - 1) / MU_KCH), d);   // partials of the mask update
        want(d->mom3, 4, 2);
    } else if (phase >= BP_BIN0 && use_fast_bins) {
        // fast bins across ranks (the same engine switch as the selections: DCR_ITER_RADIX_ONLY takes the radix bins)
        NuthDev* nd = &d->nuth;
        unsigned* gh = w.nw.ghist;
        const int nsm = dcr_num_sms();
        if (phase == BP_BIN0) {
            band_moments_final_kernel<<<1, 256, 0, s>>>(d, 1);   // trim of y always on (dem_align.py:147)
            DcrCtx* ctx = dcr_ctx();
            if (!ctx) return -1;
            const size_t smemA = (size_t)NB * FB_COARSE * sizeof(unsigned);
            if (!(ctx->attrs & DCR_ATTR_FBA)) {   // per device
                DCR_CUDA_OK(cudaFuncSetAttribute(fb_passA_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA));
                ctx->attrs |= DCR_ATTR_FBA;
            }
            DCR_CUDA_OK(cudaMemsetAsync(gh, 0, (size_t)FB_BAND_WORDS * sizeof(unsigned), s));
            fb_passA_kernel<<<nsm * 2, 512, smemA, s>>>(w.nw.y, w.nw.bin, n, nd, gh);
            want(gh + FB_HA_OFF, NB * FB_COARSE, 0);
        } else if (phase == BP_BIN1) {
            fb_findA_kernel<<<(NB + 127) / 128, 128, 0, s>>>(nd, gh);     // global counts + coarse cells, same on every rank
            fb_passB_kernel<<<nsm * 16, 256, 0, s>>>(w.nw.y, w.nw.bin, n, nd, gh, w.nw.cand, w.nw.cand_cap);
            fb_band_flag_kernel<<<1, 32, 0, s>>>(nd, w.nw.cand_cap, gh + FB_FLAG_OFF);
            want(gh, FB_HB_WORDS, 0);
            want(gh + FB_FLAG_OFF, 1, 0);
        } else if (phase == BP_BIN2) {
            fb_findB_kernel<<<(NB + 7) / 8, 256, 0, s>>>(nd, gh, w.nw.cand_cap, gh + FB_FLAG_OFF);
            fb_passC_kernel<<<nsm * 8, 256, 0, s>>>(nd, w.nw.cand, gh + FB_LSLOT_OFF, gh + FB_LSLOTN_OFF);
            fb_band_counts_kernel<<<(NB * 2 + 255) / 256, 256, 0, s>>>(gh + FB_LSLOTN_OFF, gh + FB_RCNT_OFF, rank);
            want(gh + FB_RCNT_OFF, (int64_t)nranks * NB * 2, 0);
        } else if (phase == BP_BIN3) {
            fb_band_scatter_kernel<<<(NB * 2 + 7) / 8, 256, 0, s>>>(gh + FB_LSLOT_OFF, gh + FB_RCNT_OFF, rank, nranks,
                                                                    gh + FB_SLOT_OFF, gh + FB_SLOTN_OFF, nd);
            want(gh + FB_SLOT_OFF, (int64_t)NB * 2 * FB_SLOT, 0);
        } else {
            fb_passD_kernel<<<(NB + 7) / 8, 256, 0, s>>>(nd, gh);
        }
    } else if (phase == BP_BIN0) {
        band_moments_final_kernel<<<1, 256, 0, s>>>(d, 1);   // trim of y always on (dem_align.py:147)
        DCR_CUDA_OK(cudaMemsetAsync(w.nw.ghist, 0, (size_t)2 * NB * 2048 * sizeof(unsigned), s));
        bin_hist_kernel<0><<<grid, 256, 0, s>>>(w.nw.y, w.nw.bin, n, &d->nuth, w.nw.ghist);
        want(w.nw.ghist, (int64_t)2 * NB * 2048, 0);
        want(d->nuth.bin_count, NB, 1);
    } else if (phase == BP_BIN1) {
        bin_find_kernel<0><<<NB, 256, 0, s>>>(&d->nuth, w.nw.ghist);
        bin_hist_kernel<1><<<grid, 256, 0, s>>>(w.nw.y, w.nw.bin, n, &d->nuth, w.nw.ghist);
        want(w.nw.ghist, (int64_t)2 * NB * 2048, 0);
    } else if (phase == BP_BIN2) {
        bin_find_kernel<1><<<NB, 256, 0, s>>>(&d->nuth, w.nw.ghist);
        bin_hist_kernel<2><<<grid, 256, 0, s>>>(w.nw.y, w.nw.bin, n, &d->nuth, w.nw.ghist);
        want(w.nw.ghist, (int64_t)2 * NB * 2048, 0);
    } else if (phase == BP_BIN3) {
        bin_find_kernel<2><<<NB, 256, 0, s>>>(&d->nuth, w.nw.ghist);
        nuth_count_final_kernel<<<1, 32, 0, s>>>(&d->nuth);
    }
    if (rc) return rc;
    DCR_LAUNCH_OK("band phase kernels");
    return 0;
}

// Runs phase `phase` of the iteration on this rank's band (front->row_begin/row_end).  On return reduce_ptr[k] /
// reduce_count[k] / reduce_kind[k] (k < *nreduce <= 8) describe device buffers inside `workspace` that must be
// summed over all ranks before the next phase: kind 0 = int32 (histograms), 1 = int64 (counters), 2 = float64.
//
// With the one-pass selections (no DCR_ITER_RADIX_ONLY) the elementary phases that do not depend on each other share a
// phase boundary, so that their collectives are issued together -- 16 boundaries instead of 34: the front end with the
// sample of the first selections; the moment sums with the mask update; the compress step, the selections over the final
// mask and the bin statistics in lockstep.  The remaining phase numbers are empty then.
extern "C" int dcr_band_phase(const dcr_front_args* front, int flags, int rank, int nranks, int phase, void* workspace,
                              size_t workspace_bytes, void** reduce_ptr, int64_t* reduce_count, int* reduce_kind,
                              int* nreduce, void* stream) {
    DCR_ARG(front && workspace && reduce_ptr && reduce_count && reduce_kind && nreduce, "null pointer");
    DCR_ARG(nranks >= 1 && nranks <= 8 && rank >= 0 && rank < nranks, "1 <= nranks <= 8");
    DCR_ARG(phase >= 0 && phase < BP_COUNT, "bad phase");
    *nreduce = 0;
    if (flags & DCR_ITER_RADIX_ONLY)
        return band_phase_one(front, flags, rank, nranks, phase, workspace, workspace_bytes, reduce_ptr, reduce_count,
                              reduce_kind, nreduce, stream);
    static const signed char plan[16][3] = {
        {BP_FRONT, BP_SEL0 + 0, -1},
        {BP_SEL0 + 1, -1, -1}, {BP_SEL0 + 2, -1, -1}, {BP_SEL0 + 3, -1, -1}, {BP_SEL0 + 4, -1, -1},
        {BP_SEL1 + 0, -1, -1}, {BP_SEL1 + 1, -1, -1}, {BP_SEL1 + 2, -1, -1}, {BP_SEL1 + 3, -1, -1}, {BP_SEL1 + 4, -1, -1},
        {BP_MASK, BP_PREPARE, -1},
        {BP_COMPRESS, BP_SEL2 + 0, BP_BIN0},
        {BP_SEL2 + 1, BP_BIN1, -1}, {BP_SEL2 + 2, BP_BIN2, -1}, {BP_SEL2 + 3, BP_BIN3, -1}, {BP_SEL2 + 4, BP_BIN4, -1}};
    if (phase >= 16) return 0;
    for (int q = 0; q < 3; ++q) {
        if (plan[phase][q] < 0) break;
        const int rc = band_phase_one(front, flags, rank, nranks, plan[phase][q], workspace, workspace_bytes, reduce_ptr,
                                      reduce_count, reduce_kind, nreduce, stream);
        if (rc) return rc;
    }
    return 0;
}

// After the last phase: copy the (identical on every rank) result block to the host and finish like dcr_nuth_iteration.
extern "C" int dcr_band_result(const dcr_front_args* front, void* workspace, size_t workspace_bytes, dcr_iter_result* out,
                               void* stream) {
    DCR_ARG(front && workspace && out, "null pointer");
    const int rb = front->row_end > 0 ? front->row_begin : 0, re = front->row_end > 0 ? front->row_end : front->h;
    const long long n = (long long)(re - rb) * front->w;
    const int grid = dcr_num_sms() * 8;
    IterWs w;
    if (band_carve(workspace, workspace_bytes, n, grid, &w)) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    DcrCtx* ctx = dcr_ctx();
    if (!ctx) return -1;
    IterDev* hd = (IterDev*)dcr_ctx_pinned(ctx, sizeof(IterDev));
    if (!hd) { dcr_set_error("pinned host allocation failed"); return -1; }
    DCR_CUDA_OK(cudaMemcpyAsync(hd, w.d, sizeof(IterDev), cudaMemcpyDeviceToHost, s));
    DCR_CUDA_OK(cudaStreamSynchronize(s));
    memset(out, 0, sizeof(*out));
    const int engine = (hd->selpad & 1) ? 1 : 0, fast_bins = (hd->selpad & 2) ? 1 : 0;
    {
        // a one-pass selection (bracket missed / heavy ties / overflow) or the fast bins (heavy ties / degenerate 3-sigma
        // range) declined -- on all-reduced data, so on every rank alike: the caller reruns the iteration with
        // DCR_ITER_RADIX_ONLY resp. DCR_ITER_RADIX_BINS; select_engine / bin_engine say which one it was
        bool sfb = false;
        for (int k = 0; engine && k < 5; ++k) sfb = sfb || (hd->bsel[k].fallback != 0);
        const bool bfb = fast_bins && hd->nuth.fast_fallback != 0;
        if (sfb || bfb) {
            out->status = DCR_STATUS_RERUN_RADIX;
            out->select_engine = sfb ? 0 : engine;
            out->bin_engine = bfb ? 0 : fast_bins;
            return 0;
        }
    }
    return iter_finalize(hd, engine, out, fast_bins);
}
