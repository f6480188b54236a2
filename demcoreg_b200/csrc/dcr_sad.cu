// K7b SAD offset search, batched over the whole (2p+1)^2 shift window.
//
// Replaces reference coreglib.compute_offset_sad (coreglib.py:65-115): for every offset (i, j) of the window,
//   diff = dem1[i:i+kh, j:j+kw] - dem2[p:-p, p:-p]            (float32, masked)
//   diff = masked_outside(diff, *malib.calcperc(diff, (25, 75)))   (exact percentiles, linear interpolation)
//   m[i, j] = abs(diff).sum() / diff.count()
// The reference runs the (2p+1)^2 offsets one after the other (two exact quantiles of kh*kw values each:
// 1.4 s per offset at 4096^2).  Here ALL offsets are processed by the same few passes over the two rasters:
//
//   prep     masked pixels -> NaN (a NaN difference takes part in nothing), kernel = src[p:-p, p:-p] made contiguous
//   sample   one CTA per offset: 32768 stratified samples of diff -> a bracket [lo, hi] around each of the two
//            order statistics (sample ranks -+ 5 sigma), localised by three rounds of a 1024-bin histogram
//   pass 1   every difference of every offset, out of shared-memory tiles (thread = pixel column, 8 consecutive
//            offsets unrolled so that the bracket bounds and the counters of an offset live in registers): exact
//            counts below / inside the brackets; the ~5 % inside are parked in thread-private shared-memory slots and
//            drained into a 512-bin histogram of their bracket (shared-memory atomics)
//   find     per offset: exact integer ranks of P25 / P75 (float64 virtual index, as numpy) -> the histogram bins
//            that hold them, VERIFIED to lie inside the brackets; the bin edges as exact float thresholds (bisection
//            over the monotone bin function)
//   pass 2   every difference again: sum |diff| strictly between the two candidate intervals (FP32 per 8 rows,
//            FP64 across), and the few hundred differences inside the candidate bins appended per offset
//   final    per offset: sort the candidates, P25 / P75 by numpy's lerp, trimmed sum and count -> m
//
// An offset whose verification fails (heavy ties, degenerate or overlapping brackets, candidate overflow) is
// redone by the exact per-offset path (materialised diff + two K3 selections), which is also what
// DCR_SAD_PER_OFFSET runs for every offset (tests compare the two).  Both passes are FP32/integer-issue bound
// out of shared memory (about 12 instructions per difference), not HBM bound: the rasters are read
// (2p+1)^2 / 8 times from L2.
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include "dcr_common.cuh"
#include "dcr_internal.cuh"

#define SAD_B 512        // histogram bins per bracket
#define SAD_JC 8         // consecutive column offsets per CTA (per-thread register state scales with it)
#define SAD_TU 8         // tile rows
#define SAD_TV 256       // tile columns = threads per CTA
#define SAD_DRAIN 4      // rows between two drains of the parked values (SAD_DRAIN * SAD_JC private slots per thread)
#define SAD_PRIV (SAD_DRAIN * SAD_JC)
#define SAD_CAP 2048     // candidates per (offset, quantile)
#define SAD_S 32768      // samples per offset
#define SAD_QNAN __int_as_float(0x7fc00000)

struct SadOff {
    // sample kernel
    float lo[2], hi[2], sc[2];   // closed brackets of P25 / P75 and their bin scales SAD_B / (hi - lo)
    float mid;                   // a value between the two brackets (which bracket a parked value belongs to)
    int flag;                    // != 0: batched path not applicable / failed verification -> per-offset exact path
    // pass 1
    unsigned long long n_lt, n_ge, n_rng;   // d < lo[0] ; d >= lo[0] ; lo[0] <= d <= hi[1]
    // find
    float ta[2], tb[2];          // candidate intervals [ta, tb) of P25 / P75; the untouched middle is [tb[0], ta[1])
    long long n_mid;             // differences inside the middle
    unsigned k[2], k2[2];        // ranks of the two neighbouring order statistics inside the sorted candidates
    double g[2];                 // interpolation weights
    unsigned want[2];            // candidates pass 2 must find
    unsigned ncand[2];           // candidates pass 2 found
    long long count;             // unmasked differences
    int done;                    // m already final (no unmasked difference)
    int pad;
};

__device__ __forceinline__ unsigned sad_trunc_bits(float d) { return __float_as_uint(d) & ~7u; }
// bin of a parked difference inside its bracket: monotone non-decreasing in d (truncation of the low mantissa bits,
// subtraction, multiplication, conversion and clamping all are)
__device__ __forceinline__ int sad_bin(float dt, float lo, float sc) {
    const float t = (dt - lo) * sc;
    int b = (int)t;
    return b < 0 ? 0 : (b > SAD_B - 1 ? SAD_B - 1 : b);
}

// ------------------------------------------------------------------------------------------
// prep: NaN-filled copies
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sad_prep_kernel(const float* __restrict__ x, const unsigned char* __restrict__ m,
                                                       long long stride, int r0, int c0, int h, int w,
                                                       float* __restrict__ out) {
    const long long n = (long long)h * w;
    for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < n; k += (long long)gridDim.x * 256) {
        const int i = (int)(k / w), j = (int)(k - (long long)i * w);
        const long long s = (long long)(r0 + i) * stride + c0 + j;
        out[k] = (m && m[s]) ? SAD_QNAN : x[s];
    }
}

// ------------------------------------------------------------------------------------------
// sample: brackets of the two quantiles of every offset
// ------------------------------------------------------------------------------------------
// Key range [*base, *base + *span) that holds the sample ranks ra..rb (three rounds of a 1024-bin histogram).
__device__ void sad_localize(const unsigned* __restrict__ s, long long ra, long long rb, unsigned kmin, unsigned kmax,
                             unsigned* hist, unsigned* s_scratch, unsigned* s_bin, unsigned* s_ex, unsigned* base_out,
                             unsigned long long* span_out) {
    unsigned base = kmin;
    unsigned long long span = (unsigned long long)kmax - kmin + 1ull;
    for (int round = 0; round < 3; ++round) {
        int shift = 0;
        while ((span >> shift) > 1024ull) ++shift;
        hist[threadIdx.x] = 0u;
        __syncthreads();
        for (int q = 0; q < SAD_S / 1024; ++q) {
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
        unsigned long long nt = (unsigned long long)base + (((unsigned long long)s_bin[1] + 1ull) << shift);
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

__global__ void __launch_bounds__(1024) sad_sample_kernel(const float* __restrict__ rz, const float* __restrict__ kz, int W,
                                                          int kh, int kw, int J, SadOff* __restrict__ off) {
    extern __shared__ unsigned s_keys[];   // SAD_S
    __shared__ unsigned hist[1024];
    __shared__ unsigned s_scratch[33];
    __shared__ unsigned s_valid, s_kmin, s_kmax;
    __shared__ unsigned s_bin[2], s_ex[2];
    const int o = blockIdx.x, oi = o / J, oj = o - oi * J;
    const long long n = (long long)kh * kw;
    const bool all = n <= SAD_S;
    if (threadIdx.x == 0) { s_valid = 0u; s_kmin = 0xFFFFFFFFu; s_kmax = 0u; }
    __syncthreads();
    unsigned nv = 0, kmn = 0xFFFFFFFFu, kmx = 0u;
    for (int q0 = 0; q0 < SAD_S / 1024; q0 += 8) {
        float a[8], b[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {   // 16 independent gathers in flight per thread
            const int k = (q0 + u) * 1024 + threadIdx.x;
            long long i = k;
            if (!all) {
                const unsigned long long lo = ((unsigned long long)k * (unsigned long long)n) / SAD_S;
                const unsigned long long hi = ((unsigned long long)(k + 1) * (unsigned long long)n) / SAD_S;
                const unsigned long long hsh = ((unsigned long long)k * 0x9E3779B97F4A7C15ull) >> 20;
                i = (long long)(lo + hsh % (hi - lo));
            }
            a[u] = b[u] = SAD_QNAN;
            if (i < n) {
                const int pu = (int)(i / kw), pv = (int)(i - (long long)pu * kw);
                a[u] = __ldg(rz + (long long)(pu + oi) * W + pv + oj);
                b[u] = __ldg(kz + i);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float d = a[u] - b[u];
            unsigned key = 0xFFFFFFFFu;
            if (d == d) { key = dcr_f2key(d + 0.0f); ++nv; kmn = min(kmn, key); kmx = max(kmx, key); }
            s_keys[(q0 + u) * 1024 + threadIdx.x] = key;
        }
    }
    nv = __reduce_add_sync(0xffffffffu, nv);
    kmn = __reduce_min_sync(0xffffffffu, kmn);
    kmx = __reduce_max_sync(0xffffffffu, kmx);
    if ((threadIdx.x & 31) == 0) { if (nv) atomicAdd(&s_valid, nv); atomicMin(&s_kmin, kmn); atomicMax(&s_kmax, kmx); }
    __syncthreads();
    const long long m = s_valid;
    unsigned klo[2] = {0u, 0u}, khi[2] = {0u, 0u};
    bool open = m < 256;   // too few samples for a meaningful bracket
    if (!open) {
        for (int r = 0; r < 2; ++r) {
            const double f = r == 0 ? 0.25 : 0.75;
            const double rk = f * (double)(m - 1);
            const double dd = 5.0 * sqrt((double)m * f * (1.0 - f)) + 8.0;
            const long long a0 = (long long)floor(rk - dd), b0 = (long long)ceil(rk + dd);
            if (a0 < 0 || b0 >= m) { open = true; break; }
            unsigned base;
            unsigned long long span;
            sad_localize(s_keys, a0, b0, s_kmin, s_kmax, hist, s_scratch, s_bin, s_ex, &base, &span);
            klo[r] = base;
            khi[r] = (unsigned)((unsigned long long)base + span - 1ull);
        }
    }
    if (threadIdx.x == 0) {
        SadOff q;
        memset(&q, 0, sizeof(q));
        q.flag = open ? 1 : 0;
        if (!open) {
            for (int r = 0; r < 2; ++r) {
                q.lo[r] = dcr_key2f(klo[r]);
                q.hi[r] = dcr_key2f(khi[r]);
                const float w = q.hi[r] - q.lo[r];
                q.sc[r] = (w > 0.0f && !isinf(w)) ? (float)SAD_B / w : 0.0f;
            }
            // the brackets must be well separated: a parked value is attributed to one of them by `mid`, after its low
            // mantissa bits were cleared (<= 7 ulp towards zero)
            q.mid = 0.5f * q.hi[0] + 0.5f * q.lo[1];
            const float big = fmaxf(fabsf(q.hi[0]), fabsf(q.lo[1]));
            if (!(q.lo[1] - q.hi[0] > 1e-5f * big) || !(q.mid > q.hi[0]) || !(q.mid < q.lo[1]) || isinf(q.lo[0]) || isinf(q.hi[1]))
                q.flag = 1;
        }
        off[o] = q;
    }
}

// ------------------------------------------------------------------------------------------
// the tile loop shared by pass 1 and pass 2: thread = pixel column, SAD_JC consecutive column offsets unrolled
// ------------------------------------------------------------------------------------------
struct SadTiles {
    float* ks;   // [SAD_TU][SAD_TV]
    float* rs;   // [SAD_TU][SAD_TV + SAD_JC]
};
#define SAD_RP (SAD_TV + SAD_JC)

__device__ __forceinline__ void sad_load_tile(const float* __restrict__ rz, const float* __restrict__ kz, int W, int kh, int kw,
                                              int oi, int j0, int u0, int v0, const SadTiles& t) {
    const int tid = threadIdx.x;
    const bool cok = v0 + tid < kw;
#pragma unroll
    for (int r = 0; r < SAD_TU; ++r) {
        const bool rok = u0 + r < kh;
        t.ks[r * SAD_TV + tid] = (rok && cok) ? __ldg(kz + (long long)(u0 + r) * kw + v0 + tid) : SAD_QNAN;
        const float* rrow = rz + (long long)(u0 + r + oi) * W + v0 + j0;
        t.rs[r * SAD_RP + tid] = (rok && v0 + j0 + tid < W) ? __ldg(rrow + tid) : SAD_QNAN;
        if (tid < SAD_JC) t.rs[r * SAD_RP + SAD_TV + tid] = (rok && v0 + j0 + SAD_TV + tid < W) ? __ldg(rrow + SAD_TV + tid) : SAD_QNAN;
    }
}

// ------------------------------------------------------------------------------------------
// pass 1
// ------------------------------------------------------------------------------------------
#define SAD_P1_SMEM ((SAD_TU * SAD_TV + SAD_TU * SAD_RP + SAD_JC * 2 * SAD_B + SAD_PRIV * SAD_TV + SAD_JC * 8 + SAD_JC * 4) * 4)

__global__ void __launch_bounds__(SAD_TV, 2) sad_pass1_kernel(const float* __restrict__ rz, const float* __restrict__ kz, int W,
                                                              int kh, int kw, int J, int nchunk, SadOff* __restrict__ off,
                                                              unsigned* __restrict__ ghist) {
    extern __shared__ __align__(16) unsigned char sad_smem[];
    float* ks = reinterpret_cast<float*>(sad_smem);
    float* rs = ks + SAD_TU * SAD_TV;
    unsigned* sh_hist = reinterpret_cast<unsigned*>(rs + SAD_TU * SAD_RP);   // [SAD_JC][2][SAD_B]
    unsigned* sh_priv = sh_hist + SAD_JC * 2 * SAD_B;                        // slot k of thread t at [k * SAD_TV + t]
    float* sh_b = reinterpret_cast<float*>(sh_priv + SAD_PRIV * SAD_TV);     // [SAD_JC][8]: lo0 hi0 lo1 hi1 sc0 sc1 mid -
    unsigned* sh_cnt = reinterpret_cast<unsigned*>(sh_b + SAD_JC * 8);       // [SAD_JC][4] (3 used)
    const int tid = threadIdx.x;
    const int oi = blockIdx.y / nchunk, j0 = (blockIdx.y - oi * nchunk) * SAD_JC;
    const float inf = INFINITY;
    float lo0[SAD_JC], hi0[SAD_JC], lo1[SAD_JC], hi1[SAD_JC];
    unsigned n_lt[SAD_JC], n_ge[SAD_JC], n_rng[SAD_JC];
    bool any = false;
#pragma unroll
    for (int jj = 0; jj < SAD_JC; ++jj) {
        lo0[jj] = hi0[jj] = lo1[jj] = hi1[jj] = inf;   // nothing is counted or parked for an absent / flagged offset
        n_lt[jj] = n_ge[jj] = n_rng[jj] = 0u;
        if (j0 + jj < J) {
            const SadOff* q = off + oi * J + j0 + jj;
            if (!q->flag) { lo0[jj] = q->lo[0]; hi0[jj] = q->hi[0]; lo1[jj] = q->lo[1]; hi1[jj] = q->hi[1]; any = true; }
        }
    }
    if (!any) return;   // CTA-uniform
    for (int k = tid; k < SAD_JC * 2 * SAD_B; k += SAD_TV) sh_hist[k] = 0u;
    if (tid < SAD_JC) {
        const int jj = tid;
        float b[8] = {inf, inf, inf, inf, 0.f, 0.f, inf, 0.f};
        if (j0 + jj < J) {
            const SadOff* q = off + oi * J + j0 + jj;
            if (!q->flag) { b[0] = q->lo[0]; b[1] = q->hi[0]; b[2] = q->lo[1]; b[3] = q->hi[1]; b[4] = q->sc[0]; b[5] = q->sc[1]; b[6] = q->mid; }
        }
        for (int k = 0; k < 8; ++k) sh_b[jj * 8 + k] = b[k];
        sh_cnt[jj * 4 + 0] = sh_cnt[jj * 4 + 1] = sh_cnt[jj * 4 + 2] = sh_cnt[jj * 4 + 3] = 0u;
    }
    SadTiles tl;
    tl.ks = ks; tl.rs = rs;
    // thread-private parking slots: slot k of thread t at sh_priv[k * SAD_TV + t]; `ptr` = shared address of the next free one
    const unsigned mine = (unsigned)__cvta_generic_to_shared(sh_priv + tid);
    unsigned ptr = mine;
    // tag = jj, read back from memory so that it stays in a register (a compile-time constant would be re-materialised
    // before every use): clearing the low mantissa bits and inserting the tag is then ONE three-input logic instruction
    unsigned tag[SAD_JC];
#pragma unroll
    for (int jj = 0; jj < SAD_JC; ++jj) tag[jj] = (unsigned)jj + (unsigned)off[oi * J + (j0 + jj < J ? j0 + jj : J - 1)].pad;
    auto drain = [&]() {
        const unsigned c = (ptr - mine) / (SAD_TV * 4);
        const unsigned cmax = __reduce_max_sync(0xffffffffu, c);
        for (unsigned k = 0; k < cmax; ++k) {
            if (k < c) {
                unsigned pk;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(pk) : "r"(mine + k * (SAD_TV * 4)));
                const int jj = (int)(pk & 7u);
                const float dt = __uint_as_float(pk & ~7u);
                const float* b = sh_b + jj * 8;
                const int r = dt > b[6] ? 1 : 0;
                atomicAdd(&sh_hist[(jj * 2 + r) * SAD_B + sad_bin(dt, b[2 * r], b[4 + r])], 1u);
            }
        }
        ptr = mine;
    };
    const int ntv = (kw + SAD_TV - 1) / SAD_TV, ntu = (kh + SAD_TU - 1) / SAD_TU;
    for (int t = blockIdx.x; t < ntv * ntu; t += gridDim.x) {
        const int tu = t / ntv, tv = t - tu * ntv;
        __syncthreads();   // the previous tile is no longer read (also orders the histogram / bound set-up on the first trip)
        sad_load_tile(rz, kz, W, kh, kw, oi, j0, tu * SAD_TU, tv * SAD_TV, tl);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < SAD_TU; ++r) {
            const float kv = ks[r * SAD_TV + tid];
            const float* rr = rs + r * SAD_RP + tid;
#pragma unroll
            for (int jj = 0; jj < SAD_JC; ++jj) {
                const float d = rr[jj] - kv;
                // The classification step in PTX, so that every counter update is ONE predicated add and the parking ONE
                // predicated store + pointer bump (the compiler's own if-conversion emitted select pairs: 19 instructions
                // per difference instead of 13).  NaN (a masked pixel) fails every comparison.
                //   ge  = d >= lo0          rng = ge && d <= hi1          lt = d < lo0
                //   in  = rng && (d <= hi0 || d >= lo1)  -> park (bits(d) & ~7) | tag
                asm volatile(
                    "{\n\t"
                    ".reg .pred pge, prng, plt, pa, pb, pin;\n\t"
                    ".reg .b32 t;\n\t"
                    "setp.ge.f32 pge, %4, %5;\n\t"
                    "setp.le.and.f32 prng, %4, %8, pge;\n\t"
                    "setp.lt.f32 plt, %4, %5;\n\t"
                    "setp.le.f32 pa, %4, %6;\n\t"
                    "setp.ge.or.f32 pb, %4, %7, pa;\n\t"
                    "and.pred pin, prng, pb;\n\t"
                    "@pge add.u32 %0, %0, 1;\n\t"
                    "@prng add.u32 %1, %1, 1;\n\t"
                    "@plt add.u32 %2, %2, 1;\n\t"
                    "mov.b32 t, %4;\n\t"
                    "lop3.b32 t, t, 0xfffffff8, %9, 0xEA;\n\t"
                    "@pin st.shared.u32 [%3], t;\n\t"
                    "@pin add.u32 %3, %3, %10;\n\t"
                    "}"
                    : "+r"(n_ge[jj]), "+r"(n_rng[jj]), "+r"(n_lt[jj]), "+r"(ptr)
                    : "f"(d), "f"(lo0[jj]), "f"(hi0[jj]), "f"(lo1[jj]), "f"(hi1[jj]), "r"(tag[jj]), "n"(SAD_TV * 4));
            }
            if ((r % SAD_DRAIN) == SAD_DRAIN - 1) drain();
        }
    }
    // counters -> shared -> global
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < SAD_JC; ++jj) {
        const unsigned a = __reduce_add_sync(0xffffffffu, n_lt[jj]);
        const unsigned b = __reduce_add_sync(0xffffffffu, n_ge[jj]);
        const unsigned d = __reduce_add_sync(0xffffffffu, n_rng[jj]);
        if ((tid & 31) == 0) {
            if (a) atomicAdd(&sh_cnt[jj * 4 + 0], a);
            if (b) atomicAdd(&sh_cnt[jj * 4 + 1], b);
            if (d) atomicAdd(&sh_cnt[jj * 4 + 2], d);
        }
    }
    __syncthreads();
    if (tid < SAD_JC && j0 + tid < J) {
        SadOff* q = off + oi * J + j0 + tid;
        if (sh_cnt[tid * 4 + 0]) atomicAdd(&q->n_lt, (unsigned long long)sh_cnt[tid * 4 + 0]);
        if (sh_cnt[tid * 4 + 1]) atomicAdd(&q->n_ge, (unsigned long long)sh_cnt[tid * 4 + 1]);
        if (sh_cnt[tid * 4 + 2]) atomicAdd(&q->n_rng, (unsigned long long)sh_cnt[tid * 4 + 2]);
    }
    for (int k = tid; k < SAD_JC * 2 * SAD_B; k += SAD_TV) {
        const unsigned v = sh_hist[k];
        const int jj = k / (2 * SAD_B);
        if (v && j0 + jj < J) atomicAdd(&ghist[(size_t)(oi * J + j0) * 2 * SAD_B + k], v);
    }
}

// ------------------------------------------------------------------------------------------
// find: ranks -> histogram bins -> exact float thresholds
// ------------------------------------------------------------------------------------------
// smallest float x in [lo, hi] with sad_bin(trunc(x), lo, sc) >= b; nextafter(hi) when there is none
__device__ float sad_threshold(int b, float lo, float hi, float sc) {
    unsigned a = dcr_f2key(lo), z = dcr_f2key(hi);
    auto pred = [&](unsigned key) { return sad_bin(__uint_as_float(sad_trunc_bits(dcr_key2f(key))), lo, sc) >= b; };
    if (!pred(z)) return nextafterf(hi, INFINITY);
    while (a < z) {   // invariant: pred(z) true
        const unsigned mid = a + ((z - a) >> 1);
        if (pred(mid)) z = mid; else a = mid + 1;
    }
    return dcr_key2f(z);
}

__global__ void __launch_bounds__(256) sad_find_kernel(SadOff* __restrict__ off, const unsigned* __restrict__ ghist,
                                                       double* __restrict__ m_out) {
    __shared__ unsigned s_cum[2][SAD_B + 1];
    __shared__ unsigned s_scratch[33];
    const int o = blockIdx.x;
    SadOff* q = off + o;
    if (q->flag) return;
    for (int r = 0; r < 2; ++r) {
        const unsigned* h = ghist + ((size_t)o * 2 + r) * SAD_B;
        const unsigned a = h[2 * threadIdx.x], b = h[2 * threadIdx.x + 1];
        unsigned tot;
        const unsigned ex = dcr_block_excl_scan(a + b, s_scratch, &tot);
        s_cum[r][2 * threadIdx.x] = ex;
        s_cum[r][2 * threadIdx.x + 1] = ex + a;
        if (threadIdx.x == 0) s_cum[r][SAD_B] = tot;
        __syncthreads();
    }
    if (threadIdx.x) return;
    const long long N = (long long)(q->n_lt + q->n_ge);
    q->count = N;
    if (N == 0) { q->done = 1; m_out[o] = __longlong_as_double(0x7ff8000000000000LL); return; }
    const long long in0 = s_cum[0][SAD_B], in1 = s_cum[1][SAD_B];
    // differences below each bracket
    const long long below[2] = {(long long)q->n_lt, (long long)q->n_lt + (long long)q->n_rng - in1};
    int blo[2], bhi[2];
    for (int r = 0; r < 2; ++r) {
        const double vi = (double)(N - 1) * (r == 0 ? 0.25 : 0.75);   // numpy's virtual index, float64
        const long long klo = (long long)floor(vi);
        const long long khi = klo + 1 < N ? klo + 1 : N - 1;
        q->g[r] = vi - (double)klo;
        const long long c0 = klo - below[r], c1 = khi - below[r];
        const long long nin = r == 0 ? in0 : in1;
        if (c0 < 0 || c1 >= nin) { q->flag = 2; return; }   // a rank outside its bracket
        int b0 = 0, b1 = 0;
        {
            int lo = 0, hi = SAD_B - 1;
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((long long)s_cum[r][mid] <= c0) lo = mid; else hi = mid - 1; }
            b0 = lo;
            lo = 0; hi = SAD_B - 1;
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((long long)s_cum[r][mid] <= c1) lo = mid; else hi = mid - 1; }
            b1 = lo;
        }
        blo[r] = b0; bhi[r] = b1;
        const unsigned want = s_cum[r][b1 + 1] - s_cum[r][b0];
        if (want > SAD_CAP) { q->flag = 3; return; }        // heavy ties: too many differences in the rank's bin(s)
        q->want[r] = want;
        q->k[r] = (unsigned)(c0 - (long long)s_cum[r][b0]);
        q->k2[r] = (unsigned)(c1 - (long long)s_cum[r][b0]);
        q->ta[r] = b0 == 0 ? q->lo[r] : sad_threshold(b0, q->lo[r], q->hi[r], q->sc[r]);
        q->tb[r] = b1 == SAD_B - 1 ? nextafterf(q->hi[r], INFINITY) : sad_threshold(b1 + 1, q->lo[r], q->hi[r], q->sc[r]);
    }
    // differences in [tb[0], ta[1]): below ta[1] minus below tb[0]
    q->n_mid = (below[1] + (long long)s_cum[1][blo[1]]) - (below[0] + (long long)s_cum[0][bhi[0] + 1]);
    q->ncand[0] = q->ncand[1] = 0u;
}

// ------------------------------------------------------------------------------------------
// pass 2
// ------------------------------------------------------------------------------------------
#define SAD_P2_SMEM ((SAD_TU * SAD_TV + SAD_TU * SAD_RP) * 4)

__global__ void __launch_bounds__(SAD_TV, 2) sad_pass2_kernel(const float* __restrict__ rz, const float* __restrict__ kz, int W,
                                                              int kh, int kw, int J, int nchunk, SadOff* off,
                                                              float* __restrict__ cand, double* __restrict__ part) {
    extern __shared__ __align__(16) unsigned char sad_smem[];
    float* ks = reinterpret_cast<float*>(sad_smem);
    float* rs = ks + SAD_TU * SAD_TV;
    __shared__ double s_red[SAD_TV / 32][SAD_JC];
    const int tid = threadIdx.x;
    const int oi = blockIdx.y / nchunk, j0 = (blockIdx.y - oi * nchunk) * SAD_JC;
    const float inf = INFINITY;
    float ta0[SAD_JC], tb0[SAD_JC], ta1[SAD_JC], tb1[SAD_JC];
    double acc[SAD_JC];
    bool any = false;
#pragma unroll
    for (int jj = 0; jj < SAD_JC; ++jj) {
        ta0[jj] = tb0[jj] = ta1[jj] = tb1[jj] = inf;
        acc[jj] = 0.0;
        if (j0 + jj < J) {
            const SadOff* q = off + oi * J + j0 + jj;
            if (!q->flag && !q->done) { ta0[jj] = q->ta[0]; tb0[jj] = q->tb[0]; ta1[jj] = q->ta[1]; tb1[jj] = q->tb[1]; any = true; }
        }
    }
    SadTiles tl;
    tl.ks = ks; tl.rs = rs;
    if (any) {
        const int ntv = (kw + SAD_TV - 1) / SAD_TV, ntu = (kh + SAD_TU - 1) / SAD_TU;
        for (int t = blockIdx.x; t < ntv * ntu; t += gridDim.x) {
            const int tu = t / ntv, tv = t - tu * ntv;
            __syncthreads();
            sad_load_tile(rz, kz, W, kh, kw, oi, j0, tu * SAD_TU, tv * SAD_TV, tl);
            __syncthreads();
            float s[SAD_JC];
#pragma unroll
            for (int jj = 0; jj < SAD_JC; ++jj) s[jj] = 0.0f;
#pragma unroll
            for (int r = 0; r < SAD_TU; ++r) {
                const float kv = ks[r * SAD_TV + tid];
                const float* rr = rs + r * SAD_RP + tid;
#pragma unroll
                for (int jj = 0; jj < SAD_JC; ++jj) {
                    const float d = rr[jj] - kv;
                    const bool c0 = (d >= ta0[jj]) && (d < tb0[jj]);
                    const bool ge = d >= tb0[jj];
                    const bool mid = ge && (d < ta1[jj]);
                    const bool c1 = ge && !mid && (d < tb1[jj]);
                    if (mid) s[jj] += fabsf(d);
                    if (c0 || c1) {   // a few hundred per offset and quantile
                        SadOff* q = off + oi * J + j0 + jj;
                        const int r2 = c1 ? 1 : 0;
                        const unsigned pos = atomicAdd(&q->ncand[r2], 1u);
                        if (pos < SAD_CAP) cand[((size_t)(oi * J + j0 + jj) * 2 + r2) * SAD_CAP + pos] = d;
                    }
                }
            }
#pragma unroll
            for (int jj = 0; jj < SAD_JC; ++jj) acc[jj] += (double)s[jj];   // <= 8 float32 terms per partial
        }
    }
    // fixed-order reduction of the per-thread FP64 partials: lanes (shuffle tree), warps (serial), one slot per CTA
#pragma unroll
    for (int jj = 0; jj < SAD_JC; ++jj) {
        const double v = dcr_warp_sum(acc[jj]);
        if ((tid & 31) == 0) s_red[tid >> 5][jj] = v;
    }
    __syncthreads();
    if (tid < SAD_JC && j0 + tid < J) {
        double v = 0.0;
        for (int w = 0; w < SAD_TV / 32; ++w) v += s_red[w][tid];
        part[(size_t)(oi * J + j0 + tid) * gridDim.x + blockIdx.x] = v;
    }
}

// ------------------------------------------------------------------------------------------
// final: candidates -> P25 / P75 -> trimmed mean
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void sad_bitonic(unsigned* s, int n /* power of two */) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i + j;
                const bool up = ((i & k) == 0);
                const unsigned a = s[i], b = s[l];
                if ((a > b) == up) { s[i] = b; s[l] = a; }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(256) sad_final_kernel(SadOff* __restrict__ off, const float* __restrict__ cand,
                                                        const double* __restrict__ part, int nparts, double* __restrict__ m_out) {
    __shared__ unsigned s_key[SAD_CAP];
    __shared__ double s_sum[256];
    __shared__ unsigned s_cnt[256];
    __shared__ float s_p[2];
    const int o = blockIdx.x;
    SadOff* q = off + o;
    if (q->flag || q->done) return;
    if (q->ncand[0] != q->want[0] || q->ncand[1] != q->want[1]) {   // pass 2 disagrees with the histograms: never expected
        if (threadIdx.x == 0) q->flag = 4;
        return;
    }
    // kept differences: P25 <= d <= P75 (np.ma.masked_outside keeps the closed interval).  The candidates of P25 all lie below
    // P75 and vice versa, so each list is tested against its own quantile; summed out of the SORTED list so that the result
    // does not depend on the order in which pass 2 appended the candidates.
    double s = 0.0;
    unsigned cnt = 0;
    for (int r = 0; r < 2; ++r) {
        const unsigned nc = q->ncand[r];
        int np2 = 32;
        while (np2 < (int)nc) np2 <<= 1;
        const float* c = cand + ((size_t)o * 2 + r) * SAD_CAP;
        for (int k = threadIdx.x; k < np2; k += 256) s_key[k] = k < (int)nc ? dcr_f2key(c[k] + 0.0f) : 0xFFFFFFFFu;
        __syncthreads();
        sad_bitonic(s_key, np2);
        if (threadIdx.x == 0) s_p[r] = dcr_lerp32(dcr_key2f(s_key[q->k[r]]), dcr_key2f(s_key[q->k2[r]]), q->g[r]);
        __syncthreads();
        const float p = s_p[r];
        for (unsigned k = threadIdx.x; k < nc; k += 256) {
            const float d = dcr_key2f(s_key[k]);
            if (r == 0 ? (d >= p) : (d <= p)) { s += (double)fabsf(d); ++cnt; }
        }
        __syncthreads();
    }
    s_sum[threadIdx.x] = s;
    s_cnt[threadIdx.x] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        unsigned long long n = 0;
        for (int k = 0; k < 256; ++k) { tot += s_sum[k]; n += s_cnt[k]; }
        double mid = 0.0;
        for (int k = 0; k < nparts; ++k) mid += part[(size_t)o * nparts + k];
        long long nm = q->n_mid;
        const double den = (double)(nm + (long long)n);
        m_out[o] = den > 0 ? (mid + tot) / den : __longlong_as_double(0x7ff8000000000000LL);
    }
}

// ------------------------------------------------------------------------------------------
// exact per-offset path: materialised diff + two K3 selections + trimmed sum
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sad_diff_kernel(const float* __restrict__ ref, const unsigned char* __restrict__ rm,
                                                       const float* __restrict__ src, const unsigned char* __restrict__ smk,
                                                       long long stride, int oi, int oj, int pad, int kh, int kw,
                                                       float* __restrict__ diff, unsigned char* __restrict__ dmask) {
    const long long n = (long long)kh * kw;
    for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < n; k += (long long)gridDim.x * 256) {
        const int u = (int)(k / kw), v = (int)(k - (long long)u * kw);
        const long long ri = (long long)(oi + u) * stride + oj + v;
        const long long si = (long long)(pad + u) * stride + pad + v;
        const bool masked = (rm && rm[ri]) || (smk && smk[si]);
        diff[k] = masked ? 0.0f : (ref[ri] - src[si]);
        dmask[k] = masked ? 1 : 0;
    }
}

struct SadPartial { unsigned long long n; double s; };

// sum of |d| over unmasked d with lo <= d <= hi
__global__ void __launch_bounds__(256) sad_sum_kernel(const float* __restrict__ diff, const unsigned char* __restrict__ dmask,
                                                      long long n, const float* __restrict__ qlo, const float* __restrict__ qhi,
                                                      SadPartial* __restrict__ part) {
    __shared__ SadPartial sp[8];
    const float lo = *qlo, hi = *qhi;
    unsigned long long cnt = 0;
    double s = 0;
    dcr_stream1(diff, dmask, n, [&](float d, unsigned mb, long long) {
        if (mb) return;
        if (d < lo || d > hi) return;
        ++cnt;
        s += (double)fabsf(d);
    });
    cnt = dcr_warp_sum_u64(cnt); s = dcr_warp_sum(s);
    if ((threadIdx.x & 31) == 0) { SadPartial q; q.n = cnt; q.s = s; sp[threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        SadPartial q = sp[0];
        for (int k = 1; k < 8; ++k) { q.n += sp[k].n; q.s += sp[k].s; }
        part[blockIdx.x] = q;
    }
}

__global__ void sad_sum_final_kernel(const SadPartial* __restrict__ part, int nparts, double* m_out) {
    if (threadIdx.x) return;
    unsigned long long n = 0; double s = 0;
    for (int k = 0; k < nparts; ++k) { n += part[k].n; s += part[k].s; }
    *m_out = n ? s / (double)n : __longlong_as_double(0x7ff8000000000000LL);
}

struct SadWs {
    float* rz; float* kz;
    SadOff* off;
    unsigned* ghist;
    float* cand;
    double* part;
    double* dm;
    // per-offset path
    float* diff; unsigned char* dmask;
    SelState* sst;
    unsigned* selhist;
    SadPartial* spart;
    int* flags_host_mirror;   // unused on the device
};

static int sad_p2_grid_x(int J) {
    const int nchunk = (J + SAD_JC - 1) / SAD_JC;
    int gx = (dcr_num_sms() * 6 + J * nchunk - 1) / (J * nchunk);
    return gx < 1 ? 1 : gx;
}

static size_t sad_ws_bytes(int h, int w, int pad) {
    const long long kh = h - 2 * pad, kw = w - 2 * pad;
    const long long n = kh * kw;
    const long long JJ = (long long)(2 * pad + 1) * (2 * pad + 1);
    return dcr_al256((size_t)h * w * 4) + dcr_al256((size_t)n * 4) + dcr_al256((size_t)JJ * sizeof(SadOff)) +
           dcr_al256((size_t)JJ * 2 * SAD_B * 4) + dcr_al256((size_t)JJ * 2 * SAD_CAP * 4) +
           dcr_al256((size_t)JJ * sad_p2_grid_x(2 * pad + 1) * 8) + dcr_al256((size_t)JJ * 8) +
           dcr_al256((size_t)n * 4) + dcr_al256((size_t)n) + dcr_al256(4 * sizeof(SelState)) +
           dcr_al256(DCR_SEL_HIST_WORDS * 4) + dcr_al256(4096 * sizeof(SadPartial)) + 8192;
}
extern "C" size_t dcr_sad_workspace_bytes(int h, int w, int pad) {
    if (h - 2 * pad <= 0 || w - 2 * pad <= 0 || pad < 1) return 0;
    return sad_ws_bytes(h, w, pad);
}

static int sad_carve(void* workspace, size_t bytes, int h, int w, int pad, SadWs* s) {
    const long long kh = h - 2 * pad, kw = w - 2 * pad, n = kh * kw;
    const long long JJ = (long long)(2 * pad + 1) * (2 * pad + 1);
    DcrBump ws(workspace, bytes);
    s->rz = ws.take<float>((size_t)h * w);
    s->kz = ws.take<float>(n);
    s->off = ws.take<SadOff>(JJ);
    s->ghist = ws.take<unsigned>((size_t)JJ * 2 * SAD_B);
    s->cand = ws.take<float>((size_t)JJ * 2 * SAD_CAP);
    s->part = ws.take<double>((size_t)JJ * sad_p2_grid_x(2 * pad + 1));
    s->dm = ws.take<double>(JJ);
    s->diff = ws.take<float>(n);
    s->dmask = ws.take<unsigned char>(n);
    s->sst = ws.take<SelState>(4);
    s->selhist = ws.take<unsigned>(DCR_SEL_HIST_WORDS);
    s->spart = ws.take<SadPartial>(4096);
    if (!s->rz || !s->kz || !s->off || !s->ghist || !s->cand || !s->part || !s->dm || !s->diff || !s->dmask || !s->sst ||
        !s->selhist || !s->spart) {
        dcr_set_error("workspace too small");
        return -1;
    }
    return 0;
}

// one offset through the exact radix selections (3 histogram passes each): enqueue only
static int sad_offset_exact_enqueue(const float* ref, const uint8_t* ref_mask, const float* src, const uint8_t* src_mask,
                                    int64_t stride, int pad, int kh, int kw, int i, int j, const SadWs& w, double* dm_o,
                                    cudaStream_t s) {
    const long long n = (long long)kh * kw;
    const int g = dcr_num_sms() * 8;
    sad_diff_kernel<<<g, 256, 0, s>>>(ref, ref_mask, src, src_mask, stride, i, j, pad, kh, kw, w.diff, w.dmask);
    SelView v;
    memset(&v, 0, sizeof(v));
    v.x = w.diff; v.m = w.dmask; v.reject = 1; v.n = n;
    int rc = dcr_sel_enqueue(v, 25.0, w.sst, w.selhist, s);
    if (!rc) rc = dcr_sel_enqueue(v, 75.0, w.sst + 1, w.selhist, s);
    if (rc) return rc;
    sad_sum_kernel<<<g, 256, 0, s>>>(w.diff, w.dmask, n, &w.sst[0].result, &w.sst[1].result, w.spart);
    sad_sum_final_kernel<<<1, 32, 0, s>>>(w.spart, g, dm_o);
    DCR_LAUNCH_OK("sad per-offset kernels");
    return 0;
}

extern "C" int dcr_sad_search_ex(const float* ref, const uint8_t* ref_mask, const float* src, const uint8_t* src_mask, int h,
                                 int w, int64_t stride, int pad, int flags, double* m_out, int* n_per_offset, float* stage_ms,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    DCR_ARG(ref && src && m_out && workspace, "null pointer");
    DCR_ARG(pad >= 1 && pad <= 127, "1 <= pad <= 127");
    const int kh = h - 2 * pad, kw = w - 2 * pad;
    DCR_ARG(kh > 0 && kw > 0, "raster smaller than the search window");
    DCR_ARG(workspace_bytes >= sad_ws_bytes(h, w, pad), "workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    const int J = 2 * pad + 1, JJ = J * J;
    SadWs ws;
    if (sad_carve(workspace, workspace_bytes, h, w, pad, &ws)) return -1;
    if (n_per_offset) *n_per_offset = 0;
    if (flags & DCR_SAD_PER_OFFSET) {
        for (int i = 0; i < J; ++i)
            for (int j = 0; j < J; ++j) {
                int rc = sad_offset_exact_enqueue(ref, ref_mask, src, src_mask, stride, pad, kh, kw, i, j, ws, ws.dm + i * J + j, s);
                if (rc) return rc;
            }
        DCR_CUDA_OK(cudaMemcpyAsync(m_out, ws.dm, (size_t)JJ * sizeof(double), cudaMemcpyDeviceToHost, s));
        DCR_CUDA_OK(cudaStreamSynchronize(s));
        if (n_per_offset) *n_per_offset = JJ;
        return 0;
    }
    // ---- batched path
    cudaEvent_t ev[7] = {nullptr};
    if (stage_ms)
        for (int k = 0; k < 7; ++k) DCR_CUDA_OK(cudaEventCreate(&ev[k]));
#define SAD_MARK(k) do { if (stage_ms) DCR_CUDA_OK(cudaEventRecord(ev[k], s)); } while (0)
    const int g = dcr_num_sms() * 8;
    const int nchunk = (J + SAD_JC - 1) / SAD_JC;
    int gx = sad_p2_grid_x(J);   // CTAs per (offset row, offset chunk): enough to fill the machine, never more than tiles
    {
        const long long ntiles = (long long)((kw + SAD_TV - 1) / SAD_TV) * ((kh + SAD_TU - 1) / SAD_TU);
        if (gx > ntiles) gx = (int)ntiles;
    }
    DCR_CUDA_OK(cudaFuncSetAttribute(sad_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SAD_S * 4));
    DCR_CUDA_OK(cudaFuncSetAttribute(sad_pass1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SAD_P1_SMEM));
    SAD_MARK(0);
    sad_prep_kernel<<<g, 256, 0, s>>>(ref, ref_mask, stride, 0, 0, h, w, ws.rz);
    sad_prep_kernel<<<g, 256, 0, s>>>(src, src_mask, stride, pad, pad, kh, kw, ws.kz);
    DCR_CUDA_OK(cudaMemsetAsync(ws.ghist, 0, (size_t)JJ * 2 * SAD_B * 4, s));
    SAD_MARK(1);
    sad_sample_kernel<<<JJ, 1024, SAD_S * 4, s>>>(ws.rz, ws.kz, w, kh, kw, J, ws.off);
    SAD_MARK(2);
    sad_pass1_kernel<<<dim3(gx, J * nchunk), SAD_TV, SAD_P1_SMEM, s>>>(ws.rz, ws.kz, w, kh, kw, J, nchunk, ws.off, ws.ghist);
    SAD_MARK(3);
    sad_find_kernel<<<JJ, 256, 0, s>>>(ws.off, ws.ghist, ws.dm);
    SAD_MARK(4);
    sad_pass2_kernel<<<dim3(gx, J * nchunk), SAD_TV, SAD_P2_SMEM, s>>>(ws.rz, ws.kz, w, kh, kw, J, nchunk, ws.off, ws.cand, ws.part);
    SAD_MARK(5);
    sad_final_kernel<<<JJ, 256, 0, s>>>(ws.off, ws.cand, ws.part, gx, ws.dm);
    SAD_MARK(6);
#undef SAD_MARK
    DCR_LAUNCH_OK("sad batched kernels");
    SadOff* hoff = (SadOff*)malloc((size_t)JJ * sizeof(SadOff));
    if (!hoff) { dcr_set_error("out of host memory"); return -1; }
    cudaError_t e = cudaMemcpyAsync(hoff, ws.off, (size_t)JJ * sizeof(SadOff), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { free(hoff); dcr_set_error("sad copy failed: %s", cudaGetErrorString(e)); return (int)e; }
    if (stage_ms) {
        for (int k = 0; k < 6; ++k) cudaEventElapsedTime(&stage_ms[k], ev[k], ev[k + 1]);
        for (int k = 0; k < 7; ++k) cudaEventDestroy(ev[k]);
    }
    // offsets the batched path declined (heavy ties, degenerate brackets, a missed rank): exact per-offset path
    int nfb = 0;
    for (int o = 0; o < JJ; ++o) {
        if (!hoff[o].flag) continue;
        ++nfb;
        if (getenv("DCR_DEBUG")) fprintf(stderr, "[dcr] sad offset %d: batched path declined (flag %d)\n", o, hoff[o].flag);
        int rc = sad_offset_exact_enqueue(ref, ref_mask, src, src_mask, stride, pad, kh, kw, o / J, o % J, ws, ws.dm + o, s);
        if (rc) { free(hoff); return rc; }
    }
    free(hoff);
    if (n_per_offset) *n_per_offset = nfb;
    DCR_CUDA_OK(cudaMemcpyAsync(m_out, ws.dm, (size_t)JJ * sizeof(double), cudaMemcpyDeviceToHost, s));
    DCR_CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

extern "C" int dcr_sad_search(const float* ref, const uint8_t* ref_mask, const float* src, const uint8_t* src_mask, int h, int w,
                              int64_t stride, int pad, double* m_out, void* workspace, size_t workspace_bytes, void* stream) {
    return dcr_sad_search_ex(ref, ref_mask, src, src_mask, h, w, stride, pad, 0, m_out, nullptr, nullptr, workspace,
                             workspace_bytes, stream);
}
