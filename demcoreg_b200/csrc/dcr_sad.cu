// K7b SAD offset search, batched over the whole (2p+1)^2 shift window.
//
// Replaces reference coreglib.compute_offset_sad (coreglib.py:65-115): for every offset (i, j) of the window,
//   diff = dem1[i:i+kh, j:j+kw] - dem2[p:-p, p:-p]            (float32, masked)
//   diff = masked_outside(diff, *malib.calcperc(diff, (25, 75)))   (exact percentiles, linear interpolation)
//   m[i, j] = abs(diff).sum() / diff.count()
// The reference runs the (2p+1)^2 offsets one after the other (two exact quantiles of kh*kw values each:
// 1.4 s per offset at 4096^2).  Here ALL offsets are processed by the same few passes over the two rasters:
//
//   prep     masked pixels -> NaN (a NaN difference takes part in nothing), kernel = src[p:-p, p:-p] made contiguous,
//            validity bit planes of both
//   count    per offset: number of unmasked differences = popcount of the AND of the shifted bit planes
//   sample   one CTA per offset: 32768 stratified samples of diff -> a bracket [lo, hi] around each of the two
//            order statistics (sample ranks -+ 4 sigma), localised by three rounds of a 1024-bin histogram
//   pass 1   every difference of every offset, out of shared-memory tiles (thread = pixel column; 8 consecutive column
//            offsets per CTA, two at a time): exact counts at / above the lower edge of each bracket; the ~4 % inside a
//            bracket are parked in thread-private shared-memory slots and drained into a 256-bin histogram of their
//            bracket (shared-memory atomics)
//   find     per offset: exact integer ranks of P25 / P75 (float64 virtual index, as numpy) -> the histogram bins
//            that hold them, VERIFIED to lie inside the brackets; the bin edges as exact float thresholds (bisection
//            over the monotone bin function)
//   pass 2   every difference again: sum |diff| strictly between the two candidate intervals (FP32 per 8 rows,
//            FP64 across), and the few hundred differences inside the candidate bins appended per offset
//   final    per offset: sort the candidates, P25 / P75 by numpy's lerp, trimmed sum and count -> m
//
// An offset whose verification fails (heavy ties, degenerate or overlapping brackets, candidate overflow) is
// redone by the exact per-offset path (materialised diff + two K3 selections), which is also what
// DCR_SAD_PER_OFFSET runs for every offset (tests compare the two).
//
// Both passes run out of shared memory and are bound by the half-rate ALU pipe (compares, predicate logic), not by HBM
// or L2: the classification step of a difference is hand-written PTX -- 4 FSETP + 1 PLOP3 + 1 VIADD on the ALU pipe, the
// subtraction and the two counter updates (FLOAT counters: predicated FADD) on the FMA pipe, 1 LDS + 1 predicated STS.
// Per-thread state is kept small (two offsets at a time, bounds re-read from shared memory per offset pair) so that
// 4 CTAs of 256 threads fit an SM and hide the tile loads.
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include "dcr_common.cuh"
#include "dcr_internal.cuh"

#define SAD_B 256        // histogram bins per bracket
#define SAD_JC 8         // consecutive column offsets per CTA
#define SAD_TU 8         // tile rows
#define SAD_TV 256       // tile columns = threads per CTA
#define SAD_PRIV (2 * SAD_TU)   // private parking slots per thread: one offset pair x SAD_TU rows can never overflow them
#define SAD_CAP 4096     // candidates per (offset, quantile)
#define SAD_S 32768      // samples per offset
#define SAD_QNAN __int_as_float(0x7fc00000)

struct SadOff {
    // sample kernel: both brackets in terms of e = fl(d - c), c = centre of the gap between them:
    //   bracket of P25 = {-b <= e <= -a}, bracket of P75 = {a <= e <= b}   (supersets of the sample brackets)
    float c, a, b, sc;           // sc = SAD_B / (b - a): bin scale of both brackets
    float lo[2], hi[2];          // the sample brackets themselves (closed, in d), informational
    int flag;                    // != 0: batched path not applicable / failed verification -> per-offset exact path
    int done;                    // m already final (no unmasked difference)
    // count + pass 1
    unsigned long long n_c0, n_gap;    // e >= -b ; |e| < a
    long long count;             // unmasked differences (popcount of the validity planes)
    // find (thresholds in e)
    float ta[2], tb[2];          // candidate intervals [ta, tb) of P25 / P75; the untouched middle is [tb[0], ta[1])
    long long n_mid;             // differences inside the middle
    unsigned k[2], k2[2];        // ranks of the two neighbouring order statistics inside the sorted candidates
    double g[2];                 // interpolation weights
    unsigned want[2];            // candidates pass 2 must find
    unsigned ncand[2];           // candidates pass 2 found
};

// bin of a difference inside its bracket, from t = e - (lower edge of the bracket in e): monotone non-decreasing in d
// (two subtractions, a multiplication, a conversion and clamping all are; the file is compiled without FMA contraction, so
// every kernel rounds identically)
__device__ __forceinline__ int sad_bin(float e, float lo_e, float sc) {
    const float t = (e - lo_e) * sc;
    int b = (int)t;
    return b < 0 ? 0 : (b > SAD_B - 1 ? SAD_B - 1 : b);
}

// ------------------------------------------------------------------------------------------
// prep: NaN-filled copies + validity bit planes; count: unmasked differences per offset
// ------------------------------------------------------------------------------------------
// one warp per 32 consecutive columns of a row; bits[row * wp + word], bit l = column 32 * word + l is valid
__global__ void __launch_bounds__(256) sad_prep_kernel(const float* __restrict__ x, const unsigned char* __restrict__ m,
                                                       long long stride, int r0, int c0, int h, int w,
                                                       float* __restrict__ out, unsigned* __restrict__ bits, int wp) {
    const int lane = threadIdx.x & 31;
    const int nwords = (w + 31) >> 5;
    const long long ntask = (long long)h * nwords;
    for (long long t = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); t < ntask; t += (long long)gridDim.x * 8) {
        const int i = (int)(t / nwords), wd = (int)(t - (long long)i * nwords);
        const int j = wd * 32 + lane;
        bool ok = false;
        if (j < w) {
            const long long s = (long long)(r0 + i) * stride + c0 + j;
            float v = x[s];
            if (m && m[s]) v = SAD_QNAN;
            ok = v == v;
            out[(long long)i * w + j] = v;
        }
        const unsigned b = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) bits[(long long)i * wp + wd] = b;
    }
}

// N(i, j) = sum over kernel pixels (u, v) of valid_k[u][v] & valid_r[u + i][v + j].  CTA = (row segment, offset row i,
// chunk of 32 offset columns); lane = offset column j: the three words of a step are read once (broadcast) and shifted by
// the lane's own j; per-lane counts leave through one atomic per lane and CTA.
#define SAD_CNT_SEG 8
__global__ void __launch_bounds__(256) sad_count_kernel(const unsigned* __restrict__ rb, int rwp, const unsigned* __restrict__ kb,
                                                        int kwp, int kh, int J, SadOff* __restrict__ off) {
    __shared__ unsigned long long s_n[8][32];
    const int oi = blockIdx.y, jb = blockIdx.z * 32;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int oj = jb + lane;
    unsigned long long n = 0;
    for (int u = blockIdx.x * 8 + wrp; u < kh; u += SAD_CNT_SEG * 8) {
        const unsigned* kr = kb + (long long)u * kwp;
        const unsigned* rr = rb + (long long)(u + oi) * rwp + (jb >> 5);
        unsigned cnt = 0;
        for (int wd = 0; wd < kwp; wd += 4) {
            unsigned kk[4], lo[4], hi[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool in = wd + q < kwp;
                kk[q] = in ? __ldg(kr + wd + q) : 0u;
                lo[q] = in ? __ldg(rr + wd + q) : 0u;
                hi[q] = in ? __ldg(rr + wd + q + 1) : 0u;   // (the planes carry one zero word of padding)
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) cnt += __popc(kk[q] & __funnelshift_r(lo[q], hi[q], lane));
        }
        n += cnt;
    }
    s_n[wrp][lane] = n;
    __syncthreads();
    if (wrp == 0 && oj < J) {
        unsigned long long t = 0;
        for (int k = 0; k < 8; ++k) t += s_n[k][lane];
        if (t) atomicAdd(reinterpret_cast<unsigned long long*>(&off[oi * J + oj].count), t);
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
            const double dd = 4.0 * sqrt((double)m * f * (1.0 - f)) + 4.0;   // a miss (6e-5 per edge) costs one per-offset redo
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
        SadOff q = off[o];        // keeps the count of the count kernel
        q.n_c0 = q.n_gap = 0ull;
        q.n_mid = 0;
        q.ncand[0] = q.ncand[1] = 0u;
        q.done = 0;
        q.flag = open ? 1 : 0;
        if (!open) {
            for (int r = 0; r < 2; ++r) { q.lo[r] = dcr_key2f(klo[r]); q.hi[r] = dcr_key2f(khi[r]); }
            // e = fl(d - c) with c in the middle of the gap; a / b from the SAME float expression, so that (fl(. - c) being
            // monotone) lo0 <= d <= hi0 implies -b <= e <= -a and lo1 <= d <= hi1 implies a <= e <= b
            q.c = 0.5f * q.hi[0] + 0.5f * q.lo[1];
            q.a = fminf(-(q.hi[0] - q.c), q.lo[1] - q.c);
            q.b = fmaxf(-(q.lo[0] - q.c), q.hi[1] - q.c);
            const float w = q.b - q.a;
            q.sc = (w > 0.0f && !isinf(w)) ? (float)SAD_B / w : 0.0f;
            const float big = fmaxf(fabsf(q.hi[0]), fabsf(q.lo[1]));
            // the brackets must be well separated and finite
            if (!(q.a > 1e-5f * big) || !(q.b > q.a) || isinf(q.b) || isnan(q.c) || !(q.sc > 0.0f)) q.flag = 1;
        }
        off[o] = q;
    }
}

// ------------------------------------------------------------------------------------------
// the tile loop shared by pass 1 and pass 2: thread = pixel column, SAD_JC consecutive column offsets, two at a time
// ------------------------------------------------------------------------------------------
#define SAD_RP (SAD_TV + SAD_JC)

// k tile [SAD_TU][SAD_TV] of the kernel raster, r tile [SAD_TU][SAD_RP] of the reference rows u0 + oi .., columns v0 + j0 ..
__device__ __forceinline__ void sad_load_tile(const float* __restrict__ rz, const float* __restrict__ kz, int W, int kh, int kw,
                                              int oi, int j0, int u0, int v0, float* __restrict__ ks, float* __restrict__ rs) {
    const int tid = threadIdx.x;
    if (((kw | W) & 3) == 0 && u0 + SAD_TU <= kh && v0 + SAD_TV <= kw && v0 + j0 + SAD_RP <= W) {
        // interior tile of 16-byte aligned rows: 16-byte loads (2 + 2..3 per thread instead of 17 scalar ones)
        for (int q = tid; q < SAD_TU * (SAD_TV / 4); q += SAD_TV) {
            const int r = q / (SAD_TV / 4), c4 = q - r * (SAD_TV / 4);
            reinterpret_cast<float4*>(ks)[q] = __ldg(reinterpret_cast<const float4*>(kz + (long long)(u0 + r) * kw + v0) + c4);
        }
        for (int q = tid; q < SAD_TU * (SAD_RP / 4); q += SAD_TV) {
            const int r = q / (SAD_RP / 4), c4 = q - r * (SAD_RP / 4);
            reinterpret_cast<float4*>(rs)[q] = __ldg(reinterpret_cast<const float4*>(rz + (long long)(u0 + r + oi) * W + v0 + j0) + c4);
        }
        return;
    }
    const bool cok = v0 + tid < kw;
#pragma unroll
    for (int r = 0; r < SAD_TU; ++r) {
        const bool rok = u0 + r < kh;
        ks[r * SAD_TV + tid] = (rok && cok) ? __ldg(kz + (long long)(u0 + r) * kw + v0 + tid) : SAD_QNAN;
        const float* rrow = rz + (long long)(u0 + r + oi) * W + v0 + j0;
        rs[r * SAD_RP + tid] = (rok && v0 + j0 + tid < W) ? __ldg(rrow + tid) : SAD_QNAN;
        if (tid < SAD_JC) rs[r * SAD_RP + SAD_TV + tid] = (rok && v0 + j0 + SAD_TV + tid < W) ? __ldg(rrow + SAD_TV + tid) : SAD_QNAN;
    }
}

// ------------------------------------------------------------------------------------------
// pass 1
// ------------------------------------------------------------------------------------------
#define SAD_P1_SMEM ((SAD_TU * SAD_TV + SAD_TU * SAD_RP + SAD_JC * 2 * SAD_B + SAD_PRIV * SAD_TV + SAD_JC * 4 + SAD_JC * 2) * 4)

// One difference d: e = d - c; counters (float, exact below 2^24: predicated FADD = FMA pipe) of e >= -b and of |e| < a
// (the gap between the brackets); e parked iff a <= |e| <= b: ONE predicated store + pointer bump.  Three compares, three
// predicates.  NaN (a masked pixel) fails every comparison.
#define SAD_P1_STEP(D, C0, CG, PTR, CC, NB, A, B, STEP)                                                             \
    asm volatile(                                                                                                    \
        "{\n\t"                                                                                                      \
        ".reg .pred p0, pg, pin;\n\t"                                                                                \
        ".reg .f32 e, ae;\n\t"                                                                                       \
        "sub.f32 e, %3, %4;\n\t"                                                                                     \
        "abs.f32 ae, e;\n\t"                                                                                         \
        "setp.ge.f32 p0, e, %5;\n\t"                                                                                 \
        "setp.lt.f32 pg, ae, %6;\n\t"                                                                                \
        "setp.le.and.f32 pin, ae, %7, !pg;\n\t"                                                                      \
        "@p0 add.f32 %0, %0, 0f3F800000;\n\t"                                                                        \
        "@pg add.f32 %1, %1, 0f3F800000;\n\t"                                                                        \
        "@pin st.shared.f32 [%2], e;\n\t"                                                                            \
        "@pin add.s32 %2, %2, %8;\n\t"                                                                               \
        "}"                                                                                                          \
        : "+f"(C0), "+f"(CG), "+r"(PTR)                                                                              \
        : "f"(D), "f"(CC), "f"(NB), "f"(A), "f"(B), "n"(STEP))

__global__ void __launch_bounds__(SAD_TV, 4) sad_pass1_kernel(const float* __restrict__ rz, const float* __restrict__ kz, int W,
                                                              int kh, int kw, int J, int nchunk, SadOff* __restrict__ off,
                                                              unsigned* __restrict__ ghist) {
    extern __shared__ __align__(16) unsigned char sad_smem[];
    float* ks = reinterpret_cast<float*>(sad_smem);
    float* rs = ks + SAD_TU * SAD_TV;
    unsigned* sh_hist = reinterpret_cast<unsigned*>(rs + SAD_TU * SAD_RP);   // [SAD_JC][2][SAD_B]
    float* sh_priv = reinterpret_cast<float*>(sh_hist + SAD_JC * 2 * SAD_B); // slot k of thread t at [k * SAD_TV + t]
    float* sh_b = sh_priv + SAD_PRIV * SAD_TV;                               // [SAD_JC][4]: c a b sc
    unsigned* sh_cnt = reinterpret_cast<unsigned*>(sh_b + SAD_JC * 4);       // [SAD_JC][2]
    const int tid = threadIdx.x;
    const int oi = blockIdx.y / nchunk, j0 = (blockIdx.y - oi * nchunk) * SAD_JC;
    const int nj = min(SAD_JC, J - j0);
    const float inf = INFINITY;
    for (int k = tid; k < SAD_JC * 2 * SAD_B; k += SAD_TV) sh_hist[k] = 0u;
    if (tid < SAD_JC) {
        float b[4] = {0.f, inf, inf, 0.f};   // absent / flagged offset (a = b = inf): nothing is parked
        if (tid < nj) {
            const SadOff* q = off + oi * J + j0 + tid;
            if (!q->flag) { b[0] = q->c; b[1] = q->a; b[2] = q->b; b[3] = q->sc; }
        }
        for (int k = 0; k < 4; ++k) sh_b[tid * 4 + k] = b[k];
        sh_cnt[tid * 2 + 0] = sh_cnt[tid * 2 + 1] = 0u;
    }
    float c0[SAD_JC], cg[SAD_JC];
#pragma unroll
    for (int jj = 0; jj < SAD_JC; ++jj) c0[jj] = cg[jj] = 0.0f;
    // parking slots of one offset pair: the first offset fills the thread's slots upwards from slot 0, the second one
    // downwards from the last slot -- no tag is needed to tell them apart and SAD_TU hits each always fit
    const unsigned mineA = (unsigned)__cvta_generic_to_shared(sh_priv + tid);
    const unsigned mineB = mineA + (SAD_PRIV - 1) * SAD_TV * 4;
    const int ntv = (kw + SAD_TV - 1) / SAD_TV, ntu = (kh + SAD_TU - 1) / SAD_TU;
    for (int t = blockIdx.x; t < ntv * ntu; t += gridDim.x) {
        const int tu = t / ntv, tv = t - tu * ntv;
        __syncthreads();   // the previous tile is no longer read (also orders the set-up above on the first trip)
        sad_load_tile(rz, kz, W, kh, kw, oi, j0, tu * SAD_TU, tv * SAD_TV, ks, rs);
        __syncthreads();
        float kv[SAD_TU];
#pragma unroll
        for (int r = 0; r < SAD_TU; ++r) kv[r] = ks[r * SAD_TV + tid];
#pragma unroll
        for (int jp = 0; jp < SAD_JC; jp += 2) {
            if (jp >= nj) break;   // CTA-uniform
            const float4 ba = *reinterpret_cast<const float4*>(sh_b + jp * 4);          // c a b sc (broadcast loads)
            const float4 bb = *reinterpret_cast<const float4*>(sh_b + (jp + 1) * 4);
            const float nba = -ba.z, nbb = -bb.z;
            unsigned pa = mineA, pb = mineB;
            const float* rr = rs + tid + jp;
#pragma unroll
            for (int r = 0; r < SAD_TU; ++r) {
                const float da = rr[r * SAD_RP] - kv[r];
                const float db = rr[r * SAD_RP + 1] - kv[r];
                SAD_P1_STEP(da, c0[jp], cg[jp], pa, ba.x, nba, ba.y, ba.z, SAD_TV * 4);
                SAD_P1_STEP(db, c0[jp + 1], cg[jp + 1], pb, bb.x, nbb, bb.y, bb.z, -SAD_TV * 4);
            }
            // drain the parked e into the bracket histograms: ONE value per trip (first the slots of the first offset, then
            // those of the second), warp-uniform trip count; e < 0 -> bracket of P25
            const unsigned na = (pa - mineA) / (SAD_TV * 4), nb = (mineB - pb) / (SAD_TV * 4);
            const unsigned nmax = __reduce_max_sync(0xffffffffu, na + nb);
            for (unsigned k = 0; k < nmax; ++k) {
                if (k < na + nb) {
                    const bool second = k >= na;
                    float e;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(e) : "r"(second ? mineB - (k - na) * (SAD_TV * 4) : mineA + k * (SAD_TV * 4)));
                    const float a_ = second ? bb.y : ba.y, nb_ = second ? nbb : nba, sc_ = second ? bb.w : ba.w;
                    const int q = e > 0.0f ? 1 : 0;
                    atomicAdd(&sh_hist[((jp + (second ? 1 : 0)) * 2 + q) * SAD_B + sad_bin(e, q ? a_ : nb_, sc_)], 1u);
                }
            }
        }
    }
    // counters -> shared -> global
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < SAD_JC; ++jj) {
        const unsigned a = __reduce_add_sync(0xffffffffu, (unsigned)c0[jj]);
        const unsigned b = __reduce_add_sync(0xffffffffu, (unsigned)cg[jj]);
        if ((tid & 31) == 0) {
            if (a) atomicAdd(&sh_cnt[jj * 2 + 0], a);
            if (b) atomicAdd(&sh_cnt[jj * 2 + 1], b);
        }
    }
    __syncthreads();
    if (tid < nj) {
        SadOff* q = off + oi * J + j0 + tid;
        if (sh_cnt[tid * 2 + 0]) atomicAdd(&q->n_c0, (unsigned long long)sh_cnt[tid * 2 + 0]);
        if (sh_cnt[tid * 2 + 1]) atomicAdd(&q->n_gap, (unsigned long long)sh_cnt[tid * 2 + 1]);
    }
    for (int k = tid; k < nj * 2 * SAD_B; k += SAD_TV) {
        const unsigned v = sh_hist[k];
        if (v) atomicAdd(&ghist[(size_t)(oi * J + j0) * 2 * SAD_B + k], v);
    }
}

// ------------------------------------------------------------------------------------------
// find: ranks -> histogram bins -> exact float thresholds
// ------------------------------------------------------------------------------------------
// smallest float e in [lo, hi] with sad_bin(e, lo, sc) >= b; nextafter(hi) when there is none
__device__ float sad_threshold(int b, float lo, float hi, float sc) {
    unsigned a = dcr_f2key(lo), z = dcr_f2key(hi);
    auto pred = [&](unsigned key) { return sad_bin(dcr_key2f(key), lo, sc) >= b; };
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
        const unsigned a = h[threadIdx.x];   // SAD_B == blockDim.x
        unsigned tot;
        const unsigned ex = dcr_block_excl_scan(a, s_scratch, &tot);
        s_cum[r][threadIdx.x] = ex;
        if (threadIdx.x == 0) s_cum[r][SAD_B] = tot;
        __syncthreads();
    }
    if (threadIdx.x) return;
    const long long N = q->count;
    if (N == 0) { q->done = 1; m_out[o] = __longlong_as_double(0x7ff8000000000000LL); return; }
    // differences below each bracket: e < -b ; e < a (= those + the P25 bracket + the gap)
    const long long below0 = N - (long long)q->n_c0;
    const long long below[2] = {below0, below0 + (long long)s_cum[0][SAD_B] + (long long)q->n_gap};
    if (below0 < 0) { q->flag = 6; return; }
    const float elo[2] = {-q->b, q->a}, ehi[2] = {-q->a, q->b};   // the brackets in e
    int blo[2], bhi[2];
    for (int r = 0; r < 2; ++r) {
        const double vi = (double)(N - 1) * (r == 0 ? 0.25 : 0.75);   // numpy's virtual index, float64
        const long long klo = (long long)floor(vi);
        const long long khi = klo + 1 < N ? klo + 1 : N - 1;
        q->g[r] = vi - (double)klo;
        const long long c0 = klo - below[r], c1 = khi - below[r];
        const long long nin = (long long)s_cum[r][SAD_B];
        if (c0 < 0 || c1 >= nin) { q->flag = 2; return; }   // a rank outside its bracket
        int b0, b1;
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
        q->ta[r] = b0 == 0 ? elo[r] : sad_threshold(b0, elo[r], ehi[r], q->sc);
        q->tb[r] = b1 == SAD_B - 1 ? nextafterf(ehi[r], INFINITY) : sad_threshold(b1 + 1, elo[r], ehi[r], q->sc);
    }
    // differences in [tb[0], ta[1]): below ta[1] minus below tb[0]
    q->n_mid = (below[1] + (long long)s_cum[1][blo[1]]) - (below[0] + (long long)s_cum[0][bhi[0] + 1]);
    q->ncand[0] = q->ncand[1] = 0u;
}

// ------------------------------------------------------------------------------------------
// pass 2
// ------------------------------------------------------------------------------------------
#define SAD_P2_SMEM ((SAD_TU * SAD_TV + SAD_TU * SAD_RP + SAD_PRIV * SAD_TV + SAD_JC * 8) * 4)

// One difference d, e = d - c, thresholds in e: ta0 <= tb0 <= ta1 <= tb1.  [tb0, ta1) is the middle (|d| summed), the rest
// of [ta0, tb1) are the candidates of the two quantiles (d parked).
#define SAD_P2_STEP(D, S, PTR, CC, TA0, TB0, TA1, TB1, STEP)                                                        \
    asm volatile(                                                                                                    \
        "{\n\t"                                                                                                      \
        ".reg .pred p0, pr, p2, pm, ph;\n\t"                                                                         \
        ".reg .f32 e, ad;\n\t"                                                                                       \
        "sub.f32 e, %2, %3;\n\t"                                                                                     \
        "setp.ge.f32 p0, e, %4;\n\t"                                                                                 \
        "setp.lt.and.f32 pr, e, %7, p0;\n\t"                                                                         \
        "setp.ge.f32 p2, e, %5;\n\t"                                                                                 \
        "setp.lt.and.f32 pm, e, %6, p2;\n\t"                                                                         \
        "and.pred ph, pr, !pm;\n\t"                                                                                  \
        "abs.f32 ad, %2;\n\t"                                                                                        \
        "@pm add.f32 %0, %0, ad;\n\t"                                                                                \
        "@ph st.shared.f32 [%1], %2;\n\t"                                                                            \
        "@ph add.s32 %1, %1, %8;\n\t"                                                                                \
        "}"                                                                                                          \
        : "+f"(S), "+r"(PTR)                                                                                         \
        : "f"(D), "f"(CC), "f"(TA0), "f"(TB0), "f"(TA1), "f"(TB1), "n"(STEP))

__global__ void __launch_bounds__(SAD_TV, 4) sad_pass2_kernel(const float* __restrict__ rz, const float* __restrict__ kz, int W,
                                                              int kh, int kw, int J, int nchunk, SadOff* off,
                                                              float* __restrict__ cand, double* __restrict__ part) {
    extern __shared__ __align__(16) unsigned char sad_smem[];
    float* ks = reinterpret_cast<float*>(sad_smem);
    float* rs = ks + SAD_TU * SAD_TV;
    float* sh_priv = rs + SAD_TU * SAD_RP;
    float* sh_b = sh_priv + SAD_PRIV * SAD_TV;      // [SAD_JC][8]: ta0 tb0 ta1 tb1 | c - - -
    __shared__ double s_red[SAD_TV / 32][SAD_JC];
    const int tid = threadIdx.x;
    const int oi = blockIdx.y / nchunk, j0 = (blockIdx.y - oi * nchunk) * SAD_JC;
    const int nj = min(SAD_JC, J - j0);
    const float inf = INFINITY;
    if (tid < SAD_JC) {
        float b[8] = {inf, inf, inf, inf, 0.f, 0.f, 0.f, 0.f};
        if (tid < nj) {
            const SadOff* q = off + oi * J + j0 + tid;
            if (!q->flag && !q->done) { b[0] = q->ta[0]; b[1] = q->tb[0]; b[2] = q->ta[1]; b[3] = q->tb[1]; b[4] = q->c; }
        }
        for (int k = 0; k < 8; ++k) sh_b[tid * 8 + k] = b[k];
    }
    double acc[SAD_JC];
#pragma unroll
    for (int jj = 0; jj < SAD_JC; ++jj) acc[jj] = 0.0;
    const unsigned mineA = (unsigned)__cvta_generic_to_shared(sh_priv + tid);
    const unsigned mineB = mineA + (SAD_PRIV - 1) * SAD_TV * 4;
    const int ntv = (kw + SAD_TV - 1) / SAD_TV, ntu = (kh + SAD_TU - 1) / SAD_TU;
    for (int t = blockIdx.x; t < ntv * ntu; t += gridDim.x) {
        const int tu = t / ntv, tv = t - tu * ntv;
        __syncthreads();
        sad_load_tile(rz, kz, W, kh, kw, oi, j0, tu * SAD_TU, tv * SAD_TV, ks, rs);
        __syncthreads();
        float kv[SAD_TU];
#pragma unroll
        for (int r = 0; r < SAD_TU; ++r) kv[r] = ks[r * SAD_TV + tid];
#pragma unroll
        for (int jp = 0; jp < SAD_JC; jp += 2) {
            if (jp >= nj) break;   // CTA-uniform
            const float4 ba = *reinterpret_cast<const float4*>(sh_b + jp * 8);
            const float4 bb = *reinterpret_cast<const float4*>(sh_b + (jp + 1) * 8);
            const float ca = sh_b[jp * 8 + 4], cb = sh_b[(jp + 1) * 8 + 4];
            unsigned pa = mineA, pb = mineB;
            float sa = 0.0f, sb = 0.0f;
            const float* rr = rs + tid + jp;
#pragma unroll
            for (int r = 0; r < SAD_TU; ++r) {
                const float da = rr[r * SAD_RP] - kv[r];
                const float db = rr[r * SAD_RP + 1] - kv[r];
                SAD_P2_STEP(da, sa, pa, ca, ba.x, ba.y, ba.z, ba.w, SAD_TV * 4);
                SAD_P2_STEP(db, sb, pb, cb, bb.x, bb.y, bb.z, bb.w, -SAD_TV * 4);
            }
            acc[jp] += (double)sa;         // <= SAD_TU float32 terms per partial
            acc[jp + 1] += (double)sb;
            // the few hundred candidates per offset and quantile: appended to the offset's global lists
            const unsigned na = (pa - mineA) / (SAD_TV * 4), nb = (mineB - pb) / (SAD_TV * 4);
            if (__any_sync(0xffffffffu, (na | nb) != 0u)) {
                for (unsigned k = 0; k < na; ++k) {
                    float d;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(d) : "r"(mineA + k * (SAD_TV * 4)));
                    const int q = (d - ca) < ba.y ? 0 : 1;
                    SadOff* so = off + oi * J + j0 + jp;
                    const unsigned pos = atomicAdd(&so->ncand[q], 1u);
                    if (pos < SAD_CAP) cand[((size_t)(oi * J + j0 + jp) * 2 + q) * SAD_CAP + pos] = d;
                }
                for (unsigned k = 0; k < nb; ++k) {
                    float d;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(d) : "r"(mineB - k * (SAD_TV * 4)));
                    const int q = (d - cb) < bb.y ? 0 : 1;
                    SadOff* so = off + oi * J + j0 + jp + 1;
                    const unsigned pos = atomicAdd(&so->ncand[q], 1u);
                    if (pos < SAD_CAP) cand[((size_t)(oi * J + j0 + jp + 1) * 2 + q) * SAD_CAP + pos] = d;
                }
            }
        }
    }
    // fixed-order reduction of the per-thread FP64 partials: lanes (shuffle tree), warps (serial), one slot per CTA
#pragma unroll
    for (int jj = 0; jj < SAD_JC; ++jj) {
        const double v = dcr_warp_sum(acc[jj]);
        if ((tid & 31) == 0) s_red[tid >> 5][jj] = v;
    }
    __syncthreads();
    if (tid < nj) {
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
    unsigned* rb; unsigned* kb;   // validity bit planes (rwp / kwp words per row)
    int rwp, kwp;
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
    int gx = (dcr_num_sms() * 12 + J * nchunk - 1) / (J * nchunk);   // ~3 waves of 4 CTAs per SM
    return gx < 1 ? 1 : gx;
}

static size_t sad_ws_bytes(int h, int w, int pad) {
    const long long kh = h - 2 * pad, kw = w - 2 * pad;
    const long long n = kh * kw;
    const long long JJ = (long long)(2 * pad + 1) * (2 * pad + 1);
    return dcr_al256((size_t)h * w * 4) + dcr_al256((size_t)n * 4) + dcr_al256((size_t)h * (w / 32 + 12) * 4) +
           dcr_al256((size_t)kh * ((kw + 31) / 32) * 4) + dcr_al256((size_t)JJ * sizeof(SadOff)) +
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
    s->rwp = w / 32 + 12;         // zero words of padding past the last (possibly partial) word: the count kernel's
                                  // 32-column offset chunks read up to 9 words past it
    s->kwp = (int)((kw + 31) / 32);
    s->rb = ws.take<unsigned>((size_t)h * s->rwp);
    s->kb = ws.take<unsigned>((size_t)kh * s->kwp);
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
    if (!s->rz || !s->kz || !s->rb || !s->kb || !s->off || !s->ghist || !s->cand || !s->part || !s->dm || !s->diff || !s->dmask || !s->sst ||
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
    DCR_CUDA_OK(cudaFuncSetAttribute(sad_pass2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SAD_P2_SMEM));
    SAD_MARK(0);
    DCR_CUDA_OK(cudaMemsetAsync(ws.rb, 0, (size_t)h * ws.rwp * 4, s));
    DCR_CUDA_OK(cudaMemsetAsync(ws.off, 0, (size_t)JJ * sizeof(SadOff), s));
    sad_prep_kernel<<<g, 256, 0, s>>>(ref, ref_mask, stride, 0, 0, h, w, ws.rz, ws.rb, ws.rwp);
    sad_prep_kernel<<<g, 256, 0, s>>>(src, src_mask, stride, pad, pad, kh, kw, ws.kz, ws.kb, ws.kwp);
    sad_count_kernel<<<dim3(SAD_CNT_SEG, J, (J + 31) / 32), 256, 0, s>>>(ws.rb, ws.rwp, ws.kb, ws.kwp, kh, J, ws.off);
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
