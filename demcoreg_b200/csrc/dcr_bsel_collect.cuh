// The streaming half of the sample-bracket select ("bsel") as a reusable device-side collector, so that a kernel
// which streams the data anyway (bsel_main_kernel, nuth_prepare_kernel) can gather the order-statistic candidates
// of one or more selections in the same pass.
//
// Per element:   add(x, ok)   -- counts the valid elements below / above the bracket in registers and parks the few
//                               per cent inside it in a thread-private shared-memory slot (no collective);
// per <= PRIV elements of a thread, at a WARP-uniform point:   drain()   -- moves the parked values of the warp to
//                               the CTA staging buffer (warp scan + one shared-memory reservation, uniform trip
//                               count) and histograms their bracket bin in shared memory;
// once per CTA after the stream:   finish()   -- appends the staging buffer to the global candidate buffer with ONE
//                               global reservation per CTA (same-address global atomics cost ~1.3 ns each: 30 000
//                               per-warp reservations added 45 us to a 76 us pass), merges counters + histogram; the
//                               LAST CTA locates the bracket-histogram bins of the two ranks and verifies the bracket.
// Callers size the grid so that a CTA expects fewer than STAGE candidates; the overflow path stays correct.
#pragma once
#include "dcr_internal.cuh"

#define BSC_THREADS 256

__device__ __forceinline__ void bsel_ranks(long long count, double q_percent, long long* klo, long long* khi, double* g) {
    const double vi = (double)(count - 1) * (q_percent / 100.0);  // exact rank in float64
    *klo = (long long)floor(vi);
    *khi = *klo + 1 < count ? *klo + 1 : count - 1;
    *g = vi - (double)(*klo);
}

__device__ __forceinline__ void bsel_publish(BselState* st) {
    if (st->publish) *st->publish = st->result;
}

// Level-0 find (run by the LAST CTA of bsel_main_kernel): the bracket-histogram bin of each of the two ranks,
// after VERIFYING that both ranks fall inside the bracket and that no candidate was dropped.
// off = key - klo; bin0 = off >> shift.  256 threads, 8 bins each.
// band != 0 (row-band mode): below / total / ghist have been summed over all ranks, the candidates stay on their ranks --
// the per-rank "no candidate dropped" checks were folded into the all-reduced overflow word ghist[2048] by the main pass.
__device__ __forceinline__ void bsel_find_block(BselState* st, unsigned* ghist, double q_percent, unsigned cap, unsigned* s_cum /*[2049]*/,
                                unsigned* s_scratch /*[33]*/, int band = 0) {
    const int t = threadIdx.x;
    unsigned v[8], sum = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { v[q] = __ldcg(ghist + t * 8 + q); sum += v[q]; }
    unsigned total;
    unsigned ex = dcr_block_excl_scan(sum, s_scratch, &total);
#pragma unroll
    for (int q = 0; q < 8; ++q) { s_cum[t * 8 + q] = ex; ex += v[q]; }
    if (t == 0) s_cum[2048] = total;
    const unsigned overflow = band ? __ldcg(ghist + 2048) : 0u;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) ghist[t * 8 + q] = 0u;
    if (band && t == 0) ghist[2048] = 0u;
    if (t == 0) {
        const long long count = (long long)st->total;
        st->count = count;
        if (count == 0) {
            st->empty = 1; st->done = 1;
            st->v_lo = st->v_hi = st->result = __int_as_float(0x7fc00000);
            bsel_publish(st);
        } else {
            long long klo, khi; double g;
            bsel_ranks(count, q_percent, &klo, &khi, &g);
            st->gamma = g;
            const long long c[2] = {klo - (long long)st->below, khi - (long long)st->below};
            const long long nin = (long long)total;
            const bool ok = band ? ((c[0] >= 0) && (c[1] < nin) && overflow == 0u)
                                 : ((c[0] >= 0) && (c[1] < nin) && (st->ncand <= cap) && ((long long)st->ncand == nin));
            if (!ok) { st->fallback = 1; st->done = 1; }
            else {
                for (int r = 0; r < 2; ++r) {
                    int lo = 0, hi = 2047;
                    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((long long)s_cum[mid] <= c[r]) lo = mid; else hi = mid - 1; }
                    st->prefix[r] = (unsigned)lo << st->shift;
                    st->rem[r] = c[r] - (long long)s_cum[lo];
                }
                st->cur_shift = st->shift;
                if (st->shift == 0) {  // bins are single keys already
                    const float x0 = dcr_key2f(st->klo + st->prefix[0]), x1 = dcr_key2f(st->klo + st->prefix[1]);
                    st->v_lo = x0; st->v_hi = x1;
                    st->result = dcr_lerp32(x0, x1, g);
                    st->done = 1;
                    bsel_publish(st);
                }
            }
        }
        st->ticket = 0;
    }
}


template <int PRIV, int STAGE>   // PRIV: elements a thread may add between two drains; STAGE: CTA staging (keys)
struct BselCollector {
    struct Shared {
        unsigned h[2048];                      // bracket histogram of this CTA
        float priv[PRIV * BSC_THREADS];        // slot k of thread t at [k * BSC_THREADS + t]
        unsigned buf[STAGE > 2049 ? STAGE : 2049];   // staged candidate keys (reused as scan scratch by the last CTA)
        unsigned long long below, total;
        unsigned n, gbase, last;
    };
    // per-thread state
    Shared* sh;
    BselState* st;
    unsigned* cand;
    unsigned cap;
    float* mine;                    // this thread's first private slot
    unsigned below, total, c;
    float flo, fhi;
    unsigned klo;
    int shift;
    bool on;

    // all threads of the CTA; the caller must __syncthreads() before the first add()
    __device__ __forceinline__ void init(Shared* s, BselState* state, unsigned* cand_buf, unsigned cand_cap) {
        sh = s; st = state; cand = cand_buf; cap = cand_cap;
        mine = s->priv + threadIdx.x;
        on = !state->done;
        for (int k = threadIdx.x; k < 2048; k += BSC_THREADS) s->h[k] = 0u;
        if (threadIdx.x == 0) { s->below = 0ull; s->total = 0ull; s->n = 0u; }
        below = total = c = 0;
        flo = state->flo; fhi = state->fhi; klo = state->klo; shift = state->shift;
        if (!on) { flo = INFINITY; fhi = INFINITY; shift = 31; }   // already solved (tiny input): collect nothing
    }

    // x: the (already transformed) value; ok: element takes part in the selection.  NaN fails every comparison.
    __device__ __forceinline__ void add(float x, bool ok) {
        // three chained compares: below the bracket | at or above its lower edge | of those, inside.  total = below +
        // nge, so a NaN (false in every comparison) takes part in neither; `total` holds nge until finish()
        const bool lo = ok && (x < flo), ge = ok && (x >= flo), in = ge && (x <= fhi);
        below += lo;
        total += ge;
        if (in) { mine[c * BSC_THREADS] = x; ++c; }
    }

    // warp-uniform: every lane of the warp must call it (after at most PRIV add() calls per thread).
    // The parked values move to the CTA staging buffer (one shared-memory reservation per warp); what does not fit
    // (a CTA that met more than STAGE candidates: very large rasters / wide brackets) goes straight to the global
    // buffer with one global reservation per warp and drain.
    __device__ __forceinline__ void drain() {
        const unsigned cmax = __reduce_max_sync(0xffffffffu, c);
        if (cmax == 0) return;
        const int lane = threadIdx.x & 31;
        const unsigned inc = dcr_warp_incl_scan(c);
        const unsigned wt = __shfl_sync(0xffffffffu, inc, 31);
        unsigned base = 0, gb = 0;
        if (lane == 31) {
            base = atomicAdd(&sh->n, wt);
            if (base + wt > STAGE) gb = atomicAdd(&st->ncand, base + wt - (base > STAGE ? base : STAGE));
        }
        base = __shfl_sync(0xffffffffu, base, 31);
        gb = __shfl_sync(0xffffffffu, gb, 31);
        unsigned pos = base + inc - c;
        if (base + wt <= STAGE) {   // warp-uniform common case: everything fits the staging buffer
            for (unsigned k = 0; k < cmax; ++k) {
                if (k < c) {
                    const unsigned key = dcr_f2key(mine[k * BSC_THREADS] + 0.0f);  // -0 -> +0
                    sh->buf[pos++] = key;
                    atomicAdd(&sh->h[(key - klo) >> shift], 1u);
                }
            }
        } else {
            const unsigned ov0 = base > STAGE ? base : STAGE;   // first position that overflows
            for (unsigned k = 0; k < cmax; ++k) {
                if (k < c) {
                    const unsigned key = dcr_f2key(mine[k * BSC_THREADS] + 0.0f);
                    if (pos < STAGE) sh->buf[pos] = key;
                    else if (gb + (pos - ov0) < cap) cand[gb + (pos - ov0)] = key;
                    ++pos;
                    atomicAdd(&sh->h[(key - klo) >> shift], 1u);
                }
            }
        }
        c = 0;
    }

    // all threads of the CTA, once, after the stream (contains barriers).  nctas = CTAs taking part in this selection.
    // s_scr: 33 words of shared scratch (used by the last CTA only).
    // band != 0: no find here -- the caller all-reduces below / total / ghist[0..2048] over the ranks first; the last CTA
    // only records whether this rank dropped candidates (ghist[2048], summed with the rest)
    __device__ __forceinline__ void finish(unsigned* ghist, double q_percent, unsigned nctas, unsigned* s_scr, int band = 0) {
        if (!on) return;
        const int lane = threadIdx.x & 31;
        drain();
        below = __reduce_add_sync(0xffffffffu, below);
        total = __reduce_add_sync(0xffffffffu, total) + below;
        if (lane == 0) {
            if (below) atomicAdd(&sh->below, (unsigned long long)below);
            if (total) atomicAdd(&sh->total, (unsigned long long)total);
        }
        __syncthreads();
        const unsigned nst = sh->n < STAGE ? sh->n : STAGE;
        if (threadIdx.x == 0) {
            sh->gbase = nst ? atomicAdd(&st->ncand, nst) : 0u;
            if (sh->below) atomicAdd(&st->below, sh->below);
            if (sh->total) atomicAdd(&st->total, sh->total);
        }
        for (int k = threadIdx.x; k < 2048; k += BSC_THREADS)
            if (sh->h[k]) atomicAdd(&ghist[k], sh->h[k]);
        __syncthreads();
        const unsigned gbs = sh->gbase;
        for (unsigned k = threadIdx.x; k < nst; k += BSC_THREADS)
            if (gbs + k < cap) cand[gbs + k] = sh->buf[k];
        // the last CTA to finish locates the ranks in the bracket histogram (saves a dependent launch)
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) sh->last = (atomicAdd(&st->ticket, 1u) == nctas - 1) ? 1u : 0u;
        __syncthreads();
        if (!sh->last) return;
        __threadfence();
        if (band) {
            if (threadIdx.x == 0) {
                const unsigned nc = *((volatile unsigned*)&st->ncand);
                ghist[2048] = (nc > cap) ? 1u : 0u;
                st->ticket = 0;
            }
            return;
        }
        bsel_find_block(st, ghist, q_percent, cap, sh->buf, s_scr);
    }
};
