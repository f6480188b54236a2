// Internal (non-ABI) interfaces shared between the translation units of libdemcoreg_b200.
#pragma once
#include "dcr_common.cuh"

// ---- per-(host thread, device) context (dcr_front.cu) ----
// Everything the library caches between calls is keyed by the CURRENT device of the calling thread: SM count, the
// >48 KB dynamic shared-memory opt-ins (cudaFuncSetAttribute is per device), the side streams / events of the iteration
// and the pinned result block.  One process may therefore drive several GPUs (one host thread per GPU, or one thread
// switching devices); there is no process-global mutable state.
#define DCR_MAX_DEVICES 64
enum { DCR_ATTR_FRONT = 1, DCR_ATTR_FBA = 2, DCR_ATTR_BSEL = 4 };
struct DcrCtx {
    int dev;                 // device ordinal this slot belongs to
    int sms;                 // multiprocessors
    unsigned attrs;          // DCR_ATTR_* opt-ins already done on this device
    cudaStream_t s2, s3;     // side streams of the iteration (created on first use)
    cudaEvent_t e_front, e_mask, e_join, e_join3, e_copy;
    cudaEvent_t prof[8];     // DCR_ITER_PROFILE stage marks
    void* pinned;            // pinned host block (result copies), pinned_bytes long
    size_t pinned_bytes;
};
DcrCtx* dcr_ctx();           // context of the calling thread's current device; nullptr (+ error text) on failure
void* dcr_ctx_pinned(DcrCtx* c, size_t bytes);   // pinned host scratch of at least `bytes` (grown on demand)

// ---- front end (dcr_front.cu) ----
// counters (device, 3 x u64, zeroed by the caller): base-valid, after max_dz, slope-valid
// bin (device, h*w, optional): also emit the aspect-bin code of every pixel (DCR_BIN_INVALID outside the masks known to
// the front end)
// loopmode: the slope raster receives tan(slope) instead of degrees (native dem_align loop; nothing else reads it)
// tile_list (device, tile_list_cap unsigned, optional): scratch that enables the fast interior-tile kernel
int dcr_front_launch(const dcr_front_args* a, unsigned long long* counters, cudaStream_t stream, unsigned short* bin = nullptr,
                     int loopmode = 0, unsigned* tile_list = nullptr, long long tile_list_cap = 0);
static inline long long dcr_front_list_cap(long long n) { return n / 512 + 65536; }

// ---- exact selection (dcr_select.cu) ----
// A "view" of the values taking part in a selection; all device pointers.
struct SelView {
    const float* x;
    const unsigned char* m;      // mask bytes or nullptr
    unsigned reject;             // element valid iff (m[i] & reject) == 0
    int transform;               // 0: x ; 1: |x - center| in float32
    const float* center_ptr;     // device scalar (takes precedence) or nullptr
    float center_imm;
    long long n;
    const long long* n_ptr;      // device element count (takes precedence) or nullptr
};

// Device-resident state/result of one selection.
struct SelState {
    unsigned prefix_lo, prefix_hi;   // radix prefixes of the two neighbouring order statistics
    long long k_lo, k_hi;            // remaining ranks inside the prefixes
    long long count;                 // participating elements
    double gamma;                    // interpolation weight
    int same;                        // prefix_lo == prefix_hi so far
    int empty;
    float v_lo, v_hi, result;        // result = numpy _lerp(v_lo, v_hi, gamma); NaN if empty
};

#define DCR_SEL_HIST_WORDS (2 * 2048)
// Enqueue an exact quantile selection (q in percent) of `v`; the result lands in *st (device).
// hist: device scratch of DCR_SEL_HIST_WORDS u32.
int dcr_sel_enqueue(const SelView& v, double q_percent, SelState* st, unsigned* hist, cudaStream_t stream);

// ---- sample-bracket selection (dcr_bsel.cu): one streaming pass, verified, with radix fallback ----
struct BselState {
    unsigned klo, khi;               // bracket keys (inclusive)
    float flo, fhi;                  // the same bracket as floats (the streaming pass compares floats, not keys)
    int shift;                       // (key - klo) >> shift indexes the 2048-bin bracket histogram
    int fallback;                    // 1: bracket missed / overflow / heavy ties -> result invalid, use the radix select
    int done, empty;
    unsigned long long below, total; // valid elements with key < klo ; all valid elements
    unsigned ncand, nsmall, ticket, ncollect;
    unsigned prefix[2];              // resolved high bits of (key - klo) for the two ranks
    int cur_shift, pad2;             // bits of (key - klo) still unresolved
    long long rem[2];                // ranks inside the resolved prefixes
    long long count;
    double gamma;
    float v_lo, v_hi, result;
    float chat;                      // dual median + MAD select: the centre estimate the streaming pass measures |x - chat| from
    float* publish;                  // optional device scalar that receives `result` as soon as it is final
};
unsigned dcr_bsel_cap(long long n);
size_t dcr_bsel_scratch_bytes(long long n);
int dcr_bsel_enqueue(const SelView& v, double q_percent, BselState* st, void* scratch, long long n_max, cudaStream_t s,
                     float* publish = nullptr);
// the same selection in three steps, for callers that fuse the streaming pass into a kernel of their own (BselCollector)
void dcr_bsel_buffers(void* scratch, long long n_max, unsigned** hist, unsigned** cand, unsigned* cap);
int dcr_bsel_sample_enqueue(const SelView& v, double q_percent, BselState* st, void* scratch, long long n_max, cudaStream_t s,
                            float* publish = nullptr);
int dcr_bsel_refine_enqueue(BselState* st, void* scratch, long long n_max, cudaStream_t s);
// median of v AND median of |v - that median| (filtlib.mad_fltr <- dem_align.py:46) with ONE streaming pass over the data:
// results into st0->result / st1->result (published to publish0 / publish1); a declined bracket sets st?->fallback
int dcr_msel_enqueue(const SelView& v, BselState* st0, BselState* st1, void* scratch0, void* scratch1, long long n_max,
                     cudaStream_t s, float* publish0, float* publish1);
// row-band mode: the selection cut into dcr_bsel_band_nsub() sub-phases with all-reduced hand-offs (see dcr_bsel.cu)
int dcr_bsel_band_nsub(void);
int dcr_bsel_band_phase(const SelView& v, double q_percent, BselState* st, void* scratch, long long n_max, int rank, int nranks,
                        int sub, cudaStream_t s, float* publish, void** red_ptr, int64_t* red_count, int* red_kind, int* nred);

// ---- nuth (dcr_nuth.cu) ----
struct NuthDev;  // device block

// ---- compression (dcr_select.cu) ----
long long dcr_nchunks(long long n);
int dcr_chunk_count_scan_enqueue(const float* x, const unsigned char* m, unsigned reject, long long n, unsigned* chunk_cnt,
                                 long long* chunk_off, long long* total, cudaStream_t s);
int dcr_compress_enqueue(const float* x, const unsigned char* m, unsigned reject, long long n, const long long* chunk_off,
                         long long stride_imm, const long long* stride_ptr, float* dense, long long cap,
                         cudaStream_t s, const long long* base_ptr = nullptr, int groups = 1);
int dcr_chunk_scan_enqueue(const unsigned* cnt, long long nchunks, long long* off, long long* total, cudaStream_t s);
int dcr_sel_hist_pass(const SelView& v, SelState* st, unsigned* hist, int pass, cudaStream_t s);
int dcr_sel_find_pass(SelState* st, unsigned* hist, double q_percent, int pass, cudaStream_t s);
