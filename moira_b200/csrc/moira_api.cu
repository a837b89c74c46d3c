// moira_api.cu -- C ABI (include/moira_b200.h) over the kernels in moira_kernels.cu:
// context, lookup tables, pass orchestration (first pass + escalation ladder, all enqueued
// without host synchronisation), and the chunked host<->device pipeline.
#include <cuda.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <new>
#include <thread>
#include <string>
#include <vector>

#include "moira_internal.h"

using namespace moira;

namespace {

thread_local char g_err[512] = "";

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(MOIRA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr uint32_t SUB_BATCH = 1u << 25;   // reads per first-pass launch when a ladder follows (bounds the queue memory)
constexpr int MAX_TIMED = 256;

struct Workspace {
    uint32_t *queues = nullptr;
    uint32_t *counts = nullptr;
    uint32_t cap = 0;      // reads per queue in the current layout
    size_t words = 0;      // allocated
    // length bucketing of ragged batches
    uint32_t *lqueue = nullptr;
    uint32_t *ltables = nullptr;   // hist | bucket_start | cursor | group_start | group_count
    uint32_t lcap = 0;
};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

// pinned parse buffers of one streaming slot (moira_filter_fastq)
struct StreamBuf {
    uint8_t *slab = nullptr;
    uint64_t *offsets = nullptr;
    uint32_t *lengths = nullptr;
    uint64_t slab_cap = 0, read_cap = 0;
    int ensure(uint64_t slab_bytes, uint64_t reads)
    {
        if (slab_bytes > slab_cap) {
            if (slab) cudaFreeHost(slab);
            slab = nullptr; slab_cap = 0;
            const uint64_t want = slab_bytes + slab_bytes / 4;
            if (cudaHostAlloc((void **)&slab, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return MOIRA_ERR_NOMEM; }
            slab_cap = want;
        }
        if (reads > read_cap) {
            if (offsets) cudaFreeHost(offsets);
            if (lengths) cudaFreeHost(lengths);
            offsets = nullptr; lengths = nullptr; read_cap = 0;
            const uint64_t want = reads + reads / 4 + 1024;
            if (cudaHostAlloc((void **)&offsets, want * 8, cudaHostAllocDefault) != cudaSuccess ||
                cudaHostAlloc((void **)&lengths, want * 4, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return MOIRA_ERR_NOMEM; }
            read_cap = want;
        }
        return MOIRA_OK;
    }
    void release()
    {
        if (slab) cudaFreeHost(slab);
        if (offsets) cudaFreeHost(offsets);
        if (lengths) cudaFreeHost(lengths);
        slab = nullptr; offsets = nullptr; lengths = nullptr; slab_cap = read_cap = 0;
    }
};

// device buffers of one chunk of read pairs (moira_filter_pairs)
struct PairBufs {
    DevBuf fseq, fqual, foff, flen, rseq, rqual, roff, rlen;
    DevBuf cseq, cqual, slab, clen, overlap, gaps, mism, status, ee, ns, flags;
    DevBuf *all[19] = {&fseq, &fqual, &foff, &flen, &rseq, &rqual, &roff, &rlen, &cseq, &cqual, &slab, &clen,
                       &overlap, &gaps, &mism, &status, &ee, &ns, &flags};
    // Pinned staging of the small per-pair results: a D2H copy into pageable user memory would block the host
    // until the chunk's kernels are done and serialise the two streams.  Flushed to the caller's arrays when the
    // stream is next synchronised.
    uint8_t *stage = nullptr;
    size_t stage_cap = 0;
    uint64_t pend_start = 0, pend_n = 0;   // chunk whose results sit in `stage`
    // Pageable callers (numpy arrays, memory-mapped files): the chunk's input bytes are gathered into `in_stage` by all host
    // threads and cross PCIe from there; the contig rows come back into `out_stage` and are copied to the caller's arrays
    // at the flush (a cudaMemcpyAsync from / to pageable memory runs at a fraction of the link and blocks the host).
    uint8_t *in_stage = nullptr, *out_stage = nullptr;
    size_t in_cap = 0, out_cap = 0;
    bool out_staged = false;
};
constexpr size_t PAIR_STAGE_BYTES = 8 + 4 * 5 + 2;   // ee | ns, clen, overlap, gaps, mism | flags, status

// one chunk of FASTQ text being parsed and filtered on the device
struct FqSlot {
    DevBuf text, bcnt, bstart, nl, soff, qoff, len, slab, ee, ns, flags, meta, marks;
    uint8_t *h_text = nullptr;     // pinned staging of the chunk's text (pageable callers)
    size_t h_text_cap = 0;
    uint8_t *h_res = nullptr;      // pinned staging of the results: ee | ns | len | flags
    size_t h_res_cap = 0;
    uint32_t *h_meta = nullptr;    // pinned: max length, min length, first bad record
    uint64_t pos = 0, bytes = 0, n_rec = 0, first_read = 0;
    bool counted = true;           // n_rec was counted by the host planner (else it is the device's, read back with the meta words)
    bool final = false;            // the chunk reaches the end of the text
    uint64_t skip = 0;             // the device buffer starts `skip` bytes before the chunk (page-aligned source of the copy)
    cudaEvent_t ev = nullptr;      // behind the chunk's latest D2H: what the host waits for, not the whole stream
    int state = 0;                 // 0 free, 1 indexed (meta on its way), 2 filtering (results on their way)
    int stream = 0;
    DevBuf *all[13] = {&text, &bcnt, &bstart, &nl, &soff, &qoff, &len, &slab, &ee, &ns, &flags, &meta, &marks};
};

struct Ticket {
    bool busy = false;
    DevBuf slab, slab6, offsets, lengths, ee, ns, flags, counters, marks;
    cudaEvent_t done[2] = {nullptr, nullptr};
    uint64_t *counters_out = nullptr;
    uint64_t *counters_pinned = nullptr;
    moira_params params;   // of the submission (feedback for the pilot-less cascade at moira_wait)
};

}  // namespace

namespace moira {
int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace moira

struct moira_ctx {
    int device = 0;
    int sm_count = 0;
    double h_p[256], h_q[256], h_e[256];
    int e_equals_p = 0;
    double *d_p = nullptr, *d_q = nullptr, *d_e = nullptr;
    double *d_sink = nullptr;
    cudaStream_t streams[2] = {nullptr, nullptr};
    cudaEvent_t meta_ready = nullptr;
    Workspace ws[3];          // [0],[1]: pipeline streams; [2]: moira_filter_device on a caller stream
    Ticket tickets[MOIRA_MAX_INFLIGHT];
    StreamBuf fq[3];
    uint64_t launches = 0;
    int use_tma = 1;
    int length_sort = 1;
    int cascade = 1;
    int cascade_multi = 1; // decisions needing 5..8 entries: first stage picked among 2..5 entries (0: two entries or none)
    double direct_gap = 0.5;
    int classify_first_k = 9;          // exact mode, first-pass K at least this: the classifier runs first and the ladder does all sweeps
    int classify_first_dec_k = 9;      // the same in decision mode
    int classify_first_sorted_k = 9;   // the same for length-bucketed (ragged) batches
    int sorted_cascade = 1;            // length-bucketed first pass in decision mode (K = 5..8): capped buckets + Newton bound + rungs, by pilot
    int sorted_cascade_cap = 4;
    int exact_sorted_k_cap = 4;        // length-bucketed first pass in exact mode: at most this many entries (0: the decision's K);
                                       // measured on C5 (100..600 bp): 4 -> 4.56 ms, 5 -> 4.68, 6 -> 4.88, none (3..8) -> 5.08, 3 -> 5.24
    int direct_rung = 1;   // the first pass picks the ladder rung of the reads it hands on (0: every one goes through the classifier)
    int timing = 0;
    int n_timed = 0;
    cudaEvent_t t0[MAX_TIMED] = {}, t1[MAX_TIMED] = {};
    const char *timed_name = "";
    // FASTQ parsing on the device (moira_filter_fastq): three chunks in flight
    FqSlot fqd[3];
    uint8_t *fq_ring[6] = {};        // pinned staging of the text of pageable callers, filled by the planner thread
    size_t fq_ring_cap[6] = {};
    DevBuf fq_counters;
    int device_parse = 1;
    int fq_guess_cuts = 1;
    // batches too small for a pilot launch take the cascade blindly (k_first <= 4); the escalated fraction the host
    // sees in a finished batch's counters switches that off for the next 32 batches when it was above the pilot's limit
    int blind_off_left = 0;
    // paired-end contig construction: per-stream device buffers, traceback scratch, posterior tables
    PairBufs pb[2];
    DevBuf trace, hbuf, post, pair_counters;
    bool post_ready = false;
    cudaEvent_t ct0[MAX_TIMED] = {}, ct1[MAX_TIMED] = {};   // around the contig kernel launches when timing is on
    int n_ctimed = 0;
    CommState *comm = nullptr;   // communicator rank for the counters' all-reduce (moira_comm.cpp), created on demand
    // dereplication on the device (moira_dedup.cu): hashes, table, labels; the FASTQ path's sequence store and its index
    DevBuf dd_hash, dd_table, dd_labels, dd_store, dd_seq_abs, dd_seq_eff;
    // groups from labels on the device (moira_groups.cu): device scratch and the pinned landing zone of its results
    DevBuf grp_dev;
    uint8_t *grp_host = nullptr;
    size_t grp_host_cap = 0;
    // single-read scratch (pinned)
    uint8_t *one_slab = nullptr;
    size_t one_cap = 0;
    double *one_out = nullptr;   // [0]=ee ; followed by ns (int32) and flags
};

namespace {

// Host-libm tables, built with the reference's own expressions:
//   p  = pow(10, Q / -10.0)                                   bernoullimodule.c:202
//   q  = pow(1 - p, 1)                                        bernoullimodule.c:140 (n == 1)
//   e  = ((1 - 1 + 1) / (1.0 * 1)) * (p / (1 - p)) * q        bernoullimodule.c:144 (i == 1, left to right)
// Q == 0 uses the Q == 1 entry (bernoullimodule.c:104-107, moira.py:814).
void build_tables(double *h_p, double *h_q, double *h_e, int *eqp)
{
    for (int b = 0; b < 256; b++) {
        if (b >= 0xFD) { h_p[b] = 0.0; h_q[b] = 1.0; h_e[b] = 0.0; continue; }
        const int Q = b == 0 ? 1 : b;
        const int n = 1, i = 1;
        volatile double p = pow(10, (Q / -10.0));
        volatile double q = pow((1 - p), n);
        volatile double ratio = p / (1 - p);
        volatile double lead = (n - i + 1) / (1.0 * i);
        volatile double t = lead * ratio;
        volatile double e = t * q;
        h_p[b] = p; h_q[b] = q; h_e[b] = e;
    }
    *eqp = 1;
    for (int b = 0; b < 256; b++) {
        volatile double omp = 1.0 - h_p[b];
        if (h_e[b] != h_p[b] || h_q[b] != omp) *eqp = 0;
    }
}

// pinned host memory that only grows
int ensure_host(uint8_t **p, size_t *cap, size_t bytes)
{
    if (bytes <= *cap) return MOIRA_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *cap = 0;
    const size_t want = bytes + bytes / 8 + 4096;
    if (cudaHostAlloc((void **)p, want, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return fail(MOIRA_ERR_NOMEM, "cudaHostAlloc of %zu bytes failed", want);
    }
    *cap = want;
    return MOIRA_OK;
}

int ensure(DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return MOIRA_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    if (cudaMalloc(&b.p, want) != cudaSuccess) {
        cudaGetLastError();
        return fail(MOIRA_ERR_NOMEM, "cudaMalloc of %zu bytes failed", want);
    }
    b.cap = want;
    return MOIRA_OK;
}

// queue memory for `nq` queues of `cap` reads each (the ladder uses NB queues, the decision cascade only queue 0)
int ensure_ws(Workspace &w, uint32_t cap, int nq = NB)
{
    if (!w.counts) CU(cudaMalloc(&w.counts, (NB + 4 + 16) * sizeof(uint32_t)));   // queue counts | first-pass policy word | pilot histogram
    const size_t words = (size_t)nq * cap;
    if (words <= w.words && w.queues) { w.cap = cap; return MOIRA_OK; }
    if (w.queues) cudaFree(w.queues);
    w.queues = nullptr; w.cap = 0; w.words = 0;
    if (cudaMalloc(&w.queues, words * sizeof(uint32_t)) != cudaSuccess) {
        cudaGetLastError();
        return fail(MOIRA_ERR_NOMEM, "cudaMalloc of escalation queues (%d x %u reads) failed", nq, cap);
    }
    w.cap = cap;
    w.words = words;
    return MOIRA_OK;
}

int ensure_lsort(Workspace &w, uint32_t cap, LenSortBufs &b)
{
    constexpr size_t TABLE_WORDS = 3 * LEN_BUCKETS + 2 * 32;
    if (!w.ltables) CU(cudaMalloc(&w.ltables, TABLE_WORDS * sizeof(uint32_t)));
    if (cap > w.lcap || !w.lqueue) {
        if (w.lqueue) cudaFree(w.lqueue);
        w.lqueue = nullptr; w.lcap = 0;
        if (cudaMalloc(&w.lqueue, (size_t)cap * sizeof(uint32_t)) != cudaSuccess) {
            cudaGetLastError();
            return fail(MOIRA_ERR_NOMEM, "cudaMalloc of the length-sort queue (%u reads) failed", cap);
        }
        w.lcap = cap;
    }
    b.queue = w.lqueue;
    b.hist = w.ltables;
    b.bucket_start = w.ltables + LEN_BUCKETS;
    b.cursor = w.ltables + 2 * LEN_BUCKETS;
    b.group_start = w.ltables + 3 * LEN_BUCKETS;
    b.group_count = b.group_start + 32;
    return MOIRA_OK;
}

// feedback for the pilot-less cascade from the counters of a finished Poisson-binomial batch (see moira_ctx::blind_off_left)
void note_escalations(moira_ctx *c, const moira_params *p, const uint64_t *counters)
{
    if (!p || p->mode != MOIRA_MODE_PB || p->exact_ee || !counters) return;
    if (c->blind_off_left > 0) { c->blind_off_left--; return; }
    const uint64_t reads = counters[MOIRA_CNT_READS], esc = counters[MOIRA_CNT_ESCALATED];
    if (counters[MOIRA_CNT_CLASSIFIED]) return;   // a classify-first batch says nothing about the cascade
    if (reads >= 4096 && esc * 100 > reads * 35) c->blind_off_left = 32;
}

int check_params(const moira_params *p)
{
    if (!p) return fail(MOIRA_ERR_BAD_ARG, "params is NULL");
    if (!(p->alpha > 0.0) || !(p->alpha < 1.0)) return fail(MOIRA_ERR_BAD_ALPHA, "Alpha must be between 0 and 1");
    volatile double oma = 1 - p->alpha;
    if (!(oma < 1.0)) return fail(MOIRA_ERR_BAD_ALPHA, "Alpha must be between 0 and 1 (1 - alpha rounds to 1)");
    if (p->mode < MOIRA_MODE_PB || p->mode > MOIRA_MODE_EXPECTED_ERROR) return fail(MOIRA_ERR_BAD_ARG, "unknown mode %d", p->mode);
    if (p->thr_kind != MOIRA_THR_UNCERT && p->thr_kind != MOIRA_THR_MAXERRORS) return fail(MOIRA_ERR_BAD_ARG, "unknown thr_kind %d", p->thr_kind);
    if (p->slab_format != MOIRA_SLAB_Q8 && p->slab_format != MOIRA_SLAB_Q6) return fail(MOIRA_ERR_BAD_ARG, "unknown slab_format %d", p->slab_format);
    if (p->ambigs < 0 || p->ambigs > MOIRA_AMBIGS_DISALLOW) return fail(MOIRA_ERR_BAD_ARG, "unknown ambigs %d", p->ambigs);
    if (!(p->thr == p->thr)) return fail(MOIRA_ERR_BAD_ARG, "threshold is NaN");
    return MOIRA_OK;
}

// Phi^-1(p) by bisection on erfc (host only; accuracy far beyond what a K estimate needs).
double normal_quantile(double p)
{
    double lo = -40.0, hi = 40.0;
    for (int i = 0; i < 200; i++) {
        const double mid = 0.5 * (lo + hi);
        const double cdf = 0.5 * erfc(-mid / sqrt(2.0));
        if (cdf < p) lo = mid; else hi = mid;
    }
    const double z = 0.5 * (lo + hi);
    return z < 0.0 ? 0.0 : z;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn()
{
    static const EncodeTiledFn fn = []() -> EncodeTiledFn {   // looked up once (thread-safe static initialisation)
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            return reinterpret_cast<EncodeTiledFn>(p);
        cudaGetLastError();
        return nullptr;
    }();
    return fn;
}

// 2-D uint8 tensor over `rows` slab rows of `stride` bytes; box = 128 bytes x 32 rows, 128-byte swizzle.
bool make_slab_tmap(CUtensorMap *tm, const uint8_t *d_rows, uint64_t stride, uint64_t rows)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || stride < 16 || (stride & 15u) || rows == 0 || rows >= (1ull << 31)) return false;
    cuuint64_t gdim[2] = {stride, rows};
    cuuint64_t gstr[1] = {stride};
    cuuint32_t box[2] = {128, 32};
    cuuint32_t estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(d_rows), gdim, gstr, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// labels of reads [0, n) (see moira_dedup.cu) into d_labels, enqueued on `stream`
int run_dedup(moira_ctx *c, const uint8_t *d_seq, const uint64_t *d_off, const uint32_t *d_len, uint64_t stride, uint32_t fixed_len,
              uint64_t n, uint32_t truncate, uint32_t *d_labels, cudaStream_t stream)
{
    if (n == 0) return MOIRA_OK;
    if (n >= 0xFFFFFFF0ull) return fail(MOIRA_ERR_BAD_ARG, "more than 2^32 reads in one dereplication");
    uint64_t slots = 1024;
    while (slots < 2 * n) slots <<= 1;
    int rc;
    if ((rc = ensure(c->dd_hash, n * 16)) || (rc = ensure(c->dd_table, slots * 4))) return rc;
    CU(cudaMemsetAsync(c->dd_table.p, 0xFF, slots * 4, stream));
    DedupArgs a;
    memset(&a, 0, sizeof(a));
    a.seq = d_seq; a.off = d_off; a.len = d_len; a.stride = stride; a.fixed_len = fixed_len; a.truncate = truncate; a.base = 0; a.n = n;
    LaunchCfg cfg{c->sm_count, stream};
    if (launch_dedup(a, (uint64_t *)c->dd_hash.p, (uint32_t *)c->dd_table.p, (uint32_t)(slots - 1), d_labels, cfg))
        return fail(MOIRA_ERR_CUDA, "dereplication launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    c->launches += 2;
    return MOIRA_OK;
}

int first_pass_k_template(int k_wanted)
{
    for (int i = 0; i < N_FIRST_K; i++) if (k_wanted <= first_pass_k(i)) return first_pass_k(i);
    return first_pass_k(N_FIRST_K - 1);
}

// Enqueue the whole filter for reads [0, n) on `stream`.  max_len = longest read if known, else 0.
int run_filter_full(moira_ctx *c, Workspace &ws, const uint8_t *d_slab, const uint64_t *d_offsets,
                    const uint32_t *d_lengths, uint64_t stride, uint32_t fixed_length, uint64_t n_reads,
                    const moira_params *p, uint32_t max_len, uint32_t min_len, double *d_ee, int32_t *d_ns, uint8_t *d_flags,
                    uint64_t *d_counters, cudaStream_t stream, const uint32_t *d_row_marks = nullptr)
{
    if (n_reads == 0) return MOIRA_OK;
    if (!d_slab || !d_ee) return fail(MOIRA_ERR_BAD_ARG, "slab / ee pointer is NULL");
    if (((uintptr_t)d_slab & 15u) != 0) return fail(MOIRA_ERR_BAD_ARG, "slab must be 16-byte aligned");
    if (!d_offsets && (stride & 15u)) return fail(MOIRA_ERR_BAD_ARG, "stride must be a multiple of 16");
    if (!d_lengths) max_len = fixed_length;
    FilterArgs a;
    memset(&a, 0, sizeof(a));
    a.slab = d_slab; a.offsets = d_offsets; a.lengths = d_lengths; a.stride = stride; a.fixed_length = fixed_length;
    a.ee = d_ee; a.ns = d_ns; a.flags = d_flags; a.counters = reinterpret_cast<unsigned long long *>(d_counters);
    volatile double oma = 1 - p->alpha;   // evaluated like (1 - alpha) at bernoullimodule.c:244
    a.oma = oma;
    a.thr = p->thr;
    a.z = normal_quantile(1.0 - p->alpha);
    a.zc = (a.z * a.z - 1.0) / 6.0 + 1.5;   // skew bound + K = j*+1 and rounding (1.5): no under-estimate in 28 000 test reads
    a.mode = p->mode; a.thr_kind = p->thr_kind; a.ambigs = p->ambigs; a.round_flag = p->round_flag;
    a.truncate = p->truncate; a.exact = p->exact_ee; a.ee_output = p->ee_output;
    a.lut_p = c->d_p; a.lut_q = c->d_q; a.lut_e = c->d_e; a.e_equals_p = c->e_equals_p;
    a.row_marks = d_row_marks;   // Ns / has-N per row from the slab's producer (over the first eff bases): the sweeps count nothing
    a.rung = -1;
    LaunchCfg cfg{c->sm_count, stream};

    // First-pass K: a decision needs floor(cutoff) + 2 PMF entries and the cutoff is
    // largest for the longest read.  Unknown lengths: start at 4, the ladder settles the rest.
    int k_first = 4;
    bool k_decides_all = false;
    if (p->mode == MOIRA_MODE_PB) {
        uint32_t eff = max_len;
        if (p->truncate && (eff == 0 || eff > p->truncate)) eff = p->truncate;
        const bool known = p->thr_kind == MOIRA_THR_MAXERRORS || eff != 0;
        if (known) {
            volatile double prod = (double)eff * p->thr;
            double cmax = p->thr_kind == MOIRA_THR_MAXERRORS ? p->thr : (double)prod;
            double kd = cmax < 0.0 ? 2.0 : floor(cmax) + 2.0;
            if (kd <= (double)max_first_pass_k()) { k_first = (int)kd; k_decides_all = true; }
            else k_first = max_first_pass_k();
        }
        k_first = first_pass_k_template(k_first);
        a.min_rung = 1;
        while (a.min_rung < NB - 1 && rung_cap(a.min_rung) <= k_first) a.min_rung++;
    }
    // Classify first: where the decision's own K is large (long reads, lax cutoffs) a first pass with that K is paid by every
    // read although most need far fewer entries -- or, in exact mode, far more, and then the pass was for nothing.  The
    // classifier (fp32 mean / variance of the error count: a quarter of a two-entry sweep) reads every row once and sends each
    // read to the rung that holds its own j* + 1 entries (decision mode: at most the decision's K); the ladder does all the
    // FP64 work.  Length-bucketed (ragged) batches: the classifier walks the sorted permutation.
    const bool cf_plain = p->mode == MOIRA_MODE_PB && k_first >= (p->exact_ee ? c->classify_first_k : c->classify_first_dec_k);
    const bool cf_sorted = p->mode == MOIRA_MODE_PB && k_first >= (p->exact_ee ? c->classify_first_sorted_k : c->classify_first_dec_k);
    const bool ladder = p->mode == MOIRA_MODE_PB && (p->exact_ee || !k_decides_all || cf_plain || cf_sorted);
    a.allow_push = ladder ? 1 : 0;
    a.direct_rung = (ladder && c->direct_rung) ? 1 : 0;
    a.direct_gap = c->direct_gap;
    // Cascaded first pass: when the decision needs only a few entries (k_first = 3..8), most reads of a typical run are
    // either clean (j* <= 1: two entries settle them exactly) or hopeless (the Newton bound on acc[k_first - 1] from
    // those two entries already rejects them).  Sweeping two entries costs 4-5 FP64 operations per base instead of
    // 3 k_first - 2; only the reads in between are swept again with k_first entries (decision mode) or walk the ladder
    // (exact mode).  Whether that pays depends on the data: a pilot launch over the first tiles measures the escalated
    // fraction on the device, and the two candidate launches for the rest read that verdict (no host synchronisation).
    const bool cascade_ok = p->mode == MOIRA_MODE_PB && k_decides_all && k_first >= 3 && k_first <= 8 && p->cascade != 2 && c->cascade && !cf_plain;
    // Length-bucketed (ragged) batches whose decisions need 5 .. 8 entries: the same idea per bucket -- at most four entries
    // first, the Newton bound at every read's own K, the rest to the rung that holds that K -- if a pilot over the first reads
    // (sorted and swept the same way) hands on few enough.
    const bool sc_possible = p->mode == MOIRA_MODE_PB && !p->exact_ee && k_decides_all && k_first >= 5 && k_first <= 8 && p->cascade != 2 &&
                             c->cascade && c->sorted_cascade && d_lengths && c->length_sort && !cf_sorted;
    const bool queues_needed = ladder || cascade_ok || sc_possible;
    if (cascade_ok) a.min_rung = 1;   // what a two-entry sweep hands on may need as few as three entries

    const uint64_t sub = ladder ? SUB_BATCH : (cascade_ok || sc_possible) ? (1ull << 26) : (1ull << 31);
    for (uint64_t start = 0; start < n_reads; start += sub) {
        const uint32_t n = (uint32_t)std::min<uint64_t>(sub, n_reads - start);
        a.base = start; a.n = n; a.queue = nullptr; a.queue_count = nullptr; a.rung = -1;
        if (queues_needed) {
            int rc = ensure_ws(ws, (uint32_t)std::min<uint64_t>(n_reads, sub), ladder ? NB : sc_possible ? 9 : 1);   // rung_cap(8) == 12
            if (rc) return rc;
            a.queues = ws.queues; a.queue_counts = ws.counts; a.queue_cap = ws.cap;
            CU(cudaMemsetAsync(ws.counts, 0, (NB + 4 + 16) * sizeof(uint32_t), stream));
        }
        // uniform-stride slab: feed the first pass with TMA tensor tiles of this sub-batch's rows
        CUtensorMap tmap;
        cfg.tmap = nullptr;
        if (!d_offsets && c->use_tma && make_slab_tmap(&tmap, d_slab + start * stride, stride, n))
            cfg.tmap = &tmap;   // tile t of the launch = rows [32 t, 32 t + 32) of this map
        const bool timed = c->timing && c->n_timed < MAX_TIMED;
        if (timed) CU(cudaEventRecord(c->t0[c->n_timed], stream));
        const char *name = "";
        int rc = 0;
        // Ragged batch: bucket the reads by length on the device first, so that a warp tile holds reads of
        // (almost) one length and every tile runs with the K its own cutoff needs.
        // auto: only when the caller's lengths are known to spread by more than a quarter (host path); the
        // device-pointer API sorts only on request (length_sort = 1)
        // (reads beyond the last length bucket would share its first-pass K: such batches are not sorted)
        const bool lsort = d_lengths && c->length_sort && n >= 32768 && p->length_sort != 2 && max_len <= 16u * (LEN_BUCKETS - 1) &&
                           (p->length_sort == 1 || (max_len && min_len * 4 < max_len * 3));
        if (lsort) {
            cfg.tmap = nullptr;   // tiles follow the sorted permutation, not the slab order
            LenSortBufs lb;
            if ((rc = ensure_lsort(ws, (uint32_t)std::min<uint64_t>(n_reads, sub), lb))) return rc;
            CU(cudaMemsetAsync(lb.hist, 0, LEN_BUCKETS * sizeof(uint32_t), stream));
            const int single = p->mode != MOIRA_MODE_PB;
            a.first_k_cap = (ladder && p->exact_ee && a.allow_push) ? c->exact_sorted_k_cap : 0;
            // decisions that need 5 .. 8 entries: capped buckets by pilot (below); sorts its own ranges
            const bool sc = sc_possible && !single && !cf_sorted && (p->cascade == 1 || n >= 16u * 2u * (uint32_t)c->sm_count * 16u * 32u);
            if (!sc) {
                if (launch_length_sort(a, lb, single || cf_sorted, cfg)) return fail(MOIRA_ERR_CUDA, "length-sort launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                c->launches += 3;
            }
            a.queue = lb.queue;
            a.min_rung = 1;
            if (single) {
                a.seg_start = lb.group_start;
                a.seg_count = lb.group_count;
                a.queue_count = lb.group_count;
                rc = launch_lambda(a, cfg, &name);
            } else if (cf_sorted) {
                // the classifier walks the length-sorted permutation, so the rung queues fill in (almost) length order and the
                // ladder's warp tiles hold reads of (almost) one length
                a.seg_start = lb.group_start;
                a.seg_count = lb.group_count;
                a.queue_count = lb.group_count;
                rc = launch_classify_first(a, cfg);
                name = "classifier over the length-sorted reads + ladder";
            } else if (sc) {
                const uint32_t pilot_n = 2u * (uint32_t)c->sm_count * 16u * 32u;
                uint32_t *policy = ws.counts + NB;
                FilterArgs as = a;
                as.allow_push = 1;
                as.first_k_cap = c->sorted_cascade_cap;
                if (p->cascade != 1) {
                    // pilot: the first reads of the sub-batch (slab order: a fair sample of the lengths), bucketed and swept with the cap
                    as.queue = nullptr; as.n = pilot_n;
                    if (launch_length_sort(as, lb, 0, cfg)) return fail(MOIRA_ERR_CUDA, "length-sort launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                    as.queue = lb.queue;
                    rc = launch_sorted_first(as, lb.group_start, lb.group_count, cfg);
                    if (rc >= 0 && launch_policy_sorted(ws.counts, (uint32_t)(pilot_n * 0.20), policy, stream)) rc = -1;
                    CU(cudaMemsetAsync(lb.hist, 0, LEN_BUCKETS * sizeof(uint32_t), stream));
                    as.cap_policy = policy; as.first_read = pilot_n;
                    c->launches += 6;
                }
                as.queue = nullptr; as.n = n;
                if (rc >= 0 && launch_length_sort(as, lb, 0, cfg)) return fail(MOIRA_ERR_CUDA, "length-sort launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                as.queue = lb.queue;
                if (rc >= 0) rc = launch_sorted_first(as, lb.group_start, lb.group_count, cfg);
                // what the capped buckets handed on: one more sweep each, with the entries the read's own decision needs
                as.queue = nullptr; as.first_read = 0; as.cap_policy = nullptr;
                if (rc >= 0 && launch_ladder_tpr(as, cfg)) rc = -1;
                name = "pb_tpr<K per length bucket, at most 4 first>";
                c->launches += 3 + 1 + 2;
            } else {
                rc = launch_sorted_first(a, lb.group_start, lb.group_count, cfg);
                name = "pb_tpr<K per length bucket>";
                c->launches++;   // two launches: K <= 12 on 16 warps, wider on 8
            }
            if (rc < 0) return fail(MOIRA_ERR_CUDA, "first-pass launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            c->launches++;
            a.queue = nullptr; a.queue_count = nullptr; a.seg_start = nullptr; a.seg_count = nullptr;
        } else if (cf_plain) {
            a.min_rung = 1;
            rc = launch_classify_first(a, cfg);
            name = "classifier + ladder";
            if (rc < 0) return fail(MOIRA_ERR_CUDA, "classifier launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            c->launches++;
        } else if (cascade_ok) {
            // two-entry launches: escalate to queue 0; in decision mode certain rejects are settled by the bound
            FilterArgs a2 = a;
            a2.allow_push = 1;
            a2.k_dec = p->exact_ee ? 0 : k_first;
            const uint32_t warps = (uint32_t)c->sm_count * 16u;            // tpr_warps(2) per CTA
            const uint32_t pilot_n = 2u * warps * 32u;                     // two tiles per warp
            const bool pilot = p->cascade != 1 && n >= 8u * pilot_n;
            const bool blind = p->cascade == 1 || (k_first <= 4 && c->blind_off_left == 0);   // without a pilot: only where k_first is small
            uint32_t *policy = ws.counts + NB;
            static const char *cname[] = {"", "", "", "pb_cascade<2,3>", "pb_cascade<2,4>", "pb_cascade<2,5>", "pb_cascade<2,6>",
                                          "pb_cascade<2,7>", "pb_cascade<2,8>"};
            static const char *mname[] = {"", "", "", "", "", "pb_cascade<2..4,5>", "pb_cascade<2..5,6>", "pb_cascade<2..5,7>",
                                          "pb_cascade<2..5,8>"};
            // decisions that need 5 .. 8 entries: the first stage is chosen among 2 .. 5 entries (two pilots, see policy_*_kernel)
            const bool multi = pilot && k_first >= 5 && n >= 16u * pilot_n && c->cascade_multi;
            if (multi) {
                uint32_t *jhist = ws.counts + NB + 4;
                a2.n = pilot_n;
                rc = launch_pb_first(a2, 2, cfg, nullptr);                 // pilot A: two entries, escalations counted
                if (rc >= 0 && launch_policy_first(ws.counts, (uint32_t)(pilot_n * 0.35), policy, stream)) rc = -1;
                FilterArgs b = a;                                          // pilot B (only if A was escalated too often):
                b.n = 2u * pilot_n; b.tile0 = pilot_n / 32u;               // k_first entries at once over the next reads,
                b.policy = policy; b.policy_want = MOIRA_POLICY_UNDECIDED; // histogram of floor(ee) for the verdict
                b.jhist = jhist;
                if (rc >= 0) rc = launch_pb_first(b, k_first, cfg, nullptr);
                if (rc >= 0 && launch_policy_second(jhist, k_first, p->exact_ee ? 1 : 0, policy, stream)) rc = -1;
                a2.n = n; a2.tile0 = pilot_n / 32u; a2.policy = policy; a2.policy_want = 2;
                if (rc >= 0) rc = launch_pb_first(a2, 2, cfg, nullptr);
                c->launches += 5;
                for (int k1 = 3; k1 <= 5 && k1 < k_first && rc >= 0; k1++) {
                    FilterArgs ak = a2;
                    ak.tile0 = 2u * pilot_n / 32u; ak.policy_want = (uint32_t)k1;
                    rc = launch_pb_first(ak, k1, cfg, nullptr);
                    c->launches++;
                }
                FilterArgs a1 = a;                                         // the last candidate: k_first entries at once
                a1.tile0 = 2u * pilot_n / 32u; a1.policy = policy; a1.policy_want = 0;
                if (rc >= 0) rc = launch_pb_first(a1, k_first, cfg, nullptr);
                c->launches++;
            } else if (pilot) {
                a2.n = pilot_n;
                rc = launch_pb_first(a2, 2, cfg, nullptr);
                if (rc >= 0 && launch_policy(ws.counts, (uint32_t)(pilot_n * 0.35), policy, stream)) rc = -1;
                a2.n = n; a2.tile0 = pilot_n / 32u; a2.policy = policy; a2.policy_want = 0;
                if (rc >= 0) rc = launch_pb_first(a2, 2, cfg, nullptr);
                FilterArgs a1 = a;                                         // the other candidate: k_first entries at once
                a1.tile0 = pilot_n / 32u; a1.policy = policy; a1.policy_want = 1;
                if (rc >= 0) rc = launch_pb_first(a1, k_first, cfg, nullptr);
                c->launches += 4;
            } else if (blind) {
                rc = launch_pb_first(a2, 2, cfg, nullptr);
                c->launches++;
            } else {
                rc = launch_pb_first(a, k_first, cfg, &name);
                c->launches++;
            }
            if (rc >= 0 && !p->exact_ee && (pilot || blind)) {
                // second sweep, k_first entries, over the escalated reads only: settles every one of them
                FilterArgs a3 = a;
                a3.queue = ws.queues; a3.queue_count = ws.counts; a3.allow_push = 0;
                LaunchCfg cfg3 = cfg;
                cfg3.tmap = nullptr;
                rc = launch_pb_first(a3, k_first, cfg3, nullptr);
                c->launches++;
            }
            if (pilot || blind) name = multi ? mname[k_first] : cname[k_first];
            if (rc < 0) return fail(MOIRA_ERR_CUDA, "first-pass launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        } else {
            rc = p->mode == MOIRA_MODE_PB ? launch_pb_first(a, k_first, cfg, &name) : launch_lambda(a, cfg, &name);
            if (rc < 0) return fail(MOIRA_ERR_CUDA, "first-pass launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            c->launches++;
        }
        if (timed) { CU(cudaEventRecord(c->t1[c->n_timed], stream)); c->n_timed++; c->timed_name = name; }
        if (ladder) {
            for (int b = 0; b < NB; b++) {
                if (b == 1) {   // every thread-per-read rung in one launch
                    if (launch_ladder_tpr(a, cfg)) return fail(MOIRA_ERR_CUDA, "ladder launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                    c->launches += 2;   // rungs 1..8 on 16 warps, 9..19 on 8
                    b = N_TPR_RUNGS;
                    continue;
                }
                if (launch_rung(a, b, cfg)) return fail(MOIRA_ERR_CUDA, "rung %d launch failed: %s", b, cudaGetErrorString(cudaGetLastError()));
                c->launches++;
            }
        }
    }
    return MOIRA_OK;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int moira_abi_version(void) { return MOIRA_ABI_VERSION; }
const char *moira_last_error(void) { return g_err; }

void moira_params_default(moira_params *p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->mode = MOIRA_MODE_PB;                 // --error_calc poisson_binomial (moira.py:655-657)
    p->thr_kind = MOIRA_THR_UNCERT;
    p->ambigs = MOIRA_AMBIGS_TREAT_AS_ERRORS;  // moira.py:658-660
    p->alpha = 0.005;                        // moira.py:668
    p->thr = 0.01;                           // moira.py:664
    p->exact_ee = 1;
    p->ee_output = MOIRA_EE_RAW;
}

static int ctx_init(moira_ctx *c, int device, int sm_count);

int moira_ctx_create(int device, moira_ctx **out)
{
    if (!out) return fail(MOIRA_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(MOIRA_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(MOIRA_ERR_BAD_ARG, "device %d out of range (0..%d)", device, ndev - 1);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(MOIRA_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    moira_ctx *c = new (std::nothrow) moira_ctx();
    if (!c) return fail(MOIRA_ERR_NOMEM, "out of host memory");
    const int rc = ctx_init(c, device, prop.multiProcessorCount);
    if (rc) { moira_ctx_destroy(c); return rc; }   // g_err keeps the failing step's message
    *out = c;
    return MOIRA_OK;
}

static int ctx_init(moira_ctx *c, int device, int sm_count)
{
    c->device = device;
    c->sm_count = sm_count;
    build_tables(c->h_p, c->h_q, c->h_e, &c->e_equals_p);
    if (const char *e = getenv("MOIRA_B200_NO_LENSORT")) c->length_sort = (e[0] == '1') ? 0 : 1;   // diagnostics
    if (const char *e = getenv("MOIRA_B200_FQ_COUNT")) c->fq_guess_cuts = (e[0] == '1') ? 0 : 1;   // diagnostics: always count the chunk cuts on the host
    if (const char *e = getenv("MOIRA_B200_HOST_PARSE")) c->device_parse = (e[0] == '1') ? 0 : 1;   // moira_filter_fastq: parse on the host cores instead
    if (const char *e = getenv("MOIRA_B200_NO_CASCADE")) c->cascade = (e[0] == '1') ? 0 : 1;   // diagnostics: full-K first pass always
    if (const char *e = getenv("MOIRA_B200_NO_CASCADE_MULTI")) c->cascade_multi = (e[0] == '1') ? 0 : 1;   // diagnostics
    if (const char *e = getenv("MOIRA_B200_DIRECT_GAP")) c->direct_gap = atof(e);   // tuning
    if (const char *e = getenv("MOIRA_B200_NO_DIRECT_RUNG")) c->direct_rung = (e[0] == '1') ? 0 : 1;   // diagnostics: classifier pass for all
    if (const char *e = getenv("MOIRA_B200_CLASSIFY_FIRST_K")) c->classify_first_k = atoi(e);   // tuning (1000: never)
    if (const char *e = getenv("MOIRA_B200_CLASSIFY_FIRST_DEC_K")) c->classify_first_dec_k = atoi(e);   // tuning
    if (const char *e = getenv("MOIRA_B200_CLASSIFY_FIRST_SORTED_K")) c->classify_first_sorted_k = atoi(e);   // tuning
    if (const char *e = getenv("MOIRA_B200_NO_SORTED_CASCADE")) c->sorted_cascade = (e[0] == '1') ? 0 : 1;   // diagnostics
    if (const char *e = getenv("MOIRA_B200_SORTED_CASCADE_CAP")) c->sorted_cascade_cap = atoi(e);   // tuning
    if (const char *e = getenv("MOIRA_B200_EXACT_SORTED_KCAP")) c->exact_sorted_k_cap = atoi(e);   // tuning
    if (const char *e = getenv("MOIRA_B200_NO_TMA")) c->use_tma = (e[0] == '1') ? 0 : 1;   // diagnostics: force the cp.async staging
    if (kernels_init(c->sm_count)) return fail(MOIRA_ERR_CUDA, "kernel attribute setup failed: %s", cudaGetErrorString(cudaGetLastError()));
    CU(cudaMalloc(&c->d_p, 256 * sizeof(double)));
    CU(cudaMalloc(&c->d_q, 256 * sizeof(double)));
    CU(cudaMalloc(&c->d_e, 256 * sizeof(double)));
    CU(cudaMalloc(&c->d_sink, 64));
    CU(cudaMemcpy(c->d_p, c->h_p, sizeof(c->h_p), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_q, c->h_q, sizeof(c->h_q), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_e, c->h_e, sizeof(c->h_e), cudaMemcpyHostToDevice));
    for (int i = 0; i < 2; i++) CU(cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->meta_ready, cudaEventDisableTiming));
    for (int i = 0; i < MAX_TIMED; i++) { CU(cudaEventCreate(&c->t0[i])); CU(cudaEventCreate(&c->t1[i])); }
    for (int i = 0; i < MAX_TIMED; i++) { CU(cudaEventCreate(&c->ct0[i])); CU(cudaEventCreate(&c->ct1[i])); }
    for (auto &t : c->tickets)
        for (int i = 0; i < 2; i++) CU(cudaEventCreateWithFlags(&t.done[i], cudaEventDisableTiming));
    return MOIRA_OK;
}

int moira_ctx_destroy(moira_ctx *c)
{
    if (!c) return MOIRA_OK;
    const bool trace = getenv("MOIRA_B200_TRACE_DESTROY") != nullptr;
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; };
    const double t_begin = now();
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (trace) fprintf(stderr, "[moira_ctx_destroy] device synchronised after %.3f s\n", now() - t_begin);
    for (auto &t : c->tickets) {
        for (DevBuf *b : {&t.slab, &t.slab6, &t.offsets, &t.lengths, &t.ee, &t.ns, &t.flags, &t.counters, &t.marks})
            if (b->p) cudaFree(b->p);
        for (int i = 0; i < 2; i++) if (t.done[i]) cudaEventDestroy(t.done[i]);
        if (t.counters_pinned) cudaFreeHost(t.counters_pinned);
    }
    for (auto &b : c->fq) b.release();
    for (auto &q : c->fqd) {
        for (DevBuf *b : q.all) if (b->p) cudaFree(b->p);
        if (q.h_text) cudaFreeHost(q.h_text);
        if (q.h_res) cudaFreeHost(q.h_res);
        if (q.h_meta) cudaFreeHost(q.h_meta);
        if (q.ev) cudaEventDestroy(q.ev);
    }
    for (uint8_t *r : c->fq_ring) if (r) cudaFreeHost(r);
    if (c->fq_counters.p) cudaFree(c->fq_counters.p);
    for (auto &pb : c->pb) {
        for (DevBuf *b : pb.all) if (b->p) cudaFree(b->p);
        if (pb.stage) cudaFreeHost(pb.stage);
        if (pb.in_stage) cudaFreeHost(pb.in_stage);
        if (pb.out_stage) cudaFreeHost(pb.out_stage);
    }
    for (DevBuf *b : {&c->trace, &c->hbuf, &c->post, &c->pair_counters, &c->dd_hash, &c->dd_table, &c->dd_labels, &c->dd_store,
                      &c->dd_seq_abs, &c->dd_seq_eff, &c->grp_dev})
        if (b->p) cudaFree(b->p);
    if (c->grp_host) cudaFreeHost(c->grp_host);
    for (auto &w : c->ws) {
        if (w.queues) cudaFree(w.queues);
        if (w.counts) cudaFree(w.counts);
        if (w.lqueue) cudaFree(w.lqueue);
        if (w.ltables) cudaFree(w.ltables);
    }
    for (int i = 0; i < MAX_TIMED; i++) {
        if (c->t0[i]) cudaEventDestroy(c->t0[i]);
        if (c->t1[i]) cudaEventDestroy(c->t1[i]);
        if (c->ct0[i]) cudaEventDestroy(c->ct0[i]);
        if (c->ct1[i]) cudaEventDestroy(c->ct1[i]);
    }
    if (c->meta_ready) cudaEventDestroy(c->meta_ready);
    for (int i = 0; i < 2; i++) if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
    if (c->one_slab) cudaFreeHost(c->one_slab);
    if (c->one_out) cudaFreeHost(c->one_out);
    comm_state_free(c->comm);
    cudaSetDevice(c->device);
    cudaFree(c->d_p); cudaFree(c->d_q); cudaFree(c->d_e); cudaFree(c->d_sink);
    cudaGetLastError();
    if (trace) fprintf(stderr, "[moira_ctx_destroy] everything freed after %.3f s\n", now() - t_begin);
    delete c;
    return MOIRA_OK;
}

int moira_device_count(int *out)
{
    if (!out) return fail(MOIRA_ERR_BAD_ARG, "out is NULL");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
    *out = n;
    return MOIRA_OK;
}

int moira_ctx_sm_count(const moira_ctx *c, int *out)
{
    if (!c || !out) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    *out = c->sm_count;
    return MOIRA_OK;
}

int moira_build_lut(double p[256], double q[256], double e[256], int *e_equals_p)
{
    double hp[256], hq[256], he[256];
    int eqp = 0;
    build_tables(hp, hq, he, &eqp);
    if (p) memcpy(p, hp, sizeof(hp));
    if (q) memcpy(q, hq, sizeof(hq));
    if (e) memcpy(e, he, sizeof(he));
    if (e_equals_p) *e_equals_p = eqp;
    return MOIRA_OK;
}

int moira_ctx_get_lut(const moira_ctx *c, double p[256], double q[256], double e[256])
{
    if (!c) return fail(MOIRA_ERR_BAD_ARG, "ctx is NULL");
    if (p) memcpy(p, c->h_p, sizeof(c->h_p));
    if (q) memcpy(q, c->h_q, sizeof(c->h_q));
    if (e) memcpy(e, c->h_e, sizeof(c->h_e));
    return MOIRA_OK;
}

int moira_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return fail(MOIRA_ERR_BAD_ARG, "ptr is NULL");
    *ptr = nullptr;
    if (cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return fail(MOIRA_ERR_NOMEM, "cudaHostAlloc of %zu bytes failed", bytes);
    }
    return MOIRA_OK;
}

int moira_host_free(void *ptr)
{
    if (ptr) CU(cudaFreeHost(ptr));
    return MOIRA_OK;
}

int moira_filter_device(moira_ctx *c, const uint8_t *d_slab, const uint64_t *d_offsets, const uint32_t *d_lengths,
                        uint64_t stride, uint32_t fixed_length, uint64_t n_reads, const moira_params *params,
                        const uint32_t *d_row_marks, double *d_ee, int32_t *d_ns, uint8_t *d_flags, uint64_t *d_counters,
                        void *stream)
{
    if (!c) return fail(MOIRA_ERR_BAD_ARG, "ctx is NULL");
    int rc = check_params(params);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    return run_filter_full(c, c->ws[2], d_slab, d_offsets, d_lengths, stride, fixed_length, n_reads, params, params->max_length,
                           params->min_length, d_ee, d_ns, d_flags, d_counters, (cudaStream_t)stream, d_row_marks);
}

int moira_collapse_device(moira_ctx *c, const uint8_t *d_seq, const uint64_t *d_offsets, const uint32_t *d_lengths, uint64_t stride,
                          uint32_t fixed_length, uint64_t n_reads, uint32_t truncate, uint32_t *d_labels, void *stream)
{
    if (!c || !d_seq || !d_labels) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    return run_dedup(c, d_seq, d_offsets, d_lengths, stride, fixed_length, n_reads, truncate, d_labels, (cudaStream_t)stream);
}

// Labels of sequences that lie anywhere in host memory (contig rows of several batches, sequence lines of a FASTA text):
// the rows are gathered back to back into pinned staging by all host threads, shipped in chunks (the gather of chunk k + 1
// runs while chunk k crosses PCIe), labelled on the device exactly like moira_collapse_device's, and the labels come back.
int moira_collapse_addr(moira_ctx *c, const uint64_t *seq_addr, const uint32_t *seq_len, uint64_t n, uint32_t truncate, uint32_t *labels_out)
{
    if (!c || (n && (!seq_addr || !seq_len || !labels_out))) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    if (n == 0) return MOIRA_OK;
    if (n >= 0xFFFFFFF0ull) return fail(MOIRA_ERR_BAD_ARG, "more than 2^32 reads in one dereplication");
    CU(cudaSetDevice(c->device));
    cudaStream_t stream = c->streams[0];
    const int T = (int)std::max(1u, std::thread::hardware_concurrency());
    // offsets of the packed rows
    std::vector<uint64_t> off(n + 1);
    std::vector<uint32_t> eff(n);
    uint64_t total = 0;
    for (uint64_t r = 0; r < n; r++) {
        const uint32_t l = (truncate && seq_len[r] > truncate) ? truncate : seq_len[r];
        off[r] = total; eff[r] = l; total += l;
    }
    off[n] = total;
    int rc;
    if ((rc = ensure(c->dd_store, total + 64)) || (rc = ensure(c->dd_seq_abs, n * 8 + 8)) || (rc = ensure(c->dd_seq_eff, n * 4 + 4)) ||
        (rc = ensure(c->dd_labels, n * 4 + 4)))
        return rc;
    constexpr uint64_t CHUNK_BYTES = 128ull << 20;
    const uint64_t stage_bytes = std::min<uint64_t>(total, CHUNK_BYTES) + (64u << 10);
    if ((rc = ensure_host(&c->grp_host, &c->grp_host_cap, 2 * stage_bytes))) return rc;
    cudaEvent_t freed[2];
    for (auto &e : freed) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    uint64_t r0 = 0;
    int slot = 0;
    bool used[2] = {false, false};
    while (r0 < n) {
        uint64_t r1 = r0;
        while (r1 < n && off[r1 + 1] - off[r0] <= stage_bytes) r1++;
        if (r1 == r0) { rc = fail(MOIRA_ERR_BAD_ARG, "a sequence of %u bytes does not fit the staging buffer", eff[r0]); break; }
        uint8_t *stage = c->grp_host + (size_t)slot * stage_bytes;
        if (used[slot]) cudaEventSynchronize(freed[slot]);
        const int parts = T * 4;
        parallel_run(parts, T, [&](int p) {
            for (uint64_t r = r0 + (r1 - r0) * (uint64_t)p / parts, e = r0 + (r1 - r0) * (uint64_t)(p + 1) / parts; r < e; r++)
                memcpy(stage + (off[r] - off[r0]), reinterpret_cast<const void *>((uintptr_t)seq_addr[r]), eff[r]);
        });
        if (cudaMemcpyAsync((uint8_t *)c->dd_store.p + off[r0], stage, off[r1] - off[r0], cudaMemcpyHostToDevice, stream) != cudaSuccess) {
            rc = fail(MOIRA_ERR_CUDA, "H2D of the sequences failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        cudaEventRecord(freed[slot], stream);
        used[slot] = true;
        slot ^= 1;
        r0 = r1;
    }
    if (!rc) {
        if (cudaMemsetAsync((uint8_t *)c->dd_store.p + total, 0, 64, stream) != cudaSuccess ||
            cudaMemcpyAsync(c->dd_seq_abs.p, off.data(), n * 8, cudaMemcpyHostToDevice, stream) != cudaSuccess ||
            cudaMemcpyAsync(c->dd_seq_eff.p, eff.data(), n * 4, cudaMemcpyHostToDevice, stream) != cudaSuccess)
            rc = fail(MOIRA_ERR_CUDA, "H2D of the sequence index failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (!rc) rc = run_dedup(c, (const uint8_t *)c->dd_store.p, (const uint64_t *)c->dd_seq_abs.p, (const uint32_t *)c->dd_seq_eff.p, 0, 0, n, 0,
                            (uint32_t *)c->dd_labels.p, stream);
    if (!rc && cudaMemcpyAsync(labels_out, c->dd_labels.p, n * 4, cudaMemcpyDeviceToHost, stream) != cudaSuccess)
        rc = fail(MOIRA_ERR_CUDA, "D2H of the labels failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (cudaStreamSynchronize(stream) != cudaSuccess && !rc) rc = fail(MOIRA_ERR_CUDA, "dereplication failed: %s", cudaGetErrorString(cudaGetLastError()));
    for (auto &e : freed) cudaEventDestroy(e);
    return rc;
}

int moira_collapse_groups(moira_ctx *c, const uint32_t *labels, const double *ee, int on_device, uint64_t n, uint64_t *n_groups_out,
                          const uint32_t **group_of_read, const uint32_t **group_rep, const uint32_t **group_size,
                          const uint32_t **member_start, const uint32_t **members, const uint32_t **abundance_order)
{
    if (!c || !n_groups_out || !group_of_read || !group_rep || !group_size || !member_start || !members || !abundance_order)
        return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    *n_groups_out = 0;
    *group_of_read = *group_rep = *group_size = *member_start = *members = *abundance_order = nullptr;
    if (n == 0) return MOIRA_OK;
    if (!labels || !ee) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure(c->grp_dev, groups_device_bytes(n)))) return rc;
    if ((rc = ensure_host(&c->grp_host, &c->grp_host_cap, groups_host_bytes(n)))) return rc;
    size_t off[6];
    uint64_t G = 0;
    if ((rc = groups_from_labels_device(c->sm_count, c->streams[0], (uint8_t *)c->grp_dev.p, c->grp_host, labels, ee, on_device, n, &G, off)))
        return rc;
    c->launches += 8;
    *n_groups_out = G;
    const uint32_t **dst[6] = {group_of_read, group_rep, group_size, member_start, members, abundance_order};
    for (int i = 0; i < 6; i++) *dst[i] = reinterpret_cast<const uint32_t *>(c->grp_host + off[i]);
    return MOIRA_OK;
}

int moira_collapse_labels_device(moira_ctx *c, const uint32_t *labels, const double *ee, uint64_t n, uint64_t *group_of_read,
                                 uint64_t *n_groups_out, uint64_t *group_rep, uint64_t *group_size, uint64_t *member_start,
                                 uint64_t *members, uint64_t *abundance_order)
{
    if (!n_groups_out) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    *n_groups_out = 0;
    if (n == 0) return MOIRA_OK;
    if (!group_of_read || !group_rep || !group_size || !member_start || !members || !abundance_order)
        return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    const uint32_t *h[6];
    uint64_t G = 0;
    const int rc = moira_collapse_groups(c, labels, ee, 0, n, &G, &h[0], &h[1], &h[2], &h[3], &h[4], &h[5]);
    if (rc) return rc;
    *n_groups_out = G;
    // uint32 on the device and over PCIe, uint64 in the caller's arrays
    uint64_t *dst[6] = {group_of_read, group_rep, group_size, member_start, members, abundance_order};
    const uint64_t cnt[6] = {n, G, G, G + 1, n, G};
    const int parts = 64;
    parallel_run(parts * 6, (int)std::max(1u, std::thread::hardware_concurrency()), [&](int t) {
        const int a = t / parts, p = t % parts;
        const uint64_t lo = cnt[a] * (uint64_t)p / parts, hi = cnt[a] * (uint64_t)(p + 1) / parts;
        for (uint64_t i = lo; i < hi; i++) dst[a][i] = h[a][i];
    });
    return MOIRA_OK;
}

int moira_count_marks_device(moira_ctx *c, const uint8_t *d_slab, const uint64_t *d_offsets, const uint32_t *d_lengths,
                             uint64_t stride, uint32_t fixed_length, uint64_t n_reads, uint32_t truncate, uint32_t *d_row_marks,
                             void *stream)
{
    if (!c || !d_slab || !d_row_marks) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    if (((uintptr_t)d_slab & 15u) != 0) return fail(MOIRA_ERR_BAD_ARG, "slab must be 16-byte aligned");
    if (!d_offsets && (stride & 15u)) return fail(MOIRA_ERR_BAD_ARG, "stride must be a multiple of 16");
    CU(cudaSetDevice(c->device));
    FilterArgs a;
    memset(&a, 0, sizeof(a));
    a.slab = d_slab; a.offsets = d_offsets; a.lengths = d_lengths; a.stride = stride; a.fixed_length = fixed_length;
    a.truncate = truncate;
    LaunchCfg cfg{c->sm_count, (cudaStream_t)stream};
    for (uint64_t start = 0; start < n_reads; start += 1ull << 30) {
        a.base = start;
        a.n = (uint32_t)std::min<uint64_t>(1ull << 30, n_reads - start);
        if (launch_count_marks(a, d_row_marks, d_lengths ? 0 : fixed_length, cfg))
            return fail(MOIRA_ERR_CUDA, "count-marks launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        c->launches++;
    }
    return MOIRA_OK;
}

static int submit_impl(moira_ctx *c, const uint8_t *slab, uint64_t slab_bytes, const uint64_t *offsets,
                       const uint32_t *lengths, uint64_t n, const moira_params *params, double *ee_out, int32_t *ns_out,
                       uint8_t *flags_out, uint64_t *counters_out, int *ticket_out);

int moira_submit(moira_ctx *c, const uint8_t *slab, uint64_t slab_bytes, const uint64_t *offsets,
                 const uint32_t *lengths, uint64_t n, const moira_params *params, double *ee_out, int32_t *ns_out,
                 uint8_t *flags_out, uint64_t *counters_out, int *ticket_out)
{
    const int rc = submit_impl(c, slab, slab_bytes, offsets, lengths, n, params, ee_out, ns_out, flags_out, counters_out, ticket_out);
    if (rc && c) {
        // a failure part-way may have left copies and kernels of earlier chunks in flight: drain them, so
        // that nothing writes into the caller's buffers after the error is returned (g_err is kept)
        for (cudaStream_t s : c->streams) if (s) cudaStreamSynchronize(s);
        cudaGetLastError();
    }
    return rc;
}

static int submit_impl(moira_ctx *c, const uint8_t *slab, uint64_t slab_bytes, const uint64_t *offsets,
                       const uint32_t *lengths, uint64_t n, const moira_params *params, double *ee_out, int32_t *ns_out,
                       uint8_t *flags_out, uint64_t *counters_out, int *ticket_out)
{
    if (!c || !ticket_out) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    *ticket_out = -1;
    int rc = check_params(params);
    if (rc) return rc;
    if (n && (!slab || !offsets || !lengths || !ee_out)) return fail(MOIRA_ERR_BAD_ARG, "NULL buffer");
    CU(cudaSetDevice(c->device));
    int ti = -1;
    for (int i = 0; i < MOIRA_MAX_INFLIGHT; i++) if (!c->tickets[i].busy) { ti = i; break; }
    if (ti < 0) return fail(MOIRA_ERR_BAD_ARG, "more than %d submissions in flight", MOIRA_MAX_INFLIGHT);
    Ticket &t = c->tickets[ti];
    t.counters_out = counters_out;
    if (!t.counters_pinned) CU(cudaHostAlloc((void **)&t.counters_pinned, MOIRA_N_COUNTERS * sizeof(uint64_t), cudaHostAllocDefault));
    memset(t.counters_pinned, 0, MOIRA_N_COUNTERS * sizeof(uint64_t));
    t.params = *params;
    if (n == 0) {
        t.busy = true;
        for (int i = 0; i < 2; i++) CU(cudaEventRecord(t.done[i], c->streams[i]));
        *ticket_out = ti;
        return MOIRA_OK;
    }
    const bool q6 = params->slab_format == MOIRA_SLAB_Q6;   // `slab` is the 3/4-size transport image; offsets/lengths are in slab units
    const uint64_t slab8_bytes = q6 ? slab_bytes / 12 * 16 : slab_bytes;
    if (q6 && ((rc = ensure(t.slab6, slab_bytes + 256)) || (rc = ensure(t.marks, n * 4)))) return rc;
    if ((rc = ensure(t.slab, slab8_bytes + 256)) || (rc = ensure(t.offsets, n * 8)) || (rc = ensure(t.lengths, n * 4)) ||
        (rc = ensure(t.ee, n * 8)) || (rc = ensure(t.ns, n * 4)) || (rc = ensure(t.flags, n)) ||
        (rc = ensure(t.counters, MOIRA_N_COUNTERS * 8)))
        return rc;
    uint8_t *d_slab = (uint8_t *)t.slab.p;
    uint64_t *d_off = (uint64_t *)t.offsets.p;
    uint32_t *d_len = (uint32_t *)t.lengths.p;
    double *d_ee = (double *)t.ee.p;
    int32_t *d_ns = (int32_t *)t.ns.p;
    uint8_t *d_fl = (uint8_t *)t.flags.p;
    uint64_t *d_cnt = (uint64_t *)t.counters.p;

    cudaStream_t s0 = c->streams[0], s1 = c->streams[1];
    CU(cudaMemsetAsync(d_cnt, 0, MOIRA_N_COUNTERS * 8, s0));
    CU(cudaEventRecord(c->meta_ready, s0));
    CU(cudaStreamWaitEvent(s1, c->meta_ready, 0));

    // Chunks of ~32 MB of slab alternate between the two streams: the H2D copy of one chunk overlaps
    // the kernels of the previous one and the D2H copy of the one before.  Rows must be in slab order.
    uint64_t CHUNK_BYTES = 32ull << 20;
    if (const char *e = getenv("MOIRA_B200_CHUNK_MB")) { const long v = atol(e); if (v > 0) CHUNK_BYTES = (uint64_t)v << 20; }   // tuning
    uint64_t start = 0;
    int ci = 0;
    while (start < n) {
        uint64_t end = start;
        const uint64_t b0 = offsets[start];
        uint32_t max_len = 0, min_len = 0xFFFFFFFFu;
        uint64_t b1 = b0;
        // uniform chunk (every row `stride0` bytes after the previous one, all of one length): offsets and
        // lengths need not travel, and the first pass can be fed by TMA tensor tiles
        bool uniform = true, same_len = true;
        const uint64_t stride0 = start + 1 < n ? offsets[start + 1] - offsets[start] : (((uint64_t)lengths[start] + 15u) & ~15ull);
        while (end < n) {
            const uint64_t row_end = offsets[end] + (((uint64_t)lengths[end] + 15u) & ~15ull);
            if (offsets[end] < b0) return fail(MOIRA_ERR_BAD_ARG, "offsets must be non-decreasing (read %llu)", (unsigned long long)end);
            if (row_end > slab8_bytes + 15) return fail(MOIRA_ERR_BAD_ARG, "read %llu extends past the slab", (unsigned long long)end);
            if (offsets[end] & 15u) return fail(MOIRA_ERR_BAD_ARG, "offset of read %llu is not a multiple of 16", (unsigned long long)end);
            if (end > start && row_end - b0 > CHUNK_BYTES) break;
            b1 = std::max(b1, row_end);
            max_len = std::max(max_len, lengths[end]);
            min_len = std::min(min_len, lengths[end]);
            same_len = same_len && lengths[end] == lengths[start];
            uniform = uniform && offsets[end] == b0 + (end - start) * stride0;
            end++;
        }
        cudaStream_t s = c->streams[ci & 1];
        const uint64_t cn = end - start;
        uniform = uniform && stride0 >= 16 && (stride0 & 15u) == 0 && stride0 >= max_len;
        const uint32_t *d_marks = nullptr;
        if (q6) {
            // image bytes [b0*3/4, b1*3/4) travel (b0, b1 are multiples of 16), then expand on the device
            const uint64_t i0 = b0 / 16 * 12, i1 = std::min<uint64_t>((b1 + 15) / 16 * 12, slab_bytes);
            uint8_t *d_img = (uint8_t *)t.slab6.p;
            if (i1 > i0) CU(cudaMemcpyAsync(d_img + i0, slab + i0, i1 - i0, cudaMemcpyHostToDevice, s));
            LaunchCfg ucfg{c->sm_count, s};
            // uniform rows: the expanding kernel sees every byte anyway and leaves Ns / has-N per row for the filter
            uint32_t *marks = nullptr;
            uint32_t eff = lengths[start];
            if (params->truncate && eff > params->truncate) eff = params->truncate;
            if (uniform && same_len) {
                marks = (uint32_t *)t.marks.p + start;
                CU(cudaMemsetAsync(marks, 0, cn * 4, s));
                d_marks = marks;
            }
            if (launch_unpack_q6(d_img + i0, d_slab + b0, b1 - b0, marks, stride0, eff, ucfg)) return fail(MOIRA_ERR_CUDA, "unpack launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            c->launches++;
        } else {
            const uint64_t copy_end = std::min<uint64_t>(b1, slab_bytes);
            if (copy_end > b0) CU(cudaMemcpyAsync(d_slab + b0, slab + b0, copy_end - b0, cudaMemcpyHostToDevice, s));
        }
        if (uniform && same_len) {
            rc = run_filter_full(c, c->ws[ci & 1], d_slab + b0, nullptr, nullptr, stride0, lengths[start], cn, params, max_len, max_len,
                                 d_ee + start, d_ns + start, d_fl + start, d_cnt, s, d_marks);
        } else if (uniform) {   // rows at a fixed pitch, lengths vary: only the lengths travel; TMA tiles still apply
            CU(cudaMemcpyAsync(d_len + start, lengths + start, cn * 4, cudaMemcpyHostToDevice, s));
            rc = run_filter_full(c, c->ws[ci & 1], d_slab + b0, nullptr, d_len + start, stride0, 0, cn, params, max_len, min_len,
                                 d_ee + start, d_ns + start, d_fl + start, d_cnt, s);
        } else {
            CU(cudaMemcpyAsync(d_off + start, offsets + start, cn * 8, cudaMemcpyHostToDevice, s));
            CU(cudaMemcpyAsync(d_len + start, lengths + start, cn * 4, cudaMemcpyHostToDevice, s));
            rc = run_filter_full(c, c->ws[ci & 1], d_slab, d_off + start, d_len + start, 0, 0, cn, params, max_len, min_len,
                                 d_ee + start, d_ns + start, d_fl + start, d_cnt, s);
        }
        if (rc) return rc;
        CU(cudaMemcpyAsync(ee_out + start, d_ee + start, cn * 8, cudaMemcpyDeviceToHost, s));
        if (ns_out) CU(cudaMemcpyAsync(ns_out + start, d_ns + start, cn * 4, cudaMemcpyDeviceToHost, s));
        if (flags_out) CU(cudaMemcpyAsync(flags_out + start, d_fl + start, cn, cudaMemcpyDeviceToHost, s));
        start = end;
        ci++;
    }
    CU(cudaEventRecord(t.done[1], s1));
    CU(cudaStreamWaitEvent(s0, t.done[1], 0));
    CU(cudaMemcpyAsync(t.counters_pinned, d_cnt, MOIRA_N_COUNTERS * 8, cudaMemcpyDeviceToHost, s0));
    CU(cudaEventRecord(t.done[0], s0));
    t.busy = true;
    *ticket_out = ti;
    return MOIRA_OK;
}

int moira_wait(moira_ctx *c, int ticket)
{
    if (!c || ticket < 0 || ticket >= MOIRA_MAX_INFLIGHT || !c->tickets[ticket].busy)
        return fail(MOIRA_ERR_BAD_ARG, "invalid ticket %d", ticket);
    Ticket &t = c->tickets[ticket];
    t.busy = false;
    CU(cudaEventSynchronize(t.done[1]));
    CU(cudaEventSynchronize(t.done[0]));
    if (t.counters_out) memcpy(t.counters_out, t.counters_pinned, MOIRA_N_COUNTERS * sizeof(uint64_t));
    note_escalations(c, &t.params, t.counters_pinned);
    CU(cudaGetLastError());
    return MOIRA_OK;
}

int moira_filter_batch(moira_ctx *c, const uint8_t *slab, uint64_t slab_bytes, const uint64_t *offsets,
                       const uint32_t *lengths, uint64_t n, const moira_params *params, double *ee_out,
                       int32_t *ns_out, uint8_t *flags_out, uint64_t *counters_out)
{
    int ticket = -1;
    int rc = moira_submit(c, slab, slab_bytes, offsets, lengths, n, params, ee_out, ns_out, flags_out, counters_out, &ticket);
    if (rc) {
        if (c) { cudaSetDevice(c->device); cudaDeviceSynchronize(); }
        return rc;
    }
    return moira_wait(c, ticket);
}

namespace {

int ensure_pinned(uint8_t **p, size_t *cap, size_t bytes)
{
    if (bytes <= *cap) return MOIRA_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *cap = 0;
    const size_t want = bytes + bytes / 8 + 4096;
    if (cudaHostAlloc((void **)p, want, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return fail(MOIRA_ERR_NOMEM, "cudaHostAlloc of %zu bytes failed", want);
    }
    *cap = want;
    return MOIRA_OK;
}

// FASTQ text -> decisions with the parsing on the device.  Per chunk of ~64 MB of text (cut behind a complete
// record by fastq_plan_chunk): H2D of the text, newline index, record table (lengths, validation), then -- once the
// host has read back the chunk's max / min length -- slab conversion, the filter, and D2H of the results.  Three
// chunks are in flight: while chunk k is being indexed, chunk k-1 is filtered and chunk k-2 is handed to the caller.
constexpr int FQ_RETRY_COUNTED = 0x7fff0001;   // internal: a guessed chunk cut failed the device's check -> run again, counting on the host

int filter_fastq_device(moira_ctx *c, const char *text, uint64_t text_bytes, int fastq_offset, int lower_n, const moira_params *params,
                        uint64_t max_reads, double *ee_out, int32_t *ns_out, uint8_t *flags_out, uint32_t *lengths_out,
                        uint64_t *seq_off_out, uint64_t *qual_off_out, uint32_t *labels_out,
                        uint64_t *counters_out, uint64_t *n_reads_out, bool guess_cuts)
{
    constexpr uint64_t RANGE = 64ull << 20;
    int rc;
    if ((rc = ensure(c->fq_counters, MOIRA_N_COUNTERS * 8))) return rc;
    const bool want_off = seq_off_out || qual_off_out;
    if (labels_out) {
        // --collapse on the device: the (truncated) sequences stay in HBM at their own text offsets until the last chunk is in
        if (text_bytes > (96ull << 30)) return fail(MOIRA_ERR_NOMEM, "text too large to keep its sequences on the device (%llu bytes)", (unsigned long long)text_bytes);
        if ((rc = ensure(c->dd_store, text_bytes + 64)) || (rc = ensure(c->dd_seq_abs, max_reads * 8 + 8)) ||
            (rc = ensure(c->dd_seq_eff, max_reads * 4 + 4)) || (rc = ensure(c->dd_labels, max_reads * 4 + 4)))
            return rc;
    }
    uint64_t *d_cnt = (uint64_t *)c->fq_counters.p;
    CU(cudaMemsetAsync(d_cnt, 0, MOIRA_N_COUNTERS * 8, c->streams[0]));
    CU(cudaEventRecord(c->meta_ready, c->streams[0]));
    CU(cudaStreamWaitEvent(c->streams[1], c->meta_ready, 0));
    cudaPointerAttributes attr;
    const bool pinned_src = text_bytes && cudaPointerGetAttributes(&attr, text) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    for (auto &q : c->fqd) q.state = 0;
    uint64_t n_assigned = 0;   // records of the chunks filtered so far (chunk order): where a chunk's results go

    // stage 2 of a chunk: its index is done -> convert, filter, results on their way
    auto filter_chunk = [&](FqSlot &q) -> int {
        cudaStream_t s = c->streams[q.stream];
        CU(cudaEventSynchronize(q.ev));
        if (!q.counted) {
            // the device's own count: a chunk cut at a guessed record start must hold whole records (the last one may
            // end with a partial record, ignored as the reference ignores it)
            if (q.h_meta[5] || (q.h_meta[4] && !q.final)) return FQ_RETRY_COUNTED;
            q.n_rec = q.h_meta[3];
        }
        q.first_read = n_assigned;
        if (n_assigned + q.n_rec > max_reads) return fail(MOIRA_ERR_BAD_ARG, "more than max_reads = %llu records", (unsigned long long)max_reads);
        n_assigned += q.n_rec;
        if (q.n_rec == 0) { q.state = 0; return MOIRA_OK; }
        {
            int r0;
            if ((r0 = ensure(q.ee, q.n_rec * 8)) || (r0 = ensure(q.ns, q.n_rec * 4)) || (r0 = ensure(q.flags, q.n_rec)) ||
                (r0 = ensure_pinned(&q.h_res, &q.h_res_cap, q.n_rec * 25 + 64)))
                return r0;
        }
        const uint32_t maxlen = q.h_meta[0], minlen = q.h_meta[1], bad = q.h_meta[2];
        if (bad != 0xFFFFFFFFu) {
            // let the host parser raise the reference's error for this range (same message as the host path)
            std::vector<uint8_t> slab(16);
            uint64_t n = 0, cap = 0;
            int r = parse_fastq_range(text + q.pos, q.bytes, fastq_offset, lower_n, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr,
                                      nullptr, 0, &n, &cap, nullptr, 1);
            if (!r) {
                slab.resize(cap + 16);
                std::vector<uint64_t> off(n + 1);
                std::vector<uint32_t> len(n + 1);
                r = parse_fastq_range(text + q.pos, q.bytes, fastq_offset, lower_n, slab.data(), slab.size(), off.data(), len.data(), nullptr,
                                      nullptr, nullptr, nullptr, n, &n, &cap, nullptr, 1);
            }
            return r ? r : fail(MOIRA_ERR_PARSE, "record %u of the range at byte %llu is malformed", bad, (unsigned long long)q.pos);
        }
        const uint32_t stride = std::max<uint32_t>(16, (maxlen + 15u) & ~15u);
        int r;
        if ((r = ensure(q.slab, (size_t)q.n_rec * stride + 256)) || (r = ensure(q.marks, q.n_rec * 4))) return r;
        if (launch_fq_convert((const uint8_t *)q.text.p, (const uint32_t *)q.soff.p, (const uint32_t *)q.qoff.p, (const uint32_t *)q.len.p,
                              (uint32_t)q.n_rec, stride, lower_n, fastq_offset, (uint8_t *)q.slab.p, (uint32_t *)q.meta.p,
                              (uint32_t *)q.marks.p, params->truncate, c->sm_count, s, labels_out ? (uint8_t *)c->dd_store.p : nullptr,
                              q.pos - q.skip, labels_out ? (uint64_t *)c->dd_seq_abs.p + q.first_read : nullptr,
                              labels_out ? (uint32_t *)c->dd_seq_eff.p + q.first_read : nullptr))
            return fail(MOIRA_ERR_CUDA, "fastq convert launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        c->launches++;
        const bool same = minlen == maxlen;
        r = run_filter_full(c, c->ws[q.stream], (const uint8_t *)q.slab.p, nullptr, same ? nullptr : (const uint32_t *)q.len.p, stride,
                            same ? maxlen : 0, q.n_rec, params, maxlen, minlen, (double *)q.ee.p, (int32_t *)q.ns.p, (uint8_t *)q.flags.p,
                            d_cnt, s, (const uint32_t *)q.marks.p);
        if (r) return r;
        const uint64_t m = q.n_rec;
        CU(cudaMemcpyAsync(q.h_res, q.ee.p, m * 8, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(q.h_res + m * 8, q.ns.p, m * 4, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(q.h_res + m * 12, q.len.p, m * 4, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(q.h_res + m * 16, q.flags.p, m, cudaMemcpyDeviceToHost, s));
        if (want_off) {
            CU(cudaMemcpyAsync(q.h_res + m * 17, q.soff.p, m * 4, cudaMemcpyDeviceToHost, s));
            CU(cudaMemcpyAsync(q.h_res + m * 21, q.qoff.p, m * 4, cudaMemcpyDeviceToHost, s));
        }
        CU(cudaMemcpyAsync(q.h_meta, q.meta.p, 12, cudaMemcpyDeviceToHost, s));     // a quality above 252 shows up here (words 0..2 only)
        CU(cudaEventRecord(q.ev, s));
        q.state = 2;
        return MOIRA_OK;
    };
    // stage 3: results -> the caller's arrays
    auto retire_chunk = [&](FqSlot &q) -> int {
        CU(cudaEventSynchronize(q.ev));
        if (q.h_meta[2] != 0xFFFFFFFFu)
            return fail(MOIRA_ERR_BAD_QUALITY, "a quality of record %u of the range at byte %llu is outside 0..252", q.h_meta[2],
                        (unsigned long long)q.pos);
        const uint64_t m = q.n_rec, at = q.first_read;
        memcpy(ee_out + at, q.h_res, m * 8);
        if (ns_out) memcpy(ns_out + at, q.h_res + m * 8, m * 4);
        if (lengths_out) memcpy(lengths_out + at, q.h_res + m * 12, m * 4);
        if (flags_out) memcpy(flags_out + at, q.h_res + m * 16, m);
        if (want_off) {   // chunk-relative 32-bit offsets -> positions in the caller's text
            const uint64_t base = q.pos - q.skip;
            const uint8_t *so = q.h_res + m * 17, *qo = q.h_res + m * 21;
            for (uint64_t i = 0; i < m; i++) {
                uint32_t a32, b32;
                memcpy(&a32, so + 4 * i, 4);
                memcpy(&b32, qo + 4 * i, 4);
                if (seq_off_out) seq_off_out[at + i] = base + a32;
                if (qual_off_out) qual_off_out[at + i] = base + b32;
            }
        }
        q.state = 0;
        return MOIRA_OK;
    };

    // The chunk plans (where each chunk ends, how many records and newlines it holds) come from a planner thread that
    // runs ahead of the copies: counting the newlines of a chunk on the host takes about as long as its H2D copy.
    // For a pageable source the same pass also copies the chunk into a ring of pinned buffers (one read of the text
    // for both); ring entry i % RING is free again once the H2D copy of chunk i - RING has completed (h2d_done).
    constexpr int RING = 6;
    if (!pinned_src)
        for (int r = 0; r < RING; r++)
            if ((rc = ensure_pinned(&c->fq_ring[r], &c->fq_ring_cap[r], std::min<uint64_t>(RANGE, text_bytes) + 64))) return rc;
    struct Plan { uint64_t pos, bytes, n_rec, n_nl; const uint8_t *staged; int rc; bool last; bool counted; };
    std::deque<Plan> plans;
    std::mutex pm;
    std::condition_variable pcv;
    bool stop = false;
    uint64_t h2d_done = 0;
    std::thread planner([&]() {
        uint64_t p = 0;
        for (uint64_t i = 0;; i++) {
            Plan pl{p, 0, 0, 0, nullptr, MOIRA_OK, false, true};
            uint8_t *ring = pinned_src ? nullptr : c->fq_ring[i % RING];
            if (ring) {
                std::unique_lock<std::mutex> lk(pm);
                pcv.wait(lk, [&] { return stop || i < h2d_done + RING - 1; });
                if (stop) return;
            }
            if (p >= text_bytes) pl.last = true;
            else {
                int ok = 0;
                if (guess_cuts) pl.rc = fastq_plan_chunk_fast(text, text_bytes, p, RANGE, ring, &pl.bytes, &ok);
                if (ok && !pl.rc) { pl.counted = false; pl.n_rec = 0xFFFFFFFFull; }
                else pl.rc = fastq_plan_chunk(text, text_bytes, p, RANGE, ring, &pl.bytes, &pl.n_rec, &pl.n_nl);
            }
            pl.staged = ring;
            if (pl.rc || pl.n_rec == 0 || (!pl.counted && pl.bytes == 0)) pl.last = true;
            {
                std::unique_lock<std::mutex> lk(pm);
                pcv.wait(lk, [&] { return stop || plans.size() < 4; });
                if (stop) return;
                plans.push_back(pl);
            }
            pcv.notify_all();
            if (pl.last) return;
            p += pl.bytes;
        }
    });
    auto next_plan = [&]() {
        std::unique_lock<std::mutex> lk(pm);
        pcv.wait(lk, [&] { return !plans.empty(); });
        Plan pl = plans.front();
        plans.pop_front();
        lk.unlock();
        pcv.notify_all();
        return pl;
    };
    struct PlannerJoin {   // whatever path leaves this function: stop and join the planner
        std::thread &t; std::mutex &m; std::condition_variable &cv; bool &stop;
        ~PlannerJoin() { { std::lock_guard<std::mutex> lk(m); stop = true; } cv.notify_all(); if (t.joinable()) t.join(); }
    } planner_join{planner, pm, pcv, stop};

    uint64_t pos = 0;
    int k = 0;
    rc = MOIRA_OK;
    const bool dbg = getenv("MOIRA_B200_FQ_DEBUG") != nullptr;
    double t_plan = 0, t_issue = 0, t_filter = 0, t_retire = 0;
    auto now = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    while (!rc) {
        FqSlot &q = c->fqd[k % 3];
        double t0 = now();
        if (q.state == 2 && (rc = retire_chunk(q))) break;          // chunk k - 3
        t_retire += now() - t0; t0 = now();
        // stage 1: plan, stage, copy, index
        const Plan pl = next_plan();
        if ((rc = pl.rc)) break;
        pos = pl.pos;
        const uint64_t bytes = pl.bytes, n_rec = pl.n_rec, n_nl = pl.n_nl;
        if (n_rec == 0) {
            if (pos >= text_bytes || pos + std::min<uint64_t>(RANGE, text_bytes - pos) >= text_bytes) break;   // a trailing partial record is ignored, as the reference does
            rc = fail(MOIRA_ERR_PARSE, "a FASTQ record is longer than the %llu-byte streaming range", (unsigned long long)RANGE);
            break;
        }
        t_plan += now() - t0; t0 = now();
        q.pos = pos; q.bytes = bytes; q.n_rec = n_rec; q.stream = k & 1;
        q.counted = pl.counted; q.final = pos + bytes >= text_bytes;
        // a pinned source is copied from its page boundary (unaligned DMA is several times slower); staging is aligned anyway
        q.skip = pinned_src ? (uint64_t)((uintptr_t)(text + pos) & 4095u) : 0;
        if (q.skip > pos) q.skip = 0;   // never read before the caller's buffer
        const uint64_t span = q.skip + bytes;
        cudaStream_t s = c->streams[q.stream];
        const uint32_t nb = fq_blocks(span);
        // sizes of the newline index and the record table: counted by the planner, or (guessed cut) an estimate the
        // kernels respect -- 8 bytes per line; a text with shorter lines fails the device's check and is counted
        const uint64_t nl_cap = pl.counted ? n_nl + 16 : span / 8 + 64;
        const uint64_t rec_cap = pl.counted ? n_rec : nl_cap / 4 + 1;
        const uint32_t extra_line = (q.final && bytes && text[pos + bytes - 1] != '\n') ? 1u : 0u;
        if (!q.ev) CU(cudaEventCreateWithFlags(&q.ev, cudaEventDisableTiming));
        if ((rc = ensure(q.text, span + 64)) || (rc = ensure(q.bcnt, (size_t)nb * 4 + 16)) || (rc = ensure(q.bstart, (size_t)nb * 4 + 16)) ||
            (rc = ensure(q.nl, (size_t)nl_cap * 4 + 64)) || (rc = ensure(q.soff, rec_cap * 4)) || (rc = ensure(q.qoff, rec_cap * 4)) ||
            (rc = ensure(q.len, rec_cap * 4)) || (rc = ensure(q.meta, 64)))
            break;
        if (!q.h_meta && cudaHostAlloc((void **)&q.h_meta, 64, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); rc = fail(MOIRA_ERR_NOMEM, "cudaHostAlloc failed"); break; }
        q.h_meta[0] = 0; q.h_meta[1] = 0xFFFFFFFFu; q.h_meta[2] = 0xFFFFFFFFu;
        q.h_meta[3] = q.h_meta[4] = q.h_meta[5] = q.h_meta[6] = q.h_meta[7] = 0;
        CU(cudaMemcpyAsync(q.meta.p, q.h_meta, 32, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(q.text.p, pinned_src ? (const void *)(text + pos - q.skip) : (const void *)pl.staged, span, cudaMemcpyHostToDevice, s));
        if (launch_fq_index((const uint8_t *)q.text.p, q.skip, span, (uint32_t *)q.bcnt.p, (uint32_t *)q.bstart.p, (uint32_t *)q.nl.p,
                            (uint32_t)std::min<uint64_t>(nl_cap, 0xFFFFFFFFull), s) ||
            launch_fq_records((const uint8_t *)q.text.p, q.skip, span, (const uint32_t *)q.nl.p, (const uint32_t *)q.bstart.p + nb, (uint32_t)n_rec,
                              extra_line, (uint32_t)std::min<uint64_t>(nl_cap, 0xFFFFFFFFull), (uint32_t)rec_cap, (uint32_t *)q.soff.p,
                              (uint32_t *)q.qoff.p, (uint32_t *)q.len.p, (uint32_t *)q.meta.p, s)) {
            rc = fail(MOIRA_ERR_CUDA, "fastq index launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        c->launches += 4;
        CU(cudaMemcpyAsync(q.h_meta, q.meta.p, 32, cudaMemcpyDeviceToHost, s));
        CU(cudaEventRecord(q.ev, s));
        q.state = 1;
        t_issue += now() - t0; t0 = now();
        // chunk k - 1 has had a whole iteration for its copy and index: filter it now
        if (k >= 1) {
            FqSlot &p1 = c->fqd[(k - 1) % 3];
            if (p1.state == 1 && (rc = filter_chunk(p1))) break;
            { std::lock_guard<std::mutex> lk(pm); h2d_done = (uint64_t)k; }   // the copy of chunk k - 1 is complete
            pcv.notify_all();
        }
        t_filter += now() - t0;
        k++;
    }
    if (dbg) fprintf(stderr, "[fq] chunks %d pinned %d plan %.2f issue %.2f filter %.2f retire %.2f ms\n", k, (int)pinned_src, t_plan, t_issue, t_filter, t_retire);
    for (int j = 0; j < 3 && !rc; j++) {                            // drain, oldest first
        FqSlot &q = c->fqd[(k + j) % 3];
        if (q.state == 1) rc = filter_chunk(q);
        if (!rc && q.state == 2) rc = retire_chunk(q);
    }
    if (rc) {
        char keep[sizeof(g_err)];
        memcpy(keep, g_err, sizeof(keep));
        cudaDeviceSynchronize();
        cudaGetLastError();
        memcpy(g_err, keep, sizeof(keep));
        return rc;
    }
    if (counters_out) {
        CU(cudaMemcpy(counters_out, d_cnt, MOIRA_N_COUNTERS * 8, cudaMemcpyDeviceToHost));
        note_escalations(c, params, counters_out);
    }
    if (labels_out && n_assigned) {
        // every chunk has left its sequences in the store: one exact dereplication over all of them
        CU(cudaStreamSynchronize(c->streams[1]));
        if ((rc = run_dedup(c, (const uint8_t *)c->dd_store.p, (const uint64_t *)c->dd_seq_abs.p, (const uint32_t *)c->dd_seq_eff.p, 0, 0,
                            n_assigned, 0, (uint32_t *)c->dd_labels.p, c->streams[0])))
            return rc;
        CU(cudaMemcpyAsync(labels_out, c->dd_labels.p, n_assigned * 4, cudaMemcpyDeviceToHost, c->streams[0]));
        CU(cudaStreamSynchronize(c->streams[0]));
    }
    *n_reads_out = n_assigned;
    return MOIRA_OK;
}

}  // namespace

// FASTQ text -> decisions, streaming: ranges of ~64 MB of text are parsed (all host threads) into
// pinned slabs and submitted asynchronously, so parsing range k+1 overlaps the H2D copy, the kernels
// and the D2H copy of range k.
int moira_filter_fastq(moira_ctx *c, const char *text, uint64_t text_bytes, int fastq_offset, int lower_n_ambiguous,
                       const moira_params *params, uint64_t max_reads, double *ee_out, int32_t *ns_out,
                       uint8_t *flags_out, uint32_t *lengths_out, uint64_t *counters_out, uint64_t *n_reads_out)
{
    return moira_filter_fastq_ex(c, text, text_bytes, fastq_offset, lower_n_ambiguous, params, max_reads, ee_out, ns_out, flags_out,
                                 lengths_out, nullptr, nullptr, nullptr, counters_out, n_reads_out);
}

int moira_filter_fastq_ex(moira_ctx *c, const char *text, uint64_t text_bytes, int fastq_offset, int lower_n_ambiguous,
                          const moira_params *params, uint64_t max_reads, double *ee_out, int32_t *ns_out,
                          uint8_t *flags_out, uint32_t *lengths_out, uint64_t *seq_off_out, uint64_t *qual_off_out,
                          uint32_t *labels_out, uint64_t *counters_out, uint64_t *n_reads_out)
{
    if (!c || (!text && text_bytes) || !ee_out || !n_reads_out) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    if ((seq_off_out || qual_off_out || labels_out) && !c->device_parse)
        return fail(MOIRA_ERR_BAD_ARG, "record offsets / labels need the device parser (MOIRA_B200_HOST_PARSE is set)");
    int rc = check_params(params);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    if (params->slab_format != MOIRA_SLAB_Q8) return fail(MOIRA_ERR_BAD_ARG, "slab_format does not apply to FASTQ text");
    if (c->device_parse) {
        // first with chunk cuts guessed from the text's local structure (no host pass over the text); if the device
        // finds a chunk that does not hold whole records, once more with the cuts counted on the host
        rc = filter_fastq_device(c, text, text_bytes, fastq_offset, lower_n_ambiguous, params, max_reads, ee_out, ns_out, flags_out,
                                 lengths_out, seq_off_out, qual_off_out, labels_out, counters_out, n_reads_out, c->fq_guess_cuts != 0);
        if (rc == FQ_RETRY_COUNTED)
            rc = filter_fastq_device(c, text, text_bytes, fastq_offset, lower_n_ambiguous, params, max_reads, ee_out, ns_out, flags_out,
                                     lengths_out, seq_off_out, qual_off_out, labels_out, counters_out, n_reads_out, false);
        return rc;
    }
    constexpr int SLOTS = 3;
    const uint64_t RANGE = 64ull << 20;
    struct Slot { int ticket = -1; uint64_t counters[MOIRA_N_COUNTERS]; };
    Slot slots[SLOTS];
    uint64_t total[MOIRA_N_COUNTERS] = {0};
    auto retire = [&](Slot &s) -> int {
        if (s.ticket < 0) return MOIRA_OK;
        int r = moira_wait(c, s.ticket);
        s.ticket = -1;
        for (int i = 0; i < MOIRA_N_COUNTERS; i++) total[i] += s.counters[i];
        return r;
    };
    uint64_t pos = 0, n_done = 0;
    int k = 0;
    while (pos < text_bytes) {
        Slot &s = slots[k % SLOTS];
        if ((rc = retire(s))) break;                       // its pinned buffers are free again
        StreamBuf &b = c->fq[k % SLOTS];
        const uint64_t len = std::min<uint64_t>(RANGE, text_bytes - pos);
        const int final_range = pos + len >= text_bytes;
        uint64_t n = 0, cap = 0, consumed = 0;
        // sizing (newline count only), then the single parsing pass into the slot's pinned buffers
        rc = parse_fastq_range(text + pos, len, fastq_offset, lower_n_ambiguous, nullptr, 0, nullptr, nullptr, nullptr,
                               nullptr, nullptr, nullptr, 0, &n, &cap, nullptr, final_range);
        if (rc) break;
        if (n_done + n > max_reads) { rc = fail(MOIRA_ERR_BAD_ARG, "more than max_reads = %llu records", (unsigned long long)max_reads); break; }
        if (n == 0) {
            if (final_range) break;
            rc = fail(MOIRA_ERR_PARSE, "a FASTQ record is longer than the %llu-byte streaming range", (unsigned long long)RANGE);
            break;
        }
        if ((rc = b.ensure(cap + 16, n))) break;
        uint64_t used = 0;
        rc = parse_fastq_range(text + pos, len, fastq_offset, lower_n_ambiguous, b.slab, b.slab_cap, b.offsets, b.lengths,
                               nullptr, nullptr, nullptr, nullptr, n, &n, &used, &consumed, final_range);
        if (rc) break;
        if (lengths_out) memcpy(lengths_out + n_done, b.lengths, n * sizeof(uint32_t));
        rc = moira_submit(c, b.slab, used ? used : 16, b.offsets, b.lengths, n, params, ee_out + n_done,
                          ns_out ? ns_out + n_done : nullptr, flags_out ? flags_out + n_done : nullptr, s.counters, &s.ticket);
        if (rc) break;
        n_done += n;
        pos += consumed ? consumed : len;
        k++;
    }
    for (auto &s : slots) { int r = retire(s); if (!rc) rc = r; }
    if (rc) { cudaDeviceSynchronize(); return rc; }
    if (counters_out) memcpy(counters_out, total, sizeof(total));
    *n_reads_out = n_done;
    return MOIRA_OK;
}

int moira_calculate_errors_PB(moira_ctx *c, const char *contig, const int32_t *quals, uint64_t length, double alpha,
                              double *ee_out, int32_t *ns_out)
{
    if (!c || !contig || (!quals && length) || !ee_out || !ns_out) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    if (alpha <= 0 || alpha >= 1) return fail(MOIRA_ERR_BAD_ALPHA, "Alpha must be between 0 and 1");          // bernoullimodule.c:79-83
    if (strlen(contig) != length) return fail(MOIRA_ERR_LENGTH_MISMATCH, "contig and contig_quals must have the same length");  // :85-90
    if (length > 0xFFFFFFF0ull) return fail(MOIRA_ERR_BAD_ARG, "read too long");
    CU(cudaSetDevice(c->device));
    const size_t padded = ((size_t)length + 15) & ~(size_t)15;
    if (padded + 16 > c->one_cap) {
        if (c->one_slab) cudaFreeHost(c->one_slab);
        c->one_slab = nullptr; c->one_cap = 0;
        CU(cudaHostAlloc((void **)&c->one_slab, padded + 4096, cudaHostAllocDefault));
        c->one_cap = padded + 4096;
    }
    if (!c->one_out) CU(cudaHostAlloc((void **)&c->one_out, 64, cudaHostAllocDefault));
    for (uint64_t i = 0; i < length; i++) {
        const char ch = contig[i];
        if (ch == 78) c->one_slab[i] = 0xFF;                       // 'N'  bernoullimodule.c:196
        else if (ch == 110) c->one_slab[i] = 0xFE;                 // 'n'
        else {
            const int32_t q = quals[i];
            if (q < 0 || q > 0xFC) return fail(MOIRA_ERR_BAD_QUALITY, "quality %d at position %llu is outside 0..252", q, (unsigned long long)i);
            c->one_slab[i] = (uint8_t)q;                           // 0 is read as 1 by the table (:104-107)
        }
    }
    for (size_t i = length; i < padded; i++) c->one_slab[i] = 0xFD;
    moira_params p;
    moira_params_default(&p);
    p.alpha = alpha;
    p.exact_ee = 1;
    p.ee_output = MOIRA_EE_RAW;
    uint64_t off = 0;
    uint32_t len = (uint32_t)length;
    double *ee = c->one_out;
    int32_t *ns = reinterpret_cast<int32_t *>(c->one_out + 1);
    uint8_t *fl = reinterpret_cast<uint8_t *>(c->one_out + 2);
    int rc = moira_filter_batch(c, c->one_slab, padded ? padded : 16, &off, &len, 1, &p, ee, ns, fl, nullptr);
    if (rc) return rc;
    if (*fl & MOIRA_FLAG_NUMERIC) return fail(MOIRA_ERR_UNRESOLVED, "cumulative probability never exceeded 1 - alpha");
    *ee_out = *ee;
    *ns_out = *ns;
    return MOIRA_OK;
}


// ---- paired-end contig construction --------------------------------------------------------------
void moira_contig_params_default(moira_contig_params *p)
{
    if (!p) return;
    p->match = 1; p->mismatch = -1; p->gap = -2;      // moira.py:630-634
    p->insert = 20; p->deltaq = 6;                    // moira.py:638-640
    p->consensus = MOIRA_CONSENSUS_BEST;              // moira.py:642
    p->qscore_cap = 40;                               // moira.py:645
    p->trim_overlap = 0;
}

static int check_contig_params(const moira_contig_params *p)
{
    if (!p) return fail(MOIRA_ERR_BAD_ARG, "contig params is NULL");
    if (p->consensus < MOIRA_CONSENSUS_BEST || p->consensus > MOIRA_CONSENSUS_POSTERIOR)
        return fail(MOIRA_ERR_BAD_ARG, "consensus_qscore must be \"best\", \"sum\" or \"posterior\".");   // moira.py:1405-1406
    if (p->insert <= 0) return fail(MOIRA_ERR_BAD_ARG, "insert must be a positive integer");              // moira.py:1411-1412
    if (p->deltaq <= 0) return fail(MOIRA_ERR_BAD_ARG, "deltaq must be a positive integer");              // moira.py:1413-1414
    if (p->qscore_cap < 0) return fail(MOIRA_ERR_BAD_ARG, "qscore_cap must be a non-negative integer");   // moira.py:1415-1416
    return MOIRA_OK;
}

// Posterior consensus qualities for every pair of input qualities, with the host libm and the reference's
// own expressions (moira.py:1391-1395, 1522-1524, 1547-1553): [0, 65536) agreeing bases, [65536, 131072)
// disagreeing bases indexed [better * 256 + worse].
static int ensure_post_tables(moira_ctx *c)
{
    if (c->post_ready) return MOIRA_OK;
    int rc = ensure(c->post, 2 * 65536 * sizeof(int16_t));
    if (rc) return rc;
    std::vector<int16_t> t(2 * 65536);
    auto qual2prob = [](int q) { volatile double p = pow(10, (q / (-10.0))); return (double)p; };
    auto prob2qual = [](double prob) { const double v = floor(-10 * log10(prob)); return (int16_t)(v > 32000 ? 32000 : (v < -32000 ? -32000 : v)); };
    for (int a = 0; a < 256; a++)
        for (int b = 0; b < 256; b++) {
            {
                const double p1 = qual2prob(a), p2 = qual2prob(b);
                volatile double num = p1 * p2 / 3;
                volatile double den = 1 - p1 - p2 + (4 * p1 * p2 / 3);
                volatile double post = num / den;
                t[a * 256 + b] = prob2qual(post);
            }
            if (a > b) {   // a: quality of the base that is kept
                const double p1 = qual2prob(a), p2 = qual2prob(b);
                volatile double num = p1 * (1 - p2 / 3);
                volatile double den = p1 + p2 - (4 * p1 * p2 / 3);
                volatile double post = num / den;
                t[65536 + a * 256 + b] = prob2qual(post);
            } else t[65536 + a * 256 + b] = 2;
        }
    CU(cudaMemcpy(c->post.p, t.data(), t.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
    c->post_ready = true;
    return MOIRA_OK;
}

static void fill_contig_args(ContigArgs &a, const moira_ctx *c, const moira_contig_params *p, int lower_n)
{
    memset(&a, 0, sizeof(a));
    a.match = p->match; a.mismatch = p->mismatch; a.gap = p->gap; a.insert = p->insert; a.deltaq = p->deltaq;
    a.consensus = p->consensus; a.qscore_cap = p->qscore_cap; a.trim_overlap = p->trim_overlap ? 1 : 0; a.lower_n = lower_n ? 1 : 0;
    a.post_match = (const int16_t *)c->post.p;
    a.post_mis = a.post_match ? a.post_match + 65536 : nullptr;
}

static int filter_pairs_impl(moira_ctx *c, const char *fwd_seq, uint64_t fwd_bytes, const uint8_t *fwd_qual, uint64_t fwd_qbytes,
                       const uint64_t *fwd_off, const uint64_t *fwd_qoff, const uint32_t *fwd_len, const char *rev_seq,
                       uint64_t rev_bytes, const uint8_t *rev_qual, uint64_t rev_qbytes, const uint64_t *rev_off,
                       const uint64_t *rev_qoff, const uint32_t *rev_len,
                       int qual_base, uint64_t n, const moira_contig_params *cp, int lower_n,
                       const moira_params *fp, uint64_t out_stride, char *contig_seq, uint8_t *contig_qual, uint32_t *contig_len,
                       int32_t *overlap, int32_t *gaps, int32_t *mismatches, uint8_t *status, double *ee_out, int32_t *ns_out,
                       uint8_t *flags_out, uint64_t *counters_out)
{
    if (!c) return fail(MOIRA_ERR_BAD_ARG, "ctx is NULL");
    int rc = check_contig_params(cp);
    if (rc) return rc;
    if (fp && (rc = check_params(fp))) return rc;
    if (fp && fp->slab_format != MOIRA_SLAB_Q8) return fail(MOIRA_ERR_BAD_ARG, "slab_format does not apply to read pairs");
    if (counters_out) memset(counters_out, 0, MOIRA_N_COUNTERS * sizeof(uint64_t));
    if (n == 0) return MOIRA_OK;
    if (!fwd_seq || !fwd_qual || !fwd_off || !fwd_len || !rev_seq || !rev_qual || !rev_off || !rev_len || !contig_seq ||
        !contig_qual || !contig_len || !status || (fp && !ee_out))
        return fail(MOIRA_ERR_BAD_ARG, "NULL buffer");
    CU(cudaSetDevice(c->device));
    uint32_t max_l1 = 0, max_l2 = 0;
    if (!fwd_qoff) fwd_qoff = fwd_off;
    if (!rev_qoff) rev_qoff = rev_off;
    for (uint64_t r = 0; r < n; r++) {
        if (fwd_off[r] + fwd_len[r] > fwd_bytes || rev_off[r] + rev_len[r] > rev_bytes || fwd_qoff[r] + fwd_len[r] > fwd_qbytes ||
            rev_qoff[r] + rev_len[r] > rev_qbytes)
            return fail(MOIRA_ERR_BAD_ARG, "pair %llu extends past its arrays", (unsigned long long)r);
        max_l1 = std::max(max_l1, fwd_len[r]);
        max_l2 = std::max(max_l2, rev_len[r]);
    }
    if (max_l1 > 4096) return fail(MOIRA_ERR_BAD_ARG, "forward reads longer than 4096 bases are not supported");
    if (max_l1 == 0) max_l1 = 1;
    if (max_l2 == 0) max_l2 = 1;
    if (max_l2 > 1024) max_l2 = 1024;   // longer reverse reads get MOIRA_PAIR_TOO_LONG
    if (out_stride < (uint64_t)max_l1 + max_l2 || (out_stride & 15u))
        return fail(MOIRA_ERR_BAD_ARG, "out_stride must be a multiple of 16 and at least max(fwd_len) + max(rev_len) = %u", max_l1 + max_l2);
    if (cp->consensus == MOIRA_CONSENSUS_POSTERIOR && (rc = ensure_post_tables(c))) return rc;
    const size_t warps = (size_t)c->sm_count * CONTIG_WARPS_PER_CTA;
    const size_t tw = contig_trace_words_per_warp(max_l1, max_l2);
    if ((rc = ensure(c->trace, warps * tw * sizeof(uint32_t)))) return rc;
    if (fp && (rc = ensure(c->pair_counters, MOIRA_N_COUNTERS * 8))) return rc;
    uint64_t *d_cnt = (uint64_t *)c->pair_counters.p;
    if (fp) CU(cudaMemsetAsync(d_cnt, 0, MOIRA_N_COUNTERS * 8, c->streams[0]));
    if (fp) { CU(cudaEventRecord(c->meta_ready, c->streams[0])); CU(cudaStreamWaitEvent(c->streams[1], c->meta_ready, 0)); }

    // Chunks of pairs alternate between the two streams: the copies of one chunk overlap the kernels of the other.
    constexpr uint64_t PAIR_CHUNK = 1u << 16;
    auto flush = [&](PairBufs &b) {
        const uint64_t st0 = b.pend_start, m = b.pend_n;
        if (!m) return;
        const uint8_t *st = b.stage;
        if (fp) {
            memcpy(ee_out + st0, st, m * 8);
            if (ns_out) memcpy(ns_out + st0, st + m * 8, m * 4);
            if (flags_out) memcpy(flags_out + st0, st + m * 28, m);
        }
        memcpy(contig_len + st0, st + m * 12, m * 4);
        if (overlap) memcpy(overlap + st0, st + m * 16, m * 4);
        if (gaps) memcpy(gaps + st0, st + m * 20, m * 4);
        if (mismatches) memcpy(mismatches + st0, st + m * 24, m * 4);
        memcpy(status + st0, st + m * 29, m);
        if (b.out_staged) {
            parallel_memcpy(contig_seq + st0 * out_stride, b.out_stage, m * out_stride);
            parallel_memcpy(contig_qual + st0 * out_stride, b.out_stage + m * out_stride, m * out_stride);
        }
        b.pend_n = 0;
    };
    // pageable inputs / outputs are staged through pinned memory (see PairBufs)
    auto pinned = [](const void *p) {
        cudaPointerAttributes attr;
        const bool yes = cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        cudaGetLastError();
        return yes;
    };
    const bool stage_in = !(pinned(fwd_seq) && pinned(fwd_qual) && pinned(rev_seq) && pinned(rev_qual));
    const bool stage_out = !(pinned(contig_seq) && pinned(contig_qual));
    c->pb[0].pend_n = c->pb[1].pend_n = 0;
    int ci = 0;
    for (uint64_t start = 0; start < n; start += PAIR_CHUNK, ci++) {
        const uint64_t cn = std::min<uint64_t>(PAIR_CHUNK, n - start);
        PairBufs &b = c->pb[ci & 1];
        cudaStream_t s = c->streams[ci & 1];
        if (ci >= 2) { CU(cudaStreamSynchronize(s)); flush(b); }   // chunk ci - 2 is complete: hand its results over, reuse its buffers
        if (cn * PAIR_STAGE_BYTES > b.stage_cap) {
            if (b.stage) cudaFreeHost(b.stage);
            b.stage = nullptr; b.stage_cap = 0;
            if (cudaHostAlloc((void **)&b.stage, PAIR_CHUNK * PAIR_STAGE_BYTES, cudaHostAllocDefault) != cudaSuccess) {
                cudaGetLastError();
                return fail(MOIRA_ERR_NOMEM, "cudaHostAlloc of the result staging buffer failed");
            }
            b.stage_cap = PAIR_CHUNK * PAIR_STAGE_BYTES;
        }
        // byte ranges of this chunk in the four host arrays; bases and qualities that live in ONE buffer (the
        // FASTQ text) travel once
        uint64_t f0 = ~0ull, f1 = 0, r0 = ~0ull, r1 = 0, fq0 = ~0ull, fq1 = 0, rq0 = ~0ull, rq1 = 0;
        for (uint64_t r = start; r < start + cn; r++) {
            f0 = std::min(f0, fwd_off[r]); f1 = std::max(f1, fwd_off[r] + fwd_len[r]);
            r0 = std::min(r0, rev_off[r]); r1 = std::max(r1, rev_off[r] + rev_len[r]);
            fq0 = std::min(fq0, fwd_qoff[r]); fq1 = std::max(fq1, fwd_qoff[r] + fwd_len[r]);
            rq0 = std::min(rq0, rev_qoff[r]); rq1 = std::max(rq1, rev_qoff[r] + rev_len[r]);
        }
        const bool f_shared = (const void *)fwd_seq == (const void *)fwd_qual, r_shared = (const void *)rev_seq == (const void *)rev_qual;
        if (f_shared) { f0 = fq0 = std::min(f0, fq0); f1 = fq1 = std::max(f1, fq1); }
        if (r_shared) { r0 = rq0 = std::min(r0, rq0); r1 = rq1 = std::max(r1, rq1); }
        if ((rc = ensure(b.fseq, f1 - f0 + 16)) || (!f_shared && (rc = ensure(b.fqual, fq1 - fq0 + 16))) ||
            (rc = ensure(b.rseq, r1 - r0 + 16)) || (!r_shared && (rc = ensure(b.rqual, rq1 - rq0 + 16))) ||
            (rc = ensure(b.foff, cn * 16)) || (rc = ensure(b.roff, cn * 16)) ||
            (rc = ensure(b.flen, cn * 4)) || (rc = ensure(b.rlen, cn * 4)) || (rc = ensure(b.cseq, cn * out_stride)) ||
            (rc = ensure(b.cqual, cn * out_stride)) || (rc = ensure(b.slab, cn * out_stride + 256)) || (rc = ensure(b.clen, cn * 4)) ||
            (rc = ensure(b.overlap, cn * 4)) || (rc = ensure(b.gaps, cn * 4)) || (rc = ensure(b.mism, cn * 4)) ||
            (rc = ensure(b.status, cn)) || (rc = ensure(b.ee, cn * 8)) || (rc = ensure(b.ns, cn * 4)) || (rc = ensure(b.flags, cn)))
            return rc;
        if (stage_in) {
            const size_t need = (f1 - f0) + (f_shared ? 0 : fq1 - fq0) + (r1 - r0) + (r_shared ? 0 : rq1 - rq0) + 64;
            if ((rc = ensure_host(&b.in_stage, &b.in_cap, need))) return rc;
            uint8_t *w = b.in_stage;
            auto ship = [&](void *dst, const void *src, size_t bytes) {
                parallel_memcpy(w, src, bytes);
                const cudaError_t e = cudaMemcpyAsync(dst, w, bytes, cudaMemcpyHostToDevice, s);
                w += (bytes + 15) & ~(size_t)15;
                return e;
            };
            CU(ship(b.fseq.p, fwd_seq + f0, f1 - f0));
            if (!f_shared) CU(ship(b.fqual.p, fwd_qual + fq0, fq1 - fq0));
            CU(ship(b.rseq.p, rev_seq + r0, r1 - r0));
            if (!r_shared) CU(ship(b.rqual.p, rev_qual + rq0, rq1 - rq0));
        } else {
            CU(cudaMemcpyAsync(b.fseq.p, fwd_seq + f0, f1 - f0, cudaMemcpyHostToDevice, s));
            if (!f_shared) CU(cudaMemcpyAsync(b.fqual.p, fwd_qual + fq0, fq1 - fq0, cudaMemcpyHostToDevice, s));
            CU(cudaMemcpyAsync(b.rseq.p, rev_seq + r0, r1 - r0, cudaMemcpyHostToDevice, s));
            if (!r_shared) CU(cudaMemcpyAsync(b.rqual.p, rev_qual + rq0, rq1 - rq0, cudaMemcpyHostToDevice, s));
        }
        uint64_t *d_foff = (uint64_t *)b.foff.p, *d_roff = (uint64_t *)b.roff.p;   // [0, cn): bases, [cn, 2 cn): qualities
        CU(cudaMemcpyAsync(d_foff, fwd_off + start, cn * 8, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(d_foff + cn, fwd_qoff + start, cn * 8, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(d_roff, rev_off + start, cn * 8, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(d_roff + cn, rev_qoff + start, cn * 8, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(b.flen.p, fwd_len + start, cn * 4, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(b.rlen.p, rev_len + start, cn * 4, cudaMemcpyHostToDevice, s));
        ContigArgs a;
        fill_contig_args(a, c, cp, lower_n);
        // offsets stay absolute: the device pointers are shifted by the chunk's first byte
        a.fseq = (const char *)b.fseq.p - f0;
        a.fqual = f_shared ? (const uint8_t *)b.fseq.p - f0 : (const uint8_t *)b.fqual.p - fq0;
        a.rseq = (const char *)b.rseq.p - r0;
        a.rqual = r_shared ? (const uint8_t *)b.rseq.p - r0 : (const uint8_t *)b.rqual.p - rq0;
        a.foff = d_foff; a.fqoff = d_foff + cn; a.flen = (const uint32_t *)b.flen.p;
        a.roff = d_roff; a.rqoff = d_roff + cn; a.rlen = (const uint32_t *)b.rlen.p;
        a.qual_base = qual_base;
        a.n_pairs = cn;
        // every chunk uses the whole trace scratch: chunks on the two streams must not run their contig kernels at once
        if (ci >= 1) CU(cudaStreamWaitEvent(s, c->tickets[0].done[(ci - 1) & 1], 0));   // recorded behind that chunk's kernels
        a.trace = (uint32_t *)c->trace.p; a.trace_words_per_warp = tw;
        a.out_stride = out_stride;
        a.cseq = (char *)b.cseq.p; a.cqual = (uint8_t *)b.cqual.p; a.slab = fp ? (uint8_t *)b.slab.p : nullptr;
        a.clen = (uint32_t *)b.clen.p; a.overlap = (int32_t *)b.overlap.p; a.gaps = (int32_t *)b.gaps.p; a.mism = (int32_t *)b.mism.p;
        a.status = (uint8_t *)b.status.p;
        a.max_l1 = max_l1; a.max_l2 = max_l2;
        LaunchCfg cfg{c->sm_count, s};
        const bool ctimed = c->timing && c->n_ctimed < MAX_TIMED;
        if (ctimed) CU(cudaEventRecord(c->ct0[c->n_ctimed], s));
        const int lr = launch_contigs(a, false, cfg);
        if (lr) return fail(lr == -2 ? MOIRA_ERR_BAD_ARG : MOIRA_ERR_CUDA, "contig kernel launch failed: %s", lr == -2 ? "reads too long" : cudaGetErrorString(cudaGetLastError()));
        if (ctimed) { CU(cudaEventRecord(c->ct1[c->n_ctimed], s)); c->n_ctimed++; }
        c->launches++;
        if (fp) {
            const uint32_t cap = (uint32_t)std::min<uint64_t>(out_stride, 0xFFFFFFF0u);
            rc = run_filter_full(c, c->ws[ci & 1], (const uint8_t *)b.slab.p, nullptr, (const uint32_t *)b.clen.p, out_stride, 0, cn, fp,
                                 cap, 0, (double *)b.ee.p, (int32_t *)b.ns.p, (uint8_t *)b.flags.p, d_cnt, s);
            if (rc) return rc;
        }
        // The next chunk's contig kernel starts behind this chunk's kernels (a persistent contig grid in between would
        // hold the filter back and with it the copies); only the D2H copies below overlap it.
        CU(cudaEventRecord(c->tickets[0].done[ci & 1], s));
        // small results -> pinned staging (sections of cn entries each), big rows straight to the caller
        uint8_t *st = b.stage;
        if (fp) {
            CU(cudaMemcpyAsync(st, b.ee.p, cn * 8, cudaMemcpyDeviceToHost, s));
            CU(cudaMemcpyAsync(st + cn * 8, b.ns.p, cn * 4, cudaMemcpyDeviceToHost, s));
            CU(cudaMemcpyAsync(st + cn * 28, b.flags.p, cn, cudaMemcpyDeviceToHost, s));
        }
        CU(cudaMemcpyAsync(st + cn * 12, b.clen.p, cn * 4, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(st + cn * 16, b.overlap.p, cn * 4, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(st + cn * 20, b.gaps.p, cn * 4, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(st + cn * 24, b.mism.p, cn * 4, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(st + cn * 29, b.status.p, cn, cudaMemcpyDeviceToHost, s));
        b.pend_start = start; b.pend_n = cn;
        b.out_staged = stage_out;
        if (stage_out) {
            if ((rc = ensure_host(&b.out_stage, &b.out_cap, 2 * cn * out_stride))) return rc;
            CU(cudaMemcpyAsync(b.out_stage, b.cseq.p, cn * out_stride, cudaMemcpyDeviceToHost, s));
            CU(cudaMemcpyAsync(b.out_stage + cn * out_stride, b.cqual.p, cn * out_stride, cudaMemcpyDeviceToHost, s));
        } else {
            CU(cudaMemcpyAsync(contig_seq + start * out_stride, b.cseq.p, cn * out_stride, cudaMemcpyDeviceToHost, s));
            CU(cudaMemcpyAsync(contig_qual + start * out_stride, b.cqual.p, cn * out_stride, cudaMemcpyDeviceToHost, s));
        }
    }
    CU(cudaStreamSynchronize(c->streams[1]));
    CU(cudaStreamSynchronize(c->streams[0]));
    flush(c->pb[0]);
    flush(c->pb[1]);
    if (fp && counters_out) CU(cudaMemcpy(counters_out, d_cnt, MOIRA_N_COUNTERS * 8, cudaMemcpyDeviceToHost));
    CU(cudaGetLastError());
    return MOIRA_OK;
}

int moira_filter_pairs(moira_ctx *c, const char *fwd_seq, uint64_t fwd_bytes, const uint8_t *fwd_qual, uint64_t fwd_qbytes,
                       const uint64_t *fwd_off, const uint64_t *fwd_qoff, const uint32_t *fwd_len, const char *rev_seq,
                       uint64_t rev_bytes, const uint8_t *rev_qual, uint64_t rev_qbytes, const uint64_t *rev_off,
                       const uint64_t *rev_qoff, const uint32_t *rev_len,
                       int qual_base, uint64_t n, const moira_contig_params *cp, int lower_n,
                       const moira_params *fp, uint64_t out_stride, char *contig_seq, uint8_t *contig_qual, uint32_t *contig_len,
                       int32_t *overlap, int32_t *gaps, int32_t *mismatches, uint8_t *status, double *ee_out, int32_t *ns_out,
                       uint8_t *flags_out, uint64_t *counters_out)
{
    const int rc = filter_pairs_impl(c, fwd_seq, fwd_bytes, fwd_qual, fwd_qbytes, fwd_off, fwd_qoff, fwd_len, rev_seq, rev_bytes, rev_qual,
                                     rev_qbytes, rev_off, rev_qoff, rev_len, qual_base, n, cp, lower_n, fp, out_stride, contig_seq,
                                     contig_qual, contig_len, overlap, gaps, mismatches, status, ee_out, ns_out, flags_out, counters_out);
    if (rc && c) {   // nothing may still be writing into the caller's arrays when the error is returned
        for (cudaStream_t s : c->streams) if (s) cudaStreamSynchronize(s);
        cudaGetLastError();
    }
    return rc;
}

// One pair through the contig kernel (the single-call entry points): inputs, the two output rows and the
// scalar results live in one device allocation.
struct SingleOut {
    std::vector<uint8_t> row_a, row_b;   // contig bases / qualities, or the two aligned strings
    uint32_t clen = 0;
    int32_t overlap = 0, gaps = 0, mism = 0, alen = 0;
    long long score = 0;
    uint8_t status = 0;
};

static int single_pair(moira_ctx *c, ContigArgs &a, const std::string &f, const std::vector<uint8_t> &fq, const std::string &r,
                       const std::vector<uint8_t> &rq, bool alignment_only, SingleOut &out)
{
    const size_t L1 = f.size(), L2 = r.size();
    const size_t cap = (L1 + L2 + 16) & ~(size_t)15;
    const size_t in_bytes = (2 * (L1 + L2) + 15) & ~(size_t)15;
    const size_t meta = 32, tail_bytes = 48;
    int rc;
    if ((rc = ensure(c->pb[0].fseq, in_bytes + meta + 2 * cap + tail_bytes))) return rc;
    uint8_t *d = (uint8_t *)c->pb[0].fseq.p;
    std::vector<uint8_t> h(in_bytes + meta, 0);
    memcpy(h.data(), f.data(), L1);
    memcpy(h.data() + L1, fq.data(), L1);
    memcpy(h.data() + 2 * L1, r.data(), L2);
    memcpy(h.data() + 2 * L1 + L2, rq.data(), L2);
    const uint32_t l1 = (uint32_t)L1, l2 = (uint32_t)L2;
    memcpy(h.data() + in_bytes + 8, &l1, 4);     // [in_bytes, +8): offset 0 of both reads
    memcpy(h.data() + in_bytes + 12, &l2, 4);
    cudaStream_t s = c->streams[0];
    CU(cudaMemcpyAsync(d, h.data(), h.size(), cudaMemcpyHostToDevice, s));
    a.fseq = (const char *)d; a.fqual = d + L1; a.rseq = (const char *)d + 2 * L1; a.rqual = d + 2 * L1 + L2;
    a.foff = a.roff = (const uint64_t *)(d + in_bytes);
    a.flen = (const uint32_t *)(d + in_bytes + 8); a.rlen = (const uint32_t *)(d + in_bytes + 12);
    a.n_pairs = 1;
    a.rev_direct = 1;
    a.max_l1 = std::max<uint32_t>(1, l1); a.max_l2 = std::max<uint32_t>(1, l2);
    if (!contig_columns_per_lane(a.max_l2) || a.max_l1 > 4096) return fail(MOIRA_ERR_BAD_ARG, "sequences longer than 4096 x 1024 bases are not supported");
    const size_t tw = contig_trace_words_per_warp(a.max_l1, a.max_l2);
    if ((rc = ensure(c->trace, (size_t)c->sm_count * CONTIG_WARPS_PER_CTA * tw * sizeof(uint32_t)))) return rc;
    a.trace = (uint32_t *)c->trace.p; a.trace_words_per_warp = tw;
    if (alignment_only) {
        if ((rc = ensure(c->hbuf, contig_hbuf_words(a.max_l1, a.max_l2) * sizeof(int32_t)))) return rc;
        a.hbuf = (int32_t *)c->hbuf.p;
    }
    uint8_t *o = d + in_bytes + meta;            // row A | row B | scalars
    uint8_t *tail = o + 2 * cap;
    a.out_stride = cap;
    a.cseq = (char *)o; a.cqual = o + cap; a.slab = nullptr;
    if (alignment_only) { a.al1 = (char *)o; a.al2 = (char *)o + cap; }
    a.clen = (uint32_t *)tail; a.overlap = (int32_t *)(tail + 4); a.gaps = (int32_t *)(tail + 8); a.mism = (int32_t *)(tail + 12);
    a.alen = (int32_t *)(tail + 16); a.score = (long long *)(tail + 24); a.status = tail + 32;
    LaunchCfg cfg{c->sm_count, s};
    const int lr = launch_contigs(a, alignment_only, cfg);
    if (lr) return fail(lr == -2 ? MOIRA_ERR_BAD_ARG : MOIRA_ERR_CUDA, "contig kernel launch failed");
    c->launches++;
    std::vector<uint8_t> back(2 * cap + tail_bytes);
    CU(cudaMemcpyAsync(back.data(), o, back.size(), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    CU(cudaGetLastError());
    out.row_a.assign(back.begin(), back.begin() + cap);
    out.row_b.assign(back.begin() + cap, back.begin() + 2 * cap);
    const uint8_t *t = back.data() + 2 * cap;
    memcpy(&out.clen, t, 4); memcpy(&out.overlap, t + 4, 4); memcpy(&out.gaps, t + 8, 4); memcpy(&out.mism, t + 12, 4);
    memcpy(&out.alen, t + 16, 4); memcpy(&out.score, t + 24, 8);
    out.status = t[32];
    return MOIRA_OK;
}

int moira_nw_align(moira_ctx *c, const char *seq_1, const char *seq_2, int match, int mismatch, int gap, char *aligned_1,
                   char *aligned_2, uint64_t *aligned_len, int64_t *score)
{
    if (!c || !seq_1 || !seq_2 || !aligned_1 || !aligned_2 || !aligned_len) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    const std::string f(seq_1), r(seq_2);
    aligned_1[0] = aligned_2[0] = 0;
    *aligned_len = 0;
    if (score) *score = 0;
    if (f.empty() || r.empty()) {   // a matrix of one row / column: the other sequence against gaps, all cells 0
        const size_t n = f.size() + r.size();
        for (size_t k = 0; k < n; k++) { aligned_1[k] = f.empty() ? '-' : f[k]; aligned_2[k] = r.empty() ? '-' : r[k]; }
        aligned_1[n] = aligned_2[n] = 0;
        *aligned_len = n;
        return MOIRA_OK;
    }
    moira_contig_params p;
    moira_contig_params_default(&p);
    p.match = match; p.mismatch = mismatch; p.gap = gap;
    ContigArgs a;
    fill_contig_args(a, c, &p, 0);
    const std::vector<uint8_t> q1(f.size(), 0), q2(r.size(), 0);
    SingleOut out;
    const int rc = single_pair(c, a, f, q1, r, q2, true, out);
    if (rc) return rc;
    memcpy(aligned_1, out.row_a.data(), out.alen);
    memcpy(aligned_2, out.row_b.data(), out.alen);
    aligned_1[out.alen] = aligned_2[out.alen] = 0;
    *aligned_len = (uint64_t)out.alen;
    if (score) *score = out.score;
    return MOIRA_OK;
}

int moira_make_contig(moira_ctx *c, const char *fwd_aligned, const int32_t *fwd_quals, uint64_t n_fq, const char *rev_aligned,
                      const int32_t *rev_quals, uint64_t n_rq, const moira_contig_params *p, char *contig, int32_t *contig_quals,
                      uint64_t *contig_len, int32_t *overlap, int32_t *gaps, int32_t *mismatches)
{
    if (!c || !fwd_aligned || !rev_aligned || (!fwd_quals && n_fq) || (!rev_quals && n_rq) || !contig || !contig_quals || !contig_len)
        return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    int rc = check_contig_params(p);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    const std::string a1(fwd_aligned), a2(rev_aligned);
    if (a1.size() != a2.size()) return fail(MOIRA_ERR_LENGTH_MISMATCH, "aligned sequences differ in length");
    std::string f, r;
    for (char ch : a1) if (ch != '-') f.push_back(ch);
    for (char ch : a2) if (ch != '-') r.push_back(ch);
    if (f.size() != n_fq || r.size() != n_rq) return fail(MOIRA_ERR_LENGTH_MISMATCH, "LengthMismatchError");   // moira.py:1407-1410
    std::vector<uint8_t> q1(n_fq), q2(n_rq);
    for (uint64_t k = 0; k < n_fq; k++) {
        if (fwd_quals[k] < 0 || fwd_quals[k] > 0xFC) return fail(MOIRA_ERR_BAD_QUALITY, "quality %d is outside 0..252", fwd_quals[k]);
        q1[k] = (uint8_t)fwd_quals[k];
    }
    for (uint64_t k = 0; k < n_rq; k++) {
        if (rev_quals[k] < 0 || rev_quals[k] > 0xFC) return fail(MOIRA_ERR_BAD_QUALITY, "quality %d is outside 0..252", rev_quals[k]);
        q2[k] = (uint8_t)rev_quals[k];
    }
    *contig_len = 0;
    contig[0] = 0;
    if (overlap) *overlap = 0;
    if (gaps) *gaps = 0;
    if (mismatches) *mismatches = 0;
    if (a1.empty()) return MOIRA_OK;
    if (f.empty() || r.empty()) return fail(MOIRA_ERR_BAD_ARG, "an aligned sequence holds gaps only");
    if (a1.size() > f.size() + r.size()) return fail(MOIRA_ERR_BAD_ARG, "alignment has columns that are gaps in both sequences");
    if (p->consensus == MOIRA_CONSENSUS_POSTERIOR && (rc = ensure_post_tables(c))) return rc;
    ContigArgs a;
    fill_contig_args(a, c, p, 0);
    if ((rc = ensure(c->pb[0].rseq, 2 * a1.size() + 64))) return rc;       // the aligned strings
    CU(cudaMemcpy(c->pb[0].rseq.p, a1.data(), a1.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy((char *)c->pb[0].rseq.p + a1.size(), a2.data(), a2.size(), cudaMemcpyHostToDevice));
    a.pre_a1 = (const char *)c->pb[0].rseq.p;
    a.pre_a2 = a.pre_a1 + a1.size();
    a.pre_len = (int32_t)a1.size();
    SingleOut out;
    rc = single_pair(c, a, f, q1, r, q2, false, out);
    if (rc) return rc;
    if (out.status == MOIRA_PAIR_BAD_QUALITY) return fail(MOIRA_ERR_BAD_QUALITY, "a consensus quality is outside 0..252");
    if (out.status != MOIRA_PAIR_OK) return fail(MOIRA_ERR_BAD_ARG, "contig construction failed (pair status %d)", out.status);
    memcpy(contig, out.row_a.data(), out.clen);
    contig[out.clen] = 0;
    for (uint32_t k = 0; k < out.clen; k++) contig_quals[k] = out.row_b[k] > 0xFC ? (int32_t)out.row_b[k] - 256 : out.row_b[k];
    *contig_len = out.clen;
    if (overlap) *overlap = out.overlap;
    if (gaps) *gaps = out.gaps;
    if (mismatches) *mismatches = out.mism;
    return MOIRA_OK;
}

int moira_fp64_peak(moira_ctx *c, int iters, double *ops_per_s_out, double *ms_out)
{
    if (!c || !ops_per_s_out) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    if (iters < 1) iters = 1;
    double ops = 0;
    cudaStream_t s = c->streams[0];
    if (launch_fp64_peak(iters / 8 + 1, c->sm_count, c->d_sink, s, &ops)) return fail(MOIRA_ERR_CUDA, "fp64_peak warm-up launch failed");
    CU(cudaEventRecord(c->t0[0], s));
    if (launch_fp64_peak(iters, c->sm_count, c->d_sink, s, &ops)) return fail(MOIRA_ERR_CUDA, "fp64_peak launch failed");
    CU(cudaEventRecord(c->t1[0], s));
    CU(cudaEventSynchronize(c->t1[0]));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, c->t0[0], c->t1[0]));
    c->launches += 2;
    *ops_per_s_out = ops / (ms * 1e-3);
    if (ms_out) *ms_out = ms;
    return MOIRA_OK;
}

// ---- multi-GPU: all-reduce of the counters (moira_comm.cpp) -----------------------------------------
static int ctx_comm(moira_ctx *c)
{
    if (!c->comm && !(c->comm = comm_state_new())) return fail(MOIRA_ERR_NOMEM, "out of host memory");
    return MOIRA_OK;
}

int moira_comm_unique_id(uint8_t id[MOIRA_COMM_ID_BYTES])
{
    if (!id) return fail(MOIRA_ERR_BAD_ARG, "id is NULL");
    return comm_unique_id(id);
}

int moira_comm_init(moira_ctx *c, const uint8_t id[MOIRA_COMM_ID_BYTES], int rank, int n_ranks)
{
    if (!c || !id) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    int rc = ctx_comm(c);
    if (rc) return rc;
    return comm_init_rank(c->comm, c->device, id, rank, n_ranks);
}

int moira_comm_init_all(moira_ctx *const *ctxs, int n)
{
    if (!ctxs || n < 1) return fail(MOIRA_ERR_BAD_ARG, "no contexts");
    std::vector<CommState *> st(n);
    std::vector<int> dev(n);
    for (int i = 0; i < n; i++) {
        if (!ctxs[i]) return fail(MOIRA_ERR_BAD_ARG, "context %d is NULL", i);
        int rc = ctx_comm(ctxs[i]);
        if (rc) return rc;
        st[i] = ctxs[i]->comm;
        dev[i] = ctxs[i]->device;
    }
    return comm_init_all(st.data(), dev.data(), n);
}

int moira_comm_info(const moira_ctx *c, int *rank_out, int *n_ranks_out)
{
    if (!c) return fail(MOIRA_ERR_BAD_ARG, "ctx is NULL");
    if (!c->comm) { if (rank_out) *rank_out = 0; if (n_ranks_out) *n_ranks_out = 1; return MOIRA_OK; }
    return comm_info(c->comm, rank_out, n_ranks_out);
}

int moira_reduce_counters_device(moira_ctx *c, uint64_t *d_counters, void *stream)
{
    if (!c || !d_counters) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    if (!c->comm) return fail(MOIRA_ERR_BAD_ARG, "context has no communicator (moira_comm_init / moira_comm_init_all)");
    c->launches++;   // the NCCL all-reduce kernel
    return comm_reduce_device(c->comm, d_counters, (cudaStream_t)stream);
}

int moira_reduce_counters(moira_ctx *c, uint64_t counters[MOIRA_N_COUNTERS])
{
    if (!c || !counters) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    if (!c->comm) return fail(MOIRA_ERR_BAD_ARG, "context has no communicator (moira_comm_init / moira_comm_init_all)");
    c->launches++;
    return comm_reduce_host(c->comm, counters);
}

int moira_reduce_counters_all(moira_ctx *const *ctxs, int n, uint64_t *const *counters)
{
    if (!ctxs || !counters || n < 1) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    std::vector<CommState *> st(n);
    for (int i = 0; i < n; i++) {
        if (!ctxs[i] || !counters[i]) return fail(MOIRA_ERR_BAD_ARG, "context / counters %d is NULL", i);
        if (!ctxs[i]->comm) return fail(MOIRA_ERR_BAD_ARG, "context %d has no communicator (moira_comm_init_all)", i);
        st[i] = ctxs[i]->comm;
        ctxs[i]->launches++;
    }
    return comm_reduce_all(st.data(), n, counters);
}

int moira_link_probe(moira_ctx *c, uint64_t bytes, int reps, double *h2d_out, double *d2h_out)
{
    if (!c || !bytes) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    void *h = nullptr, *d = nullptr;
    if (cudaHostAlloc(&h, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return fail(MOIRA_ERR_NOMEM, "cudaHostAlloc of %llu bytes failed", (unsigned long long)bytes); }
    if (cudaMalloc(&d, bytes) != cudaSuccess) { cudaGetLastError(); cudaFreeHost(h); return fail(MOIRA_ERR_NOMEM, "cudaMalloc of %llu bytes failed", (unsigned long long)bytes); }
    memset(h, 1, bytes);
    cudaStream_t s = c->streams[0];
    double best[2] = {0, 0};
    cudaError_t e = cudaSuccess;
    for (int dir = 0; dir < 2 && e == cudaSuccess; dir++)
        for (int r = 0; r < (reps < 1 ? 1 : reps) + 1 && e == cudaSuccess; r++) {   // the first repetition warms up
            cudaEventRecord(c->t0[0], s);
            e = dir == 0 ? cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s) : cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s);
            cudaEventRecord(c->t1[0], s);
            if (e == cudaSuccess) e = cudaEventSynchronize(c->t1[0]);
            float ms = 0;
            if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, c->t0[0], c->t1[0]);
            if (r > 0 && ms > 0) best[dir] = std::max(best[dir], (double)bytes / (ms * 1e-3) / 1e9);
        }
    cudaFree(d);
    cudaFreeHost(h);
    if (e != cudaSuccess) return fail(MOIRA_ERR_CUDA, "link probe failed: %s", cudaGetErrorString(e));
    if (h2d_out) *h2d_out = best[0];
    if (d2h_out) *d2h_out = best[1];
    return MOIRA_OK;
}

int moira_ctx_launch_count(const moira_ctx *c, uint64_t *out)
{
    if (!c || !out) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    *out = c->launches;
    return MOIRA_OK;
}

int moira_ctx_set_timing(moira_ctx *c, int enabled)
{
    if (!c) return fail(MOIRA_ERR_BAD_ARG, "ctx is NULL");
    c->timing = enabled ? 1 : 0;
    c->n_timed = 0;
    c->n_ctimed = 0;
    return MOIRA_OK;
}

int moira_ctx_last_contig_ms(moira_ctx *c, float *ms_out, int *launches_out)
{
    if (!c || !ms_out) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    float total = 0;
    for (int i = 0; i < c->n_ctimed; i++) {
        CU(cudaEventSynchronize(c->ct1[i]));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, c->ct0[i], c->ct1[i]));
        total += ms;
    }
    *ms_out = total;
    if (launches_out) *launches_out = c->n_ctimed;
    return MOIRA_OK;
}

int moira_ctx_last_kernel_ms(moira_ctx *c, float *ms_out, const char **name_out)
{
    if (!c || !ms_out) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    float total = 0;
    for (int i = 0; i < c->n_timed; i++) {
        CU(cudaEventSynchronize(c->t1[i]));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, c->t0[i], c->t1[i]));
        total += ms;
    }
    *ms_out = total;
    if (name_out) *name_out = c->timed_name;
    return MOIRA_OK;
}

}  // extern "C"
