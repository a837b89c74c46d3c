// moira_internal.h -- shared between the kernel TU and the C-ABI/host TU.  Not installed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <utility>
#include <vector>

#include "../../include/moira_b200.h"

namespace moira {

// ---- escalation ladder --------------------------------------------------------------------
// Reads whose decision / exact statistic is not settled by the first pass are pushed to rung 0, a
// classifier that estimates how many PMF entries K each of them needs (mean + upper quantile of
// the error count, fp32) and forwards it to the cheapest rung that holds that many.  A rung that
// still cannot settle a read hands it to the next one; the last rung (block-per-read) takes any K.
constexpr int NB = 25;
constexpr int N_TPR_RUNGS = 19;   // rungs 1..19 (one fused launch, ladder_tpr_kernel)
// rungs 1..19 thread-per-read (K = 3..64: the small ones take the reads a two-entry first pass hands on with j* = 2..6,
// at 7..19 operations per base instead of the 22 of K = 8), 20..23 warp-per-read, 24 block-per-read
__host__ __device__ constexpr int rung_cap(int b)
{
    constexpr int caps[NB] = {0, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 0x7fffffff};
    return caps[b];
}

struct FilterArgs {
    // input
    const uint8_t *slab;
    const uint64_t *offsets;     // may be null: row = read * stride
    const uint32_t *lengths;     // may be null: fixed_length
    uint64_t stride;
    uint32_t fixed_length;
    uint64_t base;               // first read of this sub-batch
    uint32_t n;                  // reads in this sub-batch (first pass)
    const uint32_t *row_marks;   // may be null: per row, Ns (bits 0..30) | has-'N' (bit 31) over the first eff bases, indexed like ee[]
    const uint32_t *queue;       // ladder passes: indices relative to base; null on the first pass
    const uint32_t *queue_count; // ladder passes: device count
    // output
    double *ee;
    int32_t *ns;
    uint8_t *flags;
    unsigned long long *counters;
    // parameters
    double oma;                  // 1 - alpha, computed on the host exactly as the reference does
    double thr;
    double z;                    // standard-normal quantile of 1-alpha (K estimate of the classifier)
    double zc;                   // Cornish-Fisher skew bound (z^2-1)/6 + rounding/safety margin
    int32_t mode;
    int32_t thr_kind;
    int32_t ambigs;
    int32_t round_flag;
    uint32_t truncate;
    int32_t exact;
    int32_t ee_output;
    // escalation
    uint32_t *queues;            // [NB][queue_cap]
    uint32_t *queue_counts;      // [NB]
    uint32_t queue_cap;
    // first pass over a length-sorted permutation: this launch covers queue[*seg_start .. +*seg_count)
    const uint32_t *seg_start;
    const uint32_t *seg_count;
    int32_t rung;                // -1 on the first pass, else this kernel's rung
    int32_t min_rung;            // lowest rung the classifier may forward to (its cap exceeds the first-pass K)
    int32_t allow_push;          // 0: no ladder follows (first-pass K provably decides everything)
    double direct_gap;           // ... as long as sum g - sum f of its bases is at most this (else: the classifier)
    int32_t direct_rung;         // first pass with a ladder behind it: a read swept completely but not settled goes straight to
                                 // the rung its two lowest entries call for (rung_from_two_entries), not through the classifier
    // cascaded first pass (decision mode): this launch tracks fewer entries than the k_dec = floor(cutoff) + 2 the
    // decision needs; reads it cannot settle exactly are rejected when the Newton bound on acc[k_dec - 1] allows it
    // and pushed to queue 0 otherwise.  0: off.
    int32_t k_dec;
    int32_t first_k_cap;         // length-bucketed first pass: no bucket is swept with more entries than this (0: no cap)
    const uint32_t *cap_policy;  // not null: the cap applies only if *cap_policy == 1 (verdict of the sorted pilot, decision mode)
    uint32_t first_read;         // length sort: reads [first_read, n) of the sub-batch (the reads before belong to the pilot)
    uint32_t tile0;              // first pass: the launch starts at this warp tile (the tiles before belong to the pilot launch)
    uint32_t *jhist;             // not null (second pilot launch of the cascade): 16 bins of floor(ee) of this launch's reads
    const uint32_t *policy;      // not null: the launch runs only if *policy == policy_want (set on the device by the pilot)
    uint32_t policy_want;
    // tables (device, 256 doubles each)
    const double *lut_p;
    const double *lut_q;
    const double *lut_e;
    int32_t e_equals_p;          // 1: e[Q] == p[Q] and q[Q] == 1 - p[Q] bitwise for all Q (checked on host)
};

struct LaunchCfg {
    int sm_count;
    cudaStream_t stream;
    const void *tmap = nullptr;   // CUtensorMap of the sub-batch's slab when the first pass may use TMA tiles
};

// first pass, thread-per-read; k_wanted is rounded up to an instantiated K (returned)
int launch_pb_first(const FilterArgs &a, int k_wanted, const LaunchCfg &cfg, const char **name);
// Poisson / expected-error, thread-per-read
int launch_lambda(const FilterArgs &a, const LaunchCfg &cfg, const char **name);
// ladder rung b over queue b (b == 0: classifier; b > N_TPR_RUNGS: warp-/block-per-read rungs)
int launch_rung(const FilterArgs &a, int b, const LaunchCfg &cfg);
// the classifier as the first pass: all reads of the sub-batch (or the queue / segment in `a`) -> rung queues
int launch_classify_first(const FilterArgs &a, const LaunchCfg &cfg);
// rungs 1..N_TPR_RUNGS in one launch
int launch_ladder_tpr(const FilterArgs &a, const LaunchCfg &cfg);
// ---- on-device length bucketing of ragged batches (counting sort by padded length) ----------------
constexpr int LEN_BUCKETS = 4096;        // bucket b holds reads with ceil16(eff)/16 == b (last bucket: anything longer)
constexpr int N_FIRST_K = 17;            // first-pass K templates, see first_pass_ks()
__host__ __device__ constexpr int first_pass_k(int i)
{
    constexpr int ks[N_FIRST_K] = {2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 18, 20, 22, 24, 28, 32};
    return ks[i];
}
struct LenSortBufs {
    uint32_t *queue;         // [n] read indices sorted by bucket
    uint32_t *hist;          // [LEN_BUCKETS]
    uint32_t *bucket_start;  // [LEN_BUCKETS]
    uint32_t *cursor;        // [LEN_BUCKETS]
    uint32_t *group_start;   // [N_FIRST_K] segment of the queue handled with first_pass_k(g)
    uint32_t *group_count;   // [N_FIRST_K]
};
// enqueue histogram + scan + scatter; single_group: every read goes to group 0 (modes without a per-length K)
int launch_length_sort(const FilterArgs &a, const LenSortBufs &b, int single_group, const LaunchCfg &cfg);
// first pass over the length-sorted permutation, all K groups in one launch (PB mode)
int launch_sorted_first(const FilterArgs &a, const uint32_t *seg_start, const uint32_t *seg_count, const LaunchCfg &cfg);
int launch_pb_first_k(const FilterArgs &a, int k_index, const LaunchCfg &cfg, const char **name);
// expand a 6-bit transport image into slab bytes [0, slab_bytes) (both device pointers, 16-byte aligned)
// d_marks != nullptr: rows lie at a uniform pitch of `stride` bytes and are `len` bases long; their marks words (zeroed by the
// caller, indexed from the first row of the range) receive Ns / has-N
int launch_unpack_q6(const uint8_t *d_image, uint8_t *d_slab, uint64_t slab_bytes, uint32_t *d_marks, uint64_t stride, uint32_t len,
                     const LaunchCfg &cfg);
// row marks of reads [a.base, a.base + a.n) -> d_marks[a.base + r] (max_len: longest read if known, else 0)
int launch_count_marks(const FilterArgs &a, uint32_t *d_marks, uint32_t max_len, const LaunchCfg &cfg);
// after the pilot launch: *policy = 1 (go on with the full-K first pass) if more than `max_pushed` of the pilot's reads were
// escalated, else 0 (go on with the cascade)
int launch_policy(const uint32_t *queue_count, uint32_t max_pushed, uint32_t *policy, cudaStream_t s);
// cascade with a first stage of 2 .. 5 entries: verdict of the two-entry pilot (2 or undecided), then of the k_first-entry pilot
constexpr uint32_t MOIRA_POLICY_UNDECIDED = 0xFFu;
int launch_policy_first(const uint32_t *queue_count, uint32_t max_pushed, uint32_t *policy, cudaStream_t s);
int launch_policy_second(const uint32_t *jhist, int k_first, int exact, uint32_t *policy, cudaStream_t s);
// length-bucketed first pass with capped K (decision mode): 1 = few of the pilot's reads were handed on, keep the cap; 0 = drop it
int launch_policy_sorted(const uint32_t *queue_counts, uint32_t max_pushed, uint32_t *policy, cudaStream_t s);
int launch_fp64_peak(int iters, int sm_count, double *d_sink, cudaStream_t s, double *ops_out);
int kernels_init(int sm_count);  // sets function attributes (dynamic smem opt-in)
int max_first_pass_k();

// ---- paired-end contig construction (moira_contig.cu) ------------------------------------------------
struct ContigArgs {
    const char *fseq; const uint8_t *fqual; const uint64_t *foff, *fqoff; const uint32_t *flen;   // forward reads
    const char *rseq; const uint8_t *rqual; const uint64_t *roff, *rqoff; const uint32_t *rlen;   // reverse reads
    int32_t qual_base;               // quality = byte - qual_base
    uint64_t n_pairs;
    int32_t match, mismatch, gap, insert, deltaq, consensus, qscore_cap, trim_overlap, lower_n;
    int32_t rev_direct;              // 1: the reverse read is already reverse-complemented (single-pair entry points)
    const int16_t *post_match;       // [256 * 256] posterior consensus quality of two agreeing bases (host glibc)
    const int16_t *post_mis;         // [256 * 256] ... of the better of two disagreeing bases, index [hi * 256 + lo]
    uint32_t *trace;                 // per-warp traceback pointer buffers
    uint64_t trace_words_per_warp;
    int32_t *hbuf;                   // score matrix (only for the nw_align entry point)
    const char *pre_a1, *pre_a2;     // make_contig entry point: the alignment is given
    int32_t pre_len;
    char *al1, *al2;                 // nw_align entry point: aligned strings, length, path score
    int32_t *alen;
    long long *score;
    uint64_t out_stride;
    char *cseq; uint8_t *cqual; uint8_t *slab;
    uint32_t *clen; int32_t *overlap, *gaps, *mism; uint8_t *status;
    uint32_t max_l1, max_l2;         // longest forward / reverse read of the batch
    uint32_t smem_s1, smem_per_warp; // filled by launch_contigs
};
int contig_columns_per_lane(uint32_t max_l2);                         // 0: reverse reads too long
size_t contig_trace_words_per_warp(uint32_t max_l1, uint32_t max_l2);
size_t contig_hbuf_words(uint32_t max_l1, uint32_t max_l2);
int contig_grid(int sm_count, uint64_t n_pairs, int warps);
constexpr int CONTIG_WARPS_PER_CTA = 28;   // the most any instantiation runs (sizes the per-warp scratch)
int launch_contigs(ContigArgs a, bool want_score, const LaunchCfg &cfg);   // 0 ok, -1 CUDA error, -2 unsupported size

// ---- FASTQ parsing on the device (moira_fastq.cu) ---------------------------------------------------------
uint32_t fq_blocks(uint64_t n);   // 4 KB text blocks of a chunk
int launch_fq_index(const uint8_t *d_text, uint64_t lo, uint64_t n, uint32_t *d_block_cnt, uint32_t *d_block_start, uint32_t *d_nl_pos,
                    uint32_t nl_cap, cudaStream_t s);
int launch_fq_records(const uint8_t *d_text, uint64_t lo, uint64_t n, const uint32_t *d_nl_pos, const uint32_t *d_n_nl, uint32_t n_rec,
                      uint32_t extra_line, uint32_t nl_cap, uint32_t rec_cap, uint32_t *d_seq_off, uint32_t *d_qual_off, uint32_t *d_len,
                      uint32_t *d_meta, cudaStream_t s);
int launch_fq_convert(const uint8_t *d_text, const uint32_t *d_seq_off, const uint32_t *d_qual_off, const uint32_t *d_len,
                      uint32_t n_rec, uint32_t stride, int lower_n, int qbase, uint8_t *d_slab, uint32_t *d_meta, uint32_t *d_marks,
                      uint32_t truncate, int sm_count, cudaStream_t s, uint8_t *d_seq_store = nullptr, uint64_t chunk_text_base = 0,
                      uint64_t *d_seq_abs = nullptr, uint32_t *d_seq_eff = nullptr);
void parallel_memcpy(void *dst, const void *src, uint64_t bytes);   // all host threads
int fastq_plan_chunk(const char *text, uint64_t text_bytes, uint64_t pos, uint64_t target_bytes, uint8_t *copy_to,
                     uint64_t *chunk_bytes_out, uint64_t *n_rec_out, uint64_t *n_newlines_out);
// the same without reading the chunk: the cut is the last line start that looks like a record start (ok = 0: none found)
int fastq_plan_chunk_fast(const char *text, uint64_t text_bytes, uint64_t pos, uint64_t target_bytes, uint8_t *copy_to,
                          uint64_t *chunk_bytes_out, int *ok);

// moira_parse_fastq with the number of text bytes consumed (moira_host.cpp)
int parse_fastq_range(const char *text, uint64_t text_bytes, int fastq_offset, int lower_n_ambiguous, uint8_t *slab,
                      uint64_t slab_capacity, uint64_t *out_offsets, uint32_t *lengths, uint64_t *hdr_off,
                      uint32_t *hdr_len, uint64_t *seq_off, uint64_t *qual_off, uint64_t max_reads,
                      uint64_t *n_reads_out, uint64_t *slab_bytes_out, uint64_t *consumed_out, int final_range);

// run fn(0..n_tasks-1) on up to n_threads threads of a persistent host worker pool (moira_host.cpp)
void parallel_run(int n_tasks, int n_threads, const std::function<void(int)> &fn, bool inline_if_busy = false);

// ---- dereplication on the device (moira_dedup.cu) ------------------------------------------------------------
struct DedupArgs {
    const uint8_t *seq;          // sequence bytes; read r at seq + (off ? off[r] : r * stride), any alignment, 8 bytes of slack behind
    const uint64_t *off;
    const uint32_t *len;         // may be null: fixed_len
    uint64_t stride;
    uint32_t fixed_len;
    uint32_t truncate;           // compare contig[:truncate] (moira.py:806-807); 0 = off
    uint64_t base;               // this launch covers reads [base, base + n)
    uint64_t n;
};
// labels[r] = index of one read with exactly r's sequence (the same one for all of them); d_hash: 2 x uint64 per read,
// d_table: table_mask + 1 slots preset to 0xFFFFFFFF (kept between launches that add reads to the same set)
int launch_dedup(const DedupArgs &a, uint64_t *d_hash, uint32_t *d_table, uint32_t table_mask, uint32_t *d_labels, const LaunchCfg &cfg);

// ---- groups, representatives, names order and abundance order from labels, on the device (moira_groups.cu) ----------
size_t groups_device_bytes(uint64_t n);
size_t groups_host_bytes(uint64_t n);
int groups_from_labels_device(int sm_count, cudaStream_t stream, uint8_t *d_buf, uint8_t *h_pinned, const uint32_t *labels, const double *ee,
                              int on_device, uint64_t n, uint64_t *n_groups_out, size_t out_off[6]);

// ---- the counters' all-reduce over GPUs (moira_comm.cpp; NCCL bound at run time) ---------------------------------
struct CommState;                       // one communicator rank; owned by a moira_ctx
CommState *comm_state_new();
void comm_state_free(CommState *s);
int comm_unique_id(uint8_t id[MOIRA_COMM_ID_BYTES]);
int comm_init_rank(CommState *s, int device, const uint8_t id[MOIRA_COMM_ID_BYTES], int rank, int n_ranks);
int comm_init_all(CommState **states, const int *devices, int n);
int comm_reduce_device(CommState *s, uint64_t *d_counters, cudaStream_t stream);
int comm_reduce_host(CommState *s, uint64_t *counters);
int comm_reduce_all(CommState **states, int n, uint64_t *const *counters);
int comm_info(const CommState *s, int *rank, int *n_ranks);

// BGZF members of `n` bytes written to fd at file_offset (moira_gz.cpp)
int gz_deflate_to_fd(const uint8_t *in, uint64_t n, int level, int n_threads, int fd, uint64_t file_offset, uint64_t *written_out);
// the same for several pieces of memory in output order; the threads share the work of all of them
int gz_deflate_segments_to_fd(const std::vector<std::pair<const uint8_t *, uint64_t>> &segs, int level, int n_threads, int fd,
                              uint64_t file_offset, uint64_t *written_out);

// sets the thread-local message returned by moira_last_error() and returns `code`
int fail(int code, const char *fmt, ...);

}  // namespace moira
