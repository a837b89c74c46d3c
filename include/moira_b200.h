/*
 * moira_b200.h -- C ABI of libmoira_b200.so: the B200 (sm_100a) replacement for moira's
 * per-read quality-filter hot path.
 *
 * What it replaces in the reference (fpusan/moira v1.3.2, paths relative to the reference root):
 *   moira/bernoullimodule.c:66-114   bernoulli.calculate_errors_PB binding (one read per call)
 *   moira/bernoullimodule.c:131-263  prob_j_errors / sum_of_binomials / interpolate / test
 *   moira/moira.py:1637-1679         calculate_errors_poisson            (mode POISSON)
 *   moira/moira.py:1654-1663         Lambda = sum p_i                    (mode EXPECTED_ERROR)
 *   moira/moira.py:806-833           filter half of process_data (truncate, +Ns, floor)
 *   moira/moira.py:872-970           accept/reject decision of write_results
 *
 * Conventions: every entry point returns an int status (0 = MOIRA_OK, < 0 = error enum below);
 * moira_last_error() gives a thread-local message for the last failing call on this thread.
 * Nothing throws or aborts across this boundary.  Signatures use plain pointers and sizes only.
 * There is no CPU fallback: without a usable CUDA device every compute entry point fails with
 * MOIRA_ERR_CUDA.
 *
 * Slab encoding (one byte per base, "in band"):
 *   0x00..0xFC  Phred quality of a called base (0 is accepted and means 1, as the reference
 *               binding maps it, bernoullimodule.c:104-107, moira.py:814)
 *   0xFD        padding: ignored entirely (never needed inside [0, length); the library masks
 *               bytes at and beyond `length` itself, so row padding may hold anything)
 *   0xFE        base 'n': skipped and counted in Ns (bernoullimodule.c:196)
 *   0xFF        base 'N': skipped, counted in Ns, and makes the read "ambiguous" for
 *               --ambigs disallow (moira.py:911)
 * Row r occupies bytes [offsets[r], offsets[r] + lengths[r]); offsets must be multiples of 16
 * and the slab must be readable up to offsets[r] + ceil16(lengths[r]).
 */
#ifndef MOIRA_B200_H
#define MOIRA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOIRA_ABI_VERSION 6

#if defined(__GNUC__)
#define MOIRA_API __attribute__((visibility("default")))
#else
#define MOIRA_API
#endif

/* ---- status codes ------------------------------------------------------------------------- */
#define MOIRA_OK                  0
#define MOIRA_ERR_BAD_ALPHA      -1  /* alpha <= 0, alpha >= 1, or 1-alpha rounds to 1.0 (bernoullimodule.c:79-83) */
#define MOIRA_ERR_LENGTH_MISMATCH -2 /* len(contig) != len(quals)            (bernoullimodule.c:85-90) */
#define MOIRA_ERR_BAD_QUALITY    -3  /* quality outside 0..252 (not representable in the slab) */
#define MOIRA_ERR_CUDA           -4  /* CUDA runtime/driver error, or no device */
#define MOIRA_ERR_BAD_ARG        -5
#define MOIRA_ERR_NOMEM          -6
#define MOIRA_ERR_UNRESOLVED     -7  /* cumulative probability never exceeded 1-alpha (reference would run off its arrays) */
#define MOIRA_ERR_PARSE          -8  /* malformed FASTQ / FASTA+QUAL input (host parsers) */
#define MOIRA_ERR_NCCL           -9  /* NCCL not loadable, or a collective failed (moira_comm_*, moira_reduce_counters*) */

/* ---- enums ---------------------------------------------------------------------------------- */
#define MOIRA_MODE_PB             0  /* --error_calc poisson_binomial   (bernoullimodule.c) */
#define MOIRA_MODE_POISSON        1  /* --error_calc poisson            (moira.py:1637-1679) */
#define MOIRA_MODE_EXPECTED_ERROR 2  /* ee = sum p_i                    (moira.py:1654-1663) */

#define MOIRA_THR_UNCERT          0  /* accept iff ee <= len * thr      (moira.py:950) */
#define MOIRA_THR_MAXERRORS       1  /* accept iff ee <= thr            (moira.py:926) */

#define MOIRA_AMBIGS_TREAT_AS_ERRORS 0 /* ee += Ns                       (moira.py:827-828) */
#define MOIRA_AMBIGS_IGNORE          1
#define MOIRA_AMBIGS_DISALLOW        2 /* reads containing 'N' rejected  (moira.py:911-922) */

#define MOIRA_SLAB_Q8             0  /* one byte per base (the format described above) */
#define MOIRA_SLAB_Q6             1  /* 6-bit transport image of a Q8 slab, see moira_pack_q6 (host-buffer entry points only) */

#define MOIRA_EE_RAW              0  /* ee_out = what calculate_errors_* returns */
#define MOIRA_EE_FINAL            1  /* ee_out = what process_data returns (after +Ns and floor) */

/* ---- per-read flag byte --------------------------------------------------------------------- */
#define MOIRA_FLAG_ACCEPT       0x01
#define MOIRA_FLAG_REASON_MASK  0x0E  /* (flags >> 1) & 7 : see MOIRA_REASON_* */
#define MOIRA_FLAG_LOWER_BOUND  0x10  /* decision mode only: read is a certain reject and ee_out is a LOWER BOUND, not exact */
#define MOIRA_FLAG_HAS_N        0x20  /* read contains an uppercase 'N' (0xFF byte) */
#define MOIRA_FLAG_NUMERIC      0x40  /* statistic not computable (see MOIRA_ERR_UNRESOLVED); read rejected */
#define MOIRA_FLAG_NEAR_CUTOFF  0x80  /* |ee - cutoff| <= 1e-12 * cutoff: decision is within FP64 tolerance of flipping */

#define MOIRA_REASON_NONE    0
#define MOIRA_REASON_ERRORS  1  /* "uncert > x" / "errors > x"  (moira.py:942, 963) */
#define MOIRA_REASON_LENGTH  2  /* "length below"               (moira.py:872-883) */
#define MOIRA_REASON_AMBIGS  3  /* "contains ambiguities"       (moira.py:911-922) */

/* ---- counters (uint64[MOIRA_N_COUNTERS]); device side counterpart of moira.py:406-408, 483-485, 509-519 */
#define MOIRA_CNT_READS        0
#define MOIRA_CNT_ACCEPTED     1
#define MOIRA_CNT_BAD_ERRORS   2
#define MOIRA_CNT_BAD_LENGTH   3
#define MOIRA_CNT_BAD_AMBIGS   4
#define MOIRA_CNT_NEAR_CUTOFF  5
#define MOIRA_CNT_LOWER_BOUND  6
#define MOIRA_CNT_NUMERIC      7
#define MOIRA_CNT_ESCALATED    8   /* reads a first pass could not settle (swept again with more entries); diagnostic: depends on
                                      the cascade setting and the mode, unlike every other counter */
#define MOIRA_CNT_FP64_OPS     9   /* FP64 operations (thread level) executed by the PMF / Lambda sweeps: the numerator of the executed-flop
                                      roofline; diagnostic like MOIRA_CNT_ESCALATED (depends on cascade / ladder choices) */
#define MOIRA_CNT_CLASSIFIED   10  /* reads of classify-first batches (the classifier ran as the first pass and the ladder did every sweep:
                                      exact mode or decisions that need 9 or more entries); diagnostic like the two above */
#define MOIRA_CNT_HIST         16  /* 64 bins of floor(final ee); last bin = >= 63 */
#define MOIRA_N_HIST           64
#define MOIRA_N_COUNTERS       80

typedef struct moira_ctx moira_ctx;

typedef struct moira_params {
    int32_t  mode;        /* MOIRA_MODE_* */
    int32_t  thr_kind;    /* MOIRA_THR_* */
    int32_t  ambigs;      /* MOIRA_AMBIGS_* */
    int32_t  round_flag;  /* --round: floor(ee) before comparing (moira.py:830-831) */
    uint32_t truncate;    /* --truncate: 0 = off (moira.py:806-807, 872) */
    int32_t  exact_ee;    /* 1: exact statistic for every read (what collapse needs, moira.py:466);
                             0: decision mode, certain rejects may carry a lower bound (MOIRA_FLAG_LOWER_BOUND) */
    int32_t  ee_output;   /* MOIRA_EE_* */
    int32_t  length_sort; /* ragged batches: 0 = bucket reads by length on the device when it pays (default), 2 = never */
    int32_t  slab_format; /* MOIRA_SLAB_*: format of the HOST slab given to moira_filter_batch / moira_submit */
    int32_t  cascade;     /* first pass of a decision that needs 3..8 PMF entries: 0 = two-entry sweep first when a pilot launch
                             says it pays (default), 1 = always, 2 = never (one sweep with all the entries) */
    uint32_t max_length;  /* moira_filter_device with d_lengths only (host entry points look at the lengths themselves): longest and */
    uint32_t min_length;  /* shortest read of the batch when the caller knows them, 0 = unknown.  max_length sizes the first
                             pass (a decision needs floor(max_length * uncert) + 2 PMF entries); without it the pass starts at
                             4 entries and the escalation ladder settles the rest */
    double   alpha;       /* --alpha */
    double   thr;         /* --uncert or --maxerrors value */
} moira_params;

/* ---- library / context ---------------------------------------------------------------------- */
MOIRA_API int moira_abi_version(void);
MOIRA_API const char *moira_last_error(void);
MOIRA_API void moira_params_default(moira_params *p);  /* reference defaults: PB, alpha .005, uncert .01, treat_as_errors */

MOIRA_API int moira_device_count(int *n_out);           /* usable CUDA devices (0 without a driver / GPU) */
MOIRA_API int moira_ctx_create(int device, moira_ctx **out);
MOIRA_API int moira_ctx_destroy(moira_ctx *ctx);
MOIRA_API int moira_ctx_sm_count(const moira_ctx *ctx, int *out);

/* Host-libm lookup tables the kernels use (filled at ctx creation with the reference formulas
 * bernoullimodule.c:202, :140, :144).  Index = slab byte; entries 0xFD..0xFF are (p=0, 1-p=1, e=0). */
/* The same tables without a context (pure host code; used by the CPU-side tests).  *e_equals_p is set
 * to 1 when e[Q] == p[Q] and one_minus_p[Q] == 1 - p[Q] bitwise for every entry (selects the 8-byte
 * table variant of the kernels). */
MOIRA_API int moira_build_lut(double p[256], double one_minus_p[256], double e[256], int *e_equals_p);
MOIRA_API int moira_ctx_get_lut(const moira_ctx *ctx, double p[256], double one_minus_p[256], double e[256]);

/* Pinned host memory for the end-to-end path (copies from/to it overlap with kernels). */
MOIRA_API int moira_host_alloc(void **ptr, size_t bytes);
MOIRA_API int moira_host_free(void *ptr);

/* ---- the hot path ---------------------------------------------------------------------------- */

/* Device-resident variant: every pointer is a DEVICE pointer, work is enqueued on `stream`
 * (a cudaStream_t; NULL = default stream) and the call returns without synchronising.
 * offsets may be NULL (row r starts at r * stride); lengths may be NULL (every read has
 * fixed_length bases).  d_counters (uint64[MOIRA_N_COUNTERS]) is ACCUMULATED into; may be NULL.
 * d_ns / d_flags may be NULL.  d_row_marks may be NULL (see moira_count_marks_device). */
MOIRA_API int moira_filter_device(moira_ctx *ctx, const uint8_t *d_slab, const uint64_t *d_offsets,
                        const uint32_t *d_lengths, uint64_t stride, uint32_t fixed_length,
                        uint64_t n_reads, const moira_params *params, const uint32_t *d_row_marks,
                        double *d_ee, int32_t *d_ns, uint8_t *d_flags, uint64_t *d_counters, void *stream);

/* Row marks: one uint32 per row, Ns (number of 0xFE / 0xFF bytes among the first min(length, truncate) bases) in bits
 * 0..30 and "contains 'N'" (a 0xFF byte) in bit 31.  Whoever writes a slab sees every byte anyway -- the library's own
 * producers (6-bit expansion, FASTQ conversion on the device) emit the marks as they go -- and a filter call that is
 * given them (d_row_marks, indexed like d_ee; NULL = count in the sweep) spends no issue slots on N/n accounting
 * (bernoullimodule.c:196-199 is the per-base test it replaces).  This entry point is the stand-alone producer for
 * slabs that came from elsewhere: one memory-bound pass, enqueued on `stream`. */
MOIRA_API int moira_count_marks_device(moira_ctx *ctx, const uint8_t *d_slab, const uint64_t *d_offsets,
                        const uint32_t *d_lengths, uint64_t stride, uint32_t fixed_length, uint64_t n_reads,
                        uint32_t truncate, uint32_t *d_row_marks, void *stream);

/* Host-buffer variant (the call a host program makes): copies the slab in chunks host->device,
 * runs the filter and copies ee / Ns / flags back, overlapping copies of one chunk with the
 * kernels of another on two streams.  Blocks until the results are in the output arrays.
 * counters_out (uint64[MOIRA_N_COUNTERS]) is overwritten; ns_out, flags_out, counters_out may be NULL.
 * slab_bytes = readable size of `slab`. */
MOIRA_API int moira_filter_batch(moira_ctx *ctx, const uint8_t *slab, uint64_t slab_bytes,
                       const uint64_t *offsets, const uint32_t *lengths, uint64_t n_reads,
                       const moira_params *params, double *ee_out, int32_t *ns_out,
                       uint8_t *flags_out, uint64_t *counters_out);

/* Asynchronous pair around the same work, so a host parser can fill the next slab while this one
 * is in flight.  At most MOIRA_MAX_INFLIGHT submissions may be outstanding per context; buffers
 * must stay valid (and should be pinned, moira_host_alloc) until the matching moira_wait. */
#define MOIRA_MAX_INFLIGHT 4
MOIRA_API int moira_submit(moira_ctx *ctx, const uint8_t *slab, uint64_t slab_bytes, const uint64_t *offsets,
                 const uint32_t *lengths, uint64_t n_reads, const moira_params *params,
                 double *ee_out, int32_t *ns_out, uint8_t *flags_out, uint64_t *counters_out,
                 int *ticket_out);
MOIRA_API int moira_wait(moira_ctx *ctx, int ticket);

/* Drop-in for ONE call of bernoulli.calculate_errors_PB(contig, quals, alpha)
 * (bernoullimodule.c:66-114): same validation (alpha in (0,1); Q == 0 -> 1), same result
 * (expected_errors, Ns), computed by the CUDA path as a batch of one. */
MOIRA_API int moira_calculate_errors_PB(moira_ctx *ctx, const char *contig, const int32_t *quals,
                              uint64_t length, double alpha, double *ee_out, int32_t *ns_out);

/* ---- host-side packing (plain C++, no GPU) ---------------------------------------------------- */

/* Merge bases + integer qualities of n reads into the in-band slab.  seq/qual rows are given by
 * in_offsets[r] .. + lengths[r]; out_offsets[r] (multiples of 16) are written by the call, which
 * returns the required slab size in *slab_bytes_out when slab == NULL (sizing pass).
 * lower_n_ambiguous: 1 = 'n' is skipped like 'N' (C core, bernoullimodule.c:196); 0 = only 'N'
 * (Python calculators, moira.py:1605, 1660). */
MOIRA_API int moira_pack_reads(const char *seq, const int32_t *quals, const uint64_t *in_offsets,
                     const uint32_t *lengths, uint64_t n_reads, int lower_n_ambiguous,
                     uint8_t *slab, uint64_t slab_capacity, uint64_t *out_offsets,
                     uint64_t *slab_bytes_out);

/* 6-bit transport format.  A Q8 slab whose bytes are all <= 60 or markers (0xFD..0xFF) -- every Illumina
 * run -- can cross PCIe at 3/4 of its size: every 16 slab bytes become 12 (four 6-bit codes per 3 bytes,
 * little endian; 61/62/63 = pad/'n'/'N').  The image keeps the slab's structure (row r starts at
 * offsets[r] * 3 / 4), so offsets[] and lengths[] are passed unchanged, in Q8 units; the library expands
 * the image on the device before filtering.  slab8_bytes must be a multiple of 16 (pad with 0xFD).
 * Returns MOIRA_ERR_BAD_QUALITY if a byte is not representable (use the Q8 slab then). */
MOIRA_API int moira_pack_q6(const uint8_t *slab8, uint64_t slab8_bytes, uint8_t *slab6, uint64_t slab6_capacity,
                            int n_threads);

/* Parse a FASTQ text buffer (4-line records, moira.py:1152-1204) straight into a slab.
 * Sizing pass when slab == NULL: returns n_reads and the needed slab bytes.
 * hdr_off/hdr_len, seq_off and qual_off give, per read, the byte ranges of the header token, the
 * sequence line and the quality line inside `text` (so the host can slice names, bases and quality
 * characters without re-parsing; sequence and quality lines are lengths[r] bytes long). */
MOIRA_API int moira_parse_fastq(const char *text, uint64_t text_bytes, int fastq_offset, int lower_n_ambiguous,
                      uint8_t *slab, uint64_t slab_capacity, uint64_t *out_offsets,
                      uint32_t *lengths, uint64_t *hdr_off, uint32_t *hdr_len, uint64_t *seq_off,
                      uint64_t *qual_off, uint64_t max_reads, uint64_t *n_reads_out, uint64_t *slab_bytes_out);

/* Parse a FASTA text and its QUAL text (one header line + one data line per record in each file,
 * moira.py:1093-1149) into a slab; same sizing convention as moira_parse_fastq (slab == NULL: returns
 * n_reads and an upper bound of the slab bytes).  Headers are compared after normalisation
 * (NameMismatchError); EmptySeqError / EmptyQualError / LengthMismatchError as in the reference.
 * qual_slab (may be NULL, same capacity and offsets as slab) receives the plain qualities (negative
 * values as 0), for the contig constructor and for writing .qual output. */
MOIRA_API int moira_parse_fasta_qual(const char *fasta, uint64_t fasta_bytes, const char *qual, uint64_t qual_bytes,
                                     int lower_n_ambiguous, uint8_t *slab, uint64_t slab_capacity, uint8_t *qual_slab,
                                     uint64_t *out_offsets, uint32_t *lengths, uint64_t *hdr_off, uint32_t *hdr_len,
                                     uint64_t *seq_off, uint64_t max_reads, uint64_t *n_reads_out, uint64_t *slab_bytes_out);

/* Byte offsets just behind the line_numbers[k]-th newline (ascending line numbers; 0 -> 0, beyond the last newline ->
 * text_bytes) and the number of newlines of the text: cuts two paired files into blocks holding the same records. */
MOIRA_API int moira_line_offsets(const char *text, uint64_t text_bytes, const uint64_t *line_numbers, uint64_t n_queries,
                                 uint64_t *offsets_out, uint64_t *n_lines_out);

/* Number of complete 4-line records in a FASTQ text buffer (to size the output arrays). */
MOIRA_API int moira_fastq_count_reads(const char *text, uint64_t text_bytes, uint64_t *n_reads_out);
/* The record table of a FASTQ text without a slab (parse_fastq's record semantics, moira.py:1152-1204: strip, 4 lines,
 * header token, Empty / LengthMismatch errors): lengths, header token and sequence / quality line byte ranges.  For
 * callers that hand the TEXT itself to the device (moira_filter_pairs reads bases and quality characters where they
 * lie).  lengths == NULL: count only. */
MOIRA_API int moira_index_fastq(const char *text, uint64_t text_bytes, uint32_t *lengths, uint64_t *hdr_off, uint32_t *hdr_len,
                                uint64_t *seq_off, uint64_t *qual_off, uint64_t max_reads, uint64_t *n_reads_out);

/* FASTQ text in, decisions out, in one call: the text is cut into ~64 MB ranges that are parsed (all
 * host threads, moira_parse_fastq semantics) into pinned slabs and submitted asynchronously, so the
 * parser of range k+1 overlaps the copies and kernels of range k (the batched replacement of the
 * reference's read-parse-process loop, moira.py:416-487).  Outputs are in input order; lengths_out
 * (may be NULL) receives the read lengths; counters_out is the sum over all ranges. */
MOIRA_API int moira_filter_fastq(moira_ctx *ctx, const char *text, uint64_t text_bytes, int fastq_offset,
                                 int lower_n_ambiguous, const moira_params *params, uint64_t max_reads,
                                 double *ee_out, int32_t *ns_out, uint8_t *flags_out, uint32_t *lengths_out,
                                 uint64_t *counters_out, uint64_t *n_reads_out);

/* moira_filter_fastq with what a host needs to go on from the decisions to the output files, without parsing the text
 * again: seq_off_out / qual_off_out (may be NULL) receive, per read, the position of its sequence and quality line in
 * `text` (both lengths_out[r] bytes long; moira_fastq_headers finds the header tokens from seq_off_out); labels_out (may
 * be NULL) receives the device-side dereplication of the (truncated) sequences: labels_out[r] is the index of one read
 * with exactly read r's sequence, the same index for all of them (moira_collapse_labels turns labels + ee into the
 * reference's groups).  Labels keep every sequence on the device until the last chunk (text_bytes of HBM). */
MOIRA_API int moira_filter_fastq_ex(moira_ctx *ctx, const char *text, uint64_t text_bytes, int fastq_offset,
                                    int lower_n_ambiguous, const moira_params *params, uint64_t max_reads,
                                    double *ee_out, int32_t *ns_out, uint8_t *flags_out, uint32_t *lengths_out,
                                    uint64_t *seq_off_out, uint64_t *qual_off_out, uint32_t *labels_out,
                                    uint64_t *counters_out, uint64_t *n_reads_out);

/* Dereplication on the device for sequences that already live there (device pointers; enqueued on `stream`): read r is
 * d_seq + (d_offsets ? d_offsets[r] : r * stride), d_lengths[r] (or fixed_length) bytes, cut at `truncate` when that is
 * not 0; any alignment, 8 readable bytes behind the last sequence.  d_labels[r] = index of one read with exactly the
 * same bytes (128-bit hashes select candidates, the bytes decide).  Replaces the hashing and comparing of the reference's
 * `uniques` dictionary (moira.py:409-411, 459-465). */
MOIRA_API int moira_collapse_device(moira_ctx *ctx, const uint8_t *d_seq, const uint64_t *d_offsets, const uint32_t *d_lengths,
                                    uint64_t stride, uint32_t fixed_length, uint64_t n_reads, uint32_t truncate,
                                    uint32_t *d_labels, void *stream);

/* The rest of --collapse from labels (equal label <=> equal sequence; any labelling with that property, label < n):
 * same outputs as moira_collapse, computed without looking at a single base. */
MOIRA_API int moira_collapse_labels(const uint32_t *labels, const double *ee, uint64_t n, uint64_t *group_of_read,
                                    uint64_t *n_groups_out, uint64_t *group_rep, uint64_t *group_size,
                                    uint64_t *member_start, uint64_t *members, uint64_t *abundance_order);
/* The same outputs computed on the device (csrc/moira_groups.cu: first-appearance numbering by atomicMin + scan, stable
 * radix sort by group, running minimum of ee by key, stable sort by size): labels / ee and all outputs are HOST arrays
 * (uint32 on the device and over PCIe).  n < 2^31.  Results are identical to moira_collapse_labels'. */
/* Labels (as moira_collapse_device's) of sequences anywhere in HOST memory: seq_addr[r] = absolute address of read r's
 * bases, seq_len[r] its length (truncate > 0: only the first `truncate` bases count).  The rows are gathered by all host
 * threads, shipped in chunks and labelled on the device; labels_out is a host array of n. */
MOIRA_API int moira_collapse_addr(moira_ctx *ctx, const uint64_t *seq_addr, const uint32_t *seq_len, uint64_t n, uint32_t truncate,
                                  uint32_t *labels_out);
/* The device version without the widening copy: labels / ee are host arrays, or -- on_device != 0 -- device arrays (the
 * outputs of moira_collapse_device and of the filter); the six result arrays are uint32 views of pinned host memory owned
 * by the context, valid until its next moira_collapse_groups / moira_collapse_labels_device call. */
MOIRA_API int moira_collapse_groups(moira_ctx *ctx, const uint32_t *labels, const double *ee, int on_device, uint64_t n,
                                    uint64_t *n_groups_out, const uint32_t **group_of_read, const uint32_t **group_rep,
                                    const uint32_t **group_size, const uint32_t **member_start, const uint32_t **members,
                                    const uint32_t **abundance_order);
MOIRA_API int moira_collapse_labels_device(moira_ctx *ctx, const uint32_t *labels, const double *ee, uint64_t n,
                                           uint64_t *group_of_read, uint64_t *n_groups_out, uint64_t *group_rep,
                                           uint64_t *group_size, uint64_t *member_start, uint64_t *members,
                                           uint64_t *abundance_order);

/* Dereplication of identical sequences, the reference's --collapse (moira.py:459-475, 491-504), on
 * all host threads.  Read r's (already truncated) sequence is text[seq_off[r] .. + seq_len[r]) (text == NULL: seq_off
 * holds absolute addresses); ee[r]
 * is its final expected-error value.  Outputs (all caller-allocated, capacity n, member_start n+1):
 *   group_of_read[r]            group of read r; groups are numbered by first appearance
 *   group_rep[g]                representative read: first read with the strictly smallest ee (moira.py:466)
 *   group_size[g]               number of reads
 *   members[member_start[g] .. member_start[g+1])   the group's reads in the order of the reference's
 *                               names list (each new representative inserted at the front, moira.py:470)
 *   abundance_order[0..G)       group ids, largest group first, ties by first appearance (moira.py:492)
 * n_threads = 0: one per hardware thread. */
MOIRA_API int moira_collapse(const char *text, const uint64_t *seq_off, const uint32_t *seq_len, const double *ee,
                             uint64_t n, int n_threads, uint64_t *group_of_read, uint64_t *n_groups_out,
                             uint64_t *group_rep, uint64_t *group_size, uint64_t *member_start, uint64_t *members,
                             uint64_t *abundance_order);

/* ---- output records: what write_results prints (moira.py:842-970), formatted natively -----------------
 * The reads' text stays where the parsers / the contig kernel left it; the host passes byte ranges.  A NULL base pointer
 * makes the matching offsets absolute addresses (records spread over several buffers). */
typedef struct moira_records {
    const char *hdr_base;      /* header token of read r (no leading '>' / '@'): hdr_base + hdr_off[r], hdr_len[r] bytes; ':' is */
    const uint64_t *hdr_off;   /* written as '_' (moira.py:1121, 1175) */
    const uint32_t *hdr_len;
    const char *seq_base;      /* bases: seq_base + seq_off[r] */
    const uint64_t *seq_off;
    const uint8_t *qual_base;  /* one byte per base: qual_base + qual_off[r]; quality = byte - qual_sub, read as 1 when <= 0 */
    const uint64_t *qual_off;  /* (moira.py:814); with qual_sub == 0 bytes 253..255 stand for -3..-1 (contig rows) */
    const uint32_t *len;       /* bases to write, i.e. after --truncate */
    int32_t qual_sub;
} moira_records;

typedef struct moira_write_opts {
    int32_t fastq;             /* 1: fastq records, 0: fasta + qual records          (--output_format) */
    int32_t fastq_offset;      /* chr(q + offset) of fastq output                    (--fastq_offset) */
    int32_t usearch;           /* header += ";ee=%.2f;size=%d;"                      (--pipeline USEARCH, moira.py:858-863) */
    int32_t names;             /* mothur names lines for the selected groups         (--collapse with --pipeline mothur) */
    const char *relabel;       /* NULL, or header = relabel + index                  (--relabel, moira.py:853-854) */
    uint64_t first_index;      /* index of the first selected record */
    const char *notes[8];      /* by reason code: what follows the header of a rejected record, e.g. "\tuncert > 0.010" */
} moira_write_opts;

#define MOIRA_BLOCK_GOOD        0   /* <prefix>.qc.good.fasta / .fastq */
#define MOIRA_BLOCK_GOOD_QUAL   1   /* <prefix>.qc.good.qual */
#define MOIRA_BLOCK_GOOD_NAMES  2   /* <prefix>.qc.good.names */
#define MOIRA_BLOCK_BAD         3
#define MOIRA_BLOCK_BAD_QUAL    4
#define MOIRA_BLOCK_BAD_NAMES   5
#define MOIRA_BLOCK_REPORT      6   /* <prefix>.contigs.report lines (when the contig statistics are given) */
#define MOIRA_BLOCK_N           7
typedef struct moira_blocks moira_blocks;

/* Format n_sel records in the order sel[0..n_sel) (read indices; NULL = 0, 1, 2, ...) on all host threads.  ee, accept,
 * reason are indexed by read.  With groups (sel_group[k] = group of sel[k], member_start / members as moira_collapse
 * returns them) the USEARCH size and the names lines list the group's members.  overlap / gaps / mismatches (may be
 * NULL, indexed by read) add the contigs report.  The result is a set of memory blocks: parts 0 .. n_parts-1 in
 * output order, each with MOIRA_BLOCK_N byte strings owned by the library until moira_blocks_free. */
MOIRA_API int moira_format_records(const moira_records *records, const moira_write_opts *opts, const uint64_t *sel, uint64_t n_sel,
                                   const double *ee, const uint8_t *accept, const uint8_t *reason, const uint64_t *sel_group,
                                   const uint64_t *member_start, const uint64_t *members, const int32_t *overlap,
                                   const int32_t *gaps, const int32_t *mismatches, int n_threads, moira_blocks **out);
MOIRA_API int moira_blocks_parts(const moira_blocks *blocks, int *n_parts_out);
MOIRA_API int moira_blocks_get(const moira_blocks *blocks, int part, int which, const char **ptr_out, uint64_t *len_out);
/* All parts of block kind `which`, in order, to the open file descriptor `fd` at byte `file_offset` (parallel pwrite; the
 * descriptor's own position is not used or moved).  *written_out = the bytes written. */
MOIRA_API int moira_blocks_write(const moira_blocks *blocks, int which, int fd, uint64_t file_offset, int n_threads,
                                 uint64_t *written_out);
/* Hands a batch back: the next moira_format_records reuses its memory (steady streams of batches).  moira_blocks_free
 * releases a batch (may be NULL) and everything that was recycled. */
/* The same through gzip (--output_compression gz, moira.py:323-370): every part is cut into 65 280-byte pieces, compressed
 * on n_threads threads (zlib, `level` 1..9) into BGZF members -- gzip members that carry their own size, so that any gunzip
 * reads the file and moira_gz_inflate / htslib read it in parallel -- and written at `file_offset`. */
MOIRA_API int moira_blocks_write_gz(const moira_blocks *blocks, int which, int fd, uint64_t file_offset, int level, int n_threads,
                                    uint64_t *written_out);
MOIRA_API int moira_blocks_recycle(moira_blocks *blocks);
MOIRA_API int moira_blocks_free(moira_blocks *blocks);

/* Header token (offset, length) of every FASTQ record from the position of its sequence line, as moira_filter_fastq_ex
 * returns it: the line before, stripped, cut at the first blank, leading '@' dropped (moira.py:1172-1175). */
MOIRA_API int moira_fastq_headers(const char *text, uint64_t text_bytes, const uint64_t *seq_off, uint64_t n_reads, int n_threads,
                                  uint64_t *hdr_off_out, uint32_t *hdr_len_out);

/* Record-aligned cut points of a FASTQ text for n_parts shards (one per GPU): cuts_out[0] = 0 <= ... <= cuts_out[n_parts] =
 * text_bytes; every inner cut is the start of a line that begins with '@' and whose second successor begins with '+' --
 * in a 4-line FASTQ only header lines do (a quality line may begin with '@', but two lines further down is a sequence). */
MOIRA_API int moira_fastq_split(const char *text, uint64_t text_bytes, int n_parts, uint64_t *cuts_out);

/* ---- paired-end contig construction ------------------------------------------
 * The producer of the contigs the filter runs on when --paired is given (process_data, moira.py:791-803):
 * reverse_complement (moira.py:1207-1236) -> nw_align with mothur's overlap refinement
 * (nw_align.pyx:49-202) -> make_contig (moira.py:1375-1558), one warp per pair on the device. */
#define MOIRA_CONSENSUS_BEST 0        /* --consensus_qscore best (default, moira.py:642) */
#define MOIRA_CONSENSUS_SUM 1
#define MOIRA_CONSENSUS_POSTERIOR 2
/* per-pair status */
#define MOIRA_PAIR_OK 0
#define MOIRA_PAIR_EMPTY 1            /* a read of length 0 */
#define MOIRA_PAIR_BAD_BASE 2         /* reverse read holds a character outside the IUPAC table (ValueError, moira.py:1229) */
#define MOIRA_PAIR_BAD_QUALITY 3      /* a consensus quality outside 0..252 (only 'sum' without a cap can get there) */
#define MOIRA_PAIR_TOO_LONG 4         /* reverse read longer than 1024 bases, or longer than the batch maximum given */

typedef struct moira_contig_params {
    int32_t match;          /* --match      1  (moira.py:630) */
    int32_t mismatch;       /* --mismatch  -1  (moira.py:632) */
    int32_t gap;            /* --gap       -2  (moira.py:634) */
    int32_t insert;         /* --insert    20  (moira.py:638) */
    int32_t deltaq;         /* --deltaq     6  (moira.py:640) */
    int32_t consensus;      /* MOIRA_CONSENSUS_* */
    int32_t qscore_cap;     /* --qscore_cap 40, 0 = no cap (moira.py:645) */
    int32_t trim_overlap;   /* --trim_overlap */
} moira_contig_params;
MOIRA_API void moira_contig_params_default(moira_contig_params *p);

/* Contigs of n_pairs read pairs, and -- when filter_params is not NULL -- the quality filter on them, all on
 * the device: the contig kernel writes the filter's slab, the filter kernels read it in place.
 * Inputs (host): bases (ASCII) and qualities of the forward and of the reverse reads as given in the files
 * (the reverse read is reverse-complemented on the device); read r of either file has its bases at
 * [off[r], off[r] + len[r]) of the base array and its qualities at [qual_off[r], ...) of the quality array
 * (qual_off == NULL: same offsets); a quality is byte - qual_base and must land in 0..252.  With the arrays
 * moira_parse_fastq returns, the FASTQ text itself is both arrays (off = seq_off, qual_off = qual_off,
 * qual_base = the FASTQ offset): nothing is repacked on the host.  *_bytes are the sizes of the arrays.
 * Outputs (host), row r at r * out_stride, out_stride >= max(fwd_len) + max(rev_len), multiple of 16:
 *   contig_seq / contig_qual   bases and consensus qualities as make_contig returns them (0..252; 253..255 stand for
 *                              -3..-1: the posterior formulas give -1 when one of the two bases has quality 0)
 *   contig_len, overlap, gaps, mismatches, status[r] (MOIRA_PAIR_*; pairs that fail give an empty contig)
 *   ee / ns / flags / counters  as moira_filter_batch, for the contigs (ignored when filter_params is NULL;
 *                               filter_params->truncate applies to the contig as in moira.py:806-807). */
MOIRA_API int moira_filter_pairs(moira_ctx *ctx, const char *fwd_seq, uint64_t fwd_seq_bytes, const uint8_t *fwd_qual,
                                 uint64_t fwd_qual_bytes, const uint64_t *fwd_off, const uint64_t *fwd_qual_off,
                                 const uint32_t *fwd_len, const char *rev_seq, uint64_t rev_seq_bytes,
                                 const uint8_t *rev_qual, uint64_t rev_qual_bytes, const uint64_t *rev_off,
                                 const uint64_t *rev_qual_off, const uint32_t *rev_len,
                                 int qual_base, uint64_t n_pairs, const moira_contig_params *contig_params,
                                 int lower_n_ambiguous, const moira_params *filter_params, uint64_t out_stride,
                                 char *contig_seq, uint8_t *contig_qual, uint32_t *contig_len, int32_t *overlap,
                                 int32_t *gaps, int32_t *mismatches, uint8_t *status, double *ee_out, int32_t *ns_out,
                                 uint8_t *flags_out, uint64_t *counters_out);

/* Drop-in for ONE call of nw_align.nw_align(seq_1, seq_2, match, mismatch, gap) (nw_align.pyx:49, refine_overlap
 * = True): aligned strings (NUL-terminated, capacity len_1 + len_2 + 1 each) and the reference's score (sum of the
 * matrix cells on the traceback path, nw_align.pyx:127). */
MOIRA_API int moira_nw_align(moira_ctx *ctx, const char *seq_1, const char *seq_2, int match, int mismatch, int gap,
                             char *aligned_1, char *aligned_2, uint64_t *aligned_len, int64_t *score);

/* Drop-in for ONE call of make_contig (moira.py:1375): aligned strings + unaligned qualities in; contig
 * (capacity strlen(aligned) + 1), qualities, overlap length, gaps, mismatches out.  MOIRA_ERR_LENGTH_MISMATCH /
 * MOIRA_ERR_BAD_ARG mirror LengthMismatchError / ValueError (moira.py:1405-1416). */
MOIRA_API int moira_make_contig(moira_ctx *ctx, const char *fwd_aligned, const int32_t *fwd_quals, uint64_t n_fwd_quals,
                                const char *rev_aligned, const int32_t *rev_quals, uint64_t n_rev_quals,
                                const moira_contig_params *params, char *contig, int32_t *contig_quals,
                                uint64_t *contig_len, int32_t *overlap, int32_t *gaps, int32_t *mismatches);

/* ---- gzip inputs and outputs on the host threads (moira_gz.cpp) ------------------------------------------------------
 * Replace gzip.GzipFile(filename).read() behind the reference's magic-byte sniffing (moira.py:1065-1068) and its gzip
 * output files (moira.py:323-370, --output_compression gz).
 *
 * moira_gz_scan: a BGZF file (bgzip / htslib / Illumina FASTQ writers: gzip members of <= 64 KB that carry their own
 * compressed size) -> the number of members and the exact inflated size; any other file -> 0, 0 (nothing is inflated).
 * moira_gz_inflate: the whole file -> a buffer owned by the library (*out, *out_bytes; release with moira_gz_free).
 * BGZF members are inflated in parallel on n_threads threads (<= 0: all), each straight into its place, CRC-32 and size
 * checked; other gzip files (single member, concatenated plain members, zero padding behind the last) by one thread.
 * MOIRA_ERR_PARSE: not gzip, truncated, corrupt.
 * moira_gz_deflate: `bytes` of data -> BGZF members written to fd at file_offset (parallel compression, then pwrite);
 * moira_gz_eof: the empty member BGZF readers expect at the end of a file (28 bytes). */
MOIRA_API int moira_gz_scan(const uint8_t *gz, uint64_t gz_bytes, uint64_t *n_members_out, uint64_t *inflated_bytes_out);
MOIRA_API int moira_gz_inflate(const uint8_t *gz, uint64_t gz_bytes, int n_threads, uint8_t **out, uint64_t *out_bytes);
MOIRA_API int moira_gz_free(uint8_t *buffer);
MOIRA_API int moira_gz_deflate(const uint8_t *data, uint64_t bytes, int level, int n_threads, int fd, uint64_t file_offset,
                               uint64_t *written_out);
MOIRA_API int moira_gz_eof(int fd, uint64_t file_offset, uint64_t *written_out);

/* ---- multi-GPU: the path's only collective ------------------------------------------------------------
 * Reads shard by contiguous chunk, one context per GPU; nothing but the MOIRA_N_COUNTERS counters (good / bad counts,
 * floor(ee) histogram: the device-side counterpart of moira.py:406-408, 483-485, 509-519) is ever exchanged.  Their
 * sum over GPUs is one NCCL all-reduce (ncclUint64, ncclSum) over NVLink / NVSwitch.  NCCL (libnccl.so.2) is bound at
 * run time by the first of these calls; without it they fail with MOIRA_ERR_NCCL.
 *
 * One process per GPU: rank 0 calls moira_comm_unique_id and hands the 128 bytes to the other ranks by any means
 * (file, socket, MPI, torch.distributed); every rank then calls moira_comm_init (collective).
 * One process, several GPUs (the CLI's --devices): moira_comm_init_all over contexts on distinct devices. */
#define MOIRA_COMM_ID_BYTES 128
MOIRA_API int moira_comm_unique_id(uint8_t id[MOIRA_COMM_ID_BYTES]);
MOIRA_API int moira_comm_init(moira_ctx *ctx, const uint8_t id[MOIRA_COMM_ID_BYTES], int rank, int n_ranks);
MOIRA_API int moira_comm_init_all(moira_ctx *const *ctxs, int n_ctx);
MOIRA_API int moira_comm_info(const moira_ctx *ctx, int *rank_out, int *n_ranks_out);   /* 0 of 1 without a communicator */
/* Sum of the counters over all ranks, in place (collective: every rank calls it).
 *   _device: d_counters is this rank's device array; the all-reduce is enqueued on `stream` behind the filter calls
 *            that accumulate into it, no host synchronisation;
 *   host:    counters is this rank's host array; blocks until the sums are in it;
 *   _all:    one process, n contexts: counters[i] is context i's host array; one NCCL group call. */
MOIRA_API int moira_reduce_counters_device(moira_ctx *ctx, uint64_t *d_counters, void *stream);
MOIRA_API int moira_reduce_counters(moira_ctx *ctx, uint64_t counters[MOIRA_N_COUNTERS]);
MOIRA_API int moira_reduce_counters_all(moira_ctx *const *ctxs, int n_ctx, uint64_t *const *counters);

/* Host threads used by moira_parse_fastq (0 = one per hardware thread, at most 64). */
MOIRA_API int moira_set_host_threads(int n);

/* ---- measurement helpers ---------------------------------------------------------------------- */

/* Register-resident FP64 issue-rate micro-benchmark (non-fused DMUL/DADD in the kernel's own
 * 7:4 ratio), the denominator of the FP64 roofline.  Returns operations per second. */
MOIRA_API int moira_fp64_peak(moira_ctx *ctx, int iters, double *ops_per_s_out, double *ms_out);

/* Host <-> device copy rate of this context's GPU right now: `bytes` from / to pinned host memory, best of `reps`,
 * CUDA events on the context's stream -- the denominator of an end-to-end number (frac_of_link).  Ranks that call it at
 * the same time measure the rate they get while sharing the host's memory system and PCIe root complexes. */
MOIRA_API int moira_link_probe(moira_ctx *ctx, uint64_t bytes, int reps, double *h2d_gb_per_s_out, double *d2h_gb_per_s_out);

/* Number of kernel launches this context has issued so far (for bench.py's gpu_launches). */
MOIRA_API int moira_ctx_launch_count(const moira_ctx *ctx, uint64_t *out);

/* Milliseconds spent in the dominant kernel (the first-pass kernel), summed over every launch since
 * timing was last switched on with moira_ctx_set_timing(ctx, 1) (at most 64 launches are kept),
 * measured with CUDA events on the stream the kernel was launched on. */
MOIRA_API int moira_ctx_set_timing(moira_ctx *ctx, int enabled);
MOIRA_API int moira_ctx_last_kernel_ms(moira_ctx *ctx, float *ms_out, const char **name_out);
/* The same for the contig kernel launches of moira_filter_pairs since moira_ctx_set_timing(ctx, 1). */
MOIRA_API int moira_ctx_last_contig_ms(moira_ctx *ctx, float *ms_out, int *launches_out);

#ifdef __cplusplus
}
#endif
#endif /* MOIRA_B200_H */
