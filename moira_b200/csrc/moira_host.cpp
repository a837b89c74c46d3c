// moira_host.cpp -- host-side packers/parsers of the C ABI (no GPU work): turn the reference's
// per-read (sequence, integer qualities) into the in-band uint8 slab the kernels consume.
//   moira_pack_reads   <- what bernoullimodule.c:92-108 does per call (list -> int[]), batched
//   moira_parse_fastq  <- record semantics of parse_fastq, moira/moira.py:1152-1204
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "moira_internal.h"

#define hfail moira::fail

// ---- persistent worker pool -------------------------------------------------------------------------
// The parsers run several short parallel phases per call; spawning threads each time costs more than
// the phases themselves.  Workers are created once and parked on a condition variable.
namespace moira {
namespace {
struct Pool {
    std::mutex run_mutex;                 // one parallel_run at a time
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> workers;
    const std::function<void(int)> *fn = nullptr;
    int n_tasks = 0, next = 0, pending = 0;
    uint64_t generation = 0;
    void worker()
    {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv_work.wait(lk, [&] { return generation != seen && next < n_tasks; });
            while (next < n_tasks) {
                const int t = next++;
                lk.unlock();
                (*fn)(t);
                lk.lock();
                if (--pending == 0) cv_done.notify_all();
            }
            seen = generation;
        }
    }
};
Pool &pool()
{
    static Pool *p = new Pool();   // never destroyed: workers outlive static destruction safely
    return *p;
}
}  // namespace

void parallel_run(int n_tasks, int n_threads, const std::function<void(int)> &fn, bool inline_if_busy)
{
    if (n_tasks <= 0) return;
    if (n_tasks == 1 || n_threads <= 1) {
        for (int t = 0; t < n_tasks; t++) fn(t);
        return;
    }
    Pool &p = pool();
    // one parallel phase at a time uses the pool.  inline_if_busy: a caller that finds it taken runs its tasks itself instead
    // of queueing -- for the staging copies of several contexts' planner threads (one process, several GPUs), which then
    // proceed side by side; every other phase waits its turn (running a big formatting job on one thread would be worse)
    std::unique_lock<std::mutex> run_lock(p.run_mutex, std::defer_lock);
    if (inline_if_busy) {
        if (!run_lock.try_lock()) {
            for (int t = 0; t < n_tasks; t++) fn(t);
            return;
        }
    } else {
        run_lock.lock();
    }
    std::unique_lock<std::mutex> lk(p.m);
    while ((int)p.workers.size() < n_threads - 1) {
        p.workers.emplace_back([&p] { p.worker(); });
        p.workers.back().detach();
    }
    p.fn = &fn;
    p.n_tasks = n_tasks;
    p.next = 0;
    p.pending = n_tasks;
    p.generation++;
    p.cv_work.notify_all();
    while (p.next < p.n_tasks) {          // the caller works too
        const int t = p.next++;
        lk.unlock();
        fn(t);
        lk.lock();
        --p.pending;
    }
    p.cv_done.wait(lk, [&] { return p.pending == 0; });
    p.fn = nullptr;
}
}  // namespace moira

namespace {

inline bool is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }

}  // namespace

extern "C" int moira_pack_reads(const char *seq, const int32_t *quals, const uint64_t *in_offsets,
                                const uint32_t *lengths, uint64_t n_reads, int lower_n_ambiguous, uint8_t *slab,
                                uint64_t slab_capacity, uint64_t *out_offsets, uint64_t *slab_bytes_out)
{
    if ((n_reads && (!seq || !quals || !in_offsets || !lengths)) || !slab_bytes_out)
        return hfail(MOIRA_ERR_BAD_ARG, "NULL argument");
    uint64_t pos = 0;
    for (uint64_t r = 0; r < n_reads; r++) {
        const uint32_t len = lengths[r];
        const uint64_t padded = ((uint64_t)len + 15u) & ~15ull;
        if (slab) {
            if (pos + padded > slab_capacity) return hfail(MOIRA_ERR_BAD_ARG, "slab capacity %llu too small", (unsigned long long)slab_capacity);
            const char *s = seq + in_offsets[r];
            const int32_t *q = quals + in_offsets[r];
            uint8_t *row = slab + pos;
            for (uint32_t i = 0; i < len; i++) {
                const char ch = s[i];
                if (ch == 'N') row[i] = 0xFF;
                else if (ch == 'n' && lower_n_ambiguous) row[i] = 0xFE;
                else {
                    const int32_t v = q[i];
                    if (v > 0xFC) return hfail(MOIRA_ERR_BAD_QUALITY, "quality %d (read %llu, position %u) is outside 0..252", v, (unsigned long long)r, i);
                    row[i] = v <= 0 ? 0 : (uint8_t)v;   // Q <= 0 -> 1 (moira.py:814); 0 is read as 1 by the table
                }
            }
            memset(row + len, 0xFD, padded - len);
        }
        if (out_offsets) out_offsets[r] = pos;
        pos += padded;
    }
    *slab_bytes_out = pos;
    return MOIRA_OK;
}

// ---- 6-bit transport image ------------------------------------------------------------------------
namespace {
// 16 slab bytes -> 12 image bytes; returns the OR of "not representable" flags
__attribute__((target_clones("avx2", "default")))
uint32_t q6_pack_range(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, uint64_t groups)
{
    uint32_t bad = 0;
    for (uint64_t g = 0; g < groups; g++) {
        const uint8_t *s = in + g * 16;
        uint8_t *d = out + g * 12;
        uint8_t c[16];
        for (int i = 0; i < 16; i++) {
            const uint8_t b = s[i];
            const uint8_t m = (uint8_t)(b - 192);                 // 0xFD..0xFF -> 61..63
            c[i] = b >= 0xFD ? m : b;
            bad |= (uint32_t)(b > 60 && b < 0xFD);
        }
        for (int k = 0; k < 4; k++) {
            const uint32_t v = (uint32_t)c[4 * k] | ((uint32_t)c[4 * k + 1] << 6) | ((uint32_t)c[4 * k + 2] << 12) | ((uint32_t)c[4 * k + 3] << 18);
            d[3 * k] = (uint8_t)v;
            d[3 * k + 1] = (uint8_t)(v >> 8);
            d[3 * k + 2] = (uint8_t)(v >> 16);
        }
    }
    return bad;
}
}  // namespace

extern "C" int moira_pack_q6(const uint8_t *slab8, uint64_t slab8_bytes, uint8_t *slab6, uint64_t slab6_capacity, int n_threads)
{
    if ((slab8_bytes && (!slab8 || !slab6)) || (slab8_bytes & 15u)) return hfail(MOIRA_ERR_BAD_ARG, "slab8_bytes must be a multiple of 16");
    if (slab6_capacity < slab8_bytes / 16 * 12) return hfail(MOIRA_ERR_BAD_ARG, "slab6 capacity too small (need %llu)", (unsigned long long)(slab8_bytes / 16 * 12));
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    const uint64_t groups = slab8_bytes / 16;
    if (groups < (1u << 16)) T = 1;
    std::vector<uint32_t> bad(T, 0);
    moira::parallel_run(T, T, [&](int t) {
        const uint64_t a = groups * (uint64_t)t / T, b = groups * (uint64_t)(t + 1) / T;
        bad[t] = q6_pack_range(slab8 + a * 16, slab6 + a * 12, b - a);
    });
    for (int t = 0; t < T; t++)
        if (bad[t]) return hfail(MOIRA_ERR_BAD_QUALITY, "a quality above 60 cannot travel in the 6-bit format; use the Q8 slab");
    return MOIRA_OK;
}

// ---- FASTQ ---------------------------------------------------------------------------------------
// Parallel over host threads: the text is cut into byte ranges, every thread counts the newlines of
// its range, a prefix sum gives each range its first line number, and a thread owns the records
// whose '@' line starts inside its range (records are 4 lines, moira.py:1152-1204).
namespace {

struct FqSeg {
    uint64_t begin = 0, end = 0;        // byte range
    uint64_t newlines = 0;              // in [begin, end)
    uint64_t first_line = 0;            // index of the first line that STARTS in the range
    uint64_t first_line_pos = 0;        // its byte position (== end if none)
    uint64_t rec_start = 0;             // byte position of the first record ('@' line, line index % 4 == 0) at or after it
    uint64_t n_reads = 0, slab_bytes = 0;
    uint64_t consumed = 0;              // text position after the last record parsed
    uint64_t read_base = 0, slab_base = 0;
    int err = 0;
    uint64_t err_local = 0;
    char msg[256] = "";
};

struct FqJob {
    const char *text;
    uint64_t text_bytes;
    int fastq_offset, lower_n;
    uint8_t *slab;
    uint64_t *out_offsets, *hdr_off, *seq_off, *qual_off;
    uint32_t *lengths, *hdr_len;
};

// One row: quality = max(q - offset, 0) (moira.py:1177, :814), overridden by the in-band markers for
// N / n.  Branch-free so the compiler vectorises it (AVX2 clone picked at load time when available).
// Returns the largest q - offset seen.
__attribute__((target_clones("avx2", "default")))
int fq_convert_row(uint8_t *__restrict__ row, const char *__restrict__ s, const unsigned char *__restrict__ q,
                   uint64_t slen, int off, uint8_t low_n)
{
    int worst = 0;
    for (uint64_t i = 0; i < slen; i++) {
        int v = (int)q[i] - off;
        worst = v > worst ? v : worst;
        v = v < 0 ? 0 : v;
        uint8_t out = (uint8_t)v;
        const uint8_t b = (uint8_t)s[i];
        out = b == low_n ? (uint8_t)0xFE : out;       // low_n == 0 never matches a sequence byte
        out = b == (uint8_t)'N' ? (uint8_t)0xFF : out;
        row[i] = out;
    }
    return worst;
}

// Parse the records owned by one range and write their rows into the range's slice of the slab.
void fq_parse_segment(const FqJob &j, FqSeg &g, uint64_t next_begin)
{
    const char *text = j.text;
    uint64_t cur = g.rec_start;
    uint64_t n = 0, pos = 0;
    g.consumed = cur;
    while (cur < j.text_bytes && cur < next_begin && n < g.n_reads) {
        uint64_t lb[4], le[4];
        int have = 0;
        uint64_t c = cur;
        while (have < 4 && c < j.text_bytes) {
            const char *nl = (const char *)memchr(text + c, '\n', j.text_bytes - c);
            const uint64_t end = nl ? (uint64_t)(nl - text) : j.text_bytes;
            uint64_t b = c, e = end;
            while (b < e && is_space((unsigned char)text[b])) b++;          // line.strip(), moira.py:1172
            while (e > b && is_space((unsigned char)text[e - 1])) e--;
            lb[have] = b; le[have] = e;
            have++;
            c = nl ? end + 1 : j.text_bytes;
        }
        if (have < 4) break;   // trailing partial record: ignored, as the reference's loop does
        cur = c;
        // header token: replace('\t',' ').split(' ')[0].lstrip('@')  (moira.py:1175); ':' -> '_' is left to the caller
        uint64_t hb = lb[0], he = lb[0];
        while (he < le[0] && text[he] != ' ' && text[he] != '\t') he++;
        while (hb < he && text[hb] == '@') hb++;
        const uint64_t slen = le[1] - lb[1], qlen = le[3] - lb[3];
        const char *what = nullptr;
        if (slen == 0) what = "EmptySeqError";
        else if (qlen == 0) what = "EmptyQualError";
        else if (slen != qlen) what = "LengthMismatchError";
        else if (slen > 0xFFFFFFF0ull) what = "LengthMismatchError";
        if (what) {
            g.err = MOIRA_ERR_PARSE; g.err_local = n;
            snprintf(g.msg, sizeof(g.msg), "%s: record (%.*s): %llu bases, %llu qualities", what, (int)(he - hb < 100 ? he - hb : 100),
                     text + hb, (unsigned long long)slen, (unsigned long long)qlen);
            break;
        }
        const uint64_t padded = (slen + 15u) & ~15ull;
        if (!j.slab) {   // index only (moira_index_fastq): where the records are, nothing converted
            const uint64_t r = g.read_base + n;
            if (j.lengths) j.lengths[r] = (uint32_t)slen;
            if (j.hdr_off) j.hdr_off[r] = hb;
            if (j.hdr_len) j.hdr_len[r] = (uint32_t)(he - hb);
            if (j.seq_off) j.seq_off[r] = lb[1];
            if (j.qual_off) j.qual_off[r] = lb[3];
        } else {
            const uint64_t r = g.read_base + n;
            uint8_t *row = j.slab + g.slab_base + pos;
            const char *s = text + lb[1];
            const unsigned char *q = (const unsigned char *)text + lb[3];
            const int off = j.fastq_offset;
            const int worst = fq_convert_row(row, s, q, slen, off, j.lower_n ? (uint8_t)'n' : (uint8_t)0);
            // a quality above 252 only matters where the base is not N/n; rare, so re-check exactly
            bool bad_q = false;
            if (worst > 0xFC)
                for (uint64_t i = 0; i < slen; i++)
                    if (s[i] != 'N' && !(s[i] == 'n' && j.lower_n) && (int)q[i] - off > 0xFC) bad_q = true;
            if (bad_q) {
                g.err = MOIRA_ERR_BAD_QUALITY; g.err_local = n;
                snprintf(g.msg, sizeof(g.msg), "a quality of record (%.*s) is outside 0..252", (int)(he - hb < 100 ? he - hb : 100), text + hb);
                break;
            }
            memset(row + slen, 0xFD, padded - slen);
            if (j.out_offsets) j.out_offsets[r] = g.slab_base + pos;
            if (j.lengths) j.lengths[r] = (uint32_t)slen;
            if (j.hdr_off) j.hdr_off[r] = hb;
            if (j.hdr_len) j.hdr_len[r] = (uint32_t)(he - hb);
            if (j.seq_off) j.seq_off[r] = lb[1];
            if (j.qual_off) j.qual_off[r] = lb[3];
        }
        pos += padded;
        n++;
        g.consumed = cur;
    }
    g.slab_bytes = pos;
}

int g_host_threads = 0;

}  // namespace

extern "C" int moira_set_host_threads(int n)
{
    g_host_threads = n < 0 ? 0 : n;
    return MOIRA_OK;
}

extern "C" int moira_parse_fastq(const char *text, uint64_t text_bytes, int fastq_offset, int lower_n_ambiguous,
                                 uint8_t *slab, uint64_t slab_capacity, uint64_t *out_offsets, uint32_t *lengths,
                                 uint64_t *hdr_off, uint32_t *hdr_len, uint64_t *seq_off, uint64_t *qual_off,
                                 uint64_t max_reads, uint64_t *n_reads_out, uint64_t *slab_bytes_out)
{
    return moira::parse_fastq_range(text, text_bytes, fastq_offset, lower_n_ambiguous, slab, slab_capacity, out_offsets,
                                    lengths, hdr_off, hdr_len, seq_off, qual_off, max_reads, n_reads_out, slab_bytes_out,
                                    nullptr, 1);
}

extern "C" int moira_index_fastq(const char *text, uint64_t text_bytes, uint32_t *lengths, uint64_t *hdr_off, uint32_t *hdr_len, uint64_t *seq_off,
                      uint64_t *qual_off, uint64_t max_reads, uint64_t *n_reads_out)
{
    uint64_t cap = 0;
    if (!n_reads_out) return hfail(MOIRA_ERR_BAD_ARG, "NULL argument");
    // lengths == NULL: count only (as moira_parse_fastq with slab == NULL)
    return moira::parse_fastq_range(text, text_bytes, 33, 1, nullptr, 0, nullptr, lengths, hdr_off, hdr_len, seq_off, qual_off, max_reads,
                                    n_reads_out, &cap, nullptr, 1);
}

extern "C" int moira_fastq_count_reads(const char *text, uint64_t text_bytes, uint64_t *n_reads_out)
{
    uint64_t n = 0, cap = 0;
    int rc = moira::parse_fastq_range(text, text_bytes, 33, 1, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr,
                                      nullptr, 0, &n, &cap, nullptr, 1);
    if (n_reads_out) *n_reads_out = n;
    return rc;
}

// consumed_out: bytes of `text` up to the end of the last complete record (what a streaming caller
// advances by; a trailing partial record is left for the next range).  final_range = 0: the range
// may end in the middle of a line, which then does not count.
int moira::parse_fastq_range(const char *text, uint64_t text_bytes, int fastq_offset, int lower_n_ambiguous,
                             uint8_t *slab, uint64_t slab_capacity, uint64_t *out_offsets, uint32_t *lengths,
                             uint64_t *hdr_off, uint32_t *hdr_len, uint64_t *seq_off, uint64_t *qual_off,
                             uint64_t max_reads, uint64_t *n_reads_out, uint64_t *slab_bytes_out, uint64_t *consumed_out,
                             int final_range)
{
    if (!text || !n_reads_out || !slab_bytes_out) return hfail(MOIRA_ERR_BAD_ARG, "NULL argument");
    int T = g_host_threads > 0 ? g_host_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    if (text_bytes < (1u << 20)) T = 1;
    if (consumed_out) *consumed_out = 0;
    std::vector<FqSeg> segs(T);
    for (int t = 0; t < T; t++) {
        segs[t].begin = text_bytes * (uint64_t)t / T;
        segs[t].end = text_bytes * (uint64_t)(t + 1) / T;
    }
    auto run = [&](const std::function<void(int)> &fn) { moira::parallel_run(T, T, fn); };
    // phase 0: newlines per range, and the first line start inside each range
    run([&](int t) {
        FqSeg &g = segs[t];
        uint64_t cnt = 0;
        const char *p = text + g.begin, *e = text + g.end;
        while (p < e) {
            const char *nl = (const char *)memchr(p, '\n', e - p);
            if (!nl) break;
            cnt++;
            p = nl + 1;
        }
        g.newlines = cnt;
        if (g.begin == 0 || text[g.begin - 1] == '\n') g.first_line_pos = g.begin;
        else {
            const char *nl = (const char *)memchr(text + g.begin, '\n', text_bytes - g.begin);
            g.first_line_pos = nl ? (uint64_t)(nl - text) + 1 : text_bytes;
        }
    });
    uint64_t before = 0;
    for (int t = 0; t < T; t++) {
        FqSeg &g = segs[t];
        // lines starting at or after g.begin have index >= (#newlines before g.begin) (+1 if begin is mid-line)
        g.first_line = before + ((g.begin == 0 || text[g.begin - 1] == '\n') ? 0 : 1);
        before += g.newlines;
    }
    // Reads per range follow from the line numbers alone (records start at lines 0, 4, 8, ...; a record
    // counts if all 4 of its lines exist), and so does an upper bound of the slab bytes a range produces
    // (each row <= half of its record's text + 15 bytes of padding).  So the ranges can be parsed in ONE
    // pass, each thread writing its rows into its own slice of the slab; offsets[] skips the slack.
    run([&](int t) {
        FqSeg &g = segs[t];
        uint64_t cur = g.first_line_pos, line = g.first_line;
        while (cur < text_bytes && (line & 3u) != 0) {
            const char *nl = (const char *)memchr(text + cur, '\n', text_bytes - cur);
            if (!nl) { cur = text_bytes; break; }
            cur = (uint64_t)(nl - text) + 1;
            line++;
        }
        g.rec_start = cur < text_bytes ? cur : text_bytes;
    });
    // an unterminated last line is a line only at the true end of the input (a streaming range may end mid-line)
    const uint64_t total_lines = before + ((final_range && text_bytes && text[text_bytes - 1] != '\n') ? 1 : 0);
    const uint64_t n_complete = total_lines / 4;
    auto records_before_line = [&](uint64_t line) { return std::min<uint64_t>((line + 3) / 4, n_complete); };
    uint64_t n = 0, cap = 0;
    for (int t = 0; t < T; t++) {
        FqSeg &g = segs[t];
        const uint64_t line_end = t + 1 < T ? segs[t + 1].first_line : total_lines;
        const uint64_t lo = records_before_line(g.first_line), hi = records_before_line(std::max(line_end, g.first_line));
        g.n_reads = hi - lo;
        g.read_base = lo;
        g.slab_base = cap;
        // the range's records are exactly the text [rec_start(t), rec_start(t+1)): each row <= half of its
        // record's bytes, plus at most 15 bytes of padding
        const uint64_t rec_end = t + 1 < T ? std::max(segs[t + 1].rec_start, g.rec_start) : text_bytes;
        cap += (((rec_end - g.rec_start) / 2 + 16 * (g.n_reads + 1)) + 15u) & ~15ull;
        n += g.n_reads;
    }
    *n_reads_out = n;
    *slab_bytes_out = cap;
    const bool index_only = !slab && lengths;     // moira_index_fastq: the record table without a slab
    if (!slab && !index_only) return MOIRA_OK;
    if (n > max_reads) return hfail(MOIRA_ERR_BAD_ARG, "more than max_reads = %llu records", (unsigned long long)max_reads);
    if (!index_only && cap > slab_capacity) return hfail(MOIRA_ERR_BAD_ARG, "slab capacity too small (need %llu)", (unsigned long long)cap);
    FqJob job{text, text_bytes, fastq_offset, lower_n_ambiguous, slab, out_offsets, hdr_off, seq_off, qual_off, lengths, hdr_len};
    auto next_begin = [&](int t) { return t + 1 < T ? segs[t + 1].rec_start : text_bytes; };
    run([&](int t) { if (segs[t].n_reads) fq_parse_segment(job, segs[t], next_begin(t)); });
    uint64_t used = 0;
    for (int t = 0; t < T; t++) {
        if (segs[t].err) return hfail(segs[t].err, "%s (record %llu)", segs[t].msg, (unsigned long long)(segs[t].read_base + segs[t].err_local));
        if (segs[t].n_reads) {
            used = segs[t].slab_base + segs[t].slab_bytes;
            if (consumed_out) *consumed_out = segs[t].consumed;
        }
    }
    *slab_bytes_out = used;   // end of the last row: what has to travel to the GPU
    return MOIRA_OK;
}

void moira::parallel_memcpy(void *dst, const void *src, uint64_t bytes)
{
    int T = g_host_threads > 0 ? g_host_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    if (bytes < (1u << 20)) T = 1;
    moira::parallel_run(T, T, [&](int t) {
        const uint64_t b = bytes * (uint64_t)t / T, e = bytes * (uint64_t)(t + 1) / T;
        memcpy((char *)dst + b, (const char *)src + b, e - b);
    }, true);
}

// Plan one chunk of a FASTQ text for the device parser: count the newlines of text[pos, pos + target) on all host
// threads (copying the bytes into `copy_to`, pinned staging, on the way when the caller's text is pageable) and cut
// the chunk behind its last complete 4-line record.  Outputs: bytes of the chunk, its records, its newlines.
int moira::fastq_plan_chunk(const char *text, uint64_t text_bytes, uint64_t pos, uint64_t target_bytes, uint8_t *copy_to,
                            uint64_t *chunk_bytes_out, uint64_t *n_rec_out, uint64_t *n_newlines_out)
{
    const uint64_t len = std::min<uint64_t>(target_bytes, text_bytes - pos);
    const bool final_range = pos + len >= text_bytes;
    int T = g_host_threads > 0 ? g_host_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    if (len < (1u << 20)) T = 1;
    std::vector<uint64_t> cnt(T, 0);
    const char *base = text + pos;
    moira::parallel_run(T, T, [&](int t) {
        const uint64_t b = len * (uint64_t)t / T, e = len * (uint64_t)(t + 1) / T;
        uint64_t c = 0;
        const char *p = base + b, *end = base + e;
        while (p < end) {
            const char *nl = (const char *)memchr(p, '\n', end - p);
            if (!nl) break;
            c++;
            p = nl + 1;
        }
        cnt[t] = c;
        if (copy_to) memcpy(copy_to + b, base + b, e - b);
    });
    uint64_t total_nl = 0;
    for (int t = 0; t < T; t++) total_nl += cnt[t];
    const uint64_t lines = total_nl + ((final_range && len && base[len - 1] != '\n') ? 1 : 0);
    const uint64_t n_rec = lines / 4;
    *n_rec_out = n_rec;
    if (n_rec == 0) { *chunk_bytes_out = final_range ? len : 0; *n_newlines_out = 0; return MOIRA_OK; }
    const uint64_t keep = 4 * n_rec;
    uint64_t end = len;
    if (keep <= total_nl) {
        // the chunk ends just behind its keep-th newline: walk back over the newlines that follow it
        uint64_t p = len;
        for (uint64_t k = 0; k <= total_nl - keep; k++) {
            const char *nl = (const char *)memrchr(base, '\n', p);
            p = (uint64_t)(nl - base);
        }
        end = p + 1;
    }
    *chunk_bytes_out = end;
    *n_newlines_out = keep <= total_nl ? keep : total_nl;
    return MOIRA_OK;
}

// Plan a chunk WITHOUT reading it: cut text[pos, pos + target) at the last line start that looks like the start of a
// record -- a line beginning with '@' whose second successor begins with '+'.  A quality line may begin with '@' too,
// but then the line two further down is a sequence line, and those never begin with '+'.  The device verifies the guess
// (a chunk cut this way must hold a multiple of four lines); the caller falls back to fastq_plan_chunk when it fails.
int moira::fastq_plan_chunk_fast(const char *text, uint64_t text_bytes, uint64_t pos, uint64_t target_bytes, uint8_t *copy_to,
                                 uint64_t *chunk_bytes_out, int *ok)
{
    const uint64_t len = std::min<uint64_t>(target_bytes, text_bytes - pos);
    const char *base = text + pos, *text_end = text + text_bytes;
    uint64_t end = len;
    *ok = 1;
    if (pos + len < text_bytes) {
        *ok = 0;
        uint64_t p = len;
        for (int tries = 0; tries < 64 && p > 0; tries++) {
            const char *nl = (const char *)memrchr(base, '\n', p);
            if (!nl) break;
            const uint64_t s = (uint64_t)(nl - base) + 1;
            p = (uint64_t)(nl - base);
            if (s >= len || base[s] != '@') continue;
            const char *l1 = (const char *)memchr(base + s, '\n', (size_t)(text_end - (base + s)));
            if (!l1 || l1 + 1 >= text_end) continue;
            const char *l2 = (const char *)memchr(l1 + 1, '\n', (size_t)(text_end - (l1 + 1)));
            if (!l2 || l2 + 1 >= text_end) continue;
            if (l2[1] == '+') { end = s; *ok = 1; break; }
        }
        if (!*ok || end == 0) { *ok = 0; return MOIRA_OK; }
    }
    if (copy_to && end) parallel_memcpy(copy_to, base, end);
    *chunk_bytes_out = end;
    return MOIRA_OK;
}

// Record-aligned shard boundaries of a FASTQ text (one shard per GPU): the same local test as above, forwards from the
// even split points.
extern "C" int moira_fastq_split(const char *text, uint64_t text_bytes, int n_parts, uint64_t *cuts_out)
{
    if (!cuts_out || n_parts < 1 || (!text && text_bytes)) return moira::fail(MOIRA_ERR_BAD_ARG, "bad argument");
    cuts_out[0] = 0;
    cuts_out[n_parts] = text_bytes;
    const char *end = text + text_bytes;
    for (int p = 1; p < n_parts; p++) {
        uint64_t at = text_bytes * (uint64_t)p / (uint64_t)n_parts;
        if (at < cuts_out[p - 1]) at = cuts_out[p - 1];
        uint64_t cut = text_bytes;
        const char *cur = text + at;
        while (cur < end) {
            const char *nl = (const char *)memchr(cur, '\n', (size_t)(end - cur));
            if (!nl || nl + 1 >= end) break;
            const char *s = nl + 1;                                  // a line start
            cur = s;
            if (*s != '@') continue;
            const char *l1 = (const char *)memchr(s, '\n', (size_t)(end - s));
            if (!l1 || l1 + 1 >= end) break;
            const char *l2 = (const char *)memchr(l1 + 1, '\n', (size_t)(end - (l1 + 1)));
            if (!l2 || l2 + 1 >= end) break;
            if (l2[1] == '+') { cut = (uint64_t)(s - text); break; }
        }
        cuts_out[p] = cut;
    }
    return MOIRA_OK;
}

// Byte offsets just behind given numbers of newlines (ascending) of a text, and the text's newline count: one parallel
// count over 1 MB blocks, then a short scan per query.  This is how blocks of whole records are cut out of two paired
// files so that they hold the same records (moira.py:1093-1204 reads them in lock step): offsets of lines 4R, 8R, ...
// in both.  An offset beyond the last newline is text_bytes.
extern "C" int moira_line_offsets(const char *text, uint64_t text_bytes, const uint64_t *line_numbers, uint64_t n_queries,
                                  uint64_t *offsets_out, uint64_t *n_lines_out)
{
    if ((!text && text_bytes) || (n_queries && (!line_numbers || !offsets_out))) return moira::fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    constexpr uint64_t BLK = 1u << 20;
    const uint64_t nb = (text_bytes + BLK - 1) / BLK;
    std::vector<uint64_t> cnt(nb + 1, 0);
    int T = g_host_threads > 0 ? g_host_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    const int parts = (int)std::min<uint64_t>(std::max<uint64_t>(nb, 1), (uint64_t)T * 8);
    moira::parallel_run(parts, T, [&](int p) {
        for (uint64_t b = nb * (uint64_t)p / parts, e = nb * (uint64_t)(p + 1) / parts; b < e; b++) {
            const char *q = text + b * BLK, *end = text + std::min(text_bytes, (b + 1) * BLK);
            uint64_t c = 0;
            while (q < end) {
                const char *nl = (const char *)memchr(q, '\n', (size_t)(end - q));
                if (!nl) break;
                c++;
                q = nl + 1;
            }
            cnt[b + 1] = c;
        }
    });
    for (uint64_t b = 0; b < nb; b++) cnt[b + 1] += cnt[b];      // cnt[b] = newlines before block b
    if (n_lines_out) *n_lines_out = cnt[nb];
    uint64_t b = 0;
    for (uint64_t k = 0; k < n_queries; k++) {
        const uint64_t want = line_numbers[k];
        if (k && want < line_numbers[k - 1]) return moira::fail(MOIRA_ERR_BAD_ARG, "line numbers must ascend");
        if (want == 0) { offsets_out[k] = 0; continue; }
        if (want > cnt[nb]) { offsets_out[k] = text_bytes; continue; }
        while (b + 1 < nb && cnt[b + 1] < want) b++;             // the want-th newline lies in block b
        const char *q = text + b * BLK, *end = text + std::min(text_bytes, (b + 1) * BLK);
        uint64_t left = want - cnt[b];
        while (left) {
            const char *nl = (const char *)memchr(q, '\n', (size_t)(end - q));
            q = nl + 1;
            left--;
        }
        offsets_out[k] = (uint64_t)(q - text);
    }
    return MOIRA_OK;
}

// ---- FASTA + QUAL ------------------------------------------------------------------------------------
// Record semantics of parse_fasta_and_qual (moira/moira.py:1093-1149, single-end): both files hold one
// header line and one data line per record ("Expects sequences and qualities to be stored in a single
// line"); headers are compared after normalisation (strip, tab -> space, first token, lstrip('>'));
// qualities are whitespace-separated integers.  Parallel over FASTA byte ranges; the matching position
// in the QUAL text is found through per-range newline counts.
namespace {

struct TwoLineIndex {
    std::vector<uint64_t> begin, nl_before;   // byte ranges and number of newlines before each range
    uint64_t total_lines = 0;
};

TwoLineIndex index_lines(const char *text, uint64_t bytes, int T)
{
    TwoLineIndex ix;
    ix.begin.resize(T + 1);
    ix.nl_before.assign(T + 1, 0);
    std::vector<uint64_t> cnt(T, 0);
    for (int t = 0; t <= T; t++) ix.begin[t] = bytes * (uint64_t)t / T;
    moira::parallel_run(T, T, [&](int t) {
        uint64_t c = 0;
        const char *p = text + ix.begin[t], *e = text + ix.begin[t + 1];
        while (p < e) {
            const char *nl = (const char *)memchr(p, '\n', e - p);
            if (!nl) break;
            c++;
            p = nl + 1;
        }
        cnt[t] = c;
    });
    for (int t = 0; t < T; t++) ix.nl_before[t + 1] = ix.nl_before[t] + cnt[t];
    ix.total_lines = ix.nl_before[T] + ((bytes && text[bytes - 1] != '\n') ? 1 : 0);
    return ix;
}

// byte position where line `line` starts (bytes if there is no such line)
uint64_t line_start(const char *text, uint64_t bytes, const TwoLineIndex &ix, uint64_t line)
{
    if (line == 0) return 0;
    int u = 0;
    const int T = (int)ix.begin.size() - 1;
    while (u + 1 < T && ix.nl_before[u + 1] < line) u++;
    uint64_t pos = ix.begin[u];
    for (uint64_t k = line - ix.nl_before[u]; k > 0; k--) {
        const char *nl = (const char *)memchr(text + pos, '\n', bytes - pos);
        if (!nl) return bytes;
        pos = (uint64_t)(nl - text) + 1;
    }
    return pos;
}

inline void next_line(const char *text, uint64_t bytes, uint64_t &cur, uint64_t &b, uint64_t &e)
{
    const char *nl = cur < bytes ? (const char *)memchr(text + cur, '\n', bytes - cur) : nullptr;
    const uint64_t end = nl ? (uint64_t)(nl - text) : bytes;
    b = cur < bytes ? cur : bytes;
    e = end;
    while (b < e && is_space((unsigned char)text[b])) b++;
    while (e > b && is_space((unsigned char)text[e - 1])) e--;
    cur = nl ? end + 1 : bytes;
}

inline void header_token(const char *text, uint64_t b, uint64_t e, uint64_t &hb, uint64_t &he)
{
    hb = b; he = b;
    while (he < e && text[he] != ' ' && text[he] != '\t') he++;
    while (hb < he && text[hb] == '>') hb++;
}

}  // namespace

extern "C" int moira_parse_fasta_qual(const char *fasta, uint64_t fasta_bytes, const char *qual, uint64_t qual_bytes,
                                      int lower_n_ambiguous, uint8_t *slab, uint64_t slab_capacity, uint8_t *qual_slab,
                                      uint64_t *out_offsets, uint32_t *lengths, uint64_t *hdr_off, uint32_t *hdr_len,
                                      uint64_t *seq_off, uint64_t max_reads, uint64_t *n_reads_out, uint64_t *slab_bytes_out)
{
    if (!fasta || !qual || !n_reads_out || !slab_bytes_out) return hfail(MOIRA_ERR_BAD_ARG, "NULL argument");
    int T = g_host_threads > 0 ? g_host_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    if (fasta_bytes < (1u << 20)) T = 1;
    const TwoLineIndex fx = index_lines(fasta, fasta_bytes, T), qx = index_lines(qual, qual_bytes, T);
    // trailing blank lines do not make records (the reference stops at the first empty header pair)
    uint64_t f_lines = fx.total_lines, q_lines = qx.total_lines;
    const uint64_t n = f_lines / 2;
    if (q_lines / 2 < n) return hfail(MOIRA_ERR_PARSE, "NameMismatchError: the qual file has %llu records, the fasta file %llu",
                                      (unsigned long long)(q_lines / 2), (unsigned long long)n);
    // records [r0(t), r0(t+1)) per thread; slab slices sized from the FASTA bytes they cover
    std::vector<uint64_t> r0(T + 1), fpos(T + 1), qpos(T + 1), sbase(T + 1, 0);
    for (int t = 0; t <= T; t++) r0[t] = n * (uint64_t)t / T;
    moira::parallel_run(T + 1, T, [&](int t) {
        fpos[t] = t == T ? fasta_bytes : line_start(fasta, fasta_bytes, fx, 2 * r0[t]);
        qpos[t] = t == T ? qual_bytes : line_start(qual, qual_bytes, qx, 2 * r0[t]);
    });
    for (int t = 0; t < T; t++)
        sbase[t + 1] = sbase[t] + (((fpos[t + 1] - fpos[t]) + 16 * (r0[t + 1] - r0[t] + 1) + 15u) & ~15ull);
    *n_reads_out = n;
    *slab_bytes_out = sbase[T];
    if (!slab) return MOIRA_OK;
    if (n > max_reads) return hfail(MOIRA_ERR_BAD_ARG, "more than max_reads = %llu records", (unsigned long long)max_reads);
    if (sbase[T] > slab_capacity) return hfail(MOIRA_ERR_BAD_ARG, "slab capacity too small (need %llu)", (unsigned long long)sbase[T]);
    struct Err { int code = 0; uint64_t rec = 0; char msg[256] = ""; };
    std::vector<Err> errs(T);
    std::vector<uint64_t> used(T, 0);
    moira::parallel_run(T, T, [&](int t) {
        uint64_t fc = fpos[t], qc = qpos[t], pos = 0;
        Err &er = errs[t];
        for (uint64_t r = r0[t]; r < r0[t + 1]; r++) {
            uint64_t hb, he, sb, se, qhb, qhe, qb, qe, a, b;
            next_line(fasta, fasta_bytes, fc, a, b);
            header_token(fasta, a, b, hb, he);
            next_line(fasta, fasta_bytes, fc, sb, se);
            next_line(qual, qual_bytes, qc, a, b);
            header_token(qual, a, b, qhb, qhe);
            next_line(qual, qual_bytes, qc, qb, qe);
            const uint64_t slen = se - sb;
            const char *what = nullptr;
            if (he - hb != qhe - qhb || memcmp(fasta + hb, qual + qhb, he - hb) != 0) what = "NameMismatchError";
            else if (slen == 0) what = "EmptySeqError";
            else if (qe == qb) what = "EmptyQualError";
            uint8_t *row = slab + sbase[t] + pos;
            uint8_t *qrow = qual_slab ? qual_slab + sbase[t] + pos : nullptr;
            uint64_t nq = 0;
            bool bad_q = false;
            if (!what) {
                uint64_t p = qb;
                while (p < qe) {                                   // map(int, line.split(' ')), moira.py:1124
                    while (p < qe && is_space((unsigned char)qual[p])) p++;
                    if (p >= qe) break;
                    bool neg = false;
                    if (qual[p] == '-' || qual[p] == '+') { neg = qual[p] == '-'; p++; }
                    long v = 0;
                    bool digits = false;
                    while (p < qe && qual[p] >= '0' && qual[p] <= '9') { v = v < 100000 ? v * 10 + (qual[p] - '0') : v; p++; digits = true; }
                    if (!digits || (p < qe && !is_space((unsigned char)qual[p]))) { what = "ValueError"; break; }
                    if (neg) v = -v;
                    if (nq < slen) {
                        const char ch = fasta[sb + nq];
                        const uint8_t qv = v <= 0 ? 0 : (v > 0xFC ? 0xFC : (uint8_t)v);
                        if (qrow) qrow[nq] = v <= 0 ? 0 : (v > 255 ? 255 : (uint8_t)v);
                        if (ch == 'N') row[nq] = 0xFF;
                        else if (ch == 'n' && lower_n_ambiguous) row[nq] = 0xFE;
                        else { row[nq] = qv; bad_q |= v > 0xFC; }
                    }
                    nq++;
                }
                if (!what && nq != slen) what = "LengthMismatchError";
            }
            if (what || bad_q) {
                er.code = what ? MOIRA_ERR_PARSE : MOIRA_ERR_BAD_QUALITY;
                er.rec = r;
                snprintf(er.msg, sizeof(er.msg), "%s: record %llu (%.*s): %llu bases, %llu qualities", what ? what : "quality outside 0..252",
                         (unsigned long long)r, (int)(he - hb < 100 ? he - hb : 100), fasta + hb, (unsigned long long)slen, (unsigned long long)nq);
                return;
            }
            const uint64_t padded = (slen + 15u) & ~15ull;
            memset(row + slen, 0xFD, padded - slen);
            if (qrow) memset(qrow + slen, 0, padded - slen);
            if (out_offsets) out_offsets[r] = sbase[t] + pos;
            if (lengths) lengths[r] = (uint32_t)slen;
            if (hdr_off) hdr_off[r] = hb;
            if (hdr_len) hdr_len[r] = (uint32_t)(he - hb);
            if (seq_off) seq_off[r] = sb;
            pos += padded;
        }
        used[t] = pos;
    });
    for (int t = 0; t < T; t++)
        if (errs[t].code) return hfail(errs[t].code, "%s", errs[t].msg);
    for (int t = T - 1; t >= 0; t--)
        if (r0[t + 1] > r0[t]) { *slab_bytes_out = sbase[t] + used[t]; break; }
    return MOIRA_OK;
}
