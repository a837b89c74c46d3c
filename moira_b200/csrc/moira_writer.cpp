// moira_writer.cpp -- native formatting of the output records: what write_results prints for every
// read or collapsed group (moira/moira.py:842-970) -- fasta + qual or fastq records, mothur .names lines, the USEARCH
// ";ee=%.2f;size=%d;" header, --relabel, the rejection reason behind the header, the contigs report -- produced from the
// byte ranges the parsers (host or device) and the contig kernel left behind, on all host threads, into memory
// blocks the host program only has to write to its files.  The accept / reject decision itself is the device's.
#include <errno.h>
#include <atomic>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "moira_internal.h"

struct moira_blocks {
    int n_parts = 0;
    std::vector<std::string> buf;   // [part][MOIRA_BLOCK_N]
};

namespace {

// formatted batches follow each other at a steady size: a finished batch handed back with moira_blocks_recycle keeps its
// strings' capacity for the next moira_format_records (fresh gigabytes of heap per batch cost more page faults than formatting)
std::mutex g_pool_mu;
std::vector<moira_blocks *> g_pool;
constexpr size_t POOL_MAX = 3;

inline void put_uint(std::string &o, uint64_t v)
{
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) o.push_back(tmp[--n]);
}

// header token with the reference's ':' -> '_' (moira.py:1121, 1175); leading '>' / '@' were cut by the parser
inline void put_header(std::string &o, const char *p, uint32_t n)
{
    const size_t at = o.size();
    o.append(p, n);
    for (size_t i = at; i < o.size(); i++)
        if (o[i] == ':') o[i] = '_';
}

// quality value as process_data returns it: Q if Q > 0 else 1 (moira.py:814); bytes 253..255 of a contig row stand for -3..-1
inline int qual_value(uint8_t b, int sub)
{
    int q = (int)b - sub;
    if (sub == 0 && b > 0xFC) q = (int)b - 256;
    return q > 0 ? q : 1;
}

}  // namespace

extern "C" {

int moira_format_records(const moira_records *rec, const moira_write_opts *opt, const uint64_t *sel, uint64_t n_sel, const double *ee,
                         const uint8_t *accept, const uint8_t *reason, const uint64_t *sel_group, const uint64_t *member_start,
                         const uint64_t *members, const int32_t *overlap, const int32_t *gaps, const int32_t *mismatches,
                         int n_threads, moira_blocks **out)
{
    using moira::fail;
    if (!rec || !opt || !out || (n_sel && (!accept || !reason || !ee))) return fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    if (n_sel && (!rec->hdr_off || !rec->hdr_len || !rec->seq_off || !rec->qual_off || !rec->len))
        return fail(MOIRA_ERR_BAD_ARG, "record arrays are NULL");
    // a NULL base makes the offsets absolute addresses (records spread over several buffers)
    const uintptr_t hb = (uintptr_t)rec->hdr_base, sb = (uintptr_t)rec->seq_base, qb = (uintptr_t)rec->qual_base;
    if (opt->names && (!sel_group || !member_start || !members)) return fail(MOIRA_ERR_BAD_ARG, "names blocks need the groups");
    *out = nullptr;
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    const int parts = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)T * 4, n_sel / 2048 + 1));
    moira_blocks *b = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (!g_pool.empty()) { b = g_pool.back(); g_pool.pop_back(); }
    }
    if (!b) b = new (std::nothrow) moira_blocks();
    if (!b) return fail(MOIRA_ERR_NOMEM, "out of host memory");
    b->n_parts = parts;
    if (b->buf.size() < (size_t)parts * MOIRA_BLOCK_N) b->buf.resize((size_t)parts * MOIRA_BLOCK_N);
    for (auto &str : b->buf) str.clear();   // keeps the capacity of a recycled batch
    const bool has_stats = overlap && gaps && mismatches;
    const size_t relabel_len = opt->relabel ? strlen(opt->relabel) : 0;
    bool failed = false;
    moira::parallel_run(parts, T, [&](int p) {
        try {
            std::string *o = &b->buf[(size_t)p * MOIRA_BLOCK_N];
            std::string hdr;
            char tmp[64];
            const uint64_t k0 = n_sel * (uint64_t)p / parts, k1 = n_sel * (uint64_t)(p + 1) / parts;
            {   // size the blocks once (growing them by doubling copies every byte again and faults fresh pages each time)
                uint64_t good_b = 0, bad_b = 0, good_n = 0, bad_n = 0, good_m = 0, bad_m = 0;
                for (uint64_t k = k0; k < k1; k++) {
                    const uint64_t r = sel ? sel[k] : k;
                    const uint64_t hl = (relabel_len ? relabel_len + 20 : rec->hdr_len[r]) + (opt->usearch ? 48 : 0) + 40;
                    uint64_t mem = 0;
                    if (opt->names) {
                        const uint64_t m0 = member_start[sel_group[k]], m1 = member_start[sel_group[k] + 1];
                        mem = hl + 2 + (m1 - m0) * (uint64_t)(rec->hdr_len[r] + 12);   // members' headers: about the representative's length
                    }
                    if (accept[r]) { good_b += rec->len[r]; good_n += hl; good_m += mem; }
                    else { bad_b += rec->len[r]; bad_n += hl; bad_m += mem; }
                }
                const uint64_t recs = k1 - k0;
                if (opt->fastq) {
                    o[MOIRA_BLOCK_GOOD].reserve(2 * good_b + good_n + 8 * recs);
                    o[MOIRA_BLOCK_BAD].reserve(2 * bad_b + bad_n + 8 * recs);
                } else {
                    o[MOIRA_BLOCK_GOOD].reserve(good_b + good_n + 4 * recs);
                    o[MOIRA_BLOCK_BAD].reserve(bad_b + bad_n + 4 * recs);
                    o[MOIRA_BLOCK_GOOD_QUAL].reserve(3 * good_b + good_b / 8 + good_n + 4 * recs);
                    o[MOIRA_BLOCK_BAD_QUAL].reserve(3 * bad_b + bad_b / 8 + bad_n + 4 * recs);
                }
                if (opt->names) {
                    o[MOIRA_BLOCK_GOOD_NAMES].reserve(good_m);
                    o[MOIRA_BLOCK_BAD_NAMES].reserve(bad_m);
                }
            }
            for (uint64_t k = k0; k < k1; k++) {
                const uint64_t r = sel ? sel[k] : k;
                const uint32_t L = rec->len[r];
                // ---- header as write_results builds it (moira.py:853-863) ----
                hdr.clear();
                if (relabel_len) { hdr.append(opt->relabel, relabel_len); put_uint(hdr, opt->first_index + k); }
                else put_header(hdr, (const char *)(hb + rec->hdr_off[r]), rec->hdr_len[r]);
                uint64_t size = 1, m0 = 0, m1 = 0;
                if (sel_group && member_start) { m0 = member_start[sel_group[k]]; m1 = member_start[sel_group[k] + 1]; size = m1 - m0; }
                if (opt->usearch) {
                    const int w = snprintf(tmp, sizeof(tmp), ";ee=%.2f;size=%llu;", ee[r], (unsigned long long)size);
                    hdr.append(tmp, (size_t)w);
                }
                if (has_stats) {   // contigs report, moira.py:866-870
                    std::string &rp = o[MOIRA_BLOCK_REPORT];
                    rp += hdr; rp.push_back('\t'); put_uint(rp, size);
                    const int w = snprintf(tmp, sizeof(tmp), "\t%d\t%d\t%d\n", overlap[r], gaps[r], mismatches[r]);
                    rp.append(tmp, (size_t)w);
                }
                const bool good = accept[r] != 0;
                const char *note = good ? "" : (opt->notes[reason[r] & 7] ? opt->notes[reason[r] & 7] : "");
                std::string &main = o[good ? MOIRA_BLOCK_GOOD : MOIRA_BLOCK_BAD];
                const char *s = (const char *)(sb + rec->seq_off[r]);
                const uint8_t *q = (const uint8_t *)(qb + rec->qual_off[r]);
                if (opt->fastq) {
                    main.push_back('@'); main += hdr; main += note; main.push_back('\n');
                    main.append(s, L);
                    main.append("\n+\n", 3);
                    const size_t at = main.size();
                    main.resize(at + L);
                    for (uint32_t i = 0; i < L; i++) main[at + i] = (char)(qual_value(q[i], rec->qual_sub) + opt->fastq_offset);
                    main.push_back('\n');
                } else {
                    std::string &ql = o[good ? MOIRA_BLOCK_GOOD_QUAL : MOIRA_BLOCK_BAD_QUAL];
                    main.push_back('>'); main += hdr; main += note; main.push_back('\n');
                    main.append(s, L);
                    main.push_back('\n');
                    ql.push_back('>'); ql += hdr; ql += note; ql.push_back('\n');
                    {   // "%d %d ... %d\n": at most four bytes per value
                        const size_t at = ql.size();
                        ql.resize(at + 4 * (size_t)L + 1);
                        char *w = &ql[at];
                        for (uint32_t i = 0; i < L; i++) {
                            if (i) *w++ = ' ';
                            const int v = qual_value(q[i], rec->qual_sub);
                            if (v >= 100) *w++ = (char)('0' + v / 100);
                            if (v >= 10) *w++ = (char)('0' + (v / 10) % 10);
                            *w++ = (char)('0' + v % 10);
                        }
                        *w++ = '\n';
                        ql.resize((size_t)(w - ql.data()));
                    }
                }
                if (opt->names) {   // "%s\t%s\n" % (header, ','.join(names_info)): the members' own headers, names order
                    std::string &nm = o[good ? MOIRA_BLOCK_GOOD_NAMES : MOIRA_BLOCK_BAD_NAMES];
                    nm += hdr; nm.push_back('\t');
                    for (uint64_t m = m0; m < m1; m++) {
                        if (m > m0) nm.push_back(',');
                        put_header(nm, (const char *)(hb + rec->hdr_off[members[m]]), rec->hdr_len[members[m]]);
                    }
                    nm.push_back('\n');
                }
            }
        } catch (...) {
            failed = true;
        }
    });
    if (failed) { delete b; return fail(MOIRA_ERR_NOMEM, "out of host memory while formatting records"); }
    *out = b;
    return MOIRA_OK;
}

int moira_blocks_parts(const moira_blocks *b, int *n_parts_out)
{
    if (!b || !n_parts_out) return moira::fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    *n_parts_out = b->n_parts;
    return MOIRA_OK;
}

int moira_blocks_get(const moira_blocks *b, int part, int which, const char **ptr_out, uint64_t *len_out)
{
    if (!b || !ptr_out || !len_out || part < 0 || part >= b->n_parts || which < 0 || which >= MOIRA_BLOCK_N)
        return moira::fail(MOIRA_ERR_BAD_ARG, "bad block index");
    const std::string &s = b->buf[(size_t)part * MOIRA_BLOCK_N + which];
    *ptr_out = s.data();
    *len_out = s.size();
    return MOIRA_OK;
}

// All parts of one block kind, in order, to file descriptor `fd` at `file_offset` (pwrite, the parts side by side on the
// host threads: page-cache copies scale with the writers).  The caller keeps the file position: *written_out is what to add.
int moira_blocks_write(const moira_blocks *b, int which, int fd, uint64_t file_offset, int n_threads, uint64_t *written_out)
{
    if (!b || !written_out || which < 0 || which >= MOIRA_BLOCK_N || fd < 0) return moira::fail(MOIRA_ERR_BAD_ARG, "bad argument");
    std::vector<uint64_t> at((size_t)b->n_parts + 1, file_offset);
    for (int p = 0; p < b->n_parts; p++) at[(size_t)p + 1] = at[p] + b->buf[(size_t)p * MOIRA_BLOCK_N + which].size();
    *written_out = at[b->n_parts] - file_offset;
    if (*written_out == 0) return MOIRA_OK;
    // own threads, not the library's pool: the pool belongs to the formatter, which is busy with the next batch meanwhile
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency() / 2;
    if (T < 1) T = 1;
    if (T > 8) T = 8;
    if (T > b->n_parts) T = b->n_parts;
    std::atomic<int> next{0}, err{0};
    auto work = [&]() {
        for (;;) {
            const int p = next.fetch_add(1);
            if (p >= b->n_parts) return;
            const std::string &s = b->buf[(size_t)p * MOIRA_BLOCK_N + which];
            size_t done = 0;
            while (done < s.size()) {
                const ssize_t w = pwrite(fd, s.data() + done, s.size() - done, (off_t)(at[p] + done));
                if (w < 0) {
                    if (errno == EINTR) continue;
                    err = errno;
                    return;
                }
                done += (size_t)w;
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work);
    work();
    for (auto &x : th) x.join();
    if (err) return moira::fail(MOIRA_ERR_BAD_ARG, "pwrite failed: %s", strerror(err.load()));
    return MOIRA_OK;
}

int moira_blocks_write_gz(const moira_blocks *b, int which, int fd, uint64_t file_offset, int level, int n_threads, uint64_t *written_out)
{
    if (!b || !written_out || which < 0 || which >= MOIRA_BLOCK_N || fd < 0) return moira::fail(MOIRA_ERR_BAD_ARG, "bad argument");
    std::vector<std::pair<const uint8_t *, uint64_t>> segs;   // parts in output order; the threads share the pieces of all of them
    for (int p = 0; p < b->n_parts; p++) {
        const std::string &s = b->buf[(size_t)p * MOIRA_BLOCK_N + which];
        if (!s.empty()) segs.emplace_back(reinterpret_cast<const uint8_t *>(s.data()), s.size());
    }
    return moira::gz_deflate_segments_to_fd(segs, level, n_threads, fd, file_offset, written_out);
}

// Hand a batch back for reuse by the next moira_format_records (at most a few are kept; the rest are freed).
int moira_blocks_recycle(moira_blocks *b)
{
    if (!b) return MOIRA_OK;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (g_pool.size() < POOL_MAX) { g_pool.push_back(b); return MOIRA_OK; }
    }
    delete b;
    return MOIRA_OK;
}

// Frees the batch and every recycled one.
int moira_blocks_free(moira_blocks *b)
{
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (moira_blocks *x : g_pool) delete x;
        g_pool.clear();
    }
    delete b;
    return MOIRA_OK;
}

// Header token of every FASTQ record from the position of its sequence line (what the device parser returns): the line
// before it, stripped, cut at the first blank, leading '@'s dropped (moira.py:1172-1175; ':' -> '_' is the writer's).
int moira_fastq_headers(const char *text, uint64_t text_bytes, const uint64_t *seq_off, uint64_t n, int n_threads, uint64_t *hdr_off,
                        uint32_t *hdr_len)
{
    if (n && (!text || !seq_off || !hdr_off || !hdr_len)) return moira::fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    const int parts = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)T * 4, n / 4096 + 1));
    bool bad = false;
    moira::parallel_run(parts, T, [&](int p) {
        auto blank = [](unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); };
        for (uint64_t r = n * (uint64_t)p / parts, e = n * (uint64_t)(p + 1) / parts; r < e; r++) {
            uint64_t s = seq_off[r];
            if (s > text_bytes) { bad = true; continue; }
            // back over the sequence line's leading blanks to the newline that ends the header line
            while (s > 0 && text[s - 1] != '\n') s--;
            uint64_t he = s > 0 ? s - 1 : 0;                       // the '\n'
            uint64_t hb = he;
            while (hb > 0 && text[hb - 1] != '\n') hb--;
            while (hb < he && blank((unsigned char)text[hb])) hb++;    // line.strip()
            while (he > hb && blank((unsigned char)text[he - 1])) he--;
            uint64_t t = hb;
            while (t < he && text[t] != ' ' && text[t] != '\t') t++;   // .replace('\t', ' ').split(' ')[0]
            while (hb < t && text[hb] == '@') hb++;                     // .lstrip('@')
            hdr_off[r] = hb;
            hdr_len[r] = (uint32_t)(t - hb);
        }
    });
    if (bad) return moira::fail(MOIRA_ERR_BAD_ARG, "a sequence offset lies outside the text");
    return MOIRA_OK;
}

}  // extern "C"
