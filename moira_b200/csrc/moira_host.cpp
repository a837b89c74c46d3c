// moira_host.cpp -- host-side packers/parsers of the C ABI (no GPU work): turn the reference's
// per-read (sequence, integer qualities) into the in-band uint8 slab the kernels consume.
//   moira_pack_reads   <- what bernoullimodule.c:92-108 does per call (list -> int[]), batched
//   moira_parse_fastq  <- record semantics of parse_fastq, moira/moira.py:1152-1204
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <thread>
#include <vector>

#include "moira_internal.h"

#define hfail moira::fail

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
    uint64_t n_reads = 0, slab_bytes = 0;
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

// Parse the records owned by one segment.  fill == false: count reads / slab bytes and validate.
void fq_parse_segment(const FqJob &j, FqSeg &g, bool fill, uint64_t next_begin)
{
    const char *text = j.text;
    uint64_t cur = g.first_line_pos;
    uint64_t line = g.first_line;
    // advance to the first record start (line index multiple of 4) inside the segment
    while (cur < j.text_bytes && (line & 3u) != 0) {
        const char *nl = (const char *)memchr(text + cur, '\n', j.text_bytes - cur);
        if (!nl) { cur = j.text_bytes; break; }
        cur = (uint64_t)(nl - text) + 1;
        line++;
    }
    uint64_t n = 0, pos = 0;
    while (cur < j.text_bytes && cur < next_begin) {
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
        if (fill) {
            const uint64_t r = g.read_base + n;
            uint8_t *row = j.slab + g.slab_base + pos;
            const char *s = text + lb[1];
            const unsigned char *q = (const unsigned char *)text + lb[3];
            bool bad_q = false;
            for (uint64_t i = 0; i < slen; i++) {
                if (s[i] == 'N') row[i] = 0xFF;
                else if (s[i] == 'n' && j.lower_n) row[i] = 0xFE;
                else {
                    const int v = (int)q[i] - j.fastq_offset;                 // moira.py:1177
                    if (v > 0xFC) bad_q = true;
                    row[i] = v <= 0 ? 0 : (uint8_t)v;                        // moira.py:814
                }
            }
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
    }
    if (!fill) { g.n_reads = n; g.slab_bytes = pos; }
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
    if (!text || !n_reads_out || !slab_bytes_out) return hfail(MOIRA_ERR_BAD_ARG, "NULL argument");
    int T = g_host_threads > 0 ? g_host_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    if (text_bytes < (1u << 20)) T = 1;
    std::vector<FqSeg> segs(T);
    for (int t = 0; t < T; t++) {
        segs[t].begin = text_bytes * (uint64_t)t / T;
        segs[t].end = text_bytes * (uint64_t)(t + 1) / T;
    }
    auto run = [&](auto &&fn) {
        if (T == 1) { fn(0); return; }
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(fn, t);
        for (auto &x : th) x.join();
    };
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
    FqJob job{text, text_bytes, fastq_offset, lower_n_ambiguous, slab, out_offsets, hdr_off, seq_off, qual_off, lengths, hdr_len};
    auto next_begin = [&](int t) { return t + 1 < T ? segs[t + 1].first_line_pos : text_bytes; };
    // a thread owns records whose first line starts in [first_line_pos(t), first_line_pos(t+1))
    run([&](int t) { if (segs[t].first_line_pos < next_begin(t) || t == T - 1) fq_parse_segment(job, segs[t], false, next_begin(t)); });
    uint64_t n = 0, bytes = 0;
    for (int t = 0; t < T; t++) {
        if (segs[t].err) return hfail(segs[t].err, "%s (record %llu)", segs[t].msg, (unsigned long long)(n + segs[t].err_local));
        segs[t].read_base = n; segs[t].slab_base = bytes;
        n += segs[t].n_reads; bytes += segs[t].slab_bytes;
    }
    *n_reads_out = n;
    *slab_bytes_out = bytes;
    if (!slab) return MOIRA_OK;
    if (n > max_reads) return hfail(MOIRA_ERR_BAD_ARG, "more than max_reads = %llu records", (unsigned long long)max_reads);
    if (bytes > slab_capacity) return hfail(MOIRA_ERR_BAD_ARG, "slab capacity too small");
    run([&](int t) { if (segs[t].n_reads) fq_parse_segment(job, segs[t], true, next_begin(t)); });
    for (int t = 0; t < T; t++)
        if (segs[t].err) return hfail(segs[t].err, "%s (record %llu)", segs[t].msg, (unsigned long long)(segs[t].read_base + segs[t].err_local));
    return MOIRA_OK;
}
