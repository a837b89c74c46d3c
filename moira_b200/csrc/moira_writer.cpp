// moira_writer.cpp -- native formatting of the output records: what write_results prints for every
// read or collapsed group (moira/moira.py:842-970) -- fasta + qual or fastq records, mothur .names lines, the USEARCH
// ";ee=%.2f;size=%d;" header, --relabel, the rejection reason behind the header, the contigs report -- produced from the
// byte ranges the parsers (host or device) and the contig kernel left behind, on all host threads, into memory
// blocks the host program only has to write to its files.  The accept / reject decision itself is the device's.
#include <stdio.h>
#include <string.h>

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
    moira_blocks *b = new (std::nothrow) moira_blocks();
    if (!b) return fail(MOIRA_ERR_NOMEM, "out of host memory");
    b->n_parts = parts;
    b->buf.resize((size_t)parts * MOIRA_BLOCK_N);
    const bool has_stats = overlap && gaps && mismatches;
    const size_t relabel_len = opt->relabel ? strlen(opt->relabel) : 0;
    bool failed = false;
    moira::parallel_run(parts, T, [&](int p) {
        try {
            std::string *o = &b->buf[(size_t)p * MOIRA_BLOCK_N];
            std::string hdr;
            char tmp[64];
            const uint64_t k0 = n_sel * (uint64_t)p / parts, k1 = n_sel * (uint64_t)(p + 1) / parts;
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
                    for (uint32_t i = 0; i < L; i++) {
                        if (i) ql.push_back(' ');
                        const int v = qual_value(q[i], rec->qual_sub);
                        if (v >= 100) ql.push_back((char)('0' + v / 100));
                        if (v >= 10) ql.push_back((char)('0' + (v / 10) % 10));
                        ql.push_back((char)('0' + v % 10));
                    }
                    ql.push_back('\n');
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

int moira_blocks_free(moira_blocks *b)
{
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
