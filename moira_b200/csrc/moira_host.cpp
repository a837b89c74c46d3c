// moira_host.cpp -- host-side packers/parsers of the C ABI (no GPU work): turn the reference's
// per-read (sequence, integer qualities) into the in-band uint8 slab the kernels consume.
//   moira_pack_reads   <- what bernoullimodule.c:92-108 does per call (list -> int[]), batched
//   moira_parse_fastq  <- record semantics of parse_fastq, moira/moira.py:1152-1204
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

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

extern "C" int moira_parse_fastq(const char *text, uint64_t text_bytes, int fastq_offset, int lower_n_ambiguous,
                                 uint8_t *slab, uint64_t slab_capacity, uint64_t *out_offsets, uint32_t *lengths,
                                 uint64_t *hdr_off, uint32_t *hdr_len, uint64_t *seq_off, uint64_t *qual_off,
                                 uint64_t max_reads,
                                 uint64_t *n_reads_out, uint64_t *slab_bytes_out)
{
    if (!text || !n_reads_out || !slab_bytes_out) return hfail(MOIRA_ERR_BAD_ARG, "NULL argument");
    uint64_t pos = 0, n = 0, cur = 0;
    uint64_t lb[4], le[4];   // stripped [begin, end) of the 4 lines of the current record
    int have = 0;
    while (cur < text_bytes) {
        const char *nl = (const char *)memchr(text + cur, '\n', text_bytes - cur);
        uint64_t end = nl ? (uint64_t)(nl - text) : text_bytes;
        uint64_t b = cur, e = end;
        while (b < e && is_space((unsigned char)text[b])) b++;          // line.strip(), moira.py:1172
        while (e > b && is_space((unsigned char)text[e - 1])) e--;
        lb[have] = b; le[have] = e;
        have++;
        cur = nl ? end + 1 : text_bytes;
        if (have < 4) continue;
        have = 0;
        // header token: replace('\t',' ').split(' ')[0].lstrip('@')  (moira.py:1175); ':' -> '_' is left to the caller
        uint64_t hb = lb[0], he = lb[0];
        while (he < le[0] && text[he] != ' ' && text[he] != '\t') he++;
        while (hb < he && text[hb] == '@') hb++;
        const uint64_t slen = le[1] - lb[1], qlen = le[3] - lb[3];
        if (slen == 0) return hfail(MOIRA_ERR_PARSE, "EmptySeqError: record %llu (%.*s) has an empty sequence", (unsigned long long)n, (int)(he - hb), text + hb);
        if (qlen == 0) return hfail(MOIRA_ERR_PARSE, "EmptyQualError: record %llu (%.*s) has no qualities", (unsigned long long)n, (int)(he - hb), text + hb);
        if (slen != qlen) return hfail(MOIRA_ERR_PARSE, "LengthMismatchError: record %llu (%.*s): %llu bases, %llu qualities", (unsigned long long)n, (int)(he - hb), text + hb, (unsigned long long)slen, (unsigned long long)qlen);
        if (slen > 0xFFFFFFF0ull) return hfail(MOIRA_ERR_PARSE, "record %llu too long", (unsigned long long)n);
        const uint64_t padded = (slen + 15u) & ~15ull;
        if (slab) {
            if (n >= max_reads) return hfail(MOIRA_ERR_BAD_ARG, "more than max_reads = %llu records", (unsigned long long)max_reads);
            if (pos + padded > slab_capacity) return hfail(MOIRA_ERR_BAD_ARG, "slab capacity too small");
            uint8_t *row = slab + pos;
            const char *s = text + lb[1];
            const unsigned char *q = (const unsigned char *)text + lb[3];
            for (uint64_t i = 0; i < slen; i++) {
                if (s[i] == 'N') row[i] = 0xFF;
                else if (s[i] == 'n' && lower_n_ambiguous) row[i] = 0xFE;
                else {
                    const int v = (int)q[i] - fastq_offset;                 // moira.py:1177
                    if (v > 0xFC) return hfail(MOIRA_ERR_BAD_QUALITY, "quality %d in record %llu is outside 0..252", v, (unsigned long long)n);
                    row[i] = v <= 0 ? 0 : (uint8_t)v;                        // moira.py:814
                }
            }
            memset(row + slen, 0xFD, padded - slen);
            if (out_offsets) out_offsets[n] = pos;
            if (lengths) lengths[n] = (uint32_t)slen;
            if (hdr_off) hdr_off[n] = hb;
            if (hdr_len) hdr_len[n] = (uint32_t)(he - hb);
            if (seq_off) seq_off[n] = lb[1];
            if (qual_off) qual_off[n] = lb[3];
        }
        pos += padded;
        n++;
    }
    *n_reads_out = n;
    *slab_bytes_out = pos;
    return MOIRA_OK;
}
