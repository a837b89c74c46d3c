// moira_inflate.h -- a raw-DEFLATE (RFC 1951) decoder for the gzip inputs of moira_gz.cpp.
//
// zlib's inflate() delivers 0.15-0.25 GB/s per thread on the hosts this runs on, and a gzip FILE that is one member (what
// `gzip reads.fastq` writes) can only be inflated by one thread: it is the slowest stage of a run on such a file by an order
// of magnitude.  This decoder does the same work with the techniques of the fast inflaters -- a 64-bit bit buffer refilled
// with one unaligned load, single-lookup decode tables (11 bits literal / length, 8 bits distance, sub-tables behind them for
// longer codes), word-wise match copies -- and is resumable at any point where the output runs full, so the caller can grow
// its buffer.  The caller checks the member's CRC-32 and size afterwards (moira_gz.cpp) and falls back to zlib on any
// disagreement, so a defect here can cost time but not correctness.  Not installed; included by moira_gz.cpp only.
#pragma once
#include <stdint.h>
#include <string.h>

namespace moira_inflate {

constexpr int LITLEN_BITS = 11, DIST_BITS = 8;
constexpr int LITLEN_SYMS = 288, DIST_SYMS = 32, PRE_SYMS = 19;
// table sizes with sub-tables: ENOUGH-style bounds (zlib's enough 288 11 15 = 2342 -> round up; enough 32 8 15 = 402)
constexpr int LITLEN_ENTRIES = 2400, DIST_ENTRIES = 416;

// entry: bits 0..7 = bits to consume by this lookup (code length, or the table bits for a sub-table pointer),
//        bits 8..15 = kind / extra-bit count, bits 16..31 = value (literal, base length, base distance, sub-table start)
constexpr uint32_t KIND_LITERAL = 0x8000, KIND_EOB = 0x4000, KIND_SUB = 0x2000, KIND_BAD = 0x1000;   // else: base + extra bits (count in bits 8..11)

struct State {
    // input
    const uint8_t *in, *in_end;
    uint64_t bitbuf = 0;
    int bitcnt = 0;
    // block state
    int phase = 0;            // 0: block header next; 1: stored block in progress; 2: huffman block in progress; 3: stream finished
    bool last = false;
    uint32_t stored_left = 0;
    // a match interrupted by a full output buffer
    uint32_t pend_len = 0, pend_dist = 0;
    uint32_t litlen[LITLEN_ENTRIES];
    uint32_t dist[DIST_ENTRIES];
};

enum Result { DONE = 0, NEED_OUTPUT = 1, BAD_DATA = -1 };

inline uint64_t load64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

// Build a decode table for `n` symbols with code lengths `lens` (0 = unused): canonical codes (RFC 1951 3.2.2), indexed by the
// bit-reversed code as it arrives LSB first.  `sym_entry(sym)` gives the kind / value part of a symbol's entry.
template <typename F>
inline bool build_table(uint32_t *table, int table_bits, int max_entries, const uint8_t *lens, int n, F sym_entry)
{
    int count[16] = {0};
    for (int i = 0; i < n; i++) count[lens[i]]++;
    if (count[0] == n) {   // no codes at all (a distance tree of a block without matches): every lookup is an error
        for (int i = 0; i < (1 << table_bits); i++) table[i] = KIND_BAD | 1;
        return true;
    }
    // over-subscribed or incomplete sets: incomplete is legal only for a single code of length 1 (RFC 1951 / zlib's rule)
    int left = 1;
    for (int l = 1; l <= 15; l++) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return false;
    }
    int used = n - count[0];
    if (left > 0 && !(used == 1 && count[1] == 1)) return false;
    uint16_t next_code[16];
    {
        int code = 0;
        count[0] = 0;
        for (int l = 1; l <= 15; l++) { code = (code + count[l - 1]) << 1; next_code[l] = (uint16_t)code; }
    }
    const int main_size = 1 << table_bits;
    for (int i = 0; i < main_size; i++) table[i] = KIND_BAD | 1;
    int sub_next = main_size;
    // symbols in order of (length, symbol): assign codes; long codes go to sub-tables keyed by their low table_bits bits
    // first pass: how many extra bits each sub-table needs (the longest code sharing the prefix)
    // (two passes over the symbols keep this simple: sizes, then fill)
    auto reverse = [](uint32_t c, int l) { uint32_t r = 0; for (int i = 0; i < l; i++) { r = (r << 1) | (c & 1); c >>= 1; } return r; };
    uint16_t codes[LITLEN_SYMS];
    for (int s = 0; s < n; s++) if (lens[s]) codes[s] = next_code[lens[s]]++;
    // sub-table sizing
    static thread_local uint8_t sub_bits[1 << LITLEN_BITS];
    memset(sub_bits, 0, (size_t)1 << table_bits);
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (l > table_bits) {
            const uint32_t r = reverse(codes[s], l);
            const uint32_t pre = r & (uint32_t)(main_size - 1);
            if (l - table_bits > sub_bits[pre]) sub_bits[pre] = (uint8_t)(l - table_bits);
        }
    }
    for (int pre = 0; pre < main_size; pre++) {
        if (sub_bits[pre]) {
            const int size = 1 << sub_bits[pre];
            if (sub_next + size > max_entries) return false;
            table[pre] = ((uint32_t)sub_next << 16) | KIND_SUB | ((uint32_t)sub_bits[pre] << 8 & 0x0F00) | (uint32_t)table_bits;
            for (int i = 0; i < size; i++) table[sub_next + i] = KIND_BAD | 1;
            sub_next += size;
        }
    }
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t r = reverse(codes[s], l);
        const uint32_t e = sym_entry(s);
        if (l <= table_bits) {
            for (uint32_t i = r; i < (uint32_t)main_size; i += 1u << l) table[i] = e | (uint32_t)l;
        } else {
            const uint32_t pre = r & (uint32_t)(main_size - 1);
            const uint32_t start = table[pre] >> 16;
            const int sb = sub_bits[pre], rest = l - table_bits;
            for (uint32_t i = r >> table_bits; i < (1u << sb); i += 1u << rest) table[start + i] = e | (uint32_t)rest;
        }
    }
    return true;
}

inline uint32_t litlen_entry(int sym)
{
    static const uint16_t base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    if (sym < 256) return ((uint32_t)sym << 16) | KIND_LITERAL;
    if (sym == 256) return KIND_EOB;
    if (sym > 285) return KIND_BAD;
    return ((uint32_t)base[sym - 257] << 16) | ((uint32_t)extra[sym - 257] << 8);
}
inline uint32_t dist_entry(int sym)
{
    static const uint16_t base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    if (sym > 29) return KIND_BAD;
    return ((uint32_t)base[sym] << 16) | ((uint32_t)extra[sym] << 8);
}

// ---- bit input: bytes beyond in_end read as zero (a valid stream never consumes them; a truncated one ends as BAD_DATA) -----
inline void refill(State &s)
{
    if (s.in_end - s.in >= 8) {
        s.bitbuf |= load64(s.in) << s.bitcnt;
        s.in += (63 - s.bitcnt) >> 3;
        s.bitcnt |= 56;
    } else {
        while (s.bitcnt <= 56) {   // the last bytes one by one; behind them virtual zero bytes, counted so that over-reads are detected
            if (s.in < s.in_end) s.bitbuf |= (uint64_t)*s.in << s.bitcnt;
            s.in++;
            s.bitcnt += 8;
        }
    }
}
inline uint32_t peek(const State &s, int n) { return (uint32_t)(s.bitbuf & ((1ull << n) - 1)); }
inline void drop(State &s, int n) { s.bitbuf >>= n; s.bitcnt -= n; }
inline bool overran(const State &s) { return s.in > s.in_end && (int)(s.in - s.in_end) * 8 > s.bitcnt; }   // consumed bits that were never there

inline bool read_dynamic_header(State &s)
{
    refill(s);
    const int hlit = (int)peek(s, 5) + 257; drop(s, 5);
    const int hdist = (int)peek(s, 5) + 1; drop(s, 5);
    const int hclen = (int)peek(s, 4) + 4; drop(s, 4);
    if (hlit > 286 || hdist > 30) return false;
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t pre_lens[PRE_SYMS] = {0};
    for (int i = 0; i < hclen; i++) {
        if (s.bitcnt < 3) refill(s);
        pre_lens[order[i]] = (uint8_t)peek(s, 3);
        drop(s, 3);
    }
    uint32_t pre[128 + 8];
    if (!build_table(pre, 7, 128 + 8, pre_lens, PRE_SYMS, [](int sym) { return (uint32_t)sym << 16; })) return false;
    uint8_t lens[LITLEN_SYMS + DIST_SYMS] = {0};
    int i = 0;
    while (i < hlit + hdist) {
        refill(s);
        const uint32_t e = pre[peek(s, 7)];
        if (e & KIND_BAD) return false;
        drop(s, (int)(e & 0xFF));
        const int sym = (int)(e >> 16);
        if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
        int rep, val = 0;
        if (sym == 16) {
            if (i == 0) return false;
            val = lens[i - 1];
            rep = 3 + (int)peek(s, 2); drop(s, 2);
        } else if (sym == 17) { rep = 3 + (int)peek(s, 3); drop(s, 3); }
        else { rep = 11 + (int)peek(s, 7); drop(s, 7); }
        if (i + rep > hlit + hdist) return false;
        while (rep--) lens[i++] = (uint8_t)val;
    }
    if (overran(s) || lens[256] == 0) return false;
    if (!build_table(s.litlen, LITLEN_BITS, LITLEN_ENTRIES, lens, hlit, litlen_entry)) return false;
    return build_table(s.dist, DIST_BITS, DIST_ENTRIES, lens + hlit, hdist, dist_entry);
}

inline bool set_fixed_tables(State &s)
{
    uint8_t lens[LITLEN_SYMS];
    for (int i = 0; i < 144; i++) lens[i] = 8;
    for (int i = 144; i < 256; i++) lens[i] = 9;
    for (int i = 256; i < 280; i++) lens[i] = 7;
    for (int i = 280; i < 288; i++) lens[i] = 8;
    uint8_t dl[DIST_SYMS];
    for (int i = 0; i < 32; i++) dl[i] = 5;
    return build_table(s.litlen, LITLEN_BITS, LITLEN_ENTRIES, lens, 288, litlen_entry) &&
           build_table(s.dist, DIST_BITS, DIST_ENTRIES, dl, 32, dist_entry);
}

inline void start(State &s, const uint8_t *in, const uint8_t *in_end)
{
    s.in = in; s.in_end = in_end; s.bitbuf = 0; s.bitcnt = 0; s.phase = 0; s.last = false; s.stored_left = 0; s.pend_len = 0;
}
// bytes of input consumed so far, whole bytes still in the bit buffer given back (valid once DONE: the stream ends at a byte boundary rule of gzip)
inline const uint8_t *input_position(const State &s) { return s.in - (s.bitcnt >> 3); }

// copy a match of `len` bytes from `dist` back; out_start is the start of the whole output (the window)
inline void copy_match(uint8_t *out, uint32_t dist, uint32_t len, bool room16)
{
    const uint8_t *src = out - dist;
    if (room16 && dist >= 16) {
        uint8_t *end = out + len;
        do { memcpy(out, src, 16); out += 16; src += 16; } while (out < end);    // may write up to 15 bytes past the match: room16 guarantees the space
    } else if (room16 && dist >= 8) {
        uint8_t *end = out + len;
        do { memcpy(out, src, 8); out += 8; src += 8; } while (out < end);
    } else if (room16 && dist == 1) {
        memset(out, *src, len);
    } else {
        for (uint32_t i = 0; i < len; i++) out[i] = src[i];
    }
}

// Inflate from the state's input into [out, out_end), the window being [out_start, out).  Returns DONE at the end of the
// DEFLATE stream (*out_pos = end of output), NEED_OUTPUT when the buffer is full (call again with more room: same out_start
// contents, new pointers), BAD_DATA for an invalid or truncated stream.
inline Result run(State &s, uint8_t *out_start, uint8_t *out, uint8_t *out_end, uint8_t **out_pos)
{
    for (;;) {
        if (s.pend_len) {   // finish an interrupted match
            const uint32_t can = (uint32_t)((out_end - out) < (ptrdiff_t)s.pend_len ? (out_end - out) : s.pend_len);
            copy_match(out, s.pend_dist, can, false);
            out += can; s.pend_len -= can;
            if (s.pend_len) { *out_pos = out; return NEED_OUTPUT; }
        }
        if (s.phase == 3) { *out_pos = out; return DONE; }
        if (s.phase == 0) {
            if (s.last) { s.phase = 3; continue; }
            refill(s);
            s.last = peek(s, 1) != 0; drop(s, 1);
            const uint32_t type = peek(s, 2); drop(s, 2);
            if (type == 0) {
                drop(s, s.bitcnt & 7);                       // to the byte boundary
                refill(s);
                const uint32_t len = peek(s, 16); drop(s, 16);
                const uint32_t nlen = peek(s, 16); drop(s, 16);
                if ((len ^ nlen) != 0xFFFFu || overran(s)) return BAD_DATA;
                s.stored_left = len;
                s.phase = 1;
            } else if (type == 1) {
                if (!set_fixed_tables(s)) return BAD_DATA;
                s.phase = 2;
            } else if (type == 2) {
                if (!read_dynamic_header(s)) return BAD_DATA;
                s.phase = 2;
            } else return BAD_DATA;
        }
        if (s.phase == 1) {
            // stored bytes: first what sits in the bit buffer (whole bytes), then straight from the input
            while (s.stored_left && s.bitcnt >= 8) {
                if (out == out_end) { *out_pos = out; return NEED_OUTPUT; }
                *out++ = (uint8_t)s.bitbuf; drop(s, 8); s.stored_left--;
            }
            if (s.stored_left) {
                s.bitbuf = 0; s.bitcnt = 0;
                if (s.in > s.in_end) return BAD_DATA;
                const size_t avail = (size_t)(s.in_end - s.in), room = (size_t)(out_end - out);
                size_t n = s.stored_left;
                if (n > avail) return BAD_DATA;
                if (n > room) n = room;
                memcpy(out, s.in, n);
                out += n; s.in += n; s.stored_left -= (uint32_t)n;
                if (s.stored_left) { *out_pos = out; return NEED_OUTPUT; }
            }
            s.phase = 0;
            continue;
        }
        // ---- Huffman block ----
        // Fast loop: while at least 16 real input bytes and 274 output bytes are left, nothing needs a bounds check, and the
        // bit buffer, the pointers and the table bases live in locals (byte stores through `out` may alias the state, so working
        // on the state itself reloads every field after every literal).
        {
            uint64_t bb = s.bitbuf;
            int bc = s.bitcnt;
            const uint8_t *in = s.in;
            const uint8_t *const in_fast = s.in_end - 16;
            const uint32_t *const lt = s.litlen, *const dt = s.dist;
            bool eob = false, bad = false;
            while (out_end - out >= 258 + 16 && in <= in_fast && s.in_end - s.in >= 16) {
                bb |= load64(in) << bc;
                in += (63 - bc) >> 3;
                bc |= 56;
                uint32_t e = lt[bb & ((1u << LITLEN_BITS) - 1)];
                if (e & KIND_LITERAL) {
                    bb >>= (e & 0xFF); bc -= (int)(e & 0xFF);
                    *out++ = (uint8_t)(e >> 16);
                    e = lt[bb & ((1u << LITLEN_BITS) - 1)];
                    if (e & KIND_LITERAL) {
                        bb >>= (e & 0xFF); bc -= (int)(e & 0xFF);
                        *out++ = (uint8_t)(e >> 16);
                        e = lt[bb & ((1u << LITLEN_BITS) - 1)];
                        if (e & KIND_LITERAL) {
                            bb >>= (e & 0xFF); bc -= (int)(e & 0xFF);
                            *out++ = (uint8_t)(e >> 16);
                        }
                    }
                    continue;
                }
                if (e & KIND_SUB) {
                    bb >>= LITLEN_BITS; bc -= LITLEN_BITS;
                    e = lt[(e >> 16) + (uint32_t)(bb & ((1u << ((e >> 8) & 0xF)) - 1))];
                    if (e & KIND_LITERAL) {
                        bb >>= (e & 0xFF); bc -= (int)(e & 0xFF);
                        *out++ = (uint8_t)(e >> 16);
                        continue;
                    }
                }
                if (e & KIND_BAD) { bad = true; break; }
                bb >>= (e & 0xFF); bc -= (int)(e & 0xFF);
                if (e & KIND_EOB) { eob = true; break; }
                const int lx = (int)((e >> 8) & 0xF);
                const uint32_t len = (e >> 16) + (uint32_t)(bb & ((1u << lx) - 1));
                bb >>= lx; bc -= lx;                       // <= 15 + 11 + 5 bits gone: at least 25 left, the distance needs up to 28
                if (bc < 15 + 13) { bb |= load64(in) << bc; in += (63 - bc) >> 3; bc |= 56; }
                uint32_t d = dt[bb & ((1u << DIST_BITS) - 1)];
                if (d & KIND_SUB) {
                    bb >>= DIST_BITS; bc -= DIST_BITS;
                    d = dt[(d >> 16) + (uint32_t)(bb & ((1u << ((d >> 8) & 0xF)) - 1))];
                }
                if (d & (KIND_BAD | KIND_LITERAL | KIND_EOB)) { bad = true; break; }
                bb >>= (d & 0xFF); bc -= (int)(d & 0xFF);
                const int dx = (int)((d >> 8) & 0xF);
                const uint32_t dist = (d >> 16) + (uint32_t)(bb & ((1u << dx) - 1));
                bb >>= dx; bc -= dx;
                if (dist > (size_t)(out - out_start)) { bad = true; break; }
                copy_match(out, dist, len, true);
                out += len;
            }
            s.bitbuf = bb; s.bitcnt = bc; s.in = in;
            if (bad) return BAD_DATA;
            if (eob) { s.phase = 0; continue; }
        }
        for (;;) {
            const bool room = out_end - out >= 258 + 16;      // fast copies allowed
            if (s.in > s.in_end && overran(s)) return BAD_DATA;   // bits that were never there: a truncated stream (also ends garbage)
            refill(s);
            uint32_t e = s.litlen[peek(s, LITLEN_BITS)];
            if (e & KIND_SUB) {
                if (out == out_end) {   // full output: only an end-of-block symbol may follow; look without consuming
                    const uint32_t e2 = s.litlen[(e >> 16) + ((uint32_t)(s.bitbuf >> LITLEN_BITS) & ((1u << ((e >> 8) & 0xF)) - 1))];
                    if (!(e2 & KIND_EOB)) { *out_pos = out; return (e2 & KIND_BAD) ? BAD_DATA : NEED_OUTPUT; }
                }
                drop(s, LITLEN_BITS);
                e = s.litlen[(e >> 16) + peek(s, (int)((e >> 8) & 0xF))];
            } else if (out == out_end && !(e & KIND_EOB)) {
                *out_pos = out;
                return (e & KIND_BAD) ? BAD_DATA : NEED_OUTPUT;
            }
            if (e & KIND_LITERAL) {
                drop(s, (int)(e & 0xFF));
                *out++ = (uint8_t)(e >> 16);
                if (room) {   // up to two more literals on the same refill (>= 56 - 15 bits were there)
                    uint32_t e2 = s.litlen[peek(s, LITLEN_BITS)];
                    if (e2 & KIND_LITERAL) {
                        drop(s, (int)(e2 & 0xFF));
                        *out++ = (uint8_t)(e2 >> 16);
                        e2 = s.litlen[peek(s, LITLEN_BITS)];
                        if (e2 & KIND_LITERAL) {
                            drop(s, (int)(e2 & 0xFF));
                            *out++ = (uint8_t)(e2 >> 16);
                        }
                    }
                }
                continue;
            }
            if (e & KIND_BAD) return BAD_DATA;
            drop(s, (int)(e & 0xFF));
            if (e & KIND_EOB) {
                if (overran(s)) return BAD_DATA;
                s.phase = 0;
                break;
            }
            const int lx = (int)((e >> 8) & 0xF);
            const uint32_t len = (e >> 16) + peek(s, lx);
            drop(s, lx);
            if (s.bitcnt < 15 + 13) refill(s);
            uint32_t d = s.dist[peek(s, DIST_BITS)];
            if (d & KIND_SUB) {
                drop(s, DIST_BITS);
                d = s.dist[(d >> 16) + peek(s, (int)((d >> 8) & 0xF))];
            }
            if (d & (KIND_BAD | KIND_LITERAL | KIND_EOB)) return BAD_DATA;
            drop(s, (int)(d & 0xFF));
            const int dx = (int)((d >> 8) & 0xF);
            const uint32_t dist = (d >> 16) + peek(s, dx);
            drop(s, dx);
            if (dist > (size_t)(out - out_start) || overran(s)) return BAD_DATA;
            if (room) {
                copy_match(out, dist, len, true);
                out += len;
            } else {
                const uint32_t can = (uint32_t)((out_end - out) < (ptrdiff_t)len ? (out_end - out) : len);
                copy_match(out, dist, can, false);
                out += can;
                if (can < len) { s.pend_len = len - can; s.pend_dist = dist; *out_pos = out; return NEED_OUTPUT; }
            }
        }
    }
}

}  // namespace moira_inflate
