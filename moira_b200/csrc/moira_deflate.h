// moira_deflate.h -- a fast raw-DEFLATE (RFC 1951) compressor for the gzip outputs of moira_gz.cpp (level 1).
//
// zlib's deflate() runs at 18 MB/s per thread at level 6 (60 at level 1) on the hosts this runs on: with --output_compression gz
// the compressor, not the filter, the parser or the formatter, is what a run waits for.  One call compresses one piece of at
// most 65 535 bytes (a BGZF member's payload) into ONE dynamic-Huffman block: greedy LZ77 with a single-probe hash table of
// 4-byte strings (positions fit 16 bits: the piece is the whole window), symbol statistics, length-limited Huffman codes
// (two-queue construction; frequencies are halved until no code exceeds its limit), RFC 1951 3.2.7 header with the run-length
// alphabet.  The caller (moira_gz.cpp) inflates every piece again and compares before it is written, and hands the piece to
// zlib if anything is off or it did not shrink.  Not installed; included by moira_gz.cpp only.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>

namespace moira_deflate {

constexpr int HASH_BITS = 15;
constexpr uint32_t MAX_PIECE = 65535;

struct Workspace {
    uint16_t head[1 << HASH_BITS];       // hash of 4 bytes -> position + 1 (0: empty)
    uint16_t sym_lit[MAX_PIECE + 1];     // per symbol: literal byte, or match length (3..258)
    uint16_t sym_dist[MAX_PIECE + 1];    // 0 = literal, else match distance
    uint32_t n_sym;
    uint32_t lfreq[288], dfreq[32];
    uint8_t llen[288], dlen[32];
    uint16_t lcode[288], dcode[32];
};

struct BitWriter {
    uint8_t *p, *end;
    uint64_t acc = 0;
    int n = 0;
    bool overflow = false;
    inline void put(uint32_t v, int bits)
    {
        acc |= (uint64_t)v << n;
        n += bits;
        if (n >= 32) {
            if (end - p < 4) { overflow = true; n -= 32; acc >>= 32; return; }
            const uint32_t w = (uint32_t)acc;
            memcpy(p, &w, 4);
            p += 4;
            acc >>= 32;
            n -= 32;
        }
    }
    inline void finish()
    {
        while (n > 0) {
            if (p >= end) { overflow = true; return; }
            *p++ = (uint8_t)acc;
            acc >>= 8;
            n -= 8;
        }
        n = 0;
    }
};

inline uint32_t load32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t load64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

// length 3..258 -> symbol 257..285, extra bits; distance 1..32768 -> symbol 0..29, extra bits
struct Tables {
    uint8_t len_sym[259], len_xbits[259];
    uint16_t len_base[259];
    uint8_t dist_sym_lo[256], dist_sym_hi[256];
    uint8_t dist_xbits[30];
    uint16_t dist_base[30];
    Tables()
    {
        static const uint16_t lb[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint8_t lx[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        for (int s = 0; s < 29; s++) {
            const int hi = s == 28 ? 258 : (s + 1 < 29 ? lb[s + 1] - 1 : 258);
            for (int l = lb[s]; l <= hi && l <= 258; l++) { len_sym[l] = (uint8_t)s; len_xbits[l] = lx[s]; len_base[l] = lb[s]; }
        }
        len_sym[258] = 28; len_xbits[258] = 0; len_base[258] = 258;
        static const uint16_t db[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        static const uint8_t dx[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        for (int s = 0; s < 30; s++) { dist_xbits[s] = dx[s]; dist_base[s] = db[s]; }
        for (int d = 1; d <= 256; d++) { int s = 0; while (s + 1 < 30 && db[s + 1] <= d) s++; dist_sym_lo[d - 1] = (uint8_t)s; }
        for (int k = 0; k < 256; k++) {   // distances 257..32768 by (d - 1) >> 7
            const int d = (k << 7) + 1;
            int s = 0;
            while (s + 1 < 30 && db[s + 1] <= d) s++;
            dist_sym_hi[k] = (uint8_t)s;
        }
    }
    inline int dist_symbol(uint32_t d) const { return d <= 256 ? dist_sym_lo[d - 1] : dist_sym_hi[(d - 1) >> 7]; }
};
inline const Tables &tables() { static const Tables t; return t; }

// Code lengths of a Huffman code for `n` symbols with frequencies `freq`, none longer than `maxbits`.
inline void huffman_lengths(const uint32_t *freq_in, int n, int maxbits, uint8_t *lens)
{
    struct Node { uint32_t f; int16_t left, right; };
    uint32_t freq[288];
    for (int i = 0; i < n; i++) freq[i] = freq_in[i];
    for (;;) {
        int order[288], m = 0;
        for (int i = 0; i < n; i++) { lens[i] = 0; if (freq[i]) order[m++] = i; }
        if (m == 0) return;
        if (m == 1) { lens[order[0]] = 1; return; }
        std::sort(order, order + m, [&](int a, int b) { return freq[a] != freq[b] ? freq[a] < freq[b] : a < b; });
        Node nodes[2 * 288];
        for (int i = 0; i < m; i++) nodes[i] = Node{freq[order[i]], -1, -1};
        int leaf = 0, inner = m, made = m;          // two queues: sorted leaves [leaf, m), inner nodes [inner, made) in creation (= weight) order
        auto take = [&]() {
            if (leaf < m && (inner >= made || nodes[leaf].f <= nodes[inner].f)) return leaf++;
            return inner++;
        };
        while ((m - leaf) + (made - inner) > 1) {
            const int a = take(), b = take();
            nodes[made] = Node{nodes[a].f + nodes[b].f, (int16_t)a, (int16_t)b};
            made++;
        }
        // depths: parents come after their children, so walk from the root down
        uint8_t depth[2 * 288];
        if (made < 2 || made > 2 * 288) return;      // (m >= 2 here, so made = 2 m - 1 >= 3: for the compiler's range analysis)
        depth[made - 1] = 0;
        int deepest = 0;
        for (int i = made - 1; i >= m; i--) {
            depth[nodes[i].left] = depth[nodes[i].right] = (uint8_t)(depth[i] + 1);
            if (depth[i] + 1 > deepest) deepest = depth[i] + 1;
        }
        if (deepest <= maxbits) {
            for (int i = 0; i < m; i++) lens[order[i]] = depth[i];
            return;
        }
        for (int i = 0; i < n; i++) if (freq[i]) freq[i] = (freq[i] + 1) >> 1;   // flatter statistics, shallower tree
    }
}

// canonical codes, bit-reversed for LSB-first output
inline void assign_codes(const uint8_t *lens, int n, uint16_t *codes)
{
    int count[16] = {0}, next[16];
    for (int i = 0; i < n; i++) count[lens[i]]++;
    count[0] = 0;
    int code = 0;
    for (int l = 1; l <= 15; l++) { code = (code + count[l - 1]) << 1; next[l] = code; }
    for (int i = 0; i < n; i++) {
        const int l = lens[i];
        if (!l) { codes[i] = 0; continue; }
        uint32_t c = (uint32_t)next[l]++, r = 0;
        for (int k = 0; k < l; k++) { r = (r << 1) | (c & 1); c >>= 1; }
        codes[i] = (uint16_t)r;
    }
}

// Compress in[0, n) (n <= 65535) into one final DEFLATE block at out[0, cap).  Returns the bytes written, 0 if it did not fit.
inline size_t compress_piece(const uint8_t *in, uint32_t n, uint8_t *out, size_t cap, Workspace &w)
{
    if (n > MAX_PIECE) return 0;
    if (n == 0) {   // an empty final block with the fixed code: BFINAL = 1, BTYPE = 01, end-of-block (7 zero bits)
        if (cap < 2) return 0;
        out[0] = 0x03; out[1] = 0x00;
        return 2;
    }
    const Tables &T = tables();
    memset(w.head, 0, sizeof(w.head));
    memset(w.lfreq, 0, sizeof(w.lfreq));
    memset(w.dfreq, 0, sizeof(w.dfreq));
    uint32_t ns = 0, i = 0;
    // ---- greedy LZ77, one probe per position ----
    const uint32_t last_hashable = n >= 4 ? n - 4 : 0;
    while (n >= 4 && i <= last_hashable) {
        const uint32_t v = load32(in + i);
        const uint32_t h = (v * 2654435761u) >> (32 - HASH_BITS);
        const uint32_t cand1 = w.head[h];
        w.head[h] = (uint16_t)(i + 1);
        if (cand1 && load32(in + cand1 - 1) == v && i - (cand1 - 1) <= 32768) {
            const uint32_t c = cand1 - 1;
            uint32_t len = 4;
            const uint32_t maxlen = std::min<uint32_t>(258, n - i);
            while (len + 8 <= maxlen) {
                const uint64_t x = load64(in + i + len) ^ load64(in + c + len);
                if (x) { len += (uint32_t)(__builtin_ctzll(x) >> 3); goto matched; }
                len += 8;
            }
            while (len < maxlen && in[i + len] == in[c + len]) len++;
        matched:
            w.sym_lit[ns] = (uint16_t)len; w.sym_dist[ns] = (uint16_t)(i - c); ns++;
            w.lfreq[257 + T.len_sym[len]]++;
            w.dfreq[T.dist_symbol(i - c)]++;
            // a few positions inside the match stay findable (all of them would cost more than they find)
            const uint32_t stop = std::min(i + len, last_hashable + 1);
            for (uint32_t k = i + 1; k < stop && k < i + 8; k++) w.head[(load32(in + k) * 2654435761u) >> (32 - HASH_BITS)] = (uint16_t)(k + 1);
            i += len;
        } else {
            w.sym_lit[ns] = in[i]; w.sym_dist[ns] = 0; ns++;
            w.lfreq[in[i]]++;
            i++;
        }
    }
    for (; i < n; i++) { w.sym_lit[ns] = in[i]; w.sym_dist[ns] = 0; ns++; w.lfreq[in[i]]++; }
    w.lfreq[256] = 1;
    w.n_sym = ns;
    // ---- codes ----
    huffman_lengths(w.lfreq, 286, 15, w.llen);
    huffman_lengths(w.dfreq, 30, 15, w.dlen);
    // a complete code is required for literals / lengths (at least two symbols are in use: a literal and end-of-block);
    // one or no distance code: give symbol 0 (and 1) a length so that every decoder accepts the tree
    {
        int used = 0;
        for (int s = 0; s < 30; s++) used += w.dlen[s] != 0;
        if (used == 0) { w.dlen[0] = 1; w.dlen[1] = 1; }
        else if (used == 1) { for (int s = 0; s < 30; s++) if (!w.dlen[s]) { w.dlen[s] = 1; break; } }
    }
    assign_codes(w.llen, 286, w.lcode);
    assign_codes(w.dlen, 30, w.dcode);
    int hlit = 286, hdist = 30;
    while (hlit > 257 && w.llen[hlit - 1] == 0) hlit--;
    while (hdist > 1 && w.dlen[hdist - 1] == 0) hdist--;
    // ---- header: code lengths in the run-length alphabet (RFC 1951 3.2.7) ----
    uint8_t seq[286 + 30];
    for (int s = 0; s < hlit; s++) seq[s] = w.llen[s];
    for (int s = 0; s < hdist; s++) seq[hlit + s] = w.dlen[s];
    const int total = hlit + hdist;
    uint8_t rl_sym[316], rl_extra[316];
    int nrl = 0;
    uint32_t pfreq[19] = {0};
    for (int k = 0; k < total;) {
        const int v = seq[k];
        int run = 1;
        while (k + run < total && seq[k + run] == v) run++;
        int left = run;
        if (v == 0) {
            while (left >= 11) { const int r = std::min(left, 138); rl_sym[nrl] = 18; rl_extra[nrl++] = (uint8_t)(r - 11); pfreq[18]++; left -= r; }
            if (left >= 3) { rl_sym[nrl] = 17; rl_extra[nrl++] = (uint8_t)(left - 3); pfreq[17]++; left = 0; }
            while (left-- > 0) { rl_sym[nrl] = 0; rl_extra[nrl++] = 0; pfreq[0]++; }
        } else {
            rl_sym[nrl] = (uint8_t)v; rl_extra[nrl++] = 0; pfreq[v]++; left--;
            while (left >= 3) { const int r = std::min(left, 6); rl_sym[nrl] = 16; rl_extra[nrl++] = (uint8_t)(r - 3); pfreq[16]++; left -= r; }
            while (left-- > 0) { rl_sym[nrl] = (uint8_t)v; rl_extra[nrl++] = 0; pfreq[v]++; }
        }
        k += run;
    }
    uint8_t plen[19];
    uint16_t pcode[19];
    huffman_lengths(pfreq, 19, 7, plen);
    {
        int used = 0;
        for (int s = 0; s < 19; s++) used += plen[s] != 0;
        if (used == 1) { for (int s = 0; s < 19; s++) if (!plen[s]) { plen[s] = 1; break; } }   // a complete code
    }
    assign_codes(plen, 19, pcode);
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int hclen = 19;
    while (hclen > 4 && plen[order[hclen - 1]] == 0) hclen--;
    BitWriter bw{out, out + cap};
    bw.put(1, 1);            // BFINAL
    bw.put(2, 2);            // dynamic Huffman
    bw.put((uint32_t)(hlit - 257), 5);
    bw.put((uint32_t)(hdist - 1), 5);
    bw.put((uint32_t)(hclen - 4), 4);
    for (int k = 0; k < hclen; k++) bw.put(plen[order[k]], 3);
    for (int k = 0; k < nrl; k++) {
        const int s = rl_sym[k];
        bw.put(pcode[s], plen[s]);
        if (s == 16) bw.put(rl_extra[k], 2);
        else if (s == 17) bw.put(rl_extra[k], 3);
        else if (s == 18) bw.put(rl_extra[k], 7);
    }
    // ---- symbols ----
    for (uint32_t k = 0; k < ns && !bw.overflow; k++) {
        const uint32_t d = w.sym_dist[k];
        if (!d) {
            const uint32_t c = w.sym_lit[k];
            bw.put(w.lcode[c], w.llen[c]);
        } else {
            const uint32_t len = w.sym_lit[k];
            const int ls = 257 + T.len_sym[len];
            bw.put(w.lcode[ls], w.llen[ls]);
            if (T.len_xbits[len]) bw.put(len - T.len_base[len], T.len_xbits[len]);
            const int ds = T.dist_symbol(d);
            bw.put(w.dcode[ds], w.dlen[ds]);
            if (T.dist_xbits[ds]) bw.put(d - T.dist_base[ds], T.dist_xbits[ds]);
        }
    }
    bw.put(w.lcode[256], w.llen[256]);
    bw.finish();
    if (bw.overflow) return 0;
    return (size_t)(bw.p - out);
}

}  // namespace moira_deflate
