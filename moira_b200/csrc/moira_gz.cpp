// moira_gz.cpp -- gzip inputs and outputs on all host threads.
//
// The reference sniffs gzip by its magic bytes and reads the file through Python's gzip module, one thread
// (moira/moira.py:1065-1068, input; :323-370 with --output_compression gz, output).  A deflate stream cannot be split, but
// a gzip FILE may be a concatenation of members (RFC 1952 2.2), and the blocked flavour written by bgzip / htslib and by
// Illumina's FASTQ writers (BGZF: members of at most 64 KB, each carrying its own compressed size in a 'BC' extra
// subfield) can be cut without inflating anything:
//   * input:  moira_gz_inflate walks the member headers, takes every member's inflated size from its trailer, and the host
//             threads inflate the members straight into their places; any other gzip file (single member, or members
//             without the size field) is inflated by one thread, member after member, like gzip.GzipFile.read();
//   * output: moira_gz_deflate / moira_blocks_write_gz compress 65 280-byte pieces into BGZF members on all threads and
//             write them at the offsets their sizes add up to -- a file any gunzip reads, and this reader (or htslib) reads
//             in parallel.
// zlib does the deflate work (a library call for library work); CRC-32 and ISIZE of every member are checked on input.
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "moira_deflate.h"
#include "moira_inflate.h"
#include "moira_internal.h"

namespace {

constexpr uint32_t BGZF_PIECE = 0xff00;     // input bytes per member (bgzip's choice: the member stays below 64 KB even when stored)
constexpr uint32_t BGZF_HEAD = 18, BGZF_TAIL = 8, BGZF_MAX = 65536;

struct Member {
    uint64_t at;        // offset of the member in the file
    uint32_t head;      // bytes before the deflate data
    uint32_t size;      // whole member
    uint32_t isize;     // inflated bytes (trailer)
    uint64_t out;       // where they go
};

inline uint32_t le16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
inline uint32_t le32(const uint8_t *p) { return le16(p) | (le16(p + 2) << 16); }
inline void put16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
inline void put32(uint8_t *p, uint32_t v) { put16(p, v & 0xffffu); put16(p + 2, v >> 16); }

// Member list of a pure BGZF file.  false: some member has no 'BC' subfield (or the walk ran off the file): not BGZF.
bool bgzf_members(const uint8_t *gz, uint64_t n, std::vector<Member> &out, uint64_t *total)
{
    out.clear();
    uint64_t p = 0, o = 0;
    while (p < n) {
        if (n - p < BGZF_HEAD + BGZF_TAIL) {
            for (uint64_t i = p; i < n; i++) if (gz[i]) return false;   // zero padding behind the last member is tolerated (gzip does)
            break;
        }
        const uint8_t *h = gz + p;
        if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) {
            for (uint64_t i = p; i < n; i++) if (gz[i]) return false;
            break;
        }
        if (h[3] & ~4u) return false;                 // FNAME / FCOMMENT / FHCRC: variable-length fields, not a BGZF writer's header
        const uint32_t xlen = le16(h + 10);
        if (12ull + xlen + BGZF_TAIL > n - p) return false;
        uint32_t bsize = 0, x = 0;
        while (x + 4 <= xlen) {
            const uint8_t *s = h + 12 + x;
            const uint32_t slen = le16(s + 2);
            if (s[0] == 'B' && s[1] == 'C' && slen == 2 && x + 6 <= xlen) bsize = le16(s + 4) + 1;
            x += 4 + slen;
        }
        if (!bsize || bsize < 12 + xlen + BGZF_TAIL || bsize > n - p) return false;
        Member m;
        m.at = p; m.head = 12 + xlen; m.size = bsize; m.isize = le32(h + bsize - 4); m.out = o;
        if (m.isize > BGZF_MAX) return false;
        out.push_back(m);
        o += m.isize;
        p += bsize;
    }
    *total = o;
    return !out.empty() || n == 0;
}

int threads_for(int n_threads)
{
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    return T < 1 ? 1 : (T > 64 ? 64 : T);
}

template <typename F>
void run_threads(int T, F &&work)
{
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work);
    work();
    for (auto &x : th) x.join();
}

// ---- buffers handed to the caller: anonymous mappings (grown with mremap: no copy), remembered with their sizes ----------
std::mutex g_mu;
std::map<uint8_t *, size_t> g_maps;

uint8_t *map_new(size_t bytes)
{
    if (!bytes) bytes = 1;
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) return nullptr;
#ifdef MADV_HUGEPAGE
    madvise(p, bytes, MADV_HUGEPAGE);   // gigabytes of first-touch output: 2 MB pages where the kernel grants them (a hint, may fail)
#endif
    return (uint8_t *)p;
}
uint8_t *map_grow(uint8_t *p, size_t old_bytes, size_t new_bytes)
{
    void *q = mremap(p, old_bytes, new_bytes, MREMAP_MAYMOVE);
    return q == MAP_FAILED ? nullptr : (uint8_t *)q;
}

bool use_own_inflate()
{
    static const bool yes = [] { const char *e = getenv("MOIRA_B200_ZLIB_INFLATE"); return !(e && e[0] == '1'); }();
    return yes;
}

// CRC-32 of a large buffer on several threads (zlib's crc32 per slice, crc32_combine for the result)
uint32_t crc32_parallel(const uint8_t *p, uint64_t n, int T)
{
    if (n < (8u << 20) || T <= 1) {
        uLong c = crc32(0L, Z_NULL, 0);
        for (uint64_t at = 0; at < n; at += 1u << 30) c = crc32(c, p + at, (uInt)std::min<uint64_t>(1u << 30, n - at));
        return (uint32_t)c;
    }
    std::vector<uLong> part((size_t)T);
    std::vector<uint64_t> cut((size_t)T + 1);
    for (int t = 0; t <= T; t++) cut[t] = n * (uint64_t)t / T;
    std::atomic<int> next{0};
    run_threads(T, [&]() {
        for (;;) {
            const int t = next.fetch_add(1);
            if (t >= T) return;
            uLong c = crc32(0L, Z_NULL, 0);
            for (uint64_t at = cut[t]; at < cut[t + 1]; at += 1u << 30) c = crc32(c, p + at, (uInt)std::min<uint64_t>(1u << 30, cut[t + 1] - at));
            part[t] = c;
        }
    });
    uLong c = part[0];
    for (int t = 1; t < T; t++) c = crc32_combine(c, part[t], (z_off_t)(cut[t + 1] - cut[t]));
    return (uint32_t)c;
}

// length of a gzip member header (RFC 1952 2.3), 0 if it is malformed or runs off the buffer
uint64_t gzip_header_bytes(const uint8_t *p, uint64_t n)
{
    if (n < 18 || p[0] != 31 || p[1] != 139 || p[2] != 8 || (p[3] & 0xE0)) return 0;
    const int flg = p[3];
    uint64_t at = 10;
    if (flg & 4) { if (at + 2 > n) return 0; at += 2 + le16(p + at); }
    if (flg & 8) { while (at < n && p[at]) at++; at++; }
    if (flg & 16) { while (at < n && p[at]) at++; at++; }
    if (flg & 2) at += 2;
    return at + 8 <= n ? at : 0;
}

// One thread, member after member, with this library's own DEFLATE decoder (moira_inflate.h); every member's CRC-32 and size
// are checked against its trailer.  Returns 1 when the file was inflated, 0 when anything looked wrong -- the caller then
// takes the zlib route from the start, which also words the error.
int inflate_stream_own(const uint8_t *gz, uint64_t n, int n_threads, uint8_t **out, uint64_t *out_bytes)
{
    size_t cap = (size_t)(n * 4 + (1u << 20));
    uint8_t *buf = map_new(cap);
    if (!buf) return 0;
    uint64_t in_at = 0, out_at = 0;
    bool ok = true;
    static thread_local moira_inflate::State st;
    for (;;) {
        while (in_at < n && gz[in_at] == 0) in_at++;           // zero padding between / behind members
        if (in_at >= n) break;
        const uint64_t hb = gzip_header_bytes(gz + in_at, n - in_at);
        if (!hb) { ok = false; break; }
        moira_inflate::start(st, gz + in_at + hb, gz + n);
        const uint64_t member_out = out_at;
        for (;;) {
            uint8_t *pos = nullptr;
            // the window of a member is its own output (a member never refers to the one before)
            const moira_inflate::Result r = moira_inflate::run(st, buf + member_out, buf + out_at, buf + cap, &pos);
            if (r == moira_inflate::BAD_DATA) { ok = false; break; }
            out_at = (uint64_t)(pos - buf);
            if (r == moira_inflate::DONE) break;
            const size_t ncap = cap + cap / 2 + (64u << 20);
            uint8_t *nb = map_grow(buf, cap, ncap);
            if (!nb) { ok = false; break; }
            buf = nb; cap = ncap;
        }
        if (!ok) break;
        const uint8_t *tail = moira_inflate::input_position(st);
        if (tail > gz + n || (uint64_t)(gz + n - tail) < 8) { ok = false; break; }
        const uint64_t mlen = out_at - member_out;
        if (le32(tail + 4) != (uint32_t)mlen || le32(tail) != crc32_parallel(buf + member_out, mlen, threads_for(n_threads))) { ok = false; break; }
        in_at = (uint64_t)(tail - gz) + 8;
    }
    if (!ok) { munmap(buf, cap); return 0; }
    std::lock_guard<std::mutex> lk(g_mu);
    g_maps[buf] = cap;
    *out = buf;
    *out_bytes = out_at;
    return 1;
}

// one thread, member after member (what gzip.GzipFile(...).read() does)
int inflate_stream(const uint8_t *gz, uint64_t n, uint8_t **out, uint64_t *out_bytes)
{
    size_t cap = (size_t)(n * 4 + (1u << 20));
    uint8_t *buf = map_new(cap);
    if (!buf) return moira::fail(MOIRA_ERR_NOMEM, "cannot map %zu bytes for the inflated text", cap);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, 15 + 16) != Z_OK) { munmap(buf, cap); return moira::fail(MOIRA_ERR_NOMEM, "inflateInit2 failed"); }
    uint64_t in_at = 0, out_at = 0;
    int rc = MOIRA_OK;
    bool in_member = false;
    for (;;) {
        if (!in_member) {
            // between members: zero padding and nothing else may follow the last one
            while (in_at < n && gz[in_at] == 0) in_at++;
            if (in_at >= n) break;
            if (n - in_at < 2 || gz[in_at] != 31 || gz[in_at + 1] != 139) {
                rc = moira::fail(MOIRA_ERR_PARSE, in_at ? "not a gzip member at byte %llu (gzip.GzipFile: 'Not a gzipped file')" : "not a gzip file",
                                 (unsigned long long)in_at);
                break;
            }
            inflateReset(&zs);
            in_member = true;
        }
        if (out_at == cap) {
            const size_t ncap = cap + cap / 2 + (64u << 20);
            uint8_t *nb = map_grow(buf, cap, ncap);
            if (!nb) { rc = moira::fail(MOIRA_ERR_NOMEM, "cannot grow the inflated text to %zu bytes", ncap); break; }
            buf = nb; cap = ncap;
        }
        const uint64_t in_step = std::min<uint64_t>(n - in_at, 1u << 30), out_step = std::min<uint64_t>(cap - out_at, 1u << 30);
        zs.next_in = const_cast<Bytef *>(gz + in_at); zs.avail_in = (uInt)in_step;
        zs.next_out = buf + out_at; zs.avail_out = (uInt)out_step;
        const int z = inflate(&zs, Z_NO_FLUSH);
        in_at += in_step - zs.avail_in;
        out_at += out_step - zs.avail_out;
        if (z == Z_STREAM_END) { in_member = false; continue; }
        if (z == Z_OK || (z == Z_BUF_ERROR && zs.avail_out == 0)) {
            if (in_at >= n && zs.avail_out != 0) { rc = moira::fail(MOIRA_ERR_PARSE, "gzip stream ends inside a member (truncated file)"); break; }
            continue;
        }
        if (z == Z_BUF_ERROR) { rc = moira::fail(MOIRA_ERR_PARSE, "gzip stream ends inside a member (truncated file)"); break; }
        rc = moira::fail(MOIRA_ERR_PARSE, "corrupt gzip data at byte %llu: %s", (unsigned long long)in_at, zs.msg ? zs.msg : "inflate error");
        break;
    }
    inflateEnd(&zs);
    if (rc) { munmap(buf, cap); return rc; }
    std::lock_guard<std::mutex> lk(g_mu);
    g_maps[buf] = cap;
    *out = buf;
    *out_bytes = out_at;
    return MOIRA_OK;
}

// compress `n` bytes into BGZF members appended to `out`; `z0` (level 0, set up on first use) takes the pieces that do not
// shrink enough to fit a member
bool use_own_deflate()
{
    static const bool yes = [] { const char *e = getenv("MOIRA_B200_ZLIB_DEFLATE"); return !(e && e[0] == '1'); }();
    return yes;
}

// `fast`: this library's own compressor (moira_deflate.h) first; every piece it makes is inflated again (moira_inflate.h)
// and compared with the input before it is accepted -- a compressed output file is not the place for an unverified encoder.
bool deflate_range(z_stream &zs, z_stream &z0, bool &z0_ready, const uint8_t *in, uint64_t n, std::string &out, bool fast)
{
    uint8_t block[BGZF_MAX];
    constexpr uint32_t ROOM = BGZF_MAX - BGZF_HEAD - BGZF_TAIL;
    static thread_local moira_deflate::Workspace ws;
    static thread_local moira_inflate::State ist;
    static thread_local uint8_t back[BGZF_MAX];
    for (uint64_t at = 0; at < n; at += BGZF_PIECE) {
        const uint32_t len = (uint32_t)std::min<uint64_t>(BGZF_PIECE, n - at);
        uint32_t clen = 0;
        bool done = false;
        if (fast) {
            const size_t c = moira_deflate::compress_piece(in + at, len, block + BGZF_HEAD, ROOM, ws);
            if (c && c < (size_t)len + 16) {
                moira_inflate::start(ist, block + BGZF_HEAD, block + BGZF_HEAD + c);
                uint8_t *pos = nullptr;
                if (moira_inflate::run(ist, back, back, back + len, &pos) == moira_inflate::DONE && pos == back + len &&
                    memcmp(back, in + at, len) == 0 && moira_inflate::input_position(ist) == block + BGZF_HEAD + c) {
                    clen = (uint32_t)c;
                    done = true;
                }
            }
        }
        if (!done) {
            if (deflateReset(&zs) != Z_OK) return false;
            zs.next_in = const_cast<Bytef *>(in + at); zs.avail_in = len;
            zs.next_out = block + BGZF_HEAD; zs.avail_out = ROOM;
            if (deflate(&zs, Z_FINISH) == Z_STREAM_END) clen = ROOM - zs.avail_out;
            else {   // incompressible: stored
                if (!z0_ready) {
                    memset(&z0, 0, sizeof(z0));
                    if (deflateInit2(&z0, 0, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
                    z0_ready = true;
                } else if (deflateReset(&z0) != Z_OK) return false;
                z0.next_in = const_cast<Bytef *>(in + at); z0.avail_in = len;
                z0.next_out = block + BGZF_HEAD; z0.avail_out = ROOM;
                if (deflate(&z0, Z_FINISH) != Z_STREAM_END) return false;
                clen = ROOM - z0.avail_out;
            }
        }
        static const uint8_t head[BGZF_HEAD] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0, 0, 0};
        memcpy(block, head, BGZF_HEAD);
        const uint32_t total = BGZF_HEAD + clen + BGZF_TAIL;
        put16(block + 16, total - 1);
        put32(block + BGZF_HEAD + clen, (uint32_t)crc32(crc32(0L, Z_NULL, 0), in + at, len));
        put32(block + BGZF_HEAD + clen + 4, len);
        out.append((const char *)block, total);
    }
    return true;
}

}  // namespace

namespace moira {

int gz_deflate_segments_to_fd(const std::vector<std::pair<const uint8_t *, uint64_t>> &segs, int level, int n_threads, int fd,
                              uint64_t file_offset, uint64_t *written_out)
{
    *written_out = 0;
    if (level < 0 || level > 9) level = 6;
    const bool fast = level == 1 && use_own_deflate();     // level 1: the library's own compressor, zlib (level 1) behind it
    // tasks of at most 4 MB of input, never across a segment boundary (a member holds bytes of one segment only); the
    // threads share the tasks of ALL segments
    constexpr uint64_t RANGE = 64ull * BGZF_PIECE;
    std::vector<std::pair<const uint8_t *, uint64_t>> tasks;
    for (const auto &sg : segs)
        for (uint64_t a = 0; a < sg.second; a += RANGE) tasks.emplace_back(sg.first + a, std::min<uint64_t>(RANGE, sg.second - a));
    const uint64_t n_ranges = tasks.size();
    if (!n_ranges) return MOIRA_OK;
    std::vector<std::string> parts(n_ranges);
    std::atomic<uint64_t> next{0};
    std::atomic<int> bad{0};
    int T = threads_for(n_threads);
    if ((uint64_t)T > n_ranges) T = (int)n_ranges;
    run_threads(T, [&]() {
        z_stream zs, z0;
        bool z0_ready = false;
        memset(&zs, 0, sizeof(zs));
        if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { bad = 1; return; }
        for (;;) {
            const uint64_t r = next.fetch_add(1);
            if (r >= n_ranges || bad) break;
            parts[r].reserve((size_t)(tasks[r].second / 3));
            if (!deflate_range(zs, z0, z0_ready, tasks[r].first, tasks[r].second, parts[r], fast)) { bad = 1; break; }
        }
        deflateEnd(&zs);
        if (z0_ready) deflateEnd(&z0);
    });
    if (bad) return fail(MOIRA_ERR_NOMEM, "deflate failed");
    std::vector<uint64_t> at(n_ranges + 1, file_offset);
    for (uint64_t r = 0; r < n_ranges; r++) at[r + 1] = at[r] + parts[r].size();
    next = 0;
    std::atomic<int> err{0};
    run_threads(std::min(T, 8), [&]() {
        for (;;) {
            const uint64_t r = next.fetch_add(1);
            if (r >= n_ranges || err) return;
            size_t done = 0;
            while (done < parts[r].size()) {
                const ssize_t w = pwrite(fd, parts[r].data() + done, parts[r].size() - done, (off_t)(at[r] + done));
                if (w < 0) {
                    if (errno == EINTR) continue;
                    err = errno;
                    return;
                }
                done += (size_t)w;
            }
        }
    });
    if (err) return fail(MOIRA_ERR_BAD_ARG, "pwrite failed: %s", strerror(err.load()));
    *written_out = at[n_ranges] - file_offset;
    return MOIRA_OK;
}

int gz_deflate_to_fd(const uint8_t *in, uint64_t n, int level, int n_threads, int fd, uint64_t file_offset, uint64_t *written_out)
{
    std::vector<std::pair<const uint8_t *, uint64_t>> segs;
    if (n) segs.emplace_back(in, n);
    return gz_deflate_segments_to_fd(segs, level, n_threads, fd, file_offset, written_out);
}

}  // namespace moira

extern "C" {

int moira_gz_scan(const uint8_t *gz, uint64_t gz_bytes, uint64_t *n_members_out, uint64_t *inflated_bytes_out)
{
    if ((!gz && gz_bytes) || !n_members_out || !inflated_bytes_out) return moira::fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    std::vector<Member> m;
    uint64_t total = 0;
    if (gz_bytes && bgzf_members(gz, gz_bytes, m, &total)) { *n_members_out = m.size(); *inflated_bytes_out = total; }
    else { *n_members_out = 0; *inflated_bytes_out = 0; }
    return MOIRA_OK;
}

int moira_gz_inflate(const uint8_t *gz, uint64_t gz_bytes, int n_threads, uint8_t **out, uint64_t *out_bytes)
{
    if ((!gz && gz_bytes) || !out || !out_bytes) return moira::fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    *out = nullptr;
    *out_bytes = 0;
    if (gz_bytes < 2 || gz[0] != 31 || gz[1] != 139) return moira::fail(MOIRA_ERR_PARSE, "not a gzip file");
    std::vector<Member> mem;
    uint64_t total = 0;
    if (!bgzf_members(gz, gz_bytes, mem, &total)) {
        if (use_own_inflate() && inflate_stream_own(gz, gz_bytes, n_threads, out, out_bytes)) return MOIRA_OK;
        if (getenv("MOIRA_B200_GZ_DEBUG")) fprintf(stderr, "[moira_gz] own decoder declined the stream: zlib\n");
        return inflate_stream(gz, gz_bytes, out, out_bytes);      // zlib: the reference decoder, and the one that words the errors
    }

    uint8_t *buf = map_new((size_t)total);
    if (!buf) return moira::fail(MOIRA_ERR_NOMEM, "cannot map %llu bytes for the inflated text", (unsigned long long)total);
    constexpr size_t PER_TASK = 64;
    const size_t n_tasks = (mem.size() + PER_TASK - 1) / PER_TASK;
    std::atomic<size_t> next{0};
    std::atomic<long long> bad_at{-1};
    int T = threads_for(n_threads);
    if ((size_t)T > n_tasks) T = (int)std::max<size_t>(1, n_tasks);
    const bool own = use_own_inflate();
    run_threads(T, [&]() {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) { bad_at = 0; return; }
        for (;;) {
            const size_t t = next.fetch_add(1);
            if (t >= n_tasks || bad_at >= 0) break;
            for (size_t i = t * PER_TASK; i < std::min(mem.size(), (t + 1) * PER_TASK); i++) {
                const Member &m = mem[i];
                const uint32_t crc_member = le32(gz + m.at + m.size - 8);
                if (own) {   // this library's decoder first; zlib below if it disagrees with the member's trailer in any way
                    static thread_local moira_inflate::State st;
                    moira_inflate::start(st, gz + m.at + m.head, gz + m.at + m.size - BGZF_TAIL);
                    uint8_t *pos = nullptr;
                    if (moira_inflate::run(st, buf + m.out, buf + m.out, buf + m.out + m.isize, &pos) == moira_inflate::DONE &&
                        pos == buf + m.out + m.isize && (uint32_t)crc32(crc32(0L, Z_NULL, 0), buf + m.out, m.isize) == crc_member)
                        continue;
                }
                inflateReset(&zs);
                zs.next_in = const_cast<Bytef *>(gz + m.at + m.head); zs.avail_in = m.size - m.head - BGZF_TAIL;
                zs.next_out = buf + m.out; zs.avail_out = m.isize;
                const int z = inflate(&zs, Z_FINISH);
                const uint32_t crc_want = le32(gz + m.at + m.size - 8);
                if (z != Z_STREAM_END || zs.avail_out != 0 ||
                    (uint32_t)crc32(crc32(0L, Z_NULL, 0), buf + m.out, m.isize) != crc_want) { bad_at = (long long)m.at; break; }
            }
        }
        inflateEnd(&zs);
    });
    if (bad_at >= 0) {
        munmap(buf, total ? (size_t)total : 1);
        return moira::fail(MOIRA_ERR_PARSE, "corrupt gzip member at byte %lld (inflate error, size or CRC-32 mismatch)", bad_at.load());
    }
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_maps[buf] = total ? (size_t)total : 1;
    }
    *out = buf;
    *out_bytes = total;
    return MOIRA_OK;
}

int moira_gz_free(uint8_t *p)
{
    if (!p) return MOIRA_OK;
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_maps.find(p);
        if (it == g_maps.end()) return moira::fail(MOIRA_ERR_BAD_ARG, "not a buffer of moira_gz_inflate");
        bytes = it->second;
        g_maps.erase(it);
    }
    munmap(p, bytes);
    return MOIRA_OK;
}

int moira_gz_deflate(const uint8_t *data, uint64_t bytes, int level, int n_threads, int fd, uint64_t file_offset, uint64_t *written_out)
{
    if ((!data && bytes) || !written_out || fd < 0) return moira::fail(MOIRA_ERR_BAD_ARG, "bad argument");
    return moira::gz_deflate_to_fd(data, bytes, level, n_threads, fd, file_offset, written_out);
}

int moira_gz_eof(int fd, uint64_t file_offset, uint64_t *written_out)
{
    static const uint8_t eof[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (fd < 0 || !written_out) return moira::fail(MOIRA_ERR_BAD_ARG, "bad argument");
    size_t done = 0;
    while (done < sizeof(eof)) {
        const ssize_t w = pwrite(fd, eof + done, sizeof(eof) - done, (off_t)(file_offset + done));
        if (w < 0) {
            if (errno == EINTR) continue;
            return moira::fail(MOIRA_ERR_BAD_ARG, "pwrite failed: %s", strerror(errno));
        }
        done += (size_t)w;
    }
    *written_out = sizeof(eof);
    return MOIRA_OK;
}

}  // extern "C"
