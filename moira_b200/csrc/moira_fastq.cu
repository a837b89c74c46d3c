// moira_fastq.cu -- FASTQ text -> filter slab on the device: the text travels over PCIe as
// it is and is parsed where the filter runs, so the end-to-end rate from FASTQ is bound by the link, not by host
// cores.
//
// Record semantics are parse_fastq's (moira/moira.py:1152-1204, single-end): a record is four lines, every line
// is strip()ped (moira.py:1172), the sequence is line 1 and the qualities are `ord(x) - fastq_offset` of line 3
// (moira.py:1176-1177); an empty sequence, empty qualities or differing lengths are errors (moira.py:1178-1183).
// The kernels only DETECT a bad record (smallest record index); the host parser (moira_host.cpp) then re-reads
// that range to raise the reference's error with its message, so both paths fail identically.
//
//   fq_count_nl / fq_scan_blocks / fq_write_nl   positions of all '\n' of a text chunk (count, scan, write)
//   fq_records                                   per record: sequence / quality byte ranges after strip(), length,
//                                                validation, max / min length of the chunk
//   fq_convert                                   one warp per record: bases + quality characters -> slab row
//                                                (0xFF 'N', 0xFE 'n', padding 0xFD), uniform stride
#include <cuda_runtime.h>
#include <stdint.h>

#include "moira_internal.h"

namespace moira {
namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr int NL_THREADS = 256;               // 16 bytes per thread: 4 KB of text per block

// bit 7 of every byte of w that equals '\n' (exact zero-byte test of w ^ 0x0A0A0A0A)
__device__ __forceinline__ uint32_t nl_bits(uint32_t w)
{
    const uint32_t x = w ^ 0x0A0A0A0Au;
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
}
// newline bits of the 16 bytes at text[16 tid ..); only bytes in [lo, n) count (the buffer starts at a page
// boundary of the caller's text, the chunk itself at byte lo of it)
__device__ __forceinline__ void load_nl(const uint4 *text16, uint64_t tid, uint64_t lo, uint64_t n, uint32_t (&m)[4])
{
    m[0] = m[1] = m[2] = m[3] = 0;
    const uint64_t base = tid * 16;
    if (base >= n || base + 16 <= lo) return;
    const uint4 v = text16[tid];
    m[0] = nl_bits(v.x); m[1] = nl_bits(v.y); m[2] = nl_bits(v.z); m[3] = nl_bits(v.w);
    if (base + 16 > n || base < lo) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int64_t wb = (int64_t)(base + 4 * k);
            const int64_t hi_valid = (int64_t)n - wb;      // bytes of word k below n
            const int64_t lo_skip = (int64_t)lo - wb;      // bytes of word k below lo
            uint32_t keep = 0xFFFFFFFFu;
            if (hi_valid <= 0) keep = 0;
            else if (hi_valid < 4) keep &= (1u << (8 * hi_valid)) - 1u;
            if (lo_skip >= 4) keep = 0;
            else if (lo_skip > 0) keep &= ~((1u << (8 * lo_skip)) - 1u);
            m[k] &= keep;
        }
    }
}

__global__ void __launch_bounds__(NL_THREADS) fq_count_nl(const uint4 *text16, uint64_t lo, uint64_t n, uint32_t *block_cnt)
{
    uint32_t m[4];
    load_nl(text16, (uint64_t)blockIdx.x * NL_THREADS + threadIdx.x, lo, n, m);
    uint32_t c = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
    c = __reduce_add_sync(FULL, c);
    __shared__ uint32_t s[NL_THREADS / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < NL_THREADS / 32; w++) t += s[w];
        block_cnt[blockIdx.x] = t;
    }
}

// one block: exclusive scan of block_cnt[0, nb) -> block_start[0, nb], block_start[nb] = total
__global__ void __launch_bounds__(1024) fq_scan_blocks(const uint32_t *block_cnt, uint32_t nb, uint32_t *block_start)
{
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nb ? block_cnt[i] : 0;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, inc, o);
            if ((threadIdx.x & 31) >= o) inc += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_sum[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, w, o);
                if (threadIdx.x >= o) w += t;
            }
            warp_sum[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t before = carry + (threadIdx.x >= 32 ? warp_sum[(threadIdx.x >> 5) - 1] : 0);
        if (i < nb) block_start[i] = before + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_start[nb] = carry;
}

__global__ void __launch_bounds__(NL_THREADS) fq_write_nl(const uint4 *text16, uint64_t lo, uint64_t n, const uint32_t *block_start, uint32_t *nl_pos,
                                                          uint32_t cap)
{
    uint32_t m[4];
    const uint64_t tid = (uint64_t)blockIdx.x * NL_THREADS + threadIdx.x;
    load_nl(text16, tid, lo, n, m);
    const uint32_t c = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, inc, o);
        if ((threadIdx.x & 31) >= o) inc += t;
    }
    __shared__ uint32_t s[NL_THREADS / 32];
    if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint32_t before = block_start[blockIdx.x];
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) before += s[w];
    uint32_t out = before + inc - c;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t b = m[k];
        while (b) {
            const int bit = __ffs(b) - 1;       // 7, 15, 23 or 31
            b &= b - 1;
            if (out < cap) nl_pos[out] = (uint32_t)(tid * 16 + 4 * k + (bit >> 3));   // cap: the index was sized by an estimate
            out++;
        }
    }
}

__device__ __forceinline__ bool is_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }   // str.strip()'s set

// meta: [0] max length, [1] min length, [2] smallest bad record index (0xFFFFFFFF: none), and for chunks whose record
// count the host did not establish (n_rec == 0xFFFFFFFF: it is lines / 4 of the device's own newline count):
// [3] records, [4] lines % 4, [5] 1 if the newline index or the record table would overflow (nothing is parsed then)
__global__ void __launch_bounds__(256) fq_records(const uint8_t *text, uint64_t lo, uint64_t n, const uint32_t *nl_pos, const uint32_t *n_nl,
                                                  uint32_t n_rec, uint32_t extra_line, uint32_t nl_cap, uint32_t rec_cap, uint32_t *seq_off,
                                                  uint32_t *qual_off, uint32_t *len, uint32_t *meta)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t l = 0, lmin = 0xFFFFFFFFu;
    bool bad = false;
    if (n_rec == 0xFFFFFFFFu) {
        const uint32_t lines = *n_nl + extra_line;
        n_rec = lines / 4;
        const bool over = *n_nl > nl_cap || n_rec > rec_cap;
        if (r == 0) { meta[3] = n_rec; meta[4] = lines & 3u; meta[5] = over ? 1u : 0u; }
        if (over) n_rec = 0;
    }
    if (r < n_rec) {
        const uint32_t total = *n_nl;
        auto line = [&](uint32_t k, uint32_t &b, uint32_t &e) {
            b = k == 0 ? (uint32_t)lo : nl_pos[k - 1] + 1u;
            e = k < total ? nl_pos[k] : (uint32_t)n;              // an unterminated last line ends with the text
            while (b < e && is_space(text[b])) b++;               // line.strip(), moira.py:1172
            while (e > b && is_space(text[e - 1])) e--;
        };
        uint32_t sb, se, qb, qe;
        line(4 * r + 1, sb, se);
        line(4 * r + 3, qb, qe);
        l = se - sb;
        bad = l == 0 || qe == qb || (qe - qb) != l;               // EmptySeq / EmptyQual / LengthMismatch, moira.py:1178-1183
        seq_off[r] = sb;
        qual_off[r] = qb;
        len[r] = bad ? 0u : l;
        lmin = l;
    }
    const uint32_t wmax = __reduce_max_sync(FULL, l), wmin = __reduce_min_sync(FULL, lmin);
    if ((threadIdx.x & 31) == 0) {
        if (wmax) atomicMax(&meta[0], wmax);
        atomicMin(&meta[1], wmin);
    }
    if (bad) atomicMin(&meta[2], r);
}

__global__ void __launch_bounds__(256) fq_convert(const uint8_t *text, const uint32_t *seq_off, const uint32_t *qual_off,
                                                  const uint32_t *len, uint32_t n_rec, uint32_t stride, int lower_n, int qbase,
                                                  uint8_t *slab, uint32_t *meta, uint32_t *marks, uint32_t truncate,
                                                  uint8_t *seq_store, uint64_t chunk_text_base, uint64_t *seq_abs, uint32_t *seq_eff)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = warp; r < n_rec; r += warps) {
        const uint32_t l = len[r];
        const uint8_t *s = text + seq_off[r], *q = text + qual_off[r];
        uint8_t *row = slab + (uint64_t)r * stride;
        bool bad = false;
        // row marks for the filter (Ns | has-'N' << 31 over the bases that survive --truncate): the sweep then counts nothing
        const uint32_t eff = (truncate && l > truncate) ? truncate : l;
        uint32_t ns = 0, up = 0;
        for (uint32_t i0 = 0; i0 < stride; i0 += 32) {   // every lane runs every step (the ballots below)
            const uint32_t i = i0 + lane;
            uint8_t out = 0xFD;
            if (i < l) {
                const uint8_t b = s[i];
                // --collapse: the (truncated) sequence stays on the device, at its own place in the text, for moira_dedup.cu
                if (seq_store && i < eff) seq_store[chunk_text_base + seq_off[r] + i] = b;
                int v = (int)q[i] - qbase;                        // ord(x) - fastq_offset, moira.py:1177
                if (b == 'N') out = 0xFF;
                else if (b == 'n' && lower_n) out = 0xFE;
                else {
                    bad |= v > 0xFC;
                    out = (uint8_t)(v < 0 ? 0 : v);               // <= 0 is read as 1 by the table (moira.py:814)
                }
            }
            if (i < stride) row[i] = out;
            ns += __popc(__ballot_sync(FULL, out >= 0xFE && i < eff));
            up |= __ballot_sync(FULL, out == 0xFF && i < eff);
        }
        if (marks && lane == 0) marks[r] = ns | (up ? 0x80000000u : 0u);
        if (seq_store && lane == 0) { seq_abs[r] = chunk_text_base + seq_off[r]; seq_eff[r] = eff; }
        if (__any_sync(FULL, bad) && lane == 0) atomicMin(&meta[2], r);
    }
}

}  // namespace

uint32_t fq_blocks(uint64_t n) { return (uint32_t)((n + NL_THREADS * 16 - 1) / (NL_THREADS * 16)); }

// the chunk is bytes [lo, n) of the device buffer d_text (16-byte aligned)
int launch_fq_index(const uint8_t *d_text, uint64_t lo, uint64_t n, uint32_t *d_block_cnt, uint32_t *d_block_start, uint32_t *d_nl_pos,
                    uint32_t nl_cap, cudaStream_t s)
{
    const uint32_t nb = fq_blocks(n);
    if (nb == 0) return 0;
    fq_count_nl<<<nb, NL_THREADS, 0, s>>>(reinterpret_cast<const uint4 *>(d_text), lo, n, d_block_cnt);
    fq_scan_blocks<<<1, 1024, 0, s>>>(d_block_cnt, nb, d_block_start);
    fq_write_nl<<<nb, NL_THREADS, 0, s>>>(reinterpret_cast<const uint4 *>(d_text), lo, n, d_block_start, d_nl_pos, nl_cap);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// n_rec == 0xFFFFFFFF: the record count is the device's (see fq_records); the grid then covers rec_cap records
int launch_fq_records(const uint8_t *d_text, uint64_t lo, uint64_t n, const uint32_t *d_nl_pos, const uint32_t *d_n_nl, uint32_t n_rec,
                      uint32_t extra_line, uint32_t nl_cap, uint32_t rec_cap, uint32_t *d_seq_off, uint32_t *d_qual_off, uint32_t *d_len,
                      uint32_t *d_meta, cudaStream_t s)
{
    if (n_rec == 0) return 0;
    const uint32_t cover = n_rec == 0xFFFFFFFFu ? rec_cap : n_rec;
    fq_records<<<(cover + 255) / 256, 256, 0, s>>>(d_text, lo, n, d_nl_pos, d_n_nl, n_rec, extra_line, nl_cap, rec_cap, d_seq_off, d_qual_off,
                                                 d_len, d_meta);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_fq_convert(const uint8_t *d_text, const uint32_t *d_seq_off, const uint32_t *d_qual_off, const uint32_t *d_len,
                      uint32_t n_rec, uint32_t stride, int lower_n, int qbase, uint8_t *d_slab, uint32_t *d_meta, uint32_t *d_marks,
                      uint32_t truncate, int sm_count, cudaStream_t s, uint8_t *d_seq_store, uint64_t chunk_text_base, uint64_t *d_seq_abs,
                      uint32_t *d_seq_eff)
{
    if (n_rec == 0) return 0;
    const uint32_t want = (n_rec + 7) / 8;
    const uint32_t grid = want < (uint32_t)sm_count * 8 ? want : (uint32_t)sm_count * 8;
    fq_convert<<<grid, 256, 0, s>>>(d_text, d_seq_off, d_qual_off, d_len, n_rec, stride, lower_n, qbase, d_slab, d_meta, d_marks, truncate,
                                    d_seq_store, chunk_text_base, d_seq_abs, d_seq_eff);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace moira
