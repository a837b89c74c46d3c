// moira_contig.cu -- paired-end contig construction on the device: reverse complement,
// global alignment with mothur's overlap refinement, traceback and consensus, one warp per read pair.
//
// Reference semantics (reproduced exactly; checked against the CPU restatement in the tests):
//   reverse_complement  moira/moira.py:1207-1236
//   nw_align            moira/nw_align.pyx:49-145 (first row / column 0, tie order diagonal > up > left)
//   nw_overlap          moira/nw_align.pyx:148-202
//   make_contig         moira/moira.py:1375-1558
//
// Mapping.  The DP matrix has the forward read on its rows and the (reverse-complemented) reverse read on
// its columns.  Lane l of the warp owns the C columns [l C + 1, l C + C] and walks down the rows one step
// behind lane l - 1 (a skewed wavefront): at step t it fills row t - l + 1 of its strip from registers
// (the strip's previous row) plus one boundary value handed over by a shuffle.  Per step the warp stores
// ONE coalesced 128-byte line of traceback bits (word [t][lane]: a 'diagonal' and an 'up' bit per cell) to a per-warp trace buffer in
// global memory (L2-resident: it is rewritten for every pair).  The traceback then walks the path with the
// trace staged through a 32-row window in shared memory, and the consensus runs over the aligned positions
// 32 at a time with ballot-compacted output.  Integer work only, except the posterior quality tables, which
// the host builds with glibc (exact parity with the reference's pow / log10).
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include "moira_internal.h"

namespace moira {
namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;
// warps per CTA: 28 while the strip is narrow enough for 72 registers per thread, else 16
__host__ __device__ constexpr int contig_warps(int c) { return c <= 8 ? CONTIG_WARPS_PER_CTA : 16; }

// moira.py:1210-1213; 0 = not an IUPAC code (the reference raises ValueError, moira.py:1228-1229)
__device__ __forceinline__ int complement_of(int b)
{
    switch (b) {
    case 'A': return 'T'; case 'C': return 'G'; case 'T': return 'A'; case 'G': return 'C'; case 'N': return 'N';
    case 'W': return 'W'; case 'S': return 'S'; case 'R': return 'Y'; case 'Y': return 'R'; case 'M': return 'K';
    case 'K': return 'M'; case 'B': return 'V'; case 'V': return 'B'; case 'D': return 'H'; case 'H': return 'D';
    case '-': return '-'; case '.': return '.';
    default: return 0;
    }
}

struct PairView {
    const char *f;        // forward bases
    const uint8_t *fq;
    const char *r;        // reverse bases as given
    const uint8_t *rq;
    int L1, L2;
    int qbase;            // quality = byte - qbase
    bool direct;          // reverse read already reverse-complemented (single-pair entry points)
    // base / quality of column j (1-based) of the matrix, i.e. of the reverse-complemented reverse read
    __device__ __forceinline__ int rbase(int j) const { return direct ? (int)(uint8_t)r[j - 1] : complement_of((uint8_t)r[L2 - j]); }
    __device__ __forceinline__ int rqual(int j) const { return (int)(direct ? rq[j - 1] : rq[L2 - j]) - qbase; }
    __device__ __forceinline__ int fqual(int i) const { return (int)fq[i - 1] - qbase; }
};

// One warp, one pair.  Returns the per-pair status (MOIRA_PAIR_*), identical in every lane.
template <int C, bool SCORE>
__device__ int pair_to_contig(const ContigArgs &a, uint64_t pair, const PairView &v, uint8_t *smem_warp, uint32_t *trace,
                              int32_t *hbuf, int lane)
{
    const int L1 = v.L1, L2 = v.L2;
    // per-warp shared memory: forward bases | 32-row trace window | aligned row index | aligned column index
    char *s1 = reinterpret_cast<char *>(smem_warp);
    uint32_t *win = reinterpret_cast<uint32_t *>(smem_warp + a.smem_s1);
    constexpr int W = C <= 16 ? 1 : 2;   // trace words per lane and step
    uint16_t *A1 = reinterpret_cast<uint16_t *>(smem_warp + a.smem_s1 + 4096 * W);
    uint16_t *A2 = A1 + (a.max_l1 + a.max_l2);
    int n = 0;   // alignment length

    if (a.pre_a1) {
        // make_contig entry point: the alignment is given (aligned strings, moira.py:1375); lane 0 turns it into
        // the same (row, column) index lists the traceback produces, last position first
        if (lane == 0) {
            int i = 0, j = 0;
            for (int p = 0; p < a.pre_len; p++) {
                const bool g1 = a.pre_a1[p] == '-', g2 = a.pre_a2[p] == '-';
                if (!g1) i++;
                if (!g2) j++;
                A1[a.pre_len - 1 - p] = g1 ? 0 : (uint16_t)i;
                A2[a.pre_len - 1 - p] = g2 ? 0 : (uint16_t)j;
            }
        }
        n = a.pre_len;
        __syncwarp();
    } else {
        for (int k = lane; k < L1; k += 32) s1[k] = v.f[k];
        // this lane's strip of the column sequence
        int s2c[C];
        int bad_base = 0;
#pragma unroll
        for (int c = 0; c < C; c++) {
            const int j = lane * C + c + 1;
            int b = -1;
            if (j <= L2) {
                b = v.rbase(j);
                if (b == 0) bad_base = 1;
            }
            s2c[c] = b;
        }
        if (__any_sync(FULL, bad_base)) return MOIRA_PAIR_BAD_BASE;
        __syncwarp();

        const int nl = (L2 + C - 1) / C;                 // lanes that own at least one column
        const int lc = (L2 - 1) / C, c_last = (L2 - 1) % C;   // owner of the last column
        int hprev[C];
#pragma unroll
        for (int c = 0; c < C; c++) hprev[c] = 0;        // row 0 (nw_align.pyx:77-83)
        int hlast = 0, lb_prev = 0;
        int bc = 0, bci = 0;                             // best of the last column so far; row 0 holds 0
        int br = lane == 0 ? 0 : INT_MIN, bri = lane == 0 ? 0 : -1;   // best of the last row; column 0 holds 0
        const int steps = L1 + nl - 1;
        const int sub_eq = a.match, sub_ne = a.mismatch, gap = a.gap;
#pragma unroll 2
        for (int t = 0; t < steps; t++) {
            const int i = t - lane + 1;
            const int lb_in = __shfl_up_sync(FULL, hlast, 1);
            const int lb = lane == 0 ? 0 : lb_in;        // H[i][l C]: column 0 is all zeros (nw_align.pyx:70-76)
            if (i >= 1 && i <= L1 && lane < nl) {
                const int ch = (uint8_t)s1[i - 1];
                int left = lb, diagp = lb_prev;
                // One bit per cell and mask instead of a 2-bit code: D = the diagonal wins (it wins every tie,
                // nw_align.pyx:96-97), U = up beats left (ties to up, :107).  h = max(d, max(up, left) + gap).
                uint32_t mask_d = 0, mask_u = 0;
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const int d = diagp + (ch == s2c[c] ? sub_eq : sub_ne);          // nw_align.pyx:88-91
                    const int up = hprev[c];
                    const int h = __viaddmax_s32(max(up, left), gap, d);
                    if (h == d) mask_d |= 1u << c;
                    if (up >= left) mask_u |= 1u << c;
                    diagp = up;
                    hprev[c] = h;
                    left = h;
                    if (SCORE) hbuf[((size_t)t * 32 + lane) * C + c] = h;
                }
                hlast = left;
                lb_prev = lb;
                if (C <= 16) trace[(size_t)t * 32 + lane] = mask_d | (mask_u << 16);
                else {
                    trace[((size_t)t * 2) * 32 + lane] = mask_d;
                    trace[((size_t)t * 2 + 1) * 32 + lane] = mask_u;
                }
                if (lane == lc) {                        // last column: ">=" keeps the last of equal scores (nw_align.pyx:169-173)
                    int h = hprev[0];
#pragma unroll
                    for (int c = 1; c < C; c++) if (c == c_last) h = hprev[c];
                    if (h >= bc) { bc = h; bci = i; }
                }
                if (i == L1) {                           // last row, once per lane (nw_align.pyx:177-181)
#pragma unroll
                    for (int c = 0; c < C; c++)
                        if (lane * C + c + 1 <= L2 && hprev[c] >= br) { br = hprev[c]; bri = lane * C + c + 1; }
                }
            }
        }
        // nw_overlap's decision (nw_align.pyx:184-202)
        bc = __shfl_sync(FULL, bc, lc);
        bci = __shfl_sync(FULL, bci, lc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {               // max score, ties to the larger column (">=" scan order)
            const int obr = __shfl_xor_sync(FULL, br, o), obi = __shfl_xor_sync(FULL, bri, o);
            if (obr > br || (obr == br && obi > bri)) { br = obr; bri = obi; }
        }
        const int mode = (bci == L1 && bri == L2) ? 0 : (bc > br ? 1 : 2);
        __syncwarp();   // the trace stores of every lane are visible to the whole warp from here on

        // ---- traceback (nw_align.pyx:122-143); every lane walks the same path ----
        int i = L1, j = L2, wlo = INT_MAX;
        int l = (L2 - 1) / C, c = (L2 - 1) % C;          // owner lane and strip column of column j, kept incrementally
        long long score = 0;
        auto cell_score = [&](int ii, int jj, int ll, int cc) -> long long {
            return (ii > 0 && jj > 0) ? (long long)hbuf[((size_t)(ii - 1 + ll) * 32 + ll) * C + cc] : 0ll;
        };
        // nw_overlap's rewritten cells come first on the path and only there: the last column above the best row
        // (mode 1, nw_align.pyx:191-194) or the last row right of the best column (mode 2, :196-199)
        if (mode == 1) {
            while (i > bci) {
                if (SCORE) score += cell_score(i, j, l, c);
                if (lane == 0) { A1[n] = (uint16_t)i; A2[n] = 0; }
                n++; i--;
            }
        } else if (mode == 2) {
            while (j > bri) {
                if (SCORE) score += cell_score(i, j, l, c);
                if (lane == 0) { A1[n] = 0; A2[n] = (uint16_t)j; }
                n++; j--;
                if (--c < 0) { c = C - 1; l--; }
            }
        }
        while (i > 0 && j > 0) {
            const int t = i - 1 + l;
            if (t < wlo) {                               // t never increases along the path
                __syncwarp();
                wlo = t >= 31 ? t - 31 : 0;
                for (int r = 0; r < 32 * W; r++) {
                    const size_t row = (size_t)wlo * W + r;
                    win[r * 32 + lane] = row < (size_t)steps * W ? __ldcg(trace + row * 32 + lane) : 0u;
                }
                __syncwarp();
            }
            bool dbit, ubit;
            if (W == 1) {
                const uint32_t word = win[(t - wlo) * 32 + l] >> c;
                dbit = word & 1u; ubit = word & 0x10000u;
            } else {
                dbit = (win[((t - wlo) * 2) * 32 + l] >> c) & 1u; ubit = (win[((t - wlo) * 2 + 1) * 32 + l] >> c) & 1u;
            }
            const bool mv_i = dbit || ubit, mv_j = dbit || !ubit;    // diagonal: both; up: i only; left: j only
            if (SCORE) score += cell_score(i, j, l, c);
            if (lane == 0) {
                A1[n] = mv_i ? (uint16_t)i : 0;
                A2[n] = mv_j ? (uint16_t)j : 0;
            }
            n++;
            if (mv_i) i--;
            if (mv_j) { j--; if (--c < 0) { c = C - 1; l--; } }
        }
        // what is left runs along the first column (points up, nw_align.pyx:75-76) or the first row (left, :82-83)
        for (int k = lane; k < i; k += 32) { A1[n + k] = (uint16_t)(i - k); A2[n + k] = 0; }
        for (int k = lane; k < j; k += 32) { A1[n + k] = 0; A2[n + k] = (uint16_t)(j - k); }
        n += i + j;
        __syncwarp();
        if (a.al1) {   // nw_align entry point: hand the aligned strings and the path score back
            for (int p = lane; p < n; p += 32) {
                const int ia = A1[n - 1 - p], ja = A2[n - 1 - p];
                a.al1[p] = ia ? s1[ia - 1] : '-';
                a.al2[p] = ja ? (char)v.rbase(ja) : '-';
            }
            if (lane == 0) { *a.alen = n; *a.score = score; }
            return MOIRA_PAIR_OK;
        }
    }

    // ---- make_contig (moira.py:1375-1558) over aligned positions p = 0 .. n-1 (A*[n-1-p]) ----
    int fs = n, rs = n, fe = -1, re = -1;
    for (int p = lane; p < n; p += 32) {
        if (A1[n - 1 - p]) { fs = min(fs, p); fe = max(fe, p); }
        if (A2[n - 1 - p]) { rs = min(rs, p); re = max(re, p); }
    }
    fs = __reduce_min_sync(FULL, fs); rs = __reduce_min_sync(FULL, rs);
    fe = __reduce_max_sync(FULL, fe); re = __reduce_max_sync(FULL, re);
    const bool seqs_reversed = !(fs < rs);                                          // moira.py:1461-1468
    const int ostart = seqs_reversed ? fs : rs, oend = seqs_reversed ? re : fe;
    const uint64_t out0 = pair * a.out_stride;
    int n_out = 0, gaps = 0, mism = 0, bad_q = 0;
    for (int p0 = 0; p0 < n; p0 += 32) {
        const int p = p0 + lane;
        bool emit = false, is_gap = false, is_mis = false;
        int base = 'N', q = 2;
        if (p < n) {
            const int ia = A1[n - 1 - p], ja = A2[n - 1 - p];
            const int f = ia ? (int)(uint8_t)v.f[ia - 1] : '-', fqv = ia ? v.fqual(ia) : 0;
            const int r = ja ? v.rbase(ja) : '-', rqv = ja ? v.rqual(ja) : 0;
            if (fqv < 0 || fqv > 0xFC || rqv < 0 || rqv > 0xFC) bad_q = 1;   // input quality outside the table
            if (p < ostart) {                                                       // moira.py:1478-1485
                emit = !a.trim_overlap;
                base = seqs_reversed ? r : f; q = seqs_reversed ? rqv : fqv;
            } else if (p > oend) {                                                  // moira.py:1486-1493
                emit = !a.trim_overlap;
                base = seqs_reversed ? f : r; q = seqs_reversed ? fqv : rqv;
            } else if (!ia) {                                                       // moira.py:1495-1504
                is_gap = true;
                if (a.consensus == MOIRA_CONSENSUS_POSTERIOR) emit = true;
                else if (rqv > a.insert) { emit = true; base = r; q = rqv; }
            } else if (!ja) {                                                       // moira.py:1506-1515
                is_gap = true;
                if (a.consensus == MOIRA_CONSENSUS_POSTERIOR) emit = true;
                else if (fqv > a.insert) { emit = true; base = f; q = fqv; }
            } else if (f == r) {                                                    // moira.py:1517-1528
                emit = true; base = f;
                if (a.consensus == MOIRA_CONSENSUS_SUM) q = fqv + rqv;
                else if (a.consensus == MOIRA_CONSENSUS_POSTERIOR) q = a.post_match[(fqv & 255) * 256 + (rqv & 255)];
                else q = fqv >= rqv ? fqv : rqv;
            } else {                                                                // moira.py:1530-1554
                is_mis = true; emit = true;
                if (a.consensus != MOIRA_CONSENSUS_POSTERIOR) {
                    if (abs(fqv - rqv) >= a.deltaq) { base = fqv >= rqv ? f : r; q = fqv >= rqv ? fqv : rqv; }
                } else if (fqv != rqv) {
                    base = fqv > rqv ? f : r;
                    q = a.post_mis[(max(fqv, rqv) & 255) * 256 + (min(fqv, rqv) & 255)];
                }
            }
            if (a.qscore_cap && !(q < a.qscore_cap)) q = a.qscore_cap;              // moira.py:1555-1556
        }
        const uint32_t em = __ballot_sync(FULL, emit);
        gaps += __popc(__ballot_sync(FULL, is_gap));
        mism += __popc(__ballot_sync(FULL, is_mis));
        if (emit) {
            const uint64_t o = out0 + n_out + __popc(em & ((1u << lane) - 1u));
            // -1 is a value the reference's posterior formulas do return (one base of quality 0: the ratio rounds to
            // just above 1); it travels as 255 in the plain quality row and is quality 1 to the filter (moira.py:814)
            if (q < -3 || q > 0xFC) { bad_q = 1; q = 0xFC; }
            a.cseq[o] = (char)base;
            a.cqual[o] = (uint8_t)q;
            if (a.slab) a.slab[o] = base == 'N' ? 0xFF : (base == 'n' && a.lower_n ? 0xFE : (uint8_t)(q < 0 ? 0 : q));
        }
        n_out += __popc(em);
    }
    if (a.slab)   // pad the rest of the row: the filter kernels treat 0xFD as "no base"
        for (uint64_t k = n_out + lane; k < a.out_stride; k += 32) a.slab[out0 + k] = 0xFD;
    if (lane == 0) {
        a.clen[pair] = (uint32_t)n_out;
        if (a.overlap) a.overlap[pair] = oend - ostart;                             // moira.py:1470
        if (a.gaps) a.gaps[pair] = gaps;
        if (a.mism) a.mism[pair] = mism;
    }
    return __any_sync(FULL, bad_q) ? MOIRA_PAIR_BAD_QUALITY : MOIRA_PAIR_OK;
}

template <int C, bool SCORE>
__global__ void __launch_bounds__(contig_warps(C) * 32, 1) contig_kernel(const ContigArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wpc = blockDim.x >> 5;   // warps per CTA: 16, fewer when long reads need more shared memory each
    const uint32_t gw = blockIdx.x * wpc + warp, total = gridDim.x * wpc;
    uint8_t *smem_warp = smem + (size_t)warp * a.smem_per_warp;
    uint32_t *trace = a.trace + (size_t)gw * a.trace_words_per_warp;
    for (uint64_t pair = gw; pair < a.n_pairs; pair += total) {
        PairView v;
        v.L1 = (int)a.flen[pair];
        v.L2 = (int)a.rlen[pair];
        v.f = a.fseq + a.foff[pair]; v.fq = a.fqual + (a.fqoff ? a.fqoff[pair] : a.foff[pair]);
        v.r = a.rseq + a.roff[pair]; v.rq = a.rqual + (a.rqoff ? a.rqoff[pair] : a.roff[pair]);
        v.qbase = a.qual_base;
        v.direct = a.rev_direct != 0;
        int status;
        if (!a.pre_a1 && (v.L1 <= 0 || v.L2 <= 0)) status = MOIRA_PAIR_EMPTY;
        else if (v.L1 > (int)a.max_l1 || v.L2 > (int)a.max_l2 || v.L2 > 32 * C) status = MOIRA_PAIR_TOO_LONG;
        else status = pair_to_contig<C, SCORE>(a, pair, v, smem_warp, trace, a.hbuf, lane);
        if (status != MOIRA_PAIR_OK && status != MOIRA_PAIR_BAD_QUALITY && !a.al1) {
            // no contig: an empty row, so that whatever runs next sees a read of length 0
            if (a.slab) for (uint64_t k = lane; k < a.out_stride; k += 32) a.slab[pair * a.out_stride + k] = 0xFD;
            if (lane == 0) {
                a.clen[pair] = 0;
                if (a.overlap) a.overlap[pair] = 0;
                if (a.gaps) a.gaps[pair] = 0;
                if (a.mism) a.mism[pair] = 0;
            }
        }
        if (lane == 0 && a.status) a.status[pair] = (uint8_t)status;
        __syncwarp();
    }
}

template <int C, bool SCORE>
int launch_c(const ContigArgs &a, int grid, int warps, size_t smem, cudaStream_t s)
{
    // per device and cheap: set on every launch rather than remembered (one process may hold contexts on several GPUs)
    if (cudaFuncSetAttribute(contig_kernel<C, SCORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return -1;
    contig_kernel<C, SCORE><<<grid, warps * 32, smem, s>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace

int contig_columns_per_lane(uint32_t max_l2)
{
    for (int c : {4, 8, 10, 16, 32}) if (max_l2 <= 32u * c) return c;
    return 0;
}

size_t contig_trace_words_per_warp(uint32_t max_l1, uint32_t max_l2)
{
    const int c = contig_columns_per_lane(max_l2);
    return (size_t)(max_l1 + 32) * 32 * (c <= 16 ? 1 : 2);
}

size_t contig_hbuf_words(uint32_t max_l1, uint32_t max_l2)
{
    return (size_t)(max_l1 + 32) * 32 * contig_columns_per_lane(max_l2);
}

int contig_grid(int sm_count, uint64_t n_pairs, int warps)
{
    const uint64_t ctas = (n_pairs + warps - 1) / warps;
    return (int)(ctas < (uint64_t)sm_count ? (ctas ? ctas : 1) : (uint64_t)sm_count);
}

int launch_contigs(ContigArgs a, bool want_score, const LaunchCfg &cfg)
{
    const int c = contig_columns_per_lane(a.max_l2);
    if (!c) return -2;
    a.smem_s1 = (a.max_l1 + 15u) & ~15u;
    a.smem_per_warp = (a.smem_s1 + 4096u * (c <= 16 ? 1 : 2) + 4u * (a.max_l1 + a.max_l2) + 15u) & ~15u;
    int warps = contig_warps(c);
    while (warps > 1 && (size_t)a.smem_per_warp * warps > 227 * 1024) warps--;
    const size_t smem = (size_t)a.smem_per_warp * warps;
    if (smem > 227 * 1024) return -2;
    const int grid = contig_grid(cfg.sm_count, a.n_pairs, warps);
#define MOIRA_CONTIG_CASE(CC)                                                                   \
    case CC: return want_score ? launch_c<CC, true>(a, grid, warps, smem, cfg.stream) : launch_c<CC, false>(a, grid, warps, smem, cfg.stream);
    switch (c) {
        MOIRA_CONTIG_CASE(4)
        MOIRA_CONTIG_CASE(8)
        MOIRA_CONTIG_CASE(10)
        MOIRA_CONTIG_CASE(16)
        MOIRA_CONTIG_CASE(32)
    }
#undef MOIRA_CONTIG_CASE
    return -2;
}

}  // namespace moira
