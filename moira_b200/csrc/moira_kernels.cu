// moira_kernels.cu -- sm_100a kernels of the moira read filter.
//
// What is computed (reference: /root/reference/moira/bernoullimodule.c:182-263):
// per read, the Poisson-binomial PMF of the number of sequencing errors, swept base by base
// while keeping only the first K entries P[0..K-1]:
//      P[j] <- fl( fl(q * P[j]) + fl(e * P[j-1]) )   (j = K-1 .. 1),      P[0] <- fl(q * P[0])
// with q = 1-p and e = ((1/1.0) * (p/(1-p))) * (1-p) taken from a host-libm table.  Every cell is
// bitwise the cell the reference computes (bernoullimodule.c:160 with n == 1; the i >= 2 terms
// are +0.0), so no FMA contraction is allowed: all PMF arithmetic uses __dmul_rn / __dadd_rn.
// Then acc[j] = acc[j-1] + P[j] until acc[j] > 1-alpha (:233-244) and the linear interpolation
// of bernoullimodule.c:170-178.
//
// Kernels:
//   tpr_kernel<K,0>  "pb_tpr": thread-per-read, P[0..K-1] in registers, persistent grid.  Rows are staged
//                 global->shared in 128-byte chunks, double-buffered per warp: by TMA tensor tiles
//                 (one UTMALDG per warp and chunk, 128-byte swizzle) for uniform-stride slabs, by
//                 cooperative cp.async (LDGSTS) otherwise.  The Q->(q,e) table lives in shared memory at a
//                 64 KB-aligned address, 16 replicas per row, so that one PRMT builds the lookup address
//                 and a warp's 32 random lookups never conflict.
//   tpr_kernel<1,1>  "lambda_tpr": same staging, accumulates Lambda = sum p_i sequentially (Poisson and
//                 expected-error modes, moira.py:1637-1679).
//   tpr_kernel<2,2>  ladder classifier: mean/variance of the error count -> rung.
//   wpr_kernel<M>    warp-per-read for reads that need many PMF entries (K = 32*M): lane l owns
//                 P[l*M .. l*M+M-1], the neighbour entry travels by warp shuffle.
//   blk_kernel       block-per-read, P in shared memory, any K up to 24576: last rung.
//   len_*_kernel     counting sort of read indices by length (ragged batches).
//   fp64_peak_kernel register-resident DMUL/DADD issue-rate probe (roofline denominator).
#include <cuda.h>
#include <math.h>

#include <algorithm>
#include <type_traits>

#include "fact_table.h"
#include "moira_internal.h"

namespace moira {
namespace {

constexpr unsigned FULL = 0xffffffffu;

// ---- thread-per-read geometry ---------------------------------------------------------------
#ifndef MOIRA_CHUNK
#define MOIRA_CHUNK 128
#endif
#ifndef MOIRA_WARPS_SMALLK
#define MOIRA_WARPS_SMALLK 16
#endif
#ifndef MOIRA_WARPS16_MAXK
#define MOIRA_WARPS16_MAXK 12   // largest K that still runs 16 warps per CTA (<= 128 registers per thread)
#endif
#ifndef MOIRA_PAIR_LUT
#define MOIRA_PAIR_LUT 1   // 1: first-pass kernels look up (q, e) pairs (no DSUB, LDS.128); 0: p only (LDS.64 + DSUB)
#endif
#ifndef MOIRA_PL_MAXK
#define MOIRA_PL_MAXK 2    // ... but K <= this looks up p only: with so little arithmetic per base the 16-byte lookups bind
#endif
constexpr int CHUNK = MOIRA_CHUNK;               // bytes of one row staged per pipeline stage (64 or 128)
// row stride = CHUNK + 16 B (144 B = 36 words / 80 B = 20 words): the 8 lanes of an LDS.128 phase
// start at word offsets lane*36 (or lane*20) mod 32 = distinct multiples of 4 -> conflict-free
constexpr int ROW_STRIDE = CHUNK + 16;
constexpr int STAGE_BYTES = 32 * ROW_STRIDE;     // one warp, one stage
// Q -> probability table: 256 rows of 256 bytes at a 64 KB-aligned shared address, so that one PRMT
// builds the lane's lookup address  table | Q << 8 | replica  (byte 1 <- quality byte) and the
// replicas make a warp's 32 random lookups conflict-free: 32 x double (p) or 16 x double2 (q, e).
constexpr int LUT_BYTES = 256 * 256;
constexpr int TPR_SMEM = 227 * 1024;             // whole opt-in shared memory; laid out at run time
// P[0..K-1] lives in registers: 16 warps per CTA up to K = 12, 8 above (<= 255 registers; measured best for the ladder).
__host__ __device__ constexpr int tpr_warps(int k) { return k <= 4 ? MOIRA_WARPS_SMALLK : k <= MOIRA_WARPS16_MAXK ? 16 : 8; }

constexpr int WPR_THREADS = 256;
constexpr int BLK_THREADS = 256;
constexpr int BLK_CAP = 24576;                   // PMF entries held in shared memory by pb_blk
constexpr int BLK_SMEM = BLK_CAP * 8 + 256 * 16;

__constant__ double c_fact[MOIRA_FACT_N] = MOIRA_FACT_TABLE;
__constant__ double c_ratio[9] = {1.0, 1.0, 0.5, 0.66666666666666674, 0.75, 0.80000000000000004, 0.83333333333333337, 0.85714285714285721, 0.875};   // (j - 1) / j, rounded up

// ---- PTX helpers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// Ampere-style asynchronous 16-byte copy global -> shared (SASS: LDGSTS); src_size 0 zero-fills.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_size)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_size) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// TMA tensor-tile copy global -> shared (SASS: UTMALDG): one 2-D box {x .. x+127 bytes, y .. y+31 rows}
// of the slab, written with the 128-byte swizzle; completion counted in bytes on an mbarrier.
__device__ __forceinline__ void tma_tile_g2s(uint32_t dst, const CUtensorMap *tmap, int x, int y, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// ---- byte handling -------------------------------------------------------------------------------
// Replace bytes at positions >= nvalid (0..4) of a little-endian word by the padding code 0xFD.
__device__ __forceinline__ uint32_t mask_word(uint32_t w, int nvalid)
{
    if (nvalid >= 4) return w;
    if (nvalid <= 0) return 0xFDFDFDFDu;
    uint32_t keep = (1u << (8 * nvalid)) - 1u;
    return (w & keep) | (0xFDFDFDFDu & ~keep);
}
// Bit 7 of every byte that is >= 0xFE ('n' / 'N' marker): 3 integer ops per word, no branch.
__device__ __forceinline__ uint32_t mark_bits(uint32_t w)
{
    return ((w & 0x7F7F7F7Fu) + 0x02020202u) & w & 0x80808080u;
}
// N/n accounting for one 16-byte vector.  Real qualities never have bit 7 set, so one OR/AND over
// the four words filters out every vector without markers (and without tail padding); the exact
// count runs only for the others.  ns128 accumulates 128 per marker byte.
__device__ __forceinline__ void count_marks4(const uint32_t (&w)[4], uint32_t &ns, uint32_t &upper_n)
{
#ifdef MOIRA_EXPERIMENT_NO_MARKS   // tuning experiments only: what the N/n accounting costs (results are wrong for reads with N)
    return;
#endif
    if ((w[0] | w[1] | w[2] | w[3]) & 0x80808080u) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t m = mark_bits(w[i]);
            ns += __popc(m);
            upper_n |= m & (w[i] << 7);     // marker byte with bit 0 set == 0xFF ('N')
        }
    }
}

// ---- read geometry -------------------------------------------------------------------------------
struct ReadGeom {
    uint64_t off;
    uint32_t len;      // original length
    uint32_t eff;      // after --truncate (moira.py:806-807)
};
__device__ __forceinline__ ReadGeom read_geom(const FilterArgs &a, uint32_t r_local)
{
    ReadGeom g;
    uint64_t r = a.base + r_local;
    g.off = a.offsets ? a.offsets[r] : r * a.stride;
    g.len = a.lengths ? a.lengths[r] : a.fixed_length;
    g.eff = (a.truncate && g.len > a.truncate) ? a.truncate : g.len;
    return g;
}

// ---- per-read epilogue: process_data's +Ns / floor (moira.py:827-831), write_results' decision
//      (moira.py:872-970), counters, and escalation of reads this pass could not settle -----------
struct ReadResult {
    double ee_raw;     // exact statistic, or a lower bound when !resolved
    uint32_t processed;  // bases swept before a (warp-wide) early exit
    int ns;
    bool has_n;
    bool resolved;
    bool numeric;
    bool escalate;     // cascade: not settled by this launch's entries and not a certain reject either -> next sweep, whatever
                       // the other rules say (so that the cascade writes exactly what the single sweep writes)
    int next_rung;     // where an unsettled read goes: the next rung, or the one the first pass picked itself (direct_rung)
};

// Smallest rung whose capacity holds k entries, in closed form: a search through rung_cap()'s table indexes a constexpr array
// with a run-time value, which the compiler keeps in LOCAL memory -- ncu showed the classifier waiting on those LDL/STL
// (20 % of its warp-state samples).
__host__ __device__ constexpr int rung_for(int k)
{
    if (k <= 8) return k <= 3 ? 1 : k - 2;
    if (k <= 24) return 6 + ((k - 8 + 1) >> 1);
    if (k <= 28) return 15;
    if (k <= 32) return 16;
    if (k <= 40) return 17;
    if (k <= 48) return 18;
    if (k <= 64) return 19;
    if (k <= 128) return 20;
    if (k <= 256) return 21;
    if (k <= 512) return 22;
    if (k <= 1024) return 23;
    return NB - 1;
}
constexpr bool rung_for_matches_table()
{
    for (int k = 1; k <= 1100; k++) {
        const int b = rung_for(k);
        if (b < 1 || b > NB - 1 || rung_cap(b) < k || (b > 1 && rung_cap(b - 1) >= k)) return false;
    }
    return true;
}
static_assert(rung_for_matches_table(), "rung_for() and rung_cap() out of step");
__device__ __forceinline__ int pick_rung(const FilterArgs &a, int kneed)
{
    const int lo = a.min_rung > 1 ? a.min_rung : 1, b = rung_for(kneed);
    return b > lo ? b : lo;
}

// The rung of a read the first pass swept completely without settling it, from the two entries every sweep tracks
// (no classifier pass over the row): with f_i = -ln(1 - p_i) and g_i = p_i / (1 - p_i),
//     A = -ln P[0] = sum f_i        B = P[1] / P[0] = sum g_i        (P[1] = prod(1 - p_i) * sum g_i)
// and  1.34 f - 0.34 g >= p  for every p <= 10^-0.1 (Q >= 1; equality to second order at p -> 0, +0.012 at Q = 1), so
//     mu_hat = 1.34 A - 0.34 B >= mu = sum p_i >= sigma^2 ,
// which goes into the classifier's own quantile estimate (Cornish-Fisher with the Poisson skew bound).  An over-estimate
// costs entries (about one on MiSeq-like noisy reads), an under-estimate one more rung; results never depend on it.
// P[0] too small for the quotient: rung 0, the classifier.
__device__ __forceinline__ int rung_from_two_entries(const FilterArgs &a, double p0, double p1, uint32_t eff, uint32_t ns, int k_tried)
{
    if (!(p0 > 1e-280)) return 0;
    const double A = -log(p0), B = p1 / p0;
    // B - A = sum(p^2 / 2 + 2 p^3 / 3 + ...) measures how loose the estimate is (sigma^2 = mu - sum p^2 is taken as mu): reads
    // with many low-quality bases are worth the classifier's pass over the row
    if (!(B - A <= a.direct_gap)) return 0;
    double mu = 1.34 * A - 0.34 * B;
    if (!(mu > 0.0)) mu = 0.0;
    double kn = mu + a.z * sqrt(mu) + a.zc;
    if (!a.exact) {   // a decision needs at most floor(cutoff on the raw statistic) + 2 entries
        double c = (a.thr_kind == MOIRA_THR_MAXERRORS) ? a.thr : __dmul_rn((double)eff, a.thr);
        if (a.ambigs == MOIRA_AMBIGS_TREAT_AS_ERRORS) c -= (double)ns;
        const double kd = floor(c) + 2.0;
        if (kd < kn) kn = kd;
    }
    int kneed = kn > 1.0e6 ? 1000000 : (kn < 2.0 ? 2 : (int)kn);
    if (kneed <= k_tried) kneed = k_tried + 1;   // k_tried entries did not settle it
    return pick_rung(a, kneed);
}

// Append read r_local to the queue of `rung` (warp-aggregated).  Called convergently; lanes with
// push == false only take part in the ballot.
__device__ __forceinline__ void push_read(const FilterArgs &a, bool push, int rung, uint32_t r_local, int lane)
{
    const unsigned pm = __ballot_sync(FULL, push);
    if (push) {
        const unsigned peers = __match_any_sync(pm, rung);
        const int leader = __ffs(peers) - 1;
        uint32_t basepos = 0;
        if (lane == leader) basepos = atomicAdd(&a.queue_counts[rung], (uint32_t)__popc(peers));
        basepos = __shfl_sync(peers, basepos, leader);
        const uint32_t pos = basepos + __popc(peers & ((1u << lane) - 1u));
        if (pos < a.queue_cap) a.queues[(size_t)rung * a.queue_cap + pos] = r_local;
    }
}

// Called by all 32 lanes of a warp (convergent).  `valid` lanes carry a read.
__device__ __forceinline__ void finish_read(const FilterArgs &a, bool valid, uint32_t r_local, const ReadGeom &g,
                                            const ReadResult &res, uint32_t *s_cnt, uint32_t *s_hist, int lane)
{
    const double nsd = (double)res.ns;
    double ee_fin = res.ee_raw;
    if (a.ambigs == MOIRA_AMBIGS_TREAT_AS_ERRORS) ee_fin = __dadd_rn(ee_fin, nsd);   // moira.py:828
    if (a.round_flag) ee_fin = floor(ee_fin);                                        // moira.py:831
    const double cutoff = (a.thr_kind == MOIRA_THR_MAXERRORS) ? a.thr : __dmul_rn((double)g.eff, a.thr);  // :926 / :950
    bool ok = ee_fin <= cutoff;
    int reason = MOIRA_REASON_NONE;
    if (a.truncate && g.len < a.truncate) reason = MOIRA_REASON_LENGTH;               // :872
    else if (res.has_n && a.ambigs == MOIRA_AMBIGS_DISALLOW) reason = MOIRA_REASON_AMBIGS;  // :911

    bool undecided = false;
    if (!res.resolved && !res.numeric) {
        // ee_raw is a lower bound.  The read is settled only if a decision is all that is asked
        // for and the bound already exceeds the cutoff (or another rule rejects it anyway).
        undecided = a.exact || res.escalate || (reason == MOIRA_REASON_NONE && ok);
        ok = false;
    }
    if (reason != MOIRA_REASON_NONE) ok = false;
    else if (!ok) reason = MOIRA_REASON_ERRORS;

    bool numeric = res.numeric;
    const bool push = valid && undecided && a.allow_push && a.rung < NB - 1;
    if (valid && undecided && !push) numeric = true;   // nowhere left to go (see MOIRA_ERR_UNRESOLVED)
    if (numeric) { ok = false; if (reason == MOIRA_REASON_NONE) reason = MOIRA_REASON_ERRORS; }

    // ---- escalate: first pass -> classifier (rung 0); a rung -> the next one -------------------
    if (a.rung < 0) {   // first pass: how many reads it hands on (diagnostic counter)
        const unsigned pm = __ballot_sync(FULL, push);
        if (pm && lane == 0) atomicAdd(&s_cnt[MOIRA_CNT_ESCALATED], (uint32_t)__popc(pm));
    }
    push_read(a, push, res.next_rung, r_local, lane);
    if (!valid || push) return;

    // ---- write + count ----------------------------------------------------------------------
    const bool lower = !res.resolved && !numeric;
    const bool near_cut = res.resolved && !numeric && fabs(ee_fin - cutoff) <= 1e-12 * cutoff;
    uint8_t fl = (ok ? MOIRA_FLAG_ACCEPT : 0) | (uint8_t)(reason << 1) | (lower ? MOIRA_FLAG_LOWER_BOUND : 0) |
                 (res.has_n ? MOIRA_FLAG_HAS_N : 0) | (numeric ? MOIRA_FLAG_NUMERIC : 0) |
                 (near_cut ? MOIRA_FLAG_NEAR_CUTOFF : 0);
    const uint64_t r = a.base + r_local;
    a.ee[r] = (a.ee_output == MOIRA_EE_FINAL) ? ee_fin : res.ee_raw;
    if (a.ns) a.ns[r] = res.ns;
    if (a.flags) a.flags[r] = fl;
    const int bin = ee_fin >= 63.0 ? 63 : (ee_fin > 0.0 ? (int)ee_fin : 0);
    // (measured dead end: warp-aggregating these shared atomics by MATCH.ANY / ballots -- one lane per distinct counter adds
    // the group's size -- ran 2 % slower on the two-entry sweep: 1.046 vs 1.026 ms per 10 M reads)
    atomicAdd(&s_cnt[MOIRA_CNT_READS], 1u);
    atomicAdd(&s_cnt[ok ? MOIRA_CNT_ACCEPTED : (MOIRA_CNT_BAD_ERRORS + reason - 1)], 1u);
    if (near_cut) atomicAdd(&s_cnt[MOIRA_CNT_NEAR_CUTOFF], 1u);
    if (lower) atomicAdd(&s_cnt[MOIRA_CNT_LOWER_BOUND], 1u);
    if (numeric) atomicAdd(&s_cnt[MOIRA_CNT_NUMERIC], 1u);
    atomicAdd(&s_hist[bin], 1u);
}

__device__ __forceinline__ void flush_counters(const FilterArgs &a, const uint32_t *s_cnt, const uint32_t *s_hist)
{
    if (!a.counters) return;
    for (int i = threadIdx.x; i < 16 + MOIRA_N_HIST; i += blockDim.x) {
        uint32_t v = i < 16 ? s_cnt[i] : s_hist[i - 16];
        if (v) atomicAdd(&a.counters[i], (unsigned long long)v);
    }
    // pilot launch of the cascade: its own histogram of floor(ee) (bins >= 15 together) for the policy kernel
    if (a.jhist)
        for (int i = threadIdx.x; i < MOIRA_N_HIST; i += blockDim.x)
            if (s_hist[i]) atomicAdd(&a.jhist[i < 15 ? i : 15], s_hist[i]);
}

// Upper bound of acc[kd - 1] = P(X <= kd - 1) from the K < kd entries a sweep tracks.  The PMF of a Poisson-binomial count
// is e_j(w) * prod(1 - p_i) with w_i = p_i / (1 - p_i), and Newton's inequalities for the elementary symmetric
// polynomials give rho_{j+1} <= (j / (j + 1)) * rho_j for the ratios rho_j = P[j] / P[j-1].  So beyond the tracked entries
//     P[K] <= P[K-1] * rho_{K-1} * (K-1)/K ,   P[K+1] <= P[K] * rho_{K-1} * (K-1)/(K+1) , ...
// (K = 2: P[j] <= P[0] * r^j / j! with r = P[1] / P[0]).  The prefix version holds too (acc[kd - 1] never increases as bases
// are added), so the bound may be taken after any number of bases.  Returns 1 when the last-but-one entry is too small for
// the quotient to be trusted.
template <int K>
__device__ __forceinline__ double newton_bound(const double (&P)[K], int kd)
{
    static_assert(K >= 2, "needs two tracked entries");
    if (!(P[K - 2] > 1e-280)) return 1.0;
    double rho = P[K - 1] / P[K - 2];
    double term = P[K - 1], sum = P[0];
#pragma unroll
    for (int j = 1; j < K; j++) sum += P[j];
#pragma unroll 1
    for (int j = K; j < kd; j++) {   // kd is launch-uniform: a real loop, not eight predicated copies
        rho = rho * c_ratio[j];      // (j - 1) / j rounded up; the 1e-9 margin of the callers covers every rounding here many times over
        term = term * rho;
        sum += term;
    }
    return sum == sum ? sum : 1.0;
}

template <int K>
__device__ __forceinline__ double cascade_bound(const double (&P)[K], int kd)
{
    if constexpr (K >= 2) return newton_bound<K>(P, kd);
    else return 1.0;
}

// bernoullimodule.c:233-254: cumulative sum in index order, strict '>' against 1-alpha, then
// interpolate().  j* == 0 gives 0 (the interpolation is negative there and clamps, :173-176).
template <int K>
__device__ __forceinline__ bool cdf_quantile(const double (&P)[K], double oma, double &ee)
{
    double acc = P[0];
    if (acc > oma) { ee = 0.0; return true; }
#pragma unroll
    for (int j = 1; j < K; j++) {
        double prev = acc;
        acc = __dadd_rn(prev, P[j]);
        if (acc > oma) {
            ee = __dadd_rn((double)(j - 1), __ddiv_rn(__dsub_rn(oma, prev), __dsub_rn(acc, prev)));
            return true;
        }
    }
    ee = (double)(K - 1);   // j* >= K  =>  ee >= K-1
    return false;
}

// Load one 16-byte vector of a staged row, mask what lies beyond the read to padding, account N/n.
template <bool MARKS>
__device__ __forceinline__ void load_vec(uint32_t addr, int rem, uint32_t (&w)[4], uint32_t &ns, uint32_t &has_n)
{
    const uint4 q4 = lds128(addr);
    w[0] = q4.x; w[1] = q4.y; w[2] = q4.z; w[3] = q4.w;
    if (rem < 16) {
#pragma unroll
        for (int i = 0; i < 4; i++) w[i] = mask_word(w[i], rem - 4 * i);
    }
    if (MARKS) count_marks4(w, ns, has_n);
}

// Lookup address of byte b of w: table base (64 KB aligned) | Q << 8 | lane replica offset, in one PRMT.
template <int B>
__device__ __forceinline__ uint32_t lut_addr_of(uint32_t w, uint32_t lut_lane)
{
    return __byte_perm(w, lut_lane, 0x7604u | (B << 4));
}
// The per-base work on the four bases of one quality word.
template <int K, int MODE, bool PL, typename T>
__device__ __forceinline__ void sweep_word(uint32_t wi, uint32_t lut_lane, T (&P)[K])
{
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const uint32_t addr = __byte_perm(wi, lut_lane, 0x7604u | (b << 4));   // table | Q << 8 | replica
        if (MODE == 0) {
            double q, e;
            if (PL) {
                e = lds_f64(addr);
                q = __dsub_rn(1.0, e);                 // (1 - p), bernoullimodule.c:140
            } else {
                const double2 qe = lds_f64x2(addr);
                q = qe.x; e = qe.y;
            }
#pragma unroll
            for (int j = K - 1; j >= 1; j--)
                P[j] = __dadd_rn(__dmul_rn(q, P[j]), __dmul_rn(e, P[j - 1]));
            P[0] = __dmul_rn(q, P[0]);
        } else if (MODE == 1) {
            P[0] = __dadd_rn(P[0], lds_f64(addr));     // moira.py:1663, in index order
        } else {
            const float pv = lds_f32(addr);            // classifier: sum p and sum p^2 (fp32: an estimate) -> mean, variance = sum p - sum p^2
            P[0] += pv;
            P[1] = __fmaf_rn(pv, pv, P[1]);
        }
    }
}
// The per-base work on one 16-byte vector already in registers.  ROLL = 1 / 2 keeps the loop over the four words rolled
// (one / two words per iteration): the ladder kernel -- every K in one kernel, 14 .. 50 KB of unrolled 16-base body per K --
// streamed its code through the instruction caches (ncu: 9 % of its warp states were no_instruction).  The single-K first
// pass kernels keep the unrolled body (rolled: -4 % on the K = 18 sweep of 1 500-bp reads).
#ifndef MOIRA_LADDER_SPLIT
#define MOIRA_LADDER_SPLIT 8   // last rung of the 16-warp ladder launch
#endif
#ifndef MOIRA_LADDER_ROLL
#define MOIRA_LADDER_ROLL 1
#endif
#ifndef MOIRA_ROLL_MINK
#define MOIRA_ROLL_MINK 12
#endif
template <int K, int MODE, bool PL, int ROLL, typename T>
__device__ __forceinline__ void sweep_vec(const uint32_t (&w)[4], uint32_t lut_lane, T (&P)[K])
{
    if constexpr (MODE == 0 && ROLL == 1 && K >= MOIRA_ROLL_MINK) {
#pragma unroll 1
        for (int i = 0; i < 4; i++) {
            const uint32_t wi = i == 0 ? w[0] : (i == 1 ? w[1] : (i == 2 ? w[2] : w[3]));
            sweep_word<K, MODE, PL, T>(wi, lut_lane, P);
        }
    } else if constexpr (MODE == 0 && ROLL == 2 && K >= MOIRA_ROLL_MINK) {
#pragma unroll 1
        for (int i = 0; i < 2; i++) {
            sweep_word<K, MODE, PL, T>(i == 0 ? w[0] : w[2], lut_lane, P);
            sweep_word<K, MODE, PL, T>(i == 0 ? w[1] : w[3], lut_lane, P);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) sweep_word<K, MODE, PL, T>(w[i], lut_lane, P);
    }
}

// One staged chunk of one row (thread-per-read): `cend` bytes starting at shared address `row`, of
// which the first `rem0` belong to this lane's read (the rest is masked to padding).  The first
// `cfull` bytes (warp-uniform, multiple of 16) are inside the read for EVERY lane: no masking there.
// MATH = false keeps only the N/n accounting (after a warp-wide early exit).
// `swz` is the lane's XOR term of the TMA 128-byte swizzle ((lane & 7) << 4), 0 for the padded layout.
// MARKS = false: Ns / has-N of every row came with the slab (FilterArgs::row_marks), nothing is counted here.
template <int K, int MODE, bool PL, bool MATH, bool MARKS, int ROLL, typename T>
__device__ __forceinline__ void sweep_chunk(uint32_t row, uint32_t swz, uint32_t cfull, uint32_t cend, int rem0,
                                            uint32_t lut_lane, T (&P)[K], uint32_t &ns, uint32_t &has_n)
{
    if (!MATH && !MARKS) return;
    uint32_t v = 0;
    for (; v < cfull; v += 16) {
        const uint4 q4 = lds128(row + (v ^ swz));
        const uint32_t w[4] = {q4.x, q4.y, q4.z, q4.w};
        if (MARKS) count_marks4(w, ns, has_n);
        if (MATH) sweep_vec<K, MODE, PL, ROLL, T>(w, lut_lane, P);
    }
    for (; v < cend; v += 16) {
        uint32_t w[4];
        load_vec<MARKS>(row + (v ^ swz), rem0 - (int)v, w, ns, has_n);
        if (MATH) sweep_vec<K, MODE, PL, ROLL, T>(w, lut_lane, P);
    }
}

// ==================================================================================================
// thread-per-read kernels
// ==================================================================================================
// Shared-memory context of a thread-per-read CTA.
struct TprCtx {
    uint32_t lut_lane;     // shared address of this lane's replica column of the lookup table
    uint32_t stage_warp;   // shared address of this warp's two stage buffers
    uint32_t bar0;         // shared address of this warp's two mbarriers (TMA staging)
    uint32_t *s_cnt, *s_hist;
    int lane, warp;
};

// One-time CTA setup: run-time layout of the dynamic shared memory, replicated lookup table, counters,
// mbarriers.  PL: the table holds p only (32 x 8 B per row) instead of (q, e) pairs (16 x 16 B).
template <int TPR_WARPS, int MODE, bool PL, bool TMA>
__device__ __forceinline__ TprCtx tpr_setup(const FilterArgs &a, uint8_t *smem)
{
    constexpr int TPR_THREADS = TPR_WARPS * 32;
    constexpr uint32_t STG = TMA ? 32u * CHUNK : (uint32_t)STAGE_BYTES;   // bytes of one warp-stage
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // ---- run-time layout of the dynamic shared memory (TPR_SMEM bytes) ----------------------
    //   [base, lut)            stage buffers of the first nA warps (whatever fits below the table)
    //   [lut, lut + 64 KB)     lookup table, 64 KB aligned (see LUT_BYTES)
    //   [lut + 64 KB, ...)     stage buffers of the remaining warps, then counters + histogram
    const uint32_t base = smem_u32(smem);
    const uint32_t lut = (base + 0xFFFFu) & ~0xFFFFu;
    const uint32_t below0 = (base + 1023u) & ~1023u;           // stage buffers are 1 KB aligned (TMA swizzle atom)
    const uint32_t n_below = min((lut - below0) / (2u * STG), (uint32_t)TPR_WARPS);
    const uint32_t above = lut + LUT_BYTES;
    const uint32_t cnt_addr = above + (TPR_WARPS - n_below) * 2u * STG;
    const uint32_t bar_addr = cnt_addr + (16 + MOIRA_N_HIST) * 4;
    if (bar_addr + TPR_WARPS * 16 - base > (uint32_t)TPR_SMEM) __trap();   // cannot happen for base <= 1 KB
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(smem + (cnt_addr - base));
    uint32_t *s_hist = s_cnt + 16;
    uint8_t *lut_ptr = smem + (lut - base);

    // ---- one-time setup: replicated lookup table, counters ----
    if (PL) {
        double *t = reinterpret_cast<double *>(lut_ptr);
        for (int i = threadIdx.x; i < 256 * 32; i += TPR_THREADS) t[i] = a.lut_p[i >> 5];
    } else if (MODE == 0) {
        double2 *t = reinterpret_cast<double2 *>(lut_ptr);
        for (int i = threadIdx.x; i < 256 * 16; i += TPR_THREADS) t[i] = make_double2(a.lut_q[i >> 4], a.lut_e[i >> 4]);
    } else {   // classifier: p in fp32, 64 replicas of 4 bytes per row: lane l reads bank l whatever the quality (LDS.32, one
               // wavefront per warp lookup -- with (p, p (1 - p)) pairs the pass was bound by the shared-memory pipe)
        float *t = reinterpret_cast<float *>(lut_ptr);
        for (int i = threadIdx.x; i < 256 * 64; i += TPR_THREADS) t[i] = (float)a.lut_p[i >> 6];
    }
    for (int i = threadIdx.x; i < 16 + MOIRA_N_HIST; i += TPR_THREADS) s_cnt[i] = 0;
    const uint32_t bar0 = bar_addr + warp * 16;
    if (TMA) {
        if (lane == 0) {
            mbar_init(bar0, 1);
            mbar_init(bar0 + 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    TprCtx c;
    c.lut_lane = lut + (MODE == 2 ? lane * 4 : PL ? lane * 8 : (lane & 15) * 16);
    c.stage_warp = (uint32_t)warp < n_below ? below0 + warp * 2 * STG : above + (warp - n_below) * 2 * STG;
    c.bar0 = bar0;
    c.s_cnt = s_cnt;
    c.s_hist = s_hist;
    c.lane = lane;
    c.warp = warp;
    return c;
}

// The persistent tile loop of one warp: tiles first_tile, first_tile + tile_step, ... of `count` reads
// (taken through `queue` when it is not null), K PMF entries per read.  All staging state is local, so
// a CTA may call this several times with different K (ladder kernel).
// PERREAD (length-bucketed first pass): the decision's K is taken per read, floor(cutoff_r) + 2, instead of FilterArgs::k_dec --
// a bucket may be swept with fewer entries than its reads' decisions need (first_k_cap); what those entries cannot settle
// is rejected by the Newton bound at the read's own K or pushed to the rung that holds it.
template <int K, int MODE, bool PL, bool TMA, int ROLL = 0, bool PERREAD = false>
__device__ __forceinline__ void tpr_tiles(const FilterArgs &a, const CUtensorMap *tmap_ptr, const TprCtx &ctx,
                                          const uint32_t *queue, uint32_t count, uint32_t first_tile, uint32_t total_warps)
{
    constexpr uint32_t STG = TMA ? 32u * CHUNK : (uint32_t)STAGE_BYTES;
    const int lane = ctx.lane;
    const uint32_t lut_lane = ctx.lut_lane, stage_warp = ctx.stage_warp, bar0 = ctx.bar0;
    uint32_t *s_cnt = ctx.s_cnt, *s_hist = ctx.s_hist;
    const uint32_t stage0 = stage_warp + lane * (TMA ? CHUNK : ROW_STRIDE);
    const uint32_t swz = TMA ? (lane & 7) << 4 : 0u;
    const uint32_t n_tiles = (count + 31) >> 5;
    uint32_t tile = first_tile;
    uint32_t it = 0;   // stage jobs issued == consumed so far by this warp
    // Always copy cooperatively (8 consecutive lanes fetch one 128-byte line of one row): 4 L1 tags per
    // LDGSTS instead of 32 when every lane fetches from its own row -- also for scattered ladder rows.
    const bool coop = true;

    // Ns / has-N of every row given with the slab (produced by whoever wrote it): the sweep counts nothing
    const bool marks_given = a.row_marks != nullptr;
    uint32_t vec_steps = 0;   // 16-base vector steps this warp swept with the FP64 recurrence (warp-uniform)
    // entries the decision on one read needs: floor(cutoff on the raw statistic) + 2, at least this sweep's K
    auto kd_of = [&](uint32_t eff, uint32_t n_marks) -> int {
        double c = (a.thr_kind == MOIRA_THR_MAXERRORS) ? a.thr : __dmul_rn((double)eff, a.thr);
        if (a.ambigs == MOIRA_AMBIGS_TREAT_AS_ERRORS) c -= (double)n_marks;
        const double kd = floor(c) + 2.0;
        return kd > (double)K ? (kd > 1.0e6 ? 1000000 : (int)kd) : K;
    };
    auto tile_read = [&](uint32_t t, bool &valid, uint32_t &r_local, ReadGeom &g) {
        uint32_t i = t * 32 + lane;
        valid = t < n_tiles && i < count;
        r_local = 0;
        g.off = 0; g.len = 0; g.eff = 0;
        if (valid) {
            r_local = queue ? queue[i] : i;
            g = read_geom(a, r_local);
        }
    };
    // Stage chunk c of the tile's 32 rows into stage `s` with LDGSTS (cp.async), 16 bytes per lane
    // and instruction.  First pass: the warp copies cooperatively -- 8 consecutive lanes fetch one
    // whole 128-byte line of one row, so every global request is a full line.  Ladder passes (rows
    // scattered through the slab): every lane copies its own row.  One commit group per job.
    auto issue = [&](const ReadGeom &g, uint32_t t, uint32_t c, uint32_t s) {
        if (TMA) {
            if (lane == 0) {
                const uint32_t bar = bar0 + (s & 1) * 8;
                mbar_arrive_expect_tx(bar, 32u * CHUNK);
                tma_tile_g2s(stage_warp + (s & 1) * STG, tmap_ptr, (int)(c * CHUNK), (int)(t * 32u), bar);
            }
            return;
        }
        const uint32_t padded = (g.eff + 15u) & ~15u;
        const uint32_t begin = c * CHUNK;
        const uint32_t dst_stage = (s & 1) * STG;
        if (coop && !a.offsets && !a.lengths && !queue) {
            // uniform stride and length, rows in slab order: addresses follow from the tile's first row (lane 0)
            const uint32_t col = (lane & (CHUNK / 16 - 1)) * 16;
            constexpr int ROWS_PER_INST = 32 / (CHUNK / 16);
            const uint64_t off0 = __shfl_sync(FULL, g.off, 0);
            const uint32_t nrows = __popc(__ballot_sync(FULL, g.eff != 0u || g.len != 0u));
            const bool col_in = begin + col < __shfl_sync(FULL, padded, 0);   // every row has lane 0's length here
            const uint8_t *src = a.slab + off0 + begin + col + (uint64_t)(lane / (CHUNK / 16)) * a.stride;
            const uint32_t dst = stage_warp + dst_stage + (lane / (CHUNK / 16)) * ROW_STRIDE + col;
#pragma unroll
            for (int i = 0; i < 32 / ROWS_PER_INST; i++) {
                const bool in = col_in && (uint32_t)(lane / (CHUNK / 16) + ROWS_PER_INST * i) < nrows;
                cp_async16(dst + i * ROWS_PER_INST * ROW_STRIDE, in ? src + (uint64_t)i * ROWS_PER_INST * a.stride : a.slab,
                           in ? 16u : 0u);
            }
        } else if (coop) {
            const uint32_t col = (lane & (CHUNK / 16 - 1)) * 16;
            constexpr int ROWS_PER_INST = 32 / (CHUNK / 16);
#pragma unroll
            for (int i = 0; i < 32 / ROWS_PER_INST; i++) {
                const int rr = lane / (CHUNK / 16) + ROWS_PER_INST * i;
                const uint64_t o = __shfl_sync(FULL, g.off, rr);
                const uint32_t pd = __shfl_sync(FULL, padded, rr);
                const bool in = begin + col < pd;
                cp_async16(stage_warp + dst_stage + rr * ROW_STRIDE + col, a.slab + (in ? o + begin + col : 0), in ? 16u : 0u);
            }
        } else {
#pragma unroll
            for (int i = 0; i < CHUNK / 16; i++) {
                const bool in = begin + i * 16 < padded;
                cp_async16(stage0 + dst_stage + i * 16, a.slab + (in ? g.off + begin + i * 16 : 0), in ? 16u : 0u);
            }
        }
        cp_async_commit();
    };
    // wait for the oldest outstanding job of this lane, then make every lane's copies visible
    auto consume = [&](bool next_issued, uint32_t job) {
        if (TMA) {
            mbar_wait(bar0 + (job & 1) * 8, (job >> 1) & 1);
            return;
        }
        if (next_issued) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncwarp();
    };

    bool valid, nvalid;
    uint32_t r_local, nr_local;
    ReadGeom g, ng;
    tile_read(tile, valid, r_local, g);
    if (tile < n_tiles) issue(g, tile, 0, it);
    // longest / shortest read of the tile: reduced where the tile's geometry is complete anyway (before the loop, and at the
    // end of the previous tile), so that the loop head does not wait on loads
    uint32_t maxeff = __reduce_max_sync(FULL, g.eff);
    uint32_t mineff = __reduce_min_sync(FULL, g.eff);

    while (tile < n_tiles) {
        const uint32_t next_tile = tile + total_warps;
        tile_read(next_tile, nvalid, nr_local, ng);
        // this tile's marks: requested here, first needed by the epilogue -- the sweep in between hides the latency (a load
        // carried from the previous iteration made the loop head wait for the loads just issued for the next tile)
        const uint32_t marks = (marks_given && valid) ? __ldg(a.row_marks + a.base + r_local) : 0u;
        uint32_t nch = (maxeff + CHUNK - 1) / CHUNK;

        using acc_t = typename std::conditional<MODE == 2, float, double>::type;
        acc_t P[K];
#pragma unroll
        for (int j = 0; j < K; j++) P[j] = 0;
        if (MODE == 0) P[0] = 1;
        uint32_t ns = 0, has_n = 0;
        uint32_t processed = 0;
        bool skip_math = false;   // warp-uniform

        if (nch == 0) {   // tile of empty reads: its (empty) stage job still has to be consumed
            const bool nx = next_tile < n_tiles;
            __syncwarp();
            if (nx) issue(ng, next_tile, 0, it + 1);
            consume(nx, it);
            it++;
        }
        for (uint32_t c = 0; c < nch; c++) {
            const bool nx = c + 1 < nch || next_tile < n_tiles;
            __syncwarp();   // every lane is done reading the stage the next job overwrites
            if (c + 1 < nch) issue(g, tile, c + 1, it + 1);
            else if (next_tile < n_tiles) issue(ng, next_tile, 0, it + 1);
            consume(nx, it);
            const uint32_t row = stage0 + (it & 1) * STG;
            it++;

            const uint32_t cbeg = c * CHUNK;
            const uint32_t cend = maxeff - cbeg < CHUNK ? maxeff - cbeg : CHUNK;   // warp-uniform
            const uint32_t cfull = mineff > cbeg ? min((mineff - cbeg) & ~15u, cend) : 0u;   // warp-uniform
            if (!skip_math) {
                if (marks_given) sweep_chunk<K, MODE, PL, true, false, ROLL, acc_t>(row, swz, cfull, cend, (int)g.eff - (int)cbeg, lut_lane, P, ns, has_n);
                else sweep_chunk<K, MODE, PL, true, true, ROLL, acc_t>(row, swz, cfull, cend, (int)g.eff - (int)cbeg, lut_lane, P, ns, has_n);
                processed = cbeg + cend;
                vec_steps += (cend + 15u) >> 4;
            } else if (!marks_given) {
                sweep_chunk<K, MODE, PL, false, true, ROLL, acc_t>(row, swz, cfull, cend, (int)g.eff - (int)cbeg, lut_lane, P, ns, has_n);
            }
            if constexpr (MODE == 0) if (c + 1 < nch && !skip_math) {
                // Early exit: sum_{j<K} P_k[j] never increases with k, so once it is safely below
                // 1-alpha these K entries cannot reach the quantile.  Taken warp-wide only: the
                // remaining chunks are still staged and scanned for N/n (Ns stays exact) but the
                // FP64 sweep -- the binding resource -- is skipped.
                double tracked = P[0];
#pragma unroll
                for (int j = 1; j < K; j++) tracked += P[j];
                bool certain = valid && tracked < a.oma - 1e-9;
                const bool finished = !valid || processed >= g.eff;
                bool all_certain = __all_sync(FULL, certain || finished) && __any_sync(FULL, certain);
                if constexpr (PERREAD) {
                    if (K >= 2 && all_certain && !a.exact && a.first_k_cap) {
                        // the read's own K without its N/n (not all counted yet): at least the K of the epilogue, so the test is on the safe side
                        const int kdl = kd_of(g.eff, 0u);
                        if (kdl > K) certain = valid && kdl <= 8 && cascade_bound<K>(P, kdl) < a.oma - 1e-9;
                        all_certain = __all_sync(FULL, certain || finished) && __any_sync(FULL, certain);
                    }
                } else if (K >= 2 && a.k_dec > K && all_certain) {
                    // a cascade launch stops only where the reject is certain at the decision's k_dec (the bound is at
                    // least the tracked mass, so the test above is a cheap necessary condition)
                    certain = valid && cascade_bound<K>(P, a.k_dec) < a.oma - 1e-9;
                    all_certain = __all_sync(FULL, certain || finished) && __any_sync(FULL, certain);
                }
                if (all_certain) {
                    skip_math = true;
                    // nothing left to count either: the chunk already in flight is consumed and the tile ends there
                    if (marks_given) nch = c + 2;
                }
            }
        }

        // ---- per-read epilogue ----
        if (marks_given) { ns = marks & 0x7FFFFFFFu; has_n = marks >> 31; }
        ReadResult res;
        res.ns = (int)ns;
        res.has_n = has_n != 0u;
        res.numeric = false;
        res.escalate = false;
        res.next_rung = a.rung + 1;
        res.processed = processed < g.eff ? processed : g.eff;
        if constexpr (MODE == 2) {
            // Upper quantile of the error count by Cornish-Fisher with the Poisson skew bound:
            // j* <~ mu + z*sigma + (z^2-1)/6; K = j* + 1 plus margin (zc holds the constants).
            // An under-estimate only costs one more rung; the result is never affected.
            const double var = (double)P[0] > (double)P[1] ? (double)P[0] - (double)P[1] : 0.0;   // sum p (1 - p)
            double kn = (double)P[0] + a.z * sqrt(var) + a.zc + 1e-4 * (double)P[0];   // fp32 sums: relative error << 1e-4
            bool hopeless = false;
            if (!a.exact) {   // a decision needs at most floor(cutoff on the raw statistic) + 2 entries
                double c = (a.thr_kind == MOIRA_THR_MAXERRORS) ? a.thr : __dmul_rn((double)g.eff, a.thr);
                if (a.ambigs == MOIRA_AMBIGS_TREAT_AS_ERRORS) c -= (double)ns;
                const double kd = floor(c) + 2.0;
                if (kd < kn) kn = kd;
                // Certain rejects without any FP64 sweep.  X = number of errors is a sum of independent Bernoulli(p_i) with mean
                // mu, so for 0 < t < mu the Chernoff bound gives  P(X <= t) <= exp(t - mu + t ln(mu / t))  (t = 0: exp(-mu)).
                // With t = kd - 1: if that is below 1 - alpha, then acc[kd - 1] <= 1 - alpha, j* >= kd and the statistic is at
                // least kd - 1 = floor(c) + 1 > c -- the read is rejected whatever its exact value, and it is written here as the
                // kd-entry sweep would write it (lower bound kd - 1, MOIRA_FLAG_LOWER_BOUND).  mu_lo is a rigorous lower bound of
                // mu from the fp32 sum: (eff + 2) roundings of at most 2^-24 each, doubled; the 1e-9 margin covers the rounding
                // of the reference's own double sums (<= L 2^-52 relative) by orders of magnitude.
                const double t = kd - 1.0;
                const double mu_lo = (double)P[0] * (1.0 - ((double)g.eff + 2.0) * 1.2e-7);
                if (kd >= 1.0 && kd <= 1.0e6 && mu_lo > t) {
                    const double bound = t > 0.0 ? exp(t - mu_lo + t * log(mu_lo / t)) : exp(-mu_lo);
                    hopeless = bound < a.oma - 1e-9;
                    res.ee_raw = t;
                }
            }
            const int kneed = kn > 1.0e6 ? 1000000 : (kn < 2.0 ? 2 : (int)kn);
            if (a.rung < 0) {   // classifier run as the first pass (classify-first): every read is handed on
                const unsigned vm = __ballot_sync(FULL, valid);
                if (vm && lane == 0) atomicAdd(&s_cnt[MOIRA_CNT_CLASSIFIED], (uint32_t)__popc(vm));
            }
            if (!a.exact && __any_sync(FULL, valid && hopeless)) {
                res.resolved = false;
                finish_read(a, valid && hopeless, r_local, g, res, s_cnt, s_hist, lane);
            }
            push_read(a, valid && !hopeless, pick_rung(a, kneed), r_local, lane);
            tile = next_tile;
            valid = nvalid; r_local = nr_local; g = ng;
            maxeff = __reduce_max_sync(FULL, g.eff);
            mineff = __reduce_min_sync(FULL, g.eff);
            continue;
        } else {
        if constexpr (MODE == 0) {
            res.resolved = cdf_quantile<K>(P, a.oma, res.ee_raw);
            const int kd = PERREAD ? ((K >= 2 && !a.exact && a.first_k_cap) ? kd_of(g.eff, ns) : K)
                                   : ((K >= 2 && a.k_dec > K) ? a.k_dec : K);
            if (K >= 2 && a.direct_rung && a.rung < 0 && !res.resolved && res.processed >= g.eff)   // first pass only: the rungs have their queues read already
                res.next_rung = rung_from_two_entries(a, P[0], P[K >= 2 ? 1 : 0], g.eff, ns, K);
            if (res.processed < g.eff) { res.resolved = false; res.ee_raw = (double)(kd - 1); }
            else if (kd > K && !res.resolved) {
                // (c_ratio[] of the bound reaches kd = 8; a read that needs more -- longer than the caller said -- goes to its rung)
                if (kd <= 8 && cascade_bound<K>(P, kd) < a.oma - 1e-9) res.ee_raw = (double)(kd - 1);   // j* >= kd: a certain reject, without the kd-entry sweep
                else {
                    res.escalate = true;
                    if constexpr (PERREAD) res.next_rung = pick_rung(a, kd);   // swept once more, with the entries its own decision needs
                }
            }
        } else if constexpr (MODE == 1) {
            const double lam = P[0];
            if (a.mode == MOIRA_MODE_EXPECTED_ERROR) {
                res.ee_raw = lam;
                res.resolved = true;
            } else {
                // moira.py:1668-1677.  Decision mode needs terms j = 0 .. floor(c)+1 only.
                int jlim = MOIRA_FACT_N - 1;
                if (!a.exact) {
                    double cutoff = (a.thr_kind == MOIRA_THR_MAXERRORS) ? a.thr : __dmul_rn((double)g.eff, a.thr);
                    if (a.ambigs == MOIRA_AMBIGS_TREAT_AS_ERRORS) cutoff -= (double)ns;
                    double jl = floor(cutoff) + 1.0;
                    if (jl < (double)jlim) jlim = jl < 0.0 ? -1 : (int)jl;
                }
                const double el = exp(-lam);
                double acc = 0.0, prev = 0.0;
                res.resolved = false;
                res.ee_raw = jlim > 0 ? (double)jlim : 0.0;
                for (int j = 0; valid && j <= jlim; j++) {
                    double pw = pow(lam, (double)j);
                    if (isinf(pw)) { res.numeric = true; break; }     // reference raises OverflowError here
                    double pmf = __ddiv_rn(__dmul_rn(el, pw), c_fact[j]);
                    prev = acc;
                    acc = __dadd_rn(prev, pmf);
                    if (acc > a.oma) {
                        double e = __dadd_rn((double)(j - 1), __ddiv_rn(__dsub_rn(a.oma, prev), __dsub_rn(acc, prev)));
                        res.ee_raw = e < 0.0 ? 0.0 : e;
                        res.resolved = true;
                        break;
                    }
                }
                if (!res.resolved && !res.numeric && (a.exact || jlim == MOIRA_FACT_N - 1)) res.numeric = true;
            }
        }
        finish_read(a, valid, r_local, g, res, s_cnt, s_hist, lane);
        }

        tile = next_tile;
        valid = nvalid; r_local = nr_local; g = ng;
        maxeff = __reduce_max_sync(FULL, g.eff);
        mineff = __reduce_min_sync(FULL, g.eff);
    }
    // FP64 operations the sweeps of this warp executed (thread level: 32 lanes x 16 positions per vector step)
    constexpr uint32_t OPB = MODE == 0 ? (uint32_t)(3 * K - 2) + (PL ? 1u : 0u) : (MODE == 1 ? 1u : 0u);   // the classifier sums in fp32
    if (lane == 0 && vec_steps && a.counters) atomicAdd(&a.counters[MOIRA_CNT_FP64_OPS], (unsigned long long)vec_steps * (512ull * OPB));
}

// MODE 0: Poisson-binomial with K entries.  MODE 1: Lambda accumulation (K == 1).
// MODE 2: ladder classifier (K == 2): mean/variance of the error count -> rung.
// TMA: uniform-stride first pass; tiles arrive as swizzled 2-D tensor boxes (one UTMALDG per warp and
// chunk) instead of cp.async pieces.
template <int K, int MODE, bool EQP, bool TMA>
__global__ void __launch_bounds__(tpr_warps(K) * 32, 1) tpr_kernel(const FilterArgs a, const __grid_constant__ CUtensorMap tmap)
{
    constexpr int TPR_WARPS = tpr_warps(K);
    constexpr bool PL = (EQP && MODE == 0 && !(MOIRA_PAIR_LUT && K <= 8 && K > MOIRA_PL_MAXK)) || MODE == 1;
    extern __shared__ __align__(128) uint8_t smem[];
    if (a.queue && (a.seg_count ? *a.seg_count : *a.queue_count) == 0) return;   // empty rung / segment: nothing to set up
    if (a.policy && *a.policy != a.policy_want) return;                          // the pilot chose the other first pass
    const TprCtx ctx = tpr_setup<TPR_WARPS, MODE, PL, TMA>(a, smem);
    const uint32_t *queue = a.queue && a.seg_start ? a.queue + *a.seg_start : a.queue;
    const uint32_t count = a.queue ? (a.seg_count ? *a.seg_count : *a.queue_count) : a.n;
    tpr_tiles<K, MODE, PL, TMA>(a, &tmap, ctx, queue, count, a.tile0 + blockIdx.x * TPR_WARPS + ctx.warp, gridDim.x * TPR_WARPS);
    __syncthreads();
    flush_counters(a, ctx.s_cnt, ctx.s_hist);
}

__global__ void policy_kernel(const uint32_t *queue_count, uint32_t max_pushed, uint32_t *policy)
{
    *policy = *queue_count > max_pushed ? 1u : 0u;
}

// Cascade with a first stage of K1 = 2 .. 5 entries (decisions that need k_first = 5 .. 8).  Verdict of the first pilot
// (two-entry sweep over n_pilot reads, escalations counted in queue 0): at most `max_pushed` handed on -> *policy = 2.
// Otherwise MOIRA_POLICY_UNDECIDED: the second pilot (k_first-entry sweep, with a histogram of floor(ee)) runs.
__global__ void policy_first_kernel(const uint32_t *queue_count, uint32_t max_pushed, uint32_t *policy)
{
    *policy = *queue_count > max_pushed ? MOIRA_POLICY_UNDECIDED : 2u;
}
// Length-bucketed first pass, decision mode: verdict of the pilot that swept its reads with at most first_k_cap entries per bucket.
// What those entries did not settle sits in the rung queues; few enough -> the rest of the batch is swept the same way (1),
// else every bucket gets the K of its own cutoff (0).
__global__ void policy_sorted_kernel(const uint32_t *queue_counts, uint32_t max_pushed, uint32_t *policy)
{
    uint32_t pushed = 0;
    for (int r = 0; r <= N_TPR_RUNGS; r++) pushed += queue_counts[r];
    *policy = pushed > max_pushed ? 0u : 1u;
}
// Verdict of the second pilot: the cheapest first stage in issue cycles per base, 2 F + O with F = 3 K - 2 FP64 operations
// and O = 4.3 others (profiles/r02_k2_regions_before_fix.md), counting every read with j* >= K1 as swept again with k_first
// entries through the queue (x 1.1: gathered rows, no TMA) -- the Newton bound settles some of them, so the estimate is on
// the safe side.  floor(ee) >= K1 - 1  <=>  j* >= K1 (up to reads with N, which need fewer entries).  0: single sweep.
// Exact mode (every read's statistic is wanted): the same choice with the ladder's cost for the reads K1 entries do not settle.
__global__ void policy_second_kernel(const uint32_t *jhist, int k_first, int exact, uint32_t *policy)
{
    if (*policy != MOIRA_POLICY_UNDECIDED) return;
    double total = 0.0, tail[17];
    for (int i = 0; i < 16; i++) total += (double)jhist[i];
    tail[16] = 0.0;
    for (int i = 15; i >= 0; i--) tail[i] = tail[i + 1] + (double)jhist[i];
    auto cyc = [](int k) { return 2.0 * (3 * k - 2) + 4.3; };
    const double full = cyc(k_first);
    double best = full;
    uint32_t choice = 0;
    for (int k1 = 3; k1 <= 5 && k1 < k_first; k1++) {
        double cost = cyc(k1);
        if (!exact) {
            const double esc = total > 0.0 ? tail[k1 - 1] / total : 1.0;
            cost += esc * full * 1.1;
        } else if (total > 0.0) {
            // exact mode: a read with floor(ee) = j >= k1 - 1 (j* = j + 1 >= k1) goes on to the rung that holds j* + 1 entries
            // plus about one of estimate; what the pilot's k_first entries did not settle either costs the same whatever k1 is
            for (int j = k1 - 1; j <= k_first - 2 && j < 16; j++) cost += 1.1 * ((double)jhist[j] / total) * cyc(j + 3);
        }
        if (cost < best * 0.97) { best = cost; choice = (uint32_t)k1; }
    }
    *policy = choice;
}

// Two launches like the length-bucketed first pass: rungs 1..8 (K <= 12) on 16 warps per CTA, rungs 9..19 on 8.
template <bool EQP, bool WIDE>
__global__ void __launch_bounds__(WIDE ? 256 : 512, 1) ladder_tpr_kernel(const FilterArgs a0)
{
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int WARPS = WIDE ? 8 : 16;
    constexpr int R0 = WIDE ? MOIRA_LADDER_SPLIT + 1 : 1, R1 = WIDE ? N_TPR_RUNGS : MOIRA_LADDER_SPLIT;   // rung_cap(8) == 12, rung_cap(14) == 24
    uint32_t counts[N_TPR_RUNGS + 1];
    uint32_t any = 0;
#pragma unroll
    for (int r = R0; r <= R1; r++) {
        const uint32_t c = a0.queue_counts[r];
        counts[r] = c < a0.queue_cap ? c : a0.queue_cap;
        any |= counts[r];
    }
    if (!any) return;
    const TprCtx ctx = tpr_setup<WARPS, 0, EQP, false>(a0, smem);
    FilterArgs a = a0;
    a.rung = N_TPR_RUNGS;                      // escalations of this kernel land in rung N_TPR_RUNGS + 1
    const uint32_t W = gridDim.x * WARPS, gw = blockIdx.x * WARPS + ctx.warp;
    uint32_t before = 0;                       // tiles of the rungs already walked: keeps the load balanced
#define MOIRA_RUNG(r, k)                                                                                      \
    if constexpr ((r) >= R0 && (r) <= R1) {                                                                   \
        if (counts[r]) {                                                                                      \
            const uint32_t first = (gw + W - before % W) % W;                                                 \
            tpr_tiles<k, 0, EQP, false, MOIRA_LADDER_ROLL>(a, nullptr, ctx, a0.queues + (size_t)(r) * a0.queue_cap, counts[r], first, W); \
            before += (counts[r] + 31) >> 5;                                                                  \
        }                                                                                                     \
    }
    MOIRA_RUNG(1, 3) MOIRA_RUNG(2, 4) MOIRA_RUNG(3, 5) MOIRA_RUNG(4, 6) MOIRA_RUNG(5, 7)
    MOIRA_RUNG(6, 8) MOIRA_RUNG(7, 10) MOIRA_RUNG(8, 12) MOIRA_RUNG(9, 14) MOIRA_RUNG(10, 16) MOIRA_RUNG(11, 18)
    MOIRA_RUNG(12, 20) MOIRA_RUNG(13, 22) MOIRA_RUNG(14, 24) MOIRA_RUNG(15, 28) MOIRA_RUNG(16, 32) MOIRA_RUNG(17, 40)
    MOIRA_RUNG(18, 48) MOIRA_RUNG(19, 64)
    static_assert(N_TPR_RUNGS == 19 && rung_cap(19) == 64 && rung_cap(6) == 8 && rung_cap(8) == 12, "rung table and ladder kernel out of step");
#undef MOIRA_RUNG
    __syncthreads();
    flush_counters(a, ctx.s_cnt, ctx.s_hist);
}

// First pass over a length-sorted permutation, every K template in ONE launch: segment g of the queue
// (seg_start[g], seg_count[g]; written by len_scan_kernel) is swept with first_pass_k(g) entries.
// WIDE = false: the groups with K <= 12 on 16 warps per CTA (<= 128 registers per thread, four warps per scheduler like the
// single-K kernels); WIDE = true: K = 14 .. 32 on 8 warps.  Two launches; one whose groups are all empty returns before
// any set-up.
template <bool EQP, bool WIDE>
__global__ void __launch_bounds__(WIDE ? 256 : 512, 1) sorted_first_kernel(const FilterArgs a, const uint32_t *seg_start, const uint32_t *seg_count)
{
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int WARPS = WIDE ? 8 : 16;
    constexpr int G0 = WIDE ? 9 : 0, G1 = WIDE ? N_FIRST_K : 9;   // first_pass_k(8) == 12
    static_assert(first_pass_k(8) == 12 && first_pass_k(9) == 14 && N_FIRST_K == 17, "group table and sorted first pass out of step");
    uint32_t any = 0;
#pragma unroll
    for (int g = G0; g < G1; g++) any |= seg_count[g];
    if (!any) return;
    const TprCtx ctx = tpr_setup<WARPS, 0, EQP, false>(a, smem);
    const uint32_t W = gridDim.x * WARPS, gw = blockIdx.x * WARPS + ctx.warp;
    uint32_t before = 0;
#define MOIRA_GROUP(g, k)                                                                              \
    if constexpr ((g) >= G0 && (g) < G1) {                                                             \
        const uint32_t cnt = seg_count[g];                                                             \
        if (cnt) {                                                                                     \
            const uint32_t first = (gw + W - before % W) % W;                                          \
            tpr_tiles<k, 0, EQP, false, 0, true>(a, nullptr, ctx, a.queue + seg_start[g], cnt, first, W); \
            before += (cnt + 31) >> 5;                                                                 \
        }                                                                                              \
    }
    MOIRA_GROUP(0, 2) MOIRA_GROUP(1, 3) MOIRA_GROUP(2, 4) MOIRA_GROUP(3, 5) MOIRA_GROUP(4, 6) MOIRA_GROUP(5, 7)
    MOIRA_GROUP(6, 8) MOIRA_GROUP(7, 10) MOIRA_GROUP(8, 12) MOIRA_GROUP(9, 14) MOIRA_GROUP(10, 16) MOIRA_GROUP(11, 18)
    MOIRA_GROUP(12, 20) MOIRA_GROUP(13, 22) MOIRA_GROUP(14, 24) MOIRA_GROUP(15, 28) MOIRA_GROUP(16, 32)
#undef MOIRA_GROUP
    __syncthreads();
    flush_counters(a, ctx.s_cnt, ctx.s_hist);
}

// ==================================================================================================
// warp-per-read: K = 32*M entries, lane l owns P[l*M + m]
// ==================================================================================================
template <int M>
__global__ void __launch_bounds__(WPR_THREADS) wpr_kernel(const FilterArgs a)
{
    __shared__ double2 s_lut[256];
    __shared__ uint32_t s_cnt[16 + MOIRA_N_HIST];
    if (*a.queue_count == 0) return;
    for (int i = threadIdx.x; i < 256; i += WPR_THREADS) s_lut[i] = make_double2(a.lut_q[i], a.lut_e[i]);
    for (int i = threadIdx.x; i < 16 + MOIRA_N_HIST; i += WPR_THREADS) s_cnt[i] = 0;
    __syncthreads();
    uint32_t *s_hist = s_cnt + 16;
    const int lane = threadIdx.x & 31;
    const uint32_t count = *a.queue_count < a.queue_cap ? *a.queue_count : a.queue_cap;
    const uint32_t total_warps = gridDim.x * (WPR_THREADS / 32);
    constexpr int K = 32 * M;

    for (uint32_t i = blockIdx.x * (WPR_THREADS / 32) + (threadIdx.x >> 5); i < count; i += total_warps) {
        const uint32_t r_local = a.queue[i];
        const ReadGeom g = read_geom(a, r_local);
        const uint8_t *row = a.slab + g.off;
        double P[M];
#pragma unroll
        for (int m = 0; m < M; m++) P[m] = 0.0;
        if (lane == 0) P[0] = 1.0;
        uint32_t ns = 0, has_n = 0;
        bool dead = false;
        uint32_t pos = 0;
        uint4 cur = make_uint4(0xFDFDFDFDu, 0xFDFDFDFDu, 0xFDFDFDFDu, 0xFDFDFDFDu);
        if (g.eff) cur = __ldg(reinterpret_cast<const uint4 *>(row));
        uint32_t swept = 0;   // bases covered by the FP64 sweep
        uint32_t n_swept = 0; // called bases swept (warp-uniform)
        for (; pos < g.eff; pos += 16) {
            uint4 nxt = cur;
            if (pos + 16 < g.eff) nxt = __ldg(reinterpret_cast<const uint4 *>(row + pos + 16));
            const int rem = (int)g.eff - (int)pos;
            uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
            if (rem < 16) {
#pragma unroll
                for (int k = 0; k < 4; k++) w[k] = mask_word(w[k], rem - 4 * k);
            }
            count_marks4(w, ns, has_n);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (dead) continue;                          // keep counting N/n, skip the FP64 sweep
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const uint32_t q8 = (w[k] >> (8 * b)) & 0xFFu;
                    if (q8 >= 0xFDu) continue;               // warp-uniform: padding / N / n leave P untouched
                    const double2 qe = s_lut[q8];
                    n_swept++;
                    double up = __shfl_up_sync(FULL, P[M - 1], 1);
                    if (lane == 0) up = 0.0;
#pragma unroll
                    for (int m = M - 1; m >= 1; m--)
                        P[m] = __dadd_rn(__dmul_rn(qe.x, P[m]), __dmul_rn(qe.y, P[m - 1]));
                    P[0] = __dadd_rn(__dmul_rn(qe.x, P[0]), __dmul_rn(qe.y, up));
                }
            }
            cur = nxt;
            if (!dead) swept = pos + 16;
            if (!dead && ((pos >> 4) & 3) == 3) {   // every 64 bases: can these K entries still reach 1-alpha?
                double t = 0.0;
#pragma unroll
                for (int m = 0; m < M; m++) t += P[m];
#pragma unroll
                for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
                dead = t < a.oma - 1e-9;
            }
        }
        const uint32_t processed = swept < g.eff ? swept : g.eff;

        // cumulative sum in index order across lanes (bernoullimodule.c:233-244)
        double acc = 0.0, prev = 0.0, ee = (double)(K - 1);
        bool found = false;
        if (!dead) {
            for (int l = 0; l < 32 && !found; l++) {
                double my_prev = 0.0, my_acc = acc;
                int my_j = -1;
                if (lane == l) {
#pragma unroll
                    for (int m = 0; m < M; m++) {
                        if (my_j < 0) {
                            double p2 = my_acc;
                            my_acc = __dadd_rn(p2, P[m]);    // acc[0] = 0 + P[0] == P[0] bitwise
                            if (my_acc > a.oma) { my_j = l * M + m; my_prev = p2; }
                        }
                    }
                }
                acc = __shfl_sync(FULL, my_acc, l);
                int j = __shfl_sync(FULL, my_j, l);
                if (j >= 0) {
                    prev = __shfl_sync(FULL, my_prev, l);
                    found = true;
                    ee = (j == 0) ? 0.0
                                  : __dadd_rn((double)(j - 1), __ddiv_rn(__dsub_rn(a.oma, prev), __dsub_rn(acc, prev)));
                }
            }
        }
        ReadResult res;
        res.ee_raw = ee;
        res.processed = processed;
        res.ns = (int)ns;
        res.has_n = has_n != 0u;
        res.resolved = found;
        res.numeric = false;
        res.escalate = false;
        finish_read(a, lane == 0, r_local, g, res, s_cnt, s_hist, lane);
        if (lane == 0 && a.counters) atomicAdd(&a.counters[MOIRA_CNT_FP64_OPS], (unsigned long long)n_swept * (96ull * M));
    }
    __syncthreads();
    flush_counters(a, s_cnt, s_hist);
}

// ==================================================================================================
// block-per-read, P in shared memory: any K up to BLK_CAP (last rung)
// ==================================================================================================
__global__ void __launch_bounds__(BLK_THREADS) blk_kernel(const FilterArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    double *P = reinterpret_cast<double *>(smem);
    double2 *s_lut = reinterpret_cast<double2 *>(smem + (size_t)BLK_CAP * 8);
    __shared__ uint32_t s_cnt[16 + MOIRA_N_HIST];
    __shared__ double s_res[4];
    __shared__ int s_j;
    if (*a.queue_count == 0) return;
    for (int i = threadIdx.x; i < 256; i += BLK_THREADS) s_lut[i] = make_double2(a.lut_q[i], a.lut_e[i]);
    for (int i = threadIdx.x; i < 16 + MOIRA_N_HIST; i += BLK_THREADS) s_cnt[i] = 0;
    __syncthreads();
    uint32_t *s_hist = s_cnt + 16;
    const int lane = threadIdx.x & 31;
    const uint32_t count = *a.queue_count < a.queue_cap ? *a.queue_count : a.queue_cap;

    for (uint32_t i = blockIdx.x; i < count; i += gridDim.x) {
        const uint32_t r_local = a.queue[i];
        const ReadGeom g = read_geom(a, r_local);
        const uint8_t *row = a.slab + g.off;
        const int K = (int)(g.eff + 1 < (uint32_t)BLK_CAP ? g.eff + 1 : (uint32_t)BLK_CAP);
        for (int j = threadIdx.x; j < K; j += BLK_THREADS) P[j] = j == 0 ? 1.0 : 0.0;
        __syncthreads();
        int ns = 0, nn = 0;   // nn = non-N bases swept so far: entries above nn are still zero
        bool has_n = false;
        for (uint32_t pos = 0; pos < g.eff; pos++) {
            const uint32_t q8 = row[pos];
            if (q8 >= 0xFEu) { ns++; has_n |= (q8 == 0xFFu); continue; }
            if (q8 == 0xFDu) continue;
            const double2 qe = s_lut[q8];
            const int hi = nn + 1 < K - 1 ? nn + 1 : K - 1;
            for (int top = hi; top >= 0; top -= BLK_THREADS) {
                const int j = top - (int)threadIdx.x;
                double nv = 0.0;
                if (j >= 0) {
                    double pj = P[j];
                    double pm = j > 0 ? P[j - 1] : 0.0;
                    nv = j > 0 ? __dadd_rn(__dmul_rn(qe.x, pj), __dmul_rn(qe.y, pm)) : __dmul_rn(qe.x, pj);
                }
                __syncthreads();
                if (j >= 0) P[j] = nv;
                __syncthreads();
            }
            nn++;
        }
        if (threadIdx.x == 0) {
            double acc = 0.0, prev = 0.0;
            int js = -1;
            for (int j = 0; j < K; j++) {
                prev = acc;
                acc = __dadd_rn(prev, P[j]);
                if (acc > a.oma) { js = j; break; }
            }
            s_j = js;
            s_res[0] = prev;
            s_res[1] = acc;
        }
        __syncthreads();
        ReadResult res;
        const int js = s_j;
        res.resolved = js >= 0;
        res.ee_raw = js < 0 ? (double)(K - 1)
                   : js == 0 ? 0.0
                             : __dadd_rn((double)(js - 1), __ddiv_rn(__dsub_rn(a.oma, s_res[0]), __dsub_rn(s_res[1], s_res[0])));
        res.escalate = false;
        res.processed = g.eff;
        res.ns = ns;
        res.has_n = has_n;
        res.numeric = false;
        if (threadIdx.x < 32) finish_read(a, threadIdx.x == 0, r_local, g, res, s_cnt, s_hist, lane);
        __syncthreads();
    }
    __syncthreads();
    flush_counters(a, s_cnt, s_hist);
}

// ==================================================================================================
// length bucketing of ragged batches: counting sort of the read indices by padded length, so that the
// 32 reads of a warp tile have (almost) the same length and every tile runs with the K its own
// cutoff needs.  Three tiny kernels, no host synchronisation.
// ==================================================================================================
__device__ __forceinline__ uint32_t len_bucket(uint32_t eff)
{
    const uint32_t b = (eff + 15u) >> 4;
    return b < LEN_BUCKETS ? b : LEN_BUCKETS - 1;
}

// warp-aggregated increment of counter[key]: one atomic per distinct key and warp; returns the old
// value + the lane's rank among the lanes with the same key (a unique slot).  Convergent call.
__device__ __forceinline__ uint32_t warp_agg_inc(uint32_t *counter, uint32_t key, bool active, int lane)
{
    const unsigned am = __ballot_sync(FULL, active);
    uint32_t slot = 0;
    if (active) {
        const unsigned peers = __match_any_sync(am, key);
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&counter[key], (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        slot = base + __popc(peers & ((1u << lane) - 1u));
    }
    return slot;
}

constexpr int LS_CHUNK = 4096;   // reads per block iteration of the length-sort kernels

__global__ void __launch_bounds__(256) len_hist_kernel(const FilterArgs a, uint32_t *hist)
{
    __shared__ uint32_t s_h[LEN_BUCKETS];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < LEN_BUCKETS; i += 256) s_h[i] = 0;
    __syncthreads();
    const uint32_t cnt = a.n - a.first_read, n_round = (cnt + 255u) & ~255u;     // reads [first_read, n) of the sub-batch
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n_round; i += gridDim.x * 256) {
        const bool ok = i < cnt;
        warp_agg_inc(s_h, ok ? len_bucket(read_geom(a, a.first_read + i).eff) : 0u, ok, lane);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LEN_BUCKETS; i += 256)
        if (s_h[i]) atomicAdd(&hist[i], s_h[i]);
}

// one block: exclusive scan of the histogram, and the queue segment of every first-pass K template
__global__ void __launch_bounds__(1024) len_scan_kernel(const FilterArgs a, const uint32_t *hist, uint32_t *bucket_start,
                                                        uint32_t *cursor, uint32_t *group_start, uint32_t *group_count,
                                                        int single_group)
{
    __shared__ uint32_t s_part[1024];
    __shared__ uint32_t s_gs[N_FIRST_K], s_gc[N_FIRST_K];
    constexpr int PER = LEN_BUCKETS / 1024;
    const int t = threadIdx.x;
    // the cap may hang on a verdict left on the device by a pilot (decision mode: capped buckets only pay when few reads escalate)
    const int cap = (a.first_k_cap && (!a.cap_policy || *a.cap_policy == 1u)) ? a.first_k_cap : 0;
    if (t < N_FIRST_K) { s_gs[t] = 0xFFFFFFFFu; s_gc[t] = 0; }
    uint32_t v[PER], sum = 0;
#pragma unroll
    for (int i = 0; i < PER; i++) { v[i] = hist[t * PER + i]; sum += v[i]; }
    s_part[t] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {       // Hillis-Steele inclusive scan of the per-thread sums
        uint32_t x = t >= o ? s_part[t - o] : 0u;
        __syncthreads();
        s_part[t] += x;
        __syncthreads();
    }
    uint32_t run = s_part[t] - sum;
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int b = t * PER + i;
        bucket_start[b] = run;
        cursor[b] = 0;
        if (v[i]) {
            int g = 0;
            if (!single_group) {
                // K a decision needs for the longest read of the bucket: floor(cutoff) + 2
                const double cutoff = a.thr_kind == MOIRA_THR_MAXERRORS ? a.thr : __dmul_rn((double)(b * 16), a.thr);
                double kd = cutoff < 0.0 ? 2.0 : floor(cutoff) + 2.0;
                if (b == LEN_BUCKETS - 1) kd = 1e9;   // the catch-all bucket (lengths unknown to the host): the largest K
                // exact mode with a ladder behind the first pass: the pass only has to settle the reads that are cheap to settle
                if (cap && kd > (double)cap) kd = (double)cap;
                while (g < N_FIRST_K - 1 && (double)first_pass_k(g) < kd) g++;
            }
            atomicMin(&s_gs[g], run);
            atomicAdd(&s_gc[g], v[i]);
        }
        run += v[i];
    }
    __syncthreads();
    if (t < N_FIRST_K) { group_start[t] = s_gc[t] ? s_gs[t] : 0u; group_count[t] = s_gc[t]; }
}

// Each block takes contiguous chunks of LS_CHUNK reads, counts them per bucket in shared memory,
// reserves one contiguous range per bucket with a single global atomic, and writes its reads there:
// neighbours in the queue stay neighbours in the slab (DRAM locality of the gather), and the global
// atomics are per (block-chunk, bucket) instead of per read.
__global__ void __launch_bounds__(256) len_scatter_kernel(const FilterArgs a, const uint32_t *bucket_start, uint32_t *cursor,
                                                          uint32_t *queue)
{
    __shared__ uint32_t s_cnt[LEN_BUCKETS];
    __shared__ uint32_t s_base[LEN_BUCKETS];
    const int lane = threadIdx.x & 31;
    for (uint32_t c0 = a.first_read + blockIdx.x * LS_CHUNK; c0 < a.n; c0 += gridDim.x * LS_CHUNK) {
        for (int i = threadIdx.x; i < LEN_BUCKETS; i += 256) s_cnt[i] = 0;
        __syncthreads();
        uint32_t slot[LS_CHUNK / 256], bkt[LS_CHUNK / 256];
#pragma unroll
        for (int k = 0; k < LS_CHUNK / 256; k++) {
            const uint32_t r = c0 + k * 256 + threadIdx.x;
            const bool ok = r < a.n;
            bkt[k] = ok ? len_bucket(read_geom(a, r).eff) : 0u;
            slot[k] = warp_agg_inc(s_cnt, bkt[k], ok, lane);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < LEN_BUCKETS; i += 256)
            if (s_cnt[i]) s_base[i] = bucket_start[i] + atomicAdd(&cursor[i], s_cnt[i]);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < LS_CHUNK / 256; k++) {
            const uint32_t r = c0 + k * 256 + threadIdx.x;
            if (r < a.n) queue[s_base[bkt[k]] + slot[k]] = r;
        }
        __syncthreads();
    }
}

// ==================================================================================================
// Row marks: Ns (bits 0..30) and has-'N' (bit 31) of every row, for slabs whose producer did not write them.
// G lanes per row (16 bytes each per step); memory-bound.
// ==================================================================================================
template <int G>
__global__ void __launch_bounds__(256) count_marks_kernel(const FilterArgs a, uint32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31, sub = lane & (G - 1);
    const uint64_t groups = (uint64_t)gridDim.x * (256 / G);
    const uint64_t n_round = ((uint64_t)a.n + (32 / G) - 1) / (32 / G) * (32 / G);   // whole warps stay convergent
    for (uint64_t r = (uint64_t)blockIdx.x * (256 / G) + threadIdx.x / G; r < n_round; r += groups) {
        uint32_t ns = 0, up = 0;
        if (r < a.n) {
            const ReadGeom g = read_geom(a, (uint32_t)r);
            const uint8_t *row = a.slab + g.off;
            for (uint32_t pos = sub * 16; pos < g.eff; pos += G * 16) {
                const uint4 q4 = __ldg(reinterpret_cast<const uint4 *>(row + pos));
                uint32_t w[4] = {q4.x, q4.y, q4.z, q4.w};
                const int rem = (int)g.eff - (int)pos;
                if (rem < 16) {
#pragma unroll
                    for (int k = 0; k < 4; k++) w[k] = mask_word(w[k], rem - 4 * k);
                }
                count_marks4(w, ns, up);
            }
        }
        up = up ? 1u : 0u;
#pragma unroll
        for (int o = G / 2; o; o >>= 1) {
            ns += __shfl_xor_sync(FULL, ns, o);
            up |= __shfl_xor_sync(FULL, up, o);
        }
        if (sub == 0 && r < a.n) out[a.base + r] = ns | (up << 31);
    }
}

// ==================================================================================================
// 6-bit transport image -> slab: 12 image bytes (16 codes) -> 16 slab bytes per thread; flat over the
// whole chunk (rows keep their positions: image offset = slab offset * 3/4).  Memory-bound.
// ==================================================================================================
__device__ __forceinline__ uint32_t q6_expand4(uint32_t v24)
{
    // four 6-bit codes in bits 0..23 -> four bytes; 61..63 -> 0xFD..0xFF
    uint32_t b = (v24 & 0x3Fu) | ((v24 & 0xFC0u) << 2) | ((v24 & 0x3F000u) << 4) | ((v24 & 0xFC0000u) << 6);
    const uint32_t hi = ((b + 0x03030303u) & 0x40404040u) >> 6;      // 1 in every byte whose code is > 60
    return b + hi * 192u;
}

// one thread per 16 slab bytes (= 12 image bytes = 3 words): exact, never writes past the range
// marks != nullptr (rows at a uniform pitch of stride16 units, all `len` bases long): the thread that expands a unit with
// 'N' / 'n' codes inside the read adds them to its row's marks word (zeroed by the caller) -- the filter then counts nothing.
__global__ void __launch_bounds__(256) unpack_q6_kernel(const uint32_t *__restrict__ in, uint4 *__restrict__ out, uint64_t n16,
                                                        uint32_t *__restrict__ marks, uint32_t stride16, uint32_t len)
{
    for (uint64_t u = blockIdx.x * 256ull + threadIdx.x; u < n16; u += gridDim.x * 256ull) {
        const uint32_t w0 = __ldg(in + 3 * u), w1 = __ldg(in + 3 * u + 1), w2 = __ldg(in + 3 * u + 2);
        uint4 o;
        o.x = q6_expand4(w0 & 0xFFFFFFu);
        o.y = q6_expand4((w0 >> 24) | ((w1 & 0xFFFFu) << 8));
        o.z = q6_expand4((w1 >> 16) | ((w2 & 0xFFu) << 16));
        o.w = q6_expand4(w2 >> 8);
        out[u] = o;
        if (marks && ((o.x | o.y | o.z | o.w) & 0x80808080u)) {
            const uint64_t row = u / stride16;
            const int rem = (int)len - (int)((u - row * stride16) * 16);
            uint32_t w[4] = {o.x, o.y, o.z, o.w};
            if (rem < 16) {
#pragma unroll
                for (int k = 0; k < 4; k++) w[k] = mask_word(w[k], rem - 4 * k);
            }
            uint32_t ns = 0, up = 0;
            count_marks4(w, ns, up);
            if (ns) {
                atomicAdd(&marks[row], ns);
                if (up) atomicOr(&marks[row], 0x80000000u);
            }
        }
    }
}

// ==================================================================================================
// FP64 issue-rate probe: per thread 4 independent copies of the K=4 update (7 DMUL + 4 DADD).
// ==================================================================================================
__global__ void __launch_bounds__(512) fp64_peak_kernel(int iters, double *sink, double p)
{
    double P[4][4], e[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        e[c] = p * (1.0 + 0.125 * c);
#pragma unroll
        for (int j = 0; j < 4; j++) P[c][j] = 1.0 / (1.0 + threadIdx.x + c + j);
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            e[c] = __dmul_rn(e[c], 1.0000001);           // 1 DMUL (keeps q loop-variant and per-copy)
            const double q = __dsub_rn(1.0, e[c]);       // 1 DADD
#pragma unroll
            for (int j = 3; j >= 1; j--) P[c][j] = __dadd_rn(__dmul_rn(q, P[c][j]), __dmul_rn(e[c], P[c][j - 1]));
            P[c][0] = __dmul_rn(q, P[c][0]);             // 7 DMUL + 3 DADD
        }
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += P[c][j];
    if (s == 123.456) sink[0] = s;
}

const CUtensorMap g_no_tmap = {};

template <int K, int MODE>
int launch_tpr(const FilterArgs &a, const LaunchCfg &cfg)
{
    constexpr int W = tpr_warps(K);
    if (a.e_equals_p) tpr_kernel<K, MODE, true, false><<<cfg.sm_count, W * 32, TPR_SMEM, cfg.stream>>>(a, g_no_tmap);
    else tpr_kernel<K, MODE, false, false><<<cfg.sm_count, W * 32, TPR_SMEM, cfg.stream>>>(a, g_no_tmap);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// first-pass variant fed by TMA tensor tiles (uniform-stride slabs)
template <int K, int MODE>
int launch_tpr_tma(const FilterArgs &a, const LaunchCfg &cfg)
{
    constexpr int W = tpr_warps(K);
    const CUtensorMap &tm = *static_cast<const CUtensorMap *>(cfg.tmap);
    if (a.e_equals_p) tpr_kernel<K, MODE, true, true><<<cfg.sm_count, W * 32, TPR_SMEM, cfg.stream>>>(a, tm);
    else tpr_kernel<K, MODE, false, true><<<cfg.sm_count, W * 32, TPR_SMEM, cfg.stream>>>(a, tm);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

template <int K, int MODE, bool TMA>
int init_tpr()
{
    if (cudaFuncSetAttribute(tpr_kernel<K, MODE, true, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, TPR_SMEM) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(tpr_kernel<K, MODE, false, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, TPR_SMEM) != cudaSuccess) return -1;
    return 0;
}

}  // namespace

#define MOIRA_FOR_EACH_K(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(10) X(12) X(14) X(16) X(18) X(20) X(22) X(24) X(28) X(32)

int max_first_pass_k() { return 32; }

int kernels_init(int)
{
#define X(k) if (init_tpr<k, 0, false>() || init_tpr<k, 0, true>()) return -1;
    MOIRA_FOR_EACH_K(X)
#undef X
    if (init_tpr<1, 1, false>() || init_tpr<1, 1, true>() || init_tpr<2, 2, false>() || init_tpr<2, 2, true>()) return -1;
    if (cudaFuncSetAttribute(sorted_first_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TPR_SMEM) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(sorted_first_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TPR_SMEM) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(sorted_first_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TPR_SMEM) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(sorted_first_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TPR_SMEM) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(ladder_tpr_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TPR_SMEM) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(ladder_tpr_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TPR_SMEM) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(ladder_tpr_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TPR_SMEM) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(ladder_tpr_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TPR_SMEM) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(blk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BLK_SMEM) != cudaSuccess) return -1;
    return 0;
}

int launch_pb_first(const FilterArgs &a, int k_wanted, const LaunchCfg &cfg, const char **name)
{
    static const char *names[] = {"pb_tpr<K=2>", "pb_tpr<K=3>", "pb_tpr<K=4>", "pb_tpr<K=5>", "pb_tpr<K=6>",
                                  "pb_tpr<K=7>", "pb_tpr<K=8>", "pb_tpr<K=10>", "pb_tpr<K=12>", "pb_tpr<K=14>",
                                  "pb_tpr<K=16>", "pb_tpr<K=18>", "pb_tpr<K=20>", "pb_tpr<K=22>", "pb_tpr<K=24>",
                                  "pb_tpr<K=28>", "pb_tpr<K=32>"};
    int idx = 0;
#define X(k)                                                   \
    if (k_wanted <= k) {                                       \
        if (name) *name = names[idx];                          \
        return (cfg.tmap ? launch_tpr_tma<k, 0>(a, cfg) : launch_tpr<k, 0>(a, cfg)) ? -1 : k; \
    }                                                          \
    idx++;
    MOIRA_FOR_EACH_K(X)
#undef X
    if (name) *name = names[16];
    return (cfg.tmap ? launch_tpr_tma<32, 0>(a, cfg) : launch_tpr<32, 0>(a, cfg)) ? -1 : 32;
}

int launch_pb_first_k(const FilterArgs &a, int k_index, const LaunchCfg &cfg, const char **name)
{
    return launch_pb_first(a, first_pass_k(k_index), cfg, name) < 0 ? -1 : 0;
}

int launch_length_sort(const FilterArgs &a, const LenSortBufs &b, int single_group, const LaunchCfg &cfg)
{
    const int grid = cfg.sm_count * 4;
    len_hist_kernel<<<grid, 256, 0, cfg.stream>>>(a, b.hist);
    len_scan_kernel<<<1, 1024, 0, cfg.stream>>>(a, b.hist, b.bucket_start, b.cursor, b.group_start, b.group_count, single_group);
    len_scatter_kernel<<<grid, 256, 0, cfg.stream>>>(a, b.bucket_start, b.cursor, b.queue);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_unpack_q6(const uint8_t *d_image, uint8_t *d_slab, uint64_t slab_bytes, uint32_t *d_marks, uint64_t stride, uint32_t len,
                     const LaunchCfg &cfg)
{
    const uint64_t n48 = slab_bytes / 16;           // units of 16 slab bytes <- 12 image bytes
    if (!n48) return 0;
    const uint64_t want = (n48 + 255) / 256;
    const int grid = (int)std::min<uint64_t>(want, (uint64_t)cfg.sm_count * 16);
    unpack_q6_kernel<<<grid, 256, 0, cfg.stream>>>(reinterpret_cast<const uint32_t *>(d_image), reinterpret_cast<uint4 *>(d_slab), n48,
                                                   d_marks, (uint32_t)(stride / 16), len);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_count_marks(const FilterArgs &a, uint32_t *d_marks, uint32_t max_len, const LaunchCfg &cfg)
{
    if (!a.n) return 0;
    const int grid = cfg.sm_count * 8;
    if (max_len && max_len <= 128) count_marks_kernel<8><<<grid, 256, 0, cfg.stream>>>(a, d_marks);
    else if (max_len && max_len <= 256) count_marks_kernel<16><<<grid, 256, 0, cfg.stream>>>(a, d_marks);
    else count_marks_kernel<32><<<grid, 256, 0, cfg.stream>>>(a, d_marks);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_sorted_first(const FilterArgs &a, const uint32_t *seg_start, const uint32_t *seg_count, const LaunchCfg &cfg)
{
    if (a.e_equals_p) {
        sorted_first_kernel<true, false><<<cfg.sm_count, 512, TPR_SMEM, cfg.stream>>>(a, seg_start, seg_count);
        sorted_first_kernel<true, true><<<cfg.sm_count, 256, TPR_SMEM, cfg.stream>>>(a, seg_start, seg_count);
    } else {
        sorted_first_kernel<false, false><<<cfg.sm_count, 512, TPR_SMEM, cfg.stream>>>(a, seg_start, seg_count);
        sorted_first_kernel<false, true><<<cfg.sm_count, 256, TPR_SMEM, cfg.stream>>>(a, seg_start, seg_count);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_ladder_tpr(const FilterArgs &a, const LaunchCfg &cfg)
{
    if (a.e_equals_p) {
        ladder_tpr_kernel<true, false><<<cfg.sm_count, 512, TPR_SMEM, cfg.stream>>>(a);
        ladder_tpr_kernel<true, true><<<cfg.sm_count, 256, TPR_SMEM, cfg.stream>>>(a);
    } else {
        ladder_tpr_kernel<false, false><<<cfg.sm_count, 512, TPR_SMEM, cfg.stream>>>(a);
        ladder_tpr_kernel<false, true><<<cfg.sm_count, 256, TPR_SMEM, cfg.stream>>>(a);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_lambda(const FilterArgs &a, const LaunchCfg &cfg, const char **name)
{
    if (name) *name = "lambda_tpr";
    return cfg.tmap ? launch_tpr_tma<1, 1>(a, cfg) : launch_tpr<1, 1>(a, cfg);
}

// The classifier as the first pass (exact mode, reads that need many entries): over all reads of the sub-batch in slab order
// (TMA tiles when cfg.tmap is set) or over the queue / segment the caller put into `a`; every read goes to the rung its
// mean / variance estimate calls for.
int launch_classify_first(const FilterArgs &a, const LaunchCfg &cfg)
{
    return cfg.tmap ? launch_tpr_tma<2, 2>(a, cfg) : launch_tpr<2, 2>(a, cfg);
}

int launch_rung(const FilterArgs &a0, int b, const LaunchCfg &cfg0)
{
    FilterArgs a = a0;
    LaunchCfg cfg = cfg0;
    cfg.tmap = nullptr;
    a.rung = b;
    a.queue = a0.queues + (size_t)b * a0.queue_cap;
    a.queue_count = a0.queue_counts + b;
    const int wpr_grid = cfg.sm_count * 4;
    switch (b) {
    case 0: return launch_tpr<2, 2>(a, cfg);     // classifier
    case N_TPR_RUNGS + 1: wpr_kernel<4><<<wpr_grid, WPR_THREADS, 0, cfg.stream>>>(a); break;
    case N_TPR_RUNGS + 2: wpr_kernel<8><<<wpr_grid, WPR_THREADS, 0, cfg.stream>>>(a); break;
    case N_TPR_RUNGS + 3: wpr_kernel<16><<<wpr_grid, WPR_THREADS, 0, cfg.stream>>>(a); break;
    case N_TPR_RUNGS + 4: wpr_kernel<32><<<wpr_grid, WPR_THREADS, 0, cfg.stream>>>(a); break;
    case NB - 1: blk_kernel<<<cfg.sm_count, BLK_THREADS, BLK_SMEM, cfg.stream>>>(a); break;
    default: return -1;   // rungs 1..N_TPR_RUNGS run fused, see launch_ladder_tpr
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_policy(const uint32_t *queue_count, uint32_t max_pushed, uint32_t *policy, cudaStream_t s)
{
    policy_kernel<<<1, 1, 0, s>>>(queue_count, max_pushed, policy);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
int launch_policy_first(const uint32_t *queue_count, uint32_t max_pushed, uint32_t *policy, cudaStream_t s)
{
    policy_first_kernel<<<1, 1, 0, s>>>(queue_count, max_pushed, policy);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
int launch_policy_sorted(const uint32_t *queue_counts, uint32_t max_pushed, uint32_t *policy, cudaStream_t s)
{
    policy_sorted_kernel<<<1, 1, 0, s>>>(queue_counts, max_pushed, policy);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
int launch_policy_second(const uint32_t *jhist, int k_first, int exact, uint32_t *policy, cudaStream_t s)
{
    policy_second_kernel<<<1, 1, 0, s>>>(jhist, k_first, exact, policy);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_fp64_peak(int iters, int sm_count, double *d_sink, cudaStream_t s, double *ops_out)
{
    const int blocks = sm_count * 2, threads = 512;
    fp64_peak_kernel<<<blocks, threads, 0, s>>>(iters, d_sink, 1e-3);
    // per iteration and thread: 4 x (8 DMUL + 4 DADD) = 48 FP64 instructions (checked in the SASS)
    *ops_out = (double)blocks * threads * (double)iters * 48.0;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace moira
