/*
 * oracle/pb_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, libm only) of moira's Poisson-binomial expected-error
 * calculation, i.e. the algorithm of /root/reference/moira/bernoullimodule.c:131-263
 * (and its Python twin moira/moira.py:1561-1634).  It is the checker the CUDA path is
 * compared against in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 * The product (moira_b200/) never imports, links or calls anything in this directory.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against the reference's
 * own known-answer tests (moira/test/test_moira.py:39-45, 63-70, 127-128), against the
 * golden partitions in moira/test/test_results/ and -- bit for bit -- against outputs of
 * the unmodified reference binary (oracle/_ref, built by oracle/Makefile from the reference
 * sources where they lie) stored in tests/golden/ref_outputs.npz.
 *
 * Two variants are provided:
 *   oracle_pb_faithful  same loop nest and operation order as the reference (row-major
 *                       over the error count j, full i = 0..j convolution sum with the
 *                       per-term Bernoulli pmf recomputed through pow()), O(L' * j*^2),
 *                       rows on the heap instead of an (L+1) x L' stack VLA.
 *   oracle_pb           same cells, computed with the two surviving terms only and two
 *                       rolling rows, O(L' * j*).  Bitwise equal to the faithful variant
 *                       because every dropped term is (+0.0 * finite) added to a
 *                       non-negative sum (verified by tests/test_oracle.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* bernoullimodule.c:202 -- Phred score to error probability, via the host libm. */
ORACLE_API double oracle_error_prob(int q)
{
    return pow(10, (q / -10.0));
}

/* bernoullimodule.c:131-148 with n observations; the reference always passes n == 1
 * (bernoullimodule.c:187).  Kept general so the restatement is checkable term by term. */
ORACLE_API double oracle_binomial_pmf(double p, int j, int n)
{
    if (j > n) return 0.0;                                   /* :134-137 */
    double v = pow((1 - p), n);                              /* :140 */
    for (int i = 1; i <= j; i++)                             /* :142-145 */
        v = ((n - i + 1) / (1.0 * i)) * (p / (1 - p)) * v;
    return v;
}

/* bernoullimodule.c:170-178 (Python twin moira.py:1723-1733). */
ORACLE_API double oracle_interpolate(int e1, double p1, int e2, double p2, double alpha)
{
    double r = e1 + ((e2 - e1) * ((1 - alpha) - p1) / (p2 - p1));
    if (r < 0) r = 0;
    return r;
}

/* Compact the non-N error probabilities (bernoullimodule.c:191-204): 'N' (78) and 'n'
 * (110) are counted and skipped.  Returns L' and writes Ns. */
static int collect_probs(const char *contig, const int *quals, int len, double *probs, int *ns_out)
{
    int ns = 0;
    for (int i = 0; i < len; i++) {
        if (contig[i] == 78 || contig[i] == 110) ns++;
        else probs[i - ns] = oracle_error_prob(quals[i]);
    }
    *ns_out = ns;
    return len - ns;
}

/* Faithful restatement of test() (bernoullimodule.c:182-263) + sum_of_binomials (:152-166).
 * Returns 0 on success, -1 on allocation failure, -2 if the cumulative probability never
 * exceeds 1 - alpha within L'+1 terms (the reference would run off its arrays there). */
ORACLE_API int oracle_pb_faithful(const char *contig, const int *quals, int len, double alpha,
                                  double *ee_out, int *ns_out)
{
    const int n = 1;                                          /* :187 */
    double *probs = (double *)malloc(sizeof(double) * (len > 0 ? len : 1));
    if (!probs) return -1;
    int ns;
    int lp = collect_probs(contig, quals, len, probs, &ns);
    *ns_out = ns;
    if (lp <= 0) { *ee_out = 0; free(probs); return 0; }     /* :217, :257-260 */

    int max_rows = len + 1;                                   /* :208 */
    double **rows = (double **)calloc(max_rows, sizeof(double *));
    double *acc = (double *)malloc(sizeof(double) * max_rows);
    int rc = 0, j = 0;
    if (!rows || !acc) { rc = -1; goto done; }
    for (;;) {                                                /* :219 */
        if (j >= max_rows) { rc = -2; break; }
        rows[j] = (double *)malloc(sizeof(double) * lp);
        if (!rows[j]) { rc = -1; break; }
        for (int k = 0; k < lp; k++) {                        /* :221 */
            if (k == 0) {
                rows[j][k] = oracle_binomial_pmf(probs[k], j, n);          /* :225 */
            } else {
                double s = 0;                                 /* :155 */
                for (int i = 0; i <= j; i++)                  /* :158-163 */
                    s += oracle_binomial_pmf(probs[k], i, n) * rows[j - i][k - 1];
                rows[j][k] = s;
            }
        }
        double last = rows[j][lp - 1];                        /* :233 */
        acc[j] = (j == 0) ? last : acc[j - 1] + last;         /* :235-242 */
        if (acc[j] > (1 - alpha)) break;                      /* :244 */
        j++;
    }
    if (rc == 0) {
        /* :254.  For j == 0 the reference reads acc[-1] (undefined); the Python twin seeds the
         * list with 0 (moira.py:1611) and the interpolation then clamps to 0, which is also
         * what the compiled reference returns in practice.  Observable contract: ee = 0. */
        double prev = (j == 0) ? 0.0 : acc[j - 1];
        *ee_out = oracle_interpolate(j - 1, prev, j, acc[j], alpha);
    }
done:
    if (rows) { for (int r = 0; r < max_rows; r++) free(rows[r]); free(rows); }
    free(acc);
    free(probs);
    return rc;
}

/* Same cells with the structurally-zero terms removed:
 *   P[j][k] = fl( fl((1-p_k) * P[j][k-1]) + fl(e_k * P[j-1][k-1]) ),  P[0][k] = fl((1-p_k) * P[0][k-1])
 * with (1-p_k) = pmf(p_k, 0, 1) and e_k = pmf(p_k, 1, 1) evaluated exactly as
 * bernoullimodule.c:140,144 do.  Two rolling rows. */
ORACLE_API int oracle_pb(const char *contig, const int *quals, int len, double alpha,
                         double *ee_out, int *ns_out)
{
    double *buf = (double *)malloc(sizeof(double) * (4 * (size_t)(len > 0 ? len : 1)));
    if (!buf) return -1;
    double *probs = buf, *prev_row = buf + len, *cur_row = buf + 2 * (size_t)len, *e = buf + 3 * (size_t)len;
    int ns;
    int lp = collect_probs(contig, quals, len, probs, &ns);
    *ns_out = ns;
    if (lp <= 0) { *ee_out = 0; free(buf); return 0; }
    for (int k = 0; k < lp; k++) {
        e[k] = oracle_binomial_pmf(probs[k], 1, 1);
        probs[k] = oracle_binomial_pmf(probs[k], 0, 1);      /* now holds 1 - p_k */
    }
    double acc_prev = 0.0, acc = 0.0;
    int j = 0, rc = 0;
    for (;;) {
        if (j > len) { rc = -2; break; }
        if (j == 0) {
            cur_row[0] = probs[0];
            for (int k = 1; k < lp; k++) cur_row[k] = probs[k] * cur_row[k - 1];
        } else {
            cur_row[0] = (j == 1) ? e[0] : 0.0;
            for (int k = 1; k < lp; k++)
                cur_row[k] = probs[k] * cur_row[k - 1] + e[k] * prev_row[k - 1];
        }
        acc_prev = acc;
        acc = (j == 0) ? cur_row[lp - 1] : acc_prev + cur_row[lp - 1];
        if (acc > (1 - alpha)) break;
        double *t = prev_row; prev_row = cur_row; cur_row = t;
        j++;
    }
    if (rc == 0) *ee_out = oracle_interpolate(j - 1, (j == 0) ? 0.0 : acc_prev, j, acc, alpha);
    free(buf);
    return rc;
}

/* Batch driver over an in-band packed slab (same encoding as include/moira_b200.h):
 * byte < 0xFD = Phred score (0 means 1, bernoullimodule.c:104-107), 0xFF = 'N', 0xFE = 'n'. */
ORACLE_API int oracle_pb_batch(const uint8_t *slab, const uint64_t *offsets, const uint32_t *lengths,
                               uint64_t n_reads, double alpha, int faithful, double *ee, int32_t *ns)
{
    uint32_t cap = 0;
    char *contig = NULL;
    int *quals = NULL;
    int rc = 0;
    for (uint64_t r = 0; r < n_reads && rc == 0; r++) {
        uint32_t len = lengths[r];
        if (len + 1 > cap) {
            cap = len + 1;
            contig = (char *)realloc(contig, cap);
            quals = (int *)realloc(quals, sizeof(int) * cap);
            if (!contig || !quals) { rc = -1; break; }
        }
        const uint8_t *row = slab + offsets[r];
        for (uint32_t i = 0; i < len; i++) {
            uint8_t b = row[i];
            if (b == 0xFF) { contig[i] = 'N'; quals[i] = 2; }
            else if (b == 0xFE) { contig[i] = 'n'; quals[i] = 2; }
            else { contig[i] = 'A'; quals[i] = b == 0 ? 1 : b; }
        }
        contig[len] = 0;
        int nsv = 0;
        double eev = 0;
        rc = faithful ? oracle_pb_faithful(contig, quals, (int)len, alpha, &eev, &nsv)
                      : oracle_pb(contig, quals, (int)len, alpha, &eev, &nsv);
        ee[r] = eev;
        ns[r] = nsv;
    }
    free(contig);
    free(quals);
    return rc;
}

/* Host-libm lookup tables the product must reproduce: q1mp[Q] = pmf(p,0,1), e[Q] = pmf(p,1,1),
 * p[Q] = 10^(-Q/10) for Q = 0..255 (Q = 0 is remapped to 1 by the caller, not here). */
ORACLE_API void oracle_tables(double p[256], double q1mp[256], double e[256])
{
    for (int q = 0; q < 256; q++) {
        p[q] = oracle_error_prob(q);
        q1mp[q] = oracle_binomial_pmf(p[q], 0, 1);
        e[q] = oracle_binomial_pmf(p[q], 1, 1);
    }
}
