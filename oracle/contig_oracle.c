/*
 * oracle/contig_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of moira's paired-end contig constructor:
 *   reverse_complement   moira/moira.py:1207-1236
 *   nw_align             moira/nw_align.pyx:49-145   (Python twin moira.py:1239-1331)
 *   nw_overlap           moira/nw_align.pyx:148-202  (Python twin moira.py:1334-1372)
 *   make_contig          moira/moira.py:1375-1558
 * as process_data chains them (moira.py:791-803).  It is the checker the CUDA contig kernels are
 * compared against in tests/; the product (moira_b200/) never imports, links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks it against the reference's own known-answer
 * vectors (testRC2, test_aligned incl. the path score 13431, test_contig, test_PairedProcess:
 * moira/test/test_moira.py:49-59, 67-70, 118-128), against the 400 golden contigs of the paired
 * full-pipeline test (moira/test/test_results/paired.qc.*), and -- string for string -- against the
 * UNMODIFIED reference aligner (oracle/_ref/nw_align*.so, cythonised from nw_align.pyx where it lies
 * by oracle/Makefile) on seeded random pairs.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

enum { CONSENSUS_BEST = 0, CONSENSUS_SUM = 1, CONSENSUS_POSTERIOR = 2 };

/* moira.py:1210-1213 -- the IUPAC complement table; anything else is a ValueError (moira.py:1228-1229). */
static int complement_of(char b)
{
    switch (b) {
    case 'A': return 'T'; case 'C': return 'G'; case 'T': return 'A'; case 'G': return 'C'; case 'N': return 'N';
    case 'W': return 'W'; case 'S': return 'S'; case 'R': return 'Y'; case 'Y': return 'R'; case 'M': return 'K';
    case 'K': return 'M'; case 'B': return 'V'; case 'V': return 'B'; case 'D': return 'H'; case 'H': return 'D';
    case '-': return '-'; case '.': return '.';
    default: return -1;
    }
}

/* returns 0, or 1 + index of the first unrecognised base */
ORACLE_API int oracle_reverse_complement(const char *seq, const int32_t *quals, int n, char *out_seq, int32_t *out_quals)
{
    for (int k = 0; k < n; k++) {                       /* sequence[::-1], then complement (moira.py:1223-1229) */
        const int c = complement_of(seq[n - 1 - k]);
        if (c < 0) return 1 + (n - 1 - k);
        out_seq[k] = (char)c;
        if (quals) out_quals[k] = quals[n - 1 - k];     /* quals.reverse() (moira.py:1232) */
    }
    return 0;
}

/* Global alignment with mothur's overlap refinement (refine_overlap = True, the only way process_data
 * calls it, moira.py:794-798).  a1 / a2 need room for L1 + L2 characters.  *score is the sum of the
 * matrix cells along the traceback path (nw_align.pyx:127), which is what the reference returns. */
ORACLE_API int oracle_nw_align(const char *s1, int L1, const char *s2, int L2, int match, int mismatch, int gap,
                               char *a1, char *a2, int *alen, long *score_out)
{
    const int R = L1 + 1, C = L2 + 1;                   /* the leading ' ' of both strings (nw_align.pyx:60-61) */
    int *H = (int *)malloc((size_t)R * C * sizeof(int));
    signed char *rd = (signed char *)malloc((size_t)R * C), *cd = (signed char *)malloc((size_t)R * C);
    if (!H || !rd || !cd) { free(H); free(rd); free(cd); return -1; }
#define AT(i, j) ((size_t)(i) * C + (j))
    for (int i = 0; i < R; i++) { H[AT(i, 0)] = 0; rd[AT(i, 0)] = -1; cd[AT(i, 0)] = 0; }     /* nw_align.pyx:70-76 */
    for (int j = 1; j < C; j++) { H[AT(0, j)] = 0; rd[AT(0, j)] = 0; cd[AT(0, j)] = -1; }     /* nw_align.pyx:77-83 */
    for (int i = 1; i < R; i++)
        for (int j = 1; j < C; j++) {                                                          /* nw_align.pyx:87-116 */
            const int diag = H[AT(i - 1, j - 1)] + (s1[i - 1] == s2[j - 1] ? match : mismatch);
            const int up = H[AT(i - 1, j)] + gap, left = H[AT(i, j - 1)] + gap;
            if (diag >= up) {
                if (diag >= left) { H[AT(i, j)] = diag; rd[AT(i, j)] = -1; cd[AT(i, j)] = -1; }
                else { H[AT(i, j)] = left; rd[AT(i, j)] = 0; cd[AT(i, j)] = -1; }
            } else {
                if (up >= left) { H[AT(i, j)] = up; rd[AT(i, j)] = -1; cd[AT(i, j)] = 0; }
                else { H[AT(i, j)] = left; rd[AT(i, j)] = 0; cd[AT(i, j)] = -1; }
            }
        }
    /* nw_overlap (nw_align.pyx:163-202): best cell of the last column and of the last row (">=": the last
     * of equal scores wins); unless both are the corner, the line with the larger best score -- ties go to
     * the row -- gets its cells beyond the best one redirected along the line. */
    int bc = -10000, bci = 0, br = -10000, bri = 0;
    for (int i = 0; i < R; i++) if (H[AT(i, C - 1)] >= bc) { bc = H[AT(i, C - 1)]; bci = i; }
    for (int j = 0; j < C; j++) if (H[AT(R - 1, j)] >= br) { br = H[AT(R - 1, j)]; bri = j; }
    if (bci == R - 1 && bri == C - 1) {
    } else if (bc > br) {
        for (int i = R - 1; i > bci; i--) { rd[AT(i, C - 1)] = -1; cd[AT(i, C - 1)] = 0; }
    } else {
        for (int j = C - 1; j > bri; j--) { rd[AT(R - 1, j)] = 0; cd[AT(R - 1, j)] = -1; }
    }
    /* traceback (nw_align.pyx:122-143), filled backwards then reversed */
    long score = 0;
    int n = 0, i = R - 1, j = C - 1;
    while (i > 0 || j > 0) {
        const int r = rd[AT(i, j)], c = cd[AT(i, j)];
        score += H[AT(i, j)];
        a1[n] = r == -1 ? s1[i - 1] : '-';
        a2[n] = c == -1 ? s2[j - 1] : '-';
        n++;
        i += r;
        j += c;
    }
    for (int k = 0; k < n / 2; k++) {
        char t = a1[k]; a1[k] = a1[n - 1 - k]; a1[n - 1 - k] = t;
        t = a2[k]; a2[k] = a2[n - 1 - k]; a2[n - 1 - k] = t;
    }
    *alen = n;
    if (score_out) *score_out = score;
#undef AT
    free(H); free(rd); free(cd);
    return 0;
}

static double qual2prob(int q) { return pow(10, (q / (-10.0))); }            /* moira.py:1391-1392 */
/* moira.py:1394-1395; log10 of a non-positive number raises in Python: reported as failure */
static int prob2qual(double prob, int *ok)
{
    if (!(prob > 0.0)) { *ok = 0; return 0; }
    return (int)floor(-10 * log10(prob));
}

/* make_contig (moira.py:1375-1558).  a1 / a2: the two aligned strings (alen characters each), q1 / q2: the
 * unaligned qualities.  Returns 0; -2 length mismatch (moira.py:1407-1410); -3 bad parameter (1411-1416);
 * -4 math domain error in the posterior formulas. */
ORACLE_API int oracle_make_contig(const char *a1, const int32_t *q1, int nq1, const char *a2, const int32_t *q2, int nq2,
                                  int alen, int insert, int deltaq, int consensus, int qscore_cap, int trim_overlap,
                                  char *contig, int32_t *cq, int *clen, int *overlap_length, int *gaps_out, int *mismatches_out)
{
    int n1 = 0, n2 = 0;
    for (int p = 0; p < alen; p++) { n1 += a1[p] != '-'; n2 += a2[p] != '-'; }
    if (n1 != nq1 || n2 != nq2) return -2;
    if (insert <= 0 || deltaq <= 0 || qscore_cap < 0 || consensus < 0 || consensus > 2) return -3;
    /* qualities fitted into the alignment (moira.py:1420-1437) */
    int32_t *fq = (int32_t *)malloc(sizeof(int32_t) * (alen + 1)), *rq = (int32_t *)malloc(sizeof(int32_t) * (alen + 1));
    for (int p = 0, u = 0; p < alen; p++) fq[p] = a1[p] == '-' ? 0 : q1[u++];
    for (int p = 0, u = 0; p < alen; p++) rq[p] = a2[p] == '-' ? 0 : q2[u++];
    /* first / last non-gap position of both strings (moira.py:1441-1458) */
    int fs = 0, rs = 0, fe = alen - 1, re = alen - 1;
    while (fs < alen && a1[fs] == '-') fs++;
    while (rs < alen && a2[rs] == '-') rs++;
    while (fe >= 0 && a1[fe] == '-') fe--;
    while (re >= 0 && a2[re] == '-') re--;
    int ostart, oend, reversed;                                                  /* moira.py:1461-1468 */
    if (fs < rs) { ostart = rs; oend = fe; reversed = 0; }
    else { ostart = fs; oend = re; reversed = 1; }
    *overlap_length = oend - ostart;                                             /* moira.py:1470 */
    int n = 0, gaps = 0, mism = 0, ok = 1;
#define EMIT(b, q) do { contig[n] = (b); cq[n] = (q); n++; } while (0)
    for (int p = 0; p < alen; p++) {
        const char f = a1[p], r = a2[p];
        if (p < ostart) {                                                        /* moira.py:1478-1485 */
            if (!trim_overlap) { if (reversed) EMIT(r, rq[p]); else EMIT(f, fq[p]); }
        } else if (p > oend) {                                                   /* moira.py:1486-1493 */
            if (!trim_overlap) { if (reversed) EMIT(f, fq[p]); else EMIT(r, rq[p]); }
        } else if (f == '-') {                                                   /* moira.py:1495-1504 */
            gaps++;
            if (consensus == CONSENSUS_POSTERIOR) EMIT('N', 2);
            else if (rq[p] > insert) EMIT(r, rq[p]);
        } else if (r == '-') {                                                   /* moira.py:1506-1515 */
            gaps++;
            if (consensus == CONSENSUS_POSTERIOR) EMIT('N', 2);
            else if (fq[p] > insert) EMIT(f, fq[p]);
        } else if (f == r) {                                                     /* moira.py:1517-1528 */
            if (consensus == CONSENSUS_SUM) EMIT(f, fq[p] + rq[p]);
            else if (consensus == CONSENSUS_POSTERIOR) {
                const double p1 = qual2prob(fq[p]), p2 = qual2prob(rq[p]);
                const double post = (p1 * p2 / 3) / (1 - p1 - p2 + (4 * p1 * p2 / 3));
                EMIT(f, prob2qual(post, &ok));
            } else EMIT(f, fq[p] >= rq[p] ? fq[p] : rq[p]);
        } else {                                                                 /* moira.py:1530-1554 */
            mism++;
            if (consensus != CONSENSUS_POSTERIOR) {
                if (abs(fq[p] - rq[p]) < deltaq) EMIT('N', 2);
                else if (fq[p] >= rq[p]) EMIT(f, fq[p]);
                else EMIT(r, rq[p]);
            } else if (fq[p] == rq[p]) EMIT('N', 2);
            else {
                double p1, p2;
                char b;
                if (fq[p] > rq[p]) { p1 = qual2prob(fq[p]); p2 = qual2prob(rq[p]); b = f; }
                else { p2 = qual2prob(fq[p]); p1 = qual2prob(rq[p]); b = r; }
                const double post = p1 * (1 - p2 / 3) / (p1 + p2 - (4 * p1 * p2 / 3));
                EMIT(b, prob2qual(post, &ok));
            }
        }
    }
#undef EMIT
    if (qscore_cap)                                                              /* moira.py:1555-1556 */
        for (int k = 0; k < n; k++) if (!(cq[k] < qscore_cap)) cq[k] = qscore_cap;
    *clen = n; *gaps_out = gaps; *mismatches_out = mism;
    free(fq); free(rq);
    return ok ? 0 : -4;
}

/* The paired branch of process_data (moira.py:791-803) for one pair: reverse-complement the reverse
 * read, align, build the contig.  contig / cq need room for L1 + L2 entries.
 * Returns 0, -1 (non-IUPAC base in the reverse read), or oracle_make_contig's codes. */
ORACLE_API int oracle_pair_to_contig(const char *fwd, const int32_t *fq, int L1, const char *rev, const int32_t *rq, int L2,
                                     int match, int mismatch, int gap, int insert, int deltaq, int consensus, int qscore_cap,
                                     int trim_overlap, char *contig, int32_t *cq, int *clen, int *overlap, int *gaps, int *mism)
{
    char *rc = (char *)malloc(L2 + 1), *a1 = (char *)malloc(L1 + L2 + 1), *a2 = (char *)malloc(L1 + L2 + 1);
    int32_t *rcq = (int32_t *)malloc(sizeof(int32_t) * (L2 + 1));
    int rcode = 0, alen = 0;
    if (oracle_reverse_complement(rev, rq, L2, rc, rcq)) rcode = -1;
    if (!rcode) rcode = oracle_nw_align(fwd, L1, rc, L2, match, mismatch, gap, a1, a2, &alen, NULL);
    if (!rcode) rcode = oracle_make_contig(a1, fq, L1, a2, rcq, L2, alen, insert, deltaq, consensus, qscore_cap, trim_overlap,
                                           contig, cq, clen, overlap, gaps, mism);
    free(rc); free(a1); free(a2); free(rcq);
    return rcode;
}
