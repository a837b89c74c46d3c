/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles the UNMODIFIED reference translation unit moira/bernoullimodule.c where it
 * lies under /root/reference (passed as -DREF_SOURCE="...") into oracle/_ref/bernoulli.so.
 * No reference source is copied into this repository: the file is #include'd by path at
 * build time and only the resulting binary (git-ignored) is kept.
 *
 * The reference uses two Python-2-only C-API names (PyInt_AsLong, bernoullimodule.c:97;
 * Py_InitModule / initbernoulli, bernoullimodule.c:122-125).  They are mapped to their
 * Python-3 equivalents by macro, and a PyInit_bernoulli is appended that registers the
 * reference's own method table (bernoullimodule.c:117-120), so the reference's
 * calculate_errors_PB binding (bernoullimodule.c:66-114) runs untouched under python 3.12.
 *
 * Also exported (plain C, for bulk timing without Python boxing): ref_pb_batch(), which
 * loops the reference's own test() (bernoullimodule.c:182-263) over an in-band packed slab
 * on worker threads whose stacks are large enough for the (L+1) x L' VLA at
 * bernoullimodule.c:214 (18 MB at 1500 bp; the default 8 MB stack segfaults).
 */
#include <Python.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PyInt_AsLong PyLong_AsLong
#define Py_InitModule(name, methods) ((void)0)
#undef PyMODINIT_FUNC
#define PyMODINIT_FUNC void

#ifndef REF_SOURCE
#error "build with -DREF_SOURCE=\"/root/reference/moira/bernoullimodule.c\""
#endif
#include REF_SOURCE

static struct PyModuleDef ref_module_def = {
    PyModuleDef_HEAD_INIT, "bernoulli", module_docstring, -1, module_methods,
    NULL, NULL, NULL, NULL};

__attribute__((visibility("default"))) PyObject *PyInit_bernoulli(void)
{
    return PyModule_Create(&ref_module_def);
}

/* ---- bulk driver over the reference's test() ------------------------------------------- */

/* In-band slab encoding shared with include/moira_b200.h: byte < 0xFD is a Phred score,
 * 0xFF is an 'N' base, 0xFE an 'n' base (both skipped and counted, bernoullimodule.c:196). */
typedef struct {
    const uint8_t *slab;
    const uint64_t *offsets;
    const uint32_t *lengths;
    uint64_t begin, end;
    double alpha;
    double *ee;
    int32_t *ns;
} ref_job;

static void *ref_worker(void *arg)
{
    ref_job *job = (ref_job *)arg;
    uint32_t cap = 0;
    char *contig = NULL;
    int *quals = NULL;
    for (uint64_t r = job->begin; r < job->end; r++) {
        uint32_t len = job->lengths[r];
        if (len + 1 > cap) {
            cap = len + 1;
            contig = (char *)realloc(contig, cap);
            quals = (int *)realloc(quals, cap * sizeof(int));
        }
        const uint8_t *row = job->slab + job->offsets[r];
        for (uint32_t i = 0; i < len; i++) {
            uint8_t b = row[i];
            if (b == 0xFF) { contig[i] = 'N'; quals[i] = 2; }
            else if (b == 0xFE) { contig[i] = 'n'; quals[i] = 2; }
            else { contig[i] = 'A'; quals[i] = b == 0 ? 1 : b; } /* Q==0 -> 1, bernoullimodule.c:104-107 */
        }
        contig[len] = '\0';
        struct tuple res = test(contig, quals, job->alpha);
        job->ee[r] = res.expected_errors;
        job->ns[r] = res.Ns;
    }
    free(contig);
    free(quals);
    return NULL;
}

__attribute__((visibility("default"))) int ref_pb_batch(
    const uint8_t *slab, const uint64_t *offsets, const uint32_t *lengths, uint64_t n_reads,
    double alpha, double *ee, int32_t *ns, int n_threads, uint64_t stack_bytes)
{
    if (n_threads < 1) n_threads = 1;
    if (stack_bytes < (64u << 20)) stack_bytes = 64u << 20;
    pthread_t *tids = (pthread_t *)calloc(n_threads, sizeof(pthread_t));
    ref_job *jobs = (ref_job *)calloc(n_threads, sizeof(ref_job));
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    if (pthread_attr_setstacksize(&attr, stack_bytes) != 0) return -1;
    int started = 0, rc = 0;
    for (int t = 0; t < n_threads; t++) {
        jobs[t].slab = slab; jobs[t].offsets = offsets; jobs[t].lengths = lengths;
        jobs[t].begin = n_reads * (uint64_t)t / n_threads;
        jobs[t].end = n_reads * (uint64_t)(t + 1) / n_threads;
        jobs[t].alpha = alpha; jobs[t].ee = ee; jobs[t].ns = ns;
        if (pthread_create(&tids[t], &attr, ref_worker, &jobs[t]) != 0) { rc = -2; break; }
        started++;
    }
    for (int t = 0; t < started; t++) pthread_join(tids[t], NULL);
    pthread_attr_destroy(&attr);
    free(tids);
    free(jobs);
    return rc;
}
