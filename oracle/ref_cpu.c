/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
 *
 * CPU restatement (plain C) of the reference's Csr x Dense hot path,
 * /root/reference/src/sparse.rs:426-446 (Csr::mul_dense) with the construction and row
 * access routines it calls (sparse.rs:116-132, 193-265) and the column-major Dense operand
 * (/root/reference/src/dense.rs:4-47).  See ref_cpu_impl.h for the line-by-line citations.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker / the reported CPU baseline.
 *
 * Parity status: PINNED for the integer instantiation by the reference's own KATs
 * (tests/golden/reference_kats.json, transcribed from sparse.rs tests by
 * tests/golden/make_golden.py); the f32/f64 instantiations are the same code with T swapped
 * (no floating-point golden vector exists in the reference for this path).
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define T int32_t
#define SFX i32
#include "ref_cpu_impl.h"
#undef T
#undef SFX

#define T float
#define SFX f32
#include "ref_cpu_impl.h"
#undef T
#undef SFX

#define T double
#define SFX f64
#include "ref_cpu_impl.h"
#undef T
#undef SFX

/* src/sparse.rs:468-482 — Csr::mul_vector (i32 instantiation, for the KAT at 1501-1529).
 * The reference transposes (296-318: columns ascending, stored order inside a column) and
 * then, per output row, sums v*rhs[col] over that row's entries of the transpose, i.e. over
 * the row's entries in ascending column order (ties in stored order), starting from 0.
 * Returns 3 = MatErr::IncorrectDimensions (469-471). */
int ocsr_mul_vector_i32(const ocsr_i32 *a, const int32_t *rhs, size_t rhs_len, int32_t *out,
                        size_t out_len)
{
    if (a->cols != rhs_len || a->rows != out_len) return 3;
    for (size_t r = 0; r < a->rows; ++r) {
        size_t s = a->row_index[r], e = a->row_index[r + 1];
        int32_t sum = 0;
        /* ascending-column walk, stable: for each column value in order pick matching entries */
        size_t done = 0, n = e - s;
        size_t last_col = 0;
        int have_last = 0;
        while (done < n) {
            /* find the smallest column > last_col (or >= 0 on the first pass) */
            size_t best = (size_t)-1;
            for (size_t k = s; k < e; ++k) {
                size_t c = a->col_index[k];
                if ((!have_last || c > last_col) && c < best) best = c;
            }
            for (size_t k = s; k < e; ++k)
                if (a->col_index[k] == best) {
                    sum = sum + a->v[k] * rhs[best];
                    ++done;
                }
            last_col = best;
            have_last = 1;
        }
        out[r] = sum;
    }
    return 0;
}

/* monotonic seconds, for the CPU-baseline timing legs of bench.py */
double oracle_now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Timed faithful mul_dense on the row range [row_begin,row_end) of `a` (f64): builds the
 * sub-result through the same get_row_compact(cols-capacity) + insert + finalise path and
 * returns elapsed seconds for the multiply only, like the reference bench's timed closure
 * (/root/reference/benches/sparse_dense_mul.rs:30-32). nnz of the sub-result via *nnz_out. */
#define DEFINE_TIMED(SFXX, TT)                                                                    \
    double ocsr_time_mul_dense_rows_##SFXX(const ocsr_##SFXX *a, const TT *const *rhs_cols,       \
                                           size_t rhs_row_count, size_t rhs_col_count,            \
                                           size_t row_begin, size_t row_end, int faithful,        \
                                           size_t *nnz_out)                                       \
    {                                                                                             \
        /* view of rows [row_begin,row_end) sharing a's arrays; row_index is absolute, which    \
         * get_row_compact handles because it indexes v/col_index absolutely */                  \
        ocsr_##SFXX view = *a;                                                                    \
        view.rows = row_end - row_begin;                                                          \
        view.row_index = a->row_index + row_begin;                                                \
        view.ri_len = view.rows + 1;                                                              \
        ocsr_##SFXX *res = NULL;                                                                  \
        double t0 = oracle_now();                                                                 \
        int rc = ocsr_mul_dense_##SFXX(&view, rhs_cols, rhs_row_count, rhs_col_count, faithful,   \
                                       &res);                                                     \
        double t1 = oracle_now();                                                                 \
        if (rc != 0) return -1.0;                                                                 \
        if (nnz_out) *nnz_out = ocsr_get_nnz_##SFXX(res);                                         \
        ocsr_free_##SFXX(res);                                                                    \
        return t1 - t0;                                                                           \
    }
DEFINE_TIMED(f64, double)
DEFINE_TIMED(f32, float)

/* NOT the reference: a "lean" multi-threaded variant for context next to the reference's number —
 * same arithmetic (stored order, multiply then add) and the same column-major gather, but rows are
 * split over POSIX threads, the product is written densely (row-major) and nothing is allocated per
 * row. Returns elapsed seconds. */
#include <pthread.h>
#define DEFINE_LEAN(SFXX, TT)                                                                          \
    typedef struct {                                                                                   \
        const ocsr_##SFXX *a;                                                                          \
        const TT *const *rhs_cols;                                                                     \
        size_t ncols, row_begin, r0, r1;                                                               \
        TT *out;                                                                                       \
    } lean_job_##SFXX;                                                                                 \
    static void *lean_worker_##SFXX(void *arg)                                                         \
    {                                                                                                  \
        const lean_job_##SFXX *j = (const lean_job_##SFXX *)arg;                                       \
        for (size_t r = j->r0; r < j->r1; ++r) {                                                       \
            const size_t s = j->a->row_index[r], e = j->a->row_index[r + 1];                           \
            for (size_t c = 0; c < j->ncols; ++c) {                                                    \
                TT value = (TT)0;                                                                      \
                const TT *col = j->rhs_cols[c];                                                        \
                for (size_t k = s; k < e; ++k) {                                                       \
                    TT prod = j->a->v[k] * col[j->a->col_index[k]];                                    \
                    value = value + prod;                                                              \
                }                                                                                      \
                j->out[(r - j->row_begin) * j->ncols + c] = value;                                     \
            }                                                                                          \
        }                                                                                              \
        return NULL;                                                                                   \
    }                                                                                                  \
    double ocsr_time_lean_parallel_##SFXX(const ocsr_##SFXX *a, const TT *const *rhs_cols, size_t rhs_col_count, \
                                          size_t row_begin, size_t row_end, TT *out_rowmajor, int threads)       \
    {                                                                                                  \
        if (threads < 1) threads = 1;                                                                  \
        if (threads > 256) threads = 256;                                                              \
        pthread_t tid[256];                                                                            \
        lean_job_##SFXX job[256];                                                                      \
        const size_t rows = row_end - row_begin;                                                       \
        double t0 = oracle_now();                                                                      \
        for (int t = 0; t < threads; ++t) {                                                            \
            job[t].a = a;                                                                              \
            job[t].rhs_cols = rhs_cols;                                                                \
            job[t].ncols = rhs_col_count;                                                              \
            job[t].row_begin = row_begin;                                                              \
            job[t].r0 = row_begin + rows * (size_t)t / (size_t)threads;                                \
            job[t].r1 = row_begin + rows * (size_t)(t + 1) / (size_t)threads;                          \
            job[t].out = out_rowmajor;                                                                 \
            pthread_create(&tid[t], NULL, lean_worker_##SFXX, &job[t]);                                \
        }                                                                                              \
        for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);                                  \
        return oracle_now() - t0;                                                                      \
    }
DEFINE_LEAN(f64, double)
DEFINE_LEAN(f32, float)
