/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
 *
 * CPU producer of BASELINE config 5 ("Cholesky-solve residual check"): a BANDED restatement of the
 * reference's f32 solver,
 *     solve                  /root/reference/src/lib.rs:11-24
 *     Csr::cholesky_decomp   /root/reference/src/sparse.rs:682-714
 *     forward_substitution   /root/reference/src/lib.rs:28-46
 *     backward_substitution  /root/reference/src/lib.rs:49-65
 * The reference as coded is O(n^4) (dense-row materialisation inside a triple loop) and cannot run at
 * 10^6 rows; for a matrix of half-bandwidth hb every term outside the band multiplies a structural
 * zero of L (`sum += 0.0*x` leaves a finite sum unchanged; zero results are never stored,
 * sparse.rs:229), so restricting the loops to the band is bit-equivalent for finite data.
 * With hb = n-1 nothing is skipped and the routines ARE the reference's loops: that is how they are
 * pinned on the reference's own f32 KATs (cholesky_decomposition_0/1 sparse.rs:1030-1080,
 * forward_/backward_substitution_test_0 and solve_test lib.rs:73-137) by tests/test_oracle_solve.py.
 *
 * Band storage (row-major, hb+1 slots per row): M[i][j] for i-hb <= j <= i lives at
 * band[i*(hb+1) + (j - i + hb)]; slots with j < 0 are unused.
 * Compile with -ffp-contract=off (rustc never contracts), link libm: powf is the same libm routine
 * Rust's f32::powf lowers to.
 */
#include <math.h>
#include <stddef.h>

#define BAND(m, i, j) (m)[(size_t)(i) * (hb + 1) + ((j) + hb - (i))]

/* sparse.rs:682-714. Returns 0, or 2 = MatErr::NonSquareMatrix is impossible here (band is square). */
int osolve_cholesky_band_f32(size_t n, size_t hb, const float *a_band, float *l_band)
{
    for (size_t i = 0; i < n; ++i) {                                   /* :688 */
        const size_t j0 = i > hb ? i - hb : 0;
        for (size_t j = j0; j <= i; ++j) {                             /* :689  (j < j0: L[i][j] is a structural zero) */
            float sum = 0.0f;                                          /* :690 */
            for (size_t k = j0; k < j; ++k) {                          /* :691  (k < j0: L[i][k] == 0) */
                float a = BAND(l_band, i, k);                          /* :692-696 */
                float b = BAND(l_band, j, k);                          /* :697-701  (k >= j-hb because j0 >= j-hb) */
                sum += a * b;                                          /* :702 */
            }
            float val;
            if (i == j) {
                val = powf(BAND(a_band, i, i) - sum, 0.5f);            /* :705 */
            } else {
                float temp = BAND(a_band, i, j) - sum;                 /* :707 */
                float a = BAND(l_band, j, j);                          /* :708 */
                val = (1.0f / a) * temp;                               /* :709  reciprocal, then multiply */
            }
            BAND(l_band, i, j) = val;                                  /* :711 (zeros would not be stored; they stay zeros here) */
        }
    }
    return 0;
}

/* lib.rs:28-46 — Ly = b for one right-hand side; only stored (non-zero) entries of L take part */
void osolve_forward_band_f32(size_t n, size_t hb, const float *l_band, const float *b, float *y)
{
    for (size_t r = 0; r < n; ++r) {                                   /* :33 */
        float l_x = 0.0f;                                              /* :35 */
        const size_t j0 = r > hb ? r - hb : 0;
        for (size_t j = j0; j < r; ++j) {                              /* :37-41 entries with col != row */
            float v = BAND(l_band, r, j);
            if (v != 0.0f) l_x += v * y[j];
        }
        y[r] = (b[r] - l_x) / BAND(l_band, r, r);                      /* :42  row.last() is the diagonal */
    }
}

/* lib.rs:49-65 — L* x = y; row r of L* holds L[j][r], j = r .. r+hb, diagonal first (skip(1)) */
void osolve_backward_band_f32(size_t n, size_t hb, const float *l_band, const float *y, float *x)
{
    for (size_t rr = n; rr-- > 0;) {                                   /* :54 rows reversed */
        float l_x = 0.0f;                                              /* :56 */
        const size_t j1 = rr + hb < n - 1 ? rr + hb : n - 1;
        for (size_t j = rr + 1; j <= j1; ++j) {                        /* :58-60 */
            float v = BAND(l_band, j, rr);
            if (v != 0.0f) l_x += v * x[j];
        }
        x[rr] = (y[rr] - l_x) / BAND(l_band, rr, rr);                  /* :61  row[0] is the diagonal */
    }
}

/* lib.rs:11-24 — solve(a, b): nrhs right-hand sides, column-major (b_cols[c*n + r]) */
int osolve_band_f32(size_t n, size_t hb, const float *a_band, size_t nrhs, const float *b_cols, float *x_cols, float *l_band,
                    float *y_scratch)
{
    int rc = osolve_cholesky_band_f32(n, hb, a_band, l_band);          /* :20 */
    if (rc) return rc;
    for (size_t c = 0; c < nrhs; ++c) {                                /* the per-column loops of :32 and :53 */
        osolve_forward_band_f32(n, hb, l_band, b_cols + c * n, y_scratch);   /* :22 */
        osolve_backward_band_f32(n, hb, l_band, y_scratch, x_cols + c * n);  /* :23 */
    }
    return 0;
}
