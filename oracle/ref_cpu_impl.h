/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
 *
 * Type-generic body of the CPU restatement of the reference's Csr x Dense hot path.
 * Included once per element type by ref_cpu.c with
 *     #define T      element type
 *     #define SFX    symbol suffix (i32 / f32 / f64)
 *
 * Every function cites the reference lines (under /root/reference/) it restates.
 * The reference is Rust and cannot be compiled in this image (no rustc/cargo), so this
 * is a "port" oracle. It is pinned against the reference's own integer KATs
 * (src/sparse.rs:1082-1109 test_dense_mul, 1153-1178 test_nnz, 815-868 and 1111-1151
 * structure KATs, 1501-1529 test_mul_vector) by tests/test_oracle_kats.py.
 * No floating-point KAT of mul_dense exists in the reference; f32/f64 parity means
 * agreement with this restated sequential sum (same code path as the pinned i32 one).
 *
 * Compile with -ffp-contract=off: Rust never contracts a*b+c into an FMA.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)
#define OCSR FN(ocsr)
#define OENTRY FN(ocsr_entry)

/* src/sparse.rs:68-78 — struct Csr<T>. usize == size_t (8 bytes on the target). */
typedef struct OCSR {
    size_t rows, cols;      /* dims: MatDim                       sparse.rs:70 */
    T *v;                   /* v: Vec<T>                          sparse.rs:71 */
    size_t *col_index;      /* col_index: Vec<usize>              sparse.rs:72 */
    size_t *row_index;      /* row_index: Vec<usize>              sparse.rs:73 */
    size_t v_len, v_cap;    /* v and col_index always have equal length */
    size_t ri_len, ri_cap;
    int is_finalised;       /*                                    sparse.rs:74 */
} OCSR;

/* src/sparse.rs:80-85 — CsrEntry<&T>: {v:&T, col_index:usize, row_index:usize} = 24 bytes */
typedef struct OENTRY {
    const T *v;
    size_t col_index;
    size_t row_index;
} OENTRY;

static void FN(push_v)(OCSR *m, T value, size_t col)
{
    if (m->v_len == m->v_cap) { /* Vec amortised doubling; capacity 0 start (sparse.rs:117-119) */
        size_t nc = m->v_cap ? m->v_cap * 2 : 4;
        m->v = (T *)realloc(m->v, nc * sizeof(T));
        m->col_index = (size_t *)realloc(m->col_index, nc * sizeof(size_t));
        m->v_cap = nc;
    }
    m->v[m->v_len] = value;
    m->col_index[m->v_len] = col;
    m->v_len++;
}

static void FN(push_ri)(OCSR *m, size_t x)
{
    if (m->ri_len == m->ri_cap) {
        size_t nc = m->ri_cap ? m->ri_cap * 2 : 4;
        m->row_index = (size_t *)realloc(m->row_index, nc * sizeof(size_t));
        m->ri_cap = nc;
    }
    m->row_index[m->ri_len++] = x;
}

/* src/sparse.rs:117-132 — Csr::new / new_with_capacity: empty v/col_index, row_index = [0] */
OCSR *FN(ocsr_new)(size_t rows, size_t cols, size_t capacity)
{
    OCSR *m = (OCSR *)calloc(1, sizeof(OCSR));
    m->rows = rows;
    m->cols = cols;
    if (capacity) {
        m->v = (T *)malloc(capacity * sizeof(T));
        m->col_index = (size_t *)malloc(capacity * sizeof(size_t));
        m->v_cap = capacity;
    }
    FN(push_ri)(m, 0);
    return m;
}

void FN(ocsr_free)(OCSR *m)
{
    if (!m) return;
    free(m->v);
    free(m->col_index);
    free(m->row_index);
    free(m);
}

/* src/sparse.rs:237-250 — insert_unchecked: append; grow row_index when `row` advances.
 * Out-of-order rows are NOT re-sorted: an entry whose row is <= the current last row is
 * simply appended to the current last row (the reference bench relies on this). */
static void FN(insert_unchecked)(OCSR *m, T value, size_t row, size_t col)
{
    FN(push_v)(m, value, col);                       /* :238-239 */
    if (row > m->ri_len - 1) {                       /* :240 */
        if (row > m->ri_len) {                       /* :241 */
            FN(push_ri)(m, m->v_len - 1);            /* :242 */
            /* :243-245  for _ in row_index.len()..(row+1) { push(*last) }  — the range is
             * evaluated once, before the pushes. */
            size_t from = m->ri_len, to = row + 1;
            for (size_t i = from; i < to; ++i) FN(push_ri)(m, m->row_index[m->ri_len - 1]);
        } else {
            FN(push_ri)(m, m->v_len - 1);            /* :247 */
        }
    }
}

/* src/sparse.rs:222-233 — insert: Err(MatrixFinalised) when finalised (returns 1);
 * values equal to T::default() are silently skipped (-0.0 == 0.0 is skipped, NaN is kept). */
int FN(ocsr_insert)(OCSR *m, T value, size_t row, size_t col)
{
    if (m->is_finalised) return 1;                   /* MatErr::MatrixFinalised  :223-225 */
    if (value != (T)0) FN(insert_unchecked)(m, value, row, col); /* :229-231 */
    return 0;
}

/* src/sparse.rs:206-219 — finalise: pad row_index with nnz up to rows+1 entries.
 * Returns 2 where the reference panics ("big eek", :210-212). */
int FN(ocsr_finalise)(OCSR *m)
{
    if (!m->is_finalised) {
        m->is_finalised = 1;
        if (m->rows < m->ri_len) return 2;           /* panic!("big eek") */
        size_t required_spacers = m->rows - m->ri_len;
        for (size_t i = 0; i < required_spacers; ++i) FN(push_ri)(m, m->v_len);
        FN(push_ri)(m, m->v_len);
    }
    return 0;
}

/* src/sparse.rs:162-164 — get_nnz = *row_index.last() */
size_t FN(ocsr_get_nnz)(const OCSR *m) { return m->ri_len ? m->row_index[m->ri_len - 1] : 0; }

/* src/sparse.rs:193-203 — from_data: data[r] is ROW r; every cell goes through insert */
OCSR *FN(ocsr_from_data)(const T *row_major, size_t rows, size_t cols)
{
    OCSR *m = FN(ocsr_new)(rows, cols, 0);
    for (size_t i = 0; i < rows; ++i)
        for (size_t j = 0; j < cols; ++j) FN(ocsr_insert)(m, row_major[i * cols + j], i, j);
    FN(ocsr_finalise)(m);
    return m;
}

/* src/sparse.rs:252-265 — get_row_compact. `faithful` != 0 reproduces the per-call
 * Vec::with_capacity(self.dims.cols) (24 bytes per slot, :254); otherwise the buffer is sized
 * to the row (same contents, used by the "lean" timing variant).
 * Returns the number of entries; *out must be free()d by the caller. */
static size_t FN(get_row_compact)(const OCSR *m, size_t index, int faithful, OENTRY **out)
{
    size_t row_start = m->row_index[index];                         /* :255 */
    size_t row_end = (index == m->ri_len - 1) ? m->v_len            /* :256-257 */
                                              : m->row_index[index + 1]; /* :259 */
    /* Vec::with_capacity(self.dims.cols) (:254); a Vec grows when a row holds more entries than
     * that (duplicate columns), modelled by sizing the buffer to the larger of the two */
    size_t cap = faithful ? m->cols : (row_end - row_start);
    if (cap < row_end - row_start) cap = row_end - row_start;
    OENTRY *row = (OENTRY *)malloc((cap ? cap : 1) * sizeof(OENTRY)); /* :254 */
    size_t n = 0;
    for (size_t e = row_start; e < row_end; ++e) {                  /* :261-263 */
        row[n].v = &m->v[e];
        row[n].col_index = m->col_index[e];
        row[n].row_index = index;
        ++n;
    }
    *out = row;
    return n;
}

/* exported wrapper for tests: copies (value, col_index) of one row into caller buffers */
size_t FN(ocsr_get_row_compact)(const OCSR *m, size_t index, T *vals, size_t *cols, size_t cap)
{
    OENTRY *row;
    size_t n = FN(get_row_compact)(m, index, 0, &row);
    for (size_t i = 0; i < n && i < cap; ++i) {
        vals[i] = *row[i].v;
        cols[i] = row[i].col_index;
    }
    free(row);
    return n;
}

/* src/sparse.rs:426-446 — Csr::mul_dense, the hot path.
 * rhs is the reference's Dense<T>: column-major, rhs_cols[c] points at column c
 * (src/dense.rs:5-9,31-33), rhs_row_count rows, rhs_col_count columns.
 * Returns 0 and a new finalised result Csr in *out, or 3 = MatErr::IncorrectDimensions.
 * Loop order row -> output column -> stored entry; value starts at T::default();
 * c = a*b then value = value + c (two roundings); every output goes through insert
 * (zero-drop); finalise at the end. */
int FN(ocsr_mul_dense)(const OCSR *a, const T *const *rhs_cols, size_t rhs_row_count,
                       size_t rhs_col_count, int faithful, OCSR **out)
{
    if (a->cols != rhs_row_count) return 3;                          /* :427-429 */
    OCSR *result = FN(ocsr_new)(a->rows, rhs_col_count, 0);          /* :430 */
    for (size_t row_index = 0; row_index < a->rows; ++row_index) {   /* :431 */
        OENTRY *row;
        size_t row_len = FN(get_row_compact)(a, row_index, faithful, &row); /* :432 */
        for (size_t col_index = 0; col_index < rhs_col_count; ++col_index) { /* :433 */
            T value = (T)0;                                          /* :434 */
            const T *col = rhs_cols[col_index];                      /* :437 get_col */
            for (size_t k = 0; k < row_len; ++k) {                   /* :435 */
                T av = *row[k].v;                                    /* :436 */
                T b = col[row[k].col_index];                         /* :437 */
                T c = av * b;                                        /* :438 */
                value = value + c;                                   /* :439 */
            }
            FN(ocsr_insert)(result, value, row_index, col_index);    /* :442 */
        }
        free(row);
    }
    FN(ocsr_finalise)(result);                                       /* :445 */
    *out = result;
    return 0;
}

/* Same arithmetic as ocsr_mul_dense (identical order and rounding) for the row range
 * [row_begin,row_end), written densely (row-major, ld = rhs_col_count) without the result
 * Csr. Used to check sampled row blocks of full-size workloads and by the sharded gloo tests. */
int FN(ocsr_mul_dense_rows)(const OCSR *a, const T *const *rhs_cols, size_t rhs_row_count,
                            size_t rhs_col_count, size_t row_begin, size_t row_end, T *out_rowmajor)
{
    if (a->cols != rhs_row_count) return 3;
    for (size_t row_index = row_begin; row_index < row_end; ++row_index) {
        OENTRY *row;
        size_t row_len = FN(get_row_compact)(a, row_index, 0, &row);
        for (size_t col_index = 0; col_index < rhs_col_count; ++col_index) {
            T value = (T)0;
            const T *col = rhs_cols[col_index];
            for (size_t k = 0; k < row_len; ++k) {
                T c = (*row[k].v) * col[row[k].col_index];
                value = value + c;
            }
            out_rowmajor[(row_index - row_begin) * rhs_col_count + col_index] = value;
        }
        free(row);
    }
    return 0;
}

/* Adopt raw CSR arrays (copied) — the restatement of a finalised Csr built elsewhere
 * (e.g. by the synthetic generators). row_index must have rows+1 entries. */
OCSR *FN(ocsr_from_raw)(size_t rows, size_t cols, size_t nnz, const T *v, const size_t *col_index,
                        const size_t *row_index)
{
    OCSR *m = (OCSR *)calloc(1, sizeof(OCSR));
    m->rows = rows;
    m->cols = cols;
    m->v = (T *)malloc((nnz ? nnz : 1) * sizeof(T));
    m->col_index = (size_t *)malloc((nnz ? nnz : 1) * sizeof(size_t));
    m->row_index = (size_t *)malloc((rows + 1) * sizeof(size_t));
    memcpy(m->v, v, nnz * sizeof(T));
    memcpy(m->col_index, col_index, nnz * sizeof(size_t));
    memcpy(m->row_index, row_index, (rows + 1) * sizeof(size_t));
    m->v_len = m->v_cap = nnz;
    m->ri_len = m->ri_cap = rows + 1;
    m->is_finalised = 1;
    return m;
}

/* field accessors for ctypes */
size_t FN(ocsr_rows)(const OCSR *m) { return m->rows; }
size_t FN(ocsr_cols)(const OCSR *m) { return m->cols; }
size_t FN(ocsr_v_len)(const OCSR *m) { return m->v_len; }
size_t FN(ocsr_row_index_len)(const OCSR *m) { return m->ri_len; }
int FN(ocsr_is_finalised)(const OCSR *m) { return m->is_finalised; }
const T *FN(ocsr_v)(const OCSR *m) { return m->v; }
const size_t *FN(ocsr_col_index)(const OCSR *m) { return m->col_index; }
const size_t *FN(ocsr_row_index)(const OCSR *m) { return m->row_index; }

#undef OENTRY
#undef OCSR
#undef FN
#undef CAT
#undef CAT_
