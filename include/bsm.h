/*
 * bsm.h — C ABI of the B200-native Csr x Dense (SpMM / SpMV) path.
 *
 * This is the drop-in boundary for ONE reference routine and its operand types:
 *     Csr<T>::mul_dense(&self, rhs:&Dense<T>) -> Result<Csr<T>,MatErr>
 *         /root/reference/src/sparse.rs:426-446
 * plus the sibling front-end Csr<T>::mul_vector (sparse.rs:468-482).
 * The reference has no FFI of its own (pure Rust, zero dependencies); these entry points are
 * exactly what a `gpu` module inside the crate binds with `extern "C"`.
 * Plain pointers and sizes only — no CUDA, torch or C++ types in any signature.
 *
 * Conventions
 *   - every function returns a bsm_status (0 = ok); bsm_last_error_string() describes the
 *     last failure on the calling thread;
 *   - `usize` of the reference is uint64_t here (x86-64 / aarch64);
 *   - handles own device memory (HBM) unless created by a *_borrow function;
 *   - all work is enqueued on one CUDA stream per process (bsm_set_stream can adopt an
 *     external one, e.g. a torch stream, passed as an opaque pointer); functions that return
 *     host data synchronise that stream before returning;
 *   - threading: one process drives one GPU (bsm_init) from one host thread at a time, like the
 *     single-threaded reference; error strings and bsm_last_launch_info are per thread, the device,
 *     stream and allocator state are per process and NOT locked — serialise calls from several threads;
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point fails
 *     with BSM_ERR_NO_DEVICE / BSM_ERR_CUDA.
 */
#ifndef BSM_H
#define BSM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSM_ABI_VERSION 1

typedef struct bsm_csr bsm_csr;     /* device-resident CSR: vals[nnz] T, col_idx[nnz] u32, row_ptr[rows+1] u32 */
typedef struct bsm_dense bsm_dense; /* device-resident dense, ROW-major, leading dimension ld (elements)       */
typedef struct bsm_comm bsm_comm;   /* NCCL communicator for the optional all-gather of C row blocks            */

typedef enum bsm_status {
    BSM_OK = 0,
    /* maps to MatErr::IncorrectDimensions (src/util.rs:47-55; returned at sparse.rs:427-429, 469-471) */
    BSM_ERR_INCORRECT_DIMENSIONS = 1,
    /* operand Csr was not finalised (row_index shorter than rows+1; reference would panic on index) */
    BSM_ERR_NOT_FINALISED = 2,
    /* a col_index >= cols (reference: slice-index panic at sparse.rs:437) / MatErr::OutOfBounds */
    BSM_ERR_OUT_OF_BOUNDS = 3,
    /* nnz or a dimension does not fit the device's u32 indices */
    BSM_ERR_INDEX_OVERFLOW = 4,
    BSM_ERR_INVALID_ARGUMENT = 5,
    BSM_ERR_DTYPE_MISMATCH = 6,
    BSM_ERR_CUDA = 7,
    BSM_ERR_NCCL = 8,
    BSM_ERR_NO_DEVICE = 9,
    BSM_ERR_NOT_SUPPORTED = 10
} bsm_status;

typedef enum bsm_dtype { BSM_F32 = 0, BSM_F64 = 1 } bsm_dtype;

/* kernel family (north_star): warp-per-row vector CSR, or nnz-balanced merge-path */
typedef enum bsm_algo {
    BSM_ALGO_AUTO = 0,
    BSM_ALGO_VECTOR = 1,   /* row-parallel vector CSR: stored order, unfused -> bit-identical to the reference */
    BSM_ALGO_MERGE = 2,    /* nnz-balanced merge path (power-law rows)                                     */
    BSM_ALGO_ROWBLOCK = 3  /* vector CSR for band-like matrices (every row a run of consecutive columns): blocks of
                              consecutive rows share their B-row loads; same order and rounding as BSM_ALGO_VECTOR;
                              BSM_ERR_NOT_SUPPORTED when the matrix has a row that is not such a run           */
} bsm_algo;

/* bsm_tuning.flags */
#define BSM_TUNE_A_EVICT_FIRST 0x1u   /* L2 evict-first policy on the TMA bulk copies of col_idx/values */
#define BSM_TUNE_C_STREAMING   0x2u   /* st.global.cs for C rows                                       */
#define BSM_TUNE_FUSED         0x4u   /* OPT-IN: one fused multiply-add per product instead of the reference's separately
                                         rounded multiply and add (sparse.rs:438-439). Halves the FP instructions of the
                                         row-block kernel; results then agree with the reference within north_star's
                                         tolerance (1e-12 f64 / 1e-5 f32 of sum|a*b|) instead of bit for bit. Never a
                                         default. The merge-path kernel is always fused.                               */
#define BSM_TUNE_DEFAULT_FLAGS (BSM_TUNE_A_EVICT_FIRST | BSM_TUNE_C_STREAMING)

/* Launch tuning; all-zero = library heuristics. Used by the bench sweeps and tests. */
typedef struct bsm_tuning {
    int32_t algo;            /* bsm_algo                                                            */
    int32_t col_tile;        /* columns per pass over A (0 = heuristic, -1 = all n in one pass)     */
    int32_t rows_per_slice;  /* vector kernel: rows of one warp staged per TMA slice (0 = heuristic);
                                row-block kernel: rows per lane group, 4 or 8 (else heuristic)       */
    int32_t stages;          /* vector kernel: TMA ring depth per warp (0 = heuristic)              */
    int32_t warps_per_cta;   /* compute warps per CTA (0 = heuristic)                               */
    int32_t ctas_per_sm;     /* persistent grid = SMs x this (0 = heuristic)                        */
    int32_t merge_items;     /* merge-path: (rows+nnz) items per lane group (0 = heuristic)         */
    uint32_t flags;          /* BSM_TUNE_* (0 = BSM_TUNE_DEFAULT_FLAGS); bit31 set = take literally */
    int32_t rows_per_warp;   /* vector kernel: consecutive rows a warp owns inside a CTA's
                                super-batch (0 = heuristic: the matrix's dominant row stride)       */
    int32_t prefer_wide_rows;/* vector kernel: 1 = a full warp per row even when 128-bit loads need fewer
                                lanes; merge-path does that by default, -1 turns it off there       */
    int32_t reg_flavour;     /* vector kernel: register-budget variant. 0 = heuristic; 1 = CTAs of <= 512
                                threads, 1 per SM; 5 = CTAs of <= 256 threads, 3 per SM, scalar reads of the staged
                                col_idx / values; 7 = one CTA of <= 768 threads per SM, scalar reads; 8 = 7 with a
                                window of 10 gathers (one register tile per lane); 9 = four 128-bit lanes per row walking
                                flat entry streams (the default for 64-byte output rows on short regular rows; other
                                shapes run it as 8). 2, 3, 4, 6 = variants the round-1
                                sweeps rejected (deeper window, LDS.128 reads): no longer built, they run as 1, 5, 5, 7
                                (see csrc/spmm_rows_inst.cuh)                                        */
    int32_t lanes_per_row;   /* vector kernel: lanes that share one output row (power of two <= 32). 0 = heuristic.
                                Fewer lanes than the 128-bit loads need -> 2 or 4 register tiles per lane and
                                32/lanes rows side by side, each lane group walking its own flat entry stream:
                                one LDS of the A stream then feeds 32/lanes rows                            */
    int32_t reserved[4];
} bsm_tuning;

/* what the last bsm_spmm* call on this thread actually launched */
typedef struct bsm_launch_info {
    int32_t algo;            /* BSM_ALGO_VECTOR, BSM_ALGO_MERGE or BSM_ALGO_ROWBLOCK                */
    int32_t kernels;         /* kernels launched by the call                                        */
    int32_t vec_elems;       /* elements per lane per load (V)                                      */
    int32_t lanes_per_row;   /* G                                                                   */
    int32_t reg_tiles;       /* NT                                                                  */
    int32_t grid, block;     /* of the main kernel                                                  */
    int32_t smem_bytes;
    int32_t rows_per_slice, stages, capacity;   /* capacity 0 = col_idx/values not staged */
    int32_t passes;          /* column-tile passes                                                  */
    int32_t merge_items, merge_chunks;
    int32_t rows_per_warp, reg_flavour;
    int32_t col_tile;        /* columns of the FIRST pass (same meaning in bsm_spmm* and bsm_plan_vector)          */
    int32_t reserved[1];
} bsm_launch_info;

/* ------------------------------------------------------------------------------------------
 * runtime
 * ---------------------------------------------------------------------------------------- */
int bsm_abi_version(void);
/* Select the CUDA device of this process (one process per GPU) and create the stream. */
int bsm_init(int device);
int bsm_device_count(int *count);
/* Adopt an external cudaStream_t (opaque pointer; NULL = back to the library's own stream). */
int bsm_set_stream(void *cuda_stream);
int bsm_sync(void);
const char *bsm_last_error_string(void);
const char *bsm_status_string(int status);
int bsm_device_info(int *sm_count, size_t *l2_bytes, size_t *hbm_bytes, int *cc_major, int *cc_minor);

/* ------------------------------------------------------------------------------------------
 * Csr<T> operand  (reference: struct Csr, src/sparse.rs:68-78: v:Vec<T>, col_index:Vec<usize>,
 * row_index:Vec<usize>; finalised => row_index.len() == rows+1, sparse.rs:206-219)
 * ---------------------------------------------------------------------------------------- */
/* Upload a finalised host Csr. Indices are narrowed usize -> u32 ON THE DEVICE; the column
 * bound is checked there too. Column order inside a row is preserved (may be unsorted / hold
 * duplicates — the reference never sorts, sparse.rs:237-250). */
int bsm_csr_upload_f64(uint64_t rows, uint64_t cols, uint64_t nnz, const double *v,
                       const uint64_t *col_index, const uint64_t *row_index,
                       uint64_t row_index_len, bsm_csr **out);
int bsm_csr_upload_f32(uint64_t rows, uint64_t cols, uint64_t nnz, const float *v,
                       const uint64_t *col_index, const uint64_t *row_index,
                       uint64_t row_index_len, bsm_csr **out);
/* Adopt device-resident arrays already in device format (u32 indices). copy=0 borrows the
 * pointers (caller keeps them alive; each array must be 16-byte aligned and readable up to the
 * next multiple of 4 entries); copy=1 copies into library-owned HBM. */
int bsm_csr_from_device(int dtype, uint64_t rows, uint64_t cols, uint64_t nnz, const void *d_vals,
                        const uint32_t *d_col_idx, const uint32_t *d_row_ptr, int copy,
                        bsm_csr **out);
int bsm_csr_free(bsm_csr *a);
int bsm_csr_info(const bsm_csr *a, int *dtype, uint64_t *rows, uint64_t *cols, uint64_t *nnz,
                 uint64_t *max_row_nnz);
/* The per-matrix statistics the dispatch heuristics use, computed once at upload by one kernel over ALL rows: smallest / largest
 * stored column, and the stencil line length the majority of rows suggests (0 = not stencil-like) — the number of consecutive
 * rows a warp of the vector kernel owns so that the warps of a CTA sweep adjacent grid lines. Any pointer may be NULL. */
int bsm_csr_stats(const bsm_csr *a, uint64_t *max_row_nnz, uint64_t *col_min, uint64_t *col_max, uint64_t *line_length);
/* device pointers of the three arrays (for zero-copy views, e.g. torch.from_dlpack-free use) */
int bsm_csr_device_ptrs(const bsm_csr *a, const void **d_vals, const uint32_t **d_col_idx,
                        const uint32_t **d_row_ptr);
/* Download to the reference layout (usize indices). Buffers sized nnz / nnz / rows+1. */
int bsm_csr_download_f64(const bsm_csr *a, double *v, uint64_t *col_index, uint64_t *row_index);
int bsm_csr_download_f32(const bsm_csr *a, float *v, uint64_t *col_index, uint64_t *row_index);

/* ------------------------------------------------------------------------------------------
 * Dense<T> operand  (reference: struct Dense, src/dense.rs:4-9: data:Vec<Vec<T>>, COLUMN-major,
 * data[c][r]; get_dims() = {rows:row_count, cols:col_count}, dense.rs:40-47)
 * ---------------------------------------------------------------------------------------- */
/* col_ptrs[c] points at host column c (length rows): the Vec<Vec<T>> as it lies in memory.
 * The columns are copied to HBM and transposed on the device into the row-major layout the
 * kernels gather from. */
int bsm_dense_upload_f64(uint64_t rows, uint64_t cols, const double *const *col_ptrs, bsm_dense **out);
int bsm_dense_upload_f32(uint64_t rows, uint64_t cols, const float *const *col_ptrs, bsm_dense **out);
int bsm_dense_alloc(int dtype, uint64_t rows, uint64_t cols, bsm_dense **out);
/* Borrow an existing row-major device buffer (ld in elements, ld >= cols). */
int bsm_dense_borrow(int dtype, uint64_t rows, uint64_t cols, void *d_rowmajor, uint64_t ld,
                     bsm_dense **out);
int bsm_dense_free(bsm_dense *d);
int bsm_dense_zero(bsm_dense *d);   /* all elements (padding columns included) = 0, on the library stream */
int bsm_dense_info(const bsm_dense *d, int *dtype, uint64_t *rows, uint64_t *cols, uint64_t *ld,
                   void **d_ptr);
/* Download into host columns (col_ptrs[c] has room for `rows` elements). */
int bsm_dense_download_f64(const bsm_dense *d, double *const *col_ptrs);
int bsm_dense_download_f32(const bsm_dense *d, float *const *col_ptrs);
/* Row-major host copies (tests / C++ callers). dst/src hold rows*cols elements, ld = cols. */
int bsm_dense_download_rowmajor(const bsm_dense *d, void *dst);
int bsm_dense_upload_rowmajor(int dtype, uint64_t rows, uint64_t cols, const void *src, bsm_dense **out);

/* ------------------------------------------------------------------------------------------
 * THE HOT PATH — replaces Csr::mul_dense, src/sparse.rs:426-446
 * C (rows(A) x cols(B), device-resident, row-major) = A * B.
 *   A.cols != B.rows                      -> BSM_ERR_INCORRECT_DIMENSIONS (sparse.rs:427-429)
 *   C dims != (A.rows, B.cols)            -> BSM_ERR_INCORRECT_DIMENSIONS
 * Summation order: BSM_ALGO_VECTOR walks each row's stored entries in stored order with a
 * separately rounded multiply and add (sparse.rs:434-439) — bit-identical to the reference.
 * BSM_ALGO_MERGE uses FMA and splits long rows (deterministic, tolerance-level agreement).
 * ---------------------------------------------------------------------------------------- */
int bsm_spmm(const bsm_csr *a, const bsm_dense *b, bsm_dense *c, int algo);
int bsm_spmm_tuned(const bsm_csr *a, const bsm_dense *b, bsm_dense *c, const bsm_tuning *tuning);
int bsm_last_launch_info(bsm_launch_info *info);
/* Dry run of the vector kernel's launch heuristics for a matrix described by its statistics (row_stride = line length
 * of a stencil-like matrix, 0 = none) on a device described by its SM count and opt-in shared memory per CTA.
 * Pure host arithmetic: needs no device. Operands are assumed 16-byte aligned. */
/* The stencil line length one row suggests (what the upload samples from three rows to set `row_stride`): the smallest
 * distance > 1 of a stored column from the diagonal `diag`; nx for a box stencil storing nx-1, nx, nx+1. Pure host. */
uint32_t bsm_line_length_of_row(const uint32_t *cols, uint32_t len, uint64_t diag);
int bsm_plan_vector(int dtype, uint64_t rows, uint64_t nnz, uint64_t max_row_nnz, uint32_t row_stride, uint64_t n_cols,
                    const bsm_tuning *tuning, int sm_count, uint64_t smem_optin_bytes, bsm_launch_info *out);
/* cumulative number of this library's kernels launched by this process (bench "gpu_launches") */
uint64_t bsm_kernel_launch_count(void);

/* Result construction of mul_dense: every output goes through Csr::insert, which drops values
 * equal to T::default() (sparse.rs:442 -> 222-233; -0.0 dropped, NaN kept), then finalise
 * (206-219). Device-side count -> scan -> scatter; returns a device Csr (rows x cols of d). */
int bsm_dense_to_csr(const bsm_dense *d, bsm_csr **out);

/* Residual check of BASELINE config 5 (||AX - B||): Frobenius norms ||ax - b|| and ||b|| of two
 * device-resident dense matrices of equal shape, accumulated in f64 with a fixed reduction tree. */
int bsm_dense_residual_norm(const bsm_dense *ax, const bsm_dense *b, double *resid_fro, double *b_fro);

/* The two substitutions of the reference's `solve` (src/lib.rs:11-24), on the device, for device-resident operands:
 *   bsm_forward_substitution  = forward_substitution(l, b)        lib.rs:28-46   L y = b, rows ascending:
 *       l_x = sum, in stored order, of v * y[col] over the stored entries of row r whose column != r;
 *       y[r] = (b[r] - l_x) / (last stored entry of row r);
 *   bsm_backward_substitution = backward_substitution(l_star, y)  lib.rs:49-65   L* x = y, rows descending:
 *       the first stored entry of a row is skipped, x[r] = (y[r] - l_x) / (first stored entry of row r).
 * Every product and sum is rounded separately, in the reference's order -> bit-identical to the reference for any data
 * (the reference is f32 only; f64 is offered with the same semantics). One lane per right-hand-side column, rows in
 * sequence: latency-bound by construction. l must be square with rows(l) == rows(b); y / x: rows(b) x cols(b), distinct
 * from b. A row without a stored entry (the reference panics there) -> BSM_ERR_INVALID_ARGUMENT. The factorisation itself
 * (Csr::cholesky_decomp, sparse.rs:682-714, and transpose) is out of scope and stays with the caller. */
int bsm_forward_substitution(const bsm_csr *l, const bsm_dense *b, bsm_dense *y);
int bsm_backward_substitution(const bsm_csr *l_star, const bsm_dense *y, bsm_dense *x);
/* Which substitution kernel a factor gets (probed on the device once per handle, cached). *lower_hb / *upper_hb = the
 * half-bandwidth when the matrix is a PROPER lower / upper band factor — row r stores exactly the columns
 * max(0, r-hb) .. r (diagonal last) / r .. min(n-1, r+hb) (diagonal first), what cholesky_decomp and transpose() produce for a
 * band matrix (sparse.rs:682-714, 296-318) — else -1. Proper band factors of half-bandwidth 8, 16 or 32 with at least 4 hb rows
 * run a specialised kernel (one solver warp, no hand-overs between warps; same arithmetic, same bits); everything else the general
 * one (environment: BSM_SOLVE_GENERAL=1 forces the general kernel — the tests compare the two). The band kernels divide f32 by a
 * shortcut that is the correctly rounded quotient while numerators stay in [2^-90, 2^90) (zeros included) and divisors in
 * [2^-30, 2^30); when an operand leaves these ranges (infinity, NaN, subnormal, a solution that decays towards zero) the call
 * discards the result and runs the general kernel: slower, same bits. Either pointer may be NULL. */
int bsm_csr_band_structure(const bsm_csr *a, int32_t *lower_hb, int32_t *upper_hb);

/* Host-to-host convenience = the literal reference call: uploads A and B, multiplies, compacts
 * and returns the zero-dropped result Csr in reference layout. The result arrays are allocated
 * by the library; release with bsm_host_free. */
int bsm_mul_dense_host_f64(uint64_t rows, uint64_t cols, uint64_t nnz, const double *v,
                           const uint64_t *col_index, const uint64_t *row_index,
                           uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                           const double *const *rhs_col_ptrs, int algo, uint64_t *out_nnz,
                           double **out_v, uint64_t **out_col_index, uint64_t **out_row_index);
int bsm_mul_dense_host_f32(uint64_t rows, uint64_t cols, uint64_t nnz, const float *v,
                           const uint64_t *col_index, const uint64_t *row_index,
                           uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                           const float *const *rhs_col_ptrs, int algo, uint64_t *out_nnz,
                           float **out_v, uint64_t **out_col_index, uint64_t **out_row_index);
/* The same literal call into CALLER-PROVIDED result arrays (out_v / out_col_index: room for `capacity` entries;
 * out_row_index: rows + 1): what a binding uses to fill a Csr whose Vecs it allocated (Vec::with_capacity), and what
 * bench.py times end to end with pinned arrays. rows * rhs_cols entries always suffice; too small a capacity fails
 * with BSM_ERR_INVALID_ARGUMENT (nothing useful is left in the arrays). The call is pipelined: B travels host -> device
 * in chunks of rows, each block of output rows is multiplied as soon as the B rows its columns reach have landed, and
 * its zero-dropped entries travel device -> host while the next block is computed: the values and the row_index piece as
 * they are, the column indices as one keep-bit per output, which a few threads inside the call expand into out_col_index
 * (a full row — the normal case of Csr x Dense — is the pattern 0 .. rhs_cols-1). The arrays hold exactly what the reference's
 * insert / finalise produce. Use pinned host memory for the copies to overlap.
 * Environment: BSM_PIPE_EXPAND_THREADS = expansion threads (default 8, at most half of the hardware threads; 0 = usize
 * columns are written on the device and copied, 8 bytes per entry); BSM_PIPE_EXPAND_NT=0 = plain instead of non-temporal
 * stores; BSM_PIPE_EXPAND_MIN_ENTRIES = smallest result (rows x rhs_cols, default 4 Mi) that takes the masks — smaller
 * products are launch-bound and copy their usize columns; BSM_PIPE_BLOCK_BYTES / BSM_PIPE_CHUNK_BYTES = sizes of the row
 * blocks / B chunks (tests). */
int bsm_mul_dense_host_into_f64(uint64_t rows, uint64_t cols, uint64_t nnz, const double *v,
                                const uint64_t *col_index, const uint64_t *row_index,
                                uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                                const double *const *rhs_col_ptrs, int algo, uint64_t capacity, double *out_v,
                                uint64_t *out_col_index, uint64_t *out_row_index, uint64_t *out_nnz);
int bsm_mul_dense_host_into_f32(uint64_t rows, uint64_t cols, uint64_t nnz, const float *v,
                                const uint64_t *col_index, const uint64_t *row_index,
                                uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                                const float *const *rhs_col_ptrs, int algo, uint64_t capacity, float *out_v,
                                uint64_t *out_col_index, uint64_t *out_row_index, uint64_t *out_nnz);
/* Same product with a DENSE result in the reference's column-major layout: host Csr and host Dense
 * columns in, host Dense columns out (out_col_ptrs[c] has room for `rows` elements). The same pipeline
 * (chunks of B rows in, blocks of output rows out); use pinned host buffers for the overlap to materialise. */
int bsm_mul_dense_host_dense_f64(uint64_t rows, uint64_t cols, uint64_t nnz, const double *v,
                                 const uint64_t *col_index, const uint64_t *row_index,
                                 uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                                 const double *const *rhs_col_ptrs, double *const *out_col_ptrs, int algo);
int bsm_mul_dense_host_dense_f32(uint64_t rows, uint64_t cols, uint64_t nnz, const float *v,
                                 const uint64_t *col_index, const uint64_t *row_index,
                                 uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                                 const float *const *rhs_col_ptrs, float *const *out_col_ptrs, int algo);
void bsm_host_free(void *p);

/* Csr::mul_vector(&self, rhs:&[T], out:&mut [T]) -> Result<(),MatErr>  (sparse.rs:468-482):
 * host slices in and out, dense result (no zero-drop);
 * a.cols != rhs_len || a.rows != out_len -> BSM_ERR_INCORRECT_DIMENSIONS (469-471). */
int bsm_mul_vector_f64(const bsm_csr *a, const double *rhs, uint64_t rhs_len, double *out, uint64_t out_len);
int bsm_mul_vector_f32(const bsm_csr *a, const float *rhs, uint64_t rhs_len, float *out, uint64_t out_len);

/* ------------------------------------------------------------------------------------------
 * Row-partitioned multi-GPU (one process per GPU). B is replicated; no data-path collective.
 * ---------------------------------------------------------------------------------------- */
/* nnz-balanced contiguous row split: bounds[p] = first row of part p, bounds[parts] = rows.
 * Pure host function over the reference's row_index (usize). */
int bsm_partition_rows(const uint64_t *row_index, uint64_t rows, int parts, uint64_t *bounds);
/* Upload rows [row_begin,row_end) of a host Csr as a device Csr with rebased row_ptr. */
int bsm_csr_upload_rows_f64(uint64_t rows, uint64_t cols, const double *v, const uint64_t *col_index,
                            const uint64_t *row_index, uint64_t row_begin, uint64_t row_end,
                            bsm_csr **out);
int bsm_csr_upload_rows_f32(uint64_t rows, uint64_t cols, const float *v, const uint64_t *col_index,
                            const uint64_t *row_index, uint64_t row_begin, uint64_t row_end,
                            bsm_csr **out);
/* Optional gathered result: NCCL all-gather(v) of the C row blocks over NVLink.
 * unique_id is NCCL_UNIQUE_ID_BYTES (128) bytes, created on rank 0 and shipped to the other
 * ranks by the caller (torch.distributed / any transport). */
int bsm_comm_unique_id(char id[128]);
int bsm_comm_init(const char id[128], int nranks, int rank, bsm_comm **out);
int bsm_comm_free(bsm_comm *c);
/* full (rows_total x cols, row-major, on every rank) <- concat over ranks of local blocks;
 * bounds[] as returned by bsm_partition_rows (nranks+1 entries). */
int bsm_allgather_rows(bsm_comm *c, const bsm_dense *local_block, const uint64_t *bounds, bsm_dense *full);

/* Gathered product WITHOUT a collective: C_full[0] is this rank's full (rows_total x n) result buffer,
 * C_full[1..ndest-1] are the other ranks' full buffers mapped into this process (bsm_dense_ipc_open).
 * The SpMM kernel stores every row it finishes (global row = row_offset + local row) to ALL of them —
 * multiply and all-gather fused into one kernel, P2P stores over NVLink instead of NCCL. Follow with
 * bsm_comm_barrier (or any cross-rank barrier) before reading a full buffer. Vector-CSR kernel only:
 * the merge-path kernel revisits C rows and returns BSM_ERR_NOT_SUPPORTED here. */
int bsm_spmm_scatter(const bsm_csr *a, const bsm_dense *b, bsm_dense *const *c_full, int ndest,
                     uint64_t row_offset, int algo);
/* CUDA IPC plumbing for the above: export a library-allocated dense buffer as 64 opaque bytes, ship
 * them to the other processes of the box, open them there (free with bsm_dense_free). */
int bsm_dense_ipc_export(const bsm_dense *d, char handle[64]);
int bsm_dense_ipc_open(const char handle[64], int dtype, uint64_t rows, uint64_t cols, uint64_t ld,
                       bsm_dense **out);
int bsm_comm_barrier(bsm_comm *c);

/* ------------------------------------------------------------------------------------------
 * Synthetic workloads generated directly in HBM (bench / test support; counter-based hash so
 * the CPU oracle regenerates any element independently — see basic_sparse_matrix_b200/gen.py)
 * ---------------------------------------------------------------------------------------- */
/* value mode: 0 = "exact" dyadic rationals k/1024 (every product and sum exact in f64),
 *             1 = "real" uniform [0,1) + offset */
int bsm_gen_dense(int dtype, uint64_t rows, uint64_t cols, uint64_t seed, int mode, double offset,
                  bsm_dense **out);
/* 2-D 5-point / 3-D 7-point Laplacian rows [row_begin,row_end) of the nx*ny(*nz) grid
 * (nz = 1 for 2-D), diagonal 4 / 6, off-diagonals -1, columns ascending; rebased row_ptr. */
int bsm_gen_laplacian(int dtype, uint64_t nx, uint64_t ny, uint64_t nz, uint64_t row_begin,
                      uint64_t row_end, bsm_csr **out);
/* SPD band, half-bandwidth hb: a_ij = -1/(1+|i-j|), a_ii = 1 + sum_j |a_ij|. */
int bsm_gen_band(int dtype, uint64_t n, uint64_t hb, uint64_t row_begin, uint64_t row_end, bsm_csr **out);
/* R-MAT: 2^scale rows/cols, `edges` draws with (a,b,c,d), sorted by (row,col), duplicates
 * kept; values from the hash in the given mode. */
int bsm_gen_rmat(int dtype, int scale, uint64_t edges, double a, double b, double c, uint64_t seed,
                 int mode, bsm_csr **out);

/* Where the time of the host-to-host calls goes (diagnostics; off unless BSM_PHASE_TIMERS=1 or enabled here).
 * bsm_phase_timers_read copies the accumulated seconds of the first `count` phases (bsm_phase_name(i), "" past the
 * last) and optionally resets them: wall time of the A upload and of the whole call and of the host's waits, device
 * busy time of the H2D copies, transposes, SpMM, result construction and D2H copies. */
int bsm_phase_timers_enable(int on);
int bsm_phase_timers_read(double *seconds, int count, int reset);
const char *bsm_phase_name(int phase);

/* raw device memory helpers for callers that keep their own buffers */
int bsm_l2_flush(void);   /* overwrite a buffer larger than L2 (timing hygiene) */

#ifdef __cplusplus
}
#endif
#endif /* BSM_H */
