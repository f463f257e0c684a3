// kernels.h — host-callable launchers of the sm_100a kernels (internal interface).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bsm {

// How one output row is spread over a warp: each lane owns V consecutive columns per register
// tile, G lanes form the group that owns a row, NT register tiles per lane.
// One pass covers G*V*NT columns.
struct Shape {
    int V = 1, G = 1, NT = 1;
};

// ---- vector-CSR ("row") kernel -----------------------------------------------------------
struct RowParams {
    const uint32_t *row_ptr;
    const uint32_t *col_idx;
    const void *vals;
    const void *B;   // already offset to the first column of this pass
    void *C;         // likewise
    uint32_t rows;
    uint32_t n;      // columns in this pass
    uint32_t ldb, ldc;
    uint32_t P;          // consecutive rows owned by one warp inside a super-batch (>= R; a multiple of R for the row-by-row shapes)
    uint32_t R;          // rows per TMA slice of one warp (multiple of 4 and of 32/G)
    uint32_t num_super;  // super-batches = ceil(rows / (warps * P))
    uint32_t cap;        // staged entries per slice (multiple of 4)
    uint32_t stages;     // TMA ring depth per warp
    uint32_t flags;      // BSM_TUNE_*
    uint32_t n_peers;    // scatter variant: further destinations of every C row (0 = none)
    char *peers[7];      // their C pointers, offset like C (first column of the pass, this rank's first row)
};
size_t row_kernel_smem_bytes(int dtype, const RowParams &p, int warps);
// flavour: register-budget variant of the kernel (-1 = unstaged), see spmm_rows_inst.cuh; multi: scatter variant
int row_kernel_occupancy(int dtype, Shape sh, uint32_t n, int flavour, bool multi, int block, size_t smem, int *blocks_per_sm);
int launch_spmm_rows(int dtype, Shape sh, const RowParams &p, int flavour, bool multi, int grid, int block, size_t smem,
                     int ctas_per_sm, cudaStream_t stream);

// ---- merge-path kernel ---------------------------------------------------------------------
struct MergeParams {
    const uint32_t *row_ptr;
    const uint32_t *col_idx;
    const void *vals;
    const void *B;
    void *C;
    const uint32_t *part_rows;  // [num_chunks+1]
    void *carry_vals;           // [num_chunks][ldcar]
    uint32_t *long_rows;        // [long_cap][2]: (row, first carrying chunk) of runs too long for one warp
    uint32_t *long_count;       // [1]
    uint32_t rows, nnz;
    uint32_t n, ldb, ldc, ldcar;
    uint32_t items;             // merge items (rows + nnz) per lane group
    uint32_t num_chunks;
    uint32_t long_cap;
    uint32_t flags;
};
size_t merge_kernel_smem_bytes(int dtype, Shape sh, int block, uint32_t items);
int launch_merge_partition(const uint32_t *row_ptr, uint32_t rows, uint32_t nnz, uint32_t items,
                           uint32_t num_chunks, uint32_t *part_rows, cudaStream_t stream);
int launch_spmm_merge(int dtype, Shape sh, const MergeParams &p, int block, size_t smem, int ctas_per_sm /* 0 = as many as fit */,
                      cudaStream_t stream, int *grid_out);
int launch_merge_fixup(int dtype, const MergeParams &p, cudaStream_t stream, int *launched);

// ---- row-block kernel for band-like matrices (spmm_rowblock.cu) --------------------------------------
constexpr uint32_t kRowBlockRows = 8;   // consecutive rows one lane group accumulates side by side
struct RowBlockParams {
    const uint32_t *row_ptr;
    const uint32_t *col_idx;
    const void *vals;
    const void *B;   // already offset to the first column of this pass
    void *C;         // likewise
    uint32_t rows, n, ldb, ldc, flags;
    uint32_t cap;    // staged values per warp (set by the launcher)
};
// out[0] += sum over blocks of kRowBlockRows rows of the length of the union of their column ranges,
// out[1] += blocks holding a row whose stored columns are not a run of consecutive indices
int launch_rowblock_probe(const uint32_t *row_ptr, const uint32_t *col_idx, uint64_t rows, unsigned long long *out, cudaStream_t stream);
int launch_spmm_rowblock(int dtype, Shape sh, const RowBlockParams &p, int rows_per_block /* 4 or 8 */, uint64_t max_row_nnz, int sm_count, size_t smem_max, cudaStream_t stream,
                         int *grid_out, int *block_out, int *smem_out, int *rb_out);

// ---- format conversion / construction (convert.cu) ---------------------------------------------
int launch_narrow_u64(const uint64_t *src, uint32_t *dst, uint64_t count, uint64_t bound, uint64_t subtract,
                      uint32_t *flag /* set to 1 when (v - subtract) >= bound */, cudaStream_t stream);
// per-matrix statistics in one kernel; `stats` = kStatWords u32 (line_length.h), zeroed except [kStatColMin] = ~0u
int launch_csr_stats(const uint32_t *row_ptr, const uint32_t *col_idx, uint64_t rows, uint64_t nnz, uint64_t row_offset, uint32_t *stats,
                     cudaStream_t stream);
int launch_residual_norms(int dtype, const void *x, uint64_t ldx, const void *y, uint64_t ldy, uint64_t rows, uint64_t cols,
                          double *partial, cudaStream_t stream);   // result in partial[scratch-2], partial[scratch-1]
int residual_norm_scratch_doubles();
int launch_transpose_cm2rm(int dtype, const void *colmajor, void *rowmajor, uint64_t rows, uint64_t cols, uint64_t ld,
                           cudaStream_t stream);
int launch_transpose_rm2cm(int dtype, const void *rowmajor, void *colmajor, uint64_t rows, uint64_t cols, uint64_t ld,
                           cudaStream_t stream);
int launch_widen_u32(const uint32_t *src, uint64_t *dst, uint64_t count, cudaStream_t stream);
int exclusive_scan_u32(const uint32_t *in, uint32_t *out, uint64_t count, cudaStream_t stream);  // out may alias in
int launch_count_nonzero(int dtype, const void *dense, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t *counts,
                         unsigned long long *total, cudaStream_t stream, uint64_t *masks = nullptr /* [rows][ceil(cols/64)] keep-bits, optional */);
int launch_scatter_nonzero(int dtype, const void *dense, uint64_t rows, uint64_t cols, uint64_t ld,
                           const uint32_t *row_ptr, void *vals, uint32_t *col_idx, cudaStream_t stream);
int launch_block_col_range(const uint32_t *row_ptr, const uint32_t *col_idx, uint64_t rows, uint64_t block_rows, uint32_t nblocks,
                           uint32_t *out /* [2*nblocks], preset to {~0u, 0} pairs */, cudaStream_t stream);
int launch_scatter_nonzero64(int dtype, const void *dense, uint64_t rows, uint64_t cols, uint64_t ld, const uint32_t *row_ptr, void *vals,
                             uint64_t *col_index, cudaStream_t stream);
int launch_row_index_piece(const uint32_t *local_rp, uint64_t rows, unsigned long long *tot /* [nblocks+1] */, uint32_t k, uint64_t *out64,
                           unsigned long long *host_tot /* mapped pinned mirror of tot, or null */, cudaStream_t stream);
int launch_fill_u32(uint32_t *dst, uint64_t count, uint32_t value, cudaStream_t stream);

// ---- synthetic generators (gen.cu) -----------------------------------------------------------------
int launch_gen_dense(int dtype, void *data, uint64_t rows, uint64_t cols, uint64_t ld, uint64_t seed, int mode,
                     double offset, cudaStream_t stream);
int launch_laplacian_counts(uint64_t nx, uint64_t ny, uint64_t nz, uint64_t row_begin, uint64_t row_end,
                            uint32_t *counts, cudaStream_t stream);
int launch_laplacian_fill(int dtype, uint64_t nx, uint64_t ny, uint64_t nz, uint64_t row_begin, uint64_t row_end,
                          const uint32_t *row_ptr, uint32_t *col_idx, void *vals, cudaStream_t stream);
int launch_band_counts(uint64_t n, uint64_t hb, uint64_t row_begin, uint64_t row_end, uint32_t *counts,
                       cudaStream_t stream);
int launch_band_fill(int dtype, uint64_t n, uint64_t hb, uint64_t row_begin, uint64_t row_end, const uint32_t *row_ptr,
                     uint32_t *col_idx, void *vals, cudaStream_t stream);
int gen_rmat_device(int dtype, int scale, uint64_t edges, double a, double b, double c, uint64_t seed, int mode,
                    uint32_t *row_ptr /* [2^scale+1] */, uint32_t *col_idx, void *vals, cudaStream_t stream);

}  // namespace bsm
