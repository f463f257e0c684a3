//! Raw bindings of include/bsm.h (ABI version 1). usize == u64 on the supported targets.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct bsm_csr { _p: [u8; 0] }
#[repr(C)] pub struct bsm_dense { _p: [u8; 0] }
#[repr(C)] pub struct bsm_comm { _p: [u8; 0] }

pub const BSM_OK: c_int = 0;
pub const BSM_ERR_INCORRECT_DIMENSIONS: c_int = 1;
pub const BSM_ERR_NOT_FINALISED: c_int = 2;
pub const BSM_ERR_OUT_OF_BOUNDS: c_int = 3;
pub const BSM_ERR_INDEX_OVERFLOW: c_int = 4;
pub const BSM_ERR_CUDA: c_int = 7;
pub const BSM_ERR_NCCL: c_int = 8;
pub const BSM_ERR_NO_DEVICE: c_int = 9;

pub const BSM_F32: c_int = 0;
pub const BSM_F64: c_int = 1;
pub const BSM_ALGO_AUTO: c_int = 0;
pub const BSM_ALGO_VECTOR: c_int = 1;
pub const BSM_ALGO_MERGE: c_int = 2;

extern "C" {
    pub fn bsm_abi_version() -> c_int;
    pub fn bsm_init(device: c_int) -> c_int;
    pub fn bsm_sync() -> c_int;
    pub fn bsm_last_error_string() -> *const c_char;

    pub fn bsm_csr_upload_f64(rows: u64, cols: u64, nnz: u64, v: *const f64, col_index: *const u64,
                              row_index: *const u64, row_index_len: u64, out: *mut *mut bsm_csr) -> c_int;
    pub fn bsm_csr_upload_f32(rows: u64, cols: u64, nnz: u64, v: *const f32, col_index: *const u64,
                              row_index: *const u64, row_index_len: u64, out: *mut *mut bsm_csr) -> c_int;
    pub fn bsm_csr_upload_rows_f64(rows: u64, cols: u64, v: *const f64, col_index: *const u64, row_index: *const u64,
                                   row_begin: u64, row_end: u64, out: *mut *mut bsm_csr) -> c_int;
    pub fn bsm_csr_upload_rows_f32(rows: u64, cols: u64, v: *const f32, col_index: *const u64, row_index: *const u64,
                                   row_begin: u64, row_end: u64, out: *mut *mut bsm_csr) -> c_int;
    pub fn bsm_csr_info(a: *const bsm_csr, dtype: *mut c_int, rows: *mut u64, cols: *mut u64, nnz: *mut u64,
                        max_row_nnz: *mut u64) -> c_int;
    pub fn bsm_csr_download_f64(a: *const bsm_csr, v: *mut f64, col_index: *mut u64, row_index: *mut u64) -> c_int;
    pub fn bsm_csr_download_f32(a: *const bsm_csr, v: *mut f32, col_index: *mut u64, row_index: *mut u64) -> c_int;
    pub fn bsm_csr_free(a: *mut bsm_csr) -> c_int;

    pub fn bsm_dense_upload_f64(rows: u64, cols: u64, col_ptrs: *const *const f64, out: *mut *mut bsm_dense) -> c_int;
    pub fn bsm_dense_upload_f32(rows: u64, cols: u64, col_ptrs: *const *const f32, out: *mut *mut bsm_dense) -> c_int;
    pub fn bsm_dense_alloc(dtype: c_int, rows: u64, cols: u64, out: *mut *mut bsm_dense) -> c_int;
    pub fn bsm_dense_info(d: *const bsm_dense, dtype: *mut c_int, rows: *mut u64, cols: *mut u64, ld: *mut u64,
                          d_ptr: *mut *mut c_void) -> c_int;
    pub fn bsm_dense_download_f64(d: *const bsm_dense, col_ptrs: *const *mut f64) -> c_int;
    pub fn bsm_dense_download_f32(d: *const bsm_dense, col_ptrs: *const *mut f32) -> c_int;
    pub fn bsm_dense_free(d: *mut bsm_dense) -> c_int;

    /// THE HOT PATH: replaces the loop nest of Csr::mul_dense (src/sparse.rs:431-444).
    pub fn bsm_spmm(a: *const bsm_csr, b: *const bsm_dense, c: *mut bsm_dense, algo: c_int) -> c_int;
    /// Result construction of mul_dense: insert's zero-drop + finalise (sparse.rs:442, 222-233, 206-219).
    pub fn bsm_dense_to_csr(d: *const bsm_dense, out: *mut *mut bsm_csr) -> c_int;

    /// host Csr x host Dense -> host Dense (column-major), pipelined H2D | multiply | D2H per column group
    pub fn bsm_mul_dense_host_dense_f64(rows: u64, cols: u64, nnz: u64, v: *const f64, col_index: *const u64, row_index: *const u64,
                                        row_index_len: u64, rhs_rows: u64, rhs_cols: u64, rhs_col_ptrs: *const *const f64,
                                        out_col_ptrs: *const *mut f64, algo: c_int) -> c_int;
    pub fn bsm_mul_dense_host_dense_f32(rows: u64, cols: u64, nnz: u64, v: *const f32, col_index: *const u64, row_index: *const u64,
                                        row_index_len: u64, rhs_rows: u64, rhs_cols: u64, rhs_col_ptrs: *const *const f32,
                                        out_col_ptrs: *const *mut f32, algo: c_int) -> c_int;
    pub fn bsm_mul_vector_f64(a: *const bsm_csr, rhs: *const f64, rhs_len: u64, out: *mut f64, out_len: u64) -> c_int;
    pub fn bsm_mul_vector_f32(a: *const bsm_csr, rhs: *const f32, rhs_len: u64, out: *mut f32, out_len: u64) -> c_int;

    pub fn bsm_partition_rows(row_index: *const u64, rows: u64, parts: c_int, bounds: *mut u64) -> c_int;
    pub fn bsm_comm_unique_id(id: *mut c_char) -> c_int;
    pub fn bsm_comm_init(id: *const c_char, nranks: c_int, rank: c_int, out: *mut *mut bsm_comm) -> c_int;
    pub fn bsm_comm_free(c: *mut bsm_comm) -> c_int;
    /// multiply + all-gather fused: every C row is stored to all destinations (own + peer GPUs, P2P over NVLink)
    pub fn bsm_spmm_scatter(a: *const bsm_csr, b: *const bsm_dense, c_full: *const *mut bsm_dense, ndest: c_int,
                            row_offset: u64, algo: c_int) -> c_int;
    pub fn bsm_dense_ipc_export(d: *const bsm_dense, handle: *mut c_char) -> c_int;
    pub fn bsm_dense_ipc_open(handle: *const c_char, dtype: c_int, rows: u64, cols: u64, ld: u64, out: *mut *mut bsm_dense) -> c_int;
    pub fn bsm_comm_barrier(c: *mut bsm_comm) -> c_int;
    pub fn bsm_allgather_rows(c: *mut bsm_comm, local_block: *const bsm_dense, bounds: *const u64,
                              full: *mut bsm_dense) -> c_int;
}
