//! `sparse_matrix::gpu` — device-resident Csr/Dense and the B200 `mul_dense`.
//!
//! Same multiplication API as the CPU types (src/sparse.rs:426-446): `DeviceCsr::mul_dense(&self,
//! &DeviceDense) -> Result<DeviceDense, GpuError>` keeps operands and product in HBM between calls;
//! `DeviceDense::into_csr()` + `DeviceCsr::to_host()` give the reference's zero-dropped result `Csr`.
//! `Csr::mul_dense_gpu` is the literal drop-in: host Csr/Dense in, host Csr out.
//! There is no CPU fallback: every call fails with GpuError::Cuda/NoDevice when no B200 is usable.
mod ffi;

use crate::dense::Dense;
use crate::sparse::Csr;
use crate::util::{GetDims, MatDim, MatErr};
use std::ffi::CStr;
use std::marker::PhantomData;
use std::ptr;

/// MatErr is an exhaustive public enum (util.rs:47-55): device errors get their own type instead
/// of new MatErr variants, which would break downstream `match`es.
#[derive(Debug, PartialEq)]
pub enum GpuError {
    Mat(MatErr),
    IndexOverflow,
    Cuda(String),
    Nccl(String),
    NoDevice(String),
    Other(i32, String),
}

fn check(status: i32) -> Result<(), GpuError> {
    if status == ffi::BSM_OK { return Ok(()); }
    let msg = unsafe { CStr::from_ptr(ffi::bsm_last_error_string()) }.to_string_lossy().into_owned();
    Err(match status {
        ffi::BSM_ERR_INCORRECT_DIMENSIONS => GpuError::Mat(MatErr::IncorrectDimensions),
        ffi::BSM_ERR_NOT_FINALISED => GpuError::Mat(MatErr::MatrixNotFinalised),
        ffi::BSM_ERR_OUT_OF_BOUNDS => GpuError::Mat(MatErr::OutOfBounds),
        ffi::BSM_ERR_INDEX_OVERFLOW => GpuError::IndexOverflow,
        ffi::BSM_ERR_CUDA => GpuError::Cuda(msg),
        ffi::BSM_ERR_NCCL => GpuError::Nccl(msg),
        ffi::BSM_ERR_NO_DEVICE => GpuError::NoDevice(msg),
        s => GpuError::Other(s, msg),
    })
}

/// Element types with device kernels (the reference's `T` is generic; the GPU scope is f32/f64).
pub trait GpuScalar: Copy + Default + PartialEq + std::fmt::Debug {
    const DTYPE: i32;
    unsafe fn csr_upload(rows: u64, cols: u64, nnz: u64, v: *const Self, ci: *const u64, ri: *const u64, ri_len: u64,
                         out: *mut *mut ffi::bsm_csr) -> i32;
    unsafe fn csr_download(a: *const ffi::bsm_csr, v: *mut Self, ci: *mut u64, ri: *mut u64) -> i32;
    unsafe fn dense_upload(rows: u64, cols: u64, cols_ptr: *const *const Self, out: *mut *mut ffi::bsm_dense) -> i32;
    unsafe fn dense_download(d: *const ffi::bsm_dense, cols_ptr: *const *mut Self) -> i32;
    unsafe fn mul_vector(a: *const ffi::bsm_csr, rhs: *const Self, n: u64, out: *mut Self, m: u64) -> i32;
    #[allow(clippy::too_many_arguments)]
    unsafe fn mul_dense_host_into(rows: u64, cols: u64, nnz: u64, v: *const Self, ci: *const u64, ri: *const u64, ri_len: u64,
                                  rhs_rows: u64, rhs_cols: u64, rhs_cols_ptr: *const *const Self, algo: i32, capacity: u64,
                                  out_v: *mut Self, out_ci: *mut u64, out_ri: *mut u64, out_nnz: *mut u64) -> i32;
}
macro_rules! impl_scalar {
    ($t:ty, $code:expr, $up:ident, $down:ident, $dup:ident, $ddown:ident, $mv:ident) => {
        impl GpuScalar for $t {
            const DTYPE: i32 = $code;
            unsafe fn csr_upload(r: u64, c: u64, n: u64, v: *const Self, ci: *const u64, ri: *const u64, l: u64,
                                 o: *mut *mut ffi::bsm_csr) -> i32 { ffi::$up(r, c, n, v, ci, ri, l, o) }
            unsafe fn csr_download(a: *const ffi::bsm_csr, v: *mut Self, ci: *mut u64, ri: *mut u64) -> i32 { ffi::$down(a, v, ci, ri) }
            unsafe fn dense_upload(r: u64, c: u64, p: *const *const Self, o: *mut *mut ffi::bsm_dense) -> i32 { ffi::$dup(r, c, p, o) }
            unsafe fn dense_download(d: *const ffi::bsm_dense, p: *const *mut Self) -> i32 { ffi::$ddown(d, p) }
            unsafe fn mul_vector(a: *const ffi::bsm_csr, x: *const Self, n: u64, y: *mut Self, m: u64) -> i32 { ffi::$mv(a, x, n, y, m) }
            unsafe fn mul_dense_host_into(r: u64, c: u64, n: u64, v: *const Self, ci: *const u64, ri: *const u64, l: u64, br: u64, bc: u64,
                                          bp: *const *const Self, algo: i32, cap: u64, ov: *mut Self, oc: *mut u64, or: *mut u64,
                                          on: *mut u64) -> i32 { ffi::$into(r, c, n, v, ci, ri, l, br, bc, bp, algo, cap, ov, oc, or, on) }
        }
    };
}
impl_scalar!(f64, ffi::BSM_F64, bsm_csr_upload_f64, bsm_csr_download_f64, bsm_dense_upload_f64, bsm_dense_download_f64, bsm_mul_vector_f64, bsm_mul_dense_host_into_f64);
impl_scalar!(f32, ffi::BSM_F32, bsm_csr_upload_f32, bsm_csr_download_f32, bsm_dense_upload_f32, bsm_dense_download_f32, bsm_mul_vector_f32, bsm_mul_dense_host_into_f32);

#[derive(Clone, Copy, Debug, PartialEq)]
pub enum Algo { Auto = 0, VectorCsr = 1, MergePath = 2, RowBlock = 3 }

/// Select the GPU of this process (one process per GPU).
pub fn init(device: i32) -> Result<(), GpuError> { check(unsafe { ffi::bsm_init(device) }) }

pub struct DeviceCsr<T: GpuScalar> { h: *mut ffi::bsm_csr, dims: MatDim, _t: PhantomData<T> }
pub struct DeviceDense<T: GpuScalar> { h: *mut ffi::bsm_dense, dims: MatDim, _t: PhantomData<T> }

impl<T: GpuScalar> Drop for DeviceCsr<T> { fn drop(&mut self) { unsafe { ffi::bsm_csr_free(self.h); } } }
impl<T: GpuScalar> Drop for DeviceDense<T> { fn drop(&mut self) { unsafe { ffi::bsm_dense_free(self.h); } } }
impl<T: GpuScalar> GetDims for DeviceCsr<T> { fn get_dims(&self) -> MatDim { self.dims } }
impl<T: GpuScalar> GetDims for DeviceDense<T> { fn get_dims(&self) -> MatDim { self.dims } }

impl<T: GpuScalar> DeviceCsr<T> {
    /// Upload a finalised Csr (usize indices are narrowed to u32 on the device).
    pub fn from_host(m: &Csr<T>) -> Result<Self, GpuError> {
        if !m.is_finalised() { return Err(GpuError::Mat(MatErr::MatrixNotFinalised)); }
        let (v, ci, ri) = m.raw_parts();
        let dims = m.get_dims();
        let mut h = ptr::null_mut();
        check(unsafe { T::csr_upload(dims.rows as u64, dims.cols as u64, v.len() as u64, v.as_ptr(),
                                     ci.as_ptr() as *const u64, ri.as_ptr() as *const u64, ri.len() as u64, &mut h) })?;
        Ok(Self { h, dims, _t: PhantomData })
    }
    pub fn get_nnz(&self) -> usize {
        let mut nnz = 0u64;
        unsafe { ffi::bsm_csr_info(self.h, ptr::null_mut(), ptr::null_mut(), ptr::null_mut(), &mut nnz, ptr::null_mut()); }
        nnz as usize
    }
    pub fn to_host(&self) -> Result<Csr<T>, GpuError> {
        let nnz = self.get_nnz();
        let mut v = vec![T::default(); nnz];
        let mut ci = vec![0usize; nnz];
        let mut ri = vec![0usize; self.dims.rows + 1];
        check(unsafe { T::csr_download(self.h, v.as_mut_ptr(), ci.as_mut_ptr() as *mut u64, ri.as_mut_ptr() as *mut u64) })?;
        Ok(Csr::from_raw_parts(self.dims, v, ci, ri))
    }
    /// `Csr::mul_dense` on device-resident operands (src/sparse.rs:426-446); the product stays in HBM.
    pub fn mul_dense(&self, rhs: &DeviceDense<T>) -> Result<DeviceDense<T>, GpuError> { self.mul_dense_with(rhs, Algo::Auto) }
    pub fn mul_dense_with(&self, rhs: &DeviceDense<T>, algo: Algo) -> Result<DeviceDense<T>, GpuError> {
        if self.dims.cols != rhs.dims.rows { return Err(GpuError::Mat(MatErr::IncorrectDimensions)); } // sparse.rs:427-429
        let out = DeviceDense::<T>::alloc(self.dims.rows, rhs.dims.cols)?;
        check(unsafe { ffi::bsm_spmm(self.h, rhs.h, out.h, algo as i32) })?;
        Ok(out)
    }
    /// `Csr::mul_vector` (src/sparse.rs:468-482): host slices in/out, dense result.
    pub fn mul_vector(&self, rhs: &[T], out: &mut [T]) -> Result<(), GpuError> {
        check(unsafe { T::mul_vector(self.h, rhs.as_ptr(), rhs.len() as u64, out.as_mut_ptr(), out.len() as u64) })
    }
}

impl<T: GpuScalar> DeviceDense<T> {
    pub fn alloc(rows: usize, cols: usize) -> Result<Self, GpuError> {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::bsm_dense_alloc(T::DTYPE, rows as u64, cols as u64, &mut h) })?;
        Ok(Self { h, dims: MatDim { rows, cols }, _t: PhantomData })
    }
    /// Upload the column-major Vec<Vec<T>> (dense.rs:5-9); transposed to row-major on the device.
    pub fn from_host(d: &Dense<T>) -> Result<Self, GpuError> {
        let dims = d.get_dims();
        let ptrs: Vec<*const T> = d.columns().iter().map(|c| c.as_ptr()).collect();
        let mut h = ptr::null_mut();
        check(unsafe { T::dense_upload(dims.rows as u64, dims.cols as u64, ptrs.as_ptr(), &mut h) })?;
        Ok(Self { h, dims, _t: PhantomData })
    }
    pub fn to_host(&self) -> Result<Dense<T>, GpuError> where T: Clone {
        let mut d = Dense::<T>::new_default_with_dims(self.dims.cols, self.dims.rows);   // (cols, rows): dense.rs:13
        let ptrs: Vec<*mut T> = d.columns_mut().iter_mut().map(|c| c.as_mut_ptr()).collect();
        check(unsafe { T::dense_download(self.h, ptrs.as_ptr()) })?;
        Ok(d)
    }
    /// The reference's result construction: every output through `insert` (zero-drop), then `finalise`.
    pub fn into_csr(&self) -> Result<DeviceCsr<T>, GpuError> {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::bsm_dense_to_csr(self.h, &mut h) })?;
        Ok(DeviceCsr { h, dims: self.dims, _t: PhantomData })
    }
}

/// The literal drop-in for `Csr::mul_dense(&self, rhs:&Dense<T>) -> Result<Csr<T>,MatErr>` (sparse.rs:426-446): host Csr and
/// host Dense in, zero-dropped finalised host Csr out, through ONE pipelined C-ABI call (`bsm_mul_dense_host_into_*`: B up
/// in row chunks, SpMM + result construction per block of rows, values / usize columns / row_index down, overlapped).
/// The result Vecs are sized for the worst case (every output non-zero; untouched capacity costs no pages) and trimmed.
impl<T: GpuScalar> Csr<T> {
    pub fn mul_dense_gpu(&self, rhs: &Dense<T>) -> Result<Csr<T>, GpuError> {
        let (a, b) = (self.get_dims(), rhs.get_dims());
        if a.cols != b.rows { return Err(GpuError::Mat(MatErr::IncorrectDimensions)); }          // sparse.rs:427-429
        if !self.is_finalised() { return Err(GpuError::Mat(MatErr::MatrixNotFinalised)); }
        let (v, ci, ri) = self.raw_parts();
        let cols: Vec<*const T> = rhs.columns().iter().map(|c| c.as_ptr()).collect();
        let cap = a.rows * b.cols;
        let (mut ov, mut oc, mut or) = (Vec::<T>::with_capacity(cap), Vec::<usize>::with_capacity(cap), vec![0usize; a.rows + 1]);
        let mut nnz = 0u64;
        check(unsafe { T::mul_dense_host_into(a.rows as u64, a.cols as u64, v.len() as u64, v.as_ptr(), ci.as_ptr() as *const u64,
                                              ri.as_ptr() as *const u64, ri.len() as u64, b.rows as u64, b.cols as u64, cols.as_ptr(),
                                              Algo::Auto as i32, cap as u64, ov.as_mut_ptr(), oc.as_mut_ptr() as *mut u64,
                                              or.as_mut_ptr() as *mut u64, &mut nnz) })?;
        unsafe { ov.set_len(nnz as usize); oc.set_len(nnz as usize); }
        ov.shrink_to_fit();
        oc.shrink_to_fit();
        Ok(Csr::from_raw_parts((a.rows, b.cols), ov, oc, or))
    }
}

/// The substitution half of `solve` (lib.rs:11-24) on device-resident operands: `forward_substitution(l, b)` (lib.rs:28-46)
/// and `backward_substitution(l_star, y)` (lib.rs:49-65), bit-identical to the reference's loops; the factorisation
/// (`cholesky_decomp`, `transpose`) stays on the CPU.
impl<T: GpuScalar> DeviceCsr<T> {
    pub fn forward_substitution(&self, b: &DeviceDense<T>) -> Result<DeviceDense<T>, GpuError> {
        let y = DeviceDense::<T>::alloc(b.dims.rows, b.dims.cols)?;
        check(unsafe { ffi::bsm_forward_substitution(self.h, b.h, y.h) })?;
        Ok(y)
    }
    pub fn backward_substitution(&self, y: &DeviceDense<T>) -> Result<DeviceDense<T>, GpuError> {
        let x = DeviceDense::<T>::alloc(y.dims.rows, y.dims.cols)?;
        check(unsafe { ffi::bsm_backward_substitution(self.h, y.h, x.h) })?;
        Ok(x)
    }
    /// Half-bandwidths `(lower, upper)` when the matrix is a proper lower / upper band factor (what `cholesky_decomp` and
    /// `transpose` return for a band matrix): such factors run the specialised substitution kernels. `None` otherwise.
    pub fn band_structure(&self) -> Result<(Option<usize>, Option<usize>), GpuError> {
        let (mut lo, mut up) = (-1i32, -1i32);
        check(unsafe { ffi::bsm_csr_band_structure(self.h, &mut lo, &mut up) })?;
        let opt = |v: i32| if v >= 0 { Some(v as usize) } else { None };
        Ok((opt(lo), opt(up)))
    }
}

/// nnz-balanced contiguous row split for the row-partitioned multi-GPU path (B replicated).
pub fn partition_rows<T: GpuScalar>(m: &Csr<T>, parts: usize) -> Result<Vec<usize>, GpuError> {
    let (_, _, ri) = m.raw_parts();
    let mut bounds = vec![0usize; parts + 1];
    check(unsafe { ffi::bsm_partition_rows(ri.as_ptr() as *const u64, m.get_dims().rows as u64, parts as i32,
                                           bounds.as_mut_ptr() as *mut u64) })?;
    Ok(bounds)
}
