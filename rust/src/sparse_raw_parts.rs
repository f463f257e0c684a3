// Additions to src/sparse.rs and src/dense.rs of the reference crate. The fields of Csr and Dense
// are private to their modules (sparse.rs:69-78, dense.rs:5-9), so the sibling `gpu` module needs
// crate-visible raw accessors; uploading through get_row_compact would allocate 24*cols bytes per
// row (sparse.rs:254).

// ---- in impl<T: Copy+Default+PartialEq+Debug> Csr<T> (src/sparse.rs) ----
impl<T: Copy + Default + PartialEq + std::fmt::Debug> Csr<T> {
    /// (v, col_index, row_index); row_index has rows+1 entries once finalised (sparse.rs:206-219).
    pub(crate) fn raw_parts(&self) -> (&[T], &[usize], &[usize]) {
        (&self.v, &self.col_index, &self.row_index)
    }
    pub(crate) fn is_finalised(&self) -> bool { self.is_finalised }
    /// A finalised Csr from arrays produced by the device-side zero-dropping compaction; iterator
    /// cursors at 0 so that derived PartialEq (sparse.rs:68) matches a CPU-built result.
    pub(crate) fn from_raw_parts<D: Into<MatDim>>(dims: D, v: Vec<T>, col_index: Vec<usize>, row_index: Vec<usize>) -> Self {
        Self { dims: dims.into(), v, col_index, row_index, is_finalised: true, iter_v_index: 0, iter_row_index: 0 }
    }
}

// ---- in impl<T> Dense<T> (src/dense.rs) ----
impl<T> Dense<T> {
    /// The columns as they lie in memory (data[c][r], dense.rs:5-9).
    pub(crate) fn columns(&self) -> &[Vec<T>] { &self.data }
    pub(crate) fn columns_mut(&mut self) -> &mut [Vec<T>] { &mut self.data }
}
