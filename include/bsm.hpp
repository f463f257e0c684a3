// bsm.hpp — C++ host mirror of the reference crate's hot-path surface over the C ABI (bsm.h).
//
// The reference is Rust (crate `sparse_matrix`); no Rust toolchain exists in the build image, so the
// host side above the C ABI is written in C++ with the SAME names, argument order and error
// behaviour as the reference, so that tests read like the reference's own:
//
//     reference (Rust)                                     here (C++)
//     util::MatDim / MatErr / GetDims   util.rs:11-55      sparse_matrix::MatDim / MatErr / get_dims()
//     dense::Dense<T>                   dense.rs:4-47      sparse_matrix::Dense<T>
//     dense_static::DenseS<T,R,C>       dense_static.rs    sparse_matrix::DenseS<T,R,C>
//     sparse::Csr<T>, CsrEntry          sparse.rs:68-265   sparse_matrix::Csr<T>, CsrEntry<T>
//     Csr::mul_dense                    sparse.rs:426-446  Csr<T>::mul_dense  (runs on the B200)
//     Csr::mul_dense_s                  sparse.rs:448-466  Csr<T>::mul_dense_s (runs on the B200)
//     Csr::mul_vector                   sparse.rs:468-482  Csr<T>::mul_vector (runs on the B200)
//     new `gpu` module                    sparse_matrix::gpu::DeviceCsr<T> / DeviceDense<T>
//
// `Result<T>` plays the role of Rust's `Result<T, MatErr>`: `is_ok()`, `unwrap()`, `unwrap_err()`.
// Failures that are not a MatErr (CUDA, NCCL, no device, index overflow) throw GpuError — the
// reference's MatErr is an exhaustive enum and gets no new variants.
// There is no CPU implementation of the multiplications here: without libbsm_b200 + a GPU they fail.
#pragma once

#include <array>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "bsm.h"

namespace sparse_matrix {

// ---- util.rs ------------------------------------------------------------------------------------
struct MatDim {   // util.rs:11-15
    std::size_t rows = 0, cols = 0;
    MatDim() = default;
    MatDim(std::size_t r, std::size_t c) : rows(r), cols(c) {}                   // From<(usize,usize)>  util.rs:23-27
    MatDim transpose() const { return MatDim(cols, rows); }                      // util.rs:17-21
    bool operator==(const MatDim &o) const { return rows == o.rows && cols == o.cols; }
    bool operator!=(const MatDim &o) const { return !(*this == o); }
};

enum class MatErr {   // util.rs:47-55 — variants unchanged
    MatrixFinalised,
    MatrixNotFinalised,
    NonSquareMatrix,
    IncorrectDimensions,
    PaddingSizeSmallerThanOriginal,
    OutOfBounds
};

struct GpuError : std::runtime_error {
    int status;
    GpuError(int st, const std::string &msg) : std::runtime_error("bsm status " + std::to_string(st) + ": " + msg), status(st) {}
};

template <typename T> class Result {   // Result<T, MatErr>
    bool ok_;
    T value_;
    MatErr err_;

  public:
    Result(T v) : ok_(true), value_(std::move(v)), err_(MatErr::OutOfBounds) {}
    Result(MatErr e) : ok_(false), value_(), err_(e) {}
    bool is_ok() const { return ok_; }
    bool is_err() const { return !ok_; }
    T unwrap()
    {
        if (!ok_) throw std::logic_error("called unwrap() on an Err value");   // Rust: panic
        return std::move(value_);
    }
    MatErr unwrap_err() const
    {
        if (ok_) throw std::logic_error("called unwrap_err() on an Ok value");
        return err_;
    }
    bool operator==(const Result &o) const { return ok_ == o.ok_ && (ok_ ? value_ == o.value_ : err_ == o.err_); }
};
template <> class Result<void> {
    bool ok_;
    MatErr err_;

  public:
    Result() : ok_(true), err_(MatErr::OutOfBounds) {}
    Result(MatErr e) : ok_(false), err_(e) {}
    bool is_ok() const { return ok_; }
    bool is_err() const { return !ok_; }
    void unwrap() const
    {
        if (!ok_) throw std::logic_error("called unwrap() on an Err value");
    }
    MatErr unwrap_err() const
    {
        if (ok_) throw std::logic_error("called unwrap_err() on an Ok value");
        return err_;
    }
    bool operator==(const Result &o) const { return ok_ == o.ok_ && (ok_ || err_ == o.err_); }
};
inline Result<void> Ok() { return Result<void>(); }
inline Result<void> Err(MatErr e) { return Result<void>(e); }

namespace detail {
template <typename T> struct Abi;
template <> struct Abi<double> {
    static constexpr int dtype = BSM_F64;
    static int csr_upload(uint64_t r, uint64_t c, uint64_t n, const double *v, const uint64_t *ci, const uint64_t *ri, uint64_t l, bsm_csr **o) { return bsm_csr_upload_f64(r, c, n, v, ci, ri, l, o); }
    static int csr_download(const bsm_csr *a, double *v, uint64_t *ci, uint64_t *ri) { return bsm_csr_download_f64(a, v, ci, ri); }
    static int dense_upload(uint64_t r, uint64_t c, const double *const *p, bsm_dense **o) { return bsm_dense_upload_f64(r, c, p, o); }
    static int dense_download(const bsm_dense *d, double *const *p) { return bsm_dense_download_f64(d, p); }
    static int mul_vector(const bsm_csr *a, const double *x, uint64_t n, double *y, uint64_t m) { return bsm_mul_vector_f64(a, x, n, y, m); }
    static int host_dense(uint64_t r, uint64_t c, uint64_t n, const double *v, const uint64_t *ci, const uint64_t *ri, uint64_t l, uint64_t br, uint64_t bc,
                          const double *const *b, double *const *o) { return bsm_mul_dense_host_dense_f64(r, c, n, v, ci, ri, l, br, bc, b, o, BSM_ALGO_AUTO); }
};
template <> struct Abi<float> {
    static constexpr int dtype = BSM_F32;
    static int csr_upload(uint64_t r, uint64_t c, uint64_t n, const float *v, const uint64_t *ci, const uint64_t *ri, uint64_t l, bsm_csr **o) { return bsm_csr_upload_f32(r, c, n, v, ci, ri, l, o); }
    static int csr_download(const bsm_csr *a, float *v, uint64_t *ci, uint64_t *ri) { return bsm_csr_download_f32(a, v, ci, ri); }
    static int dense_upload(uint64_t r, uint64_t c, const float *const *p, bsm_dense **o) { return bsm_dense_upload_f32(r, c, p, o); }
    static int dense_download(const bsm_dense *d, float *const *p) { return bsm_dense_download_f32(d, p); }
    static int mul_vector(const bsm_csr *a, const float *x, uint64_t n, float *y, uint64_t m) { return bsm_mul_vector_f32(a, x, n, y, m); }
    static int host_dense(uint64_t r, uint64_t c, uint64_t n, const float *v, const uint64_t *ci, const uint64_t *ri, uint64_t l, uint64_t br, uint64_t bc,
                          const float *const *b, float *const *o) { return bsm_mul_dense_host_dense_f32(r, c, n, v, ci, ri, l, br, bc, b, o, BSM_ALGO_AUTO); }
};
// status -> MatErr where the reference has a matching variant, else throw
inline bool status_to_materr(int st, MatErr *e)
{
    switch (st) {
        case BSM_ERR_INCORRECT_DIMENSIONS: *e = MatErr::IncorrectDimensions; return true;
        case BSM_ERR_NOT_FINALISED: *e = MatErr::MatrixNotFinalised; return true;
        case BSM_ERR_OUT_OF_BOUNDS: *e = MatErr::OutOfBounds; return true;
    }
    return false;
}
inline void throw_status(int st) { throw GpuError(st, bsm_last_error_string()); }
static_assert(sizeof(std::size_t) == sizeof(uint64_t), "usize must be 64-bit (the C ABI takes uint64_t indices)");
}  // namespace detail

// ---- dense.rs -------------------------------------------------------------------------------------
template <typename T> class Dense {   // dense.rs:4-9 — COLUMN-major Vec<Vec<T>>: data[c][r]
    std::size_t col_count_ = 0, row_count_ = 0;
    std::vector<std::vector<T>> data_;

  public:
    Dense() = default;
    static Dense new_default_with_dims(std::size_t col_count, std::size_t row_count)   // dense.rs:13-15 — (cols, rows)
    {
        return new_with_dims(T(), col_count, row_count);
    }
    static Dense new_with_dims(T val, std::size_t col_count, std::size_t row_count)     // dense.rs:17-19
    {
        Dense d;
        d.col_count_ = col_count;
        d.row_count_ = row_count;
        d.data_.assign(col_count, std::vector<T>(row_count, val));
        return d;
    }
    static Dense from_data(const std::vector<std::vector<T>> &columns)                  // dense.rs:21-29 — data[c] is COLUMN c
    {
        Dense d;
        d.col_count_ = columns.size();
        d.row_count_ = columns.empty() ? 0 : columns[0].size();
        d.data_ = columns;
        return d;
    }
    const std::vector<T> &get_col(std::size_t col_index) const { return data_.at(col_index); }   // dense.rs:31-33
    std::vector<T> &get_col_mut(std::size_t col_index) { return data_.at(col_index); }           // dense.rs:35-37
    MatDim get_dims() const { return MatDim(row_count_, col_count_); }                           // dense.rs:40-47
    bool operator==(const Dense &o) const { return col_count_ == o.col_count_ && row_count_ == o.row_count_ && data_ == o.data_; }
    const std::vector<std::vector<T>> &columns() const { return data_; }
    std::vector<std::vector<T>> &columns_mut() { return data_; }
};

// ---- dense_static.rs --------------------------------------------------------------------------------
template <typename T, std::size_t ROWS, std::size_t COLS> class DenseS {   // dense_static.rs:4-9 — [[T; ROWS]; COLS], column-major
    std::array<std::array<T, ROWS>, COLS> data_;

  public:
    static DenseS new_default() { return new_(T()); }                            // dense_static.rs:13-15
    static DenseS new_(T val)                                                    // dense_static.rs:17-19 (`new` is a C++ keyword)
    {
        DenseS d;
        for (auto &c : d.data_) c.fill(val);
        return d;
    }
    static DenseS from_data(const std::vector<std::vector<T>> &columns)          // dense_static.rs:21-35 — data[c] is COLUMN c
    {
        DenseS d = new_default();
        for (std::size_t i = 0; i < COLS; ++i)
            for (std::size_t j = 0; j < ROWS; ++j) d.data_[i][j] = columns.at(i).at(j);
        return d;
    }
    const std::array<T, ROWS> &get_col(std::size_t col_index) const { return data_.at(col_index); }   // dense_static.rs:37-39
    std::array<T, ROWS> &get_col_mut(std::size_t col_index) { return data_.at(col_index); }           // dense_static.rs:41-43
    MatDim get_dims() const { return MatDim(ROWS, COLS); }                                            // dense_static.rs:46-53
    bool operator==(const DenseS &o) const { return data_ == o.data_; }
};

// ---- sparse.rs ------------------------------------------------------------------------------------
template <typename T> struct CsrEntry {   // sparse.rs:80-85
    T v;
    std::size_t col_index, row_index;
    bool operator==(const CsrEntry &o) const { return v == o.v && col_index == o.col_index && row_index == o.row_index; }
};

namespace gpu {
template <typename T> class DeviceCsr;
template <typename T> class DeviceDense;
}  // namespace gpu

template <typename T> class Csr {   // sparse.rs:68-78
    MatDim dims_;
    std::vector<T> v_;
    std::vector<std::size_t> col_index_;
    std::vector<std::size_t> row_index_;
    bool is_finalised_ = false;
    std::size_t iter_v_index_ = 0, iter_row_index_ = 0;

    void insert_unchecked(T value, std::size_t row, std::size_t col)   // sparse.rs:237-250
    {
        v_.push_back(value);
        col_index_.push_back(col);
        if (row > row_index_.size() - 1) {
            if (row > row_index_.size()) {
                row_index_.push_back(v_.size() - 1);
                const std::size_t from = row_index_.size();
                for (std::size_t i = from; i < row + 1; ++i) row_index_.push_back(row_index_.back());
            } else {
                row_index_.push_back(v_.size() - 1);
            }
        }
    }

  public:
    Csr() : row_index_{0} {}
    static Csr new_(MatDim dims) { return new_with_capacity(dims, 0); }              // sparse.rs:117-119 (`new` is a C++ keyword)
    static Csr new_with_capacity(MatDim dims, std::size_t capacity)                   // sparse.rs:121-132
    {
        Csr m;
        m.dims_ = dims;
        m.v_.reserve(capacity);
        m.col_index_.reserve(capacity);
        return m;
    }
    static Csr from_data(const std::vector<std::vector<T>> &rows)                     // sparse.rs:193-203 — data[r] is ROW r
    {
        Csr m = new_(MatDim(rows.size(), rows.empty() ? 0 : rows[0].size()));
        for (std::size_t i = 0; i < rows.size(); ++i)
            for (std::size_t j = 0; j < rows[i].size(); ++j) m.insert(rows[i][j], i, j).unwrap();
        return std::move(m).finalise();
    }
    // the `pub(crate)` raw constructor the gpu module needs (fields are private in the reference)
    static Csr from_raw_parts(MatDim dims, std::vector<T> v, std::vector<std::size_t> col_index, std::vector<std::size_t> row_index)
    {
        Csr m;
        m.dims_ = dims;
        m.v_ = std::move(v);
        m.col_index_ = std::move(col_index);
        m.row_index_ = std::move(row_index);
        m.is_finalised_ = true;
        return m;
    }
    Result<void> insert(T value, std::size_t row, std::size_t col)                    // sparse.rs:222-233
    {
        if (is_finalised_) return Err(MatErr::MatrixFinalised);
        if (value != T()) insert_unchecked(value, row, col);   // T::default() skipped (-0.0 too); NaN kept
        return Ok();
    }
    Csr finalise() &&                                                                 // sparse.rs:206-219 — by value, chainable
    {
        if (!is_finalised_) {
            is_finalised_ = true;
            if (dims_.rows < row_index_.size()) throw std::logic_error("big eek");    // panic!("big eek")
            const std::size_t required_spacers = dims_.rows - row_index_.size();
            for (std::size_t i = 0; i < required_spacers; ++i) row_index_.push_back(v_.size());
            row_index_.push_back(v_.size());
        }
        return std::move(*this);
    }
    Csr finalise() const & { return Csr(*this).finalise(); }
    std::size_t get_nnz() const { return row_index_.empty() ? 0 : row_index_.back(); }                 // sparse.rs:162-164
    float get_density() const { return (float)v_.size() / (float)(dims_.rows * dims_.cols); }          // sparse.rs:166-168
    MatDim get_dims() const { return dims_; }                                                          // sparse.rs:418-422
    bool is_finalised() const { return is_finalised_; }
    std::vector<CsrEntry<T>> get_row_compact(std::size_t index) const                                  // sparse.rs:252-265
    {
        std::vector<CsrEntry<T>> row;
        const std::size_t row_start = row_index_.at(index);
        const std::size_t row_end = index == row_index_.size() - 1 ? v_.size() : row_index_.at(index + 1);
        for (std::size_t e = row_start; e < row_end; ++e) row.push_back(CsrEntry<T>{v_[e], col_index_[e], index});
        return row;
    }
    const std::vector<T> &raw_v() const { return v_; }
    const std::vector<std::size_t> &raw_col_index() const { return col_index_; }
    const std::vector<std::size_t> &raw_row_index() const { return row_index_; }
    bool operator==(const Csr &o) const   // #[derive(PartialEq)] sparse.rs:68 — every field
    {
        return dims_ == o.dims_ && v_ == o.v_ && col_index_ == o.col_index_ && row_index_ == o.row_index_ &&
               is_finalised_ == o.is_finalised_ && iter_v_index_ == o.iter_v_index_ && iter_row_index_ == o.iter_row_index_;
    }

    // ---- the hot path: Csr::mul_dense(&self, rhs:&Dense<T>) -> Result<Csr<T>,MatErr>  sparse.rs:426-446 ----
    Result<Csr> mul_dense(const Dense<T> &rhs) const;
    // same product with a DENSE result in the reference's column-major layout (no zero-drop): one pipelined
    // host-to-host call (H2D | multiply | D2H overlapped per column group)
    Result<Dense<T>> mul_dense_into_dense(const Dense<T> &rhs) const;
    // Csr::mul_dense_s(&self, rhs:&DenseS<T,ROWS,COLS>) -> Result<Csr<T>,MatErr>  sparse.rs:448-466 — the same
    // loop nest against the stack-array operand; same kernels
    template <std::size_t ROWS, std::size_t COLS> Result<Csr> mul_dense_s(const DenseS<T, ROWS, COLS> &rhs) const
    {
        if (dims_.cols != ROWS) return Result<Csr>(MatErr::IncorrectDimensions);   // sparse.rs:449
        std::vector<std::vector<T>> cols;
        for (std::size_t c = 0; c < COLS; ++c) cols.emplace_back(rhs.get_col(c).begin(), rhs.get_col(c).end());
        return mul_dense(Dense<T>::from_data(cols));
    }
    // Csr::mul_vector(&self, rhs:&[T], out:&mut [T]) -> Result<(),MatErr>  sparse.rs:468-482
    Result<void> mul_vector(const std::vector<T> &rhs, std::vector<T> &out) const;
};

// ---- the new `gpu` module: device-resident operands ------------------------------------------------------
namespace gpu {

inline void init(int device)
{
    const int st = bsm_init(device);
    if (st != BSM_OK) detail::throw_status(st);
}
enum class Algo { Auto = BSM_ALGO_AUTO, VectorCsr = BSM_ALGO_VECTOR, MergePath = BSM_ALGO_MERGE, RowBlock = BSM_ALGO_ROWBLOCK };

template <typename T> class DeviceDense {
    bsm_dense *h_ = nullptr;
    MatDim dims_;
    friend class DeviceCsr<T>;

  public:
    DeviceDense() = default;
    DeviceDense(bsm_dense *h, MatDim d) : h_(h), dims_(d) {}
    DeviceDense(const DeviceDense &) = delete;
    DeviceDense &operator=(const DeviceDense &) = delete;
    DeviceDense(DeviceDense &&o) noexcept : h_(o.h_), dims_(o.dims_) { o.h_ = nullptr; }
    DeviceDense &operator=(DeviceDense &&o) noexcept
    {
        if (this != &o) {
            bsm_dense_free(h_);
            h_ = o.h_;
            dims_ = o.dims_;
            o.h_ = nullptr;
        }
        return *this;
    }
    ~DeviceDense() { bsm_dense_free(h_); }   // Drop
    static DeviceDense alloc(std::size_t rows, std::size_t cols)
    {
        bsm_dense *h = nullptr;
        const int st = bsm_dense_alloc(detail::Abi<T>::dtype, rows, cols, &h);
        if (st != BSM_OK) detail::throw_status(st);
        return DeviceDense(h, MatDim(rows, cols));
    }
    static DeviceDense from_host(const Dense<T> &d)   // column-major Vec<Vec<T>> -> row-major HBM (device transpose)
    {
        std::vector<const T *> ptrs;
        for (const auto &c : d.columns()) ptrs.push_back(c.data());
        bsm_dense *h = nullptr;
        const MatDim dims = d.get_dims();
        const int st = detail::Abi<T>::dense_upload(dims.rows, dims.cols, ptrs.data(), &h);
        if (st != BSM_OK) detail::throw_status(st);
        return DeviceDense(h, dims);
    }
    Dense<T> to_host() const
    {
        Dense<T> d = Dense<T>::new_default_with_dims(dims_.cols, dims_.rows);
        std::vector<T *> ptrs;
        for (auto &c : d.columns_mut()) ptrs.push_back(c.data());
        const int st = detail::Abi<T>::dense_download(h_, ptrs.data());
        if (st != BSM_OK) detail::throw_status(st);
        return d;
    }
    DeviceCsr<T> into_csr() const;   // zero-dropping insert + finalise on the device (sparse.rs:442, 222-233, 206-219)
    MatDim get_dims() const { return dims_; }
    bsm_dense *handle() const { return h_; }
};

template <typename T> class DeviceCsr {
    bsm_csr *h_ = nullptr;
    MatDim dims_;

  public:
    DeviceCsr() = default;
    DeviceCsr(bsm_csr *h, MatDim d) : h_(h), dims_(d) {}
    DeviceCsr(const DeviceCsr &) = delete;
    DeviceCsr &operator=(const DeviceCsr &) = delete;
    DeviceCsr(DeviceCsr &&o) noexcept : h_(o.h_), dims_(o.dims_) { o.h_ = nullptr; }
    ~DeviceCsr() { bsm_csr_free(h_); }   // Drop
    static Result<DeviceCsr> from_host(const Csr<T> &m)
    {
        bsm_csr *h = nullptr;
        const MatDim dims = m.get_dims();
        const int st = detail::Abi<T>::csr_upload(dims.rows, dims.cols, m.raw_v().size(), m.raw_v().data(),
                                                  reinterpret_cast<const uint64_t *>(m.raw_col_index().data()),
                                                  reinterpret_cast<const uint64_t *>(m.raw_row_index().data()),
                                                  m.raw_row_index().size(), &h);
        MatErr e;
        if (st != BSM_OK) {
            if (detail::status_to_materr(st, &e)) return Result<DeviceCsr>(e);
            detail::throw_status(st);
        }
        return Result<DeviceCsr>(DeviceCsr(h, dims));
    }
    std::size_t get_nnz() const
    {
        uint64_t nnz = 0;
        bsm_csr_info(h_, nullptr, nullptr, nullptr, &nnz, nullptr);
        return (std::size_t)nnz;
    }
    Csr<T> to_host() const
    {
        const std::size_t nnz = get_nnz();
        std::vector<T> v(nnz);
        std::vector<std::size_t> ci(nnz), ri(dims_.rows + 1);
        const int st = detail::Abi<T>::csr_download(h_, v.data(), reinterpret_cast<uint64_t *>(ci.data()), reinterpret_cast<uint64_t *>(ri.data()));
        if (st != BSM_OK) detail::throw_status(st);
        return Csr<T>::from_raw_parts(dims_, std::move(v), std::move(ci), std::move(ri));
    }
    // mul_dense on device-resident operands; the dense product stays in HBM
    Result<DeviceDense<T>> mul_dense(const DeviceDense<T> &rhs, Algo algo = Algo::Auto) const
    {
        if (dims_.cols != rhs.dims_.rows) return Result<DeviceDense<T>>(MatErr::IncorrectDimensions);   // sparse.rs:427-429
        DeviceDense<T> out = DeviceDense<T>::alloc(dims_.rows, rhs.dims_.cols);
        const int st = bsm_spmm(h_, rhs.h_, out.h_, (int)algo);
        MatErr e;
        if (st != BSM_OK) {
            if (detail::status_to_materr(st, &e)) return Result<DeviceDense<T>>(e);
            detail::throw_status(st);
        }
        return Result<DeviceDense<T>>(std::move(out));
    }
    Result<void> mul_vector(const std::vector<T> &rhs, std::vector<T> &out) const
    {
        const int st = detail::Abi<T>::mul_vector(h_, rhs.data(), rhs.size(), out.data(), out.size());
        MatErr e;
        if (st != BSM_OK) {
            if (detail::status_to_materr(st, &e)) return Err(e);
            detail::throw_status(st);
        }
        return Ok();
    }
    MatDim get_dims() const { return dims_; }
    bsm_csr *handle() const { return h_; }
};

template <typename T> DeviceCsr<T> DeviceDense<T>::into_csr() const
{
    bsm_csr *h = nullptr;
    const int st = bsm_dense_to_csr(h_, &h);
    if (st != BSM_OK) detail::throw_status(st);
    return DeviceCsr<T>(h, dims_);
}

}  // namespace gpu

template <typename T> Result<Csr<T>> Csr<T>::mul_dense(const Dense<T> &rhs) const
{
    static_assert(std::is_same<T, float>::value || std::is_same<T, double>::value, "the GPU path computes in f32 or f64");
    if (dims_.cols != rhs.get_dims().rows) return Result<Csr<T>>(MatErr::IncorrectDimensions);   // sparse.rs:427-429
    auto a = gpu::DeviceCsr<T>::from_host(*this);
    if (a.is_err()) return Result<Csr<T>>(a.unwrap_err());
    gpu::DeviceCsr<T> da = a.unwrap();
    gpu::DeviceDense<T> db = gpu::DeviceDense<T>::from_host(rhs);
    auto c = da.mul_dense(db);
    if (c.is_err()) return Result<Csr<T>>(c.unwrap_err());
    return Result<Csr<T>>(c.unwrap().into_csr().to_host());
}

template <typename T> Result<Dense<T>> Csr<T>::mul_dense_into_dense(const Dense<T> &rhs) const
{
    const MatDim bd = rhs.get_dims();
    if (dims_.cols != bd.rows) return Result<Dense<T>>(MatErr::IncorrectDimensions);   // sparse.rs:427-429
    Dense<T> out = Dense<T>::new_default_with_dims(bd.cols, dims_.rows);
    std::vector<const T *> in_ptrs;
    std::vector<T *> out_ptrs;
    for (const auto &c : rhs.columns()) in_ptrs.push_back(c.data());
    for (auto &c : out.columns_mut()) out_ptrs.push_back(c.data());
    const int st = detail::Abi<T>::host_dense(dims_.rows, dims_.cols, v_.size(), v_.data(), reinterpret_cast<const uint64_t *>(col_index_.data()),
                                              reinterpret_cast<const uint64_t *>(row_index_.data()), row_index_.size(), bd.rows, bd.cols,
                                              in_ptrs.data(), out_ptrs.data());
    MatErr e;
    if (st != BSM_OK) {
        if (detail::status_to_materr(st, &e)) return Result<Dense<T>>(e);
        detail::throw_status(st);
    }
    return Result<Dense<T>>(std::move(out));
}

template <typename T> Result<void> Csr<T>::mul_vector(const std::vector<T> &rhs, std::vector<T> &out) const
{
    if (dims_.cols != rhs.size() || dims_.rows != out.size()) return Err(MatErr::IncorrectDimensions);   // sparse.rs:469-471
    auto a = gpu::DeviceCsr<T>::from_host(*this);
    if (a.is_err()) return Err(a.unwrap_err());
    return a.unwrap().mul_vector(rhs, out);
}

}  // namespace sparse_matrix
