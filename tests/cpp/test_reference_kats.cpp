// C++ rendition of the reference's own unit tests for the hot path, written against the C++ host
// mirror (include/bsm.hpp) so that they read like the originals:
//   test_dense_mul   src/sparse.rs:1082-1109      test_nnz         src/sparse.rs:1153-1178
//   test_mul_vector  src/sparse.rs:1501-1529      structure KATs   src/sparse.rs:815-868, 1111-1151
//   dense init/get_col  src/dense.rs:68-90
// The reference instantiates these with i32; the GPU path computes in f32/f64, and every value here
// is a small integer, so the expected results are exact in both.
//
//   test_reference_kats --host-only     structure / construction tests only (no GPU needed)
//   test_reference_kats                 everything; multiplications run on cuda:0 through the C ABI
#include <cstdio>
#include <cstring>
#include <vector>

#include "bsm.hpp"

using namespace sparse_matrix;

static int failures = 0;
#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);      \
            ++failures;                                                      \
        }                                                                    \
    } while (0)

template <typename T> static void structure_tests()
{
    // example_mat_0 (sparse.rs:815-827)
    Csr<T> m0 = Csr<T>::from_data({{5, 0, 0, 0}, {0, 8, 0, 0}, {0, 0, 3, 0}, {0, 6, 0, 0}});
    CHECK((m0.raw_v() == std::vector<T>{5, 8, 3, 6}));
    CHECK((m0.raw_col_index() == std::vector<std::size_t>{0, 1, 2, 1}));
    CHECK((m0.raw_row_index() == std::vector<std::size_t>{0, 1, 2, 3, 4}));
    CHECK(m0.get_nnz() == 4);
    // example_mat_1 (sparse.rs:829-841)
    Csr<T> m1 = Csr<T>::from_data({{10, 20, 0, 0, 0, 0}, {0, 30, 0, 40, 0, 0}, {0, 0, 50, 60, 70, 0}, {0, 0, 0, 0, 0, 80}});
    CHECK((m1.raw_v() == std::vector<T>{10, 20, 30, 40, 50, 60, 70, 80}));
    CHECK((m1.raw_col_index() == std::vector<std::size_t>{0, 1, 1, 3, 2, 3, 4, 5}));
    CHECK((m1.raw_row_index() == std::vector<std::size_t>{0, 2, 4, 7, 8}));
    // create_mat_by_insert (sparse.rs:854-868): insert in order, finalise, equals from_data
    Csr<T> mi = Csr<T>::new_(MatDim(4, 4));
    mi.insert(5, 0, 0).unwrap();
    mi.insert(8, 1, 1).unwrap();
    mi.insert(3, 2, 2).unwrap();
    mi.insert(6, 3, 1).unwrap();
    mi = std::move(mi).finalise();
    CHECK(mi == m0);
    CHECK(mi.insert(1, 0, 0) == Err(MatErr::MatrixFinalised));   // sparse.rs:223-225
    // empty rows at the top / in the middle (sparse.rs:1111-1151)
    Csr<T> top = Csr<T>::from_data({{0, 0, 0}, {1, 0, 2}, {0, 3, 0}});
    CHECK((top.raw_row_index() == std::vector<std::size_t>{0, 0, 2, 3}));
    Csr<T> mid = Csr<T>::from_data({{1, 0, 2}, {0, 0, 0}, {0, 3, 0}});
    CHECK((mid.raw_row_index() == std::vector<std::size_t>{0, 2, 2, 3}));
    // get_row_compact (sparse.rs:252-265)
    auto row = m1.get_row_compact(2);
    CHECK(row.size() == 3 && row[0] == (CsrEntry<T>{50, 2, 2}) && row[2] == (CsrEntry<T>{70, 4, 2}));
    // zero values are skipped by insert, -0.0 too (sparse.rs:229)
    Csr<T> z = Csr<T>::new_(MatDim(1, 3));
    z.insert(T(0), 0, 0).unwrap();
    z.insert(T(-0.0), 0, 1).unwrap();
    z.insert(T(2), 0, 2).unwrap();
    CHECK(std::move(z).finalise().get_nnz() == 1);
    // dense.rs:68-90 — (cols, rows) argument order, column-major
    Dense<T> d = Dense<T>::new_default_with_dims(5, 7);
    CHECK(d == Dense<T>::from_data(std::vector<std::vector<T>>(5, std::vector<T>(7, 0))));
    CHECK((d.get_dims() == MatDim(7, 5)));
    Dense<T> d2 = Dense<T>::from_data({{1, 2, 3}, {4, 5, 6}, {7, 8, 9}});
    CHECK((d2.get_col(2) == std::vector<T>{7, 8, 9}));
    CHECK((MatDim(2, 3).transpose() == MatDim(3, 2)));
    // dense_static.rs:71-96 — same layout rules for the stack-array twin
    DenseS<T, 7, 5> ds = DenseS<T, 7, 5>::new_default();
    CHECK((ds.get_dims() == MatDim(7, 5)));
    DenseS<T, 3, 3> ds2 = DenseS<T, 3, 3>::from_data({{1, 2, 3}, {4, 5, 6}, {7, 8, 9}});
    CHECK((ds2.get_col(2) == std::array<T, 3>{7, 8, 9}));
}

template <typename T> static void multiplication_tests()
{
    {   // test_dense_mul (sparse.rs:1082-1109)
        Dense<T> a = Dense<T>::from_data({{1, 2, 3, 4}, {5, 6, 7, 8}, {9, 10, 11, 12}});
        Csr<T> m = Csr<T>::from_data({{3, 0, 2, 0}, {7, 0, 0, 0}, {0, 2, 0, 1}, {0, 0, 1, 0}, {1, 0, 0, 0}});
        Csr<T> output_ref = Csr<T>::from_data({{9, 29, 49}, {7, 35, 63}, {8, 20, 32}, {3, 7, 11}, {1, 5, 9}});
        Csr<T> output = m.mul_dense(a).unwrap();
        CHECK(output_ref == output);
        // the same product through the pipelined host-to-host call, dense (column-major) result
        CHECK(m.mul_dense_into_dense(a).unwrap() == Dense<T>::from_data({{9, 7, 8, 3, 1}, {29, 35, 20, 7, 5}, {49, 63, 32, 11, 9}}));
    }
    {   // mul_dense_s (sparse.rs:448-466): the test_dense_mul operands as a DenseS<T,4,3>
        DenseS<T, 4, 3> a = DenseS<T, 4, 3>::from_data({{1, 2, 3, 4}, {5, 6, 7, 8}, {9, 10, 11, 12}});
        Csr<T> m = Csr<T>::from_data({{3, 0, 2, 0}, {7, 0, 0, 0}, {0, 2, 0, 1}, {0, 0, 1, 0}, {1, 0, 0, 0}});
        CHECK(m.mul_dense_s(a).unwrap() == Csr<T>::from_data({{9, 29, 49}, {7, 35, 63}, {8, 20, 32}, {3, 7, 11}, {1, 5, 9}}));
        DenseS<T, 3, 1> bad = DenseS<T, 3, 1>::new_default();
        CHECK(m.mul_dense_s(bad).unwrap_err() == MatErr::IncorrectDimensions);
    }
    {   // test_nnz (sparse.rs:1153-1178): zero outputs are dropped by insert
        Dense<T> a = Dense<T>::from_data({{1, 0, 3, 4}, {8, 0, 0, 5}});
        Csr<T> m = Csr<T>::from_data({{5, 2, 1, 3}, {7, 0, 1, 3}, {0, 1, 0, 0}, {0, 7, 4, 0}});
        Csr<T> output_ref = Csr<T>::from_data({{20, 55}, {22, 71}, {0, 0}, {12, 0}});
        Csr<T> output = m.mul_dense(a).unwrap();
        CHECK(output_ref == output);
        CHECK(output.get_nnz() == 5);
        CHECK((output.raw_row_index() == std::vector<std::size_t>{0, 2, 4, 4, 5}));
    }
    {   // dims mismatch -> Err(IncorrectDimensions) (sparse.rs:427-429)
        Dense<T> a = Dense<T>::from_data({{1, 2, 3}});
        Csr<T> m = Csr<T>::from_data({{1, 0, 0, 0}});
        CHECK(m.mul_dense(a).unwrap_err() == MatErr::IncorrectDimensions);
    }
    {   // test_mul_vector (sparse.rs:1501-1529)
        std::vector<T> v{0, 1, 2, 3, 4};
        Csr<T> bad = Csr<T>::from_data({{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}});
        std::vector<T> out(3);
        CHECK(bad.mul_vector(v, out) == Err(MatErr::IncorrectDimensions));
        Csr<T> eye = Csr<T>::from_data({{1, 0, 0, 0, 0}, {0, 1, 0, 0, 0}, {0, 0, 1, 0, 0}, {0, 0, 0, 1, 0}, {0, 0, 0, 0, 1}});
        std::vector<T> out5(5);
        CHECK(eye.mul_vector(v, out5) == Ok());
        CHECK(out5 == v);
        Csr<T> m = Csr<T>::from_data({{1, 0, 2, 0, 3}, {0, 1, 0, 2, 0}});
        std::vector<T> out2(2);
        CHECK(m.mul_vector(v, out2) == Ok());
        CHECK((out2 == std::vector<T>{16, 7}));
    }
    {   // device-resident use of the gpu module: operands stay in HBM between the two products
        Csr<T> m = Csr<T>::from_data({{2, 0}, {0, 3}});
        gpu::DeviceCsr<T> dm = gpu::DeviceCsr<T>::from_host(m).unwrap();
        gpu::DeviceDense<T> x = gpu::DeviceDense<T>::from_host(Dense<T>::from_data({{1, 1}, {2, 4}}));
        gpu::DeviceDense<T> y = dm.mul_dense(x).unwrap();
        gpu::DeviceDense<T> y2 = dm.mul_dense(y).unwrap();
        CHECK(y2.to_host() == Dense<T>::from_data({{4, 9}, {8, 36}}));
        CHECK(y2.into_csr().to_host() == Csr<T>::from_data({{4, 8}, {9, 36}}));
    }
}

int main(int argc, char **argv)
{
    const bool host_only = argc > 1 && std::strcmp(argv[1], "--host-only") == 0;
    structure_tests<double>();
    structure_tests<float>();
    if (!host_only) {
        gpu::init(0);
        multiplication_tests<double>();
        multiplication_tests<float>();
    }
    std::printf("%s: %d failure(s)%s\n", failures ? "FAILED" : "ok", failures, host_only ? " (host-only)" : "");
    return failures ? 1 : 0;
}
