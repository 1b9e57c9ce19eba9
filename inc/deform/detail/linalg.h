// deform/detail/linalg.h -- picks the linear-algebra types the API is written against:
// real Eigen when it is installed (what users of cheind/mesh-deform have), the bundled stand-in otherwise.
#ifndef DEFORM_DETAIL_LINALG_H
#define DEFORM_DETAIL_LINALG_H

#if !defined(DEFORM_FORCE_MINI_EIGEN) && defined(__has_include)
#if __has_include(<Eigen/Core>)
#define DEFORM_HAVE_EIGEN 1
#endif
#endif

#ifdef DEFORM_HAVE_EIGEN
#include <Eigen/Core>
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <Eigen/Sparse>
#else
#include <deform/detail/mini_eigen.h>
#endif

#include <vector>

namespace deform {
namespace detail {

// Row-major compressed sparse matrix as the engine hands it out (rowptr / ascending colidx / values).
// Stands in for Eigen::SparseMatrix<Scalar, RowMajor> (reference inc/deform/arap.h:445) when Eigen is absent;
// converts to a dense matrix the way `Eigen::MatrixXf sp = sparse;` does in reference tests/test_cotan.cpp:41.
template <class S>
class CsrMatrix {
public:
    CsrMatrix() : rows_(0), cols_(0) {}
    CsrMatrix(int rows, int cols, std::vector<int> rowptr, std::vector<int> colidx, std::vector<S> values)
        : rows_(rows), cols_(cols), rowptr_(std::move(rowptr)), colidx_(std::move(colidx)), values_(std::move(values)) {}
    int rows() const { return rows_; }
    int cols() const { return cols_; }
    int nonZeros() const { return (int)colidx_.size(); }
    int outerSize() const { return rows_; }
    const int *outerIndexPtr() const { return rowptr_.data(); }
    const int *innerIndexPtr() const { return colidx_.data(); }
    const S *valuePtr() const { return values_.data(); }
    S coeff(int r, int c) const {
        for (int k = rowptr_[(size_t)r]; k < rowptr_[(size_t)r + 1]; ++k) if (colidx_[(size_t)k] == c) return values_[(size_t)k];
        return S(0);
    }
    template <class T>
    operator Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic>() const {
        Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> d(rows_, cols_);
        d.setZero();
        for (int r = 0; r < rows_; ++r)
            for (int k = rowptr_[(size_t)r]; k < rowptr_[(size_t)r + 1]; ++k) d(r, colidx_[(size_t)k]) = (T)values_[(size_t)k];
        return d;
    }
#ifdef DEFORM_HAVE_EIGEN
    operator Eigen::SparseMatrix<S, Eigen::RowMajor>() const {
        return Eigen::Map<const Eigen::SparseMatrix<S, Eigen::RowMajor>>(rows_, cols_, nonZeros(), rowptr_.data(), colidx_.data(), values_.data());
    }
#endif
private:
    int rows_, cols_;
    std::vector<int> rowptr_, colidx_;
    std::vector<S> values_;
};

}  // namespace detail
}  // namespace deform

#endif
