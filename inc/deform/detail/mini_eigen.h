// deform/detail/mini_eigen.h -- a very small stand-in for the parts of Eigen that the deform API
// surface mentions (Matrix<S,3,1>, Matrix<int,3,1>, Matrix4, MatrixX, Transform<S,3,Affine>,
// Translation, AngleAxis). It is ONLY used when <Eigen/Core> is not installed (this build image has
// no Eigen); with Eigen present, deform/detail/linalg.h includes the real thing and this file is
// never seen. Same names, same call syntax for the subset, so the API headers and the restated
// reference tests compile unchanged against either.
#ifndef DEFORM_DETAIL_MINI_EIGEN_H
#define DEFORM_DETAIL_MINI_EIGEN_H

#include <cassert>
#include <cmath>
#include <cstddef>
#include <vector>

namespace Eigen {

enum { Dynamic = -1 };
enum TransformTraits { Isometry = 0x1, Affine = 0x2, AffineCompact = 0x10 | Affine, Projective = 0x20 };
typedef std::ptrdiff_t DenseIndex;

template <class S, int R, int C>
class Matrix {
public:
    typedef S Scalar;

    Matrix() : rows_(R == Dynamic ? 0 : R), cols_(C == Dynamic ? 0 : C), d_((size_t)(rows_ * cols_), S(0)) {}
    Matrix(int r, int c) : rows_(r), cols_(c), d_((size_t)(r * c), S(0)) {}
    Matrix(S x, S y, S z) : rows_(3), cols_(1), d_(3) { static_assert(R == 3 && C == 1, "3-vector ctor"); d_[0] = x; d_[1] = y; d_[2] = z; }
    template <class S2> Matrix(const Matrix<S2, R, C> &o) : rows_(o.rows()), cols_(o.cols()), d_((size_t)(rows_ * cols_)) {
        for (int i = 0; i < rows_; ++i) for (int j = 0; j < cols_; ++j) (*this)(i, j) = (S)o(i, j);
    }

    static Matrix Zero() { return Matrix(); }
    static Matrix Zero(int r, int c) { return Matrix(r, c); }
    static Matrix Identity() { Matrix m; for (int i = 0; i < m.rows_ && i < m.cols_; ++i) m(i, i) = S(1); return m; }
    static Matrix UnitX() { Matrix m; m(0) = S(1); return m; }
    static Matrix UnitY() { Matrix m; m(1) = S(1); return m; }
    static Matrix UnitZ() { Matrix m; m(2) = S(1); return m; }

    int rows() const { return rows_; }
    int cols() const { return cols_; }
    void resize(int r, int c) { rows_ = r; cols_ = c; d_.assign((size_t)(r * c), S(0)); }
    void setZero() { for (size_t i = 0; i < d_.size(); ++i) d_[i] = S(0); }

    S &operator()(int i, int j) { return d_[(size_t)(i * cols_ + j)]; }
    const S &operator()(int i, int j) const { return d_[(size_t)(i * cols_ + j)]; }
    S &operator()(int i) { return d_[(size_t)i]; }
    const S &operator()(int i) const { return d_[(size_t)i]; }
    S &operator[](int i) { return d_[(size_t)i]; }
    const S &operator[](int i) const { return d_[(size_t)i]; }
    S x() const { return d_[0]; }
    S y() const { return d_[1]; }
    S z() const { return d_[2]; }
    const S *data() const { return d_.data(); }     // NB: row-major here (Eigen's default is column-major)
    S *data() { return d_.data(); }

    template <class T> Matrix<T, R, C> cast() const { return Matrix<T, R, C>(*this); }

    Matrix operator+(const Matrix &o) const { Matrix r(*this); for (size_t i = 0; i < d_.size(); ++i) r.d_[i] += o.d_[i]; return r; }
    Matrix operator-(const Matrix &o) const { Matrix r(*this); for (size_t i = 0; i < d_.size(); ++i) r.d_[i] -= o.d_[i]; return r; }
    Matrix operator-() const { Matrix r(*this); for (size_t i = 0; i < d_.size(); ++i) r.d_[i] = -r.d_[i]; return r; }
    Matrix operator*(S s) const { Matrix r(*this); for (size_t i = 0; i < d_.size(); ++i) r.d_[i] *= s; return r; }
    Matrix &operator+=(const Matrix &o) { for (size_t i = 0; i < d_.size(); ++i) d_[i] += o.d_[i]; return *this; }
    template <int C2> Matrix<S, R, C2> operator*(const Matrix<S, C, C2> &o) const {
        Matrix<S, R, C2> r(rows_, o.cols());
        for (int i = 0; i < rows_; ++i) for (int j = 0; j < o.cols(); ++j) { S s = 0; for (int k = 0; k < cols_; ++k) s += (*this)(i, k) * o(k, j); r(i, j) = s; }
        return r;
    }
    Matrix<S, C, R> transpose() const { Matrix<S, C, R> r(cols_, rows_); for (int i = 0; i < rows_; ++i) for (int j = 0; j < cols_; ++j) r(j, i) = (*this)(i, j); return r; }
    S squaredNorm() const { S s = 0; for (size_t i = 0; i < d_.size(); ++i) s += d_[i] * d_[i]; return s; }
    S norm() const { return std::sqrt(squaredNorm()); }
    S dot(const Matrix &o) const { S s = 0; for (size_t i = 0; i < d_.size(); ++i) s += d_[i] * o.d_[i]; return s; }
    Matrix cross(const Matrix &o) const { return Matrix(y() * o.z() - z() * o.y(), z() * o.x() - x() * o.z(), x() * o.y() - y() * o.x()); }
    Matrix normalized() const { return (*this) * (S(1) / norm()); }

    // Eigen's fuzzy comparison: |a - b|_F <= prec * min(|a|_F, |b|_F)
    bool isApprox(const Matrix &o, S prec = S(1e-5)) const {
        if (rows_ != o.rows_ || cols_ != o.cols_) return false;
        const S na = norm(), nb = o.norm();
        return ((*this) - o).norm() <= prec * (na < nb ? na : nb);
    }

    // comma initialiser: m << a, b, c, ...;  (row-major fill like Eigen)
    struct CommaInit {
        Matrix &m; size_t k;
        CommaInit &operator,(S v) { m.d_[k++] = v; return *this; }
    };
    CommaInit operator<<(S v) { d_[0] = v; CommaInit c = {*this, 1}; return c; }

private:
    template <class, int, int> friend class Matrix;
    int rows_, cols_;
    std::vector<S> d_;
};

template <class S, int R, int C> Matrix<S, R, C> operator*(S s, const Matrix<S, R, C> &m) { return m * s; }

typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<int, 3, 1> Vector3i;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, Dynamic, Dynamic> MatrixXf;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;

template <class S> class AngleAxis;
template <class S, int Dim> class Translation;

// Rigid/affine transform stored as a homogeneous 4x4.
template <class S, int Dim, int Mode>
class Transform {
public:
    typedef S Scalar;
    typedef Matrix<S, 4, 4> MatrixType;
    typedef Matrix<S, 3, 1> VectorType;
    Transform() : m_(MatrixType::Identity()) {}
    explicit Transform(const MatrixType &m) : m_(m) {}
    Transform(const AngleAxis<S> &aa);
    Transform(const Translation<S, Dim> &t);
    static Transform Identity() { return Transform(); }
    MatrixType &matrix() { return m_; }
    const MatrixType &matrix() const { return m_; }
    Matrix<S, 3, 3> linear() const { Matrix<S, 3, 3> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = m_(i, j); return r; }
    VectorType translation() const { return VectorType(m_(0, 3), m_(1, 3), m_(2, 3)); }
    Transform operator*(const Transform &o) const { return Transform(m_ * o.m_); }
    Transform operator*(const AngleAxis<S> &aa) const { return (*this) * Transform(aa); }
    Transform operator*(const Translation<S, Dim> &t) const { return (*this) * Transform(t); }
    VectorType operator*(const VectorType &v) const {
        return VectorType(m_(0, 0) * v(0) + m_(0, 1) * v(1) + m_(0, 2) * v(2) + m_(0, 3),
                          m_(1, 0) * v(0) + m_(1, 1) * v(1) + m_(1, 2) * v(2) + m_(1, 3),
                          m_(2, 0) * v(0) + m_(2, 1) * v(1) + m_(2, 2) * v(2) + m_(2, 3));
    }
    // inverse of a rigid transform (the only mode the deform headers use, deformation_util.h:38)
    Transform inverse(TransformTraits = Affine) const {
        Transform r;
        for (int i = 0; i < 3; ++i) {
            S t = 0;
            for (int j = 0; j < 3; ++j) { r.m_(i, j) = m_(j, i); t -= m_(j, i) * m_(j, 3); }
            r.m_(i, 3) = t;
        }
        return r;
    }
    template <class T> Transform<T, Dim, Mode> cast() const { return Transform<T, Dim, Mode>(m_.template cast<T>()); }
private:
    MatrixType m_;
};

template <class S>
class AngleAxis {
public:
    AngleAxis(S angle, const Matrix<S, 3, 1> &axis) : angle_(angle), axis_(axis) {}
    Matrix<S, 3, 3> toRotationMatrix() const {
        const S c = std::cos(angle_), s = std::sin(angle_), t = S(1) - c;
        const S x = axis_(0), y = axis_(1), z = axis_(2);
        Matrix<S, 3, 3> r;
        r(0, 0) = t * x * x + c;     r(0, 1) = t * x * y - s * z; r(0, 2) = t * x * z + s * y;
        r(1, 0) = t * x * y + s * z; r(1, 1) = t * y * y + c;     r(1, 2) = t * y * z - s * x;
        r(2, 0) = t * x * z - s * y; r(2, 1) = t * y * z + s * x; r(2, 2) = t * z * z + c;
        return r;
    }
    Matrix<S, 3, 3> matrix() const { return toRotationMatrix(); }
    template <int Mode> Transform<S, 3, Mode> operator*(const Transform<S, 3, Mode> &t) const { return Transform<S, 3, Mode>(*this) * t; }
private:
    S angle_;
    Matrix<S, 3, 1> axis_;
};

template <class S, int Dim>
class Translation {
public:
    Translation(S x, S y, S z) : v_(x, y, z) {}
    explicit Translation(const Matrix<S, 3, 1> &v) : v_(v) {}
    const Matrix<S, 3, 1> &vector() const { return v_; }
    template <int Mode> Transform<S, Dim, Mode> operator*(const Transform<S, Dim, Mode> &t) const { return Transform<S, Dim, Mode>(*this) * t; }
    Transform<S, Dim, Affine> operator*(const AngleAxis<S> &aa) const { return Transform<S, Dim, Affine>(*this) * Transform<S, Dim, Affine>(aa); }
private:
    Matrix<S, 3, 1> v_;
};

template <class S, int Dim, int Mode>
Transform<S, Dim, Mode>::Transform(const AngleAxis<S> &aa) : m_(MatrixType::Identity()) {
    const Matrix<S, 3, 3> r = aa.toRotationMatrix();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m_(i, j) = r(i, j);
}
template <class S, int Dim, int Mode>
Transform<S, Dim, Mode>::Transform(const Translation<S, Dim> &t) : m_(MatrixType::Identity()) {
    for (int i = 0; i < 3; ++i) m_(i, 3) = t.vector()(i);
}

typedef Translation<float, 3> Translation3f;
typedef Translation<double, 3> Translation3d;
typedef AngleAxis<float> AngleAxisf;
typedef AngleAxis<double> AngleAxisd;
typedef Transform<float, 3, Affine> Affine3f;
typedef Transform<double, 3, Affine> Affine3d;

}  // namespace Eigen

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define DEFORM_USING_MINI_EIGEN 1

#endif
