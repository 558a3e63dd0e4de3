/* Stand-in for the few Eigen headers the reference's hot-path sources include (<Eigen/Geometry>, <Eigen/Eigenvalues>):
 * Eigen is not installed in this image, so the reference's OWN source files (lidar_factor.cc, rotation.cc,
 * pointcloud.cc, the pose part of misc.cc) are compiled where they lie against this header (oracle/Makefile `ffref`).
 *
 * What this pins and what it does not. The reference's statements - which products, in which grouping, with which signs
 * and indices - are the reference's own text. The primitives underneath (fixed-size products, quaternion <-> matrix /
 * angle-axis, normalisation, the 5 x 3 pivoted QR solve) are written here from Eigen 3.4's documented behaviour, eagerly
 * evaluated (no expression templates, no vectorisation): a k-sum runs in ascending k from the first product. Real Eigen
 * may group a sum differently (packet code, scalar factors pulled through a product), so agreement with real Eigen is
 * expected at the rounding level, not bit for bit. The 5 x 3 solve uses column-pivoted Householder reflections written
 * independently of oracle/oracle.cc. Test infrastructure only: nothing under ff-lins_b200/ includes or links this. */
#ifndef FFLINS_ORACLE_MINI_EIGEN_H
#define FFLINS_ORACLE_MINI_EIGEN_H

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <ostream>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_ALIGN16 alignas(16)

namespace Eigen {

enum { ColMajor = 0, RowMajor = 1 };
template <class T> using aligned_allocator = std::allocator<T>;

template <class T, int R, int C, int Opt = ColMajor> struct Matrix;
template <class T> struct Quaternion;
template <class T> struct AngleAxis;
template <class M> struct Map;

namespace mini {
template <int R, int C, int Opt> constexpr int at(int r, int c) { return (Opt & RowMajor) ? r * C + c : c * R + r; }

/* writable view of BR x BC coefficients of some storage: what block<>(), row() and Map hand out */
template <class T, int BR, int BC> struct View {
    T *p;
    int rs, cs; /* strides of a row step and of a column step */
    T &operator()(int r, int c) const { return p[r * rs + c * cs]; }
    template <int O> const View &operator=(const Matrix<T, BR, BC, O> &m) const {
        for (int r = 0; r < BR; r++)
            for (int c = 0; c < BC; c++) (*this)(r, c) = m(r, c);
        return *this;
    }
    /* Eigen lets a column vector be assigned to a row view (and the reverse) */
    template <int O, int X = BR, class = typename std::enable_if<(X == 1 || BC == 1) && BR != BC>::type>
    const View &operator=(const Matrix<T, BC, BR, O> &m) const {
        for (int r = 0; r < BR; r++)
            for (int c = 0; c < BC; c++) (*this)(r, c) = m(c, r);
        return *this;
    }
    operator Matrix<T, BR, BC>() const {
        Matrix<T, BR, BC> m;
        for (int r = 0; r < BR; r++)
            for (int c = 0; c < BC; c++) m(r, c) = (*this)(r, c);
        return m;
    }
    void setZero() const {
        for (int r = 0; r < BR; r++)
            for (int c = 0; c < BC; c++) (*this)(r, c) = T(0);
    }
    void setIdentity() const {
        for (int r = 0; r < BR; r++)
            for (int c = 0; c < BC; c++) (*this)(r, c) = r == c ? T(1) : T(0);
    }
    const View &operator*=(T s) const {
        for (int r = 0; r < BR; r++)
            for (int c = 0; c < BC; c++) (*this)(r, c) = (*this)(r, c) * s;
        return *this;
    }
};

template <class T, int R, int C> struct CommaInit {
    Matrix<T, R, C> *m;
    int k;
    CommaInit &operator,(T v) {
        (*m)(k / C, k % C) = v; /* Eigen's comma initialiser fills row by row */
        k++;
        return *this;
    }
};
} // namespace mini

/* a 1 x 1 result is a scalar (Eigen: inner products convert implicitly) */
template <class D, class T, bool one> struct ScalarConv {};
template <class D, class T> struct ScalarConv<D, T, true> {
    operator T() const { return static_cast<const D *>(this)->m[0]; }
};

template <class T, int R, int C, int Opt> struct Matrix : ScalarConv<Matrix<T, R, C, Opt>, T, R * C == 1> {
    T m[R * C];
    typedef T Scalar;

    Matrix() {
        for (int i = 0; i < R * C; i++) m[i] = T(0);
    }
    template <int O2> Matrix(const Matrix<T, R, C, O2> &o) {
        for (int r = 0; r < R; r++)
            for (int c = 0; c < C; c++) (*this)(r, c) = o(r, c);
    }
    Matrix(T a, T b) {
        static_assert(R * C == 2, "two-coefficient constructor");
        m[0] = a, m[1] = b;
    }
    Matrix(T a, T b, T c) {
        static_assert(R * C == 3, "three-coefficient constructor");
        m[0] = a, m[1] = b, m[2] = c;
    }
    explicit Matrix(const Quaternion<T> &q) { *this = q.toRotationMatrix(); }

    static Matrix Zero() { return Matrix(); }
    static Matrix Identity() {
        Matrix r;
        for (int i = 0; i < (R < C ? R : C); i++) r(i, i) = T(1);
        return r;
    }
    static Matrix UnitX() { Matrix r; r.m[0] = T(1); return r; }
    static Matrix UnitY() { Matrix r; r.m[1] = T(1); return r; }
    static Matrix UnitZ() { Matrix r; r.m[2] = T(1); return r; }

    int rows() const { return R; }
    int cols() const { return C; }
    T *data() { return m; }
    const T *data() const { return m; }
    T &operator()(int r, int c) { return m[mini::at<R, C, Opt>(r, c)]; }
    const T &operator()(int r, int c) const { return m[mini::at<R, C, Opt>(r, c)]; }
    T &operator()(int i) { return m[i]; }
    const T &operator()(int i) const { return m[i]; }
    T &operator[](int i) { return m[i]; }
    const T &operator[](int i) const { return m[i]; }
    T &x() { return m[0]; }
    T &y() { return m[1]; }
    T &z() { return m[2]; }
    const T &x() const { return m[0]; }
    const T &y() const { return m[1]; }
    const T &z() const { return m[2]; }

    void setZero() { for (int i = 0; i < R * C; i++) m[i] = T(0); }
    void setOnes() { for (int i = 0; i < R * C; i++) m[i] = T(1); }
    Matrix &operator*=(T s) { for (int i = 0; i < R * C; i++) m[i] = m[i] * s; return *this; }
    Matrix &operator/=(T s) { for (int i = 0; i < R * C; i++) m[i] = m[i] / s; return *this; }
    Matrix &operator+=(const Matrix &o) { for (int i = 0; i < R * C; i++) m[i] = m[i] + o.m[i]; return *this; }
    Matrix &operator-=(const Matrix &o) { for (int i = 0; i < R * C; i++) m[i] = m[i] - o.m[i]; return *this; }

    mini::CommaInit<T, R, C> operator<<(T v) {
        static_assert(Opt == ColMajor, "comma initialiser: default storage only");
        (*this)(0, 0) = v;
        return mini::CommaInit<T, R, C>{reinterpret_cast<Matrix<T, R, C> *>(this), 1};
    }

    Matrix<T, C, R> transpose() const {
        Matrix<T, C, R> t;
        for (int r = 0; r < R; r++)
            for (int c = 0; c < C; c++) t(c, r) = (*this)(r, c);
        return t;
    }
    template <class U> Matrix<U, R, C> cast() const {
        Matrix<U, R, C> o;
        for (int r = 0; r < R; r++)
            for (int c = 0; c < C; c++) o(r, c) = static_cast<U>((*this)(r, c));
        return o;
    }
    T squaredNorm() const {
        T s = m[0] * m[0];
        for (int i = 1; i < R * C; i++) s = s + m[i] * m[i];
        return s;
    }
    T norm() const { return std::sqrt(squaredNorm()); }
    T stableNorm() const { return norm(); }
    /* Eigen 3.4 MatrixBase::normalized(): divide by sqrt(squaredNorm) when that is positive */
    Matrix normalized() const {
        T n2 = squaredNorm();
        if (n2 > T(0)) {
            Matrix o = *this;
            o /= std::sqrt(n2);
            return o;
        }
        return *this;
    }
    void normalize() { *this = normalized(); }
    T dot(const Matrix &o) const {
        T s = m[0] * o.m[0];
        for (int i = 1; i < R * C; i++) s = s + m[i] * o.m[i];
        return s;
    }
    Matrix cross(const Matrix &o) const {
        static_assert(R * C == 3, "cross product of 3-vectors");
        return Matrix(m[1] * o.m[2] - m[2] * o.m[1], m[2] * o.m[0] - m[0] * o.m[2], m[0] * o.m[1] - m[1] * o.m[0]);
    }

    template <int BR, int BC> mini::View<T, BR, BC> block(int r, int c) {
        return mini::View<T, BR, BC>{&(*this)(r, c), mini::at<R, C, Opt>(1, 0) - mini::at<R, C, Opt>(0, 0),
                                     (C > 1 ? mini::at<R, C, Opt>(0, 1) : 1) - mini::at<R, C, Opt>(0, 0)};
    }
    template <int BR, int BC> Matrix<T, BR, BC> bottomRightCorner() const {
        Matrix<T, BR, BC> o;
        for (int r = 0; r < BR; r++)
            for (int c = 0; c < BC; c++) o(r, c) = (*this)(R - BR + r, C - BC + c);
        return o;
    }
    mini::View<T, 1, C> row(int r) { return block<1, C>(r, 0); }
    mini::View<T, R, 1> col(int c) { return block<R, 1>(0, c); }

    struct ColPivQR;
    ColPivQR colPivHouseholderQr() const { return ColPivQR{*this}; }
};

/* ---- arithmetic (eager; a k-sum starts from its first product and adds in ascending k) ---- */
template <class T, int R, int C, int O1, int O2, class = typename std::enable_if<(R > 0 && C > 0)>::type>
Matrix<T, R, C> operator+(const Matrix<T, R, C, O1> &a, const Matrix<T, R, C, O2> &b) {
    Matrix<T, R, C> o;
    for (int r = 0; r < R; r++)
        for (int c = 0; c < C; c++) o(r, c) = a(r, c) + b(r, c);
    return o;
}
template <class T, int R, int C, int O1, int O2, class = typename std::enable_if<(R > 0 && C > 0)>::type>
Matrix<T, R, C> operator-(const Matrix<T, R, C, O1> &a, const Matrix<T, R, C, O2> &b) {
    Matrix<T, R, C> o;
    for (int r = 0; r < R; r++)
        for (int c = 0; c < C; c++) o(r, c) = a(r, c) - b(r, c);
    return o;
}
template <class T, int R, int C, int O, class = typename std::enable_if<(R > 0 && C > 0)>::type>
Matrix<T, R, C> operator-(const Matrix<T, R, C, O> &a) {
    Matrix<T, R, C> o;
    for (int r = 0; r < R; r++)
        for (int c = 0; c < C; c++) o(r, c) = -a(r, c);
    return o;
}
template <class T, int R, int C, int O, class = typename std::enable_if<(R > 0 && C > 0)>::type>
Matrix<T, R, C> operator*(const Matrix<T, R, C, O> &a, T s) {
    Matrix<T, R, C> o;
    for (int r = 0; r < R; r++)
        for (int c = 0; c < C; c++) o(r, c) = a(r, c) * s;
    return o;
}
template <class T, int R, int C, int O, class = typename std::enable_if<(R > 0 && C > 0)>::type>
Matrix<T, R, C> operator*(T s, const Matrix<T, R, C, O> &a) {
    Matrix<T, R, C> o;
    for (int r = 0; r < R; r++)
        for (int c = 0; c < C; c++) o(r, c) = s * a(r, c);
    return o;
}
template <class T, int R, int C, int O, class = typename std::enable_if<(R > 0 && C > 0) && !std::is_same<T, int>::value>::type>
Matrix<T, R, C> operator*(int s, const Matrix<T, R, C, O> &a) {
    return T(s) * a;
}
template <class T, int R, int C, int O, class = typename std::enable_if<(R > 0 && C > 0)>::type>
Matrix<T, R, C> operator/(const Matrix<T, R, C, O> &a, T s) {
    Matrix<T, R, C> o;
    for (int r = 0; r < R; r++)
        for (int c = 0; c < C; c++) o(r, c) = a(r, c) / s;
    return o;
}
template <class T, int R, int K, int C, int O1, int O2, class = typename std::enable_if<(R > 0 && K > 0 && C > 0)>::type>
Matrix<T, R, C> operator*(const Matrix<T, R, K, O1> &a, const Matrix<T, K, C, O2> &b) {
    Matrix<T, R, C> o;
    for (int r = 0; r < R; r++)
        for (int c = 0; c < C; c++) {
            T s = a(r, 0) * b(0, c);
            for (int k = 1; k < K; k++) s = s + a(r, k) * b(k, c);
            o(r, c) = s;
        }
    return o;
}
template <class T, int R, int C, int O> std::ostream &operator<<(std::ostream &os, const Matrix<T, R, C, O> &a) {
    for (int r = 0; r < R; r++) {
        for (int c = 0; c < C; c++) os << (c ? " " : "") << a(r, c);
        if (r + 1 < R) os << "\n";
    }
    return os;
}

/* ---- run-time sized matrices (Eigen::Dynamic): the surface residual_block_info.cc and marginalization_info.cc use -
 * resize, Zero, blocks / segments / column ranges as views, sums, differences, products, transpose, the coefficient-wise
 * array operations of the eigenvalue cut-off, asDiagonal and a symmetric eigen-solver - evaluated eagerly in the order the
 * C++ expression groups them. The eigen-solver is a cyclic Jacobi iteration (Eigen tridiagonalises and runs QR sweeps):
 * eigenvalues agree to rounding, eigenvectors up to sign / rotation inside degenerate subspaces, so what is compared
 * downstream are the products that do not depend on that choice (J0^T J0, J0^T e0). ---- */
enum { Dynamic = -1 };
template <class T> struct DynView;
template <class T, int Opt> struct DynBase {
    int r = 0, c = 0;
    std::vector<T> v;
    int rows() const { return r; }
    int cols() const { return c; }
    int size() const { return r * c; }
    T *data() { return v.data(); }
    const T *data() const { return v.data(); }
    T &operator()(int i, int j) { return v[(Opt & RowMajor) ? i * c + j : j * r + i]; }
    const T &operator()(int i, int j) const { return v[(Opt & RowMajor) ? i * c + j : j * r + i]; }
    void resize(int rows, int cols) { r = rows, c = cols, v.assign((size_t) rows * cols, T(0)); }
    void setZero() { std::fill(v.begin(), v.end(), T(0)); }
    void setZero(int rows, int cols) { resize(rows, cols); }
    T squaredNorm() const {
        T s = T(0);
        for (size_t i = 0; i < v.size(); i++) s = i ? s + v[i] * v[i] : v[i] * v[i];
        return s;
    }
    DynView<T> block(int i, int j, int nr, int nc) {
        return DynView<T>{&(*this)(i, j), nr, nc, (Opt & RowMajor) ? c : 1, (Opt & RowMajor) ? 1 : r};
    }
    DynView<T> block(int i, int j, int nr, int nc) const { return const_cast<DynBase *>(this)->block(i, j, nr, nc); }
    DynView<T> leftCols(int n) const { return block(0, 0, r, n); }
    DynView<T> middleCols(int j, int n) const { return block(0, j, r, n); }
    DynView<T> segment(int i, int n) const { return block(i, 0, n, 1); }
    template <int BR, int BC> mini::View<T, BR, BC> block(int i, int j) const {
        DynBase *self = const_cast<DynBase *>(this);
        return mini::View<T, BR, BC>{&(*self)(i, j), (Opt & RowMajor) ? c : 1, (Opt & RowMajor) ? 1 : r};
    }
    void setIdentity(int rows, int cols) {
        resize(rows, cols);
        for (int i = 0; i < (rows < cols ? rows : cols); i++) (*this)(i, i) = T(1);
    }
    Matrix<T, Dynamic, Dynamic> inverse() const;   /* Gauss-Jordan with partial pivoting (Eigen: PartialPivLU) */
};
/* a window onto a DynBase (block, segment, leftCols): readable as a matrix, assignable, += / -= */
template <class T> struct DynView {
    T *p;
    int r, c, rs, cs;
    T &operator()(int i, int j) const { return p[i * rs + j * cs]; }
    Matrix<T, Dynamic, Dynamic> eval() const;
    Matrix<T, Dynamic, Dynamic> transpose() const;
    template <int O> const DynView &operator=(const DynBase<T, O> &m) const {
        for (int i = 0; i < r; i++)
            for (int j = 0; j < c; j++) (*this)(i, j) = m(i, j);
        return *this;
    }
    template <int R2, int C2, int O, class = typename std::enable_if<(R2 > 0 && C2 > 0)>::type>
    const DynView &operator=(const Matrix<T, R2, C2, O> &m) const {
        for (int i = 0; i < r; i++)
            for (int j = 0; j < c; j++) (*this)(i, j) = m(i, j);
        return *this;
    }
    const DynView &operator=(const DynView &m) const {
        for (int i = 0; i < r; i++)
            for (int j = 0; j < c; j++) (*this)(i, j) = m(i, j);
        return *this;
    }
    template <int O> const DynView &operator+=(const DynBase<T, O> &m) const {
        for (int i = 0; i < r; i++)
            for (int j = 0; j < c; j++) (*this)(i, j) = (*this)(i, j) + m(i, j);
        return *this;
    }
    template <int O> const DynView &operator-=(const DynBase<T, O> &m) const {
        for (int i = 0; i < r; i++)
            for (int j = 0; j < c; j++) (*this)(i, j) = (*this)(i, j) - m(i, j);
        return *this;
    }
};
template <class T> struct ArrayX;
template <class T> struct DiagX {
    std::vector<T> d;
};
template <class T, int Opt> struct Matrix<T, Dynamic, Dynamic, Opt> : DynBase<T, Opt> {
    Matrix() {}
    Matrix(int rows, int cols) { this->resize(rows, cols); }
    template <int O2> Matrix(const DynBase<T, O2> &o) { *this = o; }
    Matrix(const DynView<T> &w) {
        this->resize(w.r, w.c);
        for (int i = 0; i < w.r; i++)
            for (int j = 0; j < w.c; j++) (*this)(i, j) = w(i, j);
    }
    template <class M> Matrix(const Map<M> &mp) {
        auto f = mp.eval();
        this->resize(f.rows(), f.cols());
        for (int i = 0; i < f.rows(); i++)
            for (int j = 0; j < f.cols(); j++) (*this)(i, j) = f(i, j);
    }
    static Matrix Zero(int rows, int cols) { return Matrix(rows, cols); }
    template <int O2> Matrix &operator=(const DynBase<T, O2> &o) {
        this->resize(o.r, o.c);
        for (int i = 0; i < o.r; i++)
            for (int j = 0; j < o.c; j++) (*this)(i, j) = o(i, j);
        return *this;
    }
    Matrix &operator*=(T s) { for (auto &x : this->v) x = x * s; return *this; }
    template <int O2> Matrix &operator+=(const DynBase<T, O2> &o) {
        for (int i = 0; i < o.r; i++)
            for (int j = 0; j < o.c; j++) (*this)(i, j) = (*this)(i, j) + o(i, j);
        return *this;
    }
    Matrix<T, Dynamic, Dynamic> transpose() const {
        Matrix<T, Dynamic, Dynamic> t;
        t.resize(this->c, this->r);
        for (int i = 0; i < this->r; i++)
            for (int j = 0; j < this->c; j++) t(j, i) = (*this)(i, j);
        return t;
    }
};
template <class T, int Opt> struct Matrix<T, Dynamic, 1, Opt> : DynBase<T, ColMajor> {
    Matrix() {}
    explicit Matrix(int n) { this->resize(n); }
    template <int O2> Matrix(const DynBase<T, O2> &o) { this->r = o.r, this->c = o.c, this->v = o.v; }
    Matrix(const DynView<T> &w) {
        this->resize(w.r);
        for (int i = 0; i < w.r; i++) this->v[i] = w(i, 0);
    }
    explicit Matrix(const ArrayX<T> &a);
    static Matrix Zero(int n) { return Matrix(n); }
    void resize(int n) { DynBase<T, ColMajor>::resize(n, 1); }
    void resize(int n, int m) { DynBase<T, ColMajor>::resize(n, m); }
    void setZero() { DynBase<T, ColMajor>::setZero(); }
    void setZero(int n) { resize(n); }
    int size() const { return this->r; }
    T &operator[](int i) { return this->v[i]; }
    const T &operator[](int i) const { return this->v[i]; }
    using DynBase<T, ColMajor>::operator();
    T &operator()(int i) { return this->v[i]; }
    const T &operator()(int i) const { return this->v[i]; }
    Matrix &operator*=(T s) { for (auto &x : this->v) x = x * s; return *this; }
    template <int O2> Matrix &operator+=(const DynBase<T, O2> &o) {
        for (int i = 0; i < this->r; i++) this->v[i] = this->v[i] + o(i, 0);
        return *this;
    }
    Matrix<T, Dynamic, Dynamic> transpose() const {
        Matrix<T, Dynamic, Dynamic> t;
        t.resize(1, this->r);
        for (int i = 0; i < this->r; i++) t(0, i) = this->v[i];
        return t;
    }
    template <int N> Matrix<T, N, 1> head() const {
        Matrix<T, N, 1> o;
        for (int i = 0; i < N; i++) o[i] = this->v[i];
        return o;
    }
    template <int N> mini::View<T, N, 1> segment(int i) { return mini::View<T, N, 1>{&this->v[i], 1, 1}; }
    using DynBase<T, ColMajor>::segment;
    ArrayX<T> array() const;
    Matrix cwiseSqrt() const {
        Matrix o(this->r);
        for (int i = 0; i < this->r; i++) o[i] = std::sqrt(this->v[i]);
        return o;
    }
    DiagX<T> asDiagonal() const { return DiagX<T>{this->v}; }
};
template <class T> Matrix<T, Dynamic, Dynamic> DynView<T>::eval() const { return Matrix<T, Dynamic, Dynamic>(*this); }
template <class T> Matrix<T, Dynamic, Dynamic> DynView<T>::transpose() const { return eval().transpose(); }

template <class T, int O1, int O2> Matrix<T, Dynamic, Dynamic> operator*(const DynBase<T, O1> &a, const DynBase<T, O2> &b) {
    Matrix<T, Dynamic, Dynamic> o;
    o.resize(a.r, b.c);
    for (int i = 0; i < a.r; i++)
        for (int j = 0; j < b.c; j++) {
            T s = a(i, 0) * b(0, j);
            for (int k = 1; k < a.c; k++) s = s + a(i, k) * b(k, j);
            o(i, j) = s;
        }
    return o;
}
template <class T, int O> Matrix<T, Dynamic, Dynamic> operator*(T s, const DynBase<T, O> &a) {
    Matrix<T, Dynamic, Dynamic> o;
    o.resize(a.r, a.c);
    for (int i = 0; i < a.r; i++)
        for (int j = 0; j < a.c; j++) o(i, j) = s * a(i, j);
    return o;
}
template <class T, int O> Matrix<T, Dynamic, Dynamic> operator-(const DynBase<T, O> &a) { return T(-1) * a; }
template <class T, int O1, int O2> Matrix<T, Dynamic, Dynamic> operator-(const DynBase<T, O1> &a, const DynBase<T, O2> &b) {
    Matrix<T, Dynamic, Dynamic> o;
    o.resize(a.r, a.c);
    for (int i = 0; i < a.r; i++)
        for (int j = 0; j < a.c; j++) o(i, j) = a(i, j) - b(i, j);
    return o;
}
template <class T, int O1, int O2> Matrix<T, Dynamic, Dynamic> operator+(const DynBase<T, O1> &a, const DynBase<T, O2> &b) {
    Matrix<T, Dynamic, Dynamic> o;
    o.resize(a.r, a.c);
    for (int i = 0; i < a.r; i++)
        for (int j = 0; j < a.c; j++) o(i, j) = a(i, j) + b(i, j);
    return o;
}
template <class T, int O> Matrix<T, Dynamic, Dynamic> operator+(const DynView<T> &a, const DynBase<T, O> &b) { return a.eval() + b; }
/* diag * M scales rows, M * diag scales columns */
template <class T, int O> Matrix<T, Dynamic, Dynamic> operator*(const DiagX<T> &d, const DynBase<T, O> &a) {
    Matrix<T, Dynamic, Dynamic> o;
    o.resize(a.r, a.c);
    for (int i = 0; i < a.r; i++)
        for (int j = 0; j < a.c; j++) o(i, j) = d.d[i] * a(i, j);
    return o;
}
template <class T, int O> Matrix<T, Dynamic, Dynamic> operator*(const DynBase<T, O> &a, const DiagX<T> &d) {
    Matrix<T, Dynamic, Dynamic> o;
    o.resize(a.r, a.c);
    for (int i = 0; i < a.r; i++)
        for (int j = 0; j < a.c; j++) o(i, j) = a(i, j) * d.d[j];
    return o;
}
/* coefficient-wise arrays: (a > eps).select(b, 0), a.inverse() */
template <class T> struct BoolX {
    std::vector<char> m;
    ArrayX<T> select(const ArrayX<T> &a, T other) const;
};
template <class T> struct ArrayX {
    std::vector<T> a;
    BoolX<T> operator>(T x) const {
        BoolX<T> b;
        for (T v : a) b.m.push_back(v > x);
        return b;
    }
    ArrayX inverse() const {
        ArrayX o;
        for (T v : a) o.a.push_back(T(1) / v);
        return o;
    }
};
template <class T> ArrayX<T> BoolX<T>::select(const ArrayX<T> &a, T other) const {
    ArrayX<T> o;
    for (size_t i = 0; i < m.size(); i++) o.a.push_back(m[i] ? a.a[i] : other);
    return o;
}
template <class T, int Opt> ArrayX<T> Matrix<T, Dynamic, 1, Opt>::array() const { return ArrayX<T>{this->v}; }
template <class T, int Opt> Matrix<T, Dynamic, 1, Opt>::Matrix(const ArrayX<T> &a) {
    this->r = (int) a.a.size(), this->c = 1, this->v = a.a;
}
template <class T, int Opt> Matrix<T, Dynamic, Dynamic> DynBase<T, Opt>::inverse() const {
    const int n = r;
    Matrix<T, Dynamic, Dynamic> a(n, n), inv(n, n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) a(i, j) = (*this)(i, j), inv(i, j) = i == j ? T(1) : T(0);
    for (int k = 0; k < n; k++) {
        int piv = k;
        for (int i = k + 1; i < n; i++)
            if (std::fabs(a(i, k)) > std::fabs(a(piv, k))) piv = i;
        if (piv != k)
            for (int j = 0; j < n; j++) std::swap(a(k, j), a(piv, j)), std::swap(inv(k, j), inv(piv, j));
        const T d = a(k, k);
        for (int j = 0; j < n; j++) a(k, j) = a(k, j) / d, inv(k, j) = inv(k, j) / d;
        for (int i = 0; i < n; i++) {
            if (i == k) continue;
            const T f = a(i, k);
            if (f == T(0)) continue;
            for (int j = 0; j < n; j++) a(i, j) = a(i, j) - f * a(k, j), inv(i, j) = inv(i, j) - f * inv(k, j);
        }
    }
    return inv;
}
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;

/* Cholesky factor A = L L^T (lower), of the symmetric part of the argument */
template <class M> class LLT {
    MatrixXd L_;
public:
    template <int O> explicit LLT(const DynBase<double, O> &A) {
        const int n = A.r;
        L_.resize(n, n);
        for (int j = 0; j < n; j++) {
            double s = 0.5 * (A(j, j) + A(j, j));
            for (int k = 0; k < j; k++) s -= L_(j, k) * L_(j, k);
            L_(j, j) = std::sqrt(s);
            for (int i = j + 1; i < n; i++) {
                double t = 0.5 * (A(i, j) + A(j, i));
                for (int k = 0; k < j; k++) t -= L_(i, k) * L_(j, k);
                L_(i, j) = t / L_(j, j);
            }
        }
    }
    const MatrixXd &matrixL() const { return L_; }
};

/* symmetric eigen-decomposition, eigenvalues ascending (cyclic Jacobi rotations on the lower / upper symmetric part) */
template <class M> class SelfAdjointEigenSolver {
    MatrixXd vec_;
    VectorXd val_;
public:
    template <int O> explicit SelfAdjointEigenSolver(const DynBase<double, O> &A) {
        const int n = A.r;
        MatrixXd a(n, n), V(n, n);
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) a(i, j) = 0.5 * (A(i, j) + A(j, i)), V(i, j) = i == j ? 1.0 : 0.0;
        for (int sweep = 0; sweep < 100; sweep++) {
            double off = 0.0, diag = 0.0;
            for (int i = 0; i < n; i++)
                for (int j = 0; j < n; j++) (i == j ? diag : off) += a(i, j) * a(i, j);
            if (off <= 1e-30 * diag || off == 0.0) break;
            for (int p = 0; p < n - 1; p++)
                for (int q = p + 1; q < n; q++) {
                    if (a(p, q) == 0.0) continue;
                    const double theta = (a(q, q) - a(p, p)) / (2.0 * a(p, q));
                    const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                    const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
                    for (int k = 0; k < n; k++) {
                        const double akp = a(k, p), akq = a(k, q);
                        a(k, p) = cs * akp - sn * akq, a(k, q) = sn * akp + cs * akq;
                    }
                    for (int k = 0; k < n; k++) {
                        const double apk = a(p, k), aqk = a(q, k);
                        a(p, k) = cs * apk - sn * aqk, a(q, k) = sn * apk + cs * aqk;
                    }
                    for (int k = 0; k < n; k++) {
                        const double vkp = V(k, p), vkq = V(k, q);
                        V(k, p) = cs * vkp - sn * vkq, V(k, q) = sn * vkp + cs * vkq;
                    }
                }
        }
        std::vector<int> order(n);
        for (int i = 0; i < n; i++) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int x, int y) { return a(x, x) < a(y, y); });
        val_.resize(n);
        vec_.resize(n, n);
        for (int k = 0; k < n; k++) {
            val_[k] = a(order[k], order[k]);
            for (int i = 0; i < n; i++) vec_(i, k) = V(i, order[k]);
        }
    }
    const VectorXd &eigenvalues() const { return val_; }
    const MatrixXd &eigenvectors() const { return vec_; }
};

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 3, 3> Matrix3f;

/* ---- Map: a view of caller memory with the matrix type's storage order ---- */
template <class M> struct Map;
template <class T, int R, int C, int Opt> struct Map<Matrix<T, R, C, Opt>> {
    T *p;
    explicit Map(T *ptr) : p(ptr) {}
    typedef Matrix<T, R, C> Plain;
    T &operator()(int r, int c) const { return p[mini::at<R, C, Opt>(r, c)]; }
    Plain eval() const {
        Plain o;
        for (int r = 0; r < R; r++)
            for (int c = 0; c < C; c++) o(r, c) = (*this)(r, c);
        return o;
    }
    operator Plain() const { return eval(); }
    void setZero() const { for (int i = 0; i < R * C; i++) p[i] = T(0); }
    template <int O> const Map &operator=(const Matrix<T, R, C, O> &m) const {
        for (int r = 0; r < R; r++)
            for (int c = 0; c < C; c++) (*this)(r, c) = m(r, c);
        return *this;
    }
    const Map &operator=(const Map &o) const {
        for (int i = 0; i < R * C; i++) p[i] = o.p[i];
        return *this;
    }
    template <class M2> const Map &operator=(const Map<M2> &o) const { return *this = o.eval(); }
    template <class U> Matrix<U, R, C> cast() const { return eval().template cast<U>(); }
    T squaredNorm() const { return eval().squaredNorm(); }
    T norm() const { return eval().norm(); }
    template <int BR, int BC> mini::View<T, BR, BC> block(int r, int c) const {
        return mini::View<T, BR, BC>{&(*this)(r, c), mini::at<R, C, Opt>(1, 0) - mini::at<R, C, Opt>(0, 0),
                                     (C > 1 ? mini::at<R, C, Opt>(0, 1) : 1) - mini::at<R, C, Opt>(0, 0)};
    }
    DynView<T> block(int i, int j, int nr, int nc) const {
        return DynView<T>{&(*this)(i, j), nr, nc, mini::at<R, C, Opt>(1, 0) - mini::at<R, C, Opt>(0, 0),
                          (C > 1 ? mini::at<R, C, Opt>(0, 1) : 1) - mini::at<R, C, Opt>(0, 0)};
    }
    template <int O> const Map &operator=(const DynBase<T, O> &m) const {
        for (int r = 0; r < R; r++)
            for (int c = 0; c < C; c++) (*this)(r, c) = m(r, c);
        return *this;
    }
    int rows() const { return R; }
    int cols() const { return C; }
    template <int N> mini::View<T, N, C> topRows() const { return block<N, C>(0, 0); }
    template <int N> mini::View<T, N, C> bottomRows() const { return block<N, C>(R - N, 0); }
    template <int N> mini::View<T, R, N> leftCols() const { return block<R, N>(0, 0); }
    template <int N> mini::View<T, R, N> rightCols() const { return block<R, N>(0, C - N); }
};
/* const maps (pcl's getVector3fMap() const) */
template <class T, int R, int C, int Opt> struct Map<const Matrix<T, R, C, Opt>> {
    const T *p;
    explicit Map(const T *ptr) : p(ptr) {}
    typedef Matrix<T, R, C> Plain;
    const T &operator()(int r, int c) const { return p[mini::at<R, C, Opt>(r, c)]; }
    Plain eval() const {
        Plain o;
        for (int r = 0; r < R; r++)
            for (int c = 0; c < C; c++) o(r, c) = (*this)(r, c);
        return o;
    }
    operator Plain() const { return eval(); }
    template <class U> Matrix<U, R, C> cast() const { return eval().template cast<U>(); }
    T squaredNorm() const { return eval().squaredNorm(); }
    T norm() const { return eval().norm(); }
};
template <class M1, class M2> typename Map<M1>::Plain operator-(const Map<M1> &a, const Map<M2> &b) {
    return a.eval() - b.eval();
}
template <class M1, class M2> typename Map<M1>::Plain operator+(const Map<M1> &a, const Map<M2> &b) {
    return a.eval() + b.eval();
}

template <class T, int O, class M> Matrix<T, Dynamic, Dynamic> operator*(const DynBase<T, O> &a, const Map<M> &b) {
    return a * Matrix<T, Dynamic, Dynamic>(b);
}

/* ---- AngleAxis / Quaternion: Eigen 3.4 Geometry module formulas (generic, non-vectorised paths) ---- */
template <class T> struct AngleAxis {
    T a;
    Matrix<T, 3, 1> ax;
    AngleAxis() : a(0) {}
    template <int O> AngleAxis(T angle, const Matrix<T, 3, 1, O> &axis) : a(angle), ax(axis) {}
    /* AngleAxis::operator=(QuaternionBase): n = |vec|; angle = 2 atan2(n, |w|), axis = vec / (w < 0 ? -n : n) */
    explicit AngleAxis(const Quaternion<T> &q) {
        T n = q.vec().norm();
        if (n < std::numeric_limits<T>::epsilon()) n = q.vec().stableNorm();
        if (n != T(0)) {
            a = T(2) * std::atan2(n, std::abs(q.w()));
            if (q.w() < T(0)) n = -n;
            ax = q.vec() / n;
        } else {
            a  = T(0);
            ax = Matrix<T, 3, 1>(T(1), T(0), T(0));
        }
    }
    /* AngleAxis::fromRotationMatrix: through the quaternion */
    explicit AngleAxis(const Matrix<T, 3, 3> &m) { *this = AngleAxis(Quaternion<T>(m)); }
    T angle() const { return a; }
    const Matrix<T, 3, 1> &axis() const { return ax; }
    Quaternion<T> operator*(const AngleAxis &o) const { return Quaternion<T>(*this) * Quaternion<T>(o); }
    Quaternion<T> operator*(const Quaternion<T> &o) const { return Quaternion<T>(*this) * o; }
};

template <class T> struct Quaternion {
    Matrix<T, 4, 1> cf; /* x, y, z, w as Eigen stores them */
    Quaternion() { cf.m[0] = cf.m[1] = cf.m[2] = cf.m[3] = T(0); }
    Quaternion(T w, T x, T y, T z) { cf.m[0] = x, cf.m[1] = y, cf.m[2] = z, cf.m[3] = w; }
    /* from angle-axis: w = cos(a/2), vec = sin(a/2) * axis */
    explicit Quaternion(const AngleAxis<T> &aa) {
        T ha = T(0.5) * aa.angle();
        cf.m[3] = std::cos(ha);
        Matrix<T, 3, 1> v = std::sin(ha) * aa.axis();
        cf.m[0] = v[0], cf.m[1] = v[1], cf.m[2] = v[2];
    }
    /* from a rotation matrix (quaternionbase_assign_impl<Other,3,3>, Shoemake's branches) */
    explicit Quaternion(const Matrix<T, 3, 3> &mat) {
        T t = mat(0, 0) + mat(1, 1) + mat(2, 2);
        if (t > T(0)) {
            t    = std::sqrt(t + T(1.0));
            cf.m[3] = T(0.5) * t;
            t    = T(0.5) / t;
            cf.m[0] = (mat(2, 1) - mat(1, 2)) * t;
            cf.m[1] = (mat(0, 2) - mat(2, 0)) * t;
            cf.m[2] = (mat(1, 0) - mat(0, 1)) * t;
        } else {
            int i = 0;
            if (mat(1, 1) > mat(0, 0)) i = 1;
            if (mat(2, 2) > mat(i, i)) i = 2;
            int j = (i + 1) % 3, k = (j + 1) % 3;
            t    = std::sqrt(mat(i, i) - mat(j, j) - mat(k, k) + T(1.0));
            cf.m[i] = T(0.5) * t;
            t    = T(0.5) / t;
            cf.m[3] = (mat(k, j) - mat(j, k)) * t;
            cf.m[j] = (mat(j, i) + mat(i, j)) * t;
            cf.m[k] = (mat(k, i) + mat(i, k)) * t;
        }
    }
    T &w() { return cf.m[3]; }
    T &x() { return cf.m[0]; }
    T &y() { return cf.m[1]; }
    T &z() { return cf.m[2]; }
    const T &w() const { return cf.m[3]; }
    const T &x() const { return cf.m[0]; }
    const T &y() const { return cf.m[1]; }
    const T &z() const { return cf.m[2]; }
    Matrix<T, 3, 1> vec() const { return Matrix<T, 3, 1>(cf.m[0], cf.m[1], cf.m[2]); }
    Matrix<T, 4, 1> &coeffs() { return cf; }
    const Matrix<T, 4, 1> &coeffs() const { return cf; }
    void setIdentity() { cf.m[0] = cf.m[1] = cf.m[2] = T(0), cf.m[3] = T(1); }
    T squaredNorm() const { return coeffs().squaredNorm(); }
    T norm() const { return coeffs().norm(); }
    void normalize() {
        T n = norm(); /* coeffs().normalize() */
        T n2 = squaredNorm();
        if (n2 > T(0))
            for (int i = 0; i < 4; i++) cf.m[i] = cf.m[i] / n;
    }
    Quaternion normalized() const { Quaternion q = *this; q.normalize(); return q; }
    Quaternion conjugate() const { return Quaternion(cf.m[3], -cf.m[0], -cf.m[1], -cf.m[2]); }
    Quaternion operator*(const AngleAxis<T> &aa) const { return *this * Quaternion(aa); }
    Quaternion &operator*=(const Quaternion &o) { *this = *this * o; return *this; }
    /* QuaternionBase::inverse(): conjugate / squaredNorm when that is positive, else zero */
    Quaternion inverse() const {
        T n2 = squaredNorm();
        if (n2 > T(0)) {
            Quaternion q = conjugate();
            for (int i = 0; i < 4; i++) q.cf.m[i] = q.cf.m[i] / n2;
            return q;
        }
        return Quaternion();
    }
    /* quat_product, generic path */
    Quaternion operator*(const Quaternion &b) const {
        const Quaternion &a = *this;
        return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                          a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                          a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                          a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
    }
    /* QuaternionBase::toRotationMatrix() */
    Matrix<T, 3, 3> toRotationMatrix() const {
        Matrix<T, 3, 3> res;
        const T tx = T(2) * x(), ty = T(2) * y(), tz = T(2) * z();
        const T twx = tx * w(), twy = ty * w(), twz = tz * w();
        const T txx = tx * x(), txy = ty * x(), txz = tz * x();
        const T tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
        res(0, 0) = T(1) - (tyy + tzz);
        res(0, 1) = txy - twz;
        res(0, 2) = txz + twy;
        res(1, 0) = txy + twz;
        res(1, 1) = T(1) - (txx + tzz);
        res(1, 2) = tyz - twx;
        res(2, 0) = txz - twy;
        res(2, 1) = tyz + twx;
        res(2, 2) = T(1) - (txx + tyy);
        return res;
    }
    Matrix<T, 3, 1> operator*(const Matrix<T, 3, 1> &v) const {
        /* QuaternionBase::_transformVector: v + w * uv + vec x uv, uv = 2 vec x v */
        Matrix<T, 3, 1> uv = vec().cross(v);
        uv += uv;
        return v + w() * uv + vec().cross(uv);
    }
};

typedef Quaternion<double> Quaterniond;
typedef AngleAxis<double> AngleAxisd;

/* maps of quaternions stored x, y, z, w in caller memory (pose_manifold.cc): the const map is a copy that behaves as a
 * Quaternion, the mutable one accepts an assignment */
template <class T> struct Map<const Quaternion<T>> : Quaternion<T> {
    explicit Map(const T *p) : Quaternion<T>(p[3], p[0], p[1], p[2]) {}
};
/* maps of run-time sized vectors / row-major matrices over caller memory (marginalization_factor.cc) */
template <class T> struct Map<const Matrix<T, Dynamic, 1>> : Matrix<T, Dynamic, 1> {
    Map(const T *p, int n) {
        this->resize(n);
        for (int i = 0; i < n; i++) this->v[i] = p[i];
    }
};
template <class T> struct Map<Matrix<T, Dynamic, 1>> {
    T *p;
    int n;
    Map(T *ptr, int size) : p(ptr), n(size) {}
    template <int O> const Map &operator=(const DynBase<T, O> &m) const {
        for (int i = 0; i < n; i++) p[i] = m(i, 0);
        return *this;
    }
};
template <class T> struct Map<Matrix<T, Dynamic, Dynamic, RowMajor>> {
    T *p;
    int r, c;
    Map(T *ptr, int rows, int cols) : p(ptr), r(rows), c(cols) {}
    void setZero() const { for (int i = 0; i < r * c; i++) p[i] = T(0); }
    DynView<T> leftCols(int n) const { return DynView<T>{p, r, n, c, 1}; }
};
template <class T> struct Map<Quaternion<T>> {
    T *p;
    explicit Map(T *ptr) : p(ptr) {}
    const Map &operator=(const Quaternion<T> &q) const {
        p[0] = q.x(), p[1] = q.y(), p[2] = q.z(), p[3] = q.w();
        return *this;
    }
};

/* ---- A.colPivHouseholderQr().solve(b) for a tall fixed-size A: Householder reflections with column pivoting on the
 * largest remaining column norm, then back substitution through the permutation (least squares). Written here from the
 * textbook algorithm, independently of oracle.cc's restatement. ---- */
template <class T, int R, int C, int Opt> struct Matrix<T, R, C, Opt>::ColPivQR {
    Matrix<T, R, C, Opt> A;
    Matrix<T, C, 1> solve(const Matrix<T, R, 1> &rhs) const {
        T a[R][C], b[R];
        int perm[C];
        for (int r = 0; r < R; r++) {
            b[r] = rhs[r];
            for (int c = 0; c < C; c++) a[r][c] = A(r, c);
        }
        for (int c = 0; c < C; c++) perm[c] = c;
        int rank = 0;
        T maxnorm2 = T(0);
        for (int c = 0; c < C; c++) {
            T s = T(0);
            for (int r = 0; r < R; r++) s += a[r][c] * a[r][c];
            maxnorm2 = std::max(maxnorm2, s);
        }
        const T thresh = std::numeric_limits<T>::epsilon() * T(R) * std::sqrt(maxnorm2);
        for (int k = 0; k < C; k++) {
            int best = k;
            T bestn = T(-1);
            for (int c = k; c < C; c++) {
                T s = T(0);
                for (int r = k; r < R; r++) s += a[r][c] * a[r][c];
                if (s > bestn) bestn = s, best = c;
            }
            if (best != k) {
                std::swap(perm[k], perm[best]);
                for (int r = 0; r < R; r++) std::swap(a[r][k], a[r][best]);
            }
            T nrm = std::sqrt(bestn);
            if (!(nrm > thresh)) break;
            rank++;
            T alpha = a[k][k] > T(0) ? -nrm : nrm;
            T v[R];
            for (int r = 0; r < R; r++) v[r] = r < k ? T(0) : a[r][k];
            v[k] -= alpha;
            T vv = T(0);
            for (int r = k; r < R; r++) vv += v[r] * v[r];
            if (vv > T(0)) {
                for (int c = k; c < C; c++) {
                    T d = T(0);
                    for (int r = k; r < R; r++) d += v[r] * a[r][c];
                    d = T(2) * d / vv;
                    for (int r = k; r < R; r++) a[r][c] -= d * v[r];
                }
                T d = T(0);
                for (int r = k; r < R; r++) d += v[r] * b[r];
                d = T(2) * d / vv;
                for (int r = k; r < R; r++) b[r] -= d * v[r];
            }
        }
        T y[C];
        for (int c = 0; c < C; c++) y[c] = T(0);
        for (int k = rank - 1; k >= 0; k--) {
            T s = b[k];
            for (int c = k + 1; c < rank; c++) s -= a[k][c] * y[c];
            y[k] = s / a[k][k];
        }
        Matrix<T, C, 1> x;
        for (int c = 0; c < C; c++) x[perm[c]] = y[c];
        return x;
    }
};

} // namespace Eigen
#endif
