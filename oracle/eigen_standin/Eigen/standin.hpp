// oracle/eigen_standin/Eigen/standin.hpp -- TEST INFRASTRUCTURE ONLY.
//
// A minimal stand-in for the part of Eigen's API that the reference's hot-path sources use (framework/*.cpp, slam/*.cpp,
// utils/g2o_utils.cpp under /root/reference).  Eigen3 itself is not in this image and there is no network, so the reference cannot
// be built against the real library; with this header its OWN, UNMODIFIED source files compile in place (oracle/Makefile, target
// `_ref`) and their results pin the restatement in bos_oracle.hpp (tests/test_ref_build.py, tests/golden/ref_*.npz).
//
// What is the reference's and what is the stand-in's in such a build:
//   * the reference's: every formula, sign, operand order, angle wrap, robust kernel, damping, gauge permutation, accumulation order
//     of H and b, boxplus, the triangulation equations, the g2o parser, the id <-> stix maps;
//   * the stand-in's: the arithmetic behind the operators (plain loops, float, no FMA contraction), the sparse containers, the
//     LDL^T factorisation (up-looking, natural order, no pivoting -- what SimplicialLDLT does, minus its fill-reducing ordering) and the
//     column-pivoting Householder QR (written after Eigen's published algorithm: pivot on the largest remaining column norm, stop
//     at columns whose squared norm falls below (eps * max column norm)^2 * (rows - k) / rows, basic solution for the rest).
// Nothing here is copied from Eigen; it is written against the documented behaviour of the calls the reference makes.
// Value semantics throughout (every operator returns a plain matrix): no expression templates, no aliasing rules to honour.
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <type_traits>
#include <utility>
#include <vector>

namespace Eigen {

constexpr int Dynamic = -1;
enum NoChange_t { NoChange };
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };
template <class T>
using aligned_allocator = std::allocator<T>;
typedef std::ptrdiff_t Index;

template <class D> struct traits;
template <typename S, int R, int C> class Matrix;
template <class X, int BR, int BC> class Block;
template <typename S> class SparseMatrix;
template <class M> class ColPivHouseholderQR;

// ------------------------------------------------------------------------------------------------ dense
template <class D>
class DenseBase {
   public:
    typedef typename traits<D>::Scalar Scalar;
    static constexpr int Rows = traits<D>::Rows, Cols = traits<D>::Cols;
    typedef Matrix<Scalar, Rows, Cols> PlainObject;
    const D& derived() const { return *static_cast<const D*>(this); }
    D& derived() { return *static_cast<D*>(this); }

    int size() const { return derived().rows() * derived().cols(); }
    PlainObject eval() const {
        PlainObject m;
        m.resize(derived().rows(), derived().cols());
        for (int j = 0; j < derived().cols(); j++)
            for (int i = 0; i < derived().rows(); i++) m.coeffRef(i, j) = derived().coeff(i, j);
        return m;
    }
    Scalar operator()(int i, int j) const { return derived().coeff(i, j); }
    Scalar& operator()(int i, int j) { return derived().coeffRef(i, j); }
    // vectors (either orientation) are indexed linearly
    Scalar operator()(int i) const { return derived().cols() == 1 ? derived().coeff(i, 0) : derived().coeff(0, i); }
    Scalar& operator()(int i) { return derived().cols() == 1 ? derived().coeffRef(i, 0) : derived().coeffRef(0, i); }
    Scalar operator[](int i) const { return (*this)(i); }
    Scalar& operator[](int i) { return (*this)(i); }
    Scalar x() const { return (*this)(0); }
    Scalar y() const { return (*this)(1); }
    Scalar z() const { return (*this)(2); }
    Scalar& x() { return (*this)(0); }
    Scalar& y() { return (*this)(1); }
    Scalar& z() { return (*this)(2); }

    Matrix<Scalar, Cols, Rows> transpose() const {
        Matrix<Scalar, Cols, Rows> m;
        m.resize(derived().cols(), derived().rows());
        for (int j = 0; j < derived().cols(); j++)
            for (int i = 0; i < derived().rows(); i++) m.coeffRef(j, i) = derived().coeff(i, j);
        return m;
    }

    // sub-blocks: a proxy on a mutable object, a copy on a const one
    template <int BR, int BC> Block<D, BR, BC> block(int i, int j) { return Block<D, BR, BC>(derived(), i, j, BR, BC); }
    template <int BR, int BC> Matrix<Scalar, BR, BC> block(int i, int j) const { return copy_block<BR, BC>(i, j, BR, BC); }
    template <int N> Block<D, N, 1> head() { return Block<D, N, 1>(derived(), 0, 0, N, 1); }
    template <int N> Matrix<Scalar, N, 1> head() const { return copy_block<N, 1>(0, 0, N, 1); }
    template <int N> Block<D, N, 1> tail() { return Block<D, N, 1>(derived(), size() - N, 0, N, 1); }
    template <int N> Matrix<Scalar, N, 1> tail() const { return copy_block<N, 1>(size() - N, 0, N, 1); }
    template <int N> Block<D, N, 1> segment(int i) { return Block<D, N, 1>(derived(), i, 0, N, 1); }
    template <int N> Matrix<Scalar, N, 1> segment(int i) const { return copy_block<N, 1>(i, 0, N, 1); }
    Block<D, Dynamic, 1> head(int n) { return Block<D, Dynamic, 1>(derived(), 0, 0, n, 1); }
    Matrix<Scalar, Dynamic, 1> head(int n) const { return copy_block<Dynamic, 1>(0, 0, n, 1); }
    Block<D, 1, Cols> row(int i) { return Block<D, 1, Cols>(derived(), i, 0, 1, derived().cols()); }
    Matrix<Scalar, 1, Cols> row(int i) const { return copy_block<1, Cols>(i, 0, 1, derived().cols()); }
    Block<D, Rows, 1> col(int j) { return Block<D, Rows, 1>(derived(), 0, j, derived().rows(), 1); }
    Matrix<Scalar, Rows, 1> col(int j) const { return copy_block<Rows, 1>(0, j, derived().rows(), 1); }

    D& setZero() {
        for (int j = 0; j < derived().cols(); j++)
            for (int i = 0; i < derived().rows(); i++) derived().coeffRef(i, j) = Scalar(0);
        return derived();
    }
    D& setIdentity() {
        for (int j = 0; j < derived().cols(); j++)
            for (int i = 0; i < derived().rows(); i++) derived().coeffRef(i, j) = Scalar(i == j ? 1 : 0);
        return derived();
    }
    template <class E> D& operator+=(const DenseBase<E>& o) {
        assert(o.size() == size());
        PlainObjectOf<E> t = o.eval();
        for (int j = 0; j < derived().cols(); j++)
            for (int i = 0; i < derived().rows(); i++) derived().coeffRef(i, j) += t.coeff(i, j);
        return derived();
    }
    template <class E> D& operator-=(const DenseBase<E>& o) {
        PlainObjectOf<E> t = o.eval();
        for (int j = 0; j < derived().cols(); j++)
            for (int i = 0; i < derived().rows(); i++) derived().coeffRef(i, j) -= t.coeff(i, j);
        return derived();
    }
    D& operator*=(Scalar s) {
        for (int j = 0; j < derived().cols(); j++)
            for (int i = 0; i < derived().rows(); i++) derived().coeffRef(i, j) *= s;
        return derived();
    }
    D& operator/=(Scalar s) {
        for (int j = 0; j < derived().cols(); j++)
            for (int i = 0; i < derived().rows(); i++) derived().coeffRef(i, j) /= s;
        return derived();
    }
    // dense += sparse (the reference adds a sparse N x 1 column to its dense b, slam/solver.cpp:45)
    D& operator+=(const SparseMatrix<Scalar>& s);

    Scalar norm() const { return std::sqrt(squaredNorm()); }
    Scalar squaredNorm() const {
        Scalar a = 0;
        for (int j = 0; j < derived().cols(); j++)
            for (int i = 0; i < derived().rows(); i++) a += derived().coeff(i, j) * derived().coeff(i, j);
        return a;
    }
    Scalar dot(const PlainObject& o) const {
        Scalar a = 0;
        for (int i = 0; i < size(); i++) a += (*this)(i) * o(i);
        return a;
    }
    ColPivHouseholderQR<PlainObject> colPivHouseholderQr() const { return ColPivHouseholderQR<PlainObject>(eval()); }

   protected:
    template <class E> using PlainObjectOf = Matrix<typename traits<E>::Scalar, traits<E>::Rows, traits<E>::Cols>;
    template <int BR, int BC> Matrix<Scalar, BR, BC> copy_block(int i0, int j0, int br, int bc) const {
        Matrix<Scalar, BR, BC> m;
        m.resize(br, bc);
        for (int j = 0; j < bc; j++)
            for (int i = 0; i < br; i++) m.coeffRef(i, j) = derived().coeff(i0 + i, j0 + j);
        return m;
    }
    // element-wise copy; a vector may be assigned to a vector of the other orientation (Eigen transposes those implicitly)
    template <class E> void assign_from(const DenseBase<E>& o) {
        PlainObjectOf<E> t = o.eval();
        const int r = derived().rows(), c = derived().cols();
        if (t.rows() == r && t.cols() == c) {
            for (int j = 0; j < c; j++)
                for (int i = 0; i < r; i++) derived().coeffRef(i, j) = t.coeff(i, j);
        } else {
            assert((r == 1 || c == 1) && (t.rows() == 1 || t.cols() == 1) && r * c == t.size());
            for (int i = 0; i < r * c; i++) (*this)(i) = t(i);
        }
    }
};

template <typename S, int R, int C, bool Dyn = (R == Dynamic || C == Dynamic)>
struct DenseStorage;
template <typename S, int R, int C>
struct DenseStorage<S, R, C, false> {
    S d[R * C] = {};
    int rows() const { return R; }
    int cols() const { return C; }
    void resize(int r, int c) { assert(r == R && c == C); (void)r; (void)c; }
};
template <typename S, int R, int C>
struct DenseStorage<S, R, C, true> {
    std::vector<S> d;
    int r = (R == Dynamic ? 0 : R), c = (C == Dynamic ? 0 : C);
    int rows() const { return r; }
    int cols() const { return c; }
    void resize(int r_, int c_) { r = r_; c = c_; d.assign((size_t)r * c, S(0)); }
};

template <typename S, int R, int C>
struct traits<Matrix<S, R, C>> {
    typedef S Scalar;
    static constexpr int Rows = R, Cols = C;
};

template <typename S, int R, int C>
class Matrix : public DenseBase<Matrix<S, R, C>> {
    typedef DenseBase<Matrix<S, R, C>> Base;
    DenseStorage<S, R, C> st;

   public:
    typedef S Scalar;
    Matrix() {}
    Matrix(const Matrix&) = default;
    Matrix& operator=(const Matrix&) = default;
    Matrix(S a, S b) {
        static_assert(R * C == 2, "two-coefficient constructor is for 2-vectors");
        st.d[0] = a; st.d[1] = b;
    }
    Matrix(S a, S b, S c) {
        static_assert(R * C == 3, "three-coefficient constructor is for 3-vectors");
        st.d[0] = a; st.d[1] = b; st.d[2] = c;
    }
    template <class E> Matrix(const DenseBase<E>& o) { resize_like(o); Base::assign_from(o); }
    template <class E> Matrix& operator=(const DenseBase<E>& o) { resize_like(o); Base::assign_from(o); return *this; }

    int rows() const { return st.rows(); }
    int cols() const { return st.cols(); }
    S coeff(int i, int j) const { assert(i >= 0 && i < rows() && j >= 0 && j < cols()); return st.d[i + (size_t)j * rows()]; }
    S& coeffRef(int i, int j) { assert(i >= 0 && i < rows() && j >= 0 && j < cols()); return st.d[i + (size_t)j * rows()]; }
    const S* data() const { return &st.d[0]; }
    S* data() { return &st.d[0]; }

    void resize(int r, int c) { st.resize(r, c); }
    void resize(int n) {
        static_assert(R == 1 || C == 1, "one-argument resize is for vectors");
        if (C == 1) st.resize(n, 1); else st.resize(1, n);
    }
    void resize(int r, NoChange_t) { st.resize(r, cols()); }
    void resize(NoChange_t, int c) { st.resize(rows(), c); }

    static Matrix Zero() { Matrix m; m.setZero(); return m; }
    static Matrix Zero(int n) { Matrix m; m.resize(n); return m; }
    static Matrix Identity() { Matrix m; m.setIdentity(); return m; }

    // a 1 x 1 result converts to its scalar (`float chi = e.transpose() * Omega * e;`, slam/solver.cpp:55)
    template <int R_ = R, int C_ = C, typename = typename std::enable_if<R_ == 1 && C_ == 1>::type>
    operator S() const { return st.d[0]; }

   private:
    template <class E> void resize_like(const DenseBase<E>& o) {
        const int r = o.derived().rows(), c = o.derived().cols();
        if (R == Dynamic || C == Dynamic) {
            // a dynamic vector takes the size of a vector of either orientation
            if (R != Dynamic && R != r && c == R) st.resize(c, r);
            else if (C != Dynamic && C != c && r == C) st.resize(c, r);
            else st.resize(r, c);
        }
    }
};

template <class X, int BR, int BC>
struct traits<Block<X, BR, BC>> {
    typedef typename traits<X>::Scalar Scalar;
    static constexpr int Rows = BR, Cols = BC;
};

template <class X, int BR, int BC>
class Block : public DenseBase<Block<X, BR, BC>> {
    typedef DenseBase<Block<X, BR, BC>> Base;
    X& x;
    int i0, j0, br, bc;

   public:
    typedef typename traits<X>::Scalar Scalar;
    Block(X& x_, int i, int j, int r, int c) : x(x_), i0(i), j0(j), br(r), bc(c) {
        assert(i >= 0 && j >= 0 && i + r <= x.rows() && j + c <= x.cols());
    }
    Block(const Block&) = default;
    int rows() const { return br; }
    int cols() const { return bc; }
    Scalar coeff(int i, int j) const { return const_cast<const X&>(x).coeff(i0 + i, j0 + j); }
    Scalar& coeffRef(int i, int j) { return x.coeffRef(i0 + i, j0 + j); }
    template <class E> Block& operator=(const DenseBase<E>& o) { Base::assign_from(o); return *this; }
    Block& operator=(const Block& o) { Base::assign_from(o); return *this; }
};

// ---- arithmetic (all by value) --------------------------------------------------------------------
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator+(const DenseBase<A>& a, const DenseBase<B>& b) {
    auto m = a.eval();
    m += b;
    return m;
}
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator-(const DenseBase<A>& a, const DenseBase<B>& b) {
    auto m = a.eval();
    m -= b;
    return m;
}
template <class A>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator-(const DenseBase<A>& a) {
    auto m = a.eval();
    for (int j = 0; j < m.cols(); j++)
        for (int i = 0; i < m.rows(); i++) m.coeffRef(i, j) = -m.coeff(i, j);
    return m;
}
template <class A>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator*(const DenseBase<A>& a, typename traits<A>::Scalar s) {
    auto m = a.eval();
    m *= s;
    return m;
}
template <class A>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator*(typename traits<A>::Scalar s, const DenseBase<A>& a) {
    auto m = a.eval();
    for (int j = 0; j < m.cols(); j++)
        for (int i = 0; i < m.rows(); i++) m.coeffRef(i, j) = s * m.coeff(i, j);
    return m;
}
template <class A>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator/(const DenseBase<A>& a, typename traits<A>::Scalar s) {
    auto m = a.eval();
    m /= s;
    return m;
}
// matrix product: the inner sum runs over k in increasing order, accumulated in Scalar
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> operator*(const DenseBase<A>& a_, const DenseBase<B>& b_) {
    auto a = a_.eval();
    auto b = b_.eval();
    assert(a.cols() == b.rows());
    Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> m;
    m.resize(a.rows(), b.cols());
    for (int j = 0; j < b.cols(); j++)
        for (int i = 0; i < a.rows(); i++) {
            typename traits<A>::Scalar s = 0;
            for (int k = 0; k < a.cols(); k++) s += a.coeff(i, k) * b.coeff(k, j);
            m.coeffRef(i, j) = s;
        }
    return m;
}
template <class A>
std::ostream& operator<<(std::ostream& os, const DenseBase<A>& a) {
    for (int i = 0; i < a.derived().rows(); i++) {
        for (int j = 0; j < a.derived().cols(); j++) os << (j ? " " : "") << a.derived().coeff(i, j);
        if (i + 1 < a.derived().rows()) os << "\n";
    }
    return os;
}

typedef Matrix<float, 2, 2> Matrix2f;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, Dynamic, 1> VectorXf;
typedef Matrix<int, Dynamic, 1> VectorXi;
typedef Matrix<float, Dynamic, Dynamic> MatrixXf;

// ---- column-pivoting Householder QR (slam/triangulation.cpp:56 is the only user) ---------------------
template <class M>
class ColPivHouseholderQR {
    typedef typename M::Scalar S;
    M qr;
    std::vector<S> tau;
    std::vector<int> perm;   // perm[k] = original column now at position k
    int nonzero_pivots = 0;

   public:
    explicit ColPivHouseholderQR(const M& a) : qr(a) {
        const int rows = qr.rows(), cols = qr.cols(), size = std::min(rows, cols);
        tau.assign(size, S(0));
        perm.resize(cols);
        std::vector<S> norms(cols);
        S maxnorm = 0;
        for (int j = 0; j < cols; j++) {
            perm[j] = j;
            S s = 0;
            for (int i = 0; i < rows; i++) s += qr.coeff(i, j) * qr.coeff(i, j);
            norms[j] = std::sqrt(s);
            maxnorm = std::max(maxnorm, norms[j]);
        }
        const S eps = std::numeric_limits<S>::epsilon();
        const S threshold_helper = (maxnorm * eps) * (maxnorm * eps) / S(rows);
        nonzero_pivots = size;
        for (int k = 0; k < size; k++) {
            int big = k;
            for (int j = k + 1; j < cols; j++)
                if (norms[j] > norms[big]) big = j;
            // the column norms are RECOMPUTED on the trailing rows each step (Eigen down-dates them and recomputes when the
            // down-date loses accuracy; with at most two columns the difference is a rounding of the second pivot test only)
            if (nonzero_pivots == size && norms[big] * norms[big] < threshold_helper * S(rows - k)) nonzero_pivots = k;
            if (big != k) {
                for (int i = 0; i < rows; i++) std::swap(qr.coeffRef(i, k), qr.coeffRef(i, big));
                std::swap(norms[k], norms[big]);
                std::swap(perm[k], perm[big]);
            }
            // Householder reflector of column k, rows k..: H = I - tau v v^T, v = (1, essential)
            S tail = 0;
            for (int i = k + 1; i < rows; i++) tail += qr.coeff(i, k) * qr.coeff(i, k);
            const S c0 = qr.coeff(k, k);
            S beta;
            if (tail <= std::numeric_limits<S>::min()) {
                tau[k] = 0;
                beta = c0;
                for (int i = k + 1; i < rows; i++) qr.coeffRef(i, k) = 0;
            } else {
                beta = std::sqrt(c0 * c0 + tail);
                if (c0 >= 0) beta = -beta;
                for (int i = k + 1; i < rows; i++) qr.coeffRef(i, k) /= (c0 - beta);
                tau[k] = (beta - c0) / beta;
            }
            qr.coeffRef(k, k) = beta;
            for (int j = k + 1; j < cols; j++) {
                S w = qr.coeff(k, j);
                for (int i = k + 1; i < rows; i++) w += qr.coeff(i, k) * qr.coeff(i, j);
                w *= tau[k];
                qr.coeffRef(k, j) -= w;
                for (int i = k + 1; i < rows; i++) qr.coeffRef(i, j) -= w * qr.coeff(i, k);
                S s = 0;
                for (int i = k + 1; i < rows; i++) s += qr.coeff(i, j) * qr.coeff(i, j);
                norms[j] = std::sqrt(s);
            }
        }
    }
    int nonzeroPivots() const { return nonzero_pivots; }
    // least-squares solution; columns beyond the non-zero pivots get 0 (Eigen's "basic" solution for a rank-deficient system)
    template <class E>
    Matrix<S, Dynamic, 1> solve(const DenseBase<E>& b_) const {
        const int rows = qr.rows(), cols = qr.cols();
        Matrix<S, Dynamic, 1> c = b_.eval();
        assert(c.size() == rows);
        for (int k = 0; k < nonzero_pivots; k++) {
            S w = c(k);
            for (int i = k + 1; i < rows; i++) w += qr.coeff(i, k) * c(i);
            w *= tau[k];
            c(k) -= w;
            for (int i = k + 1; i < rows; i++) c(i) -= w * qr.coeff(i, k);
        }
        for (int k = nonzero_pivots - 1; k >= 0; k--) {
            S s = c(k);
            for (int j = k + 1; j < nonzero_pivots; j++) s -= qr.coeff(k, j) * c(j);
            c(k) = s / qr.coeff(k, k);
        }
        Matrix<S, Dynamic, 1> x;
        x.resize(cols);
        for (int k = 0; k < nonzero_pivots; k++) x(perm[k]) = c(k);
        return x;
    }
};

// ------------------------------------------------------------------------------------------------ geometry
template <typename S>
class Rotation2D {
    S a;

   public:
    Rotation2D() : a(0) {}
    explicit Rotation2D(const S& angle) : a(angle) {}
    template <class E> explicit Rotation2D(const DenseBase<E>& m) : a(std::atan2(m(1, 0), m(0, 0))) {}
    S angle() const { return a; }
    S& angle() { return a; }
    // the angle folded into [-pi, pi]
    S smallestAngle() const {
        const S pi = S(3.141592653589793238462643383279502884L);
        S t = std::fmod(a, S(2) * pi);
        if (t > pi) t -= S(2) * pi;
        else if (t < -pi) t += S(2) * pi;
        return t;
    }
    Matrix<S, 2, 2> toRotationMatrix() const {
        const S s = std::sin(a), c = std::cos(a);
        Matrix<S, 2, 2> m;
        m(0, 0) = c; m(0, 1) = -s; m(1, 0) = s; m(1, 1) = c;
        return m;
    }
    Matrix<S, 2, 2> matrix() const { return toRotationMatrix(); }
};

// Transform<float, 2, Isometry>: x -> linear * x + translation
template <typename S>
class Isometry2 {
    Matrix<S, 2, 2> R;
    Matrix<S, 2, 1> t;

   public:
    Isometry2() {}
    void setIdentity() { R.setIdentity(); t.setZero(); }
    static Isometry2 Identity() { Isometry2 x; x.setIdentity(); return x; }
    const Matrix<S, 2, 2>& linear() const { return R; }
    Matrix<S, 2, 2>& linear() { return R; }
    const Matrix<S, 2, 2>& rotation() const { return R; }   // Isometry mode: the linear part IS the rotation
    const Matrix<S, 2, 1>& translation() const { return t; }
    Matrix<S, 2, 1>& translation() { return t; }
    Isometry2 inverse() const {   // Isometry mode: R^T, -R^T t
        Isometry2 x;
        x.R = R.transpose();
        x.t = -(x.R * t);
        return x;
    }
    Matrix<S, 2, 1> operator*(const Matrix<S, 2, 1>& v) const { return R * v + t; }
    Isometry2 operator*(const Isometry2& o) const {
        Isometry2 x;
        x.R = R * o.R;
        x.t = R * o.t + t;
        return x;
    }
    Matrix<S, 3, 3> matrix() const {
        Matrix<S, 3, 3> m;
        m.setIdentity();
        m.template block<2, 2>(0, 0) = R;
        m.template block<2, 1>(0, 2) = t;
        return m;
    }
};
typedef Isometry2<float> Isometry2f;

// ------------------------------------------------------------------------------------------------ sparse
template <typename S>
class Triplet {
    int r, c;
    S v;

   public:
    Triplet() : r(0), c(0), v(0) {}
    Triplet(int i, int j, const S& x = S(0)) : r(i), c(j), v(x) {}
    int row() const { return r; }
    int col() const { return c; }
    const S& value() const { return v; }
};

template <class P> struct InversePermutation { const P& p; };

template <int N = Dynamic>
class PermutationMatrix {
    VectorXi idx;   // P(idx[i], i) = 1: row i of the operand goes to row idx[i]

   public:
    PermutationMatrix() {}
    explicit PermutationMatrix(const VectorXi& indices) : idx(indices) {}
    const VectorXi& indices() const { return idx; }
    int size() const { return idx.size(); }
    InversePermutation<PermutationMatrix> transpose() const { return InversePermutation<PermutationMatrix>{*this}; }
    InversePermutation<PermutationMatrix> inverse() const { return InversePermutation<PermutationMatrix>{*this}; }
};

// Column-major sparse matrix; an entry that was ever written stays in the pattern, also with value 0 (as in Eigen: setFromTriplets,
// + and * keep explicit zeros, which is what makes the pattern of H the union of dense per-edge blocks).
template <typename S>
class SparseMatrix {
    int nr = 0, nc = 0;
    std::vector<std::map<int, S>> colv;   // per column: row -> value, rows ascending

   public:
    typedef S Scalar;
    SparseMatrix() {}
    SparseMatrix(int r, int c) { resize(r, c); }
    void resize(int r, int c) { nr = r; nc = c; colv.assign(c, std::map<int, S>()); }
    int rows() const { return nr; }
    int cols() const { return nc; }
    int outerSize() const { return nc; }
    long nonZeros() const { long n = 0; for (auto& c : colv) n += (long)c.size(); return n; }
    void setZero() { for (auto& c : colv) c.clear(); }
    void setIdentity() { setZero(); for (int i = 0; i < std::min(nr, nc); i++) colv[i][i] = S(1); }
    void makeCompressed() {}
    bool isCompressed() const { return true; }
    S coeff(int i, int j) const { auto it = colv[j].find(i); return it == colv[j].end() ? S(0) : it->second; }
    S& coeffRef(int i, int j) { assert(i >= 0 && i < nr && j >= 0 && j < nc); return colv[j][i]; }
    const std::map<int, S>& column(int j) const { return colv[j]; }

    template <class It> void setFromTriplets(It b, It e) {   // duplicates are summed, in the order given
        setZero();
        for (It it = b; it != e; ++it) {
            assert(it->row() >= 0 && it->row() < nr && it->col() >= 0 && it->col() < nc);
            auto f = colv[it->col()].find(it->row());
            if (f == colv[it->col()].end()) colv[it->col()][it->row()] = it->value();
            else f->second += it->value();
        }
    }
    SparseMatrix transpose() const {
        SparseMatrix t(nc, nr);
        for (int j = 0; j < nc; j++)
            for (auto& e : colv[j]) t.colv[e.first][j] = e.second;
        return t;
    }
    SparseMatrix topLeftCorner(int r, int c) const {
        SparseMatrix t(r, c);
        for (int j = 0; j < c; j++)
            for (auto& e : colv[j])
                if (e.first < r) t.colv[j][e.first] = e.second;
        return t;
    }
    SparseMatrix& operator+=(const SparseMatrix& o) {
        assert(o.nr == nr && o.nc == nc);
        for (int j = 0; j < nc; j++)
            for (auto& e : o.colv[j]) {
                auto f = colv[j].find(e.first);
                if (f == colv[j].end()) colv[j][e.first] = e.second;
                else f->second += e.second;
            }
        return *this;
    }
    SparseMatrix& operator*=(S s) {
        for (auto& c : colv)
            for (auto& e : c) e.second *= s;
        return *this;
    }
    class InnerIterator {
        typename std::map<int, S>::const_iterator it, end;
        int j;

       public:
        InnerIterator(const SparseMatrix& m, int outer) : it(m.colv[outer].begin()), end(m.colv[outer].end()), j(outer) {}
        operator bool() const { return it != end; }
        InnerIterator& operator++() { ++it; return *this; }
        int row() const { return it->first; }
        int col() const { return j; }
        int index() const { return it->first; }
        S value() const { return it->second; }
    };
};

template <typename S>
SparseMatrix<S> operator*(const SparseMatrix<S>& a, const SparseMatrix<S>& b) {   // column by column, k ascending within a column
    assert(a.cols() == b.rows());
    SparseMatrix<S> m(a.rows(), b.cols());
    for (int j = 0; j < b.cols(); j++)
        for (auto& bk : b.column(j))
            for (auto& ai : a.column(bk.first)) {
                S& dst = m.coeffRef(ai.first, j);   // value-initialised to 0 on first touch
                dst += ai.second * bk.second;
            }
    return m;
}
template <typename S>
SparseMatrix<S> operator*(const SparseMatrix<S>& a, S s) {
    SparseMatrix<S> m = a;
    m *= s;
    return m;
}
template <typename S>
SparseMatrix<S> operator*(S s, const SparseMatrix<S>& a) {
    SparseMatrix<S> m = a;
    m *= s;
    return m;
}
template <typename S, class E>
Matrix<S, Dynamic, traits<E>::Cols> operator*(const SparseMatrix<S>& a, const DenseBase<E>& b_) {
    auto b = b_.eval();
    assert(a.cols() == b.rows());
    Matrix<S, Dynamic, traits<E>::Cols> m;
    m.resize(a.rows(), b.cols());
    for (int c = 0; c < b.cols(); c++)
        for (int j = 0; j < a.cols(); j++)
            for (auto& e : a.column(j)) m.coeffRef(e.first, c) += e.second * b.coeff(j, c);
    return m;
}
template <class D>
D& DenseBase<D>::operator+=(const SparseMatrix<Scalar>& s) {
    assert(s.rows() == derived().rows() && s.cols() == derived().cols());
    for (int j = 0; j < s.cols(); j++)
        for (auto& e : s.column(j)) derived().coeffRef(e.first, j) += e.second;
    return derived();
}

// permutations: P * A moves row i to row idx[i]; A * P^T moves column j to column idx[j]; P^-1 * v gathers v[idx[i]]
template <typename S, int N>
SparseMatrix<S> operator*(const PermutationMatrix<N>& p, const SparseMatrix<S>& a) {
    SparseMatrix<S> m(a.rows(), a.cols());
    for (int j = 0; j < a.cols(); j++)
        for (auto& e : a.column(j)) m.coeffRef(p.indices()[e.first], j) = e.second;
    return m;
}
template <typename S, int N>
SparseMatrix<S> operator*(const SparseMatrix<S>& a, const InversePermutation<PermutationMatrix<N>>& pt) {
    SparseMatrix<S> m(a.rows(), a.cols());
    for (int j = 0; j < a.cols(); j++)
        for (auto& e : a.column(j)) m.coeffRef(e.first, pt.p.indices()[j]) = e.second;
    return m;
}
template <int N, class E>
Matrix<typename traits<E>::Scalar, traits<E>::Rows, traits<E>::Cols> operator*(const PermutationMatrix<N>& p, const DenseBase<E>& v_) {
    auto v = v_.eval();
    auto m = v;
    for (int j = 0; j < v.cols(); j++)
        for (int i = 0; i < v.rows(); i++) m.coeffRef(p.indices()[i], j) = v.coeff(i, j);
    return m;
}
template <int N, class E>
Matrix<typename traits<E>::Scalar, traits<E>::Rows, traits<E>::Cols> operator*(const InversePermutation<PermutationMatrix<N>>& pt,
                                                                               const DenseBase<E>& v_) {
    auto v = v_.eval();
    auto m = v;
    for (int j = 0; j < v.cols(); j++)
        for (int i = 0; i < v.rows(); i++) m.coeffRef(i, j) = v.coeff(pt.p.indices()[i], j);
    return m;
}

template <typename I = int> class COLAMDOrdering {};
template <typename I = int> class AMDOrdering {};
template <typename I = int> class NaturalOrdering {};

// Sparse LDL^T without pivoting, analysed once per pattern: elimination tree + column counts, then the up-looking numeric phase
// (one sparse triangular solve per row).  Natural order (SimplicialLDLT's AMD ordering changes fill and rounding, not the result).
template <class SpMat, int UpLo = 1, class Ordering = AMDOrdering<int>>
class SimplicialLDLT {
    typedef typename SpMat::Scalar S;
    int n = 0;
    std::vector<int> parent, Lp, Li;
    std::vector<S> Lx, D;
    ComputationInfo status = InvalidInput;
    bool analysed = false;

   public:
    SimplicialLDLT() {}
    void analyzePattern(const SpMat& A) {
        n = A.rows();
        parent.assign(n, -1);
        std::vector<int> flag(n), lnz(n, 0);
        for (int k = 0; k < n; k++) {
            flag[k] = k;
            for (typename SpMat::InnerIterator it(A, k); it; ++it) {
                int i = it.row();
                if (i >= k) break;   // upper triangle of column k (rows ascending)
                for (; flag[i] != k; i = parent[i]) {
                    if (parent[i] == -1) parent[i] = k;
                    lnz[i]++;
                    flag[i] = k;
                }
            }
        }
        Lp.assign(n + 1, 0);
        for (int k = 0; k < n; k++) Lp[k + 1] = Lp[k] + lnz[k];
        Li.assign(Lp[n], 0);
        Lx.assign(Lp[n], S(0));
        D.assign(n, S(0));
        analysed = true;
    }
    void factorize(const SpMat& A) {
        assert(analysed && A.rows() == n);
        std::vector<S> y(n, S(0));
        std::vector<int> pattern(n), flag(n), lnz(n, 0);
        status = Success;
        for (int k = 0; k < n; k++) {
            int top = n;
            flag[k] = k;
            y[k] = 0;
            for (typename SpMat::InnerIterator it(A, k); it; ++it) {
                int i = it.row();
                if (i > k) break;
                y[i] += it.value();
                int len = 0;
                for (; flag[i] != k; i = parent[i]) {
                    pattern[len++] = i;
                    flag[i] = k;
                }
                while (len > 0) pattern[--top] = pattern[--len];
            }
            D[k] = y[k];
            y[k] = 0;
            for (; top < n; top++) {
                const int i = pattern[top];
                const S yi = y[i];
                y[i] = 0;
                const int p2 = Lp[i] + lnz[i];
                for (int p = Lp[i]; p < p2; p++) y[Li[p]] -= Lx[p] * yi;
                const S l_ki = yi / D[i];
                D[k] -= l_ki * yi;
                Li[p2] = k;
                Lx[p2] = l_ki;
                lnz[i]++;
            }
            if (D[k] == S(0) || D[k] != D[k]) {
                status = NumericalIssue;
                return;
            }
        }
    }
    void compute(const SpMat& A) { analyzePattern(A); factorize(A); }
    ComputationInfo info() const { return status; }
    template <class E>
    Matrix<S, Dynamic, 1> solve(const DenseBase<E>& b) const {
        Matrix<S, Dynamic, 1> x = b.eval();
        assert(x.size() == n);
        for (int j = 0; j < n; j++)
            for (int p = Lp[j]; p < Lp[j + 1]; p++) x(Li[p]) -= Lx[p] * x(j);
        for (int j = 0; j < n; j++) x(j) /= D[j];
        for (int j = n - 1; j >= 0; j--)
            for (int p = Lp[j]; p < Lp[j + 1]; p++) x(j) -= Lx[p] * x(Li[p]);
        return x;
    }
};
template <class SpMat, class Ordering = COLAMDOrdering<int>> class SparseLU;   // declared for slam/solver.hpp:73-76 (not selected)
template <class SpMat, class Ordering = COLAMDOrdering<int>> class SparseQR;

}  // namespace Eigen
