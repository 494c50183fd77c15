// linalg.hpp -- the small part of Eigen's surface the reference's callers use, as plain value types.
//
// The reference is written against Eigen3 (framework/definitions.hpp:5-7), which this image does not have.  These
// types keep the member names and semantics the reference's public API exposes (Isometry2f::translation / rotation /
// linear / inverse / operator*, Vector3f::x y z, Rotation2D::smallestAngle / matrix, VectorXf::segment, a sparse row
// matrix for the Jacobian out-parameters) so that code written for the reference compiles unchanged against this host.
// Arithmetic is float, like the reference.  Define BOS_HOST_NO_EIGEN_NAMESPACE to keep these out of `namespace Eigen`.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <vector>

namespace proj02 {
namespace la {

struct Vec2f {
    float v[2] = {0.f, 0.f};
    Vec2f() {}
    Vec2f(float x, float y) { v[0] = x; v[1] = y; }
    float& x() { return v[0]; }
    float& y() { return v[1]; }
    float x() const { return v[0]; }
    float y() const { return v[1]; }
    float& operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
    float& operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    Vec2f operator+(const Vec2f& o) const { return Vec2f(v[0] + o.v[0], v[1] + o.v[1]); }
    Vec2f operator-(const Vec2f& o) const { return Vec2f(v[0] - o.v[0], v[1] - o.v[1]); }
    Vec2f operator-() const { return Vec2f(-v[0], -v[1]); }
    Vec2f operator*(float s) const { return Vec2f(v[0] * s, v[1] * s); }
    Vec2f& operator+=(const Vec2f& o) { v[0] += o.v[0]; v[1] += o.v[1]; return *this; }
    float norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1]); }
    const Vec2f& transpose() const { return *this; }
};
struct Vec3f {
    float v[3] = {0.f, 0.f, 0.f};
    Vec3f() {}
    Vec3f(float x, float y, float z) { v[0] = x; v[1] = y; v[2] = z; }
    float& x() { return v[0]; }
    float& y() { return v[1]; }
    float& z() { return v[2]; }
    float x() const { return v[0]; }
    float y() const { return v[1]; }
    float z() const { return v[2]; }
    float& operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
    float& operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    Vec3f operator+(const Vec3f& o) const { return Vec3f(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    Vec3f operator-(const Vec3f& o) const { return Vec3f(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vec3f operator*(float s) const { return Vec3f(v[0] * s, v[1] * s, v[2] * s); }
    template <int N> Vec2f head() const { static_assert(N == 2, "head<2>"); return Vec2f(v[0], v[1]); }
    const Vec3f& transpose() const { return *this; }
};
inline std::ostream& operator<<(std::ostream& o, const Vec2f& a) { return o << a.v[0] << " " << a.v[1]; }
inline std::ostream& operator<<(std::ostream& o, const Vec3f& a) { return o << a.v[0] << " " << a.v[1] << " " << a.v[2]; }

struct Mat2f {
    float m[2][2] = {{1.f, 0.f}, {0.f, 1.f}};
    float& operator()(int i, int j) { return m[i][j]; }
    float operator()(int i, int j) const { return m[i][j]; }
    Mat2f transpose() const { Mat2f t; t.m[0][0] = m[0][0]; t.m[0][1] = m[1][0]; t.m[1][0] = m[0][1]; t.m[1][1] = m[1][1]; return t; }
    Vec2f operator*(const Vec2f& a) const { return Vec2f(m[0][0] * a.v[0] + m[0][1] * a.v[1], m[1][0] * a.v[0] + m[1][1] * a.v[1]); }
    Mat2f operator*(const Mat2f& o) const {
        Mat2f r;
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++) r.m[i][j] = m[i][0] * o.m[0][j] + m[i][1] * o.m[1][j];
        return r;
    }
};
struct Mat3f {
    float m[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    float& operator()(int i, int j) { return m[i][j]; }
    float operator()(int i, int j) const { return m[i][j]; }
    static Mat3f Zero() { return Mat3f(); }
    static Mat3f Identity() { Mat3f r; r.m[0][0] = r.m[1][1] = r.m[2][2] = 1.f; return r; }
    Vec3f operator*(const Vec3f& a) const {
        return Vec3f(m[0][0] * a.v[0] + m[0][1] * a.v[1] + m[0][2] * a.v[2], m[1][0] * a.v[0] + m[1][1] * a.v[1] + m[1][2] * a.v[2],
                     m[2][0] * a.v[0] + m[2][1] * a.v[1] + m[2][2] * a.v[2]);
    }
};

// Eigen::Rotation2D<float>: stores the angle un-normalised (framework/observation.hpp:16-27 relies on that)
class Rotation2f {
 public:
    Rotation2f() : a_(0.f) {}
    Rotation2f(float angle) : a_(angle) {}
    explicit Rotation2f(const Mat2f& m) : a_(std::atan2(m(1, 0), m(0, 0))) {}
    float angle() const { return a_; }
    float& angle() { return a_; }
    // fmod(a, 2pi) folded into [-pi, pi], all in float (Eigen >= 3.3)
    float smallestAngle() const {
        const float pi = 3.14159265358979323846f, two_pi = 6.28318530717958647692f;
        float t = std::fmod(a_, two_pi);
        if (t > pi) t -= two_pi;
        else if (t < -pi) t += two_pi;
        return t;
    }
    Mat2f matrix() const {
        const float s = std::sin(a_), c = std::cos(a_);
        Mat2f r; r.m[0][0] = c; r.m[0][1] = -s; r.m[1][0] = s; r.m[1][1] = c;
        return r;
    }
    Mat2f toRotationMatrix() const { return matrix(); }
 private:
    float a_;
};

// Eigen::Isometry2f: linear part + translation, never re-orthonormalised
class Iso2f {
 public:
    Iso2f() {}
    void setIdentity() { R_ = Mat2f(); t_ = Vec2f(); }
    static Iso2f Identity() { return Iso2f(); }
    Vec2f& translation() { return t_; }
    const Vec2f& translation() const { return t_; }
    Mat2f& linear() { return R_; }
    const Mat2f& linear() const { return R_; }
    Mat2f rotation() const { return R_; }   // == linear() for Isometry mode
    Iso2f inverse() const {                  // linear = R^T, translation = -(R^T t)
        Iso2f r;
        r.R_ = R_.transpose();
        r.t_ = -(r.R_ * t_);
        return r;
    }
    Vec2f operator*(const Vec2f& p) const { return R_ * p + t_; }
    Iso2f operator*(const Iso2f& o) const { Iso2f r; r.R_ = R_ * o.R_; r.t_ = R_ * o.t_ + t_; return r; }
 private:
    Mat2f R_;
    Vec2f t_;
};

// Eigen::VectorXf, as far as State::apply_boxplus and the callers need it
class VectorXf {
 public:
    VectorXf() {}
    explicit VectorXf(std::size_t n) : d_(n, 0.f) {}
    void resize(std::size_t n) { d_.assign(n, 0.f); }
    void setZero() { std::fill(d_.begin(), d_.end(), 0.f); }
    std::size_t size() const { return d_.size(); }
    float& operator()(std::size_t i) { return d_[i]; }
    float operator()(std::size_t i) const { return d_[i]; }
    float& operator[](std::size_t i) { return d_[i]; }
    float operator[](std::size_t i) const { return d_[i]; }
    template <int N> typename std::conditional<N == 3, Vec3f, Vec2f>::type segment(std::size_t i) const;
    float* data() { return d_.data(); }
    const float* data() const { return d_.data(); }
    const VectorXf& transpose() const { return *this; }
 private:
    std::vector<float> d_;
};
template <> inline Vec3f VectorXf::segment<3>(std::size_t i) const { return Vec3f(d_[i], d_[i + 1], d_[i + 2]); }
template <> inline Vec2f VectorXf::segment<2>(std::size_t i) const { return Vec2f(d_[i], d_[i + 1]); }
inline std::ostream& operator<<(std::ostream& o, const VectorXf& a) {
    for (std::size_t i = 0; i < a.size(); i++) o << (i ? " " : "") << a(i);
    return o;
}

struct Triplet_f {
    int r, c;
    float v;
    Triplet_f(int r_ = 0, int c_ = 0, float v_ = 0.f) : r(r_), c(c_), v(v_) {}
    int row() const { return r; }
    int col() const { return c; }
    float value() const { return v; }
};

// Eigen::SparseMatrix<float> as used for the Jacobian out-parameters of Solver::error_and_jacobian: a few explicit
// entries (structural zeros kept, like setFromTriplets) of a rows x cols matrix.
class SparseMatrixXf {
 public:
    SparseMatrixXf() : rows_(0), cols_(0) {}
    SparseMatrixXf(int r, int c) : rows_(r), cols_(c) {}
    void resize(int r, int c) { rows_ = r; cols_ = c; e_.clear(); }
    int rows() const { return rows_; }
    int cols() const { return cols_; }
    void setZero() { e_.clear(); }
    int nonZeros() const { return (int)e_.size(); }
    void makeCompressed() {}
    template <class It> void setFromTriplets(It b, It e) {
        e_.clear();
        for (; b != e; ++b) coeffRef(b->row(), b->col()) += b->value();
    }
    float coeff(int r, int c) const {
        for (const Triplet_f& t : e_)
            if (t.r == r && t.c == c) return t.v;
        return 0.f;
    }
    float& coeffRef(int r, int c) {
        for (Triplet_f& t : e_)
            if (t.r == r && t.c == c) return t.v;
        e_.emplace_back(r, c, 0.f);
        return e_.back().v;
    }
    float& insert(int r, int c) { return coeffRef(r, c); }
    SparseMatrixXf operator-(const SparseMatrixXf& o) const {
        SparseMatrixXf d(*this);
        for (const Triplet_f& t : o.e_) d.coeffRef(t.r, t.c) -= t.v;
        return d;
    }
    SparseMatrixXf cwiseAbs() const { SparseMatrixXf d(*this); for (Triplet_f& t : d.e_) t.v = std::fabs(t.v); return d; }
    float sum() const { float s = 0.f; for (const Triplet_f& t : e_) s += t.v; return s; }
    float maxCoeff() const { float m = e_.empty() ? 0.f : e_[0].v; for (const Triplet_f& t : e_) m = std::max(m, t.v); return m; }
    const SparseMatrixXf& coeffs() const { return *this; }
    const std::vector<Triplet_f>& entries() const { return e_; }
 private:
    int rows_, cols_;
    std::vector<Triplet_f> e_;
};

}  // namespace la
}  // namespace proj02

#ifndef BOS_HOST_NO_EIGEN_NAMESPACE
// spelling compatibility for callers that say Eigen::Matrix3f / Eigen::VectorXf, as the reference's executables do
namespace Eigen {
typedef proj02::la::Vec2f Vector2f;
typedef proj02::la::Vec3f Vector3f;
typedef proj02::la::Mat2f Matrix2f;
typedef proj02::la::Mat3f Matrix3f;
typedef proj02::la::Iso2f Isometry2f;
typedef proj02::la::VectorXf VectorXf;
}  // namespace Eigen
#endif
