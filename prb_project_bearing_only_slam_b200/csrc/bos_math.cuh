// bos_math.cuh -- device-side SE(2) / bearing arithmetic shared by every kernel.
// Formulas follow the reference (file:line cited per function); the arrangement is ours.
#pragma once

#include <cuda_runtime.h>

namespace bos {

#define BOS_FULL_MASK 0xffffffffu

template <typename S> struct Cst;
template <> struct Cst<double> {
    // Eigen's EIGEN_PI narrowed to Scalar (Rotation2D::smallestAngle)
    static __host__ __device__ constexpr double pi() { return 3.141592653589793238462643383279502884; }
    static __host__ __device__ constexpr double two_pi() { return 2.0 * 3.141592653589793238462643383279502884; }
};
template <> struct Cst<float> {
    static __host__ __device__ constexpr float pi() { return 3.14159265358979323846f; }
    static __host__ __device__ constexpr float two_pi() { return 6.28318530717958647692f; }
};

// OpenCV's CV_PI / CV_2PI are doubles (slam/solver_jacobians.cpp:325-333)
__device__ __forceinline__ constexpr double cv_pi() { return 3.1415926535897932384626433832795; }
__device__ __forceinline__ constexpr double cv_2pi() { return 6.283185307179586476925286766559; }

// Eigen::Rotation2D::smallestAngle: fmod(a, 2pi) folded into [-pi, pi].  |a| <= pi is a fixed point
// of both steps, so the common case skips the fmod.
template <typename S>
__device__ __forceinline__ S smallest_angle(S a) {
    if (fabs(a) <= Cst<S>::pi()) return a;
    S t = fmod(a, Cst<S>::two_pi());
    if (t > Cst<S>::pi()) t -= Cst<S>::two_pi();
    else if (t < -Cst<S>::pi()) t += Cst<S>::two_pi();
    return t;
}

// Solver::normalized_angle (slam/solver_jacobians.cpp:325-333): compare/add in double, narrow to S
// after every addition; range [-pi, pi).
template <typename S>
__device__ __forceinline__ S normalized_angle(S a) {
    while ((double)a < -cv_pi()) a = (S)((double)a + cv_2pi());
    while ((double)a >= cv_pi()) a = (S)((double)a - cv_2pi());
    return a;
}

// ---- lean FP64 reciprocal and atan2 for the edge kernels ------------------------------------------------------------------------
// The H, b build is instruction-issue bound, and a third of its instructions were libdevice's double atan2 (~140 executed instructions,
// 39 of them UMOVs that materialise polynomial constants) and its guarded division.  These versions keep full double accuracy
// (max |atan2_b - atan2| = 4.4e-16 = one ulp of pi over 2 M random arguments spanning 1e-3 .. 1e3; tests/test_gpu_parity.py checks the
// kernels against glibc through the oracle at 1e-12) with ~40 instructions: MUFU.RCP64H seed + two Newton steps, octant reduction to
// |t| <= tan(pi/8) BEFORE the one reciprocal, and an 11-term polynomial whose coefficients come from the constant bank as DFMA operands.
__device__ __forceinline__ double rcp_b(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));      // >= 20 good bits
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ float rcp_b(float x) { return 1.0f / x; }

// (atan(t) / t - 1) / t^2 on t^2 in [0, tan(pi/8)^2], highest degree first (Chebyshev fit, error 3e-18 before rounding)
static __device__ __constant__ double kAtanPoly[11] = {
    -1.91754047111049285e-02, 3.92304477847930932e-02, -5.08540783458948445e-02, 5.85814090486632827e-02, -6.66451052550893625e-02,
    7.69218312537689186e-02,  -9.09090457530648960e-02, 1.11111110151870182e-01, -1.42857142846656821e-01, 1.99999999999955158e-01,
    -3.33333333333333315e-01};

__device__ __forceinline__ double atan2_b(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    const double mx = fmax(ax, ay), mn = fmin(ax, ay);
    // atan(mn / mx) = pi/4 + atan((mn - mx) / (mn + mx)) when mn / mx > tan(pi/8): either way |t| <= tan(pi/8)
    const bool big = mn > 0.41421356237309503 * mx;
    const double num = big ? mn - mx : mn, den = big ? mn + mx : mx;
    const double rc = rcp_b(den);
    double t = num * rc;
    t = fma(fma(-t, den, num), rc, t);                          // one correction step: t = num / den to the last bit or two
    const double u = t * t;
    double p = kAtanPoly[0];
#pragma unroll
    for (int k = 1; k < 11; k++) p = fma(p, u, kAtanPoly[k]);
    double r = fma(t * u, p, t);
    if (big) r += 0.78539816339744831;
    if (ay > ax) r = 1.5707963267948966 - r;
    if (x < 0.0) r = 3.141592653589793 - r;
    return copysign(r, y);
}
__device__ __forceinline__ float atan2_b(float y, float x) { return atan2f(y, x); }

template <typename S> struct Vec2T;
template <> struct Vec2T<double> { typedef double2 type; };
template <> struct Vec2T<float> { typedef float2 type; };

template <typename S>
struct PoseV { S x, y, c, s; };

template <typename S>
__device__ __forceinline__ PoseV<S> load_pose(const S* __restrict__ pose, int i);
template <>
__device__ __forceinline__ PoseV<double> load_pose<double>(const double* __restrict__ pose, int i) {
    const double2* p = reinterpret_cast<const double2*>(pose) + 2 * (size_t)i;
    double2 a = __ldg(p), b = __ldg(p + 1);
    return PoseV<double>{a.x, a.y, b.x, b.y};
}
template <>
__device__ __forceinline__ PoseV<float> load_pose<float>(const float* __restrict__ pose, int i) {
    float4 a = __ldg(reinterpret_cast<const float4*>(pose) + i);
    return PoseV<float>{a.x, a.y, a.z, a.w};
}
template <typename S>
__device__ __forceinline__ void load_lm(const S* __restrict__ lm, int j, S& lx, S& ly) {
    typedef typename Vec2T<S>::type V2;
    V2 v = __ldg(reinterpret_cast<const V2*>(lm) + j);
    lx = v.x; ly = v.y;
}

// Landmark half of the bearing Jacobian, J_lm = a R^T with a = [-gy, gx] / |g|^2 (slam/solver_jacobians.cpp:32-49, 85-89).
// The pose half follows from it: J_pose = (-J_lm[0], -J_lm[1], J_lm . (ly, -lx)).  Shared by the assembly path's
// bearing_terms and by the PCG operator, which re-derives its per-edge factors from the state.
template <typename S>
__device__ __forceinline__ void bearing_jl(const PoseV<S>& X, S lx, S ly, S& j0, S& j1) {
    const S c = X.c, s = X.s;
    const S itx = (-c) * X.x + (-s) * X.y;
    const S ity = s * X.x + (-c) * X.y;
    const S gx = (c * lx + s * ly) + itx;
    const S gy = ((-s) * lx + c * ly) + ity;
    const S f = S(1) / (gx * gx + gy * gy);
    const S a0 = f * (-gy), a1 = f * gx;
    j0 = a0 * c + a1 * (-s);
    j1 = a0 * s + a1 * c;
}

// The same landmark Jacobian without the rotation: R g = l - t =: d and rotations commute with the 90-degree turn, so
// J_lm = [-gy, gx] R^T / |g|^2 = (-dy, dx) / |d|^2.  Algebraically identical to bearing_jl (rounding differs at the 1e-16 level);
// used by the PCG operator, which only needs the pose translation this way.
template <typename S>
__device__ __forceinline__ void bearing_jl_world(S px, S py, S lx, S ly, S& j0, S& j1) {
    const S dx = lx - px, dy = ly - py;
    const S f = rcp_b(dx * dx + dy * dy);
    j0 = -(dy * f);
    j1 = dx * f;
}

// Bearing error and 1x5 Jacobian [J_pose(3) | J_lm(2)]  (slam/solver_jacobians.cpp:9-95, 301-305).
// g = X^-1 * l with Eigen's isometry inverse: R^T l + (-(R^T) t); J = a * [-R^T | R^T (ly,-lx)^T | R^T],
// a = [-gy, gx] / |g|^2.  The world-frame landmark appears in the theta column because boxplus is a
// LEFT perturbation (framework/state.hpp:11-13).
template <typename S>
__device__ __forceinline__ void bearing_terms(const PoseV<S>& X, S lx, S ly, S z, S& err, S J[5]) {
    const S c = X.c, s = X.s;
    S itx = (-c) * X.x + (-s) * X.y;
    S ity = s * X.x + (-c) * X.y;
    S gx = (c * lx + s * ly) + itx;
    S gy = ((-s) * lx + c * ly) + ity;
    S pred = atan2_b(gy, gx);
    err = normalized_angle<S>(pred - smallest_angle<S>(z));
    S f = rcp_b(gx * gx + gy * gy);
    S a0 = f * (-gy), a1 = f * gx;
    S v0 = (-s) * lx + c * ly;
    S v1 = (-c) * lx + (-s) * ly;
    J[0] = a0 * (-c) + a1 * s;
    J[1] = a0 * (-s) + a1 * (-c);
    J[2] = a0 * v0 + a1 * v1;
    J[3] = a0 * c + a1 * (-s);
    J[4] = a0 * s + a1 * c;
}

// Odometry error (3) and the SOURCE Jacobian block rows (slam/solver_jacobians.cpp:97-168, 307-323).
// J_src = [[-R_s^T, u], [0 0 -1]] with u = (DR' R_s)^T t_d;  J_dst = -J_src entry for entry
// (R_s^T DR' t_d = -u), so only A = -R_s^T and u are returned:
//   J_src rows: (-c, -s, u0), (s, -c, u1), (0, 0, -1).
// t2v's angle of a pose (framework/definitions.hpp:39-43): Rotation2D(R).smallestAngle() = smallestAngle(atan2(R10, R00))
template <typename S>
__device__ __forceinline__ S pose_theta(const PoseV<S>& X) { return smallest_angle<S>(atan2(X.s, X.c)); }

// ths / thd = pose_theta of the two poses (the assembly path caches them per pose, K7 refreshes them with the state)
template <typename S>
__device__ __forceinline__ void odometry_terms(const PoseV<S>& Xs, const PoseV<S>& Xd, S ths, S thd, S z0, S z1, S z2,
                                               S err[3], S& u0, S& u1) {
    S tx = Xd.x - Xs.x, ty = Xd.y - Xs.y;
    S p0 = Xs.c * tx + Xs.s * ty;
    S p1 = (-Xs.s) * tx + Xs.c * ty;
    S p2 = normalized_angle<S>(thd - ths);
    err[0] = p0 - z0;
    err[1] = p1 - z1;
    err[2] = normalized_angle<S>(p2 - z2);
    u0 = (-Xs.s) * Xd.x + Xs.c * Xd.y;
    u1 = (-Xs.c) * Xd.x + (-Xs.s) * Xd.y;
}
template <typename S>
__device__ __forceinline__ void odometry_terms(const PoseV<S>& Xs, const PoseV<S>& Xd, S z0, S z1, S z2, S err[3], S& u0, S& u1) {
    odometry_terms<S>(Xs, Xd, pose_theta<S>(Xs), pose_theta<S>(Xd), z0, z1, z2, err, u0, u1);
}

// M = J_s^T Omega J_s (symmetric, 6 unique: 00 01 02 11 12 22), v = J_s^T Omega e, chi = e^T Omega e.
// Omega symmetric, upper triangle om = (00 01 02 11 12 22).
template <typename S>
__device__ __forceinline__ void odometry_normal_terms(S c, S s, S u0, S u1, const S om[6], const S e[3],
                                                      S M[6], S v[3], S scale) {
    // rows of J_s
    const S j00 = -c, j01 = -s, j02 = u0;
    const S j10 = s, j11 = -c, j12 = u1;
    const S j22 = S(-1);
    // T = Omega * J_s (3x3): T[k][a] = sum_m Om[k][m] J[m][a]
    const S o00 = om[0], o01 = om[1], o02 = om[2], o11 = om[3], o12 = om[4], o22 = om[5];
    S t00 = o00 * j00 + o01 * j10, t01 = o00 * j01 + o01 * j11, t02 = o00 * j02 + o01 * j12 + o02 * j22;
    S t10 = o01 * j00 + o11 * j10, t11 = o01 * j01 + o11 * j11, t12 = o01 * j02 + o11 * j12 + o12 * j22;
    S t20 = o02 * j00 + o12 * j10, t21 = o02 * j01 + o12 * j11, t22 = o02 * j02 + o12 * j12 + o22 * j22;
    M[0] = j00 * t00 + j10 * t10;                 // (0,0)
    M[1] = j00 * t01 + j10 * t11;                 // (0,1)
    M[2] = j00 * t02 + j10 * t12;                 // (0,2)
    M[3] = j01 * t01 + j11 * t11;                 // (1,1)
    M[4] = j01 * t02 + j11 * t12;                 // (1,2)
    M[5] = j02 * t02 + j12 * t12 + j22 * t22;     // (2,2)
    // Omega * e (scaled error)
    S e0 = e[0] * scale, e1 = e[1] * scale, e2 = e[2] * scale;
    S w0 = o00 * e0 + o01 * e1 + o02 * e2;
    S w1 = o01 * e0 + o11 * e1 + o12 * e2;
    S w2 = o02 * e0 + o12 * e1 + o22 * e2;
    v[0] = j00 * w0 + j10 * w1;
    v[1] = j01 * w0 + j11 * w1;
    v[2] = j02 * w0 + j12 * w1 + j22 * w2;
}

template <typename S>
__device__ __forceinline__ S odometry_chi(const S om[6], const S e[3]) {
    S w0 = om[0] * e[0] + om[1] * e[1] + om[2] * e[2];
    S w1 = om[1] * e[0] + om[3] * e[1] + om[4] * e[2];
    S w2 = om[2] * e[0] + om[4] * e[1] + om[5] * e[2];
    return w0 * e[0] + w1 * e[1] + w2 * e[2];
}

// fire-and-forget global add (RED); result unused so ptxas emits RED, not ATOM
template <typename S>
__device__ __forceinline__ void red_add(S* addr, S v) { atomicAdd(addr, v); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BOS_FULL_MASK, v, o);
    return v;
}

// symmetric 3x3 (00 01 02 11 12 22) inverse, symmetric output
template <typename S>
__device__ __forceinline__ void sym3_inverse(const S a[6], S o[6]) {
    S c00 = a[3] * a[5] - a[4] * a[4];
    S c01 = a[4] * a[2] - a[1] * a[5];
    S c02 = a[1] * a[4] - a[3] * a[2];
    S det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    S id = S(1) / det;
    o[0] = c00 * id; o[1] = c01 * id; o[2] = c02 * id;
    o[3] = (a[0] * a[5] - a[2] * a[2]) * id;
    o[4] = (a[1] * a[2] - a[0] * a[4]) * id;
    o[5] = (a[0] * a[3] - a[1] * a[1]) * id;
}

}  // namespace bos
