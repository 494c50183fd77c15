// bos_oracle.hpp -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A scalar-templated CPU restatement of the Gauss-Newton iteration of
// torchipeppo/prb-project-bearing-only-slam.  It is the checker for the CUDA path,
// never the thing shipped or measured: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference leg may build, load or call it.
//
//   Oracle<float>   reference-faithful precision (the reference is FP32 end to end,
//                   framework/definitions.hpp:17-37) and evaluation order.
//   Oracle<double>  the FP64 parity target of the CUDA kernels (same formulas).
//
// PARITY PIN STATUS: pinned by the reference's own code, run here.  The reference needs Eigen3 and OpenCV C++ (absent from this
// image, no network), so it cannot be built against them; `make -C oracle ref` compiles its UNMODIFIED hot-path sources where they lie
// under /root/reference against a stand-in for the Eigen / OpenCV calls they make (oracle/eigen_standin, oracle/ref_driver.cpp ->
// oracle/_ref/libbos_ref.so; standin.hpp lists what the stand-in supplies -- the arithmetic behind the operators, an LDL^T, a
// column-pivoting QR -- and what therefore remains unpinned: Eigen's own rounding and its QR rank decisions).  tests/test_ref_build.py:
// from the same state Oracle<float> and that build agree BIT FOR BIT on every per-edge error, Jacobian and on b, to 1e-6 on H, on the
// sparsity pattern exactly, and along the 20 / 50-iteration trajectories of the bundled datasets and two random worlds;
// Oracle<double> agrees with it to float rounding.  Its results are committed as tests/golden/ref_*.npz
// (tests/golden/make_ref_golden.py).  The soft pins the reference itself offers stay in tests/test_oracle_pins.py: the predict_bearing
// known answers (tests/solver_stuff.cpp:25-38), the analytic-vs-numeric Jacobian statistics (:82-88, 156-162), predict_odometry ==
// measurement on the dead-reckoned initial guess (:93-114), the single-observation landmarks 69/112/114 (slam/triangulation.cpp:41),
// fixed pose 1498, the README's "~20 iterations" (README.md:22-24), and an independent numpy restatement.
//
// Third-party arithmetic restated (Eigen3 >= 3.3, unpinned; OpenCV constants):
//   Rotation2D::smallestAngle, Rotation2D(Matrix2), Rotation2D::matrix,
//   Isometry2::inverse / operator*, ColPivHouseholderQR::solve (M x 2), CV_PI/CV_2PI.
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "bos_sparse_ldlt.hpp"

namespace bos_oracle {

// OpenCV's CV_PI / CV_2PI are doubles (used by Solver::normalized_angle,
// slam/solver_jacobians.cpp:325-333).
static const double kCvPi = 3.1415926535897932384626433832795;
static const double kCv2Pi = 6.283185307179586476925286766559;
// Eigen's EIGEN_PI is a long double literal, narrowed to Scalar at each use.
static const long double kEigenPi = 3.141592653589793238462643383279502884197169399375105820974944592307816406L;

// ---- Eigen::Rotation2D<T>::smallestAngle(): fmod then fold into [-pi, pi] -------------
template <class T>
inline T smallest_angle(T a) {
    T t = std::fmod(a, T(2 * kEigenPi));
    if (t > T(kEigenPi)) t -= T(2 * kEigenPi);
    else if (t < -T(kEigenPi)) t += T(2 * kEigenPi);
    return t;
}

// ---- Solver::normalized_angle (slam/solver_jacobians.cpp:325-333) ---------------------
// The float angle is compared with, and incremented by, DOUBLE constants and narrowed
// back to T after every addition.  Range [-pi, pi).
template <class T>
inline T normalized_angle(T angle) {
    while ((double)angle < -kCvPi) angle = (T)((double)angle + kCv2Pi);
    while ((double)angle >= kCvPi) angle = (T)((double)angle - kCv2Pi);
    return angle;
}

// ---- Eigen::Isometry2<T>: 2x2 linear part + translation, never re-orthonormalised -----
template <class T>
struct Iso2 {
    T r00 = 1, r01 = 0, r10 = 0, r11 = 1, tx = 0, ty = 0;
};

// framework/definitions.hpp:45-53  (Rotation2D::matrix() = [[c,-s],[s,c]])
template <class T>
inline Iso2<T> v2t(T x, T y, T th) {
    Iso2<T> X;
    T s = std::sin(th), c = std::cos(th);
    X.r00 = c; X.r01 = -s; X.r10 = s; X.r11 = c;
    X.tx = x; X.ty = y;
    return X;
}

// framework/definitions.hpp:39-43  (Rotation2D(Matrix2) = atan2(m10, m00))
template <class T>
inline void t2v(const Iso2<T>& X, T& x, T& y, T& th) {
    x = X.tx; y = X.ty;
    th = smallest_angle<T>(std::atan2(X.r10, X.r00));
}

// Isometry * Isometry: R1 R2, R1 t2 + t1
template <class T>
inline Iso2<T> compose(const Iso2<T>& A, const Iso2<T>& B) {
    Iso2<T> C;
    C.r00 = A.r00 * B.r00 + A.r01 * B.r10;
    C.r01 = A.r00 * B.r01 + A.r01 * B.r11;
    C.r10 = A.r10 * B.r00 + A.r11 * B.r10;
    C.r11 = A.r10 * B.r01 + A.r11 * B.r11;
    C.tx = (A.r00 * B.tx + A.r01 * B.ty) + A.tx;
    C.ty = (A.r10 * B.tx + A.r11 * B.ty) + A.ty;
    return C;
}

// framework/state.hpp:11-13: boxplus = v2t(delta) * X  (LEFT perturbation)
template <class T>
inline Iso2<T> boxplus(const Iso2<T>& X, T dx, T dy, T dth) {
    return compose(v2t<T>(dx, dy, dth), X);
}

// Isometry inverse (mode Isometry): linear = R^T, translation = (-R^T) * t
template <class T>
inline Iso2<T> inverse(const Iso2<T>& X) {
    Iso2<T> I;
    I.r00 = X.r00; I.r01 = X.r10; I.r10 = X.r01; I.r11 = X.r11;
    I.tx = (-I.r00) * X.tx + (-I.r01) * X.ty;
    I.ty = (-I.r10) * X.tx + (-I.r11) * X.ty;
    return I;
}

template <class T>
inline void apply(const Iso2<T>& X, T lx, T ly, T& ox, T& oy) {
    ox = (X.r00 * lx + X.r01 * ly) + X.tx;
    oy = (X.r10 * lx + X.r11 * ly) + X.ty;
}

// framework/observation.hpp:12-81 (host AoS edge types; bearing stored un-normalised)
template <class T>
struct BearingObs { int pose_id, lm_id; T bearing; T omega; };
template <class T>
struct OdomObs { int src_id, dst_id; T z[3]; T omega[9]; };

// ---- Eigen::ColPivHouseholderQR<Matrix<T,Dynamic,2>>::solve restated -------------------
// A is M x 2 (row-major pairs), b has M entries.  Follows Eigen's computeInPlace /
// _solve_impl: pivot on the larger (running) column norm, Householder reflectors with
// beta = -sign(c0) * norm, solution uses nonzeroPivots() and leaves the non-pivot
// coordinates at zero (M = 1 => rank-1 basic solution).  slam/triangulation.cpp:59.
template <class T>
inline void colpiv_householder_solve_Mx2(std::vector<T> A, std::vector<T> b, int M, T out[2]) {
    const int rows = M, cols = 2;
    const int size = std::min(rows, cols);
    auto at = [&](int i, int j) -> T& { return A[(size_t)i * 2 + j]; };
    auto col_norm = [&](int j, int from) {
        T s = 0;
        for (int i = from; i < rows; i++) s += at(i, j) * at(i, j);
        return std::sqrt(s);
    };
    const T eps = std::numeric_limits<T>::epsilon();
    T norms_upd[2], norms_dir[2];
    for (int j = 0; j < cols; j++) norms_upd[j] = norms_dir[j] = col_norm(j, 0);
    T mx = std::max(norms_upd[0], norms_upd[1]);
    T th = mx * eps / T(rows);
    const T threshold_helper = th * th;
    const T norm_downdate_threshold = std::sqrt(eps);
    int nonzero_pivots = size;
    int perm[2] = {0, 1};
    T hcoef[2] = {0, 0};
    for (int k = 0; k < size; k++) {
        int big = k;
        for (int j = k + 1; j < cols; j++)
            if (norms_upd[j] > norms_upd[big]) big = j;
        T big_sq = norms_upd[big] * norms_upd[big];
        if (nonzero_pivots == size && big_sq < threshold_helper * T(rows - k)) nonzero_pivots = k;
        if (big != k) {
            for (int i = 0; i < rows; i++) std::swap(at(i, k), at(i, big));
            std::swap(norms_upd[k], norms_upd[big]);
            std::swap(norms_dir[k], norms_dir[big]);
            std::swap(perm[k], perm[big]);
        }
        // makeHouseholderInPlace on A[k:, k]
        T tail_sq = 0;
        for (int i = k + 1; i < rows; i++) tail_sq += at(i, k) * at(i, k);
        T c0 = at(k, k), beta, tau;
        if (tail_sq <= std::numeric_limits<T>::min()) {
            tau = 0; beta = c0;
            for (int i = k + 1; i < rows; i++) at(i, k) = 0;
        } else {
            beta = std::sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0) beta = -beta;
            for (int i = k + 1; i < rows; i++) at(i, k) = at(i, k) / (c0 - beta);
            tau = (beta - c0) / beta;
        }
        hcoef[k] = tau;
        at(k, k) = beta;
        // apply H_k on the left to the trailing columns and keep it for the rhs
        for (int j = k + 1; j < cols; j++) {
            if (rows - k == 1) { at(k, j) *= (T(1) - tau); }
            else if (tau != 0) {
                T tmp = 0;
                for (int i = k + 1; i < rows; i++) tmp += at(i, k) * at(i, j);
                tmp += at(k, j);
                at(k, j) -= tau * tmp;
                for (int i = k + 1; i < rows; i++) at(i, j) -= tau * at(i, k) * tmp;
            }
        }
        for (int j = k + 1; j < cols; j++) {
            if (norms_upd[j] != 0) {
                T temp = std::abs(at(k, j)) / norms_upd[j];
                temp = (T(1) + temp) * (T(1) - temp);
                temp = temp < 0 ? T(0) : temp;
                T r = norms_upd[j] / norms_dir[j];
                T temp2 = temp * (r * r);
                if (temp2 <= norm_downdate_threshold) {
                    norms_dir[j] = col_norm(j, k + 1);
                    norms_upd[j] = norms_dir[j];
                } else {
                    norms_upd[j] *= std::sqrt(temp);
                }
            }
        }
    }
    out[0] = out[1] = 0;
    if (nonzero_pivots == 0) return;
    // c = Q^T b : apply H_0 then H_1
    for (int k = 0; k < nonzero_pivots; k++) {
        T tau = hcoef[k];
        if (rows - k == 1) { b[k] *= (T(1) - tau); }
        else if (tau != 0) {
            T tmp = 0;
            for (int i = k + 1; i < rows; i++) tmp += at(i, k) * b[i];
            tmp += b[k];
            b[k] -= tau * tmp;
            for (int i = k + 1; i < rows; i++) b[i] -= tau * at(i, k) * tmp;
        }
    }
    // upper-triangular solve on the leading nonzero_pivots block
    T c[2] = {b[0], rows > 1 ? b[1] : T(0)};
    if (nonzero_pivots == 2) {
        c[1] = c[1] / at(1, 1);
        c[0] = (c[0] - at(0, 1) * c[1]) / at(0, 0);
    } else {
        c[0] = c[0] / at(0, 0);
    }
    for (int i = 0; i < nonzero_pivots; i++) out[perm[i]] = c[i];
}

// Scalar CSC matrix (both triangles), the form the reference hands to SimplicialLDLT.
template <class T>
struct Csc {
    int n = 0;
    std::vector<int> colptr, rowidx;
    std::vector<T> val;
};

struct IterStats {
    double chi2_bearing = 0, chi2_odometry = 0;
    int over_bearing = 0, over_odometry = 0;
    double delta_inf = 0;
    int solver_status = 0;  // 0 ok, 1 non-positive pivot seen
    int pcg_iterations = 0;
};

template <class T>
class Oracle {
public:
    // ---------------- State (framework/state.cpp:20-67) ---------------------------------
    std::vector<Iso2<T>> poses;
    std::vector<T> lms;  // x,y interleaved
    std::map<int, int> pose_id_to_stix, lm_id_to_stix;
    std::vector<int> pose_stix_to_id, lm_stix_to_id;
    std::vector<BearingObs<T>> bearings;
    std::vector<OdomObs<T>> odoms;
    int fixed_pose_id = -1;
    float bound = 0;
    int n_unrecognized = 0;
    std::vector<int> single_observation_lms;  // ids warned about by triangulate (triangulation.cpp:38-42)

    void add_pose(const Iso2<T>& X, int id) {
        poses.push_back(X);
        pose_id_to_stix[id] = (int)poses.size() - 1;  // duplicate id: map overwritten, vector grows
        pose_stix_to_id.push_back(id);
    }
    void add_pose(T x, T y, T th, int id) { add_pose(v2t<T>(x, y, th), id); }
    void add_landmark(T x, T y, int id) {
        lms.push_back(x); lms.push_back(y);
        lm_id_to_stix[id] = (int)lms.size() / 2 - 1;
        lm_stix_to_id.push_back(id);
    }
    int NP() const { return (int)poses.size(); }
    int NL() const { return (int)lms.size() / 2; }
    int pose_stix(int id) const { return pose_id_to_stix.at(id); }  // throws std::out_of_range like map::at
    int lm_stix(int id) const { return lm_id_to_stix.at(id); }
    int default_pose_id() const { return pose_stix_to_id.at(0); }

    // ---------------- parse_g2o (utils/g2o_utils.cpp:10-146) -----------------------------
    // Values go through std::stof (FLOAT) and are widened to T afterwards.
    int load_g2o(const std::string& fname) {
        bound = 0; fixed_pose_id = -1;
        std::ifstream f(fname);
        if (!f.is_open()) return 1;
        std::string line;
        while (std::getline(f, line)) {
            std::istringstream ls(line);
            std::vector<std::string> tok;
            std::string t;
            while (ls >> t) tok.push_back(t);
            if (tok.empty()) continue;
            auto F = [&](size_t i) { return std::stof(tok.at(i)); };
            auto I = [&](size_t i) { return std::stoi(tok.at(i)); };
            auto grow = [&](float v) { if (std::abs(v) > bound) bound = std::abs(v); };
            if (tok[0] == "VERTEX_SE2") {
                int id = I(1); float x = F(2), y = F(3), th = F(4);
                grow(x); grow(y);
                // State::add_pose(float...) builds the pose with v2t in FLOAT in the reference;
                // the double oracle widens the parsed floats first and does the trig in T.
                add_pose((T)x, (T)y, (T)th, id);
            } else if (tok[0] == "VERTEX_XY") {
                int id = I(1); float x = F(2), y = F(3);
                grow(x); grow(y);
                add_landmark((T)x, (T)y, id);
            } else if (tok[0] == "FIX") {
                fixed_pose_id = I(1);
            } else if (tok[0] == "EDGE_SE2") {
                OdomObs<T> o;
                o.src_id = I(1); o.dst_id = I(2);
                o.z[0] = (T)F(3); o.z[1] = (T)F(4); o.z[2] = (T)F(5);
                float u[6];
                for (int k = 0; k < 6; k++) u[k] = F(6 + k);
                // upper-triangular, row-major: 00 01 02 11 12 22, mirrored
                o.omega[0] = (T)u[0]; o.omega[1] = (T)u[1]; o.omega[2] = (T)u[2];
                o.omega[3] = (T)u[1]; o.omega[4] = (T)u[3]; o.omega[5] = (T)u[4];
                o.omega[6] = (T)u[2]; o.omega[7] = (T)u[4]; o.omega[8] = (T)u[5];
                odoms.push_back(o);
            } else if (tok[0] == "EDGE_BEARING_SE2_XY") {
                BearingObs<T> b;
                b.pose_id = I(1); b.lm_id = I(2); b.bearing = (T)F(3); b.omega = 1;  // 4th number ignored
                bearings.push_back(b);
            } else {
                n_unrecognized++;
            }
        }
        bound += 3;
        return 0;
    }

    // ---------------- triangulation (slam/triangulation.cpp:5-74) ------------------------
    // Buckets by landmark id in a std::map (ascending id => defines landmark stix).
    void triangulate_landmarks() {
        std::map<int, std::vector<const BearingObs<T>*>> by_lm;
        for (const auto& o : bearings) by_lm[o.lm_id].push_back(&o);
        for (auto& kv : by_lm) {
            const int M = (int)kv.second.size();
            if (M == 1) single_observation_lms.push_back(kv.first);
            std::vector<T> A((size_t)M * 2), rhs(M);
            for (int i = 0; i < M; i++) {
                const auto& o = *kv.second[i];
                T px, py, th;
                t2v(poses[pose_stix(o.pose_id)], px, py, th);
                T bearing = smallest_angle<T>(o.bearing);
                T s = std::sin(th + bearing), c = std::cos(th + bearing);
                A[2 * i] = s; A[2 * i + 1] = -c;
                rhs[i] = s * px - c * py;
            }
            T lm[2];
            colpiv_householder_solve_Mx2<T>(A, rhs, M, lm);
            add_landmark(lm[0], lm[1], kv.first);
        }
    }

    // ---------------- Solver (slam/solver.cpp:5-18) --------------------------------------
    T kernel_threshold = 1, damping_factor = T(0.01f);
    // opt-in beyond the reference: IRLS flavour of the same threshold kernel -- the weight w = sqrt(kt / chi2) scales Omega
    // (H and b) instead of the error alone (b is identical, H += w J^T Omega J)
    bool irls = false;
    int N = 0, fixed_stix = 0;
    // resolved indices (id -> stix once; the reference does 2-3 map::at per edge per iteration)
    std::vector<int> b_pose, b_lm, o_src, o_dst;
    // block pattern: unique unordered block pairs in the unified block index space
    // (pose i -> i, landmark j -> NP + j), sorted; off[k] = (lo, hi), lo < hi.
    std::vector<std::pair<int, int>> off_pairs;
    std::vector<int> b_slot, o_slot;  // per-edge index into off_pairs
    std::vector<char> touched;        // block has at least one edge => full diagonal block in the pattern
    // values
    std::vector<T> Hdiag_p, Hdiag_l, Hoff, bvec;  // 9/pose, 4/lm, 9-stride per off pair (3x2 uses first 6)
    std::vector<T> err_b, jac_b, err_o, jac_o;    // per-edge debug terms (pre-kernel error)
    std::vector<T> delta;
    IterStats stats;
    // Test hook for the +-pi branch cut.  A bearing residual within 1e-9 of +-pi (on the bundled data: the single edge of each of
    // the single-observation landmarks 69 / 112 / 114, which the rank-1 triangulation puts exactly behind their pose) wraps to
    // +pi or -pi depending on the last bit of atan2, in the reference's own FP32 arithmetic as much as here.  Both branches are
    // valid linearizations; this map (bearing edge index -> +1 / -1) lets a test put the oracle on the branch the device took so
    // that iteration-0 b, dx and trajectories can be compared.  Only edges inside that 1e-9 neighbourhood are ever touched
    // (wrap_branch_tol; the comparison with the reference's own FP32 run widens it to float rounding, tests/test_ref_build.py).
    std::map<int, int> wrap_branch;
    double wrap_branch_tol = 1e-9;

    void solver_init(int fixed_id) {
        fixed_pose_id = fixed_id;
        N = 3 * NP() + 2 * NL();
        fixed_stix = pose_stix(fixed_id);
        const int np = NP();
        b_pose.resize(bearings.size()); b_lm.resize(bearings.size());
        o_src.resize(odoms.size()); o_dst.resize(odoms.size());
        std::vector<std::pair<int, int>> pairs;
        touched.assign(np + NL(), 0);
        for (size_t e = 0; e < bearings.size(); e++) {
            b_pose[e] = pose_stix(bearings[e].pose_id);
            b_lm[e] = lm_stix(bearings[e].lm_id);
            pairs.emplace_back(b_pose[e], np + b_lm[e]);
            touched[b_pose[e]] = 1; touched[np + b_lm[e]] = 1;
        }
        for (size_t e = 0; e < odoms.size(); e++) {
            o_src[e] = pose_stix(odoms[e].src_id);
            o_dst[e] = pose_stix(odoms[e].dst_id);
            if (o_src[e] == o_dst[e]) throw std::invalid_argument("odometry self-loop");
            pairs.emplace_back(std::min(o_src[e], o_dst[e]), std::max(o_src[e], o_dst[e]));
            touched[o_src[e]] = 1; touched[o_dst[e]] = 1;
        }
        off_pairs = pairs;
        std::sort(off_pairs.begin(), off_pairs.end());
        off_pairs.erase(std::unique(off_pairs.begin(), off_pairs.end()), off_pairs.end());
        auto slot_of = [&](const std::pair<int, int>& p) {
            return (int)(std::lower_bound(off_pairs.begin(), off_pairs.end(), p) - off_pairs.begin());
        };
        b_slot.resize(bearings.size()); o_slot.resize(odoms.size());
        for (size_t e = 0; e < bearings.size(); e++) b_slot[e] = slot_of(pairs[e]);
        for (size_t e = 0; e < odoms.size(); e++) o_slot[e] = slot_of(pairs[bearings.size() + e]);
        Hdiag_p.assign((size_t)np * 9, 0); Hdiag_l.assign((size_t)NL() * 4, 0);
        Hoff.assign(off_pairs.size() * 9, 0); bvec.assign(N, 0);
        delta.assign(N, 0);
    }

    // ---- predict_bearing (slam/solver_jacobians.cpp:301-305) ----------------------------
    static T predict_bearing(const Iso2<T>& pose, T lx, T ly) {
        T gx, gy;
        apply(inverse(pose), lx, ly, gx, gy);
        return std::atan2(gy, gx);
    }
    // ---- predict_odometry (slam/solver_jacobians.cpp:307-323) ---------------------------
    static void predict_odometry(const Iso2<T>& src, const Iso2<T>& dst, T pred[3]) {
        T sx, sy, sth, dx, dy, dth;
        t2v(src, sx, sy, sth); t2v(dst, dx, dy, dth);
        T tx = dx - sx, ty = dy - sy;
        pred[0] = src.r00 * tx + src.r10 * ty;  // R_s^T * t
        pred[1] = src.r01 * tx + src.r11 * ty;
        pred[2] = normalized_angle<T>(dth - sth);
    }

    // ---- bearing error + Jacobian (slam/solver_jacobians.cpp:9-95) ----------------------
    // J = [J_pose(3) | J_lm(2)]
    static void bearing_error_and_jacobian(const Iso2<T>& pose, T lx, T ly, T z, T& err, T J[5]) {
        T pred = predict_bearing(pose, lx, ly);
        err = normalized_angle<T>(pred - smallest_angle<T>(z));
        T gx, gy;
        apply(inverse(pose), lx, ly, gx, gy);
        T f = T(1) / (gx * gx + gy * gy);
        T a0 = f * (-gy), a1 = f * gx;
        // R^T
        T t00 = pose.r00, t01 = pose.r10, t10 = pose.r01, t11 = pose.r11;
        // (R^T * DR'^T) * lm with DR'^T = [[0,1],[-1,0]]
        T m00 = t00 * T(0) + t01 * T(-1), m01 = t00 * T(1) + t01 * T(0);
        T m10 = t10 * T(0) + t11 * T(-1), m11 = t10 * T(1) + t11 * T(0);
        T v0 = m00 * lx + m01 * ly, v1 = m10 * lx + m11 * ly;
        T G0[5] = {-t00, -t01, v0, t00, t01};
        T G1[5] = {-t10, -t11, v1, t10, t11};
        for (int j = 0; j < 5; j++) J[j] = a0 * G0[j] + a1 * G1[j];
    }

    // ---- odometry error + Jacobian (slam/solver_jacobians.cpp:97-168) -------------------
    // J is 3x6 row-major: [J_src(3x3) | J_dst(3x3)], explicit zeros included.
    static void odometry_error_and_jacobian(const Iso2<T>& src, const Iso2<T>& dst, const T z[3], T err[3], T J[18]) {
        T pred[3];
        predict_odometry(src, dst, pred);
        err[0] = pred[0] - z[0]; err[1] = pred[1] - z[1];
        err[2] = normalized_angle<T>(pred[2] - z[2]);
        T tdx = dst.tx, tdy = dst.ty;
        // (DR' * R_s)^T * t_d, DR' = [[0,-1],[1,0]]
        T p00 = T(0) * src.r00 + T(-1) * src.r10, p01 = T(0) * src.r01 + T(-1) * src.r11;
        T p10 = T(1) * src.r00 + T(0) * src.r10, p11 = T(1) * src.r01 + T(0) * src.r11;
        T ths0 = p00 * tdx + p10 * tdy, ths1 = p01 * tdx + p11 * tdy;
        // (R_s^T * DR') * t_d
        T q00 = src.r00 * T(0) + src.r10 * T(1), q01 = src.r00 * T(-1) + src.r10 * T(0);
        T q10 = src.r01 * T(0) + src.r11 * T(1), q11 = src.r01 * T(-1) + src.r11 * T(0);
        T thd0 = q00 * tdx + q01 * tdy, thd1 = q10 * tdx + q11 * tdy;
        T rows[3][6] = {
            {-src.r00, -src.r10, ths0, src.r00, src.r10, thd0},
            {-src.r01, -src.r11, ths1, src.r01, src.r11, thd1},
            {T(0), T(0), T(-1), T(0), T(0), T(1)}};
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 6; j++) J[i * 6 + j] = rows[i][j];
    }

    // ---- numeric Jacobians, central differences eps = 1e-3 through boxplus
    //      (slam/solver_jacobians.cpp:170-299); validation only.
    static void bearing_numeric_jacobian(const Iso2<T>& pose, T lx, T ly, T z, T J[5]) {
        const T eps = T(0.001f);
        auto err_at = [&](T dx, T dy, T dth, T dlx, T dly) {
            T pred = predict_bearing(boxplus(pose, dx, dy, dth), lx + dlx, ly + dly);
            return normalized_angle<T>(pred - smallest_angle<T>(z));
        };
        for (int k = 0; k < 5; k++) {
            T s[5] = {0, 0, 0, 0, 0};
            s[k] = 1;
            T ep = err_at(eps * s[0], eps * s[1], eps * s[2], eps * s[3], eps * s[4]);
            T em = err_at(-(eps * s[0]), -(eps * s[1]), -(eps * s[2]), -(eps * s[3]), -(eps * s[4]));
            J[k] = (ep - em) / (2 * eps);
        }
    }
    static void odometry_numeric_jacobian(const Iso2<T>& src, const Iso2<T>& dst, const T z[3], T J[18]) {
        const T eps = T(0.001f);
        auto err_at = [&](const T ds[3], const T dd[3], T out[3]) {
            T pred[3];
            predict_odometry(boxplus(src, ds[0], ds[1], ds[2]), boxplus(dst, dd[0], dd[1], dd[2]), pred);
            out[0] = pred[0] - z[0]; out[1] = pred[1] - z[1];
            out[2] = normalized_angle<T>(pred[2] - z[2]);
        };
        for (int k = 0; k < 6; k++) {
            T ds[3] = {0, 0, 0}, dd[3] = {0, 0, 0}, nds[3], ndd[3];
            if (k < 3) ds[k] = eps; else dd[k - 3] = eps;
            for (int i = 0; i < 3; i++) { nds[i] = -ds[i]; ndd[i] = -dd[i]; }
            T ep[3], em[3];
            err_at(ds, dd, ep); err_at(nds, ndd, em);
            for (int i = 0; i < 3; i++) J[i * 6 + k] = (ep[i] - em[i]) / (2 * eps);
        }
    }

    // ---- H, b accumulation (slam/solver.cpp:27-69) --------------------------------------
    // Same per-entry arithmetic as H += J^T * Omega * J, b += J^T * Omega * e, sequential in
    // edge order (all bearing edges, then all odometry edges, then damping), but O(1) per edge
    // through precomputed block slots instead of the reference's O(N + nnz H) sparse merge.
    void linearize() {
        const int np = NP();
        std::fill(Hdiag_p.begin(), Hdiag_p.end(), T(0));
        std::fill(Hdiag_l.begin(), Hdiag_l.end(), T(0));
        std::fill(Hoff.begin(), Hoff.end(), T(0));
        std::fill(bvec.begin(), bvec.end(), T(0));
        err_b.resize(bearings.size()); jac_b.resize(bearings.size() * 5);
        err_o.resize(odoms.size() * 3); jac_o.resize(odoms.size() * 18);
        stats = IterStats();
        for (size_t e = 0; e < bearings.size(); e++) {
            const int p = b_pose[e], l = b_lm[e];
            T err, J[5];
            bearing_error_and_jacobian(poses[p], lms[2 * l], lms[2 * l + 1], bearings[e].bearing, err, J);
            if (!wrap_branch.empty()) {   // test hook: see wrap_branch
                auto it = wrap_branch.find((int)e);
                if (it != wrap_branch.end() && std::abs(std::abs((double)err) - kCvPi) < wrap_branch_tol) err = (T)(it->second * std::abs((double)err));
            }
            err_b[e] = err;
            for (int j = 0; j < 5; j++) jac_b[e * 5 + j] = J[j];
            const T om = bearings[e].omega;
            T chi = err * om * err;
            stats.chi2_bearing += (double)chi;
            T om_h = om;     // the omega the H terms see
            if (chi > kernel_threshold) {
                const T wgt = std::sqrt(kernel_threshold / chi);
                err *= wgt;                       // slam/solver.cpp:37-41 (b sees w e either way)
                if (irls) om_h = om * wgt;
                stats.over_bearing++;
            }
            T* Hp = &Hdiag_p[(size_t)p * 9];
            T* Hl = &Hdiag_l[(size_t)l * 4];
            T* Hpl = &Hoff[(size_t)b_slot[e] * 9];
            for (int a = 0; a < 3; a++) {
                T ja = J[a] * om, jh = J[a] * om_h;
                for (int c = 0; c < 3; c++) Hp[a * 3 + c] += jh * J[c];
                for (int c = 0; c < 2; c++) Hpl[a * 2 + c] += jh * J[3 + c];
                bvec[3 * p + a] += ja * err;
            }
            for (int a = 0; a < 2; a++) {
                T ja = J[3 + a] * om, jh = J[3 + a] * om_h;
                for (int c = 0; c < 2; c++) Hl[a * 2 + c] += jh * J[3 + c];
                bvec[3 * np + 2 * l + a] += ja * err;
            }
        }
        for (size_t e = 0; e < odoms.size(); e++) {
            const int s = o_src[e], d = o_dst[e];
            T err[3], J[18];
            odometry_error_and_jacobian(poses[s], poses[d], odoms[e].z, err, J);
            for (int i = 0; i < 3; i++) err_o[e * 3 + i] = err[i];
            for (int i = 0; i < 18; i++) jac_o[e * 18 + i] = J[i];
            const T* Om = odoms[e].omega;
            T eo[3];
            for (int k = 0; k < 3; k++) eo[k] = err[0] * Om[0 * 3 + k] + err[1] * Om[1 * 3 + k] + err[2] * Om[2 * 3 + k];
            T chi = eo[0] * err[0] + eo[1] * err[1] + eo[2] * err[2];
            stats.chi2_odometry += (double)chi;
            T wh = 1;        // weight of the H terms
            if (chi > kernel_threshold) {
                T sc = std::sqrt(kernel_threshold / chi);
                for (int i = 0; i < 3; i++) err[i] *= sc;
                if (irls) wh = sc;
                stats.over_odometry++;
            }
            // JtO = J^T * Omega (6x3)
            T JtO[6][3];
            for (int a = 0; a < 6; a++)
                for (int k = 0; k < 3; k++)
                    JtO[a][k] = J[0 * 6 + a] * Om[0 * 3 + k] + J[1 * 6 + a] * Om[1 * 3 + k] + J[2 * 6 + a] * Om[2 * 3 + k];
            auto HJ = [&](int a, int c) { return wh * (JtO[a][0] * J[0 * 6 + c] + JtO[a][1] * J[1 * 6 + c] + JtO[a][2] * J[2 * 6 + c]); };
            T* Hs = &Hdiag_p[(size_t)s * 9];
            T* Hd = &Hdiag_p[(size_t)d * 9];
            T* Ho = &Hoff[(size_t)o_slot[e] * 9];
            const bool src_is_lo = s < d;
            for (int a = 0; a < 3; a++)
                for (int c = 0; c < 3; c++) {
                    Hs[a * 3 + c] += HJ(a, c);
                    Hd[a * 3 + c] += HJ(3 + a, 3 + c);
                    // canonical block is H[lo][hi]; for src > dst it is (J_s^T O J_d)^T = J_d^T O J_s
                    if (src_is_lo) Ho[a * 3 + c] += HJ(a, 3 + c);
                    else Ho[a * 3 + c] += HJ(3 + a, c);
                }
            for (int a = 0; a < 3; a++) {
                bvec[3 * s + a] += JtO[a][0] * err[0] + JtO[a][1] * err[1] + JtO[a][2] * err[2];
                bvec[3 * d + a] += JtO[3 + a][0] * err[0] + JtO[3 + a][1] * err[1] + JtO[3 + a][2] * err[2];
            }
        }
        // damping (slam/solver.cpp:64-69): H += damping_factor * I, every iteration
        for (int i = 0; i < np; i++)
            for (int a = 0; a < 3; a++) Hdiag_p[(size_t)i * 9 + a * 4] += damping_factor;
        for (int j = 0; j < NL(); j++)
            for (int a = 0; a < 2; a++) Hdiag_l[(size_t)j * 4 + a * 3] += damping_factor;
    }

    // scalar index of block b in the full delta vector, and its dimension
    int blk_start(int b) const { return b < NP() ? 3 * b : 3 * NP() + 2 * (b - NP()); }
    int blk_dim(int b) const { return b < NP() ? 3 : 2; }
    // gauge fix (slam/solver.cpp:72-73, 99-125): scalar index after deleting rows/cols 3f..3f+2
    int nofixed_index(int i) const {
        const int f3 = 3 * fixed_stix;
        if (i < f3) return i;
        if (i < f3 + 3) return -1;
        return i - 3;
    }

    // H_nofixed as scalar CSC with both triangles and sorted row indices, b_nofixed.
    void export_csc(Csc<T>& out, std::vector<T>& b_nofixed) const {
        const int nb = NP() + NL();
        const int n = N - 3;
        struct Ent { int row; T v; };
        std::vector<std::vector<Ent>> cols(n);
        auto put = [&](int gi, int gj, T v) {
            int i = nofixed_index(gi), j = nofixed_index(gj);
            if (i < 0 || j < 0) return;
            cols[j].push_back({i, v});
        };
        for (int b = 0; b < nb; b++) {
            const int s0 = blk_start(b), d = blk_dim(b);
            const T* blk = b < NP() ? &Hdiag_p[(size_t)b * 9] : &Hdiag_l[(size_t)(b - NP()) * 4];
            for (int a = 0; a < d; a++)
                for (int c = 0; c < d; c++)
                    if (touched[b] || a == c) put(s0 + a, s0 + c, blk[a * d + c]);
        }
        for (size_t k = 0; k < off_pairs.size(); k++) {
            const int lo = off_pairs[k].first, hi = off_pairs[k].second;
            const int s0 = blk_start(lo), s1 = blk_start(hi), d0 = blk_dim(lo), d1 = blk_dim(hi);
            const T* blk = &Hoff[k * 9];
            for (int a = 0; a < d0; a++)
                for (int c = 0; c < d1; c++) {
                    put(s0 + a, s1 + c, blk[a * d1 + c]);
                    put(s1 + c, s0 + a, blk[a * d1 + c]);
                }
        }
        out.n = n; out.colptr.assign(n + 1, 0); out.rowidx.clear(); out.val.clear();
        for (int j = 0; j < n; j++) {
            std::sort(cols[j].begin(), cols[j].end(), [](const Ent& x, const Ent& y) { return x.row < y.row; });
            for (const auto& en : cols[j]) { out.rowidx.push_back(en.row); out.val.push_back(en.v); }
            out.colptr[j + 1] = (int)out.rowidx.size();
        }
        b_nofixed.assign(n, 0);
        for (int i = 0; i < N; i++) {
            int k = nofixed_index(i);
            if (k >= 0) b_nofixed[k] = bvec[i];
        }
    }

    // ---- dense LDL^T solve of H_nofixed * dx = -b_nofixed (slam/solver.cpp:77-94) -------
    // SimplicialLDLT semantics: no pivoting, lower triangle only; Eigen's AMD ordering only
    // changes rounding.  Dense storage: meant for n up to a few thousand.
    void solve_dense_ldlt() {
        Csc<T> A; std::vector<T> bn;
        export_csc(A, bn);
        const int n = A.n;
        std::vector<T> L((size_t)n * n, 0);
        for (int j = 0; j < n; j++)
            for (int k = A.colptr[j]; k < A.colptr[j + 1]; k++)
                if (A.rowidx[k] >= j) L[(size_t)A.rowidx[k] * n + j] = A.val[k];
        std::vector<T> D(n);
        stats.solver_status = 0;
        for (int j = 0; j < n; j++) {
            T d = L[(size_t)j * n + j];
            for (int k = 0; k < j; k++) d -= L[(size_t)j * n + k] * L[(size_t)j * n + k] * D[k];
            D[j] = d;
            if (!(d > 0)) stats.solver_status = 1;
            for (int i = j + 1; i < n; i++) {
                T* Li = &L[(size_t)i * n];
                const T* Lj = &L[(size_t)j * n];
                T s = Li[j];
                for (int k = 0; k < j; k++) s -= Li[k] * Lj[k] * D[k];
                Li[j] = s / d;
            }
        }
        std::vector<T> x(n);
        for (int i = 0; i < n; i++) {
            T s = -bn[i];
            for (int k = 0; k < i; k++) s -= L[(size_t)i * n + k] * x[k];
            x[i] = s;
        }
        for (int i = 0; i < n; i++) x[i] /= D[i];
        for (int i = n - 1; i >= 0; i--) {
            T s = x[i];
            for (int k = i + 1; k < n; k++) s -= L[(size_t)k * n + i] * x[k];
            x[i] = s;
        }
        scatter_delta(x);
    }

    // ---- sparse LDL^T solve of H_nofixed * dx = -b_nofixed: the reference's own solver (slam/solver.hpp:72, solver.cpp:75-94) ----
    // SimplicialLDLT restated in bos_sparse_ldlt.hpp: fill-reducing minimum-degree ordering + symbolic analysis ONCE per problem
    // (analyzePattern, guarded by analyzed_H in the reference), numeric up-looking factorisation + solve per step.
    SparseLdlt<T> ldlt;
    double t_order = 0, t_analyze = 0, t_factor = 0, t_trisolve = 0, t_export = 0;
    bool solve_sparse_ldlt(double deadline_s = 0.0) {
        using clk = std::chrono::steady_clock;
        auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
        auto t0 = clk::now();
        Csc<T> A; std::vector<T> bn;
        export_csc(A, bn);
        auto t1 = clk::now();
        t_export = secs(t0, t1);
        if (!ldlt.analyzed) {
            // block graph without the fixed pose: node = block, weight = its scalar dimension
            const int nb = NP() + NL();
            std::vector<int> node_of(nb, -1), blocks;
            for (int b = 0; b < nb; b++)
                if (b != fixed_stix) { node_of[b] = (int)blocks.size(); blocks.push_back(b); }
            const int nn = (int)blocks.size();
            std::vector<int> xadj(nn + 1, 0), weight(nn);
            for (int k = 0; k < nn; k++) weight[k] = blk_dim(blocks[k]);
            for (const auto& pr : off_pairs) {
                const int a = node_of[pr.first], b = node_of[pr.second];
                if (a >= 0 && b >= 0) { xadj[a + 1]++; xadj[b + 1]++; }
            }
            for (int k = 0; k < nn; k++) xadj[k + 1] += xadj[k];
            std::vector<int> adjncy(xadj[nn]), cur(xadj.begin(), xadj.end() - 1);
            for (const auto& pr : off_pairs) {
                const int a = node_of[pr.first], b = node_of[pr.second];
                if (a >= 0 && b >= 0) { adjncy[cur[a]++] = b; adjncy[cur[b]++] = a; }
            }
            std::vector<int> order = min_degree_order(nn, xadj, adjncy, weight);
            std::vector<int> perm;
            perm.reserve(A.n);
            for (int k : order) {
                const int b = blocks[k], s0 = blk_start(b);
                for (int a = 0; a < blk_dim(b); a++) perm.push_back(nofixed_index(s0 + a));
            }
            auto t2 = clk::now();
            t_order = secs(t1, t2);
            ldlt.analyze(A.n, A.colptr, A.rowidx, perm);
            t_analyze = secs(t2, clk::now());
        }
        auto t3 = clk::now();
        if (!ldlt.factorize(A.colptr, A.rowidx, A.val, deadline_s)) { t_factor = secs(t3, clk::now()); return false; }
        auto t4 = clk::now();
        t_factor = secs(t3, t4);
        stats.solver_status = ldlt.status;
        std::vector<T> rhs(A.n), x;
        for (int i = 0; i < A.n; i++) rhs[i] = -bn[i];
        ldlt.solve(rhs, x);
        t_trisolve = secs(t4, clk::now());
        scatter_delta(x);
        return true;
    }

    // ---- the LITERAL accumulation of slam/solver.cpp:31-62: every edge builds its J^T Omega J as an N x N column-major sparse
    // matrix (O(N) column pointers) and H += that merges ALL of H (O(N + nnz H)); b is dense.  Same result as linearize(); kept
    // for timing the reference's actual per-iteration cost on the bundled datasets (infeasible at the synthetic sizes).
    struct LitSparse { std::vector<int> colptr, rowidx; std::vector<T> val; };
    void literal_merge(LitSparse& Hm, const int* cols, int nc, const T* blk) const {   // blk: nc x nc dense, indices ascending
        LitSparse Tm;
        Tm.colptr.assign(N + 1, 0);
        for (int c = 0; c < nc; c++) Tm.colptr[cols[c] + 1] = nc;
        for (int j = 0; j < N; j++) Tm.colptr[j + 1] += Tm.colptr[j];
        Tm.rowidx.resize((size_t)nc * nc); Tm.val.resize((size_t)nc * nc);
        for (int c = 0; c < nc; c++)
            for (int r = 0; r < nc; r++) { Tm.rowidx[Tm.colptr[cols[c]] + r] = cols[r]; Tm.val[Tm.colptr[cols[c]] + r] = blk[r * nc + c]; }
        LitSparse R;
        R.colptr.assign(N + 1, 0);
        R.rowidx.reserve(Hm.rowidx.size() + Tm.rowidx.size()); R.val.reserve(Hm.val.size() + Tm.val.size());
        for (int j = 0; j < N; j++) {
            int a = Hm.colptr[j], ae = Hm.colptr[j + 1], b = Tm.colptr[j], be = Tm.colptr[j + 1];
            while (a < ae || b < be) {
                if (b >= be || (a < ae && Hm.rowidx[a] < Tm.rowidx[b])) { R.rowidx.push_back(Hm.rowidx[a]); R.val.push_back(Hm.val[a]); a++; }
                else if (a >= ae || Tm.rowidx[b] < Hm.rowidx[a]) { R.rowidx.push_back(Tm.rowidx[b]); R.val.push_back(Tm.val[b]); b++; }
                else { R.rowidx.push_back(Hm.rowidx[a]); R.val.push_back(Hm.val[a] + Tm.val[b]); a++; b++; }
            }
            R.colptr[j + 1] = (int)R.rowidx.size();
        }
        Hm.colptr.swap(R.colptr); Hm.rowidx.swap(R.rowidx); Hm.val.swap(R.val);
    }
    LitSparse Hlit;
    void linearize_literal() {
        const int np = NP();
        Hlit.colptr.assign(N + 1, 0); Hlit.rowidx.clear(); Hlit.val.clear();
        std::fill(bvec.begin(), bvec.end(), T(0));
        for (size_t e = 0; e < bearings.size(); e++) {
            const int p = b_pose[e], l = b_lm[e];
            T err, J[5];
            bearing_error_and_jacobian(poses[p], lms[2 * l], lms[2 * l + 1], bearings[e].bearing, err, J);
            const T om = bearings[e].omega;
            T chi = err * om * err;
            if (chi > kernel_threshold) err *= std::sqrt(kernel_threshold / chi);
            const int cols[5] = {3 * p, 3 * p + 1, 3 * p + 2, 3 * np + 2 * l, 3 * np + 2 * l + 1};
            T blk[25];
            for (int a = 0; a < 5; a++)
                for (int c = 0; c < 5; c++) blk[a * 5 + c] = J[a] * om * J[c];
            literal_merge(Hlit, cols, 5, blk);
            for (int a = 0; a < 5; a++) bvec[cols[a]] += J[a] * om * err;
        }
        for (size_t e = 0; e < odoms.size(); e++) {
            const int s = o_src[e], d = o_dst[e];
            T err[3], J[18];
            odometry_error_and_jacobian(poses[s], poses[d], odoms[e].z, err, J);
            const T* Om = odoms[e].omega;
            T eo[3];
            for (int k = 0; k < 3; k++) eo[k] = err[0] * Om[0 * 3 + k] + err[1] * Om[1 * 3 + k] + err[2] * Om[2 * 3 + k];
            T chi = eo[0] * err[0] + eo[1] * err[1] + eo[2] * err[2];
            if (chi > kernel_threshold) { T sc = std::sqrt(kernel_threshold / chi); for (int i = 0; i < 3; i++) err[i] *= sc; }
            const int lo = std::min(s, d), hi = std::max(s, d);
            const int off_s = s < d ? 0 : 3, off_d = s < d ? 3 : 0;   // position of each pose's columns in the ascending index list
            int cols[6];
            for (int a = 0; a < 3; a++) { cols[a] = 3 * lo + a; cols[3 + a] = 3 * hi + a; }
            T JtO[6][3], Jc[3][6];
            for (int i = 0; i < 3; i++)
                for (int a = 0; a < 3; a++) { Jc[i][off_s + a] = J[i * 6 + a]; Jc[i][off_d + a] = J[i * 6 + 3 + a]; }
            for (int a = 0; a < 6; a++)
                for (int k = 0; k < 3; k++) JtO[a][k] = Jc[0][a] * Om[0 * 3 + k] + Jc[1][a] * Om[1 * 3 + k] + Jc[2][a] * Om[2 * 3 + k];
            T blk[36];
            for (int a = 0; a < 6; a++)
                for (int c = 0; c < 6; c++) blk[a * 6 + c] = JtO[a][0] * Jc[0][c] + JtO[a][1] * Jc[1][c] + JtO[a][2] * Jc[2][c];
            literal_merge(Hlit, cols, 6, blk);
            for (int a = 0; a < 6; a++) bvec[cols[a]] += JtO[a][0] * err[0] + JtO[a][1] * err[1] + JtO[a][2] * err[2];
        }
        // damping: H += damping_factor * I (solver.cpp:64-69), one more full merge
        LitSparse R;
        R.colptr.assign(N + 1, 0);
        for (int j = 0; j < N; j++) {
            bool done = false;
            for (int q = Hlit.colptr[j]; q < Hlit.colptr[j + 1]; q++) {
                if (!done && Hlit.rowidx[q] > j) { R.rowidx.push_back(j); R.val.push_back(damping_factor); done = true; }
                R.rowidx.push_back(Hlit.rowidx[q]);
                R.val.push_back(Hlit.val[q] + ((Hlit.rowidx[q] == j) ? damping_factor : T(0)));
                if (Hlit.rowidx[q] == j) done = true;
            }
            if (!done) { R.rowidx.push_back(j); R.val.push_back(damping_factor); }
            R.colptr[j + 1] = (int)R.rowidx.size();
        }
        Hlit.colptr.swap(R.colptr); Hlit.rowidx.swap(R.rowidx); Hlit.val.swap(R.val);
    }

    void scatter_delta(const std::vector<T>& x_nofixed) {
        delta.assign(N, 0);
        double m = 0;
        for (int i = 0; i < N; i++) {
            int k = nofixed_index(i);
            delta[i] = k >= 0 ? x_nofixed[k] : T(0);
            m = std::max(m, std::abs((double)delta[i]));
        }
        stats.delta_inf = m;
    }

    // ---- Schur complement on the landmarks + block-Jacobi PCG on the pose system --------
    // Same mathematics as the exact solve (elimination of the 2x2 landmark blocks is exact);
    // this is the CPU restatement of the CUDA large-problem solver, used as its cross-check and
    // as the cpu_baseline at sizes where a dense factorisation is infeasible.
    // The fixed pose is handled by zeroing its couplings: its 3x3 block is then only damping
    // and its rhs is zero, so dx_fixed = 0 exactly as if its rows/cols had been deleted.
    int solve_schur_pcg(int max_iters, double rtol) {
        const int np = NP(), nl = NL();
        std::vector<T> Linv((size_t)nl * 4);
        for (int j = 0; j < nl; j++) {
            const T* h = &Hdiag_l[(size_t)j * 4];
            T det = h[0] * h[3] - h[1] * h[2];
            T* o = &Linv[(size_t)j * 4];
            o[0] = h[3] / det; o[1] = -h[1] / det; o[2] = -h[2] / det; o[3] = h[0] / det;
        }
        auto is_fixed = [&](int b) { return b == fixed_stix; };
        // reduced rhs g = -(b_p - Hpl Hll^-1 b_l)
        std::vector<T> g((size_t)3 * np), tl((size_t)2 * nl), ul((size_t)2 * nl);
        for (int j = 0; j < nl; j++) {
            const T* o = &Linv[(size_t)j * 4];
            T b0 = bvec[3 * np + 2 * j], b1 = bvec[3 * np + 2 * j + 1];
            ul[2 * j] = o[0] * b0 + o[1] * b1; ul[2 * j + 1] = o[2] * b0 + o[3] * b1;
        }
        for (int i = 0; i < 3 * np; i++) g[i] = -bvec[i];
        for (size_t k = 0; k < off_pairs.size(); k++) {
            int lo = off_pairs[k].first, hi = off_pairs[k].second;
            if (hi < np || is_fixed(lo)) continue;
            int j = hi - np;
            const T* B = &Hoff[k * 9];
            for (int a = 0; a < 3; a++) g[3 * lo + a] += B[a * 2] * ul[2 * j] + B[a * 2 + 1] * ul[2 * j + 1];
        }
        for (int a = 0; a < 3; a++) g[3 * fixed_stix + a] = 0;
        // block-Jacobi preconditioner: inverse of diag blocks of S
        std::vector<T> Sd((size_t)np * 9);
        for (int i = 0; i < np; i++)
            for (int a = 0; a < 9; a++) Sd[(size_t)i * 9 + a] = Hdiag_p[(size_t)i * 9 + a];
        for (size_t k = 0; k < off_pairs.size(); k++) {
            int lo = off_pairs[k].first, hi = off_pairs[k].second;
            if (hi < np || is_fixed(lo)) continue;
            const T* B = &Hoff[k * 9];
            const T* o = &Linv[(size_t)(hi - np) * 4];
            T Y[6];
            for (int a = 0; a < 3; a++) {
                Y[a * 2] = B[a * 2] * o[0] + B[a * 2 + 1] * o[2];
                Y[a * 2 + 1] = B[a * 2] * o[1] + B[a * 2 + 1] * o[3];
            }
            for (int a = 0; a < 3; a++)
                for (int c = 0; c < 3; c++) Sd[(size_t)lo * 9 + a * 3 + c] -= Y[a * 2] * B[c * 2] + Y[a * 2 + 1] * B[c * 2 + 1];
        }
        if (true) {  // the fixed pose keeps only damping on its diagonal block
            T* s = &Sd[(size_t)fixed_stix * 9];
            for (int a = 0; a < 9; a++) s[a] = 0;
            s[0] = s[4] = s[8] = damping_factor;
        }
        std::vector<T> Minv((size_t)np * 9);
        for (int i = 0; i < np; i++) inv3(&Sd[(size_t)i * 9], &Minv[(size_t)i * 9]);
        auto matvec = [&](const std::vector<T>& x, std::vector<T>& y) {
            std::fill(tl.begin(), tl.end(), T(0));
            for (int i = 0; i < np; i++) {
                const T* h = is_fixed(i) ? nullptr : &Hdiag_p[(size_t)i * 9];
                for (int a = 0; a < 3; a++)
                    y[3 * i + a] = h ? h[a * 3] * x[3 * i] + h[a * 3 + 1] * x[3 * i + 1] + h[a * 3 + 2] * x[3 * i + 2]
                                     : damping_factor * x[3 * i + a];
            }
            for (size_t k = 0; k < off_pairs.size(); k++) {
                int lo = off_pairs[k].first, hi = off_pairs[k].second;
                const T* B = &Hoff[k * 9];
                if (hi < np) {
                    if (is_fixed(lo) || is_fixed(hi)) continue;
                    for (int a = 0; a < 3; a++)
                        for (int c = 0; c < 3; c++) {
                            y[3 * lo + a] += B[a * 3 + c] * x[3 * hi + c];
                            y[3 * hi + c] += B[a * 3 + c] * x[3 * lo + a];
                        }
                } else {
                    if (is_fixed(lo)) continue;
                    int j = hi - np;
                    for (int a = 0; a < 3; a++) {
                        tl[2 * j] += B[a * 2] * x[3 * lo + a];
                        tl[2 * j + 1] += B[a * 2 + 1] * x[3 * lo + a];
                    }
                }
            }
            for (int j = 0; j < nl; j++) {
                const T* o = &Linv[(size_t)j * 4];
                ul[2 * j] = o[0] * tl[2 * j] + o[1] * tl[2 * j + 1];
                ul[2 * j + 1] = o[2] * tl[2 * j] + o[3] * tl[2 * j + 1];
            }
            for (size_t k = 0; k < off_pairs.size(); k++) {
                int lo = off_pairs[k].first, hi = off_pairs[k].second;
                if (hi < np || is_fixed(lo)) continue;
                int j = hi - np;
                const T* B = &Hoff[k * 9];
                for (int a = 0; a < 3; a++) y[3 * lo + a] -= B[a * 2] * ul[2 * j] + B[a * 2 + 1] * ul[2 * j + 1];
            }
        };
        const int n = 3 * np;
        std::vector<T> x(n, 0), r = g, z(n), p(n), Sp(n);
        auto precond = [&](const std::vector<T>& rr, std::vector<T>& zz) {
            for (int i = 0; i < np; i++) {
                const T* m = &Minv[(size_t)i * 9];
                for (int a = 0; a < 3; a++) zz[3 * i + a] = m[a * 3] * rr[3 * i] + m[a * 3 + 1] * rr[3 * i + 1] + m[a * 3 + 2] * rr[3 * i + 2];
            }
        };
        auto dot = [&](const std::vector<T>& u, const std::vector<T>& v) {
            double s = 0;
            for (int i = 0; i < n; i++) s += (double)u[i] * (double)v[i];
            return s;
        };
        precond(r, z); p = z;
        double rz = dot(r, z), rz0 = rz;
        int it = 0;
        for (; it < max_iters && rz > rtol * rtol * rz0 && rz > 0; it++) {
            matvec(p, Sp);
            double alpha = rz / dot(p, Sp);
            for (int i = 0; i < n; i++) { x[i] += (T)alpha * p[i]; r[i] -= (T)alpha * Sp[i]; }
            precond(r, z);
            double rz_new = dot(r, z), beta = rz_new / rz;
            rz = rz_new;
            for (int i = 0; i < n; i++) p[i] = z[i] + (T)beta * p[i];
        }
        stats.pcg_iterations = it;
        // back-substitution: dx_l = Hll^-1 (-b_l - Hlp dx_p)
        delta.assign(N, 0);
        for (int i = 0; i < n; i++) delta[i] = x[i];
        for (int a = 0; a < 3; a++) delta[3 * fixed_stix + a] = 0;
        for (int j = 0; j < nl; j++) { tl[2 * j] = -bvec[3 * np + 2 * j]; tl[2 * j + 1] = -bvec[3 * np + 2 * j + 1]; }
        for (size_t k = 0; k < off_pairs.size(); k++) {
            int lo = off_pairs[k].first, hi = off_pairs[k].second;
            if (hi < np || is_fixed(lo)) continue;
            int j = hi - np;
            const T* B = &Hoff[k * 9];
            for (int a = 0; a < 3; a++) {
                tl[2 * j] -= B[a * 2] * delta[3 * lo + a];
                tl[2 * j + 1] -= B[a * 2 + 1] * delta[3 * lo + a];
            }
        }
        double m = 0;
        for (int j = 0; j < nl; j++) {
            const T* o = &Linv[(size_t)j * 4];
            delta[3 * np + 2 * j] = o[0] * tl[2 * j] + o[1] * tl[2 * j + 1];
            delta[3 * np + 2 * j + 1] = o[2] * tl[2 * j] + o[3] * tl[2 * j + 1];
        }
        for (int i = 0; i < N; i++) m = std::max(m, std::abs((double)delta[i]));
        stats.delta_inf = m;
        return it;
    }

    static void inv3(const T* a, T* o) {
        T c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
        T det = a[0] * c00 + a[1] * c01 + a[2] * c02;
        T id = T(1) / det;
        o[0] = c00 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
        o[3] = c01 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
        o[6] = c02 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
    }

    // ---- State::apply_boxplus (framework/state.cpp:69-80) -------------------------------
    void apply_boxplus() {
        const int np = NP();
        for (int i = 0; i < np; i++) poses[i] = boxplus(poses[i], delta[3 * i], delta[3 * i + 1], delta[3 * i + 2]);
        for (int j = 0; j < NL(); j++) {
            lms[2 * j] += delta[3 * np + 2 * j];
            lms[2 * j + 1] += delta[3 * np + 2 * j + 1];
        }
    }

    // ---- Solver::step (slam/solver.cpp:27-97): exactly one GN iteration incl. update ----
    // solver_kind 0: dense LDL^T on H_nofixed (the reference's mathematics);
    //             1: Schur + block-Jacobi PCG.
    //             2: sparse LDL^T with a cached symbolic phase (the reference's SimplicialLDLT).
    void step(int solver_kind, int pcg_max_iters = 2000, double pcg_rtol = 1e-12) {
        linearize();
        if (solver_kind == 0) solve_dense_ldlt();
        else if (solver_kind == 2) solve_sparse_ldlt();
        else solve_schur_pcg(pcg_max_iters, pcg_rtol);
        apply_boxplus();
    }
};

}  // namespace bos_oracle
