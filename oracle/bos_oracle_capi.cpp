// bos_oracle_capi.cpp -- flat C entry points over bos_oracle.hpp for ctypes.
// CPU ORACLE. TEST INFRASTRUCTURE ONLY (see the header of bos_oracle.hpp): loaded by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
// All arrays cross this boundary as double / int32 whatever the internal scalar is.
#include "bos_oracle.hpp"

#include <chrono>
#include <cstring>
#include <memory>

using namespace bos_oracle;

namespace {

struct IOracle {
    virtual ~IOracle() {}
    virtual int load_g2o(const char* path) = 0;
    virtual void set_poses(int n, const int* ids, const double* xyt) = 0;
    virtual void set_landmarks(int n, const int* ids, const double* xy) = 0;
    virtual void set_bearings(int n, const int* pose_ids, const int* lm_ids, const double* z, const double* om) = 0;
    virtual void set_odometry(int n, const int* src, const int* dst, const double* z3, const double* om9) = 0;
    virtual void set_fixed(int id) = 0;
    virtual void triangulate() = 0;
    virtual void counts(int out[8]) = 0;
    virtual double get_bound() = 0;
    virtual void get_ids(int* pose_ids, int* lm_ids) = 0;
    virtual void get_state(double* poses_xycs, double* lms_xy) = 0;
    virtual void set_state(const double* poses_xycs, const double* lms_xy) = 0;
    virtual void get_state_xyt(double* poses_xyt) = 0;
    virtual void get_edges(int* bp, int* bl, double* bz, double* bom, int* os, int* od, double* oz, double* oom) = 0;
    virtual int get_single_obs(int* ids, int cap) = 0;
    virtual int solver_init(int fixed_id) = 0;
    virtual void set_params(double kt, double damp) = 0;
    virtual void set_irls(int on) = 0;
    virtual void get_edge_stix(int* bp, int* bl, int* os, int* od) = 0;
    virtual void linearize() = 0;
    virtual int n_off() = 0;
    virtual void get_blocks(double* hp, double* hl, double* hoff, int* off_lo, int* off_hi, double* b) = 0;
    virtual void get_edge_terms(double* eb, double* jb, double* eo, double* jo) = 0;
    virtual int csc_nnz() = 0;
    virtual void get_csc(int* colptr, int* rowidx, double* val, double* b) = 0;
    virtual void solve(int kind, int max_iters, double rtol) = 0;
    virtual void get_delta(double* d) = 0;
    virtual void set_delta(const double* d) = 0;
    virtual void apply_boxplus() = 0;
    virtual void step(int kind, int max_iters, double rtol) = 0;
    virtual void get_stats(double out[8]) = 0;
    virtual double predict_bearing(double x, double y, double th, double lx, double ly) = 0;
    virtual void predict_odometry(const double* s_xyt, const double* d_xyt, double* out3) = 0;
    virtual void bearing_jacobians(int e, double* ana5, double* num5) = 0;
    virtual void odometry_jacobians(int e, double* ana18, double* num18) = 0;
    virtual double time_linearize(int reps) = 0;
    virtual int solve_sparse(double deadline_s, double* info10) = 0;
    virtual double time_linearize_literal(int reps) = 0;
    virtual double literal_max_diff() = 0;
    virtual void set_wrap_branch(int n, const int* edges, const int* signs) = 0;
    virtual void set_wrap_branch_tol(double tol) = 0;
};

template <class T>
struct Impl : IOracle {
    Oracle<T> o;
    Csc<T> csc; std::vector<T> bn; bool csc_valid = false;

    int load_g2o(const char* path) override { return o.load_g2o(path); }
    void set_poses(int n, const int* ids, const double* xyt) override {
        for (int i = 0; i < n; i++) o.add_pose((T)xyt[3 * i], (T)xyt[3 * i + 1], (T)xyt[3 * i + 2], ids[i]);
    }
    void set_landmarks(int n, const int* ids, const double* xy) override {
        for (int i = 0; i < n; i++) o.add_landmark((T)xy[2 * i], (T)xy[2 * i + 1], ids[i]);
    }
    void set_bearings(int n, const int* p, const int* l, const double* z, const double* om) override {
        o.bearings.reserve(o.bearings.size() + n);
        for (int i = 0; i < n; i++) o.bearings.push_back({p[i], l[i], (T)z[i], om ? (T)om[i] : T(1)});
    }
    void set_odometry(int n, const int* s, const int* d, const double* z3, const double* om9) override {
        o.odoms.reserve(o.odoms.size() + n);
        for (int i = 0; i < n; i++) {
            OdomObs<T> e; e.src_id = s[i]; e.dst_id = d[i];
            for (int k = 0; k < 3; k++) e.z[k] = (T)z3[3 * i + k];
            for (int k = 0; k < 9; k++) e.omega[k] = (T)om9[9 * i + k];
            o.odoms.push_back(e);
        }
    }
    void set_fixed(int id) override { o.fixed_pose_id = id; }
    void triangulate() override { o.triangulate_landmarks(); }
    void counts(int out[8]) override {
        out[0] = o.NP(); out[1] = o.NL(); out[2] = (int)o.bearings.size(); out[3] = (int)o.odoms.size();
        out[4] = o.fixed_pose_id; out[5] = o.n_unrecognized; out[6] = o.N; out[7] = o.fixed_stix;
    }
    double get_bound() override { return o.bound; }
    void get_ids(int* pid, int* lid) override {
        if (pid) std::copy(o.pose_stix_to_id.begin(), o.pose_stix_to_id.end(), pid);
        if (lid) std::copy(o.lm_stix_to_id.begin(), o.lm_stix_to_id.end(), lid);
    }
    void get_state(double* P, double* L) override {
        if (P) for (int i = 0; i < o.NP(); i++) {
            P[4 * i] = o.poses[i].tx; P[4 * i + 1] = o.poses[i].ty; P[4 * i + 2] = o.poses[i].r00; P[4 * i + 3] = o.poses[i].r10;
        }
        if (L) for (size_t i = 0; i < o.lms.size(); i++) L[i] = o.lms[i];
    }
    void set_state(const double* P, const double* L) override {
        if (P) for (int i = 0; i < o.NP(); i++) {
            auto& X = o.poses[i];
            X.tx = (T)P[4 * i]; X.ty = (T)P[4 * i + 1];
            X.r00 = (T)P[4 * i + 2]; X.r10 = (T)P[4 * i + 3]; X.r01 = -X.r10; X.r11 = X.r00;
        }
        if (L) for (size_t i = 0; i < o.lms.size(); i++) o.lms[i] = (T)L[i];
    }
    void get_state_xyt(double* P) override {
        for (int i = 0; i < o.NP(); i++) {
            T x, y, th; t2v(o.poses[i], x, y, th);
            P[3 * i] = x; P[3 * i + 1] = y; P[3 * i + 2] = th;
        }
    }
    void get_edges(int* bp, int* bl, double* bz, double* bom, int* os, int* od, double* oz, double* oom) override {
        for (size_t e = 0; e < o.bearings.size(); e++) {
            if (bp) bp[e] = o.bearings[e].pose_id;
            if (bl) bl[e] = o.bearings[e].lm_id;
            if (bz) bz[e] = o.bearings[e].bearing;
            if (bom) bom[e] = o.bearings[e].omega;
        }
        for (size_t e = 0; e < o.odoms.size(); e++) {
            if (os) os[e] = o.odoms[e].src_id;
            if (od) od[e] = o.odoms[e].dst_id;
            if (oz) for (int k = 0; k < 3; k++) oz[3 * e + k] = o.odoms[e].z[k];
            if (oom) for (int k = 0; k < 9; k++) oom[9 * e + k] = o.odoms[e].omega[k];
        }
    }
    int get_single_obs(int* ids, int cap) override {
        int n = (int)o.single_observation_lms.size();
        for (int i = 0; i < n && i < cap; i++) ids[i] = o.single_observation_lms[i];
        return n;
    }
    int solver_init(int fixed_id) override {
        try { o.solver_init(fixed_id); } catch (const std::exception&) { return 1; }
        csc_valid = false;
        return 0;
    }
    void set_params(double kt, double damp) override { o.kernel_threshold = (T)kt; o.damping_factor = (T)damp; }
    void set_irls(int on) override { o.irls = on != 0; }
    void get_edge_stix(int* bp, int* bl, int* os, int* od) override {
        if (bp) std::copy(o.b_pose.begin(), o.b_pose.end(), bp);
        if (bl) std::copy(o.b_lm.begin(), o.b_lm.end(), bl);
        if (os) std::copy(o.o_src.begin(), o.o_src.end(), os);
        if (od) std::copy(o.o_dst.begin(), o.o_dst.end(), od);
    }
    void linearize() override { o.linearize(); csc_valid = false; }
    int n_off() override { return (int)o.off_pairs.size(); }
    void get_blocks(double* hp, double* hl, double* hoff, int* lo, int* hi, double* b) override {
        if (hp) for (size_t i = 0; i < o.Hdiag_p.size(); i++) hp[i] = o.Hdiag_p[i];
        if (hl) for (size_t i = 0; i < o.Hdiag_l.size(); i++) hl[i] = o.Hdiag_l[i];
        if (hoff) for (size_t i = 0; i < o.Hoff.size(); i++) hoff[i] = o.Hoff[i];
        for (size_t k = 0; k < o.off_pairs.size(); k++) {
            if (lo) lo[k] = o.off_pairs[k].first;
            if (hi) hi[k] = o.off_pairs[k].second;
        }
        if (b) for (size_t i = 0; i < o.bvec.size(); i++) b[i] = o.bvec[i];
    }
    void get_edge_terms(double* eb, double* jb, double* eo, double* jo) override {
        if (eb) for (size_t i = 0; i < o.err_b.size(); i++) eb[i] = o.err_b[i];
        if (jb) for (size_t i = 0; i < o.jac_b.size(); i++) jb[i] = o.jac_b[i];
        if (eo) for (size_t i = 0; i < o.err_o.size(); i++) eo[i] = o.err_o[i];
        if (jo) for (size_t i = 0; i < o.jac_o.size(); i++) jo[i] = o.jac_o[i];
    }
    void ensure_csc() { if (!csc_valid) { o.export_csc(csc, bn); csc_valid = true; } }
    int csc_nnz() override { ensure_csc(); return (int)csc.rowidx.size(); }
    void get_csc(int* colptr, int* rowidx, double* val, double* b) override {
        ensure_csc();
        if (colptr) std::copy(csc.colptr.begin(), csc.colptr.end(), colptr);
        if (rowidx) std::copy(csc.rowidx.begin(), csc.rowidx.end(), rowidx);
        if (val) for (size_t i = 0; i < csc.val.size(); i++) val[i] = csc.val[i];
        if (b) for (size_t i = 0; i < bn.size(); i++) b[i] = bn[i];
    }
    void solve(int kind, int max_iters, double rtol) override {
        if (kind == 0) o.solve_dense_ldlt(); else if (kind == 2) o.solve_sparse_ldlt(); else o.solve_schur_pcg(max_iters, rtol);
    }
    int solve_sparse(double deadline_s, double* info8) override {   // info8: 10 doubles
        const bool ok = o.solve_sparse_ldlt(deadline_s);
        if (info8) {
            info8[0] = o.t_order; info8[1] = o.t_analyze; info8[2] = o.t_factor; info8[3] = o.t_trisolve; info8[4] = (double)o.ldlt.nnzL;
            info8[5] = o.ldlt.flops; info8[6] = o.t_export; info8[7] = o.ldlt.status; info8[8] = o.ldlt.flops_done; info8[9] = o.ldlt.rows_done;
        }
        return ok ? 0 : 1;
    }
    double time_linearize_literal(int reps) override {
        auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < reps; r++) o.linearize_literal();
        auto t1 = std::chrono::steady_clock::now();
        return std::chrono::duration<double>(t1 - t0).count() / reps;
    }
    // max |H_literal - H_blocks| over the entries of the exported CSC (with the fixed pose's rows / columns kept in the literal matrix)
    double literal_max_diff() override {
        std::vector<T> bsave = o.bvec;
        o.linearize_literal();
        std::vector<T> blit = o.bvec;
        o.linearize();
        double worst = 0;
        for (size_t i = 0; i < blit.size(); i++) worst = std::max(worst, std::abs((double)blit[i] - (double)o.bvec[i]));
        Csc<T> A; std::vector<T> bnv;
        o.export_csc(A, bnv);
        // index map nofixed -> full
        std::vector<int> full(A.n);
        for (int i = 0, k = 0; i < o.N; i++) if (o.nofixed_index(i) >= 0) full[k++] = i;
        for (int j = 0; j < A.n; j++)
            for (int q = A.colptr[j]; q < A.colptr[j + 1]; q++) {
                const int gi = full[A.rowidx[q]], gj = full[j];
                double v = 0; bool found = false;
                for (int t = o.Hlit.colptr[gj]; t < o.Hlit.colptr[gj + 1]; t++)
                    if (o.Hlit.rowidx[t] == gi) { v = (double)o.Hlit.val[t]; found = true; break; }
                if (!found) return 1e300;
                worst = std::max(worst, std::abs(v - (double)A.val[q]));
            }
        csc_valid = false;
        (void)bsave;
        return worst;
    }
    void set_wrap_branch_tol(double tol) override { o.wrap_branch_tol = tol; }
    void set_wrap_branch(int n, const int* edges, const int* signs) override {
        o.wrap_branch.clear();
        for (int i = 0; i < n; i++) o.wrap_branch[edges[i]] = signs[i] >= 0 ? 1 : -1;
    }
    void get_delta(double* d) override { for (size_t i = 0; i < o.delta.size(); i++) d[i] = o.delta[i]; }
    void set_delta(const double* d) override { o.delta.resize(o.N); for (int i = 0; i < o.N; i++) o.delta[i] = (T)d[i]; }
    void apply_boxplus() override { o.apply_boxplus(); }
    void step(int kind, int max_iters, double rtol) override { o.step(kind, max_iters, rtol); csc_valid = false; }
    void get_stats(double out[8]) override {
        out[0] = o.stats.chi2_bearing; out[1] = o.stats.chi2_odometry; out[2] = o.stats.over_bearing;
        out[3] = o.stats.over_odometry; out[4] = o.stats.delta_inf; out[5] = o.stats.solver_status;
        out[6] = o.stats.pcg_iterations; out[7] = 0;
    }
    double predict_bearing(double x, double y, double th, double lx, double ly) override {
        return Oracle<T>::predict_bearing(v2t<T>((T)x, (T)y, (T)th), (T)lx, (T)ly);
    }
    void predict_odometry(const double* s, const double* d, double* out3) override {
        T p[3];
        Oracle<T>::predict_odometry(v2t<T>((T)s[0], (T)s[1], (T)s[2]), v2t<T>((T)d[0], (T)d[1], (T)d[2]), p);
        for (int k = 0; k < 3; k++) out3[k] = p[k];
    }
    void bearing_jacobians(int e, double* ana5, double* num5) override {
        const auto& ob = o.bearings[e];
        const auto& X = o.poses[o.pose_stix(ob.pose_id)];
        int l = o.lm_stix(ob.lm_id);
        T err, Ja[5], Jn[5];
        Oracle<T>::bearing_error_and_jacobian(X, o.lms[2 * l], o.lms[2 * l + 1], ob.bearing, err, Ja);
        Oracle<T>::bearing_numeric_jacobian(X, o.lms[2 * l], o.lms[2 * l + 1], ob.bearing, Jn);
        for (int k = 0; k < 5; k++) { ana5[k] = Ja[k]; num5[k] = Jn[k]; }
    }
    void odometry_jacobians(int e, double* ana18, double* num18) override {
        const auto& ob = o.odoms[e];
        const auto& S = o.poses[o.pose_stix(ob.src_id)];
        const auto& D = o.poses[o.pose_stix(ob.dst_id)];
        T err[3], Ja[18], Jn[18];
        Oracle<T>::odometry_error_and_jacobian(S, D, ob.z, err, Ja);
        Oracle<T>::odometry_numeric_jacobian(S, D, ob.z, Jn);
        for (int k = 0; k < 18; k++) { ana18[k] = Ja[k]; num18[k] = Jn[k]; }
    }
    double time_linearize(int reps) override {
        auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < reps; r++) o.linearize();
        auto t1 = std::chrono::steady_clock::now();
        return std::chrono::duration<double>(t1 - t0).count() / reps;
    }
};

inline IOracle* H(void* h) { return static_cast<IOracle*>(h); }

}  // namespace

extern "C" {

void* orc_new(int use_double) { return use_double ? (IOracle*)new Impl<double>() : (IOracle*)new Impl<float>(); }
void orc_free(void* h) { delete H(h); }
int orc_load_g2o(void* h, const char* path) {
    try { return H(h)->load_g2o(path); } catch (const std::exception&) { return 2; }
}
void orc_set_poses(void* h, int n, const int* ids, const double* xyt) { H(h)->set_poses(n, ids, xyt); }
void orc_set_landmarks(void* h, int n, const int* ids, const double* xy) { H(h)->set_landmarks(n, ids, xy); }
void orc_set_bearings(void* h, int n, const int* p, const int* l, const double* z, const double* om) { H(h)->set_bearings(n, p, l, z, om); }
void orc_set_odometry(void* h, int n, const int* s, const int* d, const double* z3, const double* om9) { H(h)->set_odometry(n, s, d, z3, om9); }
void orc_set_fixed(void* h, int id) { H(h)->set_fixed(id); }
int orc_triangulate(void* h) {
    try { H(h)->triangulate(); } catch (const std::exception&) { return 1; }
    return 0;
}
void orc_counts(void* h, int* out8) { H(h)->counts(out8); }
double orc_bound(void* h) { return H(h)->get_bound(); }
void orc_get_ids(void* h, int* pose_ids, int* lm_ids) { H(h)->get_ids(pose_ids, lm_ids); }
void orc_get_state(void* h, double* poses_xycs, double* lms_xy) { H(h)->get_state(poses_xycs, lms_xy); }
void orc_set_state(void* h, const double* poses_xycs, const double* lms_xy) { H(h)->set_state(poses_xycs, lms_xy); }
void orc_get_state_xyt(void* h, double* poses_xyt) { H(h)->get_state_xyt(poses_xyt); }
void orc_get_edges(void* h, int* bp, int* bl, double* bz, double* bom, int* os, int* od, double* oz, double* oom) {
    H(h)->get_edges(bp, bl, bz, bom, os, od, oz, oom);
}
int orc_get_single_obs(void* h, int* ids, int cap) { return H(h)->get_single_obs(ids, cap); }
int orc_solver_init(void* h, int fixed_id) { return H(h)->solver_init(fixed_id); }
void orc_set_params(void* h, double kt, double damp) { H(h)->set_params(kt, damp); }
void orc_set_irls(void* h, int on) { H(h)->set_irls(on); }
void orc_get_edge_stix(void* h, int* bp, int* bl, int* os, int* od) { H(h)->get_edge_stix(bp, bl, os, od); }
void orc_linearize(void* h) { H(h)->linearize(); }
int orc_n_off(void* h) { return H(h)->n_off(); }
void orc_get_blocks(void* h, double* hp, double* hl, double* hoff, int* lo, int* hi, double* b) { H(h)->get_blocks(hp, hl, hoff, lo, hi, b); }
void orc_get_edge_terms(void* h, double* eb, double* jb, double* eo, double* jo) { H(h)->get_edge_terms(eb, jb, eo, jo); }
int orc_csc_nnz(void* h) { return H(h)->csc_nnz(); }
void orc_get_csc(void* h, int* colptr, int* rowidx, double* val, double* b) { H(h)->get_csc(colptr, rowidx, val, b); }
void orc_solve(void* h, int kind, int max_iters, double rtol) { H(h)->solve(kind, max_iters, rtol); }
void orc_get_delta(void* h, double* d) { H(h)->get_delta(d); }
void orc_set_delta(void* h, const double* d) { H(h)->set_delta(d); }
void orc_apply_boxplus(void* h) { H(h)->apply_boxplus(); }
void orc_step(void* h, int kind, int max_iters, double rtol) { H(h)->step(kind, max_iters, rtol); }
void orc_get_stats(void* h, double* out8) { H(h)->get_stats(out8); }
double orc_predict_bearing(void* h, double x, double y, double th, double lx, double ly) { return H(h)->predict_bearing(x, y, th, lx, ly); }
void orc_predict_odometry(void* h, const double* s, const double* d, double* out3) { H(h)->predict_odometry(s, d, out3); }
void orc_bearing_jacobians(void* h, int e, double* ana5, double* num5) { H(h)->bearing_jacobians(e, ana5, num5); }
void orc_odometry_jacobians(void* h, int e, double* ana18, double* num18) { H(h)->odometry_jacobians(e, ana18, num18); }
double orc_time_linearize(void* h, int reps) { return H(h)->time_linearize(reps); }
int orc_solve_sparse(void* h, double deadline_s, double* info8) { return H(h)->solve_sparse(deadline_s, info8); }
double orc_time_linearize_literal(void* h, int reps) { return H(h)->time_linearize_literal(reps); }
double orc_literal_max_diff(void* h) { return H(h)->literal_max_diff(); }
void orc_set_wrap_branch(void* h, int n, const int* edges, const int* signs) { H(h)->set_wrap_branch(n, edges, signs); }
void orc_set_wrap_branch_tol(void* h, double tol) { H(h)->set_wrap_branch_tol(tol); }
double orc_smallest_angle(int use_double, double a) { return use_double ? smallest_angle<double>(a) : (double)smallest_angle<float>((float)a); }
double orc_normalized_angle(int use_double, double a) { return use_double ? normalized_angle<double>(a) : (double)normalized_angle<float>((float)a); }
void orc_colpiv_solve(int use_double, int M, const double* A, const double* b, double* out2) {
    if (use_double) {
        std::vector<double> a(A, A + 2 * M), r(b, b + M);
        colpiv_householder_solve_Mx2<double>(a, r, M, out2);
    } else {
        std::vector<float> a(2 * M), r(M);
        for (int i = 0; i < 2 * M; i++) a[i] = (float)A[i];
        for (int i = 0; i < M; i++) r[i] = (float)b[i];
        float o[2];
        colpiv_householder_solve_Mx2<float>(a, r, M, o);
        out2[0] = o[0]; out2[1] = o[1];
    }
}

}  // extern "C"
