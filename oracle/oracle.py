"""ctypes front end of the CPU oracle (oracle/bos_oracle.hpp).

CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg; never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "libbos_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("bos_oracle_capi.cpp", "bos_oracle.hpp", "bos_sparse_ldlt.hpp")]
    stale = not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libbos_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.orc_new.restype = C.c_void_p
        L.orc_new.argtypes = [C.c_int]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_load_g2o.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_bound.restype = C.c_double
        L.orc_bound.argtypes = [C.c_void_p]
        L.orc_predict_bearing.restype = C.c_double
        L.orc_predict_bearing.argtypes = [C.c_void_p] + [C.c_double] * 5
        L.orc_time_linearize.restype = C.c_double
        L.orc_time_linearize.argtypes = [C.c_void_p, C.c_int]
        L.orc_smallest_angle.restype = C.c_double
        L.orc_smallest_angle.argtypes = [C.c_int, C.c_double]
        L.orc_normalized_angle.restype = C.c_double
        L.orc_normalized_angle.argtypes = [C.c_int, C.c_double]
        L.orc_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.orc_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.orc_set_params.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_set_irls.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_wrap_branch_tol.argtypes = [C.c_void_p, C.c_double]
        L.orc_solve_sparse.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        L.orc_time_linearize_literal.restype = C.c_double
        L.orc_time_linearize_literal.argtypes = [C.c_void_p, C.c_int]
        L.orc_literal_max_diff.restype = C.c_double
        L.orc_literal_max_diff.argtypes = [C.c_void_p]
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One problem + solver instance; dtype 'f64' (parity target) or 'f32' (reference-faithful)."""

    def __init__(self, dtype="f64"):
        self.L = lib()
        self.dbl = 1 if dtype == "f64" else 0
        self.h = C.c_void_p(self.L.orc_new(self.dbl))

    def __del__(self):
        try:
            self.L.orc_free(self.h)
        except Exception:
            pass

    # ---- problem -------------------------------------------------------------------------
    def load_g2o(self, path):
        rc = self.L.orc_load_g2o(self.h, path.encode())
        if rc != 0:
            raise IOError("oracle: cannot parse %s (rc=%d)" % (path, rc))

    def set_problem(self, pose_ids, poses_xyt, b_pose_id, b_lm_id, b_z, o_src_id, o_dst_id, o_z, o_omega,
                    lm_ids=None, lms_xy=None, b_omega=None, fixed_id=-1):
        c = lambda a, t: np.ascontiguousarray(a, dtype=t)
        pose_ids = c(pose_ids, np.int32); poses_xyt = c(poses_xyt, np.float64)
        self.L.orc_set_poses(self.h, len(pose_ids), _p(pose_ids), _p(poses_xyt))
        if lm_ids is not None and len(lm_ids):
            lm_ids = c(lm_ids, np.int32); lms_xy = c(lms_xy, np.float64)
            self.L.orc_set_landmarks(self.h, len(lm_ids), _p(lm_ids), _p(lms_xy))
        b_pose_id = c(b_pose_id, np.int32); b_lm_id = c(b_lm_id, np.int32); b_z = c(b_z, np.float64)
        bo = None if b_omega is None else c(b_omega, np.float64)
        self.L.orc_set_bearings(self.h, len(b_z), _p(b_pose_id), _p(b_lm_id), _p(b_z), _p(bo))
        o_src_id = c(o_src_id, np.int32); o_dst_id = c(o_dst_id, np.int32)
        o_z = c(o_z, np.float64); o_omega = c(o_omega, np.float64)
        self.L.orc_set_odometry(self.h, len(o_src_id), _p(o_src_id), _p(o_dst_id), _p(o_z), _p(o_omega))
        self.L.orc_set_fixed(self.h, int(fixed_id))

    def triangulate(self):
        if self.L.orc_triangulate(self.h) != 0:
            raise KeyError("oracle: triangulation hit an unknown pose id")

    def counts(self):
        out = np.zeros(8, np.int32)
        self.L.orc_counts(self.h, _p(out))
        return dict(NP=int(out[0]), NL=int(out[1]), Eb=int(out[2]), Eo=int(out[3]), fixed_pose_id=int(out[4]),
                    unrecognized=int(out[5]), N=int(out[6]), fixed_stix=int(out[7]))

    def bound(self):
        return float(self.L.orc_bound(self.h))

    def ids(self):
        c = self.counts()
        p = np.zeros(c["NP"], np.int32); l = np.zeros(c["NL"], np.int32)
        self.L.orc_get_ids(self.h, _p(p), _p(l))
        return p, l

    def state(self):
        c = self.counts()
        P = np.zeros((c["NP"], 4)); Lm = np.zeros((c["NL"], 2))
        self.L.orc_get_state(self.h, _p(P), _p(Lm))
        return P, Lm

    def set_state(self, P, Lm):
        P = np.ascontiguousarray(P, np.float64); Lm = np.ascontiguousarray(Lm, np.float64)
        self.L.orc_set_state(self.h, _p(P), _p(Lm))

    def state_xyt(self):
        c = self.counts()
        P = np.zeros((c["NP"], 3))
        self.L.orc_get_state_xyt(self.h, _p(P))
        return P

    def edges(self):
        c = self.counts()
        bp = np.zeros(c["Eb"], np.int32); bl = np.zeros(c["Eb"], np.int32)
        bz = np.zeros(c["Eb"]); bom = np.zeros(c["Eb"])
        os_ = np.zeros(c["Eo"], np.int32); od = np.zeros(c["Eo"], np.int32)
        oz = np.zeros((c["Eo"], 3)); oom = np.zeros((c["Eo"], 9))
        self.L.orc_get_edges(self.h, _p(bp), _p(bl), _p(bz), _p(bom), _p(os_), _p(od), _p(oz), _p(oom))
        return dict(b_pose_id=bp, b_lm_id=bl, b_z=bz, b_omega=bom, o_src_id=os_, o_dst_id=od, o_z=oz, o_omega=oom)

    def single_observation_landmarks(self):
        ids = np.zeros(4096, np.int32)
        n = self.L.orc_get_single_obs(self.h, _p(ids), len(ids))
        return ids[:n].tolist()

    # ---- solver --------------------------------------------------------------------------
    def solver_init(self, fixed_id):
        if self.L.orc_solver_init(self.h, int(fixed_id)) != 0:
            raise KeyError("oracle: unknown id while resolving edges / fixed pose")

    def set_params(self, kernel_threshold=1.0, damping=0.01):
        self.L.orc_set_params(self.h, float(kernel_threshold), float(damping))

    def set_irls(self, on=True):
        """opt-in IRLS flavour of the threshold kernel (the weight scales Omega; not in the reference)."""
        self.L.orc_set_irls(self.h, 1 if on else 0)

    def edge_stix(self):
        c = self.counts()
        bp = np.zeros(c["Eb"], np.int32); bl = np.zeros(c["Eb"], np.int32)
        os_ = np.zeros(c["Eo"], np.int32); od = np.zeros(c["Eo"], np.int32)
        self.L.orc_get_edge_stix(self.h, _p(bp), _p(bl), _p(os_), _p(od))
        return bp, bl, os_, od

    def linearize(self):
        self.L.orc_linearize(self.h)

    def blocks(self):
        c = self.counts()
        n_off = self.L.orc_n_off(self.h)
        hp = np.zeros((c["NP"], 9)); hl = np.zeros((c["NL"], 4)); hoff = np.zeros((n_off, 9))
        lo = np.zeros(n_off, np.int32); hi = np.zeros(n_off, np.int32); b = np.zeros(c["N"])
        self.L.orc_get_blocks(self.h, _p(hp), _p(hl), _p(hoff), _p(lo), _p(hi), _p(b))
        return dict(Hpp=hp, Hll=hl, Hoff=hoff, off_lo=lo, off_hi=hi, b=b)

    def edge_terms(self):
        c = self.counts()
        eb = np.zeros(c["Eb"]); jb = np.zeros((c["Eb"], 5)); eo = np.zeros((c["Eo"], 3)); jo = np.zeros((c["Eo"], 18))
        self.L.orc_get_edge_terms(self.h, _p(eb), _p(jb), _p(eo), _p(jo))
        return eb, jb, eo, jo

    def csc(self):
        c = self.counts()
        n = c["N"] - 3
        nnz = self.L.orc_csc_nnz(self.h)
        colptr = np.zeros(n + 1, np.int32); rowidx = np.zeros(nnz, np.int32); val = np.zeros(nnz); b = np.zeros(n)
        self.L.orc_get_csc(self.h, _p(colptr), _p(rowidx), _p(val), _p(b))
        return colptr, rowidx, val, b

    def solve(self, kind=0, max_iters=2000, rtol=1e-12):
        self.L.orc_solve(self.h, kind, max_iters, rtol)

    def solve_sparse(self, deadline_s=0.0):
        """The reference's own solver restated (SimplicialLDLT: minimum-degree ordering + symbolic phase once, up-looking LDL^T per
        call).  Returns a dict of timings / fill; 'finished' is False when deadline_s (seconds, 0 = none) cut the factorisation."""
        info = np.zeros(10)
        rc = self.L.orc_solve_sparse(self.h, float(deadline_s), _p(info))
        return dict(finished=(rc == 0), t_order=float(info[0]), t_analyze=float(info[1]), t_factor=float(info[2]), t_trisolve=float(info[3]),
                    nnzL=int(info[4]), flops=float(info[5]), t_export=float(info[6]), status=int(info[7]), flops_done=float(info[8]),
                    rows_done=int(info[9]))

    def time_linearize_literal(self, reps=1):
        """Seconds per H, b build with the reference's literal per-edge sparse merge (slam/solver.cpp:44,60): O(N + nnz H) per edge."""
        return float(self.L.orc_time_linearize_literal(self.h, reps))

    def literal_max_diff(self):
        return float(self.L.orc_literal_max_diff(self.h))

    def set_wrap_branch(self, edges, signs, tol=1e-9):
        """Test hook: put the bearing edges `edges` (residual within `tol` of +-pi) on the branch `signs` (+1 / -1)."""
        e = np.ascontiguousarray(edges, np.int32); s = np.ascontiguousarray(signs, np.int32)
        self.L.orc_set_wrap_branch_tol(self.h, float(tol))
        self.L.orc_set_wrap_branch(self.h, len(e), _p(e), _p(s))

    def delta(self):
        d = np.zeros(self.counts()["N"])
        self.L.orc_get_delta(self.h, _p(d))
        return d

    def set_delta(self, d):
        d = np.ascontiguousarray(d, np.float64)
        self.L.orc_set_delta(self.h, _p(d))

    def apply_boxplus(self):
        self.L.orc_apply_boxplus(self.h)

    def step(self, kind=0, max_iters=2000, rtol=1e-12):
        self.L.orc_step(self.h, kind, max_iters, rtol)

    def step_lm(self, damping, kernel_threshold=1.0, kind=0):
        """Restatement of the opt-in Levenberg-Marquardt iteration of include/bos_b200.h (bos_step_lm; an extension: the reference has
        a FIXED damping and never rejects a step, slam/solver.cpp:64-69): take a GN step with `damping`, relinearize, keep the step
        and divide the damping by 3 if the total chi2 (pre-kernel error_omeganorm sums) decreased, else restore the state and
        multiply the damping by 10 (clamped to [1e-9, 1e9]).  Returns (chi2_before, chi2_after, accepted, next_damping)."""
        P, L = self.state()
        self.set_params(kernel_threshold, damping)
        self.linearize()
        s = self.stats()
        before = s["chi2_bearing"] + s["chi2_odometry"]
        self.solve(kind)
        self.apply_boxplus()
        self.linearize()
        s = self.stats()
        after = s["chi2_bearing"] + s["chi2_odometry"]
        ok = after < before
        if ok:
            damping = max(damping / 3.0, 1e-9)
        else:
            self.set_state(P, L)
            damping = min(damping * 10.0, 1e9)
        return before, after, ok, damping

    def stats(self):
        s = np.zeros(8)
        self.L.orc_get_stats(self.h, _p(s))
        return dict(chi2_bearing=s[0], chi2_odometry=s[1], over_bearing=int(s[2]), over_odometry=int(s[3]),
                    delta_inf=s[4], solver_status=int(s[5]), pcg_iterations=int(s[6]))

    def predict_bearing(self, x, y, th, lx, ly):
        return float(self.L.orc_predict_bearing(self.h, x, y, th, lx, ly))

    def predict_odometry(self, s_xyt, d_xyt):
        s = np.ascontiguousarray(s_xyt, np.float64); d = np.ascontiguousarray(d_xyt, np.float64); o = np.zeros(3)
        self.L.orc_predict_odometry(self.h, _p(s), _p(d), _p(o))
        return o

    def bearing_jacobians(self, e):
        a = np.zeros(5); n = np.zeros(5)
        self.L.orc_bearing_jacobians(self.h, int(e), _p(a), _p(n))
        return a, n

    def odometry_jacobians(self, e):
        a = np.zeros(18); n = np.zeros(18)
        self.L.orc_odometry_jacobians(self.h, int(e), _p(a), _p(n))
        return a, n

    def time_linearize(self, reps=1):
        return float(self.L.orc_time_linearize(self.h, reps))


def smallest_angle(a, dtype="f64"):
    return float(lib().orc_smallest_angle(1 if dtype == "f64" else 0, float(a)))


def normalized_angle(a, dtype="f64"):
    return float(lib().orc_normalized_angle(1 if dtype == "f64" else 0, float(a)))


def colpiv_solve(A, b, dtype="f64"):
    A = np.ascontiguousarray(A, np.float64); b = np.ascontiguousarray(b, np.float64); out = np.zeros(2)
    lib().orc_colpiv_solve(1 if dtype == "f64" else 0, len(b), _p(A), _p(b), _p(out))
    return out
